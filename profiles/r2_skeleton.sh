# round 2: cost of tcgen05.commit in the per-tile skeleton (bits 1024 / 2048: plain mbarrier arrive instead, valid with bit 1 = no MMA)
for shape in "0 3 512 64 64 64 0 64"; do
  for dbg in 11 1035 2059 3083 9 3081 1 3073; do
    DBS_GN_PROLOGUE=0 DBS_STRIP_DBG=$dbg timeout 60 python profiles/conv_micro.py $shape 10 2>&1 | tail -1
  done
done
