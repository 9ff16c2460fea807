# round 2: cost of tcgen05.commit in the per-tile skeleton (bits 1024 / 2048: plain mbarrier arrive instead, valid with bit 1 = no MMA),
# and the epilogue-side stage release (DBS_STRIP_EPI_RELEASE=1: one commit per tile instead of two)
for shape in "0 3 512 64 64 64 0 64"; do
  for dbg in 11 1035 2059 3083; do
    DBS_STRIP_DBG=$dbg timeout 60 python profiles/conv_micro.py $shape 10 2>&1 | tail -1
  done
  for er in 0 1; do for dbg in 0 11 2 1; do
    echo -n "epi_release=$er "; DBS_STRIP_EPI_RELEASE=$er DBS_STRIP_DBG=$dbg timeout 60 python profiles/conv_micro.py $shape 10 2>&1 | tail -1
  done; done
done
for shape in "0 3 512 64 64 32 0 32" "0 3 512 64 64 32 0 64" "0 3 1024 32 32 64 0 64"; do
  for er in 0 1; do echo -n "epi_release=$er "; DBS_STRIP_EPI_RELEASE=$er timeout 60 python profiles/conv_micro.py $shape 10 2>&1 | tail -1; done
done
