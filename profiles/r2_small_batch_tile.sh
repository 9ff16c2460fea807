# tile-kernel layers of the config-4 network at small per-rank batches: mid conv 8x8 256 -> 256 (fused GN), pair mode on / off
for B in 64 128 256 512 2048; do
  for pm in 1 0; do
    echo -n "B=$B DBS_CONV_PAIR=$pm  "; DBS_CONV_PAIR=$pm python profiles/conv_micro.py 0 3 $B 8 8 256 0 256 20 2>&1 | tail -1
  done
done
