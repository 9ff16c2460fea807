# round 2: what bounds the strip epilogue at the config-4 shapes (64x64, B = 256 = the pixel count of 32x32 B = 1024 x4... )
# DBS_STRIP_DBG bits: 1 no MMA, 2 no epilogue body, 4 no stores, 8 no strip loads, 16 no neighbour shuffles/FMAs,
# 32 no statistics, 64 no neighbour-slot TMEM loads
for shape in "0 3 512 64 64 64 0 64" "0 3 512 64 64 32 0 32" "0 3 512 64 64 32 0 64"; do
  for dbg in 0 1 2 3 5 17 33 65 81 85 117 9 11; do
    DBS_STRIP_DBG=$dbg python profiles/conv_micro.py $shape 10 2>&1 | tail -1
  done
done
