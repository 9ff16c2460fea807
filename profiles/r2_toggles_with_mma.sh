export DBS_B200_LIB=$PWD/diffusion-boltzmann-sampler_b200/lib/libdbs_x.so
for shape in "0 3 512 64 64 64 0 64" "0 3 512 64 64 32 0 64"; do
  for dbg in 0 16 4 64 20 80 84 116 32 2 1; do
    DBS_STRIP_DBG=$dbg python profiles/conv_micro.py $shape 10 2>&1 | tail -1
  done
done
