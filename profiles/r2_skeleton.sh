# round 2: per-tile skeleton of the strip kernel after pinning the role invariants (needs lib/libdbs_exp.so = -DDBS_STRIP_EXPERIMENTS build)
export DBS_B200_LIB=$PWD/diffusion-boltzmann-sampler_b200/lib/libdbs_exp.so
for shape in "0 3 512 64 64 64 0 64" "0 3 512 64 64 32 0 32"; do
  for dbg in 0 1 2 3 9 11 139; do
    DBS_STRIP_DBG=$dbg timeout 60 python profiles/conv_micro.py $shape 10 2>&1 | tail -1
  done
done
