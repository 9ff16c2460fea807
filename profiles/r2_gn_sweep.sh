# round 2: GroupNorm-apply launch shape at the config-4 sizes (DBS_GN_THREADS x DBS_GN_UN x DBS_GN_CAP), total gn_apply ms per forward
for thr in 256 512; do for un in 2 4; do for cap in 4 8 16 32; do
  echo -n "threads=$thr un=$un cap=$cap: "
  DBS_GN_THREADS=$thr DBS_GN_UN=$un DBS_GN_CAP=$cap python profiles/r2_c4_layers.py 64 1024 2>&1 | grep "^{" | head -1
done; done; done
