# A/B of two builds of the library in ONE gpurun call (box-to-box variance is +-3 %): lib/libdbs_old.so vs lib/libdbs_b200.so
# usage: bash profiles/ab.sh [grep pattern for per-layer lines]
PAT=${1:-"conv TFLOP"}
for rep in 1 2; do
  for lib in old b200; do
    echo "== $lib"
    DBS_B200_LIB=$PWD/diffusion-boltzmann-sampler_b200/lib/libdbs_$lib.so python profiles/r2_c4_layers.py 64 1024 2>&1 | grep "$PAT\|^{\|B= *2048"
  done
done
