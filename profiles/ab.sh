# A/B of two builds of the library in ONE gpurun call (box-to-box variance is +-3 %): lib/libdbs_old.so vs lib/libdbs_b200.so
# usage: [LIBS="old v2 b200"] bash profiles/ab.sh [grep pattern for per-layer lines]   (lib/libdbs_<name>.so each)
PAT=${1:-"conv TFLOP"}
for rep in 1 2; do
  for lib in ${LIBS:-old b200}; do
    echo "== $lib"
    DBS_B200_LIB=$PWD/diffusion-boltzmann-sampler_b200/lib/libdbs_$lib.so python profiles/r2_c4_layers.py 64 1024 2>&1 | grep "$PAT\|^{\|B= *2048"
  done
done
