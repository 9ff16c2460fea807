# role-toggle sweep of the strip kernel on the full-resolution layer shapes (see profiles/README.md)
# usage: DBGS="0 1 2" bash profiles/micro_sweep.sh
for shape in "0 3 1024 32 32 32 0 32" "2 4 1024 16 16 64 0 64" "0 3 1024 32 32 32 0 64" "0 3 1024 32 32 64 0 64" "0 3 1024 32 32 64 64 32"; do
  for dbg in ${DBGS:-0 1 2 4 16 32 64 3}; do
    DBS_STRIP_DBG=$dbg python profiles/conv_micro.py $shape 2>&1 | tail -1
  done
done
