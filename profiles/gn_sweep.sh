for cfg in "256 2 8" "256 4 8" "256 1 8" "512 2 8" "256 2 16" "256 2 4" "128 2 8" "256 4 16"; do
  set -- $cfg
  echo -n "threads=$1 un=$2 cap=$3: "
  DBS_GN_THREADS=$1 DBS_GN_UN=$2 DBS_GN_CAP=$3 python profiles/layer_times.py 1024 32 2>&1 | grep "^{" 
done
