mkdir -p gpurun_out
for chains in 1024 2368 4096; do
  echo "== default chains=$chains"
  timeout 200 python tools/probe_water.py 32 $chains 2000 2>&1 | grep -E "step [12]|rror" | cut -c1-110
  for lib in build_variants/*.so; do
    echo "== $lib chains=$chains"
    JELLYFYSH_B200_LIBRARY=$PWD/$lib timeout 200 python tools/probe_water.py 32 $chains 2000 2>&1 | grep -E "step [12]|rror" | cut -c1-110
  done
done
