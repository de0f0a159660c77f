mkdir -p gpurun_out
for w in 12 14 16 18 20; do
  lib=build_variants/w${w}b1.so
  for chains in $((148 * w)) 4096; do
    echo "== $lib chains=$chains"
    JELLYFYSH_B200_LIBRARY=$PWD/$lib timeout 200 python tools/probe_water.py 32 $chains 2000 2>&1 | grep -E "step [2]|rror" | cut -c1-110
  done
done
