mkdir -p gpurun_out
echo "== window sweep (probe: 4096 chains x 1024 particles, 1024 events per step)"
for lib in build_variants/w*.so; do
  echo "-- $lib"
  JELLYFYSH_B200_LIBRARY=$PWD/$lib timeout 200 python tools/probe.py 4096 1024 12 1024 2>&1 | grep -E "step [2-5]|rror" | cut -c1-150
done
