mkdir -p gpurun_out
echo "== C2 kernel variants (probe: 4096 chains x 1024 particles, 1024 events per step)"
for lib in build_variants/base.so build_variants/live.so build_variants/base.so build_variants/live.so; do
  echo "-- $lib"
  JELLYFYSH_B200_LIBRARY=$PWD/$lib timeout 200 python tools/probe.py 4096 1024 12 1024 2>&1 | grep -E "step [1-5]|rror" | cut -c1-150
done
echo "== spec parity tests with the live variant"
JELLYFYSH_B200_LIBRARY=$PWD/build_variants/live.so timeout 600 python -m pytest tests/test_gpu_spec.py tests/test_gpu_full_size_parity.py -q -m gpu -x > gpurun_out/r2x_pytest.log 2>&1; echo "tests rc=$?"; tail -3 gpurun_out/r2x_pytest.log | cut -c1-200
