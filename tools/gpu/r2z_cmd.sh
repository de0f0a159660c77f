mkdir -p gpurun_out
echo "== window / rebuild fraction sweep (probe: 4096 chains x 1024 particles, 1024 events per step)"
for lib in build_variants/w*.so; do
  echo "-- $lib"
  JELLYFYSH_B200_LIBRARY=$PWD/$lib timeout 200 python tools/probe.py 4096 1024 12 1024 2>&1 | grep -E "step [2-5]|rror" | cut -c1-150
done
echo "== C5 single chain with the default library (live entries in lj_chain_kernel)"
timeout 120 python tools/probe.py 1 65536 48 50000 2>&1 | grep -E "step [2-4]|rror" | cut -c1-150
timeout 600 python -m pytest tests/test_gpu_spec.py tests/test_gpu_full_size_parity.py -q -m gpu -x > gpurun_out/r2z_pytest.log 2>&1; echo "tests rc=$?"; tail -3 gpurun_out/r2z_pytest.log | cut -c1-200
