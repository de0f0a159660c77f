# first look at the batched kernel: its own tests, the chain parity tests, speeds of the variants, then the whole suite
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_spec.py -x -q > gpurun_out/r2b_spec.log 2>&1; echo "spec rc=$?"; tail -15 gpurun_out/r2b_spec.log
timeout 300 python -m pytest tests/test_gpu_chains.py tests/test_gpu_full_size_parity.py -x -q --durations=5 > gpurun_out/r2b_chains.log 2>&1; echo "chains rc=$?"; tail -8 gpurun_out/r2b_chains.log
for v in "ECMC_SPEC=0" "ECMC_SPEC_PRUNE=0 ECMC_SPEC_LANES=4" "ECMC_SPEC_PRUNE=0 ECMC_SPEC_LANES=8" "ECMC_SPEC_PRUNE=1 ECMC_SPEC_LANES=4" "ECMC_SPEC_PRUNE=1 ECMC_SPEC_LANES=8"; do
  echo "== $v"; env $v timeout 120 python tools/probe.py 4096 1024 12 1024 2>&1 | grep -E "step [1-5]|rror" | cut -c1-150
done > gpurun_out/r2b_probe.txt 2>&1
cat gpurun_out/r2b_probe.txt
for v in "ECMC_SPEC=0" "ECMC_SPEC_PRUNE=0" "ECMC_SPEC_PRUNE=1" "ECMC_SPEC_PRUNE=1 ECMC_SPEC_LANES=8"; do
  echo "== C5 $v"; env $v timeout 120 python tools/probe.py 1 65536 48 50000 2>&1 | grep -E "step [1-5]|rror" | cut -c1-150
done > gpurun_out/r2b_probe_c5.txt 2>&1
cat gpurun_out/r2b_probe_c5.txt
timeout 400 python -m pytest tests -m gpu -x -q --durations=8 > gpurun_out/r2b_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -12 gpurun_out/r2b_pytest_gpu.log
