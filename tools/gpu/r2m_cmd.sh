mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_spec.py -q -m gpu -x > gpurun_out/r2m_pytest_spec.log 2>&1; echo "spec tests rc=$?"; tail -15 gpurun_out/r2m_pytest_spec.log | cut -c1-250
for v in "ECMC_CHAIN_BLOCKS=0" "ECMC_CHAIN_BLOCKS=1" "ECMC_CHAIN_BLOCKS=1 ECMC_SPEC_PRUNE=0"; do
  echo "== C5 $v"; env $v timeout 120 python tools/probe.py 1 65536 48 50000 2>&1 | grep -E "step [2-4]|rror" | cut -c1-150
done
for v in "ECMC_CHAIN_BLOCKS=0" "ECMC_CHAIN_BLOCKS=1"; do
  echo "== 148 chains of C2 $v"; env $v timeout 120 python tools/probe.py 148 1024 12 4096 2>&1 | grep -E "step [2-4]|rror" | cut -c1-150
done
