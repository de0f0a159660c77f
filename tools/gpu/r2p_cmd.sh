mkdir -p gpurun_out
timeout 1700 python -m pytest tests -q -m gpu -s > gpurun_out/r2p_pytest_gpu.log 2>&1; echo "tests rc=$?"
grep -E "KS dist|passed|failed|^E  |^FAILED|horizon" gpurun_out/r2p_pytest_gpu.log | cut -c1-260 | tail -40
