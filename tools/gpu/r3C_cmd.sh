mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_chains.py tests/test_abi.py -x -q -m gpu -k "leaf_cell or water or abi or composite" > gpurun_out/r3C_pytest.log 2>&1; echo "pytest rc=$?"; tail -30 gpurun_out/r3C_pytest.log
