mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_chains.py tests/test_abi.py -x -q -m gpu -k "root_unit_active or no_cells_composite or abi" > gpurun_out/r2P_pytest.log 2>&1; echo "pytest rc=$?"
tail -30 gpurun_out/r2P_pytest.log
timeout 600 python -m pytest tests/test_gpu_mediator.py -x -q -m gpu -k "shipped_dipole_config" > gpurun_out/r2P_pytest_med.log 2>&1; echo "pytest mediator rc=$?"
tail -15 gpurun_out/r2P_pytest_med.log
