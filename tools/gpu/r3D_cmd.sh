mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_mediator.py -x -q -m gpu -k "shipped_water_config and cell_bounded" > gpurun_out/r3D_pytest.log 2>&1; echo "pytest rc=$?"; tail -30 gpurun_out/r3D_pytest.log | cut -c1-400
