mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu --durations=25 > gpurun_out/r2W_pytest_gpu.txt 2>&1; echo "pytest rc=$?"; tail -40 gpurun_out/r2W_pytest_gpu.txt
