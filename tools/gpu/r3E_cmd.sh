mkdir -p gpurun_out
timeout 1200 python -m pytest tests -q -m gpu --durations=6 > gpurun_out/r3E_pytest_gpu.txt 2>&1; echo "pytest rc=$?"; tail -12 gpurun_out/r3E_pytest_gpu.txt | cut -c1-200
timeout 120 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -3
