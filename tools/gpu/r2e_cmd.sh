mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu --durations=8 > gpurun_out/r2e_pytest_gpu.log 2>&1; echo "tests rc=$?"; tail -14 gpurun_out/r2e_pytest_gpu.log
timeout 600 python bench.py > gpurun_out/r2e_bench.json 2> gpurun_out/r2e_bench.err; echo "bench rc=$?"; cat gpurun_out/r2e_bench.json
python tools/probe_e2e_slices.py > gpurun_out/r2e_e2e_slices.txt 2>&1; tail -12 gpurun_out/r2e_e2e_slices.txt
