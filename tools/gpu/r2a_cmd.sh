mkdir -p gpurun_out
timeout 500 python -m pytest tests -m gpu -x -q --durations=8 > gpurun_out/r2a_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -12 gpurun_out/r2a_pytest_gpu.log
timeout 200 python bench.py --no-cpu-baseline > gpurun_out/r2a_bench.json 2> gpurun_out/r2a_bench.err; echo "bench rc=$?"; cut -c1-300 gpurun_out/r2a_bench.json
timeout 120 ncu --set full --clock-control none --import-source on -k regex:event_kernel -s 2 -c 1 -f -o gpurun_out/r2a_event_kernel python tools/probe.py 4096 1024 12 1024 > gpurun_out/r2a_ncu_full.log 2>&1; echo "ncu rc=$?"
