mkdir -p gpurun_out
timeout 1200 python -m pytest tests -x -q -m gpu > gpurun_out/r2T_pytest_gpu.txt 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r2T_pytest_gpu.txt
timeout 600 python bench.py --workload c4 --no-cpu-baseline > gpurun_out/r2T_bench_c4.json 2> gpurun_out/r2T_bench_c4.err; echo "c4 rc=$?"
timeout 600 python bench.py --workload c4 --chains 1024 --no-cpu-baseline > gpurun_out/r2T_bench_c4_1024.json 2> gpurun_out/r2T_bench_c4_1024.err; echo "c4/1024 rc=$?"
timeout 600 python bench.py --no-cpu-baseline --no-single-chain > gpurun_out/r2T_bench_c2_1024.json 2> gpurun_out/r2T_bench_c2_1024.err; echo "c2 rc=$?"
timeout 600 python bench.py --events 4096 --no-cpu-baseline --no-single-chain > gpurun_out/r2T_bench_c2_4096.json 2> gpurun_out/r2T_bench_c2_4096.err; echo "c2/4096 rc=$?"
python - <<'PY'
import json
for name in ("c4", "c4_1024", "c2_1024", "c2_4096"):
    try:
        d = json.loads(open("gpurun_out/r2T_bench_%s.json" % name).read().strip().splitlines()[-1])
        print(name, "value %.3e ms/step %.3f e2e %.3e" % (d["value"], d["ms_per_step"], d["e2e"]["value"]), d["roofline"]["kernel"], d["config"].get("event_window_per_chain"))
    except Exception as e:
        print(name, "failed", e)
PY
