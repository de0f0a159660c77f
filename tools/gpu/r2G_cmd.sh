mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_chains.py -q -m gpu -x -k "sequential" > gpurun_out/r2G_pytest.log 2>&1; echo "tests rc=$?"; tail -2 gpurun_out/r2G_pytest.log | cut -c1-200
timeout 500 python bench.py --workload c1n > gpurun_out/r2G_bench_c1n.json 2> gpurun_out/r2G_bench_c1n.err; echo "bench c1n rc=$?"; tail -3 gpurun_out/r2G_bench_c1n.err | cut -c1-300
timeout 500 python bench.py --workload c3 --particles 512 > gpurun_out/r2G_bench_c3_512.json 2> gpurun_out/r2G_bench_c3_512.err; echo "bench c3 512 rc=$?"; tail -3 gpurun_out/r2G_bench_c3_512.err | cut -c1-300
python - <<'PY'
import json
for w in ("c1n","c3_512"):
    try:
        d=json.loads(open(f'gpurun_out/r2G_bench_{w}.json').read().strip().splitlines()[-1]); e=d['e2e']
        print(w, 'value %.4g (%.2f ms) e2e %.4g roofline %.4f targets %.1f cpu %.4g kernel %s' % (d['value'], d['ms_per_step'], e['value'], d['roofline']['frac'], d['roofline']['pair_targets_per_event'], d.get('cpu_baseline',{}).get('value',0), d['roofline']['kernel']))
    except Exception as error:
        print(w, 'failed', error)
PY
