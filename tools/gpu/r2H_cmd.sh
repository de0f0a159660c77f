mkdir -p gpurun_out
for i in 0 1 2 3 4 5 6 7; do
  CUDA_VISIBLE_DEVICES=$i timeout 200 python tools/probe_timeline.py 1 8 400 > gpurun_out/r2H_timeline_rank$i.txt 2>&1 &
done
wait
head -8 gpurun_out/r2H_timeline_rank0.txt | cut -c1-150
for i in 1 7; do sed -n 3,5p gpurun_out/r2H_timeline_rank$i.txt | cut -c1-150; done
timeout 500 python bench.py --workload c1n > gpurun_out/r2H_bench_c1n.json 2> gpurun_out/r2H_bench_c1n.err; echo "bench c1n rc=$?"; tail -2 gpurun_out/r2H_bench_c1n.err | cut -c1-300
python - <<'PY'
import json
try:
    d=json.loads(open('gpurun_out/r2H_bench_c1n.json').read().strip().splitlines()[-1]); e=d['e2e']
    print('c1n value %.4g (%.2f ms) e2e %.4g roofline %.4f targets %.1f cpu %.4g kernel %s' % (d['value'], d['ms_per_step'], e['value'], d['roofline']['frac'], d['roofline']['pair_targets_per_event'], d.get('cpu_baseline',{}).get('value',0), d['roofline']['kernel']))
except Exception as error:
    print('c1n failed', error)
PY
