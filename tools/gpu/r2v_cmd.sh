mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_chains.py -q -m gpu -x -k "continued or sparse or submitted or lennard_jones_2d" > gpurun_out/r2v_pytest.log 2>&1; echo "tests rc=$?"
grep -E "passed|failed|^E  |^FAILED" gpurun_out/r2v_pytest.log | cut -c1-260 | tail -12
timeout 900 python bench.py 2> gpurun_out/r2v_bench.err > gpurun_out/r2v_bench.json; echo "bench rc=$?"; tail -3 gpurun_out/r2v_bench.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2v_bench.json').read().strip().splitlines()[-1])
e=d['e2e']
print('value %.4g  ms/step %.3f targets %.2f' % (d['value'], d['ms_per_step'], d['roofline']['pair_targets_per_event']))
print('e2e sparse %.4g (%.2f ms/step, %d steps) d2h %d B  host GB/s %.1f targets %.2f' % (e['value'], e['ms_per_step'], e['steps'], e['d2h_bytes_per_step'], e['host_gb_per_s_per_rank'], e['pair_targets_per_event']))
f=e.get('full_copy',{}); print('e2e full %.4g (%.2f ms/step) targets %.2f' % (f.get('value',0), f.get('ms_per_step',0), f.get('pair_targets_per_event',0)))
print('sync %.4g' % e['synchronous']['value'], 'link', {k:v for k,v in e['link_gb_per_s'].items() if k!='note'})
print('single_chain', d.get('single_chain',{}).get('ns_per_event'), 'cpu', d.get('cpu_baseline',{}).get('value'), 'launches', d['gpu_launches'])
PY
for w in c3 c5; do timeout 300 python bench.py --workload $w --no-cpu-baseline 2>> gpurun_out/r2v_bench.err | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); e=d['e2e']; f=e.get('full_copy',{})
print('$w value %.4g | e2e %.4g (%.2f ms) d2h %d B | full %.4g | sync %.4g' % (d['value'], e['value'], e['ms_per_step'], e['d2h_bytes_per_step'], f.get('value',0), e['synchronous']['value']))"; done
