mkdir -p gpurun_out
timeout 900 python bench.py 2> gpurun_out/r2u_bench.err > gpurun_out/r2u_bench.json; echo "bench rc=$?"; tail -3 gpurun_out/r2u_bench.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2u_bench.json').read().strip().splitlines()[-1])
e=d['e2e']
print('value %.4g  ms/step %.3f' % (d['value'], d['ms_per_step']))
print('e2e sparse %.4g (%.2f ms/step, %d steps) d2h %d B  host GB/s %.1f' % (e['value'], e['ms_per_step'], e['steps'], e['d2h_bytes_per_step'], e['host_gb_per_s_per_rank']))
f=e.get('full_copy',{}); print('e2e full %.4g (%.2f ms/step)' % (f.get('value',0), f.get('ms_per_step',0)))
print('sync %.4g' % e['synchronous']['value'], 'link', e['link_gb_per_s'])
print('single_chain', d.get('single_chain',{}).get('ns_per_event'), 'cpu', d.get('cpu_baseline',{}).get('value'), 'launches', d['gpu_launches'])
PY
