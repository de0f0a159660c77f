mkdir -p gpurun_out
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 8 --steps 20 --warmup 3 2> gpurun_out/r2w_bench_8gpu.err > gpurun_out/r2w_bench_8gpu.json; echo "bench8 rc=$?"; tail -3 gpurun_out/r2w_bench_8gpu.err | cut -c1-200
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2w_bench_8gpu.json').read().strip().splitlines()[-1])
e=d['e2e']
print('value %.4g  ms/step %.3f' % (d['value'], d['ms_per_step']))
print('e2e sparse %.4g (%.2f ms/step, %d steps) d2h %d B  host GB/s/rank %.1f' % (e['value'], e['ms_per_step'], e['steps'], e['d2h_bytes_per_step'], e['host_gb_per_s_per_rank']))
f=e.get('full_copy',{}); print('e2e full %.4g (%.2f ms/step)' % (f.get('value',0), f.get('ms_per_step',0)))
print('sync %.4g' % e['synchronous']['value'], 'link', {k:v for k,v in e['link_gb_per_s'].items() if k!='note'})
PY
