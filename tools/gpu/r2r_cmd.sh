mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_chains.py tests/test_gpu_spec.py -q -m gpu -x -k "sparse or submitted or spec" > gpurun_out/r2r_pytest.log 2>&1; echo "tests rc=$?"
grep -E "passed|failed|^E  |^FAILED" gpurun_out/r2r_pytest.log | cut -c1-260 | tail -12
for sl in 8 32 64; do
  echo "== ECMC_FUSED_SLICES=$sl"
  ECMC_FUSED_SLICES=$sl timeout 300 python bench.py --no-cpu-baseline --no-single-chain 2>> gpurun_out/r2r.err | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); e=d['e2e']; f=e.get('full_copy',{})
print('value %.4g | e2e sparse %.4g (%.2f ms) d2h %d B | full %.4g (%.2f ms) | sync %.4g | launches %d' % (d['value'], e['value'], e['ms_per_step'], e['d2h_bytes_per_step'], f.get('value',0), f.get('ms_per_step',0), e['synchronous']['value'], d['gpu_launches']))"
done
tail -5 gpurun_out/r2r.err
