mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_chains.py -q -m gpu -x -k "sparse" > gpurun_out/r2s_pytest.log 2>&1; echo "tests rc=$?"
grep -E "passed|failed|^E  |^FAILED" gpurun_out/r2s_pytest.log | cut -c1-260 | tail -12
for sl in 4 8 16 32; do
  echo "== ECMC_FUSED_SLICES=$sl (copy engine + fused kernel)"
  ECMC_FUSED_SLICES=$sl timeout 300 python bench.py --no-cpu-baseline --no-single-chain 2>> gpurun_out/r2s.err | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); e=d['e2e']; f=e.get('full_copy',{})
print('value %.4g | e2e sparse %.4g (%.2f ms) d2h %d B | full %.4g (%.2f ms) | sync %.4g | launches %d' % (d['value'], e['value'], e['ms_per_step'], e['d2h_bytes_per_step'], f.get('value',0), f.get('ms_per_step',0), e['synchronous']['value'], d['gpu_launches']))"
done
tail -5 gpurun_out/r2s.err
echo "== C2 kernel variants (probe: 4096 chains x 1024 particles, 1024 events per step)"
for lib in build_variants/base.so build_variants/a_unroll2.so build_variants/b_prefilter.so build_variants/ab.so build_variants/base.so; do
  echo "-- $lib"
  JELLYFYSH_B200_LIBRARY=$PWD/$lib timeout 200 python tools/probe.py 4096 1024 12 1024 2>&1 | grep -E "step [2-5]|rror" | cut -c1-150
done
echo "== timeline of pipelined host steps (staged full copy)"
timeout 200 python tools/probe_timeline.py 0 3 > gpurun_out/r2s_timeline_full.txt 2>&1; head -14 gpurun_out/r2s_timeline_full.txt | cut -c1-150
echo "== timeline (sparse, fused)"
timeout 200 python tools/probe_timeline.py 1 3 > gpurun_out/r2s_timeline_sparse.txt 2>&1; head -10 gpurun_out/r2s_timeline_sparse.txt | cut -c1-150
