mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_chains.py tests/test_gpu_mediator.py -q -m gpu -x -k "submitted or estimators or at_scale" -s > gpurun_out/r2k_pytest.log 2>&1; echo "tests rc=$?"; grep -E "KS distance|passed|failed|^E " gpurun_out/r2k_pytest.log | cut -c1-300 | head
for sl in 4 8 16; do
  echo "== ECMC_HOST_SLICES=$sl"
  ECMC_HOST_SLICES=$sl timeout 300 python bench.py --no-cpu-baseline --no-single-chain 2>> gpurun_out/r2k.err | python -c "import json,sys; d=json.loads(sys.stdin.read()); e=d['e2e']; print('value %.4g e2e %.4g (%.2f ms) sync %.4g (%.2f ms) link %s' % (d['value'], e['value'], e['ms_per_step'], e['synchronous']['value'], e['synchronous']['ms_per_step'], e['link_gb_per_s']))"
done
tail -5 gpurun_out/r2k.err
