mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_spec.py -q -m gpu -x -k "coulomb" > gpurun_out/r2I_pytest_spec.log 2>&1; echo "coulomb spec tests rc=$?"; grep -E "passed|failed|Error|^E  " gpurun_out/r2I_pytest_spec.log | head -12 | cut -c1-300
timeout 900 python -m pytest tests/test_gpu_chains.py tests/test_gpu_full_size_parity.py tests/test_gpu_full_size.py -q -m gpu -k "coulomb or Coulomb or c3 or trace" > gpurun_out/r2I_pytest_chains.log 2>&1; echo "coulomb chain tests rc=$?"; grep -E "passed|failed|^E  |^FAILED" gpurun_out/r2I_pytest_chains.log | head -12 | cut -c1-300
for opt in 1 0; do
  echo "== bench c3 (ECMC_SPEC=$opt)"
  ECMC_SPEC=$opt timeout 300 python bench.py --workload c3 --no-cpu-baseline 2>> gpurun_out/r2I.err | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); e=d['e2e']
print('value %.4g (%.2f ms) e2e %.4g kernel %s' % (d['value'], d['ms_per_step'], e['value'], d['roofline']['kernel']))"
done
tail -3 gpurun_out/r2I.err | cut -c1-200
