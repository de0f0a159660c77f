mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_spec.py -q -m gpu -x -k "coulomb" > gpurun_out/r2N_pytest_spec.log 2>&1; echo "coulomb spec tests rc=$?"; grep -E "passed|failed|Error|^E  " gpurun_out/r2N_pytest_spec.log | head -12 | cut -c1-300
for n in 64 128 256; do for spec in 1 0; do
  echo "== bench c3 N=$n ECMC_SPEC=$spec"
  ECMC_SPEC=$spec timeout 300 python bench.py --workload c3 --particles $n --no-cpu-baseline 2>> gpurun_out/r2N.err | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); e=d['e2e']
print('value %.4g (%.2f ms) e2e %.4g kernel %s' % (d['value'], d['ms_per_step'], e['value'], d['roofline']['kernel']))"
done; done
tail -3 gpurun_out/r2N.err | cut -c1-200
