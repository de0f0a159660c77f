mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_spec.py tests/test_gpu_full_size_parity.py -x -q -m gpu -k "coulomb or c3" --durations=5 > gpurun_out/r2Y_pytest.log 2>&1; echo "pytest rc=$?"; tail -12 gpurun_out/r2Y_pytest.log
for n in 64 128 256 512; do
  timeout 300 python bench.py --workload c3 --particles $n --no-cpu-baseline --steps 10 --e2e-steps 2 > gpurun_out/r2Y_c3_$n.json 2> gpurun_out/r2Y_c3_$n.err
  python -c "
import json; d=json.loads(open('gpurun_out/r2Y_c3_$n.json').read().strip().splitlines()[-1]); print($n, '%.3e'%d['value'], d['roofline']['kernel'])"
done
