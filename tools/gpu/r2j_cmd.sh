mkdir -p gpurun_out
for a in 0 1; do
  for w in "c3" "c3 --particles 512" "c1"; do
    echo "== ECMC_EVENT_ALIGNED=$a $w"
    ECMC_EVENT_ALIGNED=$a timeout 300 python bench.py --workload $w --no-cpu-baseline --e2e-steps 1 2>> gpurun_out/r2j.err | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['roofline']['kernel'])"
  done
done
ECMC_EVENT_ALIGNED=1 timeout 900 python -m pytest tests/test_gpu_chains.py tests/test_gpu_full_size_parity.py tests/test_gpu_full_size.py -q -m gpu -x > gpurun_out/r2j_pytest_aligned.log 2>&1; echo "aligned tests rc=$?"; tail -3 gpurun_out/r2j_pytest_aligned.log
timeout 600 python -m pytest tests/test_gpu_mediator.py -q -m gpu -x -k "estimators" > gpurun_out/r2j_pytest_est.log 2>&1; echo "estimator test rc=$?"; tail -3 gpurun_out/r2j_pytest_est.log; grep -n "^E " gpurun_out/r2j_pytest_est.log | head
