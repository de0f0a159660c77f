mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_chains.py tests/test_gpu_mediator.py -x -q -m gpu -k "water or molecule or composite or dipole_config or single_molecule or root_unit or cell_bounding or lifting" --durations=5 > gpurun_out/r3A_pytest.log 2>&1; echo "pytest rc=$?"; tail -12 gpurun_out/r3A_pytest.log
for chains in 2368 1024; do
  echo "== chains=$chains"; timeout 200 python tools/probe_water.py 32 $chains 2000 2>&1 | grep -E "step [12]|rror" | cut -c1-110
done
