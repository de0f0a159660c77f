mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_gpu_spec.py tests/test_gpu_potentials.py tests/test_gpu_chains.py tests/test_gpu_full_size_parity.py -x -q > gpurun_out/r2d_tests.log 2>&1; echo "tests rc=$?"; tail -6 gpurun_out/r2d_tests.log
for v in "ECMC_SPEC=0" "ECMC_SPEC_PRUNE=0 ECMC_SPEC_LANES=4" "ECMC_SPEC_PRUNE=0 ECMC_SPEC_LANES=8" "ECMC_SPEC_PRUNE=1 ECMC_SPEC_LANES=4" "ECMC_SPEC_PRUNE=1 ECMC_SPEC_LANES=8"; do
  echo "== $v"; env $v timeout 120 python tools/probe.py 4096 1024 12 1024 2>&1 | grep -E "step [2-4]|rror" | cut -c1-150
done > gpurun_out/r2d_probe.txt 2>&1
cat gpurun_out/r2d_probe.txt
for v in "ECMC_SPEC_PRUNE=0" "ECMC_SPEC_PRUNE=1" "ECMC_SPEC_PRUNE=1 ECMC_SPEC_LANES=8"; do
  echo "== C5 $v"; env $v timeout 120 python tools/probe.py 1 65536 48 50000 2>&1 | grep -E "step [2-3]|rror" | cut -c1-150
done > gpurun_out/r2d_probe_c5.txt 2>&1
cat gpurun_out/r2d_probe_c5.txt
ECMC_SPEC_PRUNE=1 ECMC_SPEC_LANES=4 timeout 150 ncu --set full --clock-control none --import-source on -k regex:lj_spec -s 2 -c 1 -f -o gpurun_out/r2d_spec_prune python tools/probe.py 4096 1024 12 1024 > gpurun_out/r2d_ncu_prune.log 2>&1; echo "ncu rc=$?"
ECMC_SPEC_PRUNE=0 ECMC_SPEC_LANES=4 timeout 150 ncu --set full --clock-control none --import-source on -k regex:lj_spec -s 2 -c 1 -f -o gpurun_out/r2d_spec_full python tools/probe.py 4096 1024 12 1024 > gpurun_out/r2d_ncu_full.log 2>&1; echo "ncu rc=$?"
