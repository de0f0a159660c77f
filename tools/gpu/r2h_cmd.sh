mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu -x > gpurun_out/r2h_pytest_gpu.log 2>&1; echo "tests rc=$?"; tail -12 gpurun_out/r2h_pytest_gpu.log | cut -c1-300
for wpc in 1 2; do
  echo "== ECMC_MOLECULE_WPC=$wpc"
  ECMC_MOLECULE_WPC=$wpc timeout 300 python bench.py --workload c4 --no-cpu-baseline --e2e-steps 2 2> gpurun_out/r2h_c4_wpc$wpc.err | tee gpurun_out/r2h_c4_wpc$wpc.json | cut -c1-220
done
