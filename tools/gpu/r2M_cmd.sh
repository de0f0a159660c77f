mkdir -p gpurun_out
for w in 7 14 28; do for l in 4 8; do
  echo "== C3 N=64 ECMC_COULOMB_WARPS=$w ECMC_SPEC_LANES=$l"
  ECMC_COULOMB_WARPS=$w ECMC_SPEC_LANES=$l timeout 100 python tools/probe_c3.py 2>&1 | tail -1 | cut -c1-120
done; done
echo "== C2 probe: default, then the batch barrier for the LJ model"
timeout 100 python tools/probe.py 4096 1024 12 1024 2>&1 | grep -E "step [2-4]" | cut -c1-100
JELLYFYSH_B200_LIBRARY=$PWD/build_variants/align_lj.so timeout 100 python tools/probe.py 4096 1024 12 1024 2>&1 | grep -E "step [2-4]" | cut -c1-100
