mkdir -p gpurun_out
export NCU_SUMMARY_DIR=gpurun_out
timeout 300 ncu --set full --clock-control none --import-source on -k regex:lj_chain -s 2 -c 1 -f -o gpurun_out/r2n_c5 python tools/probe.py 1 65536 48 50000 > gpurun_out/r2n_ncu_c5.log 2>&1; echo "ncu rc=$?"
python tools/ncu_summary.py gpurun_out/r2n_c5.ncu-rep r2n_c5 50000 "lj_chain_kernel<record=0, prune=1, warps per chain=4>" > /dev/null
python tools/ncu_lines.py gpurun_out/r2n_c5.ncu-rep jellyfysh_b200/libecmc_b200.so lj_chain_kernelILb0ELb1E 50000 ecmc_spec_cta.cuh > gpurun_out/r2n_c5_lines.txt 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:lj_spec -s 2 -c 1 -f -o gpurun_out/r2n_c5w python tools/probe.py 1 65536 48 50000 > gpurun_out/r2n_ncu_c5w.log 2>&1
ECMC_CHAIN_BLOCKS=0 timeout 300 ncu --set full --clock-control none --import-source on -k regex:lj_spec -s 2 -c 1 -f -o gpurun_out/r2n_c5w python tools/probe.py 1 65536 48 50000 > gpurun_out/r2n_ncu_c5w.log 2>&1; echo "ncu rc=$?"
python tools/ncu_lines.py gpurun_out/r2n_c5w.ncu-rep jellyfysh_b200/libecmc_b200.so lj_spec_kernelILb0ELb1ELi4E 50000 ecmc_spec.cuh > gpurun_out/r2n_c5w_lines.txt 2>&1
rm -f gpurun_out/r2n_c5w.ncu-rep gpurun_out/r2n_c5.ncu-rep
