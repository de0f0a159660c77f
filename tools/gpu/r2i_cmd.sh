mkdir -p gpurun_out
export NCU_SUMMARY_DIR=gpurun_out
timeout 900 python -m pytest tests/test_gpu_mediator.py tests/test_gpu_chains.py -q -m gpu -x -k "estimators or horizon or at_scale or state_handler or several_engines" -s > gpurun_out/r2i_pytest.log 2>&1; echo "tests rc=$?"; grep -E "KS distance|horizon|passed|failed|Error" gpurun_out/r2i_pytest.log | cut -c1-300
timeout 300 ncu --set full --clock-control none --import-source on -k regex:molecule_kernel -s 3 -c 1 -f -o gpurun_out/r2i_c4 python bench.py --workload c4 --steps 2 --warmup 3 --no-cpu-baseline --e2e-steps 1 > gpurun_out/r2i_ncu_c4.log 2>&1; echo "ncu c4 rc=$?"
python tools/ncu_summary.py gpurun_out/r2i_c4.ncu-rep r2i_c4 512000 "molecule_kernel wpc2" > /dev/null
python tools/ncu_lines.py gpurun_out/r2i_c4.ncu-rep jellyfysh_b200/libecmc_b200.so molecule_kernelILi7ELi6ELi103ELi2ELb0ELi16ELb1ELi2E 512000 ecmc_molecules.cuh > gpurun_out/r2i_c4_lines.txt 2>&1
rm -f gpurun_out/r2i_c4.ncu-rep
