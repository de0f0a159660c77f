mkdir -p gpurun_out
export NCU_SUMMARY_DIR=gpurun_out
timeout 400 python bench.py > gpurun_out/r2F_bench.json 2> gpurun_out/r2F_bench.err; echo "bench rc=$?"; cut -c1-300 gpurun_out/r2F_bench.json
timeout 400 python bench.py --impl reference > gpurun_out/r2F_bench_reference.json 2> gpurun_out/r2F_bench_reference.err; echo "ref rc=$?"; cut -c1-200 gpurun_out/r2F_bench_reference.json
for w in c1 c3 c4 c5; do
  timeout 400 python bench.py --workload $w > gpurun_out/r2F_bench_$w.json 2> gpurun_out/r2F_bench_$w.err; echo "bench $w rc=$?"; cut -c1-200 gpurun_out/r2F_bench_$w.json; tail -3 gpurun_out/r2F_bench_$w.err | cut -c1-300
done
timeout 400 python bench.py --workload c3 --particles 512 > gpurun_out/r2F_bench_c3_512.json 2> gpurun_out/r2F_bench_c3_512.err; echo "bench c3 512 rc=$?"; cut -c1-200 gpurun_out/r2F_bench_c3_512.json; tail -3 gpurun_out/r2F_bench_c3_512.err | cut -c1-300
timeout 120 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2F_launches.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-single-chain > gpurun_out/r2F_bench_under_ncu.log 2>&1; echo "launch list rc=$?"
timeout 200 ncu --set full --clock-control none --import-source on -k regex:lj_spec -s 3 -c 1 -f -o gpurun_out/r2F_c2 python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-single-chain --e2e-steps 1 > gpurun_out/r2F_ncu_c2.log 2>&1; echo "ncu c2 rc=$?"
python tools/ncu_summary.py gpurun_out/r2F_c2.ncu-rep r2F_c2 4194304 "lj_spec_kernel<record=0, prune=1, lanes=4, warps=14>" > /dev/null
python tools/ncu_lines.py gpurun_out/r2F_c2.ncu-rep jellyfysh_b200/libecmc_b200.so lj_spec_kernelILb0ELb1ELi4E 4194304 > gpurun_out/r2F_c2_lines.txt 2>&1
timeout 200 ncu --set full --clock-control none --import-source on -k regex:event_kernel -s 3 -c 1 -f -o gpurun_out/r2F_c3 python bench.py --workload c3 --steps 2 --warmup 3 --no-cpu-baseline --e2e-steps 1 > gpurun_out/r2F_ncu_c3.log 2>&1; echo "ncu c3 rc=$?"
python tools/ncu_summary.py gpurun_out/r2F_c3.ncu-rep r2F_c3 4096000 "event_kernel<cand=IPCB, real=MIC, veto=MIC, record=0, warps=14>" > /dev/null
rm -f gpurun_out/r2F_c3.ncu-rep
timeout 200 ncu --set full --clock-control none --import-source on -k regex:event_kernel -s 3 -c 1 -f -o gpurun_out/r2F_c1 python bench.py --workload c1 --steps 2 --warmup 3 --no-cpu-baseline --e2e-steps 1 > gpurun_out/r2F_ncu_c1.log 2>&1; echo "ncu c1 rc=$?"
python tools/ncu_summary.py gpurun_out/r2F_c1.ncu-rep r2F_c1 16384000 "event_kernel<cand=HS, real=none, veto=none, composite, record=0, warps=14>" > /dev/null
rm -f gpurun_out/r2F_c1.ncu-rep
timeout 300 ncu --set full --clock-control none --import-source on -k regex:molecule_kernel -s 3 -c 1 -f -o gpurun_out/r2F_c4 python bench.py --workload c4 --steps 2 --warmup 3 --no-cpu-baseline --e2e-steps 1 > gpurun_out/r2F_ncu_c4.log 2>&1; echo "ncu c4 rc=$?"
python tools/ncu_summary.py gpurun_out/r2F_c4.ncu-rep r2F_c4 512000 "molecule_kernel<cand=IPCB, real=MIC, veto=MIC, record=0>" > /dev/null
python tools/ncu_lines.py gpurun_out/r2F_c4.ncu-rep jellyfysh_b200/libecmc_b200.so molecule_kernelILi7ELi6ELi103ELi2ELb0E 512000 ecmc_molecules.cuh > gpurun_out/r2F_c4_lines.txt 2>&1
timeout 1500 python -m pytest tests -q -m gpu > gpurun_out/r2F_pytest_gpu.log 2>&1; echo "tests rc=$?"; tail -5 gpurun_out/r2F_pytest_gpu.log
ls -la gpurun_out/; du -sh gpurun_out
