mkdir -p gpurun_out
export NCU_SUMMARY_DIR=gpurun_out
timeout 1500 python -m pytest tests -q -m gpu --durations=8 > gpurun_out/r2Z_pytest_gpu.txt 2>&1; echo "pytest rc=$?"; tail -14 gpurun_out/r2Z_pytest_gpu.txt
timeout 400 python bench.py --workload c4 > gpurun_out/r2Z_bench_c4.json 2> gpurun_out/r2Z_bench_c4.err; echo "bench c4 rc=$?"; cut -c1-160 gpurun_out/r2Z_bench_c4.json
timeout 400 python bench.py --workload c4 --chains 1024 --no-cpu-baseline > gpurun_out/r2Z_bench_c4_1024.json 2> gpurun_out/r2Z_bench_c4_1024.err; echo "bench c4 1024 rc=$?"; cut -c1-160 gpurun_out/r2Z_bench_c4_1024.json
timeout 300 ncu --set full --clock-control none --import-source on -k regex:molecule_kernel -s 3 -c 1 -f -o gpurun_out/r2Z_c4 python bench.py --workload c4 --steps 2 --warmup 3 --no-cpu-baseline --e2e-steps 1 > gpurun_out/r2Z_ncu_c4.log 2>&1; echo "ncu c4 rc=$?"
python tools/ncu_summary.py gpurun_out/r2Z_c4.ncu-rep r2Z_c4 1184000 "molecule_kernel<cand=IPCB, real=MIC, veto=MIC, record=0, warps=16>" > /dev/null
python tools/ncu_lines.py gpurun_out/r2Z_c4.ncu-rep jellyfysh_b200/libecmc_b200.so molecule_kernelILi7ELi6ELi103ELi2ELb0ELi16E 1184000 ecmc_molecules.cuh > gpurun_out/r2Z_c4_lines.txt 2>&1
rm -f gpurun_out/r2Z_c4.ncu-rep
python - <<'PY'
import json
s = json.load(open("gpurun_out/r2Z_c4_ncu_summary.json"))
print({k: s[k] for k in ("kernel", "duration_ms", "warp_instructions_per_event", "issue_active_pct", "fp64_pipe_pct", "registers_per_thread", "dram_bytes_per_event", "l2_hit_pct")}, s["stall_per_issue"])
PY
