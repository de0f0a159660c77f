mkdir -p gpurun_out
export NCU_SUMMARY_DIR=gpurun_out
timeout 500 python bench.py > gpurun_out/r2V_bench.json 2> gpurun_out/r2V_bench.err; echo "bench rc=$?"; cut -c1-200 gpurun_out/r2V_bench.json
timeout 500 python bench.py --impl reference > gpurun_out/r2V_bench_reference.json 2> gpurun_out/r2V_bench_reference.err; echo "ref rc=$?"; cut -c1-200 gpurun_out/r2V_bench_reference.json
for w in c1 c1n c3 c4 c5; do
  timeout 400 python bench.py --workload $w > gpurun_out/r2V_bench_$w.json 2> gpurun_out/r2V_bench_$w.err; echo "bench $w rc=$?"; cut -c1-160 gpurun_out/r2V_bench_$w.json; tail -2 gpurun_out/r2V_bench_$w.err | cut -c1-200
done
timeout 400 python bench.py --workload c3 --particles 512 > gpurun_out/r2V_bench_c3_512.json 2> gpurun_out/r2V_bench_c3_512.err; echo "bench c3 512 rc=$?"; cut -c1-160 gpurun_out/r2V_bench_c3_512.json
timeout 400 python bench.py --events 1024 --no-cpu-baseline --no-single-chain > gpurun_out/r2V_bench_1024.json 2> gpurun_out/r2V_bench_1024.err; echo "bench 1024 rc=$?"; cut -c1-160 gpurun_out/r2V_bench_1024.json
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2V_launches.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-single-chain > gpurun_out/r2V_bench_under_ncu.log 2>&1; echo "launch list rc=$?"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:lj_spec -s 3 -c 1 -f -o gpurun_out/r2V_c2 python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-single-chain --e2e-steps 1 > gpurun_out/r2V_ncu_c2.log 2>&1; echo "ncu c2 rc=$?"
python tools/ncu_summary.py gpurun_out/r2V_c2.ncu-rep r2V_c2 16777216 "lj_spec_kernel<record=0, prune=1, lanes=4, warps=14>" > /dev/null
python tools/ncu_lines.py gpurun_out/r2V_c2.ncu-rep jellyfysh_b200/libecmc_b200.so lj_spec_kernelILb0ELb1ELi4E 16777216 > gpurun_out/r2V_c2_lines.txt 2>&1
rm -f gpurun_out/r2V_c2.ncu-rep
timeout 300 ncu --set full --clock-control none --import-source on -k regex:molecule_kernel -s 3 -c 1 -f -o gpurun_out/r2V_c4 python bench.py --workload c4 --steps 2 --warmup 3 --no-cpu-baseline --e2e-steps 1 > gpurun_out/r2V_ncu_c4.log 2>&1; echo "ncu c4 rc=$?"
python tools/ncu_summary.py gpurun_out/r2V_c4.ncu-rep r2V_c4 1184000 "molecule_kernel<cand=IPCB, real=MIC, veto=MIC, record=0, warps=16>" > /dev/null
python tools/ncu_lines.py gpurun_out/r2V_c4.ncu-rep jellyfysh_b200/libecmc_b200.so molecule_kernelILi7ELi6ELi103ELi2ELb0ELi16E 1184000 ecmc_molecules.cuh > gpurun_out/r2V_c4_lines.txt 2>&1
rm -f gpurun_out/r2V_c4.ncu-rep
python - <<'PY'
import json
for n in ("c2", "c4"):
    s = json.load(open("gpurun_out/r2V_%s_ncu_summary.json" % n))
    print(n, {k: s[k] for k in ("kernel", "duration_ms", "warp_instructions_per_event", "issue_active_pct", "fp64_pipe_pct", "registers_per_thread", "dram_bytes_per_event")}, s["stall_per_issue"])
PY
ls -la gpurun_out/ | grep r2V
