mkdir -p gpurun_out
export NCU_SUMMARY_DIR=gpurun_out
timeout 300 ncu --set full --clock-control none --import-source on -k regex:lj_spec -s 3 -c 1 -f -o gpurun_out/r2O_c3 python bench.py --workload c3 --steps 2 --warmup 3 --no-cpu-baseline --e2e-steps 1 > gpurun_out/r2O_ncu_c3.log 2>&1; echo "ncu c3 rc=$?"
python tools/ncu_summary.py gpurun_out/r2O_c3.ncu-rep r2O_c3 4096000 "lj_spec_kernel<coulomb, record=0, prune=1, lanes=4, warps=28>" > /dev/null
python tools/ncu_lines.py gpurun_out/r2O_c3.ncu-rep jellyfysh_b200/libecmc_b200.so lj_spec_kernelILb0ELb1ELi4ELi14ELb0ELi1E 4096000 > gpurun_out/r2O_c3_lines.txt 2>&1
rm -f gpurun_out/r2O_c3.ncu-rep
python - <<'PY'
import json
s=json.load(open('gpurun_out/r2O_c3_ncu_summary.json'))
print({k:s[k] for k in ('duration_ms','warp_instructions_per_event','issue_active_pct','fp64_pipe_pct','registers_per_thread')}, s['stall_per_issue'])
print('LDL/event', s['local_load_instructions']/s['events_per_launch'], 'STL', s['local_store_instructions']/s['events_per_launch'])
PY
head -2 gpurun_out/r2O_c3_lines.txt; sort -k2 -n -r gpurun_out/r2O_c3_lines.txt | head -24

