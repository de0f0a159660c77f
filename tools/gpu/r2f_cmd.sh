mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_full_size.py tests/test_gpu_full_size_parity.py -q -m gpu > gpurun_out/r2f_pytest.log 2>&1; echo "tests rc=$?"; tail -3 gpurun_out/r2f_pytest.log
timeout 400 python bench.py > gpurun_out/r2f_bench.json 2> gpurun_out/r2f_bench.err; echo "bench rc=$?"; cut -c1-300 gpurun_out/r2f_bench.json
timeout 400 python bench.py --impl reference > gpurun_out/r2f_bench_reference.json 2> gpurun_out/r2f_bench_reference.err; echo "ref rc=$?"; cut -c1-200 gpurun_out/r2f_bench_reference.json
for w in c1 c3 c4 c5; do
  timeout 400 python bench.py --workload $w > gpurun_out/r2f_bench_$w.json 2> gpurun_out/r2f_bench_$w.err; echo "bench $w rc=$?"; cut -c1-200 gpurun_out/r2f_bench_$w.json
done
timeout 400 python bench.py --workload c3 --particles 512 > gpurun_out/r2f_bench_c3_512.json 2> gpurun_out/r2f_bench_c3_512.err; echo "bench c3 512 rc=$?"; cut -c1-200 gpurun_out/r2f_bench_c3_512.json
timeout 120 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2f_launches.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-single-chain > gpurun_out/r2f_bench_under_ncu.log 2>&1; echo "launch list rc=$?"
timeout 200 ncu --set full --clock-control none --import-source on -k regex:lj_spec -s 3 -c 1 -f -o gpurun_out/r2f_c2 python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-single-chain --e2e-steps 1 > gpurun_out/r2f_ncu_c2.log 2>&1; echo "ncu c2 rc=$?"
timeout 200 ncu --set full --clock-control none --import-source on -k regex:event_kernel -s 3 -c 1 -f -o gpurun_out/r2f_c3 python bench.py --workload c3 --steps 2 --warmup 3 --no-cpu-baseline --e2e-steps 1 > gpurun_out/r2f_ncu_c3.log 2>&1; echo "ncu c3 rc=$?"
timeout 200 ncu --set full --clock-control none --import-source on -k regex:event_kernel -s 3 -c 1 -f -o gpurun_out/r2f_c1 python bench.py --workload c1 --steps 2 --warmup 3 --no-cpu-baseline --e2e-steps 1 > gpurun_out/r2f_ncu_c1.log 2>&1; echo "ncu c1 rc=$?"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:molecule_kernel -s 3 -c 1 -f -o gpurun_out/r2f_c4 python bench.py --workload c4 --steps 2 --warmup 3 --no-cpu-baseline --e2e-steps 1 > gpurun_out/r2f_ncu_c4.log 2>&1; echo "ncu c4 rc=$?"
ls -la gpurun_out/*.ncu-rep
