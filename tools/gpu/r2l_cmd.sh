mkdir -p gpurun_out
N=${1:-8}
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N > gpurun_out/r2l_bench_${N}gpu.json 2> gpurun_out/r2l_bench_${N}gpu.err; echo "bench $N rc=$?"; cut -c1-250 gpurun_out/r2l_bench_${N}gpu.json
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29518 bench.py --gpus $N --workload c4 > gpurun_out/r2l_bench_c4_${N}gpu.json 2> gpurun_out/r2l_bench_c4_${N}gpu.err; echo "bench c4 $N rc=$?"; cut -c1-250 gpurun_out/r2l_bench_c4_${N}gpu.json
tail -3 gpurun_out/r2l_bench_${N}gpu.err gpurun_out/r2l_bench_c4_${N}gpu.err | cut -c1-300
