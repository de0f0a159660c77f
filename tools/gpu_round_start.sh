#!/bin/bash
# First GPU call of a round, as one gpurun command (about 8 minutes of box time):
#   gpurun --timeout 600 -- 'bash tools/gpu_round_start.sh r2'
# GPU suite, the bench line of both arms, the launch list of the bench command and one full ncu capture of the C2
# event kernel at the bench's launch size. Everything lands under gpurun_out/; afterwards, here:
#   python tools/ncu_summary.py gpurun_out/<tag>_event_kernel.ncu-rep <tag> 4194304
#   cp gpurun_out/<tag>_bench.json profiles/ ; cp gpurun_out/<tag>_launches.csv profiles/
tag=${1:-rN}
mkdir -p gpurun_out
timeout 260 python -m pytest tests -m gpu -x -q > gpurun_out/${tag}_pytest_gpu.log 2>&1
echo "pytest rc=$?"; tail -2 gpurun_out/${tag}_pytest_gpu.log
timeout 170 python bench.py > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err
echo "bench rc=$?"; cut -c1-400 gpurun_out/${tag}_bench.json
timeout 90 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv \
    --log-file gpurun_out/${tag}_launches.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline \
    > gpurun_out/${tag}_bench_under_ncu.log 2>&1
echo "launch list rc=$?"
# the probe launches the bench kernel with the bench's chain count; 1024 events per chain = the bench's launch size
timeout 120 ncu --set full --clock-control none --import-source on -k regex:event_kernel -s 2 -c 1 -f \
    -o gpurun_out/${tag}_event_kernel python tools/probe.py 4096 1024 12 1024 > gpurun_out/${tag}_ncu_full.log 2>&1
echo "ncu full rc=$?"
