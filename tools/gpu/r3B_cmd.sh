mkdir -p gpurun_out
N=8
for slices in 8 16 64; do
  ECMC_FUSED_SLICES=$slices timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 295$slices bench.py --gpus $N --steps 12 --warmup 3 --no-cpu-baseline --no-single-chain > gpurun_out/r3B_bench_${N}gpu_s$slices.json 2> gpurun_out/r3B_bench_${N}gpu_s$slices.err; echo "slices $slices rc=$?"
  python - <<PY
import json
d = json.loads(open("gpurun_out/r3B_bench_${N}gpu_s$slices.json").read().strip().splitlines()[-1])
e = d["e2e"]
print("slices $slices value %.3e e2e %.3e ms/step %.3f full %.3e sync %.3e" % (d["value"], e["value"], e["ms_per_step"], e["full_copy"]["value"], e["synchronous"]["value"]))
PY
done
