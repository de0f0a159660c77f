mkdir -p gpurun_out
N=8
timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N > gpurun_out/r2U_bench_${N}gpu.json 2> gpurun_out/r2U_bench_${N}gpu.err; echo "bench $N rc=$?"
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29518 bench.py --gpus $N --workload c4 > gpurun_out/r2U_bench_c4_${N}gpu.json 2> gpurun_out/r2U_bench_c4_${N}gpu.err; echo "bench c4 $N rc=$?"
python - <<'PY'
import json
for name in ("8gpu", "c4_8gpu"):
    try:
        d = json.loads(open("gpurun_out/r2U_bench_%s.json" % name).read().strip().splitlines()[-1])
        e = d["e2e"]
        print(name, "value %.3e ms/step %.3f e2e %.3e" % (d["value"], d["ms_per_step"], e["value"]), {k: (v if not isinstance(v, dict) else v.get("value")) for k, v in e.items() if k not in ("call", "link_gb_per_s")})
        print(json.dumps(e.get("link_gb_per_s"))[:300])
    except Exception as ex:
        print(name, "failed", ex)
PY
tail -3 gpurun_out/r2U_bench_8gpu.err | cut -c1-300
