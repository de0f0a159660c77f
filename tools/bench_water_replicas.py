"""C4 as SURVEY.md 8(d) frames it: a replica sweep of SPC/Fw water over the GPUs of one node, every rank its own chains,
the oxygen-oxygen separation histogram accumulated on the device and summed over the ranks with ONE NCCL all-reduce.

    python tools/bench_water_replicas.py                              # one GPU
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
        tools/bench_water_replicas.py                                 # two GPUs

Prints one JSON line on rank 0: events/s summed over ranks (device time, max over ranks), the histogram's sample count."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for path in (ROOT, os.path.join(ROOT, "tests"), os.path.join(ROOT, "tests", "golden")):
    if path not in sys.path:
        sys.path.insert(0, path)

import numpy as np  # noqa: E402
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

import configs  # noqa: E402
import trace_util as tu  # noqa: E402
from jellyfysh_b200 import engine, sharding  # noqa: E402
from jellyfysh_b200.program import ProgramBuilder  # noqa: E402


def main():
    rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    local_rank = int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    n_molecules, n_chains, events, samples = 32, 1024, 250, 8
    g = dict(tu.load_trace("trace_water"))
    g["meta_n"] = np.asarray(3 * n_molecules)
    pb = tu.water_builder_of(g, ProgramBuilder)
    first_chain, _ = sharding.chain_shard(rank, world, n_chains)
    roots = np.empty((n_chains, n_molecules, 3))
    leaves = np.empty((n_chains, 3 * n_molecules, 3))
    for c in range(n_chains):
        r, l = configs.water_start(n_molecules, 10.0, seed=first_chain + c)
        roots[c], leaves[c] = r, l.reshape(-1, 3)
    histogram = np.zeros(1000, dtype=np.uint64)
    with engine.Engine(pb, n_chains=n_chains, device=local_rank) as eng:
        eng.upload_positions(leaves, np.tile([0.41, -0.82, 0.41], (n_chains, n_molecules)))
        eng.upload_roots(roots)
        eng.start(first_stream=first_chain)
        eng.run(max_events=events)
        eng.sync()
        before = eng.kernel_seconds
        total = 0
        for _ in range(samples):
            eng.run(max_events=events)
            total += eng.sync()["events"]
            # oxygen-oxygen separations: every third leaf starting at 1, bins of the reference's plot script on [2, 7]
            eng.separation_histogram(1000, 2.0, 7.0, out=histogram, first=1, stride=3)
        seconds = eng.kernel_seconds - before
    device = torch.device("cuda", local_rank)
    summed = sharding.reduce_histogram(histogram.astype(np.int64), device=device)
    all_events = int(sharding.reduce_histogram([total], device=device)[0])
    slowest = float(sharding.reduce_max([seconds], device=device)[0])
    if rank == 0:
        print(json.dumps({"workload": "C4 replica sweep: SPC/Fw water, 32 molecules, 1024 chains per GPU", "n_gpus": world,
                          "events": all_events, "device_seconds_max_over_ranks": slowest,
                          "events_per_sec": all_events / slowest,
                          "oo_histogram_samples": int(summed.sum()), "oo_histogram_bins": len(summed),
                          "collective": "one all-reduce (NCCL) of the int64 histogram"}))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
