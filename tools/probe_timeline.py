"""Development tool: device timeline (CUPTI through torch.profiler) of a few pipelined host steps of C2 -- which kernels and
copies run when, per stream -- to see what bounds a step of ecmc_submit_from_host[_sparse].

    python tools/probe_timeline.py [sparse 0/1] [steps] > gpurun_out/timeline.txt
"""
import sys

import numpy as np
import torch
from torch.profiler import ProfilerActivity, profile

sys.path.insert(0, ".")
from jellyfysh_b200 import engine, workloads  # noqa: E402

sparse = int(sys.argv[1]) if len(sys.argv) > 1 else 1
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 4
n_chains, n, cells, events = 4096, 1024, 12, 1024
builder, length = workloads.lennard_jones(n_particles=n, cells_per_side=cells)
positions = workloads.lattice_start(n_chains, n, cells, length)
buffers = [engine.pinned_array(positions.shape) for _ in range(2)]
buffers[0][...] = positions
torch.cuda.init()
with engine.Engine(builder, n_chains=n_chains) as eng:
    eng.set_option(eng.OPTION_CONTINUE_HOST_STEPS, 1)
    eng.upload_positions(positions)
    eng.start(first_stream=0)
    def submit(k):
        if sparse:
            eng.submit_from_host(buffers[0], first_stream=(k + 1) * n_chains, max_events=events, out=buffers[0], sparse=True)
        else:
            eng.submit_from_host(buffers[k % 2], first_stream=(k + 1) * n_chains, max_events=events, out=buffers[(k + 1) % 2])
    for k in range(3):
        submit(k)
    eng.wait()
    with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
        for k in range(3, 3 + steps):
            submit(k)
        eng.wait()
    rows = []
    for event in prof.events():
        if event.device_type.name != "CUDA":
            continue
        rows.append((event.time_range.start, event.time_range.end, event.name[:60]))
    rows.sort()
    t0 = rows[0][0]
    print(f"{len(rows)} device activities over {(rows[-1][1] - t0) / 1e3:.3f} ms for {steps} steps")
    # busy time per activity name and the union of busy intervals of the event kernels
    by_name = {}
    for start, end, name in rows:
        total, count = by_name.get(name, (0.0, 0))
        by_name[name] = (total + (end - start), count + 1)
    for name, (total, count) in sorted(by_name.items(), key=lambda item: -item[1][0]):
        print(f"{total / 1e3:10.3f} ms  {count:5d} x  mean {total / count / 1e3:8.3f} ms  {name}")
    print("activities: start [ms], duration [ms], name")
    for start, end, name in rows[:int(sys.argv[3]) if len(sys.argv) > 3 else 80]:
        print(f"{(start - t0) / 1e3:9.3f} {(end - start) / 1e3:9.3f}  {name}")
