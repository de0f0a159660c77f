import sys, time, os
import numpy as np, torch
sys.path.insert(0, ".")
from jellyfysh_b200 import engine, workloads
n_chains, n, cells, events = 4096, 1024, 12, 1024
builder, length = workloads.lennard_jones(n_particles=n, cells_per_side=cells)
positions = workloads.lattice_start(n_chains, n, cells, length)
pin_in = torch.from_numpy(positions).pin_memory(); pin_out = torch.empty_like(pin_in).pin_memory()
host_in, host_out = pin_in.numpy(), pin_out.numpy()
for slices in (4, 8, 16):
    os.environ["ECMC_HOST_SLICES"] = str(slices)
    with engine.Engine(builder, n_chains=n_chains) as eng:
        for _ in range(2): eng.run_from_host(host_in, max_events=events, out=host_out)
        torch.cuda.synchronize(); t0 = time.perf_counter()
        for _ in range(8): eng.run_from_host(host_in, max_events=events, out=host_out)
        dt = (time.perf_counter() - t0) / 8
        print(f"slices {slices}: {1e3*dt:.2f} ms per call -> {n_chains*events/dt:.3e} events/s", flush=True)
