"""Development tool: where does the time of the host-buffer call go?"""
import sys, time
import numpy as np
import torch
sys.path.insert(0, ".")
from jellyfysh_b200 import engine, workloads

n_chains, n, cells, events = 4096, 1024, 12, 1024
builder, length = workloads.lennard_jones(n_particles=n, cells_per_side=cells)
positions = workloads.lattice_start(n_chains, n, cells, length)
pin_in = torch.from_numpy(positions).pin_memory()
pin_out = torch.empty_like(pin_in).pin_memory()
host_in, host_out = pin_in.numpy(), pin_out.numpy()
with engine.Engine(builder, n_chains=n_chains) as eng:
    for rep in range(3):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        eng.upload_positions(host_in); t1 = time.perf_counter()
        eng.start(); t2 = time.perf_counter()
        eng.run(max_events=events); eng.sync(); t3 = time.perf_counter()
        eng._check(eng._lib.ecmc_download_positions(eng._h, host_out.ctypes.data)); t4 = time.perf_counter()
        print(f"upload {1e3*(t1-t0):.2f} ms  start {1e3*(t2-t1):.2f}  run {1e3*(t3-t2):.2f}  download {1e3*(t4-t3):.2f}  total {1e3*(t4-t0):.2f}")
        t0 = time.perf_counter()
        eng.run_from_host(host_in, max_events=events, out=host_out)
        print(f"run_from_host {1e3*(time.perf_counter()-t0):.2f} ms")
    a = torch.empty(100663296 // 8, dtype=torch.float64, device="cuda")
    for rep in range(3):
        torch.cuda.synchronize(); t0 = time.perf_counter(); a.copy_(pin_in.view(-1), non_blocking=True); torch.cuda.synchronize()
        t1 = time.perf_counter(); pin_out.view(-1).copy_(a, non_blocking=True); torch.cuda.synchronize(); t2 = time.perf_counter()
        print(f"torch H2D {100.66/(t1-t0)/1e3:.1f} GB/s  D2H {100.66/(t2-t1)/1e3:.1f} GB/s")
