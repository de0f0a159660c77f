"""Development tool: replay a golden trace on the GPU and dump the records to gpurun_out/ for offline diffing."""
import sys

import numpy as np

sys.path.insert(0, ".")
sys.path.insert(0, "tests")
sys.path.insert(0, "tests/golden")
import trace_util as tu  # noqa: E402
from jellyfysh_b200 import engine  # noqa: E402
from jellyfysh_b200.program import ProgramBuilder  # noqa: E402

name = sys.argv[1]
g = tu.load_trace(name)
n = len(g["records"])
with engine.Engine(tu.builder_of(g, ProgramBuilder), n_chains=1) as eng:
    eng.upload_positions(g["positions0"][None], None if tu.charges_of(g) is None else tu.charges_of(g)[None])
    eng.start(first_stream=int(g["seed"][1]))
    rec, stats = eng.run_recorded(max_events=n, records_per_chain=n)
np.save(f"gpurun_out/{name}_records.npy", rec[0])
print(stats)
