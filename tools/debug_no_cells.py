import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests")); sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
import numpy as np
import trace_util as tu
from jellyfysh_b200 import engine
from jellyfysh_b200.program import ProgramBuilder
from oracle import oracle
g = tu.load_trace("trace_coulomb_power_bounded")
chain = oracle.OracleChain(tu.no_cells_builder_of(g, oracle.ProgramBuilder))
chain.set_positions(g["positions0"], tu.charges_of(g)); chain.start(stream=int(g["seed"][1]))
with engine.Engine(tu.no_cells_builder_of(g, ProgramBuilder), n_chains=1) as eng:
    eng.upload_positions(g["positions0"][None], tu.charges_of(g)[None])
    eng.start(first_stream=int(g["seed"][1]))
    for k in range(2000):
        rec, stats = eng.run_recorded(max_events=1, records_per_chain=1)
        n, ref = chain.run(max_events=1, record=1)
        dev = eng.download_positions()[0]
        err = np.abs(dev - chain.positions()).max()
        occ, sur = eng.cells()
        if err > 1e-12 or any(rec[0][f][0] != ref[f][0] for f in tu.DISCRETE_FIELDS):
            print("event", k, "err", err, "\n dev rec", rec[0][0], "\n ref rec", ref[0])
            print(" dev pos\n", dev, "\n ref pos\n", chain.positions())
            print(" occ", occ[0], "sur", sur[0], "state", eng.chain_states()[0])
            break
    else:
        print("no difference in 2000 single-event launches")
