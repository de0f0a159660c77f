import sys
import numpy as np
sys.path.insert(0, ".")
from jellyfysh_b200 import abi, engine
g = np.load("tests/golden/potentials.npz")
out = {}
for tag in ["ip_rep", "ip_coul", "ip_six"]:
    power, k = g[tag + "_params"]
    pot = abi.EcmcPotential.make(abi.POT_INVERSE_POWER, power, k)
    sep, du, direction = g[tag + "_sep"], g[tag + "_du"], g[tag + "_dir"]
    charges = np.stack([g[tag + "_c1"], g[tag + "_c2"]], axis=1)
    disp = np.empty(len(sep))
    for d in range(3):
        m = direction == d
        disp[m] = engine.potential_displacement(pot, 3, 12.0, d, sep[m], charges[m], du[m])
    out[tag] = disp
np.savez("gpurun_out/ip_dump.npz", **out)
