"""Store the reference's statistical fixtures (cumulative histograms of observables obtained with reversible Monte
Carlo, shipped under jellyfysh/output/) as compact arrays: tests/golden/reference_cdfs.npz.
Run where the reference checkout is available: python tests/golden/make_reference_cdf.py [/root/reference]"""
import os
import sys

import numpy as np

ref = sys.argv[1] if len(sys.argv) > 1 else "/root/reference"
out = {}
for key, path in {"coulomb_atoms": "jellyfysh/output/2018_JCP_149_064113/coulomb_atoms/ReferenceDataCoulombAtoms.dat",
                  "dipoles_px": "jellyfysh/output/hard_disk_dipoles/ReferenceDataPx_81Dipoles_NewtonianECMC.dat",
                  "dipoles_py": "jellyfysh/output/hard_disk_dipoles/ReferenceDataPy_81Dipoles_NewtonianECMC.dat",
                  "water_oo": "jellyfysh/output/2018_JCP_149_064113/water/ReferenceOOSeparation.dat",
                  "dipoles_13": "jellyfysh/output/2018_JCP_149_064113/dipoles/ReferenceDataDipoles_13.dat",
                  "dipoles_14": "jellyfysh/output/2018_JCP_149_064113/dipoles/ReferenceDataDipoles_14.dat",
                  "water_angle": "jellyfysh/output/2018_JCP_149_064113/water/ReferenceAngleSingleMolecule.dat",
                  "water_length": "jellyfysh/output/2018_JCP_149_064113/water/ReferenceLengthSingleMolecule.dat"}.items():
    data = np.loadtxt(os.path.join(ref, path))
    out[key + "_x"], out[key + "_cdf"] = data[:, 0], data[:, 1]
np.savez_compressed(os.path.join(os.path.dirname(os.path.abspath(__file__)), "reference_cdfs.npz"), **out)
print({k: v.shape for k, v in out.items()})
