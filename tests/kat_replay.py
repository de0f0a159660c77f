"""Replay helpers for the fixtures under tests/golden: turn a harvested known-answer record of the reference's
unit tests (kats.json) or a potential vector set (potentials.npz) into calls on an implementation under test."""
import json
import os

import numpy as np

from jellyfysh_b200 import abi

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load_kats():
    with open(os.path.join(GOLDEN, "kats.json")) as handle:
        return json.load(handle)["records"]


def load_npz(name):
    return np.load(os.path.join(GOLDEN, name + ".npz"))


def potential_of(cls, init):
    if cls == "LennardJonesPotential":
        return abi.EcmcPotential.make(abi.POT_LENNARD_JONES, init["prefactor"], init["characteristic_length"])
    if cls == "InversePowerPotential":
        return abi.EcmcPotential.make(abi.POT_INVERSE_POWER, init["power"], init["prefactor"])
    if cls == "DisplacedEvenPowerPotential":
        return abi.EcmcPotential.make(abi.POT_DISPLACED_EVEN_POWER, init["prefactor"], init["equilibrium_separation"],
                                      init["power"])
    if cls == "HardSpherePotential":
        return abi.EcmcPotential.make(abi.POT_HARD_SPHERE, init["radius"])
    if cls == "HardDipolePotential":
        return abi.EcmcPotential.make(abi.POT_HARD_DIPOLE, init["minimum_separation"], init["maximum_separation"])
    if cls == "MergedImageCoulombPotential":
        return abi.EcmcPotential.make(abi.POT_MERGED_IMAGE_COULOMB, init["prefactor"], init["alpha"],
                                      init["fourier_cutoff"], init["position_cutoff"])
    if cls == "InversePowerCoulombBoundingPotential":
        return abi.EcmcPotential.make(abi.POT_INVERSE_POWER_COULOMB_BOUNDING, init["prefactor"])
    raise KeyError(cls)


USES_CHARGES = {"InversePowerPotential", "MergedImageCoulombPotential", "InversePowerCoulombBoundingPotential"}
NEEDS_POTENTIAL_CHANGE = {"LennardJonesPotential", "InversePowerPotential", "DisplacedEvenPowerPotential",
                          "InversePowerCoulombBoundingPotential"}


def kat_call(record):
    """(potential, dimension, length, velocity, separation, c1, c2, potential_change, method, expected, places)."""
    cls = record["cls"]
    args = list(record["args"])
    velocity, separation = args[0], args[1]
    rest = args[2:]
    c1 = c2 = 1.0
    if cls in USES_CHARGES:
        c1, c2 = rest[0], rest[1]
        rest = rest[2:]
    du = rest[0] if (record["method"] == "displacement" and cls in NEEDS_POTENTIAL_CHANGE) else 0.0
    dimension = len(separation)
    length = record["length"] if record["length"] is not None else 1.0
    expected = float("inf") if record["expected"] == "inf" else record["expected"]
    return (potential_of(cls, record["init"]), dimension, length, velocity, separation, c1, c2, du,
            record["method"], expected, record["places"])


def kat_matches(value, expected, places):
    """unittest.assertAlmostEqual semantics: round(|a - b|, places) == 0, equal infinities match."""
    if value == expected:
        return True
    if np.isinf(expected) or np.isinf(value) or np.isnan(value):
        return False
    if places >= 16:
        return value == expected
    return round(abs(value - expected), places) == 0
