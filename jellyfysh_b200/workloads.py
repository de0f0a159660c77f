"""Synthetic configurations of SURVEY.md section 8(d) as engine programs + start configurations.

C2: 3D Lennard-Jones, N = 1024 per chain, L = 2048^(1/3) (density 0.5), LJ(prefactor 4, length 1), cells 12^3 with
    one neighbour layer and one occupant per cell, LJ inverted for the nearby cells and the surplus, cell veto
    (inner point estimator, prefactor 1.5, 4 points per side) for all other cells, chain_time 10, speed 1.
C5: the same with N = 65536, cells 48^3, one chain.
C3: Coulomb atoms: merged-image Coulomb bounded by the inverse-power Coulomb bound for nearby cells, cell veto else.
Start configurations come from numpy PCG64(seed = 1000 + chain), chain c reads random stream first_stream + c.
"""
import numpy as np

from jellyfysh_b200 import abi, tables
from jellyfysh_b200.program import ProgramBuilder

SEED = 20260117


def lennard_jones(n_particles=1024, cells_per_side=12, density=0.5, beta=1.0, chain_time=10.0, neighbor_layers=1,
                  estimator_prefactor=1.5, points_per_side=4, max_surplus=64, seed=SEED, device=0, veto=True):
    """(ProgramBuilder, system_length) of the Lennard-Jones configuration C2 / C5."""
    length = float((n_particles / density) ** (1.0 / 3.0))
    potential = abi.EcmcPotential.make(abi.POT_LENNARD_JONES, 4.0, 1.0)
    builder = ProgramBuilder(3, n_particles, length, beta, [cells_per_side] * 3, neighbor_layers, max_occupants=1,
                             max_surplus=max_surplus, chain_time=chain_time, speed=1.0, initial_direction=0,
                             initial_active=0, seed=seed)
    builder.set_pair(abi.PAIR_TWO_LEAF_UNIT, potential)
    if veto:
        geometry = tables.CellGeometry(3, length, [cells_per_side] * 3, neighbor_layers)
        bounds, far = tables.inner_point_derivative_bounds(potential, geometry, prefactor=estimator_prefactor,
                                                           points_per_side=points_per_side, device=device)
        builder.set_veto(potential, tables.veto_tables(bounds, far))
    return builder, length


def lattice_start(n_chains, n_particles, cells_per_side, length, first_chain=0, jitter=0.05):
    """Particle i sits in cell i (cell-index order) at the cell centre + U(-jitter, jitter)^3; chain c is seeded
    with PCG64(1000 + first_chain + c). Returns positions[n_chains][n_particles][3]."""
    if n_particles > cells_per_side ** 3:
        raise ValueError("more particles than cells")
    side = length / cells_per_side
    cells = np.arange(n_particles)
    centres = np.stack([(cells // cells_per_side ** d) % cells_per_side for d in range(3)], axis=1) * side + side / 2.0
    positions = np.empty((n_chains, n_particles, 3))
    for c in range(n_chains):
        rng = np.random.Generator(np.random.PCG64(1000 + first_chain + c))
        positions[c] = centres + rng.uniform(-jitter, jitter, size=(n_particles, 3))
    return positions


def coulomb_atoms(n_particles=64, cells_per_side=None, beta=2.0, alpha=3.45, fourier_cutoff=6, position_cutoff=2,
                  bounding_prefactor=1.5837, points_per_side=10, chain_time=0.78965, max_surplus=None, seed=SEED,
                  device=0):
    """(ProgramBuilder, system_length) of C3: the structure of coulomb_atoms/cell_veto.ini at N particles, L = 1."""
    length = 1.0
    if cells_per_side is None:
        cells_per_side = int(np.ceil((2 * n_particles) ** (1.0 / 3.0)))
    mic = abi.EcmcPotential.make(abi.POT_MERGED_IMAGE_COULOMB, 1.0, alpha, fourier_cutoff, position_cutoff)
    bound = abi.EcmcPotential.make(abi.POT_INVERSE_POWER_COULOMB_BOUNDING, bounding_prefactor)
    builder = ProgramBuilder(3, n_particles, length, beta, [cells_per_side] * 3, 1, max_occupants=1,
                             max_surplus=n_particles if max_surplus is None else max_surplus, chain_time=chain_time,
                             speed=1.0, initial_direction=0, initial_active=0, seed=seed)
    builder.set_pair(abi.PAIR_TWO_LEAF_UNIT_BOUNDING, mic, bound, use_charge=True)
    geometry = tables.CellGeometry(3, length, [cells_per_side] * 3, 1)
    bounds, far = tables.inner_point_derivative_bounds(mic, geometry, prefactor=1.0, points_per_side=points_per_side,
                                                       charges=(1.0, 1.0), device=device)
    builder.set_veto(mic, tables.veto_tables(bounds, far), use_charge=True, target_charge=1.0)
    return builder, length


def uniform_start(n_chains, n_particles, length, first_chain=0):
    positions = np.empty((n_chains, n_particles, 3))
    for c in range(n_chains):
        rng = np.random.Generator(np.random.PCG64(1000 + first_chain + c))
        positions[c] = rng.uniform(0.0, length, size=(n_particles, 3))
    return positions
