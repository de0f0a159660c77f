"""Build programs / oracle chains from the committed reference traces (tests/golden/trace_*.npz)."""
import numpy as np

import kat_replay as kr
from jellyfysh_b200 import abi

TRACES = ["trace_lj_small", "trace_lj_surplus", "trace_coulomb_small", "trace_coulomb_surplus"]
CELL_BOUNDING_TRACES = ["trace_coulomb_cell_bounded"]
NO_CELL_TRACES = ["trace_coulomb_power_bounded"]
DIPOLE_TRACES = ["trace_hard_disk_dipoles"]
WATER_TRACES = ["trace_water", "trace_water_dense"]
DISCRETE_FIELDS = ("kind", "target", "target_cell", "accepted", "n_candidates", "new_active", "new_direction")


def reference_tables(g, dimension=3):
    """Walker tables and bounds exactly as the reference built them (stored in the trace fixture)."""
    tables = {"upper": [], "lower": [], "bounds": g["bounds"]}
    for name in ("upper", "lower"):
        for d in range(dimension):
            tables[name].append({"cell_a": g[f"{name}{d}_cell_a"], "cell_b": g[f"{name}{d}_cell_b"],
                                 "rate_a": g[f"{name}{d}_rate_a"], "total_rate": float(g[f"{name}{d}_rates"][0]),
                                 "mean_rate": float(g[f"{name}{d}_rates"][1])})
    return tables


def potentials_of(g):
    """(pair_handler, pair potential, bounding potential or None, veto potential, use_charge)."""
    if "meta_lj" in g:
        lj = abi.EcmcPotential.make(abi.POT_LENNARD_JONES, *g["meta_lj"])
        return abi.PAIR_TWO_LEAF_UNIT, lj, None, lj, False
    mic = abi.EcmcPotential.make(abi.POT_MERGED_IMAGE_COULOMB, *g["meta_mic"])
    ipcb = abi.EcmcPotential.make(abi.POT_INVERSE_POWER_COULOMB_BOUNDING, *g["meta_ipcb"])
    return abi.PAIR_TWO_LEAF_UNIT_BOUNDING, mic, ipcb, mic, True


def builder_of(g, builder_cls, tables=None, max_surplus=128):
    handler, pot, bound, veto, use_charge = potentials_of(g)
    cps = [int(c) for c in g["meta_cells_per_side"]]
    pb = builder_cls(3, int(g["meta_n"]), float(g["meta_system_length"]), float(g["meta_beta"]), cps, 1,
                     max_occupants=1, max_surplus=max_surplus, chain_time=float(g["meta_chain_time"]),
                     seed=int(g["seed"][0]))
    pb.set_pair(handler, pot, bound, use_charge=use_charge)
    if "meta_far_field" in g and int(g["meta_far_field"]) == abi.FAR_CELL_BOUNDING:
        pb.set_cell_bounding(veto, g["bounds"] if tables is None else tables["bounds"], use_charge=use_charge,
                             target_charge=1.0)
    else:
        pb.set_veto(veto, tables if tables is not None else reference_tables(g), use_charge=use_charge,
                    target_charge=1.0)
    return pb


def no_cells_builder_of(g, builder_cls):
    """The shipped coulomb_atoms/power_bounded.ini shape: no cell system, every other atom is a candidate of every event
    (merged-image Coulomb bounded by the inverse-power Coulomb bounding potential)."""
    handler, pot, bound, _, use_charge = potentials_of(g)
    n = int(g["meta_n"])
    pb = builder_cls(3, n, float(g["meta_system_length"]), float(g["meta_beta"]), [1, 1, 1], 0, max_occupants=1,
                     max_surplus=n, chain_time=float(g["meta_chain_time"]), seed=int(g["seed"][0]), no_cells=True)
    pb.set_pair(handler, pot, bound, use_charge=use_charge)
    return pb


def dipole_factors_builder_of(g, builder_cls):
    """The shipped dipoles/dipole_factors_*.ini: composite-object Coulomb factor with a lifting scheme, harmonic bond,
    1/r^6 repulsion between unlike charges of different dipoles; no cell system."""
    pb = builder_cls(3, int(g["meta_n"]), float(g["meta_system_length"]), float(g["meta_beta"]), [1, 1, 1], 0,
                     chain_time=float(g["meta_chain_time"]), initial_active=int(g["meta_initial_active"]),
                     seed=int(g["seed"][0]), no_cells=True)
    # atom_factors.ini: the Coulomb interaction as bounded leaf-to-leaf factors between the objects
    handler = abi.PAIR_TWO_LEAF_UNIT_BOUNDING if "meta_leaf_pairs" in g else abi.PAIR_TWO_COMPOSITE_SUMMED_BOUNDING
    pb.set_pair(handler, abi.EcmcPotential.make(abi.POT_MERGED_IMAGE_COULOMB, *g["meta_mic"]),
                abi.EcmcPotential.make(abi.POT_INVERSE_POWER_COULOMB_BOUNDING, *g["meta_ipcb"]), use_charge=True)
    pb.set_composite(int(g["meta_nodes_per_root"]), bonds=[(0, 1)],
                     bond_potential=abi.EcmcPotential.make(abi.POT_DISPLACED_EVEN_POWER, *g["meta_harmonic"]))
    lifting = (abi.LIFTING_INSIDE_FIRST, abi.LIFTING_OUTSIDE_FIRST, abi.LIFTING_RATIO)[int(g["meta_lifting"])]
    pb.set_molecules(lifting, inter_factors=[(0, 1), (1, 0)],
                     inter_potential=abi.EcmcPotential.make(abi.POT_INVERSE_POWER, *g["meta_repulsive"]))
    return pb


def dipole_motion_builder_of(g, builder_cls):
    """The shipped dipoles/dipole_motion.ini: the factors of dipole_factors_inside_first.ini while a leaf unit is active,
    the root-unit-active handlers with the same potentials while a root unit is, RootLeafUnitActiveSwitcher in between."""
    pb = dipole_factors_builder_of(g, builder_cls)
    pb.set_root_mode(*[float(x) for x in g["meta_switch_chain_length"]])
    return pb


def dipole_cell_bounded_builder_of(g, builder_cls):
    """The shipped dipoles/cell_bounded.ini: composite-object Coulomb handlers for nearby cells and the surplus, the
    composite-object cell-bounding handler for all other occupied cells (bounds as the reference's estimator built
    them), harmonic bond, 1/r^6 repulsion between objects; root-level cells."""
    n = int(g["meta_n"])
    npr = int(g["meta_nodes_per_root"])
    pb = builder_cls(3, n, float(g["meta_system_length"]), float(g["meta_beta"]),
                     [int(c) for c in g["meta_cells_per_side"]], int(g["meta_neighbor_layers"]), max_occupants=1,
                     max_surplus=n // npr, chain_time=float(g["meta_chain_time"]),
                     initial_active=int(g["meta_initial_active"]), seed=int(g["seed"][0]))
    mic = abi.EcmcPotential.make(abi.POT_MERGED_IMAGE_COULOMB, *g["meta_mic"])
    pb.set_pair(abi.PAIR_TWO_COMPOSITE_SUMMED_BOUNDING, mic,
                abi.EcmcPotential.make(abi.POT_INVERSE_POWER_COULOMB_BOUNDING, *g["meta_ipcb"]), use_charge=True)
    pb.set_cell_bounding(mic, g["bounds"], use_charge=True, target_charge=1.0)
    pb.set_composite(npr, bonds=[(0, 1)],
                     bond_potential=abi.EcmcPotential.make(abi.POT_DISPLACED_EVEN_POWER, *g["meta_harmonic"]))
    pb.set_molecules(abi.LIFTING_INSIDE_FIRST, inter_factors=[(0, 1), (1, 0)],
                     inter_potential=abi.EcmcPotential.make(abi.POT_INVERSE_POWER, *g["meta_repulsive"]))
    return pb


def single_molecule_builder_of(g, builder_cls):
    """The shipped water/single_molecule.ini: two harmonic bonds and the bending factor of one molecule."""
    pb = builder_cls(3, int(g["meta_n"]), float(g["meta_system_length"]), float(g["meta_beta"]), [1, 1, 1], 0,
                     chain_time=float(g["meta_chain_time"]), initial_active=int(g["meta_initial_active"]),
                     seed=int(g["seed"][0]), no_cells=True)
    pb.set_composite(int(g["meta_nodes_per_root"]), bonds=[(0, 1), (1, 2)],
                     bond_potential=abi.EcmcPotential.make(abi.POT_DISPLACED_EVEN_POWER, *g["meta_harmonic"]))
    pb.set_molecules(abi.LIFTING_INSIDE_FIRST,
                     bending=dict(children=[0, 1, 2], separations=[1, 0, 1, 2], lifting=abi.LIFTING_RATIO,
                                  potential=abi.EcmcPotential.make(abi.POT_BENDING, *g["meta_bending"]),
                                  offset=float(g["meta_bending_offset"]),
                                  max_displacement=float(g["meta_bending_max_displacement"])))
    return pb


def water_atomic_builder_of(g, builder_cls):
    """The shipped water/coulomb_power_bounded_lj_inverted.ini: no cell system, Coulomb as bounded leaf-to-leaf factors
    between the molecules, Lennard-Jones between the oxygens, harmonic bonds, bending."""
    pb = single_molecule_builder_of(g, builder_cls)
    pb.set_pair(abi.PAIR_TWO_LEAF_UNIT_BOUNDING, abi.EcmcPotential.make(abi.POT_MERGED_IMAGE_COULOMB, *g["meta_mic"]),
                abi.EcmcPotential.make(abi.POT_INVERSE_POWER_COULOMB_BOUNDING, *g["meta_ipcb"]), use_charge=True)
    pb.set_molecules(abi.LIFTING_INSIDE_FIRST, inter_factors=[(1, 1)],
                     inter_potential=abi.EcmcPotential.make(abi.POT_LENNARD_JONES, *g["meta_lj"]),
                     bending=dict(children=[0, 1, 2], separations=[1, 0, 1, 2], lifting=abi.LIFTING_RATIO,
                                  potential=abi.EcmcPotential.make(abi.POT_BENDING, *g["meta_bending"]),
                                  offset=float(g["meta_bending_offset"]),
                                  max_displacement=float(g["meta_bending_max_displacement"])))
    return pb


COMPOSITE_CELL_BOUNDING_TRACES = ["trace_dipole_cell_bounded"]
NO_CELL_MOLECULE_TRACES = {"trace_dipole_factors_inside_first": dipole_factors_builder_of,
                           "trace_dipole_factors_outside_first": dipole_factors_builder_of,
                           "trace_dipole_factors_ratio": dipole_factors_builder_of,
                           "trace_dipole_atom_factors": dipole_factors_builder_of,
                           "trace_dipole_motion": dipole_motion_builder_of,
                           "trace_water_atomic_factors": water_atomic_builder_of,
                           "trace_water_single_molecule": single_molecule_builder_of}


def dipole_builder_of(g, builder_cls):
    """C1: hard-disk dipoles (two disks per root, leaf-level cells, hard-sphere pairs, hard-dipole tether)."""
    dimension = len(g["meta_cells_per_side"])
    pb = builder_cls(dimension, int(g["meta_n"]), float(g["meta_system_length"]), float(g["meta_beta"]),
                     [int(c) for c in g["meta_cells_per_side"]], 1, max_occupants=int(g["meta_max_occupants"]),
                     max_surplus=0, chain_time=float(g["meta_chain_time"]), seed=int(g["seed"][0]))
    pb.set_pair(abi.PAIR_TWO_LEAF_UNIT, abi.EcmcPotential.make(abi.POT_HARD_SPHERE, *g["meta_hard_sphere"]))
    pb.set_composite(int(g["meta_nodes_per_root"]), bonds=[(0, 1)],
                     bond_potential=abi.EcmcPotential.make(abi.POT_HARD_DIPOLE, *g["meta_hard_dipole"]))
    return pb


def sequential_dipole_builder_of(g, builder_cls):
    """The shipped hard_disk_dipoles.ini: hard-disk dipoles without a cell system, general velocities (the
    sequential-direction end-of-chain handler rotates the velocity by delta_phi per chain)."""
    pb = builder_cls(2, int(g["meta_n"]), float(g["meta_system_length"]), float(g["meta_beta"]), [1, 1], 0,
                     chain_time=float(g["meta_chain_time"]), seed=int(g["seed"][0]), no_cells=True)
    pb.set_composite(int(g["meta_nodes_per_root"]), bonds=[(0, 1)],
                     bond_potential=abi.EcmcPotential.make(abi.POT_HARD_DIPOLE, *g["meta_hard_dipole"]))
    pb.set_sequential_direction(float(g["meta_delta_phi_degree"]),
                                abi.EcmcPotential.make(abi.POT_HARD_SPHERE, *g["meta_hard_sphere"]),
                                [(0, 0), (0, 1), (1, 0), (1, 1)])
    return pb


def water_builder_of(g, builder_cls, max_surplus=None):
    """C4: SPC/Fw-like water (water/coulomb_cell_veto_lj_inverted.ini): composite-object Coulomb handlers with
    inside-first lifting on root-level cells, Lennard-Jones between the oxygens, harmonic bonds, bending."""
    n = int(g["meta_n"])
    pb = builder_cls(3, n, float(g["meta_system_length"]), float(g["meta_beta"]),
                     [int(c) for c in g["meta_cells_per_side"]], int(g["meta_neighbor_layers"]), max_occupants=1,
                     max_surplus=n // 3 if max_surplus is None else max_surplus, chain_time=float(g["meta_chain_time"]),
                     initial_active=int(g["meta_initial_active"]), seed=int(g["seed"][0]))
    mic = abi.EcmcPotential.make(abi.POT_MERGED_IMAGE_COULOMB, *g["meta_mic"])
    pb.set_pair(abi.PAIR_TWO_COMPOSITE_SUMMED_BOUNDING, mic,
                abi.EcmcPotential.make(abi.POT_INVERSE_POWER_COULOMB_BOUNDING, *g["meta_ipcb"]), use_charge=True)
    pb.set_veto(mic, reference_tables(g), use_charge=True, target_charge=1.0)
    pb.set_composite(int(g["meta_nodes_per_root"]), bonds=[(0, 1), (1, 2)],
                     bond_potential=abi.EcmcPotential.make(abi.POT_DISPLACED_EVEN_POWER, *g["meta_harmonic"]))
    pb.set_molecules(abi.LIFTING_INSIDE_FIRST, inter_factors=[(1, 1)],
                     inter_potential=abi.EcmcPotential.make(abi.POT_LENNARD_JONES, *g["meta_lj"]),
                     bending=dict(children=[0, 1, 2], separations=[1, 0, 1, 2], lifting=abi.LIFTING_RATIO,
                                  potential=abi.EcmcPotential.make(abi.POT_BENDING, *g["meta_bending"]),
                                  offset=float(g["meta_bending_offset"]),
                                  max_displacement=float(g["meta_bending_max_displacement"])))
    return pb


LEAF_CELL_WATER_TRACES = ["trace_water_lj_cell_bounded", "trace_water_lj_cell_bounded_dense",
                          "trace_water_lj_cell_bounded_surplus"]


def leaf_cell_water_builder_of(g, builder_cls):
    """The shipped water/coulomb_power_bounded_lj_cell_bounded.ini: composite-object Coulomb factors, bonds and bending from
    the factor type map; a cell system for the oxygens only, through which the Lennard-Jones factor between oxygens is
    found (piecewise constant bound for nearby cells + surplus, cell-bounding potential elsewhere)."""
    n = int(g["meta_n"])
    pb = builder_cls(3, n, float(g["meta_system_length"]), float(g["meta_beta"]),
                     [int(c) for c in g["meta_cells_per_side"]], int(g["meta_neighbor_layers"]), max_occupants=1,
                     max_surplus=n // 3, chain_time=float(g["meta_chain_time"]),
                     initial_active=int(g["meta_initial_active"]), seed=int(g["seed"][0]))
    pb.set_pair(abi.PAIR_TWO_COMPOSITE_SUMMED_BOUNDING, abi.EcmcPotential.make(abi.POT_MERGED_IMAGE_COULOMB, *g["meta_mic"]),
                abi.EcmcPotential.make(abi.POT_INVERSE_POWER_COULOMB_BOUNDING, *g["meta_ipcb"]), use_charge=True)
    lj = abi.EcmcPotential.make(abi.POT_LENNARD_JONES, *g["meta_lj"])
    pb.set_cell_bounding(lj, g["bounds"], use_charge=False)
    pb.set_composite(int(g["meta_nodes_per_root"]), bonds=[(0, 1), (1, 2)],
                     bond_potential=abi.EcmcPotential.make(abi.POT_DISPLACED_EVEN_POWER, *g["meta_harmonic"]))
    child = int(g["meta_cell_child"])
    pb.set_molecules(abi.LIFTING_INSIDE_FIRST, inter_factors=[(child, child)], inter_potential=lj,
                     bending=dict(children=[0, 1, 2], separations=[1, 0, 1, 2], lifting=abi.LIFTING_RATIO,
                                  potential=abi.EcmcPotential.make(abi.POT_BENDING, *g["meta_bending"]),
                                  offset=float(g["meta_bending_offset"]),
                                  max_displacement=float(g["meta_bending_max_displacement"])))
    pb.set_leaf_cells(child, float(g["meta_lj_offset"]), float(g["meta_lj_max_displacement"]))
    return pb


def charges_of(g):
    return g["charges"] if "charges" in g else None


def load_trace(name):
    return kr.load_npz(name)


def records_equal_discrete(a, b):
    return all(np.array_equal(a[f], b[f]) for f in DISCRETE_FIELDS)


def max_time_error(a, b):
    """Largest |t_a - t_b| / max(1, |t|) with t = quotient + remainder evaluated without cancellation."""
    diff = (a["time_q"] - b["time_q"]) + (a["time_r"] - b["time_r"])
    scale = np.maximum(1.0, np.abs(a["time_q"] + a["time_r"]))
    return float(np.max(np.abs(diff) / scale))
