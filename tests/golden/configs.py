"""INI texts (reference grammar, shipped classes only) of the synthetic configurations of SURVEY.md section 8(d),
and the matching start configurations. Shared by the golden generator, the tests and bench.py."""
import numpy as np

_INTERACTION_TAGS_LJ = "lj_nearby, lj_surplus, lj_cell_veto, cell_boundary"


def lennard_jones_ini(n, system_length, cells_per_side, neighbor_layers=1, beta=1.0, prefactor=4.0,
                      characteristic_length=1.0, estimator_prefactor=1.5, points_per_side=4, chain_time=10.0,
                      end_of_run_time=1.0e9, sampling_interval=None, nearby_handlers=27, surplus_handlers=16,
                      empirical_bound=None):
    """C2 / C5 of SURVEY.md 8(d): 3D Lennard-Jones atoms, LJ inverted exactly for nearby cells and the surplus,
    cell-veto (InnerPointEstimator) for all other cells."""
    tags = _INTERACTION_TAGS_LJ
    sampling = ""
    sampling_tag = ""
    if sampling_interval is not None:
        sampling_tag = ", sampling"
        sampling = f"""
[Sampling]
create = sampling
trash = sampling
event_handler = fixed_interval_sampling_event_handler

[FixedIntervalSamplingEventHandler]
sampling_interval = {sampling_interval!r}
output_handler = separation_output_handler
"""
    output_handlers = "output_handlers = separation_output_handler" if sampling_interval is not None else ""
    eb = "" if empirical_bound is None else f"empirical_bound = {empirical_bound!r}\n"
    return f"""
[Run]
mediator = single_process_mediator
setting = hypercubic_setting

[HypercubicSetting]
system_length = {system_length!r}
beta = {beta!r}
dimension = 3

[SingleProcessMediator]
state_handler = tree_state_handler
scheduler = heap_scheduler
activator = tag_activator
input_output_handler = input_output_handler

[TagActivator]
taggers =
    lj_cell_veto (cell_veto_tagger),
    lj_nearby (excluded_cells_tagger),
    cell_boundary (cell_boundary_tagger),
    lj_surplus (surplus_cells_tagger),
    {"sampling (no_in_state_tagger)," if sampling_interval is not None else ""}
    end_of_chain (active_global_state_in_state_tagger),
    end_of_run (no_in_state_tagger),
    start_of_run (no_in_state_tagger)
internal_states = single_active_cell_occupancy

[LjCellVeto]
create = {tags}
trash = {tags}
event_handler = leaf_unit_cell_veto_event_handler
internal_state_label = single_active_cell_occupancy

[LeafUnitCellVetoEventHandler]
estimator = inner_point_estimator

[InnerPointEstimator]
potential = lennard_jones_potential
prefactor = {estimator_prefactor!r}
points_per_side = {points_per_side}
{eb}
[LennardJonesPotential]
prefactor = {prefactor!r}
characteristic_length = {characteristic_length!r}

[LjNearby]
create = {tags}
trash = {tags}
internal_state_label = single_active_cell_occupancy
event_handler = two_leaf_unit_event_handler
number_event_handlers = {nearby_handlers}

[TwoLeafUnitEventHandler]
potential = lennard_jones_potential

[LjSurplus]
create = {tags}
trash = {tags}
internal_state_label = single_active_cell_occupancy
event_handler = two_leaf_unit_event_handler
number_event_handlers = {surplus_handlers}

[CellBoundary]
create = {tags}
trash = {tags}
internal_state_label = single_active_cell_occupancy
event_handler = cell_boundary_event_handler

[SingleActiveCellOccupancy]
cells = cuboid_periodic_cells
cell_level = 1

[CuboidPeriodicCells]
cells_per_side = {cells_per_side}
neighbor_layers = {neighbor_layers}
{sampling}
[EndOfChain]
create = end_of_chain, {tags}
trash = end_of_chain, {tags}
event_handler = single_independent_active_periodic_direction_end_of_chain_event_handler

[SingleIndependentActivePeriodicDirectionEndOfChainEventHandler]
chain_time = {chain_time!r}

[EndOfRun]
create = end_of_run
trash = end_of_chain, {tags}{sampling_tag}, end_of_run
event_handler = final_time_end_of_run_event_handler

[FinalTimeEndOfRunEventHandler]
end_of_run_time = {end_of_run_time!r}

[StartOfRun]
trash = start_of_run
create = end_of_chain, {tags}{sampling_tag}, end_of_run
event_handler = initial_chain_start_of_run_event_handler

[InitialChainStartOfRunEventHandler]
initial_direction_of_motion = 0
speed = 1.0
initial_active_identifier = 0

[TreeStateHandler]
physical_state = tree_physical_state
lifting_state = tree_lifting_state

[InputOutputHandler]
{output_handlers}
input_handler = random_input_handler

[RandomInputHandler]
random_node_creator = atom_random_node_creator
number_of_root_nodes = {n}

[AtomRandomNodeCreator]
{"" if sampling_interval is None else """
[SeparationOutputHandler]
filename = /tmp/jf_b200_golden_separation.dat
"""}"""


_INTERACTION_TAGS_COULOMB = "coulomb_nearby, coulomb_cell_veto, cell_boundary, coulomb_surplus"


def coulomb_atoms_ini(n, cells_per_side, system_length=1.0, beta=2.0, alpha=3.45, fourier_cutoff=6,
                      position_cutoff=2, prefactor=1.0, bounding_prefactor=1.5837, estimator_prefactor=1.0,
                      points_per_side=10, chain_time=0.78965, end_of_run_time=1.0e9, charge_values="1",
                      nearby_handlers=27, surplus_handlers=16, neighbor_layers=1, far_field="cell_veto"):
    """C3 of SURVEY.md 8(d): the structure of config_files/2018_JCP_149_064113/coulomb_atoms/cell_veto.ini, or with
    far_field="cell_bounding" of coulomb_atoms/cell_bounded.ini (TwoLeafUnitCellBoundingPotentialEventHandler)."""
    tags = _INTERACTION_TAGS_COULOMB
    cps = ", ".join(str(c) for c in cells_per_side)
    text = _coulomb_atoms_text(n, cps, system_length, beta, alpha, fourier_cutoff, position_cutoff, prefactor,
                               bounding_prefactor, estimator_prefactor, points_per_side, chain_time, end_of_run_time,
                               charge_values, nearby_handlers, surplus_handlers, neighbor_layers, tags)
    if far_field == "cell_bounding":
        text = text.replace("coulomb_cell_veto (cell_veto_tagger)", "coulomb_cell_veto (cell_bounding_potential_tagger)")
        text = text.replace("event_handler = leaf_unit_cell_veto_event_handler",
                            "event_handler = two_leaf_unit_cell_bounding_potential_event_handler\n"
                            f"number_event_handlers = {n}")
        text = text.replace("[LeafUnitCellVetoEventHandler]\nestimator = inner_point_estimator\n",
                            "[TwoLeafUnitCellBoundingPotentialEventHandler]\npotential = merged_image_coulomb_potential\n"
                            "bounding_potential = cell_bounding_potential\n")
        text = text.replace("[InnerPointEstimator]", "[CellBoundingPotential]\nestimator = inner_point_estimator\n\n"
                                                     "[InnerPointEstimator]")
    elif far_field != "cell_veto":
        raise ValueError(far_field)
    return text


def _coulomb_atoms_text(n, cps, system_length, beta, alpha, fourier_cutoff, position_cutoff, prefactor,
                        bounding_prefactor, estimator_prefactor, points_per_side, chain_time, end_of_run_time,
                        charge_values, nearby_handlers, surplus_handlers, neighbor_layers, tags):
    return f"""
[Run]
mediator = single_process_mediator
setting = hypercubic_setting

[HypercubicSetting]
system_length = {system_length!r}
beta = {beta!r}
dimension = 3

[SingleProcessMediator]
state_handler = tree_state_handler
scheduler = heap_scheduler
activator = tag_activator
input_output_handler = input_output_handler

[TagActivator]
taggers =
    coulomb_cell_veto (cell_veto_tagger),
    coulomb_nearby (excluded_cells_tagger),
    cell_boundary (cell_boundary_tagger),
    coulomb_surplus (surplus_cells_tagger),
    end_of_chain (active_global_state_in_state_tagger),
    end_of_run (no_in_state_tagger),
    start_of_run (no_in_state_tagger)
internal_states = single_active_cell_occupancy

[CoulombCellVeto]
create = {tags}
trash = {tags}
event_handler = leaf_unit_cell_veto_event_handler
internal_state_label = single_active_cell_occupancy

[LeafUnitCellVetoEventHandler]
estimator = inner_point_estimator
charge = electric_charge

[InnerPointEstimator]
potential = merged_image_coulomb_potential
prefactor = {estimator_prefactor!r}
target_charge = 1.0
points_per_side = {points_per_side}

[MergedImageCoulombPotential]
alpha = {alpha!r}
fourier_cutoff = {fourier_cutoff}
position_cutoff = {position_cutoff}
prefactor = {prefactor!r}

[InversePowerCoulombBoundingPotential]
prefactor = {bounding_prefactor!r}

[CoulombNearby]
create = {tags}
trash = {tags}
internal_state_label = single_active_cell_occupancy
event_handler = two_leaf_unit_bounding_potential_event_handler
number_event_handlers = {nearby_handlers}

[TwoLeafUnitBoundingPotentialEventHandler]
potential = merged_image_coulomb_potential
bounding_potential = inverse_power_coulomb_bounding_potential
charge = electric_charge

[CoulombSurplus]
create = {tags}
trash = {tags}
internal_state_label = single_active_cell_occupancy
event_handler = two_leaf_unit_bounding_potential_event_handler
number_event_handlers = {surplus_handlers}

[CellBoundary]
create = {tags}
trash = {tags}
internal_state_label = single_active_cell_occupancy
event_handler = cell_boundary_event_handler

[SingleActiveCellOccupancy]
cells = cuboid_periodic_cells
cell_level = 1

[CuboidPeriodicCells]
cells_per_side = {cps}
neighbor_layers = {neighbor_layers}

[EndOfChain]
create = end_of_chain, {tags}
trash = end_of_chain, {tags}
event_handler = single_independent_active_periodic_direction_end_of_chain_event_handler

[SingleIndependentActivePeriodicDirectionEndOfChainEventHandler]
chain_time = {chain_time!r}

[EndOfRun]
create = end_of_run
trash = end_of_chain, {tags}, end_of_run
event_handler = final_time_end_of_run_event_handler

[FinalTimeEndOfRunEventHandler]
end_of_run_time = {end_of_run_time!r}

[StartOfRun]
trash = start_of_run
create = end_of_chain, {tags}, end_of_run
event_handler = initial_chain_start_of_run_event_handler

[InitialChainStartOfRunEventHandler]
initial_direction_of_motion = 0
speed = 1.0
initial_active_identifier = 0

[TreeStateHandler]
physical_state = tree_physical_state
lifting_state = tree_lifting_state

[InputOutputHandler]
input_handler = random_input_handler

[RandomInputHandler]
random_node_creator = atom_random_node_creator
number_of_root_nodes = {n}

[AtomRandomNodeCreator]
charge_values = electric_charge_values (charge_values)

[ElectricChargeValues]
charge_name = electric_charge
charge_values = {charge_values}
"""


def lattice_start(n, system_length, cells_per_side, jitter=0.05, seed=1000):
    """Start configuration of C2/C5: particle i sits in cell i (flat cell order, x fastest) at
    the cell centre plus a uniform jitter in (-jitter, jitter)^3, numpy PCG64(seed)."""
    rng = np.random.Generator(np.random.PCG64(seed))
    h = system_length / cells_per_side
    idx = np.arange(n)
    ids = np.stack([idx % cells_per_side, (idx // cells_per_side) % cells_per_side,
                    idx // (cells_per_side * cells_per_side)], axis=1)
    assert ids[:, 2].max() < cells_per_side
    return (ids + 0.5) * h + rng.uniform(-jitter, jitter, size=(n, 3))


def uniform_start(n, system_length, dimension=3, seed=1000):
    rng = np.random.Generator(np.random.PCG64(seed))
    return rng.uniform(0.0, system_length, size=(n, dimension))


# ------------------------------------------------------------------------------------------------------
# C1 of SURVEY.md 8(d): the shipped hard-disk dipole configuration, unchanged except for where the start
# configuration comes from and where the output goes
# ------------------------------------------------------------------------------------------------------
def shipped_ini(ref_root, *relative):
    import os
    with open(os.path.join(ref_root, "jellyfysh", "config_files", *relative)) as file:
        return file.read()


def hard_disk_dipoles_cells_ini(ref_root, n_dipoles=81, end_of_run_time=1.0e9, sampling=False, output="/dev/null"):
    import os
    """config_files/hard_disk_dipoles/hard_disk_dipoles_cells.ini with the PDB input handler (MDAnalysis is not
    installed here) replaced by the random input handler, whose node creator the recorder feeds with the PDB's
    coordinates (ReferenceRun(composites=...)); everything that touches the hot path is the shipped text."""
    text = shipped_ini(ref_root, "hard_disk_dipoles", "hard_disk_dipoles_cells.ini")
    head, tail = text.split("[PdbInputHandler]")
    tail = tail.split("[ElectricChargeValues]", 1)[1]
    text = (head.replace("input_handler = pdb_input_handler", "input_handler = random_input_handler") +
            "[RandomInputHandler]\nrandom_node_creator = dipole_random_node_creator\n"
            f"number_of_root_nodes = {n_dipoles}\n\n"
            "[DipoleRandomNodeCreator]\ncharge_values = electric_charge_values (charge_values)\n"
            "min_initial_dipole_separation = 0.96\nmax_initial_dipole_separation = 1.04\n\n"
            "[ElectricChargeValues]" + tail)
    text = text.replace("filename = config_files/", "filename = " + os.path.join(ref_root, "jellyfysh", "config_files") + "/")
    text = text.replace("end_of_run_time = 15015000", f"end_of_run_time = {end_of_run_time!r}")
    text = text.replace("output/hard_disk_dipoles/Polarization_81Dipoles_Cells.dat", output)
    if not sampling:
        text = text.replace("    polarization_sampling (no_in_state_tagger),\n", "")
        text = text.replace(", polarization_sampling", "")
        start, rest = text.split("[PolarizationSampling]")
        rest = rest.split("[EndOfChain]", 1)[1]
        text = start + "[EndOfChain]" + rest
        text = text.replace("output_handlers = polarization_output_handler\n", "")
        start, rest = text.split("[PolarizationOutputHandler]")
        text = start
    return text


def hard_disk_dipoles_ini(ref_root, end_of_run_time=1.0e9, sampling=False, output="/dev/null", chain_time=None):
    """config_files/hard_disk_dipoles/hard_disk_dipoles.ini (no cell system: every other disk is a candidate of every
    event; the sequential-direction end-of-chain handler rotates the velocity by 20 degrees per chain) with its shipped
    PDB start configuration -- read by the reference's PdbInputHandler through MDAnalysis or jellyfysh_b200's stand-in --,
    absolute file names, the run length and optionally without the sampling handler."""
    import os
    text = shipped_ini(ref_root, "hard_disk_dipoles", "hard_disk_dipoles.ini")
    text = text.replace("filename = config_files/", "filename = " + os.path.join(ref_root, "jellyfysh", "config_files") + "/")
    text = text.replace("end_of_run_time = 15015000", f"end_of_run_time = {end_of_run_time!r}")
    text = text.replace("output/hard_disk_dipoles/Polarization_81Dipoles.dat", output)
    if chain_time is not None:
        text = text.replace("chain_time = 6.0", f"chain_time = {chain_time!r}")
    assert "input_handler = pdb_input_handler" in text and str(end_of_run_time) in text
    if not sampling:
        text = text.replace("    polarization_sampling (no_in_state_tagger),\n", "")
        text = text.replace(", polarization_sampling", "")
        text = text.replace("polarization_sampling, ", "")
        start, rest = text.split("[PolarizationSampling]")
        rest = rest.split("[EndOfChain]", 1)[1]
        text = start + "[EndOfChain]" + rest
        text = text.replace("output_handlers = polarization_output_handler\n", "")
        text = text.split("[PolarizationOutputHandler]")[0]
    return text


def read_pdb_dipoles(ref_root, system_length=12.836):
    """(roots[81][2], leaves[81][2][2]) of the shipped PDB start configuration as PdbInputHandler.read builds them
    (pdb_input_handler.py:146-190): MDAnalysis keeps coordinates as float32; the root is the barycentre over the
    shortest separations."""
    import os
    import numpy as np
    path = os.path.join(ref_root, "jellyfysh", "config_files", "hard_disk_dipoles",
                        "81dipoles_min0.952380952380952_max1.047619047619048.pdb")
    atoms = []
    for line in open(path):
        if line.startswith("ATOM"):
            atoms.append((int(line[22:26]), float(np.float32(line[30:38])), float(np.float32(line[38:46]))))
    n = max(a[0] for a in atoms)
    leaves = np.zeros((n, 2, 2))
    count = [0] * n
    for resid, x, y in atoms:
        leaves[resid - 1, count[resid - 1]] = [x % system_length, y % system_length]
        count[resid - 1] += 1
    roots = np.zeros((n, 2))
    half = system_length / 2.0
    for r in range(n):
        first, other = leaves[r, 0], leaves[r, 1]
        for d in range(2):
            shortest = (float(other[d]) - float(first[d]) + half) % system_length - half
            closest = float(first[d]) + shortest
            center = float(first[d]) * 0.5 + closest * 0.5
            roots[r, d] = center % system_length
    return roots, leaves


def patch_composite_start(composites):
    """Make the reference's random node creators return the given composite point objects instead of drawing them:
    composites = (roots[n_roots][D], leaves[n_roots][nodes_per_root][D]); successive reads of the input handler cycle
    through roots / leaves in blocks of number_of_root_nodes. Charges are assigned by the creator as usual. Returns a
    function that restores the reference's methods. Needs `jellyfysh` importable."""
    import itertools
    from jellyfysh.base.node import Node
    from jellyfysh.base.particle import Particle
    from jellyfysh.input_output_handler.input_handler.random_node_creator import (dipole_random_node_creator,
                                                                                  water_random_node_creator)
    roots, leaves = composites
    counter = itertools.cycle(range(len(roots)))

    def fill_root_node(creator, node):
        r = next(counter)
        for k, position in enumerate(leaves[r]):
            node.add_child(Node(Particle([float(x) for x in position],
                                         {cv.charge_name: cv[k] for cv in creator._charge_values})))
        node.value = Particle(position=[float(x) for x in roots[r]])

    saved = []
    for cls in (dipole_random_node_creator.DipoleRandomNodeCreator, water_random_node_creator.WaterRandomNodeCreator):
        saved.append((cls, cls.fill_root_node))
        cls.fill_root_node = fill_root_node

    def restore():
        for cls, original in saved:
            cls.fill_root_node = original
    return restore


# ------------------------------------------------------------------------------------------------------
# C4 of SURVEY.md 8(d): SPC/Fw-like water, config_files/2018_JCP_149_064113/water/coulomb_cell_veto_lj_inverted.ini
# ------------------------------------------------------------------------------------------------------
def water_ini(ref_root, n_molecules=32, end_of_run_time=1.0e9, number_trials=1000, sampling_interval=None,
              output="/dev/null", cells_per_side=None, neighbor_layers=None, system_length=None):
    """The shipped water configuration (Coulomb through composite-object handlers with cell veto, Lennard-Jones
    inverted, harmonic bonds, bending with ratio lifting) for n_molecules molecules; only the size of the system,
    the run length, the number of estimator trials and the output change."""
    import os
    text = shipped_ini(ref_root, "2018_JCP_149_064113", "water", "coulomb_cell_veto_lj_inverted.ini")
    text = text.replace("filename = config_files/", "filename = " + os.path.join(ref_root, "jellyfysh", "config_files") + "/")
    text = text.replace("number_of_root_nodes = 2", f"number_of_root_nodes = {n_molecules}")
    # the shipped file is sized for two molecules: one handler per possible partner molecule
    for section in ("[CoulombNearby]", "[CoulombSurplus]", "[LennardJones]"):
        head, tail = text.split(section)
        tail = tail.replace("number_event_handlers = 1", f"number_event_handlers = {max(n_molecules - 1, 1)}", 1)
        text = head + section + tail
    text = text.replace("end_of_run_time = 500000", f"end_of_run_time = {end_of_run_time!r}")
    text = text.replace("number_trials = 1000", f"number_trials = {number_trials}")
    text = text.replace("output/2018_JCP_149_064113/water/SamplesOfOOSeparation_CoulombCellVeto_LJInverted.dat", output)
    if cells_per_side is not None:
        text = text.replace("cells_per_side = 6, 6, 6", "cells_per_side = " + ", ".join(str(c) for c in cells_per_side))
    if neighbor_layers is not None:
        text = text.replace("neighbor_layers = 2", f"neighbor_layers = {neighbor_layers}")
    if system_length is not None:
        text = text.replace("system_length = 10", f"system_length = {system_length!r}")
    if sampling_interval is None:
        text = text.replace("    sampling (no_in_state_tagger),\n", "")
        text = text.replace(", sampling", "")
        start, rest = text.split("[Sampling]")
        rest = rest.split("[EndOfChain]", 1)[1]
        text = start + "[EndOfChain]" + rest
        text = text.replace("output_handlers = oxygen_oxygen_separation_output_handler\n", "")
        text = text.split("[OxygenOxygenSeparationOutputHandler]")[0]
    else:
        text = text.replace("sampling_interval = 2.6789", f"sampling_interval = {sampling_interval!r}")
    return text


def water_start(n_molecules, system_length, seed=1000, bond_length=1.012, bond_angle=1.9764, jitter=0.3):
    """(roots[n][3], leaves[n][3][3]) of water molecules (H, O, H; the root is the geometric centre, as
    WaterRandomNodeCreator builds them, water_random_node_creator.py:69-118) on a jittered cubic lattice with random
    orientations, so that no two molecules start on top of each other."""
    import numpy as np
    rng = np.random.default_rng(seed)
    side = int(np.ceil(n_molecules ** (1.0 / 3.0)))
    grid = np.stack(np.meshgrid(*[np.arange(side)] * 3, indexing="ij"), axis=-1).reshape(-1, 3)
    grid = grid[rng.permutation(len(grid))[:n_molecules]]
    roots = np.empty((n_molecules, 3))
    leaves = np.empty((n_molecules, 3, 3))
    for m in range(n_molecules):
        centre = (grid[m] + 0.5) * (system_length / side) + rng.uniform(-jitter, jitter, size=3)
        axis = rng.normal(size=3)
        axis /= np.linalg.norm(axis)
        other = rng.normal(size=3)
        other -= other.dot(axis) * axis
        other /= np.linalg.norm(other)
        oh_one = bond_length * (np.cos(bond_angle / 2) * axis + np.sin(bond_angle / 2) * other)
        oh_two = bond_length * (np.cos(bond_angle / 2) * axis - np.sin(bond_angle / 2) * other)
        oxygen = centre - (oh_one + oh_two) / 3.0
        roots[m] = centre % system_length
        leaves[m, 0] = (oxygen + oh_one) % system_length
        leaves[m, 1] = oxygen % system_length
        leaves[m, 2] = (oxygen + oh_two) % system_length
    return roots, leaves


def coulomb_power_bounded_ini(ref_root, n_atoms=2, end_of_run_time=1.0e9, sampling=False, output="/dev/null"):
    """The shipped coulomb_atoms/power_bounded.ini -- no cell system: the Coulomb pair factors of the active atom with
    every other atom come from a FactorTypeMapInStateTagger ("[0, 1], Coulomb") and are bounded by the inverse-power
    Coulomb bounding potential -- for n_atoms atoms; only the size of the system, the run length and the output change."""
    import os
    text = shipped_ini(ref_root, "2018_JCP_149_064113", "coulomb_atoms", "power_bounded.ini")
    text = text.replace("filename = config_files/", "filename = " + os.path.join(ref_root, "jellyfysh", "config_files") + "/")
    text = text.replace("number_of_root_nodes = 2", f"number_of_root_nodes = {n_atoms}")
    # the shipped file is sized for two atoms: one handler per possible partner
    text = text.replace("number_event_handlers = 1", f"number_event_handlers = {max(n_atoms - 1, 1)}", 1)
    text = text.replace("end_of_run_time = 100000", f"end_of_run_time = {end_of_run_time!r}")
    text = text.replace("output/2018_JCP_149_064113/coulomb_atoms/SamplesOfSeparation_PowerBounded.dat", output)
    if not sampling:
        text = text.replace("    sampling (no_in_state_tagger),\n", "")
        text = text.replace(", sampling", "")
        start, rest = text.split("[Sampling]")
        rest = rest.split("[EndOfChain]", 1)[1]
        text = start + "[EndOfChain]" + rest
        text = text.replace("output_handlers = separation_output_handler\n", "")
        text = text.split("[SeparationOutputHandler]")[0]
    return text


def shipped_without_sampling(ref_root, relative, end_of_run_time=1.0e9, replacements=()):
    """A shipped configuration file, unchanged in everything the hot path reads, with its sampling tagger, output
    handlers and output files removed and the run length changed: for event-by-event recordings.
    `relative`: path components below config_files; `replacements`: (old, new) text pairs applied first."""
    import configparser
    import io
    import os
    text = shipped_ini(ref_root, *relative)
    text = text.replace("filename = config_files/", "filename = " + os.path.join(ref_root, "jellyfysh", "config_files") + "/")
    for old, new in replacements:
        assert old in text, old
        text = text.replace(old, new)
    parser = configparser.ConfigParser()
    parser.optionxform = str
    parser.read_string(text)
    taggers = [line.strip().rstrip(",") for line in parser.get("TagActivator", "taggers").strip().splitlines()]
    sampling = [t.split()[0] for t in taggers if "sampling" in t.split()[0]]
    parser.set("TagActivator", "taggers", "\n" + ",\n".join(t for t in taggers if t.split()[0] not in sampling))
    camel = lambda name: "".join(part.capitalize() for part in name.split("_"))
    for tag in sampling:
        handler = parser.get(camel(tag), "event_handler").split()[0]
        parser.remove_section(camel(tag))
        parser.remove_section(camel(handler))
    for section in parser.sections():
        for key in ("create", "trash", "activate", "deactivate"):
            if parser.has_option(section, key):
                kept = [item.strip() for item in parser.get(section, key).split(",") if item.strip() not in sampling]
                parser.set(section, key, ", ".join(kept))
    for handler in [h.strip().split()[0] for h in parser.get("InputOutputHandler", "output_handlers").split(",")]:
        parser.remove_section(camel(handler))
    parser.remove_option("InputOutputHandler", "output_handlers")
    parser.set("FinalTimeEndOfRunEventHandler", "end_of_run_time", repr(end_of_run_time))
    out = io.StringIO()
    parser.write(out)
    return out.getvalue()


def dipole_start(n_dipoles, length=1.0, seed=900, minimum_distance=0.3, bond=0.1):
    """n_dipoles dipoles (charges +1, -1 at separation `bond`) with centres at least minimum_distance apart."""
    import numpy as np
    rng = np.random.default_rng(seed)
    centres = []
    while len(centres) < n_dipoles:
        c = rng.uniform(0.0, length, size=3)
        if all(np.linalg.norm(np.mod(c - o + length / 2, length) - length / 2) > minimum_distance for o in centres):
            centres.append(c)
    roots = np.empty((n_dipoles, 3))
    leaves = np.empty((n_dipoles, 2, 3))
    for k, c in enumerate(centres):
        axis = rng.normal(size=3)
        axis *= 0.5 * bond / np.linalg.norm(axis)
        roots[k] = c % length
        leaves[k, 0] = (c + axis) % length
        leaves[k, 1] = (c - axis) % length
    return roots, leaves
