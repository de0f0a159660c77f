"""Compile a JeLLyFysh object graph (built by the reference's own factory from an INI file) into an EcmcProgram.

The reference decides at run time, event by event, which handlers run (TagActivator + taggers). For the cell-based
configurations of SURVEY.md 8(d) that decision is static: after every event all interaction candidates are trashed and
recreated, so the tag graph is a *description* of the device program:

  ExcludedCellsTagger + SurplusCellsTagger  -> the pair factor (potential [+ bounding potential], charge)
  CellVetoTagger                            -> LeafUnitCellVetoEventHandler: potential, Walker tables, bounds, charge
  CellBoundaryTagger                        -> cell-boundary candidate
  end-of-chain / start-of-run handlers      -> chain_time, speed, initial direction / active particle
  SingleActiveCellOccupancy + CuboidPeriodicCells -> cell grid, neighbour layers, occupants per cell

This module reads those objects (public attributes where the reference has them, otherwise the private ones named
below with file:line) and fills a ProgramBuilder. Anything it does not recognise raises ConfigurationError -- a graph
is never approximated. Needs the `jellyfysh` package importable; nothing here computes on the hot path.
"""
import numpy as np

from jellyfysh_b200 import abi
from jellyfysh_b200.program import ProgramBuilder


def _class_names(obj):
    """Names of all classes in the MRO; factory aliases are dynamic subclasses named 'Alias (RealClass)'
    (jellyfysh/base/factory.py:143-149), so the real class is always among the bases."""
    return {cls.__name__ for cls in type(obj).__mro__}


def _configuration_error(message):
    from jellyfysh.base.exceptions import ConfigurationError
    return ConfigurationError("CudaBatchedMediator: " + message)


class _ChargeProbe(dict):
    """Stands in for Unit.charge to learn which charge name (if any) a handler reads."""

    def __init__(self):
        super().__init__()
        self.names = []

    def __getitem__(self, name):
        self.names.append(name)
        return 1.0


class _ProbeUnit:
    def __init__(self):
        self.charge = _ChargeProbe()


def _charge_name(charges_function):
    """The charge name used by a handler's `_charges(unit_one, unit_two)` lambda
    (two_leaf_unit_event_handler.py:84-94), or None if it passes 1.0 / nothing."""
    one, two = _ProbeUnit(), _ProbeUnit()
    charges_function(one, two)
    names = set(one.charge.names) | set(two.charge.names)
    if len(names) > 1:
        raise _configuration_error("a handler reads more than one charge: {0}".format(sorted(names)))
    return names.pop() if names else None


def _root_with_exact_square(square, scale=1.0):
    """x with scale * x * x == square in floating point (the reference stores only the squares, e.g.
    hard_sphere_potential.py:63 `4.0 * radius * radius`; the device recomputes them the same way from x)."""
    import math
    x = math.sqrt(square / scale)
    for candidate in (x, math.nextafter(x, 0.0), math.nextafter(x, math.inf)):
        if scale * candidate * candidate == square:
            return candidate
    raise _configuration_error("cannot recover a length from its stored square {0!r}".format(square))


def potential_descriptor(potential):
    """EcmcPotential of a reference potential object (jellyfysh/potential/*)."""
    names = _class_names(potential)
    if "LennardJonesPotential" in names:
        # lennard_jones_potential.py:42-61: _prefactor (Potential base), _characteristic_length
        return abi.EcmcPotential.make(abi.POT_LENNARD_JONES, potential._prefactor, potential._characteristic_length)
    if "InversePowerPotential" in names:
        return abi.EcmcPotential.make(abi.POT_INVERSE_POWER, potential._power, potential._prefactor)
    if "DisplacedEvenPowerPotential" in names:
        return abi.EcmcPotential.make(abi.POT_DISPLACED_EVEN_POWER, potential._prefactor,
                                      potential._equilibrium_separation, potential._power)
    if "HardSpherePotential" in names:
        return abi.EcmcPotential.make(abi.POT_HARD_SPHERE, _root_with_exact_square(potential._diameter_squared, 4.0))
    if "HardDipolePotential" in names:
        # hard_dipole_potential.py:72-73
        return abi.EcmcPotential.make(abi.POT_HARD_DIPOLE, _root_with_exact_square(potential._minimum_separation_squared),
                                      _root_with_exact_square(potential._maximum_separation_squared))
    if "MergedImageCoulombPotential" in names:
        # merged_image_coulomb_potential.py:111-118
        return abi.EcmcPotential.make(abi.POT_MERGED_IMAGE_COULOMB, potential._prefactor, potential._alpha,
                                      potential._fourier_cutoff, potential._position_cutoff)
    if "InversePowerCoulombBoundingPotential" in names:
        return abi.EcmcPotential.make(abi.POT_INVERSE_POWER_COULOMB_BOUNDING, potential._prefactor)
    if "BendingPotential" in names:
        return abi.EcmcPotential.make(abi.POT_BENDING, potential._prefactor, potential._equilibrium_angle)
    raise _configuration_error("potential {0} has no device implementation".format(type(potential).__name__))


def _lifting_kind(lifting):
    names = _class_names(lifting)
    for name, kind in (("InsideFirstLifting", abi.LIFTING_INSIDE_FIRST), ("OutsideFirstLifting", abi.LIFTING_OUTSIDE_FIRST),
                       ("RatioLifting", abi.LIFTING_RATIO)):
        if name in names:
            return kind
    raise _configuration_error("lifting scheme {0} has no device implementation".format(type(lifting).__name__))


def _same_potential(a, b):
    return a.kind == b.kind and list(a.params) == list(b.params)


class CompiledProgram:
    """The result: a ProgramBuilder plus what the mediator needs around it."""

    def __init__(self, builder, charge_name, control_handlers, n_particles, nodes_per_root=1):
        self.builder = builder
        self.charge_name = charge_name
        self.control_handlers = control_handlers  # sampling / end-of-run handlers, run on the host
        self.n_particles = n_particles            # leaf units
        self.nodes_per_root = nodes_per_root      # 1: point masses; > 1: composite point objects


def compile_program(activator, extracted_global_state, seed=0, max_surplus=None, occupant_capacity=8):
    """Walk the initialized activator and return a CompiledProgram.

    extracted_global_state: state_handler.extract_global_state() (root cnodes), used for N and the node structure.
    occupant_capacity: occupants stored per cell on the device when the reference's occupancy is unbounded
    (maximum_number_occupants = -1); an overflow is reported as a capacity error, never silently dropped."""
    import jellyfysh.setting as setting
    from jellyfysh.setting import hypercubic_setting

    levels = setting.number_of_node_levels
    if levels not in (1, 2):
        raise _configuration_error("only trees with one or two node levels are supported")
    nodes_per_root = setting.number_of_nodes_per_root_node if levels == 2 else 1
    for cnode in extracted_global_state:
        if len(cnode.children) != (nodes_per_root if levels == 2 else 0):
            raise _configuration_error("every root node must have the same number of leaf nodes")
    if not hypercubic_setting.initialized():
        raise _configuration_error("a hypercubic setting is required")
    dimension, length = setting.dimension, hypercubic_setting.system_length
    n_particles = len(extracted_global_state) * nodes_per_root

    # ---- cell occupancy (tag_activator.py:82-135 keeps the internal states in _internal_states)
    internal_states = list(activator._internal_states)
    # No internal state at all: the configuration has no cell system and its pair factors come from factor type maps
    # (coulomb_atoms/power_bounded.ini). On the device that is EcmcProgram.no_cells: every other unit is a candidate
    # of every event and there are no cell-boundary events.
    no_cells = not internal_states
    leaf_cell_child = None  # the child index a cell system with a charge indicator stores, if there is one
    # General velocities: the sequential-direction end-of-chain handler rotates the velocity by an angle
    # (single_independent_active_sequential_direction_end_of_chain_event_handler.py:64-122). On the device that is the
    # disk kernel: two-dimensional composite point objects without a cell system whose pair factors -- all with hard
    # potentials -- come from factor type maps (hard_disk_dipoles/hard_disk_dipoles.ini, single_hard_disk_dipole.ini).
    sequential = any("SingleIndependentActiveSequentialDirectionEndOfChainEventHandler" in _class_names(handler)
                     for handler in activator.get_event_handlers())
    if sequential and not (no_cells and levels == 2 and dimension == 2):
        raise _configuration_error("the sequential-direction end-of-chain handler is supported for two-dimensional "
                                   "composite point objects without a cell system")
    if sequential:
        molecules, max_occupants, cells_per_side, neighbor_layers, cell_objects = False, 1, [1] * setting.dimension, 0, []
    elif no_cells:
        # composite point objects without cells are handled like those in root-level cells: whole objects are the
        # candidates (dipoles/dipole_factors_*.ini, water/single_molecule.ini)
        molecules, max_occupants, cells_per_side, neighbor_layers, cell_objects = levels == 2, 1, [1] * setting.dimension, 0, []
    else:
        if len(internal_states) != 1 or "SingleActiveCellOccupancy" not in _class_names(internal_states[0]):
            raise _configuration_error("at most one internal state, a SingleActiveCellOccupancy, is supported")
        occupancy = internal_states[0]
        # single_active_cell_occupancy.py:62-92: with `charge = <name>` only units with that charge unequal zero are stored
        # (and take part as active units). The device supports that for ONE kind of leaf of composite objects -- an
        # indicator charge that is non-zero for exactly one child index, the oxygen cells of
        # water/coulomb_power_bounded_lj_cell_bounded.ini (EcmcProgram.cell_child) --; any other configuration that relies on
        # the filter is refused instead of being run with neutral units in the cells.
        probe = _ProbeUnit()
        occupancy._is_relevant_unit(probe)
        if probe.charge.names:
            stored = [[k for k, child in enumerate(cnode.children) if occupancy._is_relevant_unit(child.value)]
                      for cnode in extracted_global_state]
            if levels != 2 or occupancy.cell_level != 2 or any(len(kinds) != 1 or kinds != stored[0] for kinds in stored):
                raise _configuration_error("SingleActiveCellOccupancy with a charge filter (charge = {0}) is only supported "
                                           "as a leaf-level cell system for one child index of composite objects"
                                           .format(probe.charge.names[0]))
            leaf_cell_child = stored[0][0]
        cells = occupancy.cells
        if "CuboidPeriodicCells" not in _class_names(cells):
            raise _configuration_error("cells must be CuboidPeriodicCells")
        # composite objects in root-level cells (water), or with a cell system for one kind of leaf
        molecules = levels == 2 and (occupancy.cell_level == 1 or leaf_cell_child is not None)
        if occupancy.cell_level != levels and not molecules:
            raise _configuration_error("cell_level must be 1 or the number of node levels")
        max_occupants = occupancy._maximum_number_occupants  # cell_occupancy.py:70-72
        unbounded = max_occupants <= 0
        if unbounded:
            # the reference keeps plain lists per cell and never uses the surplus; the device keeps `occupant_capacity`
            # slots per cell and no surplus, so that an overflow surfaces as a capacity error
            max_occupants, max_surplus = occupant_capacity, 0
        cells_per_side = list(cells._cells_per_side)
        neighbor_layers = cells._neighbor_layers
        cell_objects = list(cells.yield_cells())  # flat index order (cuboid_cells.py:144-146)

    # ---- handlers by kind
    pair_handlers, veto_handlers, boundary_handlers, eoc_handlers, start_handlers, control = [], [], [], [], [], []
    bounding_handlers, bond_handlers, bending_handlers, leaf_pair_handlers = [], [], [], []
    # root-unit-active mode (dipoles/dipole_motion.ini): the handlers that run while the ROOT unit of an object is the
    # independent active unit, and the RootLeafUnitActiveSwitcher handlers that alternate between the two modes
    root_pair_handlers, root_factor_handlers, switchers = [], [], []
    near_leaf_handlers = []  # TwoLeafUnitEventHandlerWithPiecewiseConstantBoundingPotential through the leaf cells
    # handlers fed by a factor type map are intramolecular factors (factor_type_map_in_state_tagger.py:83-107)
    factor_tagger_of = {}
    for tagger in activator._taggers:
        if "FactorTypeMapInStateTagger" in _class_names(tagger):
            for handler in tagger.get_event_handlers():
                factor_tagger_of[id(handler)] = tagger
    for handler in activator.get_event_handlers():
        names = _class_names(handler)
        if "RootUnitActiveTwoCompositeObjectSummedBoundingPotentialEventHandler" in names:
            root_pair_handlers.append(handler)
        elif "RootUnitActiveTwoLeafUnitEventHandler" in names:
            root_factor_handlers.append(handler)
        elif "RootLeafUnitActiveSwitcher" in names:
            switchers.append(handler)
        elif "TwoLeafUnitEventHandlerWithPiecewiseConstantBoundingPotential" in names:
            if leaf_cell_child is None or id(handler) in factor_tagger_of:
                raise _configuration_error("the piecewise-constant-bound two-leaf handler is supported for the nearby cells "
                                           "and the surplus of a leaf-level cell system with a charge indicator")
            near_leaf_handlers.append(handler)
        elif id(handler) in factor_tagger_of and levels == 1:
            # point masses: the factor type map lists the pair factors themselves ("[0, 1], Coulomb" = the active
            # atom with every other atom, factor_type_maps.py:333-347)
            factor_map = factor_tagger_of[id(handler)]._factor_type_map
            entries = {tuple(indices) for lists in factor_map.map.values() for indices in lists}
            if not no_cells or entries != {(0, 1)}:
                raise _configuration_error("factor type maps of point masses must hold the one pair factor [0, 1] "
                                           "and need a configuration without cells")
            if not names & {"TwoLeafUnitBoundingPotentialEventHandler", "TwoLeafUnitEventHandler"}:
                raise _configuration_error("factor-type-map handler {0} has no device implementation"
                                           .format(type(handler).__name__))
            pair_handlers.append(handler)
        elif id(handler) in factor_tagger_of and (no_cells or leaf_cell_child is not None) and \
                "TwoCompositeObjectSummedBoundingPotentialEventHandler" in names:
            # "[0, 1, 2, 3], Coulomb": the object of the active leaf with every other object
            factor_map = factor_tagger_of[id(handler)]._factor_type_map
            entries = {tuple(indices) for lists in factor_map.map.values() for indices in lists}
            if entries != {tuple(range(2 * nodes_per_root))}:
                raise _configuration_error("the composite-object factor must join all leaves of two objects")
            pair_handlers.append(handler)
        elif id(handler) in factor_tagger_of and no_cells and "TwoLeafUnitBoundingPotentialEventHandler" in names:
            # "[0, 2], Coulomb", "[0, 3], Coulomb", ...: the Coulomb interaction as bounded leaf-to-leaf factors between
            # the objects; the device handles them only if every leaf of one object is paired with every leaf of the other
            factor_map = factor_tagger_of[id(handler)]._factor_type_map
            entries = {tuple(indices) for lists in factor_map.map.values() for indices in lists}
            wanted = {(a, nodes_per_root + b) for a in range(nodes_per_root) for b in range(nodes_per_root)}
            if entries != wanted:
                raise _configuration_error("bounded leaf-to-leaf factors must join every leaf of one object with every "
                                           "leaf of the other")
            leaf_pair_handlers.append(handler)
        elif id(handler) in factor_tagger_of:
            if "FixedSeparationsEventHandlerWithPiecewiseConstantBoundingPotential" in names:
                bending_handlers.append(handler)
            elif "TwoLeafUnitEventHandler" in names:
                bond_handlers.append(handler)
            else:
                raise _configuration_error("factor-type-map handler {0} has no device implementation"
                                           .format(type(handler).__name__))
        elif "TwoCompositeObjectSummedBoundingPotentialEventHandler" in names:
            pair_handlers.append(handler)
        elif names & {"TwoLeafUnitCellBoundingPotentialEventHandler", "TwoCompositeObjectCellBoundingPotentialEventHandler"}:
            bounding_handlers.append(handler)
        elif "TwoLeafUnitBoundingPotentialEventHandler" in names or "TwoLeafUnitEventHandler" in names:
            pair_handlers.append(handler)
        elif "LeafUnitCellVetoEventHandler" in names or "CompositeObjectCellVetoEventHandler" in names:
            veto_handlers.append(handler)
        elif "CellBoundaryEventHandler" in names:
            boundary_handlers.append(handler)
        elif "SingleIndependentActivePeriodicDirectionEndOfChainEventHandler" in names:
            # (the sequential-direction handler is a subclass: same chain time and draw of the next active unit)
            eoc_handlers.append(handler)
        elif "InitialChainStartOfRunEventHandler" in names:
            start_handlers.append(handler)
        elif names & {"SamplingEventHandler", "EndOfRunEventHandler", "DumpingEventHandler"}:
            # host control events; a dumping event writes the device checkpoint (CudaBatchedMediator._dump)
            control.append(handler)
        else:
            raise _configuration_error("event handler {0} has no device implementation".format(type(handler).__name__))
    if len(eoc_handlers) != 1 or len(start_handlers) != 1 or (not boundary_handlers and not no_cells):
        raise _configuration_error("exactly one end-of-chain handler, one start-of-run handler and a cell-boundary "
                                   "handler are required")
    if no_cells and (boundary_handlers or veto_handlers or bounding_handlers):
        raise _configuration_error("cell handlers without a cell system")
    start, eoc = start_handlers[0], eoc_handlers[0]
    velocity = list(start._initial_velocity)  # initial_chain_start_of_run_event_handler.py:86-88
    moving = [d for d, v in enumerate(velocity) if v != 0.0]
    if len(moving) != 1 or velocity[moving[0]] <= 0.0:
        raise _configuration_error("the initial velocity must be along one positive axis")
    initial_active = tuple(start._initial_active_identifier)
    if len(initial_active) != levels:
        raise _configuration_error("the initial active identifier must name a leaf unit")
    initial_leaf = initial_active[0] if levels == 1 else initial_active[0] * nodes_per_root + initial_active[1]
    if not any("EndOfRunEventHandler" in _class_names(h) for h in control):
        raise _configuration_error("an end-of-run handler is required")

    builder = ProgramBuilder(dimension, n_particles, length, setting.beta, cells_per_side, neighbor_layers,
                             max_occupants=max_occupants,
                             max_surplus=n_particles if max_surplus is None else max_surplus,
                             chain_time=eoc._chain_time, speed=velocity[moving[0]], initial_direction=moving[0],
                             initial_active=initial_leaf, seed=seed, no_cells=no_cells)

    # ---- composite point objects: factors of the factor type map
    inter_factors, inter_potential, bending = [], None, None
    if levels == 2:
        bonds, bond_potential = [], None
        for handler in bond_handlers:
            factor_map = factor_tagger_of[id(handler)]._factor_type_map
            if _charge_name(handler._charges) is not None:
                raise _configuration_error("factor-type-map pair factors with charges are not supported")
            potential = potential_descriptor(handler._potential)
            local = getattr(factor_map, "_local", None)
            if local is None:
                raise _configuration_error("a factor type map without entries in the factor set file is not supported")
            if local:
                if bond_potential is not None and not _same_potential(bond_potential, potential):
                    raise _configuration_error("all intramolecular pair factors must share one potential")
                bond_potential = potential
            else:
                if not molecules and not sequential:
                    raise _configuration_error("factors between composite objects need root-level cells")
                if inter_potential is not None and not _same_potential(inter_potential, potential):
                    raise _configuration_error("all pair factors between composite objects must share one potential")
                inter_potential = potential
            for child, entries in factor_map.map.items():
                for indices in entries:
                    if len(indices) != 2:
                        raise _configuration_error("only two-unit factors are supported by the pair handlers")
                    if local:
                        if tuple(sorted(indices)) not in bonds:
                            bonds.append(tuple(sorted(indices)))
                    else:
                        # "[1, 4]": the active child `child` against child (other - nodes_per_root) of every other object
                        other = [index for index in indices if index >= nodes_per_root]
                        own = [index for index in indices if index < nodes_per_root]
                        if len(other) != 1 or own != [child]:
                            raise _configuration_error("unsupported factor between composite objects: {0}".format(indices))
                        pair = (child, other[0] - nodes_per_root)
                        if pair not in inter_factors:
                            inter_factors.append(pair)
        builder.set_composite(nodes_per_root, bonds, bond_potential)
        if sequential:
            hard = (abi.POT_HARD_SPHERE, abi.POT_HARD_DIPOLE)
            if pair_handlers or leaf_pair_handlers or bending_handlers or \
                    any(potential is not None and potential.kind not in hard for potential in (bond_potential, inter_potential)):
                raise _configuration_error("general velocities are supported for hard potentials in factor type maps only")
            # the handler keeps cos / sin of the angle (:97-99); the angle in degrees is recovered for the builder, which
            # recomputes them by the handler's own expressions -- and the result is checked against the handler's values
            import math
            delta_phi_degree = math.degrees(math.atan2(eoc._sin_delta_phi, eoc._cos_delta_phi)) % 360.0
            builder.set_sequential_direction(delta_phi_degree, inter_potential, inter_factors)
            builder.program.eoc_cos, builder.program.eoc_sin = eoc._cos_delta_phi, eoc._sin_delta_phi
            inter_factors, inter_potential = [], None
        if bending_handlers:
            first = bending_handlers[0]
            factor_map = factor_tagger_of[id(first)]._factor_type_map
            entries = {tuple(indices) for lists in factor_map.map.values() for indices in lists}
            if len(entries) != 1 or len(next(iter(entries))) != 3 or not getattr(factor_map, "_local", False):
                raise _configuration_error("exactly one local three-unit (bending) factor is supported")
            bending = dict(children=list(next(iter(entries))), separations=list(first._separations),
                           potential=potential_descriptor(first._potential), offset=first._offset,
                           max_displacement=first._max_displacement, lifting=_lifting_kind(first._lifting))
            if bending["potential"].kind != abi.POT_BENDING or not molecules:
                raise _configuration_error("the three-unit factor must be a bending potential in root-level cells")
    elif bond_handlers or bending_handlers:
        raise _configuration_error("factor type maps need composite point objects")

    # ---- pair factor
    charge_names = set()
    composite_lifting = None
    if pair_handlers and "TwoCompositeObjectSummedBoundingPotentialEventHandler" in _class_names(pair_handlers[0]):
        # two_composite_object_summed_bounding_potential_event_handler.py:65-117
        if not molecules:
            raise _configuration_error("composite-object pair handlers need root-level cells")
        first = pair_handlers[0]
        potential = potential_descriptor(first._potential)
        bounding = potential_descriptor(first._bounding_potential)
        charge = _charge_name(first._potential_charges)
        composite_lifting = _lifting_kind(first._lifting)
        for handler in pair_handlers[1:]:
            if "TwoCompositeObjectSummedBoundingPotentialEventHandler" not in _class_names(handler) or \
                    not _same_potential(potential, potential_descriptor(handler._potential)) or \
                    not _same_potential(bounding, potential_descriptor(handler._bounding_potential)) or \
                    _charge_name(handler._potential_charges) != charge or \
                    _charge_name(handler._bounding_potential_charges) != charge or \
                    _lifting_kind(handler._lifting) != composite_lifting:
                raise _configuration_error("nearby and surplus composite-object handlers must be alike")
        builder.set_pair(abi.PAIR_TWO_COMPOSITE_SUMMED_BOUNDING, potential, bounding, use_charge=charge is not None)
        if charge is not None:
            charge_names.add(charge)
    elif leaf_pair_handlers:
        # molecules without cells, Coulomb as bounded factors between leaves of different objects
        if pair_handlers or not molecules:
            raise _configuration_error("bounded leaf-to-leaf factors cannot be combined with other pair handlers")
        first = leaf_pair_handlers[0]
        potential = potential_descriptor(first._potential)
        bounding = potential_descriptor(first._bounding_potential)
        charge = _charge_name(first._potential_charges)
        for handler in leaf_pair_handlers[1:]:
            if not _same_potential(potential, potential_descriptor(handler._potential)) or \
                    not _same_potential(bounding, potential_descriptor(handler._bounding_potential)) or \
                    _charge_name(handler._potential_charges) != charge:
                raise _configuration_error("the bounded leaf-to-leaf handlers must be alike")
        builder.set_pair(abi.PAIR_TWO_LEAF_UNIT_BOUNDING, potential, bounding, use_charge=charge is not None)
        if charge is not None:
            charge_names.add(charge)
    elif pair_handlers:
        if molecules:
            raise _configuration_error("leaf-unit pair handlers in root-level cells are not supported")
        first = pair_handlers[0]
        bounded = "TwoLeafUnitBoundingPotentialEventHandler" in _class_names(first)
        potential = potential_descriptor(first._potential)
        bounding = potential_descriptor(first._bounding_potential) if bounded else None
        charge = _charge_name(first._potential_charges if bounded else first._charges)
        for handler in pair_handlers[1:]:
            same_kind = ("TwoLeafUnitBoundingPotentialEventHandler" in _class_names(handler)) == bounded
            if not same_kind or not _same_potential(potential, potential_descriptor(handler._potential)) or \
                    (bounded and not _same_potential(bounding, potential_descriptor(handler._bounding_potential))) or \
                    _charge_name(handler._potential_charges if bounded else handler._charges) != charge:
                raise _configuration_error("nearby and surplus pair handlers must share potential, bounding potential "
                                           "and charge")
        builder.set_pair(abi.PAIR_TWO_LEAF_UNIT_BOUNDING if bounded else abi.PAIR_TWO_LEAF_UNIT, potential, bounding,
                         use_charge=charge is not None)
        if charge is not None:
            charge_names.add(charge)

    # ---- cell veto: the tables the reference built in CellVetoEventHandler.initialize (cell_veto_event_handler.py:134-159)
    if len(veto_handlers) > 1:
        raise _configuration_error("more than one cell-veto handler")
    if veto_handlers:
        veto = veto_handlers[0]
        index_of = {cell: index for index, cell in enumerate(cell_objects)}
        bounds = np.zeros((len(cell_objects), dimension, 2))
        for cell, per_direction in veto._derivative_bounds.items():
            for d in range(dimension):
                bounds[index_of[cell], d, 0] = per_direction[d][0]
                bounds[index_of[cell], d, 1] = per_direction[d][1]
        tables = {"upper": [], "lower": [], "bounds": bounds}
        for name, walkers in (("upper", veto._upper_bound_walker), ("lower", veto._lower_bound_walker)):
            for d in range(dimension):
                walker = walkers[d]  # walker.py:69-103: entries (small item, large item) or (item,)
                tables[name].append({
                    "cell_a": np.array([index_of[entry[0].item] for entry in walker._table], dtype=np.int32),
                    "cell_b": np.array([index_of[entry[1].item] if len(entry) > 1 else -1 for entry in walker._table],
                                       dtype=np.int32),
                    "rate_a": np.array([entry[0].rate for entry in walker._table], dtype=np.float64),
                    "total_rate": walker.total_rate, "mean_rate": walker._mean_rate})
        veto_charge = veto._charge
        target_charge = getattr(veto._estimator, "_target_charge", None)
        builder.set_veto(potential_descriptor(veto._potential), tables, use_charge=veto_charge is not None,
                         target_charge=1.0 if target_charge is None else target_charge)
        if veto_charge is not None:
            charge_names.add(veto_charge)
    # ---- cell bounding: one TwoLeafUnitCellBoundingPotentialEventHandler per possible far target, all alike
    # (two_leaf_unit_cell_bounding_potential_event_handler.py:65-110); the bounds were built by
    # CellBoundingPotential.initialize (cell_bounding_potential.py:96-153) as (upper dict, lower dict or None)
    if bounding_handlers:
        if veto_handlers:
            raise _configuration_error("cell-veto and cell-bounding handlers for the same far field")
        first = bounding_handlers[0]
        potential = potential_descriptor(first._potential)
        charge = first._charge
        composite_bounding = "TwoCompositeObjectCellBoundingPotentialEventHandler" in _class_names(first)
        if composite_bounding != (molecules and leaf_cell_child is None):
            raise _configuration_error("composite-object cell-bounding handlers need root-level cells, leaf-unit ones "
                                       "leaf-level cells")
        for handler in bounding_handlers[1:]:
            if not _same_potential(potential, potential_descriptor(handler._potential)) or handler._charge != charge or \
                    type(handler) is not type(first):
                raise _configuration_error("cell-bounding handlers must share potential and charge")
        if composite_bounding:
            # charge correction factor = active charge x max |target charges| (dipole_monte_carlo_estimator.py:158-186)
            if "DipoleMonteCarloEstimator" not in _class_names(first._bounding_potential._estimator):
                raise _configuration_error("the composite-object cell-bounding handler needs the dipole Monte Carlo estimator")
            bounding_lifting = _lifting_kind(first._lifting)
            if composite_lifting is not None and bounding_lifting != composite_lifting:
                raise _configuration_error("the composite-object handlers must share one lifting scheme")
            composite_lifting = bounding_lifting
        index_of = {cell: index for index, cell in enumerate(cell_objects)}
        bounds = np.zeros((len(cell_objects), dimension, 2))
        stored = first._bounding_potential._derivative_bounds  # a bare dict when no lower bounds were asked for
        upper, lower = (stored, None) if isinstance(stored, dict) else stored
        for cell, per_direction in upper.items():
            for d in range(dimension):
                bounds[index_of[cell], d, 0] = per_direction[d]
                if lower is not None:
                    bounds[index_of[cell], d, 1] = -lower[cell][d]
        target_charge = getattr(first._bounding_potential._estimator, "_target_charge", None)
        builder.set_cell_bounding(potential, bounds, use_charge=charge is not None,
                                  target_charge=1.0 if target_charge is None else target_charge)
        if charge is not None:
            charge_names.add(charge)
    # ---- molecules: lifting scheme of the composite-object handlers, factors between objects, bending
    if molecules:
        if veto_handlers:
            if "CompositeObjectCellVetoEventHandler" not in _class_names(veto_handlers[0]):
                raise _configuration_error("root-level cells need the composite-object cell-veto handler")
            veto_lifting = _lifting_kind(veto_handlers[0]._lifting)
            if composite_lifting is not None and veto_lifting != composite_lifting:
                raise _configuration_error("the composite-object handlers must share one lifting scheme")
            composite_lifting = veto_lifting
        # Does a cell-boundary event of the root trash the leaf-level factor handlers? (tag lists, tagger.py:163-200)
        keeps_factors = False
        if not no_cells:
            boundary_tagger = [tagger for tagger in activator._taggers
                               if any(handler is boundary_handlers[0] for handler in tagger.get_event_handlers())][0]
            factor_tags = {tagger.tag for tagger in factor_tagger_of.values()}
            trashed = set(boundary_tagger.trashes)
            if factor_tags and factor_tags & trashed and not factor_tags <= trashed:
                raise _configuration_error("a cell-boundary event must trash all or none of the factor-type-map "
                                           "handlers")
            keeps_factors = not (factor_tags and factor_tags <= trashed)
        if leaf_cell_child is not None:
            # the two-leaf factor between the stored leaves is found through the cells: nearby cells and surplus by the
            # piecewise-constant-bound handler, every other cell by the leaf-level cell-bounding handler
            if not near_leaf_handlers or inter_factors or veto_handlers or not keeps_factors or occupancy.cell_level != 2 or \
                    not pair_handlers or any(id(h) not in factor_tagger_of for h in pair_handlers):
                raise _configuration_error("a leaf-level cell system with a charge indicator needs composite-object pair "
                                           "factors from a factor type map, piecewise-constant-bound handlers for its nearby "
                                           "cells and a cell-boundary event that keeps the factor-type-map handlers")
            first = near_leaf_handlers[0]
            inter_potential = potential_descriptor(first._potential)
            for handler in near_leaf_handlers:
                if _charge_name(handler._charges) is not None or handler._offset != first._offset or handler._max_displacement != first._max_displacement or \
                        not _same_potential(inter_potential, potential_descriptor(handler._potential)):
                    raise _configuration_error("the piecewise-constant-bound handlers must be alike and chargeless")
            if bounding_handlers and (bounding_handlers[0]._charge is not None or not _same_potential(
                    inter_potential, potential_descriptor(bounding_handlers[0]._potential))):
                raise _configuration_error("the cell-bounding handler of the leaf cells must use the potential of the "
                                           "piecewise-constant-bound handlers, without charges")
            inter_factors = [(leaf_cell_child, leaf_cell_child)]
        elif near_leaf_handlers:
            raise _configuration_error("piecewise-constant-bound two-leaf handlers need a leaf-level cell system")
        builder.set_molecules(composite_lifting if composite_lifting is not None else abi.LIFTING_INSIDE_FIRST,
                              inter_factors=inter_factors, inter_potential=inter_potential, bending=bending,
                              boundary_keeps_factors=keeps_factors)
        if leaf_cell_child is not None:
            builder.set_leaf_cells(leaf_cell_child, near_leaf_handlers[0]._offset, near_leaf_handlers[0]._max_displacement)
    elif veto_handlers and "CompositeObjectCellVetoEventHandler" in _class_names(veto_handlers[0]):
        raise _configuration_error("the composite-object cell-veto handler needs root-level cells")
    # ---- root-unit-active mode: the same factors with the root unit of an object active, and the switchers
    if root_pair_handlers or root_factor_handlers or switchers:
        # root_leaf_unit_active_switcher.py:59-100 (aim modes), :102-127 (chain length)
        aims = sorted(switcher._aim_mode.name for switcher in switchers)
        if not (no_cells and molecules and nodes_per_root == 2 and aims == ["leaf_unit_active", "root_unit_active"]):
            raise _configuration_error("the root-unit-active mode needs composite objects of two leaves without a cell "
                                       "system and one RootLeafUnitActiveSwitcher per aim mode")
        leaf_pair = [h for h in pair_handlers if "TwoCompositeObjectSummedBoundingPotentialEventHandler" in _class_names(h)]
        if not leaf_pair or not root_pair_handlers or bending_handlers or leaf_pair_handlers:
            raise _configuration_error("the root-unit-active mode needs the composite-object pair handler in both modes")
        first = leaf_pair[0]
        for handler in root_pair_handlers:
            # root_unit_active_two_composite_object_summed_bounding_potential_event_handler.py:66-114
            factor_map = factor_tagger_of[id(handler)]._factor_type_map
            entries = {tuple(indices) for lists in factor_map.map.values() for indices in lists}
            if entries != {tuple(range(2 * nodes_per_root))} or \
                    not _same_potential(potential_descriptor(first._potential), potential_descriptor(handler._potential)) or \
                    not _same_potential(potential_descriptor(first._bounding_potential),
                                        potential_descriptor(handler._bounding_potential)) or \
                    _charge_name(handler._potential_charges) != _charge_name(first._potential_charges) or \
                    _charge_name(handler._bounding_potential_charges) != _charge_name(first._potential_charges):
                raise _configuration_error("the root-unit-active composite-object handler must use the factor, the "
                                           "potentials and the charge of the leaf-unit-active one")
        root_factors = []
        for handler in root_factor_handlers:
            factor_map = factor_tagger_of[id(handler)]._factor_type_map
            if _charge_name(handler._charges) is not None or inter_potential is None or getattr(factor_map, "_local", True) or \
                    not _same_potential(inter_potential, potential_descriptor(handler._potential)):
                raise _configuration_error("the root-unit-active two-leaf handlers must use the potential of the "
                                           "leaf-unit-active factors between the objects")
            for child, entries in factor_map.map.items():
                for indices in entries:
                    other = [index for index in indices if index >= nodes_per_root]
                    own = [index for index in indices if index < nodes_per_root]
                    if len(indices) != 2 or len(other) != 1 or own != [child]:
                        raise _configuration_error("unsupported factor between composite objects: {0}".format(indices))
                    root_factors.append((child, other[0] - nodes_per_root))
        if sorted(set(root_factors)) != sorted(inter_factors):
            raise _configuration_error("the root-unit-active two-leaf handlers must cover the factors between the objects "
                                       "of the leaf-unit-active mode")
        # the run starts with a leaf unit active: the start-of-run tagger activates the switcher that aims at the root unit
        tagger_of = {id(handler): tagger for tagger in activator._taggers for handler in tagger.get_event_handlers()}
        to_root = [switcher for switcher in switchers if switcher._aim_mode.name == "root_unit_active"][0]
        to_leaf = [switcher for switcher in switchers if switcher._aim_mode.name == "leaf_unit_active"][0]
        start_tagger = tagger_of[id(start)]
        if tagger_of[id(to_root)].tag not in start_tagger.activates or tagger_of[id(to_leaf)].tag in start_tagger.activates:
            raise _configuration_error("the run must start with the leaf-to-root switcher activated")
        builder.set_root_mode(to_root._chain_length, to_leaf._chain_length)
    if len(charge_names) > 1:
        raise _configuration_error("pair and cell-veto handlers use different charges: {0}".format(sorted(charge_names)))
    return CompiledProgram(builder, charge_names.pop() if charge_names else None, control, n_particles, nodes_per_root)


def positions_and_charges(extracted_global_state, charge_name):
    """(positions[N][D] of the leaf units, charges[N] or None, roots[N_root][D] or None) of root cnodes
    (tree_state_handler.py:213-230); leaves in flat order root * nodes_per_root + child."""
    composite = bool(extracted_global_state[0].children)
    leaves = ([child for cnode in extracted_global_state for child in cnode.children] if composite
              else list(extracted_global_state))
    positions = np.array([leaf.value.position for leaf in leaves], dtype=np.float64)
    charges = None
    if charge_name is not None:
        charges = np.array([leaf.value.charge[charge_name] for leaf in leaves], dtype=np.float64)
    roots = np.array([cnode.value.position for cnode in extracted_global_state], dtype=np.float64) if composite else None
    return positions, charges, roots
