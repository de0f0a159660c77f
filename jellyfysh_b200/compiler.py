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
        return abi.EcmcPotential.make(abi.POT_HARD_SPHERE, (potential._diameter_squared / 4.0) ** 0.5)
    if "MergedImageCoulombPotential" in names:
        # merged_image_coulomb_potential.py:111-118
        return abi.EcmcPotential.make(abi.POT_MERGED_IMAGE_COULOMB, potential._prefactor, potential._alpha,
                                      potential._fourier_cutoff, potential._position_cutoff)
    if "InversePowerCoulombBoundingPotential" in names:
        return abi.EcmcPotential.make(abi.POT_INVERSE_POWER_COULOMB_BOUNDING, potential._prefactor)
    raise _configuration_error("potential {0} has no device implementation".format(type(potential).__name__))


def _same_potential(a, b):
    return a.kind == b.kind and list(a.params) == list(b.params)


class CompiledProgram:
    """The result: a ProgramBuilder plus what the mediator needs around it."""

    def __init__(self, builder, charge_name, control_handlers, n_particles):
        self.builder = builder
        self.charge_name = charge_name
        self.control_handlers = control_handlers  # sampling / end-of-run handlers, run on the host
        self.n_particles = n_particles


def compile_program(activator, extracted_global_state, seed=0, max_surplus=None):
    """Walk the initialized activator and return a CompiledProgram.

    extracted_global_state: state_handler.extract_global_state() (root cnodes), used for N and the node structure."""
    import jellyfysh.setting as setting
    from jellyfysh.setting import hypercubic_setting

    if setting.number_of_node_levels != 1:
        raise _configuration_error("only point-mass (single level) systems run on the device in this version; composite "
                                   "objects (dipoles, water) need the composite-object handlers")
    for cnode in extracted_global_state:
        if cnode.children:
            raise _configuration_error("root nodes with children are not supported")
    if not hypercubic_setting.initialized():
        raise _configuration_error("a hypercubic setting is required")
    dimension, length, n_particles = setting.dimension, hypercubic_setting.system_length, len(extracted_global_state)

    # ---- cell occupancy (tag_activator.py:82-135 keeps the internal states in _internal_states)
    internal_states = list(activator._internal_states)
    if len(internal_states) != 1 or "SingleActiveCellOccupancy" not in _class_names(internal_states[0]):
        raise _configuration_error("exactly one SingleActiveCellOccupancy internal state is required")
    occupancy = internal_states[0]
    cells = occupancy.cells
    if "CuboidPeriodicCells" not in _class_names(cells):
        raise _configuration_error("cells must be CuboidPeriodicCells")
    if occupancy.cell_level != 1:
        raise _configuration_error("cell_level must be 1")
    max_occupants = occupancy._maximum_number_occupants  # cell_occupancy.py:70-72
    if max_occupants <= 0:
        raise _configuration_error("an unbounded number of occupants per cell is not supported: set maximum_number_occupants")
    cells_per_side = list(cells._cells_per_side)
    neighbor_layers = cells._neighbor_layers
    cell_objects = list(cells.yield_cells())  # flat index order (cuboid_cells.py:144-146)

    # ---- handlers by kind
    pair_handlers, veto_handlers, boundary_handlers, eoc_handlers, start_handlers, control = [], [], [], [], [], []
    bounding_handlers = []
    for handler in activator.get_event_handlers():
        names = _class_names(handler)
        if "TwoLeafUnitCellBoundingPotentialEventHandler" in names:
            bounding_handlers.append(handler)
        elif "TwoLeafUnitBoundingPotentialEventHandler" in names or "TwoLeafUnitEventHandler" in names:
            pair_handlers.append(handler)
        elif "LeafUnitCellVetoEventHandler" in names:
            veto_handlers.append(handler)
        elif "CellBoundaryEventHandler" in names:
            boundary_handlers.append(handler)
        elif "SingleIndependentActivePeriodicDirectionEndOfChainEventHandler" in names:
            eoc_handlers.append(handler)
        elif "InitialChainStartOfRunEventHandler" in names:
            start_handlers.append(handler)
        elif names & {"SamplingEventHandler", "EndOfRunEventHandler"}:
            control.append(handler)
        elif "DumpingEventHandler" in names:
            raise _configuration_error("dumping handlers are not supported (device state is not picklable)")
        else:
            raise _configuration_error("event handler {0} has no device implementation".format(type(handler).__name__))
    if len(eoc_handlers) != 1 or len(start_handlers) != 1 or not boundary_handlers:
        raise _configuration_error("exactly one end-of-chain handler, one start-of-run handler and a cell-boundary "
                                   "handler are required")
    start, eoc = start_handlers[0], eoc_handlers[0]
    velocity = list(start._initial_velocity)  # initial_chain_start_of_run_event_handler.py:86-88
    moving = [d for d, v in enumerate(velocity) if v != 0.0]
    if len(moving) != 1 or velocity[moving[0]] <= 0.0:
        raise _configuration_error("the initial velocity must be along one positive axis")
    initial_active = tuple(start._initial_active_identifier)
    if len(initial_active) != 1:
        raise _configuration_error("the initial active identifier must name a root node")
    if not any("EndOfRunEventHandler" in _class_names(h) for h in control):
        raise _configuration_error("an end-of-run handler is required")

    builder = ProgramBuilder(dimension, n_particles, length, setting.beta, cells_per_side, neighbor_layers,
                             max_occupants=max_occupants,
                             max_surplus=n_particles if max_surplus is None else max_surplus,
                             chain_time=eoc._chain_time, speed=velocity[moving[0]], initial_direction=moving[0],
                             initial_active=initial_active[0], seed=seed)

    # ---- pair factor
    charge_names = set()
    if pair_handlers:
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
        for handler in bounding_handlers[1:]:
            if not _same_potential(potential, potential_descriptor(handler._potential)) or handler._charge != charge:
                raise _configuration_error("cell-bounding handlers must share potential and charge")
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
    if len(charge_names) > 1:
        raise _configuration_error("pair and cell-veto handlers use different charges: {0}".format(sorted(charge_names)))
    return CompiledProgram(builder, charge_names.pop() if charge_names else None, control, n_particles)


def positions_and_charges(extracted_global_state, charge_name):
    """(positions[N][D], charges[N] or None) of root cnodes (tree_state_handler.py:213-230)."""
    positions = np.array([cnode.value.position for cnode in extracted_global_state], dtype=np.float64)
    charges = None
    if charge_name is not None:
        charges = np.array([cnode.value.charge[charge_name] for cnode in extracted_global_state], dtype=np.float64)
    return positions, charges
