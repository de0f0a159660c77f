"""Assembly of an EcmcProgram (include/ecmc.h) from plain Python values: the parameter-carrier side of the
host layer. Pure data, no computation; keeps every array the C struct points to alive."""
import ctypes as C

import numpy as np

from jellyfysh_b200 import abi


class ProgramBuilder:
    """Assembles an EcmcProgram and keeps every array it points to alive."""

    def __init__(self, dimension, n_particles, system_length, beta, cells_per_side, neighbor_layers=1,
                 max_occupants=1, max_surplus=64, chain_time=1.0, speed=1.0, initial_direction=0,
                 initial_active=0, seed=0, no_cells=False):
        """no_cells: the configuration has no cell system (every other unit is a candidate of every event, no
        cell-boundary events); cells_per_side and neighbor_layers are then ignored."""
        if no_cells:
            cells_per_side, neighbor_layers, max_occupants = [1] * dimension, 0, 1
            max_surplus = max(max_surplus, n_particles)
        self.program = abi.EcmcProgram()
        p = self.program
        p.abi_version = abi.ECMC_ABI_VERSION
        p.dimension = dimension
        p.n_particles = n_particles
        p.system_length = system_length
        p.beta = beta
        cps = list(cells_per_side) + [1] * (3 - len(cells_per_side))
        for d in range(3):
            p.cells_per_side[d] = cps[d] if d < dimension else 1
        p.neighbor_layers = neighbor_layers
        p.max_occupants = max_occupants
        p.max_surplus = max_surplus
        p.chain_time = chain_time
        p.speed = speed
        p.initial_direction = initial_direction
        p.initial_active = initial_active
        p.seed = seed
        p.no_cells = int(no_cells)
        self._keep = []
        self.tables = None

    def fingerprint(self) -> bytes:
        """32 bytes that identify the program (every scalar of the EcmcProgram incl. seed, potentials and handler kinds,
        plus the far-field tables): stored in checkpoints so that a dump of another program is refused on load."""
        import hashlib
        copy = abi.EcmcProgram.from_buffer_copy(bytes(self.program))
        copy.veto_tables = None  # an address, not content
        digest = hashlib.sha256(bytes(copy))
        for array in self._keep:
            if isinstance(array, np.ndarray):
                digest.update(array.tobytes())
        return digest.digest()

    @property
    def n_cells(self):
        p = self.program
        return int(np.prod([p.cells_per_side[d] for d in range(p.dimension)]))

    def set_pair(self, handler, potential, bounding=None, use_charge=False):
        p = self.program
        p.pair_handler = handler
        p.pair_potential = potential
        if bounding is not None:
            p.pair_bounding_potential = bounding
        p.pair_use_charge = int(use_charge)
        return self

    def set_composite(self, nodes_per_root, bonds=(), bond_potential=None):
        """Composite point objects: n_particles leaves in groups of nodes_per_root; `bonds` = pairs of child indices
        handled by a TwoLeafUnitEventHandler with bond_potential (factor type map, e.g. "[0, 1], Dipole")."""
        p = self.program
        if p.n_particles % nodes_per_root:
            raise ValueError("n_particles must be a multiple of nodes_per_root")
        if len(bonds) > abi.ECMC_MAX_BONDS:
            raise ValueError("too many bonds")
        p.nodes_per_root = nodes_per_root
        p.n_bonds = len(bonds)
        for i, (a, b) in enumerate(bonds):
            p.bonds[i][0], p.bonds[i][1] = a, b
        if bond_potential is not None:
            p.bond_potential = bond_potential
        return self

    def set_sequential_direction(self, delta_phi_degree, sphere_potential, sphere_factors):
        """General velocities (two-dimensional composite point objects without a cell system): the end-of-chain handler
        SingleIndependentActiveSequentialDirectionEndOfChainEventHandler rotates the velocity by delta_phi_degree;
        sphere_factors = (child a, child b) leaf factors between different objects with sphere_potential (hard spheres)."""
        import math
        p = self.program
        if p.dimension != 2 or not p.no_cells:
            raise ValueError("general velocities: two dimensions, no cell system")
        delta_phi = delta_phi_degree * math.pi / 180.0  # the handler's own expressions (:97-99)
        p.eoc_sequential = 1
        p.eoc_cos = math.cos(delta_phi)
        p.eoc_sin = math.sin(delta_phi)
        if len(sphere_factors) > abi.ECMC_MAX_INTER_FACTORS:
            raise ValueError("too many inter-object factors")
        p.n_inter_factors = len(sphere_factors)
        for i, (a, b) in enumerate(sphere_factors):
            p.inter_factors[i][0], p.inter_factors[i][1] = a, b
        if sphere_potential is not None:
            p.inter_potential = sphere_potential
        return self

    def set_molecules(self, lifting, inter_factors=(), inter_potential=None, bending=None, boundary_keeps_factors=True):
        """Composite objects in root-level cells (water): the pair handler set by set_pair(PAIR_TWO_COMPOSITE_...)
        and the cell veto act on whole objects with the given lifting scheme; inter_factors = (child a, child b) leaf
        factors between different objects with inter_potential; bending = dict(children, separations, potential,
        offset, max_displacement, lifting)."""
        p = self.program
        p.cell_level = 1
        p.composite_lifting = lifting
        p.boundary_keeps_factors = int(boundary_keeps_factors)
        if len(inter_factors) > abi.ECMC_MAX_INTER_FACTORS:
            raise ValueError("too many inter-object factors")
        p.n_inter_factors = len(inter_factors)
        for i, (a, b) in enumerate(inter_factors):
            p.inter_factors[i][0], p.inter_factors[i][1] = a, b
        if inter_potential is not None:
            p.inter_potential = inter_potential
        if bending is not None:
            p.bending_enabled = 1
            p.bending_lifting = bending["lifting"]
            for i, child in enumerate(bending["children"]):
                p.bending_children[i] = child
            for i, index in enumerate(bending["separations"]):
                p.bending_separations[i] = index
            p.bending_potential = bending["potential"]
            p.bending_offset = bending["offset"]
            p.bending_max_displacement = bending["max_displacement"]
        return self

    def set_root_mode(self, leaf_to_root_chain_length, root_to_leaf_chain_length):
        """Root-unit-active mode (dipoles/dipole_motion.ini): two RootLeafUnitActiveSwitcher handlers alternate the
        independent active unit between a leaf unit (for leaf_to_root_chain_length) and the root unit of its object (for
        root_to_leaf_chain_length); the root-unit-active handlers use the pair and inter-object potentials of the leaf
        mode. Needs set_pair(PAIR_TWO_COMPOSITE_SUMMED_BOUNDING, ...), set_composite(2, ...), no cell system."""
        p = self.program
        if p.nodes_per_root != 2 or not p.no_cells or p.pair_handler != abi.PAIR_TWO_COMPOSITE_SUMMED_BOUNDING:
            raise ValueError("root-unit-active mode: composite objects of two leaves with the composite-object pair "
                             "handler, no cell system")
        p.root_mode = 1
        p.switch_chain_length[0] = leaf_to_root_chain_length
        p.switch_chain_length[1] = root_to_leaf_chain_length
        return self

    def set_leaf_cells(self, child, offset, max_displacement):
        """A cell system that holds only the leaves with child index `child` of every object (SingleActiveCellOccupancy with
        cell_level = 2 and a charge indicator, water/coulomb_power_bounded_lj_cell_bounded.ini): the two-leaf factor between
        these leaves of different objects (inter_potential, set_molecules(inter_factors=[(child, child)])) is found through
        the cells -- nearby cells and surplus with the piecewise constant bound (offset, max_displacement), all other cells
        through set_cell_bounding --, the composite-object pair factors come from the factor type map."""
        p = self.program
        if p.no_cells or p.nodes_per_root < 2 or p.cell_level != 1 or not 0 <= child < p.nodes_per_root:
            raise ValueError("leaf cells need composite objects (set_composite, set_molecules) and a cell system")
        p.cell_child = child + 1
        p.inter_bound_offset = offset
        p.inter_bound_max_displacement = max_displacement
        return self

    def set_cell_bounding(self, potential, bounds, use_charge=False, target_charge=1.0):
        """Far field through TwoLeafUnitCellBoundingPotentialEventHandler: bounds[n_cells][dimension][2] holds
        (upper bound, -lower bound) of the derivative per relative cell (CellBoundingPotential._derivative_bounds)."""
        p = self.program
        p.veto_enabled = abi.FAR_CELL_BOUNDING
        p.veto_potential = potential
        p.veto_use_charge = int(use_charge)
        p.veto_target_charge = target_charge
        vt = abi.EcmcVetoTables()
        array = np.ascontiguousarray(np.nan_to_num(bounds, nan=0.0), dtype=np.float64)
        self._keep += [array, vt]
        vt.bounds = array.ctypes.data_as(C.POINTER(C.c_double))
        p.veto_tables = C.pointer(vt)
        self.tables = {"bounds": array}
        return self

    def set_veto(self, potential, tables, use_charge=False, target_charge=1.0):
        p = self.program
        p.veto_enabled = 1
        p.veto_potential = potential
        p.veto_use_charge = int(use_charge)
        p.veto_target_charge = target_charge
        vt = abi.EcmcVetoTables()
        for name, dst in (("upper", vt.upper), ("lower", vt.lower)):
            for d in range(p.dimension):
                t = tables[name][d]
                a = np.ascontiguousarray(t["cell_a"], dtype=np.int32)
                b = np.ascontiguousarray(t["cell_b"], dtype=np.int32)
                r = np.ascontiguousarray(t["rate_a"], dtype=np.float64)
                self._keep += [a, b, r]
                dst[d].n_entries = len(a)
                dst[d].cell_a = a.ctypes.data_as(C.POINTER(C.c_int32))
                dst[d].cell_b = b.ctypes.data_as(C.POINTER(C.c_int32))
                dst[d].rate_a = r.ctypes.data_as(C.POINTER(C.c_double))
                dst[d].total_rate = t["total_rate"]
                dst[d].mean_rate = t["mean_rate"]
        bounds = np.ascontiguousarray(np.nan_to_num(tables["bounds"], nan=0.0), dtype=np.float64)
        self._keep.append(bounds)
        vt.bounds = bounds.ctypes.data_as(C.POINTER(C.c_double))
        self._keep.append(vt)
        p.veto_tables = C.pointer(vt)
        self.tables = tables
        return self
