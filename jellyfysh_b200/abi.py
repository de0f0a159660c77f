"""ctypes mirror of include/ecmc.h (the C ABI of libecmc_b200.so). Plain data only, no computation."""
import ctypes as C

ECMC_ABI_VERSION = 6
ECMC_MAX_DIM = 3
ECMC_MAX_BONDS = 4
ECMC_MAX_INTER_FACTORS = 4

ECMC_OK = 0
ECMC_ERR_INVALID = -1
ECMC_ERR_CUDA = -2
ECMC_ERR_STATE = -3
ECMC_ERR_CAPACITY = -4

POT_NONE = 0
POT_INVERSE_POWER = 1
POT_LENNARD_JONES = 2
POT_DISPLACED_EVEN_POWER = 3
POT_HARD_SPHERE = 4
POT_HARD_DIPOLE = 5
POT_MERGED_IMAGE_COULOMB = 6
POT_INVERSE_POWER_COULOMB_BOUNDING = 7
POT_BENDING = 8

PAIR_NONE = 0
PAIR_TWO_LEAF_UNIT = 1
PAIR_TWO_LEAF_UNIT_BOUNDING = 2
PAIR_TWO_COMPOSITE_SUMMED_BOUNDING = 3

LIFTING_NONE = 0
LIFTING_INSIDE_FIRST = 1
LIFTING_OUTSIDE_FIRST = 2
LIFTING_RATIO = 3

EVENT_NONE = 0
EVENT_PAIR = 1
EVENT_CELL_VETO = 2
EVENT_CELL_BOUNDARY = 3
EVENT_END_OF_CHAIN = 4
EVENT_CELL_BOUNDING = 5
EVENT_BOND = 6
EVENT_FACTOR_PAIR = 7
EVENT_BENDING = 8
EVENT_SWITCH = 9

FAR_NONE = 0
FAR_CELL_VETO = 1
FAR_CELL_BOUNDING = 2
EVENT_NAMES = {EVENT_NONE: "none", EVENT_PAIR: "pair", EVENT_CELL_VETO: "cell_veto",
               EVENT_CELL_BOUNDARY: "cell_boundary", EVENT_END_OF_CHAIN: "end_of_chain",
               EVENT_CELL_BOUNDING: "cell_bounding", EVENT_BOND: "bond", EVENT_FACTOR_PAIR: "factor_pair",
               EVENT_BENDING: "bending", EVENT_SWITCH: "switch"}

SLOT_PAIR_TIME = 1
SLOT_VETO_TIME = 2
SLOT_VETO_CHOICE = 3
SLOT_CONFIRM = 4
SLOT_END_OF_CHAIN = 5
SLOT_LIFTING = 6
SLOT_FACTOR_TIME = 7
SLOT_BENDING_TIME = 8
SLOT_SWITCH = 9


def slot(kind: int, index: int = 0) -> int:
    """ECMC_SLOT(kind, index) of include/ecmc.h."""
    return ((kind << 24) | (index & 0xFFFFFF)) & 0xFFFFFFFF


class EcmcPotential(C.Structure):
    _fields_ = [("kind", C.c_int32), ("reserved", C.c_int32), ("params", C.c_double * 6)]

    @classmethod
    def make(cls, kind: int, *params: float) -> "EcmcPotential":
        p = cls()
        p.kind = kind
        for i, value in enumerate(params):
            p.params[i] = float(value)
        return p


class EcmcWalkerTable(C.Structure):
    _fields_ = [("n_entries", C.c_int32), ("reserved", C.c_int32),
                ("cell_a", C.POINTER(C.c_int32)), ("cell_b", C.POINTER(C.c_int32)),
                ("rate_a", C.POINTER(C.c_double)),
                ("total_rate", C.c_double), ("mean_rate", C.c_double)]


class EcmcVetoTables(C.Structure):
    _fields_ = [("upper", EcmcWalkerTable * ECMC_MAX_DIM), ("lower", EcmcWalkerTable * ECMC_MAX_DIM),
                ("bounds", C.POINTER(C.c_double))]


class EcmcProgram(C.Structure):
    _fields_ = [("abi_version", C.c_int32), ("dimension", C.c_int32), ("n_particles", C.c_int32),
                ("no_cells", C.c_int32),
                ("system_length", C.c_double), ("beta", C.c_double),
                ("cells_per_side", C.c_int32 * ECMC_MAX_DIM), ("neighbor_layers", C.c_int32),
                ("max_occupants", C.c_int32), ("max_surplus", C.c_int32),
                ("pair_handler", C.c_int32), ("pair_use_charge", C.c_int32),
                ("pair_potential", EcmcPotential), ("pair_bounding_potential", EcmcPotential),
                ("veto_enabled", C.c_int32), ("veto_use_charge", C.c_int32),
                ("veto_potential", EcmcPotential), ("veto_target_charge", C.c_double),
                ("veto_tables", C.POINTER(EcmcVetoTables)),
                ("chain_time", C.c_double), ("speed", C.c_double),
                ("initial_direction", C.c_int32), ("initial_active", C.c_int32),
                ("seed", C.c_uint32), ("reserved1", C.c_uint32),
                ("nodes_per_root", C.c_int32), ("n_bonds", C.c_int32),
                ("bonds", (C.c_int32 * 2) * ECMC_MAX_BONDS), ("bond_potential", EcmcPotential),
                ("cell_level", C.c_int32), ("composite_lifting", C.c_int32),
                ("n_inter_factors", C.c_int32), ("inter_factors", (C.c_int32 * 2) * ECMC_MAX_INTER_FACTORS),
                ("inter_potential", EcmcPotential),
                ("bending_enabled", C.c_int32), ("bending_lifting", C.c_int32),
                ("bending_children", C.c_int32 * 3), ("bending_separations", C.c_int32 * 4), ("boundary_keeps_factors", C.c_int32),
                ("bending_potential", EcmcPotential), ("bending_offset", C.c_double),
                ("bending_max_displacement", C.c_double),
                ("eoc_sequential", C.c_int32), ("root_mode", C.c_int32), ("eoc_cos", C.c_double), ("eoc_sin", C.c_double),
                ("switch_chain_length", C.c_double * 2),
                ("cell_child", C.c_int32), ("reserved3", C.c_int32), ("inter_bound_offset", C.c_double),
                ("inter_bound_max_displacement", C.c_double)]


class EcmcChainState(C.Structure):
    _fields_ = [("active", C.c_int32), ("direction", C.c_int32),
                ("time_q", C.c_double), ("time_r", C.c_double),
                ("eoc_q", C.c_double), ("eoc_r", C.c_double),
                ("eoc_next_active", C.c_int32), ("active_cell", C.c_int32),
                ("event_counter", C.c_uint64),
                ("stream", C.c_uint32), ("pending_kind", C.c_int32),
                ("pending_target", C.c_int32), ("mode", C.c_int32),
                ("pending_q", C.c_double), ("pending_r", C.c_double), ("pending_rate", C.c_double),
                ("pending_position", C.c_double), ("pending_stamp_q", C.c_double), ("pending_stamp_r", C.c_double),
                ("pending_root_position", C.c_double),
                ("kept_kind", C.c_int32), ("kept_target", C.c_int32), ("kept_q", C.c_double), ("kept_r", C.c_double),
                ("kept_rate", C.c_double), ("kept_position", C.c_double), ("kept_root_position", C.c_double),
                ("kept_stamp_q", C.c_double), ("kept_stamp_r", C.c_double),
                ("velocity", C.c_double * 2), ("root_velocity", C.c_double * 2),
                ("pending_position_y", C.c_double), ("pending_root_position_y", C.c_double),
                ("switch_q", C.c_double), ("switch_r", C.c_double), ("eoc_last_q", C.c_double), ("eoc_last_r", C.c_double)]


class EcmcEventRecord(C.Structure):
    _fields_ = [("kind", C.c_int32), ("target", C.c_int32), ("target_cell", C.c_int32), ("accepted", C.c_int32),
                ("n_candidates", C.c_int32), ("new_active", C.c_int32), ("new_direction", C.c_int32),
                ("mode", C.c_int32),
                ("time_q", C.c_double), ("time_r", C.c_double), ("active_pos", C.c_double * ECMC_MAX_DIM)]


class EcmcStats(C.Structure):
    _fields_ = [("events", C.c_uint64), ("pair_events", C.c_uint64), ("veto_events", C.c_uint64),
                ("veto_accepted", C.c_uint64), ("boundary_events", C.c_uint64),
                ("end_of_chain_events", C.c_uint64), ("candidates", C.c_uint64),
                ("bound_violations", C.c_uint64), ("capacity_errors", C.c_uint64),
                ("bond_events", C.c_uint64), ("factor_pair_events", C.c_uint64), ("pair_targets", C.c_uint64)]

    def as_dict(self):
        return {name: int(getattr(self, name)) for name, _ in self._fields_ if name != "reserved"}


# numpy dtypes with the same memory layout (records / chain states are exchanged as arrays)
def record_dtype():
    import numpy as np
    return np.dtype([("kind", "<i4"), ("target", "<i4"), ("target_cell", "<i4"), ("accepted", "<i4"),
                     ("n_candidates", "<i4"), ("new_active", "<i4"), ("new_direction", "<i4"), ("mode", "<i4"),
                     ("time_q", "<f8"), ("time_r", "<f8"), ("active_pos", "<f8", (ECMC_MAX_DIM,))])


def chain_state_dtype():
    import numpy as np
    return np.dtype([("active", "<i4"), ("direction", "<i4"), ("time_q", "<f8"), ("time_r", "<f8"),
                     ("eoc_q", "<f8"), ("eoc_r", "<f8"), ("eoc_next_active", "<i4"), ("active_cell", "<i4"),
                     ("event_counter", "<u8"), ("stream", "<u4"), ("pending_kind", "<i4"),
                     ("pending_target", "<i4"), ("mode", "<i4"),
                     ("pending_q", "<f8"), ("pending_r", "<f8"), ("pending_rate", "<f8"),
                     ("pending_position", "<f8"), ("pending_stamp_q", "<f8"), ("pending_stamp_r", "<f8"),
                     ("pending_root_position", "<f8"),
                     ("kept_kind", "<i4"), ("kept_target", "<i4"), ("kept_q", "<f8"), ("kept_r", "<f8"),
                     ("kept_rate", "<f8"), ("kept_position", "<f8"), ("kept_root_position", "<f8"),
                     ("kept_stamp_q", "<f8"), ("kept_stamp_r", "<f8"),
                     ("velocity", "<f8", (2,)), ("root_velocity", "<f8", (2,)),
                     ("pending_position_y", "<f8"), ("pending_root_position_y", "<f8"),
                     ("switch_q", "<f8"), ("switch_r", "<f8"), ("eoc_last_q", "<f8"), ("eoc_last_r", "<f8")])


assert C.sizeof(EcmcEventRecord) == 72
assert C.sizeof(EcmcChainState) == 272
