"""ctypes binding of libecmc_b200.so (include/ecmc.h): the thin host side of the B200 event-chain engine.

`Engine` owns one EcmcHandle = the chains of one GPU. Every method maps 1:1 to a C entry point; numpy arrays
are the host buffers. There is no CPU implementation behind this module: if the library or a CUDA device is
missing, calls raise.
"""
import ctypes as C
import os

import numpy as np

from jellyfysh_b200 import abi

_HERE = os.path.dirname(os.path.abspath(__file__))
# JELLYFYSH_B200_LIBRARY selects another build of the same library (kernel tuning experiments)
LIBRARY_PATH = os.environ.get("JELLYFYSH_B200_LIBRARY", os.path.join(_HERE, "libecmc_b200.so"))
_LIB = None

INF = float("inf")


class EcmcError(RuntimeError):
    """A libecmc_b200 call returned a non-zero status."""

    def __init__(self, status, message):
        super().__init__(f"libecmc_b200 status {status}: {message}")
        self.status = status


def library() -> C.CDLL:
    """Load libecmc_b200.so (built by jellyfysh_b200.build / __graft_entry__.build). No fallback."""
    global _LIB
    if _LIB is not None:
        return _LIB
    if not os.path.exists(LIBRARY_PATH):
        raise ImportError(f"{LIBRARY_PATH} is missing: run `python -m jellyfysh_b200.build` (nvcc, sm_100a). "
                          "The ECMC hot path has no CPU fallback.")
    lib = C.CDLL(LIBRARY_PATH)
    vp, d, i32, i64, u32, u64, sz = C.c_void_p, C.c_double, C.c_int32, C.c_int64, C.c_uint32, C.c_uint64, C.c_size_t
    pot = C.POINTER(abi.EcmcPotential)
    stats = C.POINTER(abi.EcmcStats)
    lib.ecmc_abi_version.restype = C.c_int
    lib.ecmc_create.argtypes = [C.POINTER(abi.EcmcProgram), C.c_int, C.c_int, C.POINTER(vp)]
    lib.ecmc_destroy.argtypes = [vp]
    lib.ecmc_destroy.restype = None
    lib.ecmc_last_error.argtypes = [vp]
    lib.ecmc_last_error.restype = C.c_char_p
    lib.ecmc_upload_positions.argtypes = [vp, vp, vp]
    lib.ecmc_download_positions.argtypes = [vp, vp]
    lib.ecmc_download_chain.argtypes = [vp, C.c_int, vp, vp, vp]
    lib.ecmc_upload_roots.argtypes = [vp, vp]
    lib.ecmc_download_roots.argtypes = [vp, vp]
    lib.ecmc_start.argtypes = [vp, vp, u32]
    lib.ecmc_upload_chain_states.argtypes = [vp, vp]
    lib.ecmc_download_chain_states.argtypes = [vp, vp]
    lib.ecmc_upload_cells.argtypes = [vp, vp, vp, vp]
    lib.ecmc_download_cells.argtypes = [vp, vp, vp, vp]
    lib.ecmc_run.argtypes = [vp, d, d, i64]
    lib.ecmc_sync.argtypes = [vp, stats]
    lib.ecmc_run_recorded.argtypes = [vp, d, d, i64, vp, i32, stats]
    lib.ecmc_run_from_host.argtypes = [vp, vp, vp, u32, d, d, i64, vp, stats]
    lib.ecmc_submit_from_host.argtypes = [vp, vp, vp, u32, d, d, i64, vp]
    lib.ecmc_submit_from_host_sparse.argtypes = [vp, vp, vp, u32, d, d, i64, vp]
    lib.ecmc_host_bytes_written.argtypes = [vp]
    lib.ecmc_host_bytes_written.restype = C.c_uint64
    lib.ecmc_host_alloc.argtypes = [C.c_size_t, C.POINTER(vp)]
    lib.ecmc_host_free.argtypes = [vp]
    lib.ecmc_wait.argtypes = [vp, stats]
    lib.ecmc_separation_histogram.argtypes = [vp, i32, d, d, vp]
    lib.ecmc_separation_histogram_subset.argtypes = [vp, i32, i32, i32, d, d, vp]
    lib.ecmc_polarization.argtypes = [vp, vp, vp]
    lib.ecmc_bond_histograms.argtypes = [vp, i32, d, d, d, d, vp, vp]
    lib.ecmc_set_option.argtypes = [vp, C.c_int, C.c_int]
    lib.ecmc_stream.argtypes = [vp]
    lib.ecmc_stream.restype = vp
    lib.ecmc_kernel_seconds.argtypes = [vp]
    lib.ecmc_kernel_seconds.restype = d
    lib.ecmc_kernel_launches.argtypes = [vp]
    lib.ecmc_kernel_launches.restype = u64
    lib.ecmc_kernel_name.argtypes = [vp, C.c_int]
    lib.ecmc_kernel_name.restype = C.c_char_p
    lib.ecmc_potential_derivative.argtypes = [pot, C.c_int, d, vp, sz, vp, vp, vp, C.c_int]
    lib.ecmc_potential_displacement.argtypes = [pot, C.c_int, d, vp, sz, vp, vp, vp, vp, C.c_int]
    lib.ecmc_random_doubles.argtypes = [u32, u32, u64, u32, u32, sz, vp]
    lib.ecmc_random_doubles.restype = None
    lib.ecmc_random_words.argtypes = [u32, u32, u64, u32, u32, sz, vp]
    lib.ecmc_random_words.restype = None
    if lib.ecmc_abi_version() != abi.ECMC_ABI_VERSION:
        raise ImportError("libecmc_b200.so was built for another ABI version: rebuild it")
    _LIB = lib
    return lib


def _ptr(array):
    return None if array is None else array.ctypes.data


def _f64(array, shape=None):
    out = np.ascontiguousarray(array, dtype=np.float64)
    return out if shape is None else out.reshape(shape)


class Engine:
    """`n_chains` independent Markov chains of one program on one CUDA device."""

    def __init__(self, builder, n_chains=1, device=0):
        self._lib = library()
        self._builder = builder  # keeps the program's arrays alive during create
        program = builder.program
        self.dimension = int(program.dimension)
        self.n_particles = int(program.n_particles)
        self.n_cells = int(np.prod([program.cells_per_side[d] for d in range(self.dimension)]))
        self.max_occupants = int(program.max_occupants)
        self.max_surplus = max(int(program.max_surplus), 1)
        self.nodes_per_root = max(int(program.nodes_per_root), 1)
        self.n_chains = int(n_chains)
        self.device = int(device)
        handle = C.c_void_p()
        status = self._lib.ecmc_create(C.byref(program), device, n_chains, C.byref(handle))
        if status != abi.ECMC_OK:
            raise EcmcError(status, self._lib.ecmc_last_error(None).decode())
        self._h = handle

    def close(self):
        if getattr(self, "_h", None):
            self._lib.ecmc_destroy(self._h)
            self._h = None

    def __del__(self):
        self.close()

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def _check(self, status):
        if status != abi.ECMC_OK:
            raise EcmcError(status, self._lib.ecmc_last_error(self._h).decode())

    # ---- state -----------------------------------------------------------------------------------------
    def upload_positions(self, positions, charges=None):
        pos = _f64(positions, (self.n_chains, self.n_particles, self.dimension))
        ch = None if charges is None else _f64(charges, (self.n_chains, self.n_particles))
        self._charges = None if ch is None else ch.copy()  # part of a checkpoint (the device records are never read back)
        self._check(self._lib.ecmc_upload_positions(self._h, _ptr(pos), _ptr(ch)))
        self._check(self._lib.ecmc_sync(self._h, None))  # the host buffers may go away after this call

    def download_positions(self):
        out = np.empty((self.n_chains, self.n_particles, self.dimension), dtype=np.float64)
        self._check(self._lib.ecmc_download_positions(self._h, _ptr(out)))
        return out

    def download_chain(self, chain):
        """ecmc_download_chain: (positions[N][D], roots[N / npr][D] or None, lifting state record) of one chain."""
        positions = np.empty((self.n_particles, self.dimension), dtype=np.float64)
        roots = (np.empty((self.n_particles // self.nodes_per_root, self.dimension), dtype=np.float64)
                 if self.nodes_per_root > 1 else None)
        state = np.zeros(1, dtype=abi.chain_state_dtype())
        self._check(self._lib.ecmc_download_chain(self._h, int(chain), _ptr(positions), _ptr(roots), _ptr(state)))
        return positions, roots, state[0]

    def upload_roots(self, roots):
        """Root-unit positions of composite objects, [n_chains][n_particles / nodes_per_root][dimension]."""
        arr = _f64(roots, (self.n_chains, self.n_particles // self.nodes_per_root, self.dimension))
        self._check(self._lib.ecmc_upload_roots(self._h, _ptr(arr)))

    def download_roots(self):
        out = np.empty((self.n_chains, self.n_particles // self.nodes_per_root, self.dimension), dtype=np.float64)
        self._check(self._lib.ecmc_download_roots(self._h, _ptr(out)))
        return out

    def start(self, streams=None, first_stream=0):
        arr = None if streams is None else np.ascontiguousarray(streams, dtype=np.uint32).reshape(self.n_chains)
        self._check(self._lib.ecmc_start(self._h, _ptr(arr), int(first_stream)))
        self._check(self._lib.ecmc_sync(self._h, None))

    def chain_states(self):
        out = np.zeros(self.n_chains, dtype=abi.chain_state_dtype())
        self._check(self._lib.ecmc_download_chain_states(self._h, _ptr(out)))
        return out

    def set_chain_states(self, states):
        arr = np.ascontiguousarray(states, dtype=abi.chain_state_dtype()).reshape(self.n_chains)
        self._check(self._lib.ecmc_upload_chain_states(self._h, _ptr(arr)))

    def cells(self):
        """(occupants[n_chains][n_cells][max_occupants], list of surplus arrays per chain)."""
        occ = np.empty((self.n_chains, self.n_cells, self.max_occupants), dtype=np.int32)
        sur = np.empty((self.n_chains, self.max_surplus), dtype=np.int32)
        n = np.empty(self.n_chains, dtype=np.int32)
        self._check(self._lib.ecmc_download_cells(self._h, _ptr(occ), _ptr(sur), _ptr(n)))
        return occ, [sur[c, :n[c]].copy() for c in range(self.n_chains)]

    def set_cells(self, occupants, surplus_lists):
        occ = np.ascontiguousarray(occupants, dtype=np.int32).reshape(self.n_chains, self.n_cells, self.max_occupants)
        sur = np.full((self.n_chains, self.max_surplus), -1, dtype=np.int32)
        n = np.zeros(self.n_chains, dtype=np.int32)
        for c, items in enumerate(surplus_lists):
            n[c] = len(items)
            sur[c, :len(items)] = items
        self._check(self._lib.ecmc_upload_cells(self._h, _ptr(occ), _ptr(sur), _ptr(n)))

    OPTION_BATCHED_EVENTS, OPTION_PRUNE_CANDIDATES, OPTION_LANES_PER_EVENT, OPTION_CHAIN_BLOCKS = 1, 2, 3, 4
    OPTION_FUSED_HOST_STEPS, OPTION_CONTINUE_HOST_STEPS = 5, 6

    def set_option(self, option, value):
        """ecmc_set_option: how the device schedules the events (batched speculative evaluation, candidate pruning)."""
        self._check(self._lib.ecmc_set_option(self._h, int(option), int(value)))

    # ---- the hot path ----------------------------------------------------------------------------------
    def run(self, until=(INF, INF), max_events=0):
        """Asynchronous: advance every chain to the time `until` (quotient, remainder) or by max_events events."""
        self._check(self._lib.ecmc_run(self._h, float(until[0]), float(until[1]), int(max_events)))

    def sync(self):
        stats = abi.EcmcStats()
        self._check(self._lib.ecmc_sync(self._h, C.byref(stats)))
        return stats.as_dict()

    def run_recorded(self, until=(INF, INF), max_events=0, records_per_chain=1):
        """Synchronous run that also returns the first records_per_chain events of every chain."""
        rec = np.zeros((self.n_chains, records_per_chain), dtype=abi.record_dtype())
        stats = abi.EcmcStats()
        self._check(self._lib.ecmc_run_recorded(self._h, float(until[0]), float(until[1]), int(max_events), _ptr(rec),
                                                int(records_per_chain), C.byref(stats)))
        return rec, stats.as_dict()

    def run_from_host(self, positions, charges=None, first_stream=0, until=(INF, INF), max_events=0, out=None):
        """Host buffers in, host buffers out: upload -> start -> run -> download in one C call."""
        pos = _f64(positions, (self.n_chains, self.n_particles, self.dimension))
        ch = None if charges is None else _f64(charges, (self.n_chains, self.n_particles))
        if out is None:
            out = np.empty_like(pos)
        stats = abi.EcmcStats()
        self._check(self._lib.ecmc_run_from_host(self._h, _ptr(pos), _ptr(ch), int(first_stream), float(until[0]),
                                                 float(until[1]), int(max_events), _ptr(out), C.byref(stats)))
        return out, stats.as_dict()

    def submit_from_host(self, positions, charges=None, first_stream=0, until=(INF, INF), max_events=0, out=None,
                         sparse=False):
        """ecmc_submit_from_host: the step of run_from_host, enqueued only. sparse: ecmc_submit_from_host_sparse -- `out`
        (normally the same array as `positions`, allocated by pinned_array) already holds the input configuration and only
        the coordinates of the particles that moved are written into it, by the device. `positions`, `charges` and `out` must be
        page-locked C-contiguous float64 arrays of the engine's shape (they are used in place, nothing is copied here) and
        stay alive until wait(); steps submitted back to back may chain through their buffers (out of one = positions of
        the next)."""
        for array, shape in ((positions, (self.n_chains, self.n_particles, self.dimension)),
                             (charges, (self.n_chains, self.n_particles)), (out, (self.n_chains, self.n_particles, self.dimension))):
            if array is not None and (array.dtype != np.float64 or not array.flags["C_CONTIGUOUS"] or array.shape != shape):
                raise ValueError("submit_from_host needs C-contiguous float64 arrays of shape {0}".format(shape))
        call = self._lib.ecmc_submit_from_host_sparse if sparse else self._lib.ecmc_submit_from_host
        self._check(call(self._h, _ptr(positions), _ptr(charges), int(first_stream), float(until[0]), float(until[1]),
                         int(max_events), _ptr(out)))

    @property
    def host_bytes_written(self):
        """ecmc_host_bytes_written: bytes the device wrote into host buffers by sparse write-backs (valid after wait())."""
        return int(self._lib.ecmc_host_bytes_written(self._h))

    def wait(self):
        """ecmc_wait: block until all submitted steps are complete; their summed counters."""
        stats = abi.EcmcStats()
        self._check(self._lib.ecmc_wait(self._h, C.byref(stats)))
        return stats.as_dict()

    def separation_histogram(self, n_bins, r_min, r_max, out=None, first=0, stride=1):
        """Add the pair-separation counts of the current configuration of all chains to `out` (uint64[n_bins]);
        first / stride select every stride-th particle (e.g. the oxygens of water: first=1, stride=3)."""
        if out is None:
            out = np.zeros(n_bins, dtype=np.uint64)
        self._check(self._lib.ecmc_separation_histogram_subset(self._h, int(first), int(stride), int(n_bins),
                                                               float(r_min), float(r_max), _ptr(out)))
        return out

    def polarization(self, charges=None):
        """ecmc_polarization: [n_chains][dimension], sum of charge x closest leaf position over all leaves of a chain;
        charges[n_particles] (the same in every chain) or None for the uploaded ones."""
        out = np.empty((self.n_chains, self.dimension), dtype=np.float64)
        ch = None if charges is None else _f64(charges, (self.n_particles,))
        self._check(self._lib.ecmc_polarization(self._h, _ptr(ch), _ptr(out)))
        return out

    def bond_histograms(self, n_bins, length_range, angle_range, out=None):
        """Add the bond lengths / bond angles of all three-leaf objects to (lengths, angles) uint64[n_bins] histograms."""
        lengths, angles = out if out is not None else (np.zeros(n_bins, dtype=np.uint64), np.zeros(n_bins, dtype=np.uint64))
        self._check(self._lib.ecmc_bond_histograms(self._h, int(n_bins), float(length_range[0]), float(length_range[1]),
                                                   float(angle_range[0]), float(angle_range[1]), _ptr(lengths), _ptr(angles)))
        return lengths, angles

    # ---- checkpoint / resume (the role of DumpingOutputHandler + resume.py, jellyfysh/resume.py) ----
    def save_checkpoint(self, path):
        """Everything that defines the future of all chains -- leaf (and root) positions, cell occupancy, surplus lists,
        lifting state incl. the random-stream counters and kept candidates -- as one .npz file. A run resumed from it
        commits bit for bit the events the uninterrupted run commits."""
        self._check(self._lib.ecmc_sync(self._h, None))
        occupants, surplus = self.cells()
        n_surplus = np.array([len(items) for items in surplus], dtype=np.int32)
        padded = np.full((self.n_chains, self.max_surplus), -1, dtype=np.int32)
        for chain, items in enumerate(surplus):
            padded[chain, :len(items)] = items
        arrays = {"positions": self.download_positions(), "chain_states": self.chain_states(), "occupants": occupants,
                  "surplus": padded, "n_surplus": n_surplus,
                  "layout": np.array([self.n_chains, self.n_particles, self.dimension, self.n_cells, self.max_occupants,
                                      self.max_surplus, self.nodes_per_root], dtype=np.int64),
                  "fingerprint": np.frombuffer(self._builder.fingerprint(), dtype=np.uint8)}
        if getattr(self, "_charges", None) is not None:
            arrays["charges"] = self._charges
        if self.nodes_per_root > 1:
            arrays["roots"] = self.download_roots()
        np.savez_compressed(path, **arrays)

    def load_checkpoint(self, path, charges=None):
        """Restore a state written by save_checkpoint into an engine of the same program and number of chains. The
        checkpoint carries the charges it was started with (`charges` overrides them) and a fingerprint of the program
        (seed, potentials, cell system, handler kinds): a dump of another program is refused."""
        with np.load(path) as data:
            layout = [self.n_chains, self.n_particles, self.dimension, self.n_cells, self.max_occupants, self.max_surplus,
                      self.nodes_per_root]
            if data["layout"].tolist() != layout:
                raise ValueError("checkpoint layout {0} does not match this engine {1}".format(data["layout"].tolist(), layout))
            if "fingerprint" in data and data["fingerprint"].tobytes() != self._builder.fingerprint():
                raise ValueError("the checkpoint was written by a different program (seed, potentials or handlers differ)")
            if charges is None and "charges" in data:
                charges = data["charges"]
            program = self._builder.program
            if charges is None and (program.pair_use_charge or program.veto_use_charge):
                raise ValueError("the program uses charges and the checkpoint holds none: pass charges=")
            self.upload_positions(data["positions"], charges)
            if self.nodes_per_root > 1:
                self.upload_roots(data["roots"])
            self.set_cells(data["occupants"], [row[:n] for row, n in zip(data["surplus"], data["n_surplus"])])
            self.set_chain_states(data["chain_states"])

    @property
    def cuda_stream(self):
        return self._lib.ecmc_stream(self._h)

    @property
    def kernel_seconds(self):
        return float(self._lib.ecmc_kernel_seconds(self._h))

    @property
    def kernel_launches(self):
        return int(self._lib.ecmc_kernel_launches(self._h))

    def kernel_name(self, record=False):
        """ecmc_kernel_name: the event kernel ecmc_run (or ecmc_run_recorded) launches for this program and options."""
        return self._lib.ecmc_kernel_name(self._h, int(bool(record))).decode()


class _PinnedBlock:
    """Owner of one ecmc_host_alloc block; frees it when the last array over it is gone."""

    def __init__(self, n_bytes):
        self.pointer = C.c_void_p()
        if library().ecmc_host_alloc(int(n_bytes), C.byref(self.pointer)) != 0:
            raise MemoryError("ecmc_host_alloc({0}) failed".format(n_bytes))
        self.buffer = (C.c_char * int(n_bytes)).from_address(self.pointer.value)

    def __del__(self):
        if self.pointer:
            library().ecmc_host_free(self.pointer)
            self.pointer = C.c_void_p()


def pinned_array(shape, dtype=np.float64):
    """numpy array over page-locked, device-addressable host memory (ecmc_host_alloc): the buffers of the *_from_host
    calls, required by the sparse write-back."""
    n_bytes = int(np.prod(shape)) * np.dtype(dtype).itemsize
    block = _PinnedBlock(max(n_bytes, 1))
    array = np.frombuffer(block.buffer, dtype=dtype, count=int(np.prod(shape))).reshape(shape)
    array.flags.writeable = True
    _PINNED_OWNERS[id(block)] = block  # np.frombuffer keeps block.buffer alive, not the block: keep the owner too
    import weakref
    weakref.finalize(array, _PINNED_OWNERS.pop, id(block), None)
    return array


_PINNED_OWNERS = {}


# ---- batched potential arithmetic and the random stream -----------------------------------------------------
def _velocity(direction_or_velocity, dimension, speed=1.0):
    if np.isscalar(direction_or_velocity):
        velocity = np.zeros(3)
        velocity[int(direction_or_velocity)] = float(speed)
        return velocity
    velocity = np.zeros(3)
    velocity[:dimension] = np.asarray(direction_or_velocity, dtype=np.float64)[:dimension]
    return velocity


def potential_derivative(potential, dimension, system_length, velocity, separations, charges=None, device=0):
    """Potential.derivative of the reference (jellyfysh/potential/potential.py:154-181) for n separations."""
    lib = library()
    seps = _f64(separations).reshape(-1, dimension)
    out = np.empty(len(seps), dtype=np.float64)
    ch = None if charges is None else _f64(charges, (len(seps), 2))
    vel = _velocity(velocity, dimension)
    status = lib.ecmc_potential_derivative(C.byref(potential), dimension, float(system_length), _ptr(vel), len(seps),
                                           _ptr(seps), _ptr(ch), _ptr(out), device)
    if status != abi.ECMC_OK:
        raise EcmcError(status, lib.ecmc_last_error(None).decode())
    return out


def potential_displacement(potential, dimension, system_length, velocity, separations, charges=None,
                           potential_changes=None, device=0):
    """InvertiblePotential.displacement of the reference (potential.py:218-301), a time, for n inputs."""
    lib = library()
    seps = _f64(separations).reshape(-1, dimension)
    out = np.empty(len(seps), dtype=np.float64)
    ch = None if charges is None else _f64(charges, (len(seps), 2))
    du = None if potential_changes is None else _f64(potential_changes, (len(seps),))
    vel = _velocity(velocity, dimension)
    status = lib.ecmc_potential_displacement(C.byref(potential), dimension, float(system_length), _ptr(vel), len(seps),
                                             _ptr(seps), _ptr(ch), _ptr(du), _ptr(out), device)
    if status != abi.ECMC_OK:
        raise EcmcError(status, lib.ecmc_last_error(None).decode())
    return out


def random_doubles(seed, stream, event, slot, first, n):
    out = np.empty(n, dtype=np.float64)
    library().ecmc_random_doubles(seed, stream, event, slot, first, n, _ptr(out))
    return out


def random_words(seed, stream, event, slot, first, n):
    out = np.empty(n, dtype=np.uint32)
    library().ecmc_random_words(seed, stream, event, slot, first, n, _ptr(out))
    return out
