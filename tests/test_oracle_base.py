"""Pin the oracle's Time arithmetic, periodic boundaries and cell geometry to outputs of the running reference
(tests/golden/base.npz; reference: jellyfysh/base/time.py, jellyfysh/setting/hypercubic_setting.py,
jellyfysh/activator/internal_state/cell_occupancy/cells/cuboid_periodic_cells.py)."""
import numpy as np
import pytest

import kat_replay as kr


def test_time_arithmetic_bit_exact(oracle):
    g = kr.load_npz("base")
    for q, r, dt, added, sub, ff in zip(g["time_q"], g["time_r"], g["time_dt"], g["time_added"], g["time_sub"],
                                        g["time_from_float"]):
        assert oracle.time_add(q, r, dt) == (added[0], added[1])
        assert oracle.time_sub(added[0], added[1], q, r) == sub
        assert oracle.time_from_float(q + r) == (ff[0], ff[1])
    assert oracle.time_add(3.0, 0.5, float("inf")) == (float("inf"), float("inf"))


@pytest.mark.parametrize("geo", range(6))
def test_cells_and_boundaries_bit_exact(oracle, geo):
    g = kr.load_npz("base")
    params = g[f"geo{geo}_params"]
    dim, length, nl = int(params[0]), float(params[1]), int(params[2])
    cps = [int(c) for c in params[3:3 + dim]]
    cmin, cmax = oracle.cells_geometry(dim, cps, length)
    limit = len(g[f"geo{geo}_cell_min"])
    assert np.array_equal(cmin[:limit], g[f"geo{geo}_cell_min"])
    assert np.array_equal(cmax[:limit], g[f"geo{geo}_cell_max"])
    for pos, cell in zip(g[f"geo{geo}_pos"], g[f"geo{geo}_pos_cell"]):
        assert oracle.position_to_cell(dim, cps, length, pos) == cell
    for k, probe in enumerate(g[f"geo{geo}_probe"]):
        assert sorted(oracle.nearby_cells(dim, cps, nl, length, int(probe))) == list(g[f"geo{geo}_nearby"][k])
        rel = int(g[f"geo{geo}_rel"][k])
        assert oracle.cells_translate(dim, cps, length, int(probe), rel) == g[f"geo{geo}_translate"][k]
        assert oracle.cells_relative(dim, cps, length, int(probe), rel) == g[f"geo{geo}_relative"][k]
    lib = oracle.lib()
    for s, so, po in zip(g[f"geo{geo}_sep_in"], g[f"geo{geo}_sep_out"], g[f"geo{geo}_pos_out"]):
        assert lib.orc_correct_separation_entry(s, length) == so
        assert lib.orc_correct_position_entry(s, length) == po


def test_time_order_matches_the_compiled_reference_heap(oracle):
    """The argmin the Scheduler would pop: the reference's own heap (oracle/_ref/libref_heap.so, compiled in place from
    jellyfysh/scheduler/heap_scheduler/heap.c) pops random Times -- equal quotients, tiny and huge remainders, stale
    entries removed lazily through the validity callback -- in exactly the order of the oracle's comparison."""
    import ctypes as C
    import functools
    import os
    path = os.path.join(os.path.dirname(os.path.abspath(oracle.__file__)), "_ref", "libref_heap.so")
    if not os.path.exists(path):
        pytest.skip("oracle/_ref not built (no reference checkout on this machine)")
    heap_lib = C.CDLL(path)

    class HeapEntry(C.Structure):
        _fields_ = [("time_quotient", C.c_double), ("time_remainder", C.c_double), ("event_handler", C.c_void_p),
                    ("counter", C.c_uint)]

    callback_type = C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_void_p, C.c_uint)
    heap_lib.construct_heap.restype = C.c_void_p
    heap_lib.destroy_heap.argtypes = [C.c_void_p]
    heap_lib.insert.argtypes = [C.c_void_p, C.c_double, C.c_double, C.c_void_p, C.c_uint]
    heap_lib.insert.restype = C.c_size_t
    heap_lib.root.argtypes = [C.c_void_p, C.c_void_p, callback_type]
    heap_lib.root.restype = HeapEntry

    rng = np.random.default_rng(7)
    for trial in range(20):
        n = 200
        times = []
        for _ in range(n):
            q, r = oracle.time_from_float(float(rng.integers(0, 4)) + float(rng.random()))
            q, r = oracle.time_add(q, r, float(rng.random()) * 10.0 ** float(rng.integers(-12, 2)))
            times.append((q, r))
        times[5] = (times[4][0], np.nextafter(times[4][1], 2.0))   # neighbours in the last bit of the remainder
        stale = set(int(i) + 1 for i in rng.choice(n, size=n // 4, replace=False))
        popped = set()
        callback = callback_type(lambda scheduler, handler, counter: int(handler in stale or handler in popped))
        heap = heap_lib.construct_heap()
        try:
            for index, (q, r) in enumerate(times):
                assert heap_lib.insert(heap, q, r, index + 1, 0) != C.c_size_t(-1).value
            order = []
            while True:
                entry = heap_lib.root(heap, None, callback)
                if entry.event_handler is None:
                    break
                order.append(entry.event_handler - 1)
                popped.add(entry.event_handler)
        finally:
            heap_lib.destroy_heap(heap)
        valid = [i for i in range(n) if i + 1 not in stale]
        assert len(set(times[i] for i in valid)) == len(valid)   # no exact ties: the order is unique
        compare = lambda a, b: -1 if oracle.time_lt(*times[a], *times[b]) else (1 if oracle.time_lt(*times[b], *times[a]) else 0)
        assert order == sorted(valid, key=functools.cmp_to_key(compare))
