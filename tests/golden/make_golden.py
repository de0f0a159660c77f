"""Generate the committed golden fixtures tests/golden/*.npz from the RUNNING reference.

Usage (in the build container, where the reference checkout exists):

    # one-time: a built copy of the reference (git-ignored), exactly as SURVEY.md 8(c) describes
    cp -r /root/reference baseline/_ref && cd baseline/_ref && \
      python jellyfysh/potential/merged_image_coulomb_potential/merged_image_coulomb_potential_build.py && \
      python jellyfysh/potential/inverse_power_coulomb_bounding_potential/inverse_power_coulomb_bounding_potential_build.py && \
      python jellyfysh/scheduler/heap_scheduler/heap_build.py
    python tests/golden/make_golden.py            # writes tests/golden/*.npz

Interpreter used for the committed fixtures: CPython 3.12.3 (builtin sum() over floats is Neumaier-compensated
there, which is visible in the last bit of norms; see oracle/ecmc_oracle.c: orc_set_sum_mode).
The fixtures hold inputs AND the reference's outputs; tests never need the reference at run time.
"""
import os
import sys
import warnings

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.join(HERE, "..", ".."))

import configs  # noqa: E402
import ref_recorder as rr  # noqa: E402

REF = os.path.abspath(rr.default_ref_root())


# ------------------------------------------------------------------------------------------------------
# potentials
# ------------------------------------------------------------------------------------------------------
def _setting(system_length, dimension=3):
    import jellyfysh.setting as setting
    from jellyfysh.setting.hypercubic_setting import HypercubicSetting
    setting.reset()
    HypercubicSetting(beta=1.0, dimension=dimension, system_length=system_length)
    setting.set_number_of_root_nodes(2)
    setting.set_number_of_nodes_per_root_node(1)
    setting.set_number_of_node_levels(1)
    return setting


def _unit(d, dimension=3, speed=1.0):
    return [speed if k == d else 0.0 for k in range(dimension)]


def potential_vectors():
    rr.import_reference(REF)
    warnings.filterwarnings("ignore")
    from jellyfysh.potential.lennard_jones_potential import LennardJonesPotential
    from jellyfysh.potential.inverse_power_potential import InversePowerPotential
    from jellyfysh.potential.displaced_even_power_potential import DisplacedEvenPowerPotential
    from jellyfysh.potential.hard_sphere_potential import HardSpherePotential
    from jellyfysh.potential.hard_dipole_potential import HardDipolePotential
    from jellyfysh.potential.merged_image_coulomb_potential import MergedImageCoulombPotential
    from jellyfysh.potential.inverse_power_coulomb_bounding_potential import InversePowerCoulombBoundingPotential
    rng = np.random.default_rng(20260117)
    out = {}
    n = 1500

    def floats(a):
        return [float(x) for x in a]

    # --- Lennard-Jones, two parameter sets (C2 and the water oxygen-oxygen values) ---
    setting = _setting(12.0)
    for tag, (k, s) in {"lj_c2": (4.0, 1.0), "lj_water": (0.6217012, 3.165492)}.items():
        pot = LennardJonesPotential(prefactor=k, characteristic_length=s)
        sep = rng.uniform(-2.3 * s, 2.3 * s, size=(n, 3))
        norms = np.linalg.norm(sep, axis=1)
        sep[norms < 0.8 * s] *= (1.0 / norms[norms < 0.8 * s])[:, None] * 0.9 * s
        # a share of points exactly aimed at the sphere / at head-on geometry
        sep[: n // 10, 1:] *= 0.05
        du = rng.exponential(0.3 * k, size=n)  # well depth is k / 4
        du[n // 2: n // 2 + n // 10] *= 1e-3
        direction = rng.integers(0, 3, size=n)
        speed = rng.choice([1.0, 0.5, 2.0], size=n)
        disp = np.array([pot.displacement(_unit(int(d), 3, float(v)), floats(x), float(u))
                         for x, u, d, v in zip(sep, du, direction, speed)])
        der = np.array([pot.derivative(_unit(int(d), 3, float(v)), floats(x))
                        for x, d, v in zip(sep, direction, speed)])
        out[tag + "_params"] = np.array([k, s])
        out[tag + "_sep"], out[tag + "_du"], out[tag + "_dir"], out[tag + "_speed"] = sep, du, direction, speed
        out[tag + "_displacement"], out[tag + "_derivative"] = disp, der

    # --- inverse power: repulsive and attractive ---
    for tag, (power, k) in {"ip_rep": (12.0, 1.0), "ip_coul": (1.0, 2.5), "ip_six": (6.0, 0.7)}.items():
        pot = InversePowerPotential(power=power, prefactor=k)
        sep = rng.uniform(-2.0, 2.0, size=(n, 3))
        sep[np.linalg.norm(sep, axis=1) < 0.3] += 0.5
        du = rng.exponential(1.0, size=n)
        direction = rng.integers(0, 3, size=n)
        c1 = rng.choice([1.0, -1.0, 0.41, -0.82], size=n)
        c2 = rng.choice([1.0, -1.0, 0.41, -0.82], size=n)
        disp = np.array([pot.displacement(_unit(int(d)), floats(x), float(a), float(b), float(u))
                         for x, u, d, a, b in zip(sep, du, direction, c1, c2)])
        der = np.array([pot.derivative(_unit(int(d)), floats(x), float(a), float(b))
                        for x, d, a, b in zip(sep, direction, c1, c2)])
        out[tag + "_params"] = np.array([power, k])
        out[tag + "_sep"], out[tag + "_du"], out[tag + "_dir"] = sep, du, direction
        out[tag + "_c1"], out[tag + "_c2"] = c1, c2
        out[tag + "_displacement"], out[tag + "_derivative"] = disp, der

    # --- displaced even power (harmonic bond of water) ---
    pot = DisplacedEvenPowerPotential(equilibrium_separation=1.012, power=2, prefactor=529.581)
    sep = rng.uniform(-1.3, 1.3, size=(n, 3))
    du = rng.exponential(1.0, size=n)
    direction = rng.integers(0, 3, size=n)
    out["dep_params"] = np.array([529.581, 1.012, 2.0])
    out["dep_sep"], out["dep_du"], out["dep_dir"] = sep, du, direction
    out["dep_displacement"] = np.array([pot.displacement(_unit(int(d)), floats(x), float(u))
                                        for x, u, d in zip(sep, du, direction)])
    out["dep_derivative"] = np.array([pot.derivative(_unit(int(d)), floats(x)) for x, d in zip(sep, direction)])

    # --- hard sphere / hard dipole with general velocities, 2D and 3D ---
    for dim in (2, 3):
        setting = _setting(12.836, dim)
        radius = 0.47619047619047616
        pot = HardSpherePotential(radius=radius)
        sep = rng.uniform(-3.0, 3.0, size=(n, dim))
        norms = np.linalg.norm(sep, axis=1)
        small = norms < 2 * radius
        sep[small] *= (2.0 * radius * (1.0 + rng.uniform(0, 1, size=small.sum())) / norms[small])[:, None]
        vel = rng.normal(size=(n, dim))
        vel[: n // 2] = 0.0
        for i in range(n // 2):
            vel[i, rng.integers(0, dim)] = 1.0
        out[f"hs{dim}_params"] = np.array([radius])
        out[f"hs{dim}_sep"], out[f"hs{dim}_vel"] = sep, vel
        out[f"hs{dim}_displacement"] = np.array([pot.displacement(floats(v), floats(x)) for v, x in zip(vel, sep)])
        lo, hi = 0.952380952380952, 1.047619047619048
        pot = HardDipolePotential(minimum_separation=lo, maximum_separation=hi)
        dirs = rng.normal(size=(n, dim))
        dirs /= np.linalg.norm(dirs, axis=1)[:, None]
        sep = dirs * rng.uniform(lo, hi, size=n)[:, None]
        out[f"hd{dim}_params"] = np.array([lo, hi])
        out[f"hd{dim}_sep"], out[f"hd{dim}_vel"] = sep, vel
        out[f"hd{dim}_displacement"] = np.array([pot.displacement(floats(v), floats(x)) for v, x in zip(vel, sep)])

    # --- merged-image Coulomb and its bounding potential ---
    for tag, (length, alpha, fc, pc, k) in {"mic_l1": (1.0, 3.45, 6, 2, 1.0), "mic_l10": (10.0, 3.45, 6, 2, 332.0),
                                            "mic_var": (2.0, 5.0, 9, 2, 1.0)}.items():
        setting = _setting(length)
        pot = MergedImageCoulombPotential(alpha=alpha, fourier_cutoff=fc, position_cutoff=pc, prefactor=k)
        sep = rng.uniform(-0.5 * length, 0.5 * length, size=(n, 3))
        sep[np.linalg.norm(sep, axis=1) < 0.02 * length] += 0.1 * length
        direction = rng.integers(0, 3, size=n)
        c1 = rng.choice([1.0, -1.0, 0.41, -0.82], size=n)
        c2 = rng.choice([1.0, -1.0, 0.41, -0.82], size=n)
        out[tag + "_params"] = np.array([k, alpha, fc, pc, length])
        out[tag + "_sep"], out[tag + "_dir"], out[tag + "_c1"], out[tag + "_c2"] = sep, direction, c1, c2
        out[tag + "_derivative"] = np.array([pot.derivative(_unit(int(d)), floats(x), float(a), float(b))
                                             for x, d, a, b in zip(sep, direction, c1, c2)])
    for tag, (length, k) in {"ipcb_l1": (1.0, 1.5837), "ipcb_l10": (10.0, 531.2)}.items():
        setting = _setting(length)
        pot = InversePowerCoulombBoundingPotential(prefactor=k)
        sep = rng.uniform(-0.5 * length, 0.5 * length, size=(n, 3))
        sep[np.linalg.norm(sep, axis=1) < 0.02 * length] += 0.1 * length
        direction = rng.integers(0, 3, size=n)
        du = rng.exponential(1.0, size=n)
        du[: n // 5] *= 20.0  # several box traversals
        c1 = rng.choice([1.0, -1.0, 0.41, -0.82], size=n)
        c2 = rng.choice([1.0, -1.0, 0.41, -0.82], size=n)
        out[tag + "_params"] = np.array([k, length])
        out[tag + "_sep"], out[tag + "_dir"], out[tag + "_du"] = sep, direction, du
        out[tag + "_c1"], out[tag + "_c2"] = c1, c2
        out[tag + "_derivative"] = np.array([pot.derivative(_unit(int(d)), floats(x), float(a), float(b))
                                             for x, d, a, b in zip(sep, direction, c1, c2)])
        out[tag + "_displacement"] = np.array([pot.displacement(_unit(int(d)), floats(x), float(a), float(b), float(u))
                                               for x, d, a, b, u in zip(sep, direction, c1, c2, du)])
    setting.reset()
    np.savez_compressed(os.path.join(HERE, "potentials.npz"), **out)
    print("potentials.npz:", len(out), "arrays")


# ------------------------------------------------------------------------------------------------------
# cells, time, periodic boundaries
# ------------------------------------------------------------------------------------------------------
def base_vectors():
    rr.import_reference(REF)
    from jellyfysh.base.time import Time
    out = {}
    rng = np.random.default_rng(7)
    # Time arithmetic (base/time.py)
    q = np.floor(rng.uniform(0, 1e6, size=2000))
    r = rng.uniform(0, 1, size=2000)
    dt = rng.exponential(0.5, size=2000)
    dt[:50] = np.array([0.0, 1.0, 2.0 ** -60, 1.0 - 2.0 ** -53, 3.999999999999999] * 10)
    added = np.array([[t.quotient, t.remainder] for t in (Time(float(a), float(b)) + float(c) for a, b, c in zip(q, r, dt))])
    sub = np.array([Time(float(a), float(b)) - Time(float(c), float(d))
                    for a, b, c, d in zip(added[:, 0], added[:, 1], q, r)])
    ff = np.array([[t.quotient, t.remainder] for t in (Time.from_float(float(x)) for x in q + r)])
    out.update(time_q=q, time_r=r, time_dt=dt, time_added=added, time_sub=sub, time_from_float=ff)
    # periodic boundaries and cells for a few geometries
    geometries = [(3, 5.0, (5, 5, 5), 1), (3, 12.699208415745595, (12, 12, 12), 1), (3, 1.0, (3, 5, 7), 1),
                  (2, 12.836, (13, 13), 1), (3, 10.0, (6, 6, 6), 2), (3, 50.79683366298238, (48, 48, 48), 1)]
    for g, (dim, length, cps, nl) in enumerate(geometries):
        setting = _setting(length, dim)
        from jellyfysh.activator.internal_state.cell_occupancy.cells.cuboid_periodic_cells import CuboidPeriodicCells
        cells = CuboidPeriodicCells(cells_per_side=list(cps), neighbor_layers=nl)
        all_cells = list(cells.yield_cells())
        limit = 4000
        out[f"geo{g}_params"] = np.array([dim, length, nl] + list(cps), dtype=np.float64)
        out[f"geo{g}_cell_min"] = np.array([c.cell_min for c in all_cells[:limit]])
        out[f"geo{g}_cell_max"] = np.array([c.cell_max for c in all_cells[:limit]])
        index = {c: i for i, c in enumerate(all_cells)}
        x = rng.uniform(0, length, size=(3000, dim))
        # positions exactly on cell walls
        for i in range(200):
            c = all_cells[rng.integers(0, len(all_cells))]
            d = rng.integers(0, dim)
            x[i, d] = c.cell_min[d] if i % 2 else c.cell_max[d]
        out[f"geo{g}_pos"] = x
        out[f"geo{g}_pos_cell"] = np.array([index[cells.position_to_cell([float(v) for v in p])] for p in x])
        probe = [all_cells[i] for i in rng.integers(0, len(all_cells), size=40)]
        out[f"geo{g}_probe"] = np.array([index[c] for c in probe])
        out[f"geo{g}_nearby"] = np.array([sorted(index[n] for n in cells.nearby_cells(c)) for c in probe])
        out[f"geo{g}_neighbor_pos"] = np.array([[index[cells.neighbor_cell(c, d, True)] for d in range(dim)]
                                                for c in probe])
        rel = [all_cells[i] for i in rng.integers(0, len(all_cells), size=40)]
        out[f"geo{g}_rel"] = np.array([index[c] for c in rel])
        out[f"geo{g}_translate"] = np.array([index[cells.translate(c, r_)] for c, r_ in zip(probe, rel)])
        out[f"geo{g}_relative"] = np.array([index[cells.relative_cell(c, r_)] for c, r_ in zip(probe, rel)])
        pb = setting.periodic_boundaries
        s = rng.uniform(-1.5 * length, 1.5 * length, size=3000)
        out[f"geo{g}_sep_in"] = s
        out[f"geo{g}_sep_out"] = np.array([pb.correct_separation_entry(float(v), 0) for v in s])
        out[f"geo{g}_pos_out"] = np.array([pb.correct_position_entry(float(v), 0) for v in s])
    setting.reset()
    np.savez_compressed(os.path.join(HERE, "base.npz"), **out)
    print("base.npz:", len(out), "arrays")


# ------------------------------------------------------------------------------------------------------
# whole-chain traces
# ------------------------------------------------------------------------------------------------------
def _tables_of(run):
    dim = run.setting.dimension
    n_cells = len(list(run._cells().yield_cells()))
    out = {}
    bounds = np.full((n_cells, dim, 2), np.nan)
    vetoes = [h for h in run.mediator._activator.get_event_handlers() if "CellVeto" in type(h).__name__]
    if not vetoes:
        # cell-bounding configuration: CellBoundingPotential._derivative_bounds = (upper dict, lower dict)
        handler = [h for h in run.mediator._activator.get_event_handlers() if "CellBounding" in type(h).__name__][0]
        stored = handler._bounding_potential._derivative_bounds  # a bare dict when no lower bounds were asked for
        upper, lower = (stored, None) if isinstance(stored, dict) else stored
        for cell, per_direction in upper.items():
            for d in range(dim):
                bounds[run._cell_index(cell), d, 0] = per_direction[d]
                bounds[run._cell_index(cell), d, 1] = -lower[cell][d] if lower is not None else 0.0
        out["bounds"] = bounds
        return out
    veto = vetoes[0]
    for cell, b in veto._derivative_bounds.items():
        for d in range(dim):
            bounds[run._cell_index(cell), d, 0] = b[d][0]
            bounds[run._cell_index(cell), d, 1] = b[d][1]
    out["bounds"] = bounds
    for name, walkers in (("upper", veto._upper_bound_walker), ("lower", veto._lower_bound_walker)):
        for d in range(dim):
            w = walkers[d]
            out[f"{name}{d}_cell_a"] = np.array([run._cell_index(e[0].item) for e in w._table], dtype=np.int32)
            out[f"{name}{d}_cell_b"] = np.array([run._cell_index(e[1].item) if len(e) > 1 else -1 for e in w._table],
                                                dtype=np.int32)
            out[f"{name}{d}_rate_a"] = np.array([e[0].rate for e in w._table])
            out[f"{name}{d}_rates"] = np.array([w.total_rate, w._mean_rate])
    return out


def _pack_snapshots(run):
    snaps = run.snapshots
    max_sur = max([len(s["surplus"]) for s in snaps] + [1])
    return {"snap_event": np.array([s["event"] for s in snaps], dtype=np.int64),
            "snap_positions": np.array([s["positions"] for s in snaps]),
            "snap_occupants": np.array([s["occupants"] for s in snaps], dtype=np.int32),
            "snap_surplus": np.array([list(s["surplus"]) + [-1] * (max_sur - len(s["surplus"])) for s in snaps],
                                     dtype=np.int32),
            "snap_n_surplus": np.array([len(s["surplus"]) for s in snaps], dtype=np.int32),
            "snap_active": np.array([s["active"] for s in snaps], dtype=np.int32),
            "snap_direction": np.array([s["direction"] for s in snaps], dtype=np.int32),
            "snap_time": np.array([[s["time_q"], s["time_r"]] for s in snaps]),
            "snap_mode": np.array([s.get("mode", 0) for s in snaps], dtype=np.int32)}


def chain_trace(name, ini, positions, seed, stream, n_events, snapshot_every, meta, charges=None, max_occupants=1,
                composites=None):
    run = rr.ReferenceRun(REF, ini, seed=seed, stream=stream, positions=positions, composites=composites)
    try:
        if composites is not None:
            positions = np.asarray(composites[1], dtype=np.float64).reshape(-1, np.shape(composites[1])[-1])
        records = run.run(max_events=n_events, snapshot_every=snapshot_every, max_occupants=max_occupants)
        out = {"records": records, "positions0": np.asarray(positions, dtype=np.float64),
               "final_positions": run.positions(), "seed": np.array([seed, stream], dtype=np.int64)}
        if composites is not None:
            out["roots0"] = np.asarray(composites[0], dtype=np.float64)
            out["final_roots"] = run.roots()
            out["snap_roots"] = np.array([s["roots"] for s in run.snapshots])
        out.update({"meta_" + k: np.asarray(v) for k, v in meta.items()})
        if charges is not None:
            out["charges"] = np.asarray(charges, dtype=np.float64)
        if meta.get("far_field", 1):
            out.update(_tables_of(run))
        out.update(_pack_snapshots(run))
        out["host_times"] = np.array(run.host_times, dtype=np.float64).reshape(-1, 3)
    finally:
        run.close()
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)
    kinds = np.bincount(records["kind"], minlength=9)
    print(f"{name}.npz: {len(records)} events, kinds pair/veto/boundary/eoc/cell-bounding/bond/factor-pair/bending = "
          f"{kinds[1:].tolist()}, "
          f"accepted = {int(records['accepted'].sum())}, snapshots = {len(run.snapshots)}, "
          f"max surplus = {int(out['snap_n_surplus'].max())}")


def chain_traces():
    # LJ, the C2 structure at small N: 40 atoms in 5^3 cells (the reference runs ~1e3 events/s)
    n, ncell, length = 40, 5, 5.0
    pos = configs.lattice_start(n, length, ncell, seed=1000)
    chain_trace("trace_lj_small", configs.lennard_jones_ini(n, length, ncell, chain_time=3.0), pos, seed=7, stream=3,
                n_events=6000, snapshot_every=250,
                meta=dict(n=n, cells_per_side=[ncell] * 3, system_length=length, beta=1.0, lj=[4.0, 1.0],
                          estimator=[1.5, 4], chain_time=3.0))
    # the same with sampling events in between: interaction candidates survive a host control event
    chain_trace("trace_lj_sampling", configs.lennard_jones_ini(n, length, ncell, chain_time=3.0,
                                                               sampling_interval=0.0317), pos, seed=7, stream=4,
                n_events=2500, snapshot_every=250,
                meta=dict(n=n, cells_per_side=[ncell] * 3, system_length=length, beta=1.0, lj=[4.0, 1.0],
                          estimator=[1.5, 4], chain_time=3.0, sampling_interval=0.0317))
    # LJ dense: 100 atoms in 4^3 cells -> crowded cells, surplus list in constant use
    n, ncell, length = 100, 4, 5.2
    pos = configs.uniform_start(n, length, seed=1001)
    chain_trace("trace_lj_surplus", configs.lennard_jones_ini(n, length, ncell, chain_time=0.7, surplus_handlers=100),
                pos, seed=8, stream=1, n_events=3000, snapshot_every=100,
                meta=dict(n=n, cells_per_side=[ncell] * 3, system_length=length, beta=1.0, lj=[4.0, 1.0],
                          estimator=[1.5, 4], chain_time=0.7))
    # Coulomb atoms, the C3 structure (shipped cell_veto.ini shape) with anisotropic cell counts
    n, cps, length = 12, [4, 5, 4], 1.0
    pos = configs.uniform_start(n, length, seed=5)
    chain_trace("trace_coulomb_small", configs.coulomb_atoms_ini(n, cps, points_per_side=4), pos, seed=11, stream=2,
                n_events=6000, snapshot_every=250, charges=np.ones(n),
                meta=dict(n=n, cells_per_side=cps, system_length=length, beta=2.0, mic=[1.0, 3.45, 6, 2],
                          ipcb=[1.5837], estimator=[1.0, 4], chain_time=0.78965))
    # Coulomb atoms with more atoms than cells can hold singly: surplus pairs with bounding potential
    n, cps, length = 48, [4, 4, 4], 1.0
    pos = configs.uniform_start(n, length, seed=6)
    chain_trace("trace_coulomb_surplus", configs.coulomb_atoms_ini(n, cps, points_per_side=4, surplus_handlers=48),
                pos, seed=12, stream=5, n_events=2500, snapshot_every=100, charges=np.ones(n),
                meta=dict(n=n, cells_per_side=cps, system_length=length, beta=2.0, mic=[1.0, 3.45, 6, 2],
                          ipcb=[1.5837], estimator=[1.0, 4], chain_time=0.78965))


def no_cell_traces():
    # Coulomb atoms without a cell system (shipped coulomb_atoms/power_bounded.ini, sized for 6 atoms): the pair factors
    # come from the factor type map, every other atom is a candidate of every event, no cell-boundary events
    n, length = 6, 1.0
    pos = configs.uniform_start(n, length, seed=21)
    chain_trace("trace_coulomb_power_bounded", configs.coulomb_power_bounded_ini(REF, n_atoms=n), pos, seed=14, stream=7,
                n_events=4000, snapshot_every=250, charges=np.ones(n),
                meta=dict(n=n, cells_per_side=[1, 1, 1], system_length=length, beta=2.0, mic=[1.0, 3.45, 6, 2],
                          ipcb=[1.5837], chain_time=0.78965, far_field=0))


def dipole_motion_traces():
    n = 3
    # the shipped dipoles/dipole_motion.ini: the independent active unit alternates between a leaf unit and the root unit
    # of a dipole (RootLeafUnitActiveSwitcher; root-unit-active composite-object and two-leaf handlers), three dipoles;
    # once without and once with the shipped sampling events (candidates that survive a host control event)
    roots, leaves = configs.dipole_start(n, seed=47)
    replacements = [("number_of_root_nodes = 2", f"number_of_root_nodes = {n}"),
                    ("number_event_handlers = 2", f"number_event_handlers = {2 * (n - 1)}"),
                    ("number_event_handlers = 1", f"number_event_handlers = {2 * (n - 1)}")]
    meta = dict(n=2 * n, nodes_per_root=2, system_length=1.0, beta=1.0, chain_time=0.78965, mic=[1.0, 3.45, 6, 2],
                ipcb=[1.5837], harmonic=[200.0, 0.1, 2.0], repulsive=[6.0, 1.0e-6], lifting=0, initial_active=0,
                far_field=0, switch_chain_length=[0.69, 0.7])
    ini = configs.shipped_without_sampling(REF, ("2018_JCP_149_064113", "dipoles", "dipole_motion.ini"),
                                           replacements=replacements)
    chain_trace("trace_dipole_motion", ini, None, seed=31, stream=21, n_events=4000, snapshot_every=250,
                composites=(roots, leaves), charges=np.tile([1.0, -1.0], n), meta=meta)
    ini = configs.shipped_ini(REF, "2018_JCP_149_064113", "dipoles", "dipole_motion.ini")
    ini = ini.replace("filename = config_files/", "filename = " + os.path.join(REF, "jellyfysh", "config_files") + "/")
    for old, new in replacements + [("sampling_interval = 0.56789", "sampling_interval = 0.0831"),
                                    ("filename = output/2018_JCP_149_064113/dipoles/SamplesOfSeparation_DipoleMotion.dat",
                                     "filename = /tmp/jf_b200_golden_separation.dat")]:
        assert old in ini, old
        ini = ini.replace(old, new)
    chain_trace("trace_dipole_motion_sampling", ini, None, seed=32, stream=22, n_events=2500, snapshot_every=250,
                composites=(roots, leaves), charges=np.tile([1.0, -1.0], n),
                meta=dict(meta, sampling_interval=0.0831))


def no_cell_molecule_traces():
    # composite point objects without a cell system: the three shipped dipoles/dipole_factors_*.ini (composite-object
    # Coulomb factor "[0, 1, 2, 3]" with each of the three lifting schemes, harmonic bond, 1/r^6 repulsion between
    # unlike charges of different dipoles), sized for three dipoles
    n = 3
    for k, lifting in enumerate(("inside_first", "outside_first", "ratio")):
        roots, leaves = configs.dipole_start(n, seed=41 + k)
        ini = configs.shipped_without_sampling(
            REF, ("2018_JCP_149_064113", "dipoles", f"dipole_factors_{lifting}.ini"),
            replacements=[("number_of_root_nodes = 2", f"number_of_root_nodes = {n}"),
                          ("number_event_handlers = 1", f"number_event_handlers = {2 * (n - 1)}")])
        chain_trace(f"trace_dipole_factors_{lifting}", ini, None, seed=15 + k, stream=9 + k, n_events=3000,
                    snapshot_every=250, composites=(roots, leaves), charges=np.tile([1.0, -1.0], n),
                    meta=dict(n=2 * n, nodes_per_root=2, system_length=1.0, beta=1.0, chain_time=0.78965,
                              mic=[1.0, 3.45, 6, 2], ipcb=[1.5837], harmonic=[200.0, 0.1, 2.0], repulsive=[6.0, 1.0e-6],
                              lifting=k, initial_active=0, far_field=0))
    # the shipped dipoles/atom_factors.ini: the Coulomb interaction as bounded leaf-to-leaf factors between the dipoles
    # ("[0, 2], Coulomb" ...: TwoLeafUnitBoundingPotentialEventHandler fed by the factor type map), three dipoles
    roots, leaves = configs.dipole_start(n, seed=44)
    ini = configs.shipped_without_sampling(
        REF, ("2018_JCP_149_064113", "dipoles", "atom_factors.ini"),
        replacements=[("number_of_root_nodes = 2", f"number_of_root_nodes = {n}"),
                      ("number_event_handlers = 2", f"number_event_handlers = {2 * (n - 1)}"),
                      ("number_event_handlers = 1", f"number_event_handlers = {2 * (n - 1)}")])
    chain_trace("trace_dipole_atom_factors", ini, None, seed=21, stream=15, n_events=3000, snapshot_every=250,
                composites=(roots, leaves), charges=np.tile([1.0, -1.0], n),
                meta=dict(n=2 * n, nodes_per_root=2, system_length=1.0, beta=1.0, chain_time=0.78965,
                          mic=[1.0, 3.45, 6, 2], ipcb=[1.5837], harmonic=[200.0, 0.1, 2.0], repulsive=[6.0, 1.0e-6],
                          lifting=0, initial_active=0, far_field=0, leaf_pairs=1))
    # the shipped water/coulomb_power_bounded_lj_inverted.ini: three water molecules without a cell system, Coulomb as
    # nine bounded leaf-to-leaf factors per pair of molecules, Lennard-Jones between the oxygens, bonds, bending
    nw = 3
    roots, leaves = configs.water_start(nw, 10.0, seed=11, jitter=0.2)
    ini = configs.shipped_without_sampling(
        REF, ("2018_JCP_149_064113", "water", "coulomb_power_bounded_lj_inverted.ini"),
        replacements=[("number_of_root_nodes = 2", f"number_of_root_nodes = {nw}"),
                      ("number_event_handlers = 3", f"number_event_handlers = {3 * (nw - 1)}"),
                      ("number_event_handlers = 1\nfactor_type_maps = factor_type_maps\n\n[LennardJonesEventHandler]",
                       f"number_event_handlers = {nw - 1}\nfactor_type_maps = factor_type_maps\n\n[LennardJonesEventHandler]")])
    chain_trace("trace_water_atomic_factors", ini, None, seed=23, stream=17, n_events=3000, snapshot_every=250,
                composites=(roots, leaves), charges=np.tile([0.41, -0.82, 0.41], nw),
                meta=dict(n=3 * nw, nodes_per_root=3, system_length=10.0, beta=1.679, chain_time=2.12345,
                          mic=[332.0, 3.45, 6, 2], ipcb=[531.2], lj=[0.6217012, 3.165492], harmonic=[529.581, 1.012, 2.0],
                          bending=[75.9, 1.9764], bending_offset=10.0, bending_max_displacement=0.1, initial_active=1,
                          far_field=0, leaf_pairs=1))
    # the shipped water/single_molecule.ini: harmonic bonds and the bending factor of one molecule, nothing else
    roots, leaves = configs.water_start(1, 10.0, seed=7)
    ini = configs.shipped_without_sampling(REF, ("2018_JCP_149_064113", "water", "single_molecule.ini"))
    chain_trace("trace_water_single_molecule", ini, None, seed=19, stream=13, n_events=3000, snapshot_every=250,
                composites=(roots, leaves), charges=np.array([0.41, -0.82, 0.41]),
                meta=dict(n=3, nodes_per_root=3, system_length=10.0, beta=1.679, chain_time=2.12345,
                          harmonic=[529.581, 1.012, 2.0], bending=[75.9, 1.9764], bending_offset=10.0,
                          bending_max_displacement=0.1, initial_active=1, far_field=0))


def composite_cell_bounding_traces():
    # the shipped dipoles/cell_bounded.ini sized for four dipoles: composite-object Coulomb handlers for nearby cells and
    # the surplus, TwoCompositeObjectCellBoundingPotentialEventHandler for every object in a cell that is not nearby
    # (3 x 5 x 7 root-level cells), harmonic bond, 1/r^6 repulsion, factors kept across cell-boundary events of the root
    n = 4
    roots, leaves = configs.dipole_start(n, seed=47, minimum_distance=0.25)
    ini = configs.shipped_without_sampling(
        REF, ("2018_JCP_149_064113", "dipoles", "cell_bounded.ini"),
        replacements=[("number_of_root_nodes = 2", f"number_of_root_nodes = {n}"),
                      ("number_event_handlers = 1", f"number_event_handlers = {2 * (n - 1)}"),
                      ("number_trials = 1000", "number_trials = 100")])
    chain_trace("trace_dipole_cell_bounded", ini, None, seed=25, stream=19, n_events=4000, snapshot_every=250,
                composites=(roots, leaves), charges=np.tile([1.0, -1.0], n),
                meta=dict(n=2 * n, nodes_per_root=2, system_length=1.0, beta=1.0, chain_time=0.78965,
                          cells_per_side=[3, 5, 7], neighbor_layers=1, mic=[1.0, 3.45, 6, 2], ipcb=[1.5837],
                          harmonic=[200.0, 0.1, 2.0], repulsive=[6.0, 1.0e-6], lifting=0, initial_active=0, far_field=2))


def cell_bounding_traces():
    # Coulomb atoms with the far field through TwoLeafUnitCellBoundingPotentialEventHandler (shipped
    # coulomb_atoms/cell_bounded.ini shape)
    n, cps, length = 10, [4, 5, 4], 1.0
    pos = configs.uniform_start(n, length, seed=9)
    charges = np.ones(n)
    chain_trace("trace_coulomb_cell_bounded",
                configs.coulomb_atoms_ini(n, cps, points_per_side=4, far_field="cell_bounding",
                                          estimator_prefactor=1.5, surplus_handlers=n),
                pos, seed=13, stream=6, n_events=5000, snapshot_every=250, charges=charges,
                meta=dict(n=n, cells_per_side=cps, system_length=length, beta=2.0, mic=[1.0, 3.45, 6, 2],
                          ipcb=[1.5837], estimator=[1.5, 4], chain_time=0.78965, far_field=2))


def lifting_vectors():
    """The three lifting schemes and the bending potential of the running reference on random inputs: tests/golden/
    lifting.npz. random.uniform is replaced by a + (b - a) * u with recorded u (what CPython computes from random())."""
    rr.import_reference(REF)
    import random as _random
    from jellyfysh.lifting.inside_first_lifting import InsideFirstLifting
    from jellyfysh.lifting.outside_first_lifting import OutsideFirstLifting
    from jellyfysh.lifting.ratio_lifting import RatioLifting
    from jellyfysh.potential.bending_potential import BendingPotential
    rng = np.random.default_rng(4711)
    n_cases, n_units = 600, 6
    rates = np.zeros((n_cases, n_units))
    active = np.zeros(n_cases, dtype=np.int32)
    uniforms = rng.random((n_cases, 2))
    expected = np.zeros((n_cases, 3), dtype=np.int32)
    saved = _random.uniform
    try:
        for case in range(n_cases):
            r = rng.normal(size=n_units) * rng.choice([0.1, 1.0, 30.0])
            a = int(rng.integers(n_units))
            r[a] = abs(r[a]) + 1e-3  # the active unit has a positive derivative
            if not (np.delete(r, a) <= 0).any():
                r[(a + 1) % n_units] = -abs(r[(a + 1) % n_units])
            rates[case], active[case] = r, a
            for k, cls in enumerate((InsideFirstLifting, OutsideFirstLifting, RatioLifting)):
                draws = iter(uniforms[case])
                _random.uniform = lambda lo, hi, _d=draws: lo + (hi - lo) * next(_d)
                scheme = cls()
                scheme.reset()
                for i in range(n_units):
                    scheme.insert(float(r[i]), (i,), i == a)
                expected[case, k] = scheme.get_active_identifier()[0]
    finally:
        _random.uniform = saved
    setting = _setting(10.0)
    bending = BendingPotential(equilibrium_angle=1.9764, prefactor=75.9)
    n_b = 400
    s1 = rng.normal(size=(n_b, 3)) * 1.0
    s2 = rng.normal(size=(n_b, 3)) * 1.0
    direction = rng.integers(3, size=n_b).astype(np.int32)
    triples = np.zeros((n_b, 3))
    for i in range(n_b):
        triples[i] = bending.derivative(_unit(int(direction[i]), speed=1.7), [float(x) for x in s1[i]],
                                        [float(x) for x in s2[i]])
    setting.reset()
    np.savez_compressed(os.path.join(HERE, "lifting.npz"), rates=rates, active=active, uniforms=uniforms,
                        expected=expected, bending_s1=s1, bending_s2=s2, bending_direction=direction,
                        bending_triples=triples, bending_params=np.array([75.9, 1.9764, 1.7]))
    print("lifting.npz:", n_cases, "lifting cases,", n_b, "bending triples")


def dipole_traces():
    # C1: the shipped hard_disk_dipoles_cells.ini from the shipped PDB start configuration (81 dipoles = 162 disks,
    # 13^2 leaf-level cells, unbounded occupancy, hard-sphere pairs + hard-dipole tether, root units follow)
    roots, leaves = configs.read_pdb_dipoles(REF)
    chain_trace("trace_hard_disk_dipoles", configs.hard_disk_dipoles_cells_ini(REF), None, seed=17, stream=9,
                n_events=6000, snapshot_every=500, max_occupants=6, composites=(roots, leaves),
                meta=dict(n=162, cells_per_side=[13, 13], system_length=12.836, beta=1.0, chain_time=1.0, far_field=0,
                          nodes_per_root=2, hard_sphere=[0.476190476190476],
                          hard_dipole=[0.952380952380952, 1.047619047619048], max_occupants=6))


def sequential_dipole_trace():
    # the shipped hard_disk_dipoles.ini: no cell system (160 hard-disk candidates + the tether per event), general
    # velocities -- the sequential-direction end-of-chain handler rotates the velocity by 20 degrees every chain time.
    # The start configuration goes through the reference's own PdbInputHandler (MDAnalysis or the stand-in).
    shims = os.path.join(HERE, "..", "..", "jellyfysh_b200", "shims")
    try:
        import MDAnalysis  # noqa: F401
    except ImportError:
        sys.path.append(os.path.abspath(shims))
    roots, leaves = configs.read_pdb_dipoles(REF)
    name, seed, stream, n_events = "trace_hard_disk_dipoles_sequential", 23, 4, 5000
    run = rr.ReferenceRun(REF, configs.hard_disk_dipoles_ini(REF, chain_time=1.5), seed=seed, stream=stream)
    try:
        positions0, roots0 = run.positions(), run.roots()
        assert np.array_equal(positions0, leaves.reshape(-1, 2)) and np.array_equal(roots0, roots)
        records = run.run(max_events=n_events, snapshot_every=250, max_occupants=1)
        out = {"records": records, "positions0": positions0, "roots0": roots0, "final_positions": run.positions(),
               "final_roots": run.roots(), "seed": np.array([seed, stream], dtype=np.int64),
               "snap_roots": np.array([s["roots"] for s in run.snapshots]),
               "snap_velocities": np.array([s["velocities"] for s in run.snapshots])}
        meta = dict(n=162, system_length=12.836, beta=1.0, chain_time=1.5, nodes_per_root=2,
                    hard_sphere=[0.476190476190476], hard_dipole=[0.952380952380952, 1.047619047619048],
                    delta_phi_degree=20.0)
        out.update({"meta_" + k: np.asarray(v) for k, v in meta.items()})
        out.update(_pack_snapshots(run))
        out["host_times"] = np.array(run.host_times, dtype=np.float64).reshape(-1, 3)
    finally:
        run.close()
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)
    kinds = np.bincount(records["kind"], minlength=9)
    print(f"{name}.npz: {len(records)} events, kinds pair/veto/boundary/eoc/cell-bounding/bond = {kinds[1:7].tolist()}, "
          f"snapshots = {len(run.snapshots)}")


def _composite_veto_tables(run):
    """Tables of the CompositeObjectCellVetoEventHandler (same layout as the leaf-unit handler's)."""
    return _tables_of(run)


def water_traces():
    # C4: the shipped water/coulomb_cell_veto_lj_inverted.ini with 32 molecules, 200 estimator trials
    n = 32
    roots, leaves = configs.water_start(n, 10.0, seed=4)
    chain_trace("trace_water", configs.water_ini(REF, n_molecules=n, number_trials=200), None, seed=23, stream=7,
                n_events=4000, snapshot_every=500, max_occupants=1, composites=(roots, leaves),
                charges=np.tile([0.41, -0.82, 0.41], n),
                meta=dict(n=3 * n, cells_per_side=[6, 6, 6], neighbor_layers=2, system_length=10.0, beta=1.679,
                          chain_time=2.12345, nodes_per_root=3, mic=[332.0, 3.45, 6, 2], ipcb=[531.2],
                          lj=[0.6217012, 3.165492], harmonic=[529.581, 1.012, 2.0], bending=[75.9, 1.9764],
                          bending_offset=10.0, bending_max_displacement=0.112321434, initial_active=1))
    # dense and small: 12 molecules in a box of 6 with 4^3 cells and one neighbour layer -> composite pair events,
    # surplus molecules, accepted cell vetoes with lifting into either molecule
    n = 12
    roots, leaves = configs.water_start(n, 6.0, seed=5, jitter=0.2)
    chain_trace("trace_water_dense", configs.water_ini(REF, n_molecules=n, number_trials=200, system_length=6.0,
                                                       cells_per_side=[4, 4, 4], neighbor_layers=1), None,
                seed=24, stream=8, n_events=4000, snapshot_every=500, max_occupants=1, composites=(roots, leaves),
                charges=np.tile([0.41, -0.82, 0.41], n),
                meta=dict(n=3 * n, cells_per_side=[4, 4, 4], neighbor_layers=1, system_length=6.0, beta=1.679,
                          chain_time=2.12345, nodes_per_root=3, mic=[332.0, 3.45, 6, 2], ipcb=[531.2],
                          lj=[0.6217012, 3.165492], harmonic=[529.581, 1.012, 2.0], bending=[75.9, 1.9764],
                          bending_offset=10.0, bending_max_displacement=0.112321434, initial_active=1))


def leaf_cell_water_traces():
    # the shipped water/coulomb_power_bounded_lj_cell_bounded.ini: a cell system for the OXYGENS only (cell level 2 with a
    # charge indicator), Lennard-Jones between the oxygens through it (piecewise constant bound for nearby cells and the
    # surplus, cell-bounding potential for all other cells, cell boundary of the oxygen), composite-object Coulomb factors,
    # bonds and bending from the factor type map. Twelve molecules in the shipped 6^3 cells; sixteen in 4^3 cells, where
    # two oxygens share a cell now and then (surplus)
    meta = dict(neighbor_layers=1, system_length=10.0, beta=1.679, chain_time=2.12345, nodes_per_root=3,
                mic=[332.0, 3.45, 6, 2], ipcb=[531.2], lj=[0.6217012, 3.165492], harmonic=[529.581, 1.012, 2.0],
                bending=[75.9, 1.9764], bending_offset=10.0, bending_max_displacement=0.112321434, initial_active=1,
                lj_offset=10.0, lj_max_displacement=0.24353253124, cell_child=1, far_field=2)
    # ... and forty in 3^3 cells: 17-18 oxygens in the surplus at all times, every cell nearby
    for name, n, cells, seed, jitter in (("trace_water_lj_cell_bounded", 12, [6, 6, 6], 13, 0.2),
                                         ("trace_water_lj_cell_bounded_dense", 16, [4, 4, 4], 14, 0.3),
                                         ("trace_water_lj_cell_bounded_surplus", 40, [3, 3, 3], 15, 0.3)):
        if os.environ.get("JF_ONLY_TRACE") not in (None, name):
            continue
        roots, leaves = configs.water_start(n, 10.0, seed=seed, jitter=jitter)
        ini = configs.shipped_without_sampling(
            REF, ("2018_JCP_149_064113", "water", "coulomb_power_bounded_lj_cell_bounded.ini"),
            replacements=[("number_of_root_nodes = 2", f"number_of_root_nodes = {n}"),
                          ("number_event_handlers = 1", f"number_event_handlers = {n - 1}"),
                          ("cells_per_side = 6, 6, 6", "cells_per_side = " + ", ".join(map(str, cells)))])
        chain_trace(name, ini, None, seed=41 + n, stream=31 + n, n_events=4000 if n < 40 else 2500, snapshot_every=250,
                    max_occupants=1,
                    composites=(roots, leaves), charges=np.tile([0.41, -0.82, 0.41], n),
                    meta=dict(meta, n=3 * n, cells_per_side=cells))


if __name__ == "__main__":  # noqa
    which = sys.argv[1:] or ["potentials", "base", "traces", "cell_bounding", "dipoles", "sequential", "water", "lifting",
                             "no_cells"]
    if "no_cells" in which:
        no_cell_traces()
        no_cell_molecule_traces()
    if "no_cells" in which or "dipole_motion" in which:
        dipole_motion_traces()
    if "water" in which:
        water_traces()
    if "water" in which or "leaf_cells" in which:
        leaf_cell_water_traces()
    if "lifting" in which:
        lifting_vectors()
    if "dipoles" in which:
        dipole_traces()
    if "sequential" in which:
        sequential_dipole_trace()
    if "cell_bounding" in which:
        cell_bounding_traces()
        composite_cell_bounding_traces()
    if "potentials" in which:
        potential_vectors()
    if "base" in which:
        base_vectors()
    if "traces" in which:
        chain_traces()
