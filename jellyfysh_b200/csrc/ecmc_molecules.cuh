// ecmc_molecules.cuh -- the event kernel for composite point objects in root-level cells (water).
//
// One warp advances one Markov chain, like ecmc::event_kernel, but the factors are those of the shipped water
// configurations (jellyfysh/config_files/2018_JCP_149_064113/water/coulomb_cell_veto_lj_inverted.ini):
//   * composite-object Coulomb pairs with every object in a nearby cell / the surplus
//     (TwoCompositeObjectSummedBoundingPotentialEventHandler, two_composite_object_summed_bounding_potential_event_handler.py:
//     119-202): one LANE per (target object, target leaf) -- the handler's minimum over the target leaves is part of the
//     warp argmin;
//   * the composite-object cell veto for all other cells (CompositeObjectCellVetoEventHandler,
//     composite_object_cell_veto_event_handler.py:110-162 on abstracts/cell_veto_event_handler.py:200-238);
//   * two-leaf factors between objects (Lennard-Jones between the oxygens, "[1, 4], LennardJones"): one lane per object;
//   * intramolecular two-leaf factors (harmonic bonds) and the three-leaf bending factor with its piecewise constant
//     bounding potential (event_handler_with_bounding_potential.py:282-332);
//   * the cell boundary of the ROOT unit, which moves with speed / nodes_per_root (cell_boundary_event_handler.py:98-173);
//   * lifting schemes over the six leaf units (event_handler_with_bounding_potential.py:170-220, lifting/*.py).
// The work of one event is a list of items in shared memory (type, target, sequence number), worked off 32 at a time.
#pragma once

#include "ecmc_kernels.cuh"

namespace ecmc {

struct MoleculeProgram {
    int composite_lifting, bending_lifting;
    int n_inter, bending_enabled;
    int inter[ECMC_MAX_INTER_FACTORS][2];
    int bending_children[3];
    int bending_separations[4];
    int boundary_keeps_factors;
    double bending_prefactor, bending_angle, bending_offset, bending_max_displacement;
    PotentialParams inter_potential;
    // EcmcProgram.root_mode: chain lengths of the leaf-to-root and of the root-to-leaf RootLeafUnitActiveSwitcher
    int root_mode, pad;
    double switch_length[2];
    // EcmcProgram.cell_child: the cells hold only the leaves with child index cell_child - 1 (0: no such system); the
    // piecewise constant bound of the two-leaf factor between those leaves
    int cell_child, pad2;
    double inter_bound_offset, inter_bound_max_displacement;
};

constexpr int kItemCapacity = 192;  // work items per chunk (two ints each: 1.5 KB per warp)
// per-warp sine / cosine scratch of the merged-image Coulomb sums: one sum up to the largest Fourier cutoff, or the three
// sums of mic_derivative_warp3 up to cutoff 6 (3 x 3 x 2 x 7 = 126 doubles; the shipped potentials use 6)
constexpr int kMoleculeTrigDoubles = kTrigDoubles3;
enum ItemType { ITEM_PAIR_LEAF = 0, ITEM_INTER = 1, ITEM_BOND = 2, ITEM_BENDING = 3, ITEM_VETO = 4, ITEM_BOUNDARY = 5,
                ITEM_FAR_OBJECT = 6, ITEM_NEAR_LEAF = 7 };

struct Vec3 {
    double x, y, z;
};
ECMC_D double vcomp(const Vec3 &v, int d) { return d == 0 ? v.x : (d == 1 ? v.y : v.z); }
ECMC_D Vec3 lab_position(const Particle &p) {
    Vec3 v;
    v.x = p.x; v.y = p.y; v.z = p.z;
    return v;
}
ECMC_D Vec3 separation_lab(const Vec3 &from, const Vec3 &to, double L, double half) {
    Vec3 s;
    s.x = correct_separation_in_box(to.x - from.x, L, half);
    s.y = correct_separation_in_box(to.y - from.y, L, half);
    s.z = correct_separation_in_box(to.z - from.z, L, half);
    return s;
}

// BendingPotential.derivative (bending_potential.py:60-138): time derivatives with respect to the units i, j, k for
// s1 = r_i - r_j, s2 = r_k - r_j along direction `dir`
__device__ __noinline__ void bending_derivative(double prefactor, double equilibrium_angle, int dir, double speed,
                                                const Vec3 &s1, const Vec3 &s2, double out[3]) {
    const double n1 = sqrt(fma(s1.x, s1.x, fma(s1.y, s1.y, s1.z * s1.z)));
    const double n2 = sqrt(fma(s2.x, s2.x, fma(s2.y, s2.y, s2.z * s2.z)));
    const double inv1 = 1.0 / n1, inv2 = 1.0 / n2;
    const double cosine = fma(s1.x, s2.x, fma(s1.y, s2.y, s1.z * s2.z)) * inv1 * inv2;
    const double angle = acos(cosine);
    const double du_dangle = prefactor * (angle - equilibrium_angle);
    // sin(angle) for an angle in [0, pi] without the sine: sqrt((1 - cos)(1 + cos))
    const double dangle_dcos = -rsqrt((1.0 - cosine) * (1.0 + cosine));
    const double a1 = vcomp(s1, dir), a2 = vcomp(s2, dir);
    const double dcos_ds1 = a2 * inv1 * inv2 - cosine * a1 * inv1 * inv1;
    const double dcos_ds2 = a1 * inv1 * inv2 - cosine * a2 * inv2 * inv2;
    const double du_ds1 = du_dangle * dangle_dcos * dcos_ds1;
    const double du_ds2 = du_dangle * dangle_dcos * dcos_ds2;
    out[0] = du_ds1 * speed;
    out[1] = (-du_ds1 - du_ds2) * speed;
    out[2] = du_ds2 * speed;
}

// Lifting.insert / get_active_identifier (lifting/lifting.py:49-91 and the three schemes), at most eight units.
// Draws come from the out-state's slot (ECMC_SLOT_CONFIRM) in call order.
// The draws of an out-state (confirmation, lifting) come from many call sites, each event uses one or two of them: one
// out-of-line Philox instead of a dozen inlined copies keeps the kernel's instruction footprint down.
__device__ __noinline__ double confirm_draw(uint32_t seed, uint32_t stream, unsigned long long event, uint32_t index) {
    const StreamKey key = {seed, stream, event};
    return stream_double(key, ECMC_SLOT(ECMC_SLOT_CONFIRM, 0), index);
}

struct Lifting {
    double negative[6];
    int ids[6];
    int n_negative;
    double random_position;
    bool active_recorded;
};
ECMC_D void lifting_reset(Lifting &l) {
    l.n_negative = 0;
    l.random_position = 0.0;
    l.active_recorded = false;
}
ECMC_D void lifting_insert(Lifting &l, double rate, int id, bool is_active, const StreamKey &key, uint32_t &draw) {
    if (rate > 0.0) {
        if (is_active) {
            l.active_recorded = true;
            const double u = confirm_draw(key.seed, key.stream, key.event, draw++);
            l.random_position += 0.0 + (rate - 0.0) * u;
        } else if (!l.active_recorded) {
            l.random_position += rate;
        }
    } else {
#pragma unroll
        for (int i = 0; i < 6; i++)
            if (i == l.n_negative) { l.negative[i] = -rate; l.ids[i] = id; }
        l.n_negative++;
    }
}
// Python's sum() over floats is Neumaier-compensated since CPython 3.12
ECMC_D double pysum_n(const double *v, int n) {
    double f = 0.0, comp = 0.0;
#pragma unroll
    for (int i = 0; i < 6; i++)
        if (i < n) {
            const double x = v[i];
            const double t = __dadd_rn(f, x);
            comp = __dadd_rn(comp, fabs(f) >= fabs(x) ? __dadd_rn(__dsub_rn(f, t), x) : __dadd_rn(__dsub_rn(x, t), f));
            f = t;
        }
    if (comp != 0.0 && isfinite(comp)) f = __dadd_rn(f, comp);
    return f;
}
ECMC_D int lifting_get(const Lifting &l, int kind, const StreamKey &key, uint32_t &draw) {
    double position = l.random_position;
    if (kind == ECMC_LIFTING_OUTSIDE_FIRST) {
        position = pysum_n(l.negative, l.n_negative) - l.random_position;
    } else if (kind == ECMC_LIFTING_RATIO) {
        const double u = confirm_draw(key.seed, key.stream, key.event, draw++);
        position = 0.0 + (pysum_n(l.negative, l.n_negative) - 0.0) * u;
    }
    double summed = 0.0;
    int chosen = -1;
#pragma unroll
    for (int i = 0; i < 6; i++)
        if (i < l.n_negative) {
            summed += l.negative[i];
            if (chosen < 0 && position <= summed) chosen = l.ids[i];
            if (i == l.n_negative - 1 && chosen < 0) chosen = l.ids[i];
        }
    return chosen;
}

// derivative of a pair potential between two leaves in the lab frame (all lanes, identical arguments)
template <int KIND>
ECMC_D double pair_derivative_lab(const PotentialParams &p, int dir, double speed, const Vec3 &from, const Vec3 &to,
                                  double c1, double c2, double L, double half, double *trig, int lane) {
    const Vec3 s = separation_lab(from, to, L, half);
    return derivative_warp<KIND>(p, dir, speed, s.x, s.y, s.z, c1, c2, trig, lane);
}

// CAND / REAL / BOND / INTER: compile-time kinds of the bounding potential of the composite pairs, of the potential the
// composite events are confirmed against (pair and cell veto), of the intramolecular pair factors and of the factors
// between objects; -1 = decided at run time (the generic instantiation is four times the code).
// ALIGNED: the warps of a CTA meet at a barrier before every event. The kernel is bound by instruction fetch (thousands
// of instructions per event, a handful of warps per SM, each somewhere else in the code); warps that walk through the
// event together fetch every instruction once for all of them.
// ROOT_MODE: the program has the root-unit-active mode of dipoles/dipole_motion.ini (EcmcProgram.root_mode: objects of two
// leaves, no cell system). While EcmcChainState.mode is 1 the ROOT unit of the object of `active` (its first leaf) is the
// independent active unit: root and both leaves move with the full velocity, the candidates are those of the
// root-unit-active handlers (see include/ecmc.h), and RootLeafUnitActiveSwitcher events alternate between the two modes.
// Compile-time, so that the other instantiations carry none of it.
// LEAF_CELLS: the cell system stores ONE kind of leaf only (EcmcProgram.cell_child, the shipped
// water/coulomb_power_bounded_lj_cell_bounded.ini: SingleActiveCellOccupancy with cell level 2 and a charge indicator). The
// composite-object pairs then come from the factor type map (every other object, every event); while a leaf of that kind is
// active, the two-leaf factor with the same leaf of other objects is found through the cells -- nearby cells and surplus by
// TwoLeafUnitEventHandlerWithPiecewiseConstantBoundingPotential, all other cells by
// TwoLeafUnitCellBoundingPotentialEventHandler --, and the cell boundary is that of the active leaf, whose events leave the
// composite pairs, the bonds and the bending factor running. Occupants / surplus are leaf identifiers. Compile-time.
template <int CAND, int REAL, int BOND, int INTER, bool RECORD, int WARPS, bool ALIGNED, bool ROOT_MODE = false,
          bool LEAF_CELLS = false>
__global__ void __launch_bounds__(WARPS * 32)
molecule_kernel(const __grid_constant__ DeviceProgram P, const __grid_constant__ MoleculeProgram M, const DeviceState S,
                const RunArgs A) {
    static_assert(!(ROOT_MODE && ALIGNED), "the root-unit-active mode runs without the per-event CTA barrier");
    static_assert(!(ROOT_MODE && LEAF_CELLS), "one special mode per instantiation");
    __shared__ double trig_all[WARPS * kMoleculeTrigDoubles];
    __shared__ int items_all[WARPS * 2 * kItemCapacity];
    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    const int chain = S.first_chain + blockIdx.x * WARPS + warp;
    if (chain >= S.first_chain + S.n_chains) {
        if (ALIGNED) while (!__syncthreads_and(1)) {}  // keep the barriers of the other warps complete
        return;
    }
    double *trig = trig_all + warp * kMoleculeTrigDoubles;
    int *item_code = items_all + warp * 2 * kItemCapacity;  // type | sequence << 4
    int *item_target = item_code + kItemCapacity;

    const int npr = P.nodes_per_root;
    const int n_roots = P.n_particles / npr;
    Particle *part = S.particles + (size_t)chain * P.n_particles;
    Particle *roots = S.roots + (size_t)chain * n_roots;
    int *occ = S.occupants + (size_t)chain * P.n_cells;
    int *sur = S.surplus + (size_t)chain * P.max_surplus;
    EcmcChainState *stp = S.chains + chain;

    int active = stp->active, dir = stp->direction;
    Time now = {stp->time_q, stp->time_r};
    Time eoc = {stp->eoc_q, stp->eoc_r};
    int eoc_next = stp->eoc_next_active;
    int active_cell = stp->active_cell;
    unsigned long long ev = stp->event_counter;
    const uint32_t stream = stp->stream;
    bool was_pending = stp->pending_kind != ECMC_EVENT_NONE;
    int n_surplus = S.n_surplus[chain];
    // candidates of the leaf-level factors kept across a cell-boundary event of the root (EcmcChainState.kept_*)
    int kept_kind = stp->kept_kind, kept_target = stp->kept_target;
    Time kept_time = {stp->kept_q, stp->kept_r};
    double kept_rate = stp->kept_rate, kept_pos = stp->kept_position, kept_root = stp->kept_root_position;
    Time kept_stamp = {stp->kept_stamp_q, stp->kept_stamp_r};

    Vec3 apos = lab_position(part[active]);     // the active leaf
    double acharge = part[active].charge;
    Vec3 rpos = lab_position(roots[active / npr]);  // its root unit
    // root-unit-active mode: which unit is active, the running switcher, the last end of chain, and -- while the root unit
    // is active -- the second leaf of its object (apos is the first one)
    int mode = ROOT_MODE ? stp->mode : 0;
    Time sw = {ROOT_MODE ? stp->switch_q : 0.0, ROOT_MODE ? stp->switch_r : 0.0};
    Time eoc_last = {ROOT_MODE ? stp->eoc_last_q : 0.0, ROOT_MODE ? stp->eoc_last_r : 0.0};
    Vec3 bpos = apos;
    double bcharge = acharge;
    if (ROOT_MODE && mode == 1) {
        const Particle second = part[active + 1];
        bpos = lab_position(second);
        bcharge = second.charge;
    }
    int cid0 = (active_cell / P.cumulative[0]) % P.per_side[0];
    int cid1 = (active_cell / P.cumulative[1]) % P.per_side[1];
    int cid2 = (active_cell / P.cumulative[2]) % P.per_side[2];

    const Time until = {A.until_q, A.until_r};
    const double L = P.length, half = P.half_length, speed = P.speed;
    const unsigned max_events = A.max_events > 0 ? (unsigned)min(A.max_events, 0x7fffffffLL) : 0x7fffffffu;
    unsigned n_events = 0, n_pair = 0, n_veto = 0, n_eoc = 0, n_bond = 0, n_factor = 0, n_boundary = 0;
    unsigned long long n_candidates = 0, n_targets = 0;  // n_targets: gathered target objects (EcmcStats.pair_targets)
    bool stopped_by_time = false;

    auto set_dir = [&](Vec3 &v, double value) { if (dir == 0) v.x = value; else if (dir == 1) v.y = value; else v.z = value; };

    // the Coulomb interaction as bounded leaf-to-leaf factors between the objects, one TwoLeafUnitBoundingPotential-
    // EventHandler per pair of leaves (dipoles/atom_factors.ini), instead of one composite-object handler per pair of
    // objects: the candidates are the same lanes, an event is confirmed against its own pair only and lifts to the target
    const bool leaf_pairs = P.pair_handler == ECMC_PAIR_TWO_LEAF_UNIT_BOUNDING;
    bool done = false;
    while (true) {
        if (ALIGNED) {
            const bool finished = done || n_events >= max_events;
            if (__syncthreads_and(finished)) break;
            if (finished) continue;
        } else if (n_events >= max_events) {
            break;
        }
        const StreamKey key = {P.seed, stream, ev};
        if (ROOT_MODE && mode == 1) {
            // ---- one event with the root unit of object `active / 2` as the independent active unit
            const int active_root = active / npr;
            Time bt = time_inf();
            int bkind = ECMC_EVENT_NONE, btarget = -1, n_cand = 0;
            // the leaf units as the handler that fires saw them in send_event_time (the confirmation of the
            // summed-bounding handler uses those; every out-state time-slices a fresh copy of the objects)
            Vec3 ia = apos, ib = bpos;
            Time in_stamp = now;
            if (was_pending) {
                bkind = stp->pending_kind;
                bt.q = stp->pending_q; bt.r = stp->pending_r;
                btarget = stp->pending_target;
                set_dir(ia, stp->pending_position);
                set_dir(ib, stp->pending_position_y);
                in_stamp.q = stp->pending_stamp_q; in_stamp.r = stp->pending_stamp_r;
            } else {
                // eight lanes per other object: (local leaf i, target leaf j) of the composite-object handler
                // (root_unit_active_two_composite_object_summed_bounding_potential_event_handler.py:116-152, its minimum
                // over the four pairs is part of the warp argmin), then the inter-object factors (a, b) of
                // RootUnitActiveTwoLeafUnitEventHandler (two_leaf_unit_event_handler.py:105-138 for the moving leaf a)
                const int n_items = (n_roots - 1) * 8;
                unsigned long long best_key = 0x7ff0000000000000ull;
                double best_x = INFINITY;
                int best_seq = kSeqNone;
                for (int base = 0; base < n_items; base += 32) {
                    const int item = base + lane;
                    const int object = item >> 3, within = item & 7;
                    const int t = object < active_root ? object : object + 1;
                    double dt = INFINITY;
                    int kind = ECMC_EVENT_NONE, target = -1;
                    if (item < n_items && (within < 4 || within - 4 < M.n_inter)) {
                        const int local = within < 4 ? within >> 1 : M.inter[within - 4][0];
                        const int leaf = t * npr + (within < 4 ? within & 1 : M.inter[within - 4][1]);
                        const Particle tp = part[leaf];
                        const Vec3 s3 = separation_lab(local == 0 ? apos : bpos, lab_position(tp), L, half);
                        const double s0 = vcomp(s3, dir);
                        const double s1 = dir == 0 ? s3.y : (dir == 1 ? s3.z : s3.x);
                        const double s2 = dir == 0 ? s3.z : (dir == 1 ? s3.x : s3.y);
                        if (within < 4) {
                            const double u = stream_double(key, ECMC_SLOT(ECMC_SLOT_PAIR_TIME, t), (uint32_t)within);
                            const double du = -log_unit_interval(1.0 - u) * P.inv_beta;
                            const double c1 = P.pair_use_charge ? (local == 0 ? acharge : bcharge) : 1.0;
                            const double c2 = P.pair_use_charge ? tp.charge : 1.0;
                            dt = displacement_time<CAND>(P.cand_potential, 0, P.inv_speed, L, s0, s1, s2, c1, c2, du);
                            kind = ECMC_EVENT_PAIR; target = t;
                        } else {
                            const double u = stream_double(key, ECMC_SLOT(ECMC_SLOT_FACTOR_TIME, leaf), (uint32_t)local);
                            const double du = -log_unit_interval(1.0 - u) * P.inv_beta;
                            dt = displacement_time<INTER>(M.inter_potential, 0, P.inv_speed, L, s0, s1, s2, 1.0, 1.0,
                                                          needs_potential_change(resolve_kind<INTER>(M.inter_potential.kind)) ? du : 0.0);
                            kind = ECMC_EVENT_FACTOR_PAIR; target = leaf;
                        }
                    }
                    const double x = now.r + dt;
                    const bool finite = kind != ECMC_EVENT_NONE && x < INFINITY;
                    // one candidate per handler: an object whose composite pair has a finite time, every finite factor
                    const unsigned pair_mask = __ballot_sync(kFull, finite && within < 4);
                    for (int g = 0; g < 4; g++) n_cand += ((pair_mask >> (8 * g)) & 0xFu) != 0u;
                    n_cand += __popc(__ballot_sync(kFull, finite && within >= 4));
                    const unsigned long long k64 = finite ? time_key(x) : 0x7ff0000000000000ull;
                    const int owner = warp_argmin(k64, finite ? item : kSeqNone, lane);
                    const unsigned long long pass_key = __shfl_sync(kFull, k64, owner);
                    const int pass_seq = __shfl_sync(kFull, finite ? item : kSeqNone, owner);
                    if (pass_key < best_key || (pass_key == best_key && pass_seq < best_seq)) {
                        best_key = pass_key; best_seq = pass_seq;
                        best_x = __shfl_sync(kFull, x, owner);
                        bkind = __shfl_sync(kFull, kind, owner);
                        btarget = __shfl_sync(kFull, target, owner);
                    }
                }
                n_targets += (unsigned long long)(n_roots - 1);
                if (best_seq != kSeqNone) {
                    const double fl = floor(best_x);
                    bt.q = now.q + fl; bt.r = best_x - fl;
                } else {
                    bkind = ECMC_EVENT_NONE;
                }
            }
            n_cand += 2;  // the end of chain and the root-to-leaf switcher, both in the scheduler all along
            Time event_time = bt;
            int kind = bkind;
            if (time_lt(eoc, event_time)) { event_time = eoc; kind = ECMC_EVENT_END_OF_CHAIN; }
            if (time_lt(sw, event_time)) { event_time = sw; kind = ECMC_EVENT_SWITCH; }
            if (!time_lt(event_time, until)) {
                if (lane == 0) {
                    stp->pending_kind = bkind;
                    stp->pending_q = bt.q; stp->pending_r = bt.r;
                    stp->pending_rate = 0.0;
                    stp->pending_target = btarget;
                    if (!was_pending) {
                        stp->pending_position = vcomp(apos, dir);
                        stp->pending_position_y = vcomp(bpos, dir);
                        stp->pending_root_position = vcomp(rpos, dir);
                        stp->pending_stamp_q = now.q; stp->pending_stamp_r = now.r;
                    }
                }
                stopped_by_time = true;
                break;
            }
            if (was_pending && lane == 0) stp->pending_kind = ECMC_EVENT_NONE;
            was_pending = false;
            {
                // time slice of the root unit and of both leaves, each with the full velocity (abstracts.py:89-107)
                const double step = __dmul_rn(speed, time_sub(event_time, now));
                set_dir(apos, correct_position_entry(__dadd_rn(vcomp(apos, dir), step), L));
                set_dir(bpos, correct_position_entry(__dadd_rn(vcomp(bpos, dir), step), L));
                set_dir(rpos, correct_position_entry(__dadd_rn(vcomp(rpos, dir), step), L));
                now = event_time;
            }
            int new_active = active, rec_target = -1;
            switch (kind) {
            case ECMC_EVENT_PAIR: {
                // send_out_state (:154-190): summed derivatives over the four pairs of leaves, in the handler's order
                rec_target = btarget;
                n_pair++;
                const double step = __dmul_rn(speed, time_sub(event_time, in_stamp));
                set_dir(ia, correct_position_entry(__dadd_rn(vcomp(ia, dir), step), L));
                set_dir(ib, correct_position_entry(__dadd_rn(vcomp(ib, dir), step), L));
                double bounding_rate = 0.0, factor_derivative = 0.0;
                for (int i = 0; i < 2; i++)
                    for (int j = 0; j < 2; j++) {
                        const Particle tp = part[btarget * npr + j];
                        const double c1 = P.pair_use_charge ? (i == 0 ? acharge : bcharge) : 1.0;
                        const double c2 = P.pair_use_charge ? tp.charge : 1.0;
                        const Vec3 from = i == 0 ? ia : ib, to = lab_position(tp);
                        const double b = pair_derivative_lab<CAND>(P.cand_potential, dir, speed, from, to, c1, c2, L, half, trig, lane);
                        bounding_rate += b > 0.0 ? b : 0.0;
                        factor_derivative += pair_derivative_lab<REAL>(P.real_potential, dir, speed, from, to, c1, c2, L, half,
                                                                       trig, lane);
                    }
                if (factor_derivative > 0.0) {
                    if (bounding_rate < factor_derivative) count_rare(A, lane, 7);
                    const double u = confirm_draw(key.seed, key.stream, key.event, 0);
                    if (0.0 + (bounding_rate - 0.0) * u < factor_derivative) new_active = btarget * npr;
                }
                break;
            }
            case ECMC_EVENT_FACTOR_PAIR:
                // RootUnitActiveTwoLeafUnitEventHandler.send_out_state (:102-125): the object of the target leaf takes over
                rec_target = btarget;
                new_active = (btarget / npr) * npr;
                n_factor++;
                break;
            case ECMC_EVENT_END_OF_CHAIN:
                new_active = eoc_next;
                rec_target = eoc_next;
                eoc_last = event_time;
                n_eoc++;
                break;
            case ECMC_EVENT_SWITCH:
                // _send_out_state_leaf_unit_active (root_leaf_unit_active_switcher.py:129-169): random.choice over the leaves
                new_active = active + (int)stream_randbelow(key, ECMC_SLOT(ECMC_SLOT_SWITCH, 0), (uint32_t)npr);
                break;
            default: break;
            }
            const int new_mode = kind == ECMC_EVENT_SWITCH ? 0 : 1;
            if (RECORD && lane == 0 && (int)n_events < A.records_per_chain) {
                EcmcEventRecord rec;
                rec.kind = kind; rec.target = rec_target; rec.target_cell = -1;
                rec.accepted = (kind == ECMC_EVENT_END_OF_CHAIN || kind == ECMC_EVENT_SWITCH) ? 1 : (new_active != active);
                rec.n_candidates = n_cand;
                rec.new_active = new_active;
                rec.new_direction = kind == ECMC_EVENT_END_OF_CHAIN ? (dir + 1) % P.dimension : dir;
                rec.mode = new_mode;
                rec.time_q = event_time.q; rec.time_r = event_time.r;
                rec.active_pos[0] = apos.x; rec.active_pos[1] = apos.y; rec.active_pos[2] = apos.z;
                A.records[(size_t)chain * A.records_per_chain + n_events] = rec;
            }
            ev++;
            n_events++;
            n_candidates += (unsigned long long)n_cand;
            if (kind == ECMC_EVENT_END_OF_CHAIN) dir = dir + 1 == P.dimension ? 0 : dir + 1;
            // commit the object, hand over
            const int new_root = new_active / npr;
            int delta = 0;
            if (lane == 0) {
                Particle p = part[active];
                p.x = apos.x; p.y = apos.y; p.z = apos.z;
                part[active] = p;
                p = part[active + 1];
                p.x = bpos.x; p.y = bpos.y; p.z = bpos.z;
                part[active + 1] = p;
                Particle r = roots[active_root];
                r.x = rpos.x; r.y = rpos.y; r.z = rpos.z;
                roots[active_root] = r;
                if (new_root != active_root) delta = occupancy_insert(occ, sur, n_surplus, 1, P.max_surplus, active_cell, active_root);
            }
            delta = __shfl_sync(kFull, delta, 0);
            if (delta == 2) count_rare(A, lane, 8); else n_surplus += delta;
            __syncwarp();
            if (new_root != active_root) {
                rpos = lab_position(roots[new_root]);
                cid0 = (int)(rpos.x / P.side_length[0]);
                cid1 = (int)(rpos.y / P.side_length[1]);
                cid2 = (int)(rpos.z / P.side_length[2]);
                active_cell = cid0 * P.cumulative[0] + cid1 * P.cumulative[1] + cid2 * P.cumulative[2];
                delta = 0;
                if (lane == 0) delta = occupancy_remove(occ, sur, n_surplus, 1, active_cell, new_root);
                delta = __shfl_sync(kFull, delta, 0);
                if (delta == 2) count_rare(A, lane, 8); else n_surplus += delta;
                __syncwarp();
            }
            active = new_active;
            mode = new_mode;
            {
                const Particle first = part[active];
                apos = lab_position(first);
                acharge = first.charge;
                if (mode == 1) {
                    const Particle second = part[active + 1];
                    bpos = lab_position(second);
                    bcharge = second.charge;
                }
            }
            if (kind == ECMC_EVENT_SWITCH) sw = time_add(event_time, M.switch_length[0]);
            if (kind == ECMC_EVENT_END_OF_CHAIN || kind == ECMC_EVENT_SWITCH) {
                // re-created by the switcher as well: (_last_committed_event_time - time stamp) + chain_time (:203-215)
                eoc = time_add(now, time_sub(eoc_last, now) + P.chain_time);
                const StreamKey next_key = {P.seed, stream, ev};
                eoc_next = mode == 1 ? (int)stream_randbelow(next_key, ECMC_SLOT(ECMC_SLOT_END_OF_CHAIN, 0), (uint32_t)n_roots) * npr
                                     : draw_end_of_chain_active(P, next_key);
            }
            continue;
        }
        const int active_root = active / npr, active_child = active - active_root * npr;
        Time bt = time_inf();
        int bkind = ECMC_EVENT_NONE, btarget = -1, bcell = -1;
        double brate = 0.0;
        int n_cand = 0;
        bool best_from_kept = false;
        // the earliest factor candidate of this iteration (what a cell-boundary event would leave running)
        Time ft = time_inf();
        int fkind = ECMC_EVENT_NONE, ftarget = -1;
        double frate = 0.0;
        double restore_pos = 0.0, restore_root = 0.0;
        Time restore_stamp = now;
        bool restore = false;
        // lower boundary of the next cell of the ROOT unit along the direction of motion
        const int id_dir = dir == 0 ? cid0 : (dir == 1 ? cid1 : cid2);
        const int nid = id_dir + 1 == P.per_side[dir] ? 0 : id_dir + 1;
        const int next_cell = active_cell + (nid - id_dir) * P.cumulative[dir];
        const double boundary = __ldg(P.cell_min_axis + dir * P.max_per_side + nid);

        if (was_pending) {
            bkind = stp->pending_kind;
            bt.q = stp->pending_q; bt.r = stp->pending_r;
            brate = stp->pending_rate;
            if (bkind == ECMC_EVENT_PAIR || bkind == ECMC_EVENT_BOND || bkind == ECMC_EVENT_FACTOR_PAIR ||
                bkind == ECMC_EVENT_CELL_BOUNDING)
                btarget = stp->pending_target;
            else bcell = stp->pending_target;
            restore = true;
            restore_pos = stp->pending_position; restore_root = stp->pending_root_position;
            restore_stamp.q = stp->pending_stamp_q; restore_stamp.r = stp->pending_stamp_r;
        } else {
            const bool factors_kept = kept_kind != ECMC_EVENT_NONE;
            // LEAF_CELLS: is a leaf of the stored kind active (does the occupancy have an active cell)?
            const bool cell_leaf_active = LEAF_CELLS && active_child == M.cell_child - 1;
            // LEAF_CELLS: one scan position per object for the composite pairs of the factor type map (unless kept), then
            // the nearby cells and the surplus of the leaf cells
            const int lc_pair_slots = LEAF_CELLS && !factors_kept ? n_roots : 0;
            const int lc_near_slots = cell_leaf_active ? P.n_nearby + n_surplus : 0;
            const int nearby_slots = LEAF_CELLS ? 0
                : ((P.pair_handler == ECMC_PAIR_TWO_COMPOSITE_SUMMED_BOUNDING || leaf_pairs) ? P.n_nearby : 0);
            const int n_pair_slots = LEAF_CELLS ? lc_pair_slots + lc_near_slots : (nearby_slots ? nearby_slots + n_surplus : 0);
            // scan positions: [0, n_pair_slots) objects, then (cell-bounding far field: one per cell) the objects in cells
            // that are not nearby, bonds, inter-object factors (one per object and factor), bending, veto, boundary
            const bool far_objects = P.veto_enabled == ECMC_FAR_CELL_BOUNDING && (!LEAF_CELLS || cell_leaf_active);
            const int far_base = n_pair_slots;
            const int bond_base = far_base + (far_objects ? P.n_cells : 0);
            const int inter_base = bond_base + (factors_kept ? 0 : P.n_bonds);
            const int bending_base = inter_base + ((factors_kept || LEAF_CELLS) ? 0 : M.n_inter * n_roots);
            const int veto_base = bending_base + ((!factors_kept && M.bending_enabled) ? 1 : 0);
            const int boundary_base = veto_base + (P.veto_enabled == ECMC_FAR_CELL_VETO ? 1 : 0);
            // no cell system, or no active cell: no cell boundary
            const int n_scan = boundary_base + ((P.no_cells || (LEAF_CELLS && !cell_leaf_active)) ? 0 : 1);
            unsigned long long best_key = 0x7ff0000000000000ull;
            double best_x = INFINITY;
            int best_seq = kSeqNone;
            unsigned long long fbest_key = 0x7ff0000000000000ull;
            int fbest_seq = kSeqNone;
            double fbest_x = INFINITY;
            int cursor = 0;
            while (cursor < n_scan) {
                int count = 0;
                while (cursor < n_scan && count <= kItemCapacity - 96) {
                    const int s = cursor + lane;
                    int type = -1, target = -1, copies = 0;
                    if (LEAF_CELLS && s < n_pair_slots) {
                        if (s < lc_pair_slots) {
                            if (s != active_root) { type = ITEM_PAIR_LEAF; target = s; copies = npr; }
                        } else if (s - lc_pair_slots < P.n_nearby) {
                            const int code = __ldg(P.nearby + (s - lc_pair_slots));
                            int x = cid0 + (code & 1023), y = cid1 + ((code >> 10) & 1023), z = cid2 + (code >> 20);
                            if (x >= P.per_side[0]) x -= P.per_side[0];
                            if (y >= P.per_side[1]) y -= P.per_side[1];
                            if (z >= P.per_side[2]) z -= P.per_side[2];
                            target = occ[x * P.cumulative[0] + y * P.cumulative[1] + z * P.cumulative[2]];
                            if (target >= 0) { type = ITEM_NEAR_LEAF; copies = 1; }
                        } else {
                            target = sur[s - lc_pair_slots - P.n_nearby];
                            type = ITEM_NEAR_LEAF; copies = 1;
                        }
                    } else if (s < nearby_slots) {
                        const int code = __ldg(P.nearby + s);
                        int x = cid0 + (code & 1023), y = cid1 + ((code >> 10) & 1023), z = cid2 + (code >> 20);
                        if (x >= P.per_side[0]) x -= P.per_side[0];
                        if (y >= P.per_side[1]) y -= P.per_side[1];
                        if (z >= P.per_side[2]) z -= P.per_side[2];
                        target = occ[x * P.cumulative[0] + y * P.cumulative[1] + z * P.cumulative[2]];
                        if (target >= 0) { type = ITEM_PAIR_LEAF; copies = npr; }
                    } else if (s < n_pair_slots) {
                        target = sur[s - nearby_slots];
                        type = ITEM_PAIR_LEAF; copies = npr;
                    } else if (s < bond_base) {
                        // CellBoundingPotentialTagger (cell_bounding_potential_tagger.py:150-155)
                        const int cell = s - far_base;
                        if (!cell_is_nearby(P, cell, cid0, cid1, cid2) && occ[cell] >= 0) {
                            type = ITEM_FAR_OBJECT; target = cell; copies = 1;
                        }
                    } else if (s < inter_base) {
                        const int b = s - bond_base;
                        const int partner = P.bonds[b][0] == active_child ? P.bonds[b][1]
                                                                          : (P.bonds[b][1] == active_child ? P.bonds[b][0] : -1);
                        if (partner >= 0) { type = ITEM_BOND; target = active_root * npr + partner; copies = 1; }
                    } else if (s < bending_base) {
                        const int f = (s - inter_base) / n_roots, r = (s - inter_base) - f * n_roots;
                        if (M.inter[f][0] == active_child && r != active_root) {
                            type = ITEM_INTER; target = r * npr + M.inter[f][1]; copies = 1;
                        }
                    } else if (s < veto_base) {
                        bool member = false;
                        for (int i = 0; i < 3; i++) member = member || M.bending_children[i] == active_child;
                        if (member) { type = ITEM_BENDING; target = 0; copies = 1; }
                    } else if (s < boundary_base) {
                        type = ITEM_VETO; target = 0; copies = 1;
                    } else if (s < n_scan) {
                        type = ITEM_BOUNDARY; target = 0; copies = 1;
                    }
                    n_targets += (unsigned long long)__popc(__ballot_sync(kFull, type == ITEM_PAIR_LEAF || type == ITEM_FAR_OBJECT));
                    // exclusive prefix sum of `copies` over the lanes: where this lane's items go
                    int offset = copies;
#pragma unroll
                    for (int o = 1; o < 32; o <<= 1) {
                        const int v = __shfl_up_sync(kFull, offset, o);
                        if (lane >= o) offset += v;
                    }
                    const int total = __shfl_sync(kFull, offset, 31);
                    offset -= copies;
                    for (int k = 0; k < copies; k++) {
                        item_code[count + offset + k] = type | ((s * 4 + k) << 4);
                        item_target[count + offset + k] = type == ITEM_PAIR_LEAF ? target * npr + k : target;
                    }
                    count += total;
                    cursor += 32;
                }
                __syncwarp();
                for (int base = 0; base < count; base += 32) {
                    const int entry = base + lane;
                    int type = -1, target = -1, seq = kSeqNone;
                    if (entry < count) {
                        const int code = item_code[entry];
                        type = code & 15; seq = code >> 4;
                        target = item_target[entry];
                    }
                    double dt = INFINITY, rate = 0.0;
                    int kind = ECMC_EVENT_NONE, cell = -1, rec_target = -1;
                    bool is_factor = false;
                    if (type == ITEM_PAIR_LEAF || type == ITEM_INTER || type == ITEM_BOND || type == ITEM_NEAR_LEAF) {
                        const Particle tp = part[target];
                        const Vec3 s3 = separation_lab(apos, lab_position(tp), L, half);
                        const double s0 = vcomp(s3, dir);
                        const double s1 = dir == 0 ? s3.y : (dir == 1 ? s3.z : s3.x);
                        const double s2 = dir == 0 ? s3.z : (dir == 1 ? s3.x : s3.y);
                        if (type == ITEM_PAIR_LEAF) {
                            // double k of (PAIR_TIME, target object) for target leaf k
                            const int root = target / npr, k = target - root * npr;
                            const double u = stream_double(key, ECMC_SLOT(ECMC_SLOT_PAIR_TIME, root), (uint32_t)k);
                            const double du = -log_unit_interval(1.0 - u) * P.inv_beta;
                            const double c1 = P.pair_use_charge ? acharge : 1.0, c2 = P.pair_use_charge ? tp.charge : 1.0;
                            dt = displacement_time<CAND>(P.cand_potential, 0, P.inv_speed, L, s0, s1, s2, c1, c2, du);
                            kind = ECMC_EVENT_PAIR;
                            rec_target = leaf_pairs ? target : root;
                            is_factor = LEAF_CELLS;  // (a cell-boundary event of the active leaf leaves these handlers running)
                        } else if (type == ITEM_NEAR_LEAF) {
                            // TwoLeafUnitEventHandlerWithPiecewiseConstantBoundingPotential.send_event_time (:103-131) on
                            // _displacement_from_piecewise_constant_bounding_potential
                            // (event_handler_with_bounding_potential.py:282-332): bound = max(derivative now, derivative after
                            // max_displacement) + offset
                            const double u = stream_double(key, ECMC_SLOT(ECMC_SLOT_FACTOR_TIME, target), 0);
                            const double du = -log_unit_interval(1.0 - u) * P.inv_beta;
                            const double one = derivative_warp<INTER>(M.inter_potential, dir, speed, s3.x, s3.y, s3.z, 1.0, 1.0,
                                                                      trig, lane);
                            Vec3 moved = apos;
                            set_dir(moved, correct_position_entry(
                                __dadd_rn(vcomp(apos, dir), __dmul_rn(speed, M.inter_bound_max_displacement)), L));
                            const Vec3 m3 = separation_lab(moved, lab_position(tp), L, half);
                            const double two = derivative_warp<INTER>(M.inter_potential, dir, speed, m3.x, m3.y, m3.z, 1.0, 1.0,
                                                                      trig, lane);
                            const double constant = (one > two ? one : two) + M.inter_bound_offset;
                            rate = -1.0;  // bounding event rate None
                            dt = M.inter_bound_max_displacement;
                            if (constant > 0.0 && du / constant < M.inter_bound_max_displacement) { rate = constant; dt = du / constant; }
                            kind = ECMC_EVENT_FACTOR_PAIR;
                            rec_target = target;
                        } else {
                            const double u = stream_double(key, ECMC_SLOT(ECMC_SLOT_FACTOR_TIME, target), 0);
                            const double du = -log_unit_interval(1.0 - u) * P.inv_beta;
                            if (type == ITEM_BOND)
                                dt = displacement_time<BOND>(P.bond_potential, 0, P.inv_speed, L, s0, s1, s2, 1.0, 1.0,
                                                             needs_potential_change(resolve_kind<BOND>(P.bond_potential.kind)) ? du : 0.0);
                            else
                                dt = displacement_time<INTER>(M.inter_potential, 0, P.inv_speed, L, s0, s1, s2, 1.0, 1.0,
                                                              needs_potential_change(resolve_kind<INTER>(M.inter_potential.kind)) ? du : 0.0);
                            kind = type == ITEM_BOND ? ECMC_EVENT_BOND : ECMC_EVENT_FACTOR_PAIR;
                            rec_target = target;
                            is_factor = true;
                        }
                    } else if (type == ITEM_FAR_OBJECT && LEAF_CELLS) {
                        // TwoLeafUnitCellBoundingPotentialEventHandler.send_event_time
                        // (two_leaf_unit_cell_bounding_potential_event_handler.py:137-177) for the leaf in a cell that is not
                        // nearby; chargeless; the draw is double 1 of the factor-time slot of the target leaf
                        const int leaf = occ[target];
                        const int relative = relative_cell_of(P, target, cid0, cid1, cid2);
                        rate = __ldg(P.bounds + (relative * P.dimension + dir) * 2);
                        const double u = stream_double(key, ECMC_SLOT(ECMC_SLOT_FACTOR_TIME, leaf), 1);
                        const double du = -log_unit_interval(1.0 - u) * P.inv_beta;
                        dt = rate > 0.0 ? du / rate * P.inv_speed : INFINITY;
                        cell = relative;
                        kind = ECMC_EVENT_CELL_BOUNDING;
                        rec_target = leaf;
                    } else if (type == ITEM_FAR_OBJECT) {
                        // TwoCompositeObjectCellBoundingPotentialEventHandler.send_event_time
                        // (two_composite_object_cell_bounding_potential_event_handler.py:152-196): constant event rate =
                        // bound of the relative cell x the estimator's charge correction factor, active charge x
                        // max |target charges| (dipole_monte_carlo_estimator.py:158-186)
                        const int root = occ[target];
                        const int relative = relative_cell_of(P, target, cid0, cid1, cid2);
                        double charge_product = 1.0;
                        if (P.veto_use_charge) {
                            double largest = 0.0;
                            for (int k = 0; k < npr; k++) largest = fmax(largest, fabs(part[root * npr + k].charge));
                            charge_product = acharge * largest;
                        }
                        const double *bound = P.bounds + (relative * P.dimension + dir) * 2;
                        rate = charge_product > 0.0 ? __ldg(bound) * charge_product : -__ldg(bound + 1) * charge_product;
                        const double u = stream_double(key, ECMC_SLOT(ECMC_SLOT_PAIR_TIME, root), 0);
                        const double du = -log_unit_interval(1.0 - u) * P.inv_beta;
                        dt = rate > 0.0 ? du / rate * P.inv_speed : INFINITY;
                        cell = relative;
                        kind = ECMC_EVENT_CELL_BOUNDING;
                        rec_target = root;
                    } else if (type == ITEM_BENDING) {
                        // _displacement_from_piecewise_constant_bounding_potential
                        // (event_handler_with_bounding_potential.py:282-332)
                        int index = 0;
                        Vec3 unit[3], moved[3];
                        for (int i = 0; i < 3; i++) {
                            const int leaf = active_root * npr + M.bending_children[i];
                            unit[i] = leaf == active ? apos : lab_position(part[leaf]);
                            moved[i] = unit[i];
                            if (leaf == active) {
                                index = i;
                                const double x = correct_position_entry(
                                    __dadd_rn(vcomp(apos, dir), __dmul_rn(speed, M.bending_max_displacement)), L);
                                set_dir(moved[i], x);
                            }
                        }
                        const int *sp = M.bending_separations;
                        double one[3], two[3];
                        bending_derivative(M.bending_prefactor, M.bending_angle, dir, speed,
                                           separation_lab(unit[sp[0]], unit[sp[1]], L, half),
                                           separation_lab(unit[sp[2]], unit[sp[3]], L, half), one);
                        bending_derivative(M.bending_prefactor, M.bending_angle, dir, speed,
                                           separation_lab(moved[sp[0]], moved[sp[1]], L, half),
                                           separation_lab(moved[sp[2]], moved[sp[3]], L, half), two);
                        const double a = index == 0 ? one[0] : (index == 1 ? one[1] : one[2]);
                        const double b = index == 0 ? two[0] : (index == 1 ? two[1] : two[2]);
                        const double constant = (a > b ? a : b) + M.bending_offset;
                        const double u = stream_double(key, ECMC_SLOT(ECMC_SLOT_BENDING_TIME, 0), 0);
                        const double du = -log_unit_interval(1.0 - u) * P.inv_beta;
                        rate = -1.0;  // bounding event rate None
                        dt = M.bending_max_displacement;
                        if (constant > 0.0 && du / constant < M.bending_max_displacement) { rate = constant; dt = du / constant; }
                        kind = ECMC_EVENT_BENDING;
                        is_factor = true;
                    } else if (type == ITEM_VETO) {
                        // CellVetoEventHandler.send_event_time (cell_veto_event_handler.py:200-238) for the active leaf
                        // of a composite object; DipoleMonteCarloEstimator.charge_correction_factor = the active charge
                        double charge_factor = 1.0;
                        if (P.veto_use_charge) {
                            charge_factor = acharge * 1.0;
                            if (P.veto_target_charge != 1.0) charge_factor = charge_factor / P.veto_target_charge;
                        }
                        const DeviceWalker *w = &P.upper[dir];
                        if (!(charge_factor > 0.0)) { charge_factor *= -1.0; w = &P.lower[dir]; }
                        uint32_t index = 0;
                        const uint32_t e = stream_randbelow_from(key, ECMC_SLOT(ECMC_SLOT_VETO_CHOICE, 0),
                                                                 (uint32_t)w->n_entries, index);
                        const WalkerEntry entry = w->entries[e];
                        const double u0 = stream_double(key, ECMC_SLOT(ECMC_SLOT_VETO_TIME, 0), 0);
                        const double u1 = stream_double(key, ECMC_SLOT(ECMC_SLOT_VETO_TIME, 0), 1);
                        const bool first = 0.0 + (w->mean_rate - 0.0) * u0 <= entry.rate_a;
                        const int relative = first ? entry.cell_a : entry.cell_b;
                        rate = (first ? entry.bound_a : entry.bound_b) * charge_factor;
                        int tx, ty, tz;
                        const int rx = relative & 1023, ry = (relative >> 10) & 1023, rz = relative >> 20;
                        if (P.translate_modular) {
                            tx = cid0 + rx; ty = cid1 + ry; tz = cid2 + rz;
                            if (tx >= P.per_side[0]) tx -= P.per_side[0];
                            if (ty >= P.per_side[1]) ty -= P.per_side[1];
                            if (tz >= P.per_side[2]) tz -= P.per_side[2];
                        } else {
                            const int mps = P.max_per_side;
                            tx = __ldg(P.translate_axis + (0 * mps + cid0) * mps + rx);
                            ty = __ldg(P.translate_axis + (1 * mps + cid1) * mps + ry);
                            tz = __ldg(P.translate_axis + (2 * mps + cid2) * mps + rz);
                        }
                        cell = tx * P.cumulative[0] + ty * P.cumulative[1] + tz * P.cumulative[2];
                        const double exponential = -log_unit_interval(1.0 - u1) * P.inv_beta;
                        dt = exponential / (w->total_rate * charge_factor * speed);
                        kind = ECMC_EVENT_CELL_VETO;
                    } else if (type == ITEM_BOUNDARY) {
                        // the unit on the cell level: the root unit, or (LEAF_CELLS) the active leaf
                        double separation = boundary - vcomp(LEAF_CELLS ? apos : rpos, dir);
                        if (separation < 0.0) separation = separation + L;
                        dt = separation / (LEAF_CELLS ? speed : P.root_speed);
                        cell = next_cell;
                        kind = ECMC_EVENT_CELL_BOUNDARY;
                    }
                    const double x = now.r + dt;
                    const bool finite = kind != ECMC_EVENT_NONE && x < INFINITY;
                    // the reference counts one candidate per handler: the first leaf of an object stands for the pair
                    n_cand += __popc(__ballot_sync(kFull, finite && !(type == ITEM_PAIR_LEAF && (seq & 3) != 0 && !leaf_pairs)));
                    const unsigned long long k64 = finite ? time_key(x) : 0x7ff0000000000000ull;
                    const int owner = warp_argmin(k64, finite ? seq : kSeqNone, lane);
                    const unsigned long long pass_key = __shfl_sync(kFull, k64, owner);
                    const int pass_seq = __shfl_sync(kFull, finite ? seq : kSeqNone, owner);
                    if (pass_key < best_key || (pass_key == best_key && pass_seq < best_seq)) {
                        best_key = pass_key; best_seq = pass_seq;
                        best_x = __shfl_sync(kFull, x, owner);
                        bkind = __shfl_sync(kFull, kind, owner);
                        btarget = __shfl_sync(kFull, rec_target, owner);
                        bcell = __shfl_sync(kFull, cell, owner);
                        brate = __shfl_sync(kFull, rate, owner);
                    }
                    // the earliest of the leaf-level factors, separately (kept if a cell-boundary event wins)
                    const bool ffinite = finite && is_factor;
                    if (__any_sync(kFull, ffinite)) {
                        const unsigned long long f64 = ffinite ? k64 : 0x7ff0000000000000ull;
                        const int fowner = warp_argmin(f64, ffinite ? seq : kSeqNone, lane);
                        const unsigned long long fkey = __shfl_sync(kFull, f64, fowner);
                        const int fseq = __shfl_sync(kFull, ffinite ? seq : kSeqNone, fowner);
                        if (fkey < fbest_key || (fkey == fbest_key && fseq < fbest_seq)) {
                            fbest_key = fkey; fbest_seq = fseq;
                            fbest_x = __shfl_sync(kFull, x, fowner);
                            fkind = __shfl_sync(kFull, kind, fowner);
                            ftarget = __shfl_sync(kFull, rec_target, fowner);
                            frate = __shfl_sync(kFull, rate, fowner);
                        }
                    }
                }
                __syncwarp();
            }
            if (best_seq != kSeqNone) {
                const double fl = floor(best_x);
                bt.q = now.q + fl; bt.r = best_x - fl;
            } else {
                bkind = ECMC_EVENT_NONE;
            }
            if (fbest_seq != kSeqNone) {
                const double fl = floor(fbest_x);
                ft.q = now.q + fl; ft.r = fbest_x - fl;
            }
            if (kept_kind > 0 && time_lt(kept_time, bt)) {
                // a factor handler that a cell-boundary event left running fires first: its in-state is the old one
                bt = kept_time; bkind = kept_kind; btarget = kept_target; bcell = -1; brate = kept_rate;
                best_from_kept = true;
            }
        }

        n_cand++;
        const bool eoc_first = time_lt(eoc, bt);
        Time event_time = eoc_first ? eoc : bt;
        int kind = eoc_first ? ECMC_EVENT_END_OF_CHAIN : bkind;
        if (ROOT_MODE) {
            // the leaf-to-root RootLeafUnitActiveSwitcher, in the scheduler since the last switch
            n_cand++;
            if (time_lt(sw, event_time)) { event_time = sw; kind = ECMC_EVENT_SWITCH; }
        }
        if (!time_lt(event_time, until)) {
            if (lane == 0) {
                stp->pending_kind = bkind;
                stp->pending_q = bt.q; stp->pending_r = bt.r;
                stp->pending_rate = brate;
                stp->pending_target = (bkind == ECMC_EVENT_PAIR || bkind == ECMC_EVENT_BOND || bkind == ECMC_EVENT_FACTOR_PAIR ||
                                       bkind == ECMC_EVENT_CELL_BOUNDING) ? btarget : bcell;
                if (!was_pending) {
                    stp->pending_position = best_from_kept ? kept_pos : vcomp(apos, dir);
                    stp->pending_root_position = best_from_kept ? kept_root : vcomp(rpos, dir);
                    stp->pending_stamp_q = best_from_kept ? kept_stamp.q : now.q;
                    stp->pending_stamp_r = best_from_kept ? kept_stamp.r : now.r;
                }
            }
            stopped_by_time = true;
            if (ALIGNED) { done = true; continue; }
            break;
        }
        if (was_pending) {
            if (lane == 0) stp->pending_kind = ECMC_EVENT_NONE;
        } else if (best_from_kept) {
            restore = true;
            restore_pos = kept_pos; restore_root = kept_root; restore_stamp = kept_stamp;
        }
        if (restore && kind != ECMC_EVENT_END_OF_CHAIN && !(ROOT_MODE && kind == ECMC_EVENT_SWITCH)) {
            set_dir(apos, restore_pos);
            set_dir(rpos, restore_root);
            now = restore_stamp;
        }
        // which handlers survive this event: a cell-boundary event of the root leaves the leaf-level factors running
        if (kind == ECMC_EVENT_CELL_BOUNDARY && M.boundary_keeps_factors) {
            if (kept_kind == ECMC_EVENT_NONE && !was_pending) {
                kept_kind = fkind == ECMC_EVENT_NONE ? -1 : fkind;
                kept_target = ftarget; kept_time = ft; kept_rate = frate;
                kept_pos = vcomp(apos, dir); kept_root = vcomp(rpos, dir); kept_stamp = now;
            }
        } else {
            kept_kind = ECMC_EVENT_NONE;
        }
        was_pending = false;

        // ---- out-state: time slice of the active leaf and its root (abstracts.py:82-101)
        {
            const double dt = time_sub(event_time, now);
            set_dir(apos, correct_position_entry(__dadd_rn(vcomp(apos, dir), __dmul_rn(speed, dt)), L));
            set_dir(rpos, correct_position_entry(__dadd_rn(vcomp(rpos, dir), __dmul_rn(P.root_speed, dt)), L));
            now = event_time;
        }
        int new_active = active, rec_target = -1;
        uint32_t draw = 0;
        switch (kind) {
        case ECMC_EVENT_PAIR:
        case ECMC_EVENT_CELL_BOUNDING:
        case ECMC_EVENT_CELL_VETO: {
            // far: TwoCompositeObjectCellBoundingPotentialEventHandler.send_out_state (:198-246) -- the cell veto's
            // out-state with a known target object; the stored rate is per length, the bound's derivative is rate x speed
            const bool far = kind == ECMC_EVENT_CELL_BOUNDING;
            const bool veto = kind == ECMC_EVENT_CELL_VETO || far;
            if (LEAF_CELLS && far) {
                // TwoLeafUnitCellBoundingPotentialEventHandler.send_out_state (:179-211) between two leaves: the stored
                // rate x speed is the bounding event rate, confirmed against the real potential; the target leaf takes
                // over. Recorded with the target's object.
                rec_target = btarget / npr;
                n_pair++;
                const double real = pair_derivative_lab<-1>(P.veto_potential, dir, speed, apos, lab_position(part[btarget]), 1.0,
                                                            1.0, L, half, trig, lane);
                const double bounding_rate = brate * speed;
                if (real > 0.0) {
                    if (bounding_rate < real) count_rare(A, lane, 7);
                    const double u = confirm_draw(key.seed, key.stream, key.event, draw++);
                    if (0.0 + (bounding_rate - 0.0) * u < real) new_active = btarget;
                }
                break;
            }
            // leaf_pairs: btarget is the target LEAF; the loops below then run over that one leaf only
            const bool one_leaf = leaf_pairs && !veto;
            const int target_root = far ? btarget : (veto ? occ[bcell] : (one_leaf ? btarget / npr : btarget));
            rec_target = one_leaf ? btarget : target_root;
            if (veto && !far) n_veto++; else n_pair++;
            if (target_root < 0) break;
            const int k_first = one_leaf ? btarget - target_root * npr : 0, k_last = one_leaf ? k_first + 1 : npr;
            const bool use_charge = veto ? P.veto_use_charge : P.pair_use_charge;
            double bounding_rate = far ? brate * speed : (veto ? brate : 0.0);
            double factor_derivative = 0.0;
            double target_derivatives[4] = {0.0, 0.0, 0.0, 0.0};
            Vec3 tpos[4];
            double tcharge[4];
            const PotentialParams &real_potential = veto ? P.veto_potential : P.real_potential;
            for (int k = k_first; k < k_last; k++) {
                const Particle tp = part[target_root * npr + k];
                tpos[k] = lab_position(tp);
                tcharge[k] = tp.charge;
                if (!veto) {
                    const double c1 = use_charge ? acharge : 1.0, c2 = use_charge ? tp.charge : 1.0;
                    const double b = pair_derivative_lab<CAND>(P.cand_potential, dir, speed, apos, tpos[k], c1, c2, L, half, trig, lane);
                    bounding_rate += b > 0.0 ? b : 0.0;
                }
            }
            // The derivatives of the real potential, all through ONE call site (the merged-image Coulomb sum is a
            // thousand instructions): first the active leaf against the target leaves (confirmation), then -- only if
            // the event is confirmed -- the other local leaves against the target leaves (_fill_lifting,
            // event_handler_with_bounding_potential.py:170-220)
            double local_derivatives[4] = {0.0, 0.0, 0.0, 0.0};
            bool confirmed = false;
            const bool three_sums = REAL == ECMC_POT_MERGED_IMAGE_COULOMB && !one_leaf && npr == 3 &&
                                    real_potential.kind == ECMC_POT_MERGED_IMAGE_COULOMB &&
                                    18 * (real_potential.mic.fourier_cutoff + 1) <= kMoleculeTrigDoubles;
            for (int i = -1; i < npr; i++) {
                const int local = i < 0 ? active : active_root * npr + i;
                if (i >= 0 && local == active) continue;
                if (i == 0 || (i == 1 && active == active_root * npr)) {
                    // between the two stages: confirm the event with the summed derivative of the active leaf
                    const double event_rate = factor_derivative > 0.0 ? factor_derivative : 0.0;
                    if (bounding_rate < event_rate) count_rare(A, lane, 7);
                    const double u = confirm_draw(key.seed, key.stream, key.event, draw++);
                    if (event_rate <= 0.0 + (bounding_rate - 0.0) * u) break;
                    confirmed = true;
                    // TwoLeafUnitBoundingPotentialEventHandler.send_out_state (:148-168): the target leaf takes over
                    if (one_leaf) break;
                    local_derivatives[active - active_root * npr] = veto ? factor_derivative : event_rate;
                }
                const Particle lp = part[local];
                const Vec3 lpos = i < 0 ? apos : lab_position(lp);
                const double lcharge = i < 0 ? acharge : lp.charge;
                if (REAL == ECMC_POT_MERGED_IMAGE_COULOMB && three_sums) {
                    // one local leaf against the three leaves of the target molecule: the three Ewald sums side by side,
                    // multiplied as derivative_warp does (prefactor x charges x sum x speed)
                    Vec3 s3[3];
                    for (int j = 0; j < 3; j++) {
                        const Vec3 lab = separation_lab(lpos, tpos[j], L, half);
                        s3[j].x = vcomp(lab, dir);
                        s3[j].y = dir == 0 ? lab.y : (dir == 1 ? lab.z : lab.x);
                        s3[j].z = dir == 0 ? lab.z : (dir == 1 ? lab.x : lab.y);
                    }
                    double sums[3];
                    mic_derivative_warp3(real_potential.mic, s3[0].x, s3[0].y, s3[0].z, s3[1].x, s3[1].y, s3[1].z, s3[2].x,
                                         s3[2].y, s3[2].z, trig, lane, sums[0], sums[1], sums[2]);
                    for (int j = 0; j < 3; j++) {
                        const double c1 = use_charge ? lcharge : 1.0, c2 = use_charge ? tcharge[j] : 1.0;
                        const double pairwise = real_potential.mic.prefactor * c1 * c2 * sums[j] * speed;
                        if (i < 0) factor_derivative += pairwise; else local_derivatives[i] += pairwise;
                        target_derivatives[j] -= pairwise;
                    }
                    continue;
                }
                for (int j = k_first; j < k_last; j++) {
                    const double c1 = use_charge ? lcharge : 1.0, c2 = use_charge ? tcharge[j] : 1.0;
                    const double pairwise = pair_derivative_lab<REAL>(real_potential, dir, speed, lpos, tpos[j], c1, c2, L,
                                                                      half, trig, lane);
                    if (i < 0) factor_derivative += pairwise; else local_derivatives[i] += pairwise;
                    target_derivatives[j] -= pairwise;
                }
            }
            if (!confirmed) break;
            if (one_leaf) { new_active = btarget; break; }
            Lifting lift;
            lifting_reset(lift);
            for (int pass = 0; pass < 2; pass++) {
                const bool local_now = (active_root < target_root) == (pass == 0);
                for (int i = 0; i < npr; i++) {
                    if (local_now)
                        lifting_insert(lift, local_derivatives[i], active_root * npr + i, active_root * npr + i == active, key, draw);
                    else
                        lifting_insert(lift, target_derivatives[i], target_root * npr + i, false, key, draw);
                }
            }
            new_active = lifting_get(lift, M.composite_lifting, key, draw);
            if (veto && !far) count_rare(A, lane, 3);
            break;
        }
        case ECMC_EVENT_BOND:
        case ECMC_EVENT_FACTOR_PAIR:
            rec_target = btarget;
            if (LEAF_CELLS && kind == ECMC_EVENT_FACTOR_PAIR) {
                // TwoLeafUnitEventHandlerWithPiecewiseConstantBoundingPotential.send_out_state (:133-152)
                n_factor++;
                if (brate < 0.0) break;  // bounding event rate None
                const double real = pair_derivative_lab<INTER>(M.inter_potential, dir, speed, apos, lab_position(part[btarget]),
                                                               1.0, 1.0, L, half, trig, lane);
                if (real > 0.0) {
                    if (brate < real) count_rare(A, lane, 7);
                    const double u = confirm_draw(key.seed, key.stream, key.event, draw++);
                    if (0.0 + (brate - 0.0) * u < real) new_active = btarget;
                }
                break;
            }
            new_active = btarget;
            if (kind == ECMC_EVENT_BOND) n_bond++; else n_factor++;
            break;
        case ECMC_EVENT_BENDING: {
            n_bond++;
            if (brate < 0.0) break;
            int index = 0;
            Vec3 unit[3];
            for (int i = 0; i < 3; i++) {
                const int leaf = active_root * npr + M.bending_children[i];
                unit[i] = leaf == active ? apos : lab_position(part[leaf]);
                if (leaf == active) index = i;
            }
            const int *sp = M.bending_separations;
            double derivatives[3];
            bending_derivative(M.bending_prefactor, M.bending_angle, dir, speed,
                               separation_lab(unit[sp[0]], unit[sp[1]], L, half),
                               separation_lab(unit[sp[2]], unit[sp[3]], L, half), derivatives);
            const double own = index == 0 ? derivatives[0] : (index == 1 ? derivatives[1] : derivatives[2]);
            if (own > 0.0) {
                if (brate < own) count_rare(A, lane, 7);
                const double u = confirm_draw(key.seed, key.stream, key.event, draw++);
                if (0.0 + (brate - 0.0) * u < own) {
                    Lifting lift;
                    lifting_reset(lift);
                    for (int i = 0; i < 3; i++)
                        lifting_insert(lift, derivatives[i], active_root * npr + M.bending_children[i], i == index, key, draw);
                    new_active = lifting_get(lift, M.bending_lifting, key, draw);
                }
            }
            break;
        }
        case ECMC_EVENT_CELL_BOUNDARY: {
            // the unit on the cell level -- the root, or (LEAF_CELLS) the active leaf -- lands exactly on the lower boundary
            // of its new cell (cell_boundary_event_handler.py:158-173)
            const double landing = __ldg(P.cell_min_axis + dir * P.max_per_side + (bcell / P.cumulative[dir]) % P.per_side[dir]);
            if (LEAF_CELLS) set_dir(apos, landing); else set_dir(rpos, landing);
            n_boundary++;
            break;
        }
        case ECMC_EVENT_END_OF_CHAIN:
            new_active = eoc_next;
            rec_target = new_active;
            if (ROOT_MODE) eoc_last = event_time;
            n_eoc++;
            break;
        case ECMC_EVENT_SWITCH:
            // _send_out_state_root_unit_active (root_leaf_unit_active_switcher.py:171-208): the root unit takes over, the
            // other leaf takes the velocity and the time stamp of the active one; `active` becomes the first leaf
            new_active = active_root * npr;
            break;
        default: break;
        }
        const bool to_root = ROOT_MODE && kind == ECMC_EVENT_SWITCH;
        if (new_active < 0) { count_rare(A, lane, 8); new_active = active; }
        if (RECORD && lane == 0 && (int)n_events < A.records_per_chain) {
            EcmcEventRecord rec;
            rec.kind = kind; rec.target = rec_target; rec.target_cell = (kind == ECMC_EVENT_END_OF_CHAIN || to_root) ? -1 : bcell;
            rec.accepted = (kind == ECMC_EVENT_END_OF_CHAIN || to_root) ? 1 : (new_active != active);
            rec.n_candidates = n_cand;
            rec.new_active = new_active;
            rec.new_direction = kind == ECMC_EVENT_END_OF_CHAIN ? (dir + 1) % P.dimension : dir;
            rec.mode = to_root ? 1 : 0;
            rec.time_q = event_time.q; rec.time_r = event_time.r;
            rec.active_pos[0] = apos.x; rec.active_pos[1] = apos.y; rec.active_pos[2] = apos.z;
            A.records[(size_t)chain * A.records_per_chain + n_events] = rec;
        }
        ev++;
        n_events++;
        n_candidates += (unsigned long long)n_cand;
        if (kind == ECMC_EVENT_END_OF_CHAIN) dir = dir + 1 == P.dimension ? 0 : dir + 1;

        // ---- commit + SingleActiveCellOccupancy.update on the cell level of the roots
        const int new_root = new_active / npr;
        if (LEAF_CELLS && new_active != active && active_child == M.cell_child - 1) {
            // the cells hold one kind of leaf: the old active leaf goes back into its cell if it is of that kind
            // (single_active_cell_occupancy.py:149-203)
            int delta = 0;
            if (lane == 0) delta = occupancy_insert(occ, sur, n_surplus, 1, P.max_surplus, active_cell, active);
            delta = __shfl_sync(kFull, delta, 0);
            if (delta == 2) count_rare(A, lane, 8); else n_surplus += delta;
            __syncwarp();
        }
        const bool active_changed = new_active != active;
        if (new_active != active) {
            if (lane == 0) {
                Particle p = part[active];
                p.x = apos.x; p.y = apos.y; p.z = apos.z;
                part[active] = p;
            }
            __syncwarp();
            const Particle np = part[new_active];
            apos = lab_position(np);
            acharge = np.charge;
        }
        if (new_root != active_root) {
            int delta = 0;
            if (lane == 0) {
                Particle r = roots[active_root];
                r.x = rpos.x; r.y = rpos.y; r.z = rpos.z;
                roots[active_root] = r;
                if (!LEAF_CELLS) delta = occupancy_insert(occ, sur, n_surplus, 1, P.max_surplus, active_cell, active_root);
            }
            delta = __shfl_sync(kFull, delta, 0);
            if (delta == 2) count_rare(A, lane, 8); else n_surplus += delta;
            __syncwarp();
            rpos = lab_position(roots[new_root]);
        }
        active = new_active;
        if (LEAF_CELLS) {
            // the active cell follows the active leaf while it is of the stored kind; a new active leaf of that kind leaves
            // its cell
            const bool relevant = active - new_root * npr == M.cell_child - 1;
            if (relevant) {
                cid0 = (int)(apos.x / P.side_length[0]);
                cid1 = (int)(apos.y / P.side_length[1]);
                cid2 = (int)(apos.z / P.side_length[2]);
                active_cell = cid0 * P.cumulative[0] + cid1 * P.cumulative[1] + cid2 * P.cumulative[2];
                if (active_changed) {
                    int delta = 0;
                    if (lane == 0) delta = occupancy_remove(occ, sur, n_surplus, 1, active_cell, active);
                    delta = __shfl_sync(kFull, delta, 0);
                    if (delta == 2) count_rare(A, lane, 8); else n_surplus += delta;
                    __syncwarp();
                }
            } else if (active_changed) {
                active_cell = 0; cid0 = cid1 = cid2 = 0;
            }
        } else {
            // the cell of the root is recomputed from its position after every event, like the oracle does
            const int old_cell = active_cell;
            cid0 = (int)(rpos.x / P.side_length[0]);
            cid1 = (int)(rpos.y / P.side_length[1]);
            cid2 = (int)(rpos.z / P.side_length[2]);
            active_cell = cid0 * P.cumulative[0] + cid1 * P.cumulative[1] + cid2 * P.cumulative[2];
            (void)old_cell;
        }
        if (!LEAF_CELLS && new_root != active_root) {
            int delta = 0;
            if (lane == 0) delta = occupancy_remove(occ, sur, n_surplus, 1, active_cell, new_root);
            delta = __shfl_sync(kFull, delta, 0);
            if (delta == 2) count_rare(A, lane, 8); else n_surplus += delta;
            __syncwarp();
        }
        if (to_root) {
            // the root unit is active from here on: the root-to-leaf switcher is created with the root's time stamp
            mode = 1;
            sw = time_add(event_time, M.switch_length[1]);
            const Particle second = part[active + 1];
            bpos = lab_position(second);
            bcharge = second.charge;
        }
        if (kind == ECMC_EVENT_END_OF_CHAIN || to_root) {
            // a switcher event re-creates the candidate between two ends of chain: (_last_committed_event_time - time
            // stamp) + chain_time, single_independent_active_periodic_direction_end_of_chain_event_handler.py:203-215
            eoc = time_add(now, (ROOT_MODE ? time_sub(eoc_last, now) : time_sub(now, now)) + P.chain_time);
            const StreamKey next_key = {P.seed, stream, ev};
            eoc_next = to_root ? (int)stream_randbelow(next_key, ECMC_SLOT(ECMC_SLOT_END_OF_CHAIN, 0), (uint32_t)n_roots) * npr
                               : draw_end_of_chain_active(P, next_key);
        }
    }

    if (stopped_by_time) {
        const double dt = time_sub(until, now);
        const bool whole = ROOT_MODE && mode == 1;  // the root unit is active: root and leaves with the full velocity
        set_dir(apos, correct_position_entry(__dadd_rn(vcomp(apos, dir), __dmul_rn(speed, dt)), L));
        if (whole) set_dir(bpos, correct_position_entry(__dadd_rn(vcomp(bpos, dir), __dmul_rn(speed, dt)), L));
        set_dir(rpos, correct_position_entry(__dadd_rn(vcomp(rpos, dir), __dmul_rn(whole ? speed : P.root_speed, dt)), L));
        now = until;
    }
    if (lane == 0) {
        Particle p = part[active];
        p.x = apos.x; p.y = apos.y; p.z = apos.z;
        part[active] = p;
        if (ROOT_MODE) {
            if (mode == 1) {
                p = part[active + 1];
                p.x = bpos.x; p.y = bpos.y; p.z = bpos.z;
                part[active + 1] = p;
            }
            stp->mode = mode;
            stp->switch_q = sw.q; stp->switch_r = sw.r;
            stp->eoc_last_q = eoc_last.q; stp->eoc_last_r = eoc_last.r;
        }
        Particle r = roots[active / npr];
        r.x = rpos.x; r.y = rpos.y; r.z = rpos.z;
        roots[active / npr] = r;
        stp->active = active; stp->direction = dir;
        stp->time_q = now.q; stp->time_r = now.r;
        stp->eoc_q = eoc.q; stp->eoc_r = eoc.r;
        stp->eoc_next_active = eoc_next; stp->active_cell = active_cell;
        stp->event_counter = ev;
        stp->kept_kind = kept_kind; stp->kept_target = kept_target;
        stp->kept_q = kept_time.q; stp->kept_r = kept_time.r;
        stp->kept_rate = kept_rate; stp->kept_position = kept_pos; stp->kept_root_position = kept_root;
        stp->kept_stamp_q = kept_stamp.q; stp->kept_stamp_r = kept_stamp.r;
        S.n_surplus[chain] = n_surplus;
        if (A.stats) {
            unsigned long long *st = reinterpret_cast<unsigned long long *>(A.stats);
            if (n_events) atomicAdd(st + 0, (unsigned long long)n_events);
            if (n_pair) atomicAdd(st + 1, (unsigned long long)n_pair);
            if (n_veto) atomicAdd(st + 2, (unsigned long long)n_veto);
            if (n_boundary) atomicAdd(st + 4, (unsigned long long)n_boundary);
            if (n_eoc) atomicAdd(st + 5, (unsigned long long)n_eoc);
            if (n_candidates) atomicAdd(st + 6, n_candidates);
            if (n_bond) atomicAdd(st + 9, (unsigned long long)n_bond);
            if (n_factor) atomicAdd(st + 10, (unsigned long long)n_factor);
            if (n_targets) atomicAdd(st + 11, n_targets);
        }
    }
}

// start of run for molecules: the cells hold the ROOT units (SingleActiveCellOccupancy.initialize with cell_level = 1,
// single_active_cell_occupancy.py:95-121), the active unit on the cell level is the root of the initial active leaf
template <int WARPS>
__global__ void __launch_bounds__(WARPS * 32)
molecule_start_kernel(const __grid_constant__ DeviceProgram P, const DeviceState S, const uint32_t *streams,
                      uint32_t first_stream, int initial_active, int initial_direction, EcmcStats *stats,
                      double first_switch, int cell_child) {
    const int lane = threadIdx.x & 31;
    const int chain = S.first_chain + blockIdx.x * WARPS + (threadIdx.x >> 5);
    if (chain >= S.first_chain + S.n_chains) return;
    const int npr = P.nodes_per_root, n_roots = P.n_particles / npr;
    const Particle *roots = S.roots + (size_t)chain * n_roots;
    int *occ = S.occupants + (size_t)chain * P.n_cells;
    int *sur = S.surplus + (size_t)chain * P.max_surplus;
    for (int i = lane; i < P.n_cells; i += 32) occ[i] = -1;
    __syncwarp();
    if (lane == 0) {
        int n_surplus = 0, overflow = 0;
        // EcmcProgram.cell_child: the cells hold only the leaves with child index cell_child - 1, by their own positions
        // (SingleActiveCellOccupancy with cell level 2 and a charge indicator, :62-121), and have an active cell only while
        // such a leaf is active
        const Particle *leaves = S.particles + (size_t)chain * P.n_particles;
        for (int r = 0; r < n_roots; r++) {
            int id[3];
            cell_identifier_of(P, cell_child ? leaves[r * npr + cell_child - 1] : roots[r], id);
            const int delta = occupancy_insert(occ, sur, n_surplus, 1, P.max_surplus, flat_cell(P, id),
                                               cell_child ? r * npr + cell_child - 1 : r);
            if (delta == 2) overflow++; else n_surplus += delta;
        }
        EcmcChainState st = {};  // every field defined: the state is downloaded, compared and checkpointed as bytes
        st.active = initial_active; st.direction = initial_direction;
        st.time_q = 0.0; st.time_r = 0.0;
        st.event_counter = 0;
        st.stream = streams ? streams[chain] : first_stream + (uint32_t)chain;
        int id[3];
        if (cell_child && initial_active % npr != cell_child - 1) {
            st.active_cell = 0;
        } else {
            cell_identifier_of(P, cell_child ? leaves[initial_active] : roots[initial_active / npr], id);
            st.active_cell = flat_cell(P, id);
            const int delta = occupancy_remove(occ, sur, n_surplus, 1, st.active_cell,
                                               cell_child ? initial_active : initial_active / npr);
            if (delta == 2) overflow++; else n_surplus += delta;
        }
        const Time now = {0.0, 0.0};
        const Time eoc = time_add(now, time_sub(now, now) + P.chain_time);
        st.eoc_q = eoc.q; st.eoc_r = eoc.r;
        const StreamKey key = {P.seed, st.stream, 0ull};
        st.eoc_next_active = draw_end_of_chain_active(P, key);
        st.pending_kind = ECMC_EVENT_NONE; st.pending_target = 0; st.mode = 0;
        st.pending_q = 0.0; st.pending_r = 0.0; st.pending_rate = 0.0; st.pending_position = 0.0;
        st.pending_stamp_q = 0.0; st.pending_stamp_r = 0.0; st.pending_root_position = 0.0;
        st.kept_kind = ECMC_EVENT_NONE; st.kept_target = 0; st.kept_q = 0.0; st.kept_r = 0.0; st.kept_rate = 0.0;
        st.kept_position = 0.0; st.kept_root_position = 0.0; st.kept_stamp_q = 0.0; st.kept_stamp_r = 0.0;
        if (first_switch > 0.0) {
            // EcmcProgram.root_mode: the leaf-to-root switcher is created at the start of the run, time stamp of the root
            // unit + chain length (root_leaf_unit_active_switcher.py:102-127)
            const Time first = time_add(now, first_switch);
            st.switch_q = first.q; st.switch_r = first.r;
        }
        S.chains[chain] = st;
        S.n_surplus[chain] = n_surplus;
        if (overflow && stats) atomicAdd(reinterpret_cast<unsigned long long *>(stats) + 8, (unsigned long long)overflow);
    }
}

}  // namespace ecmc
