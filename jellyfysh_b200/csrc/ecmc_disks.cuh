// ecmc_disks.cuh -- general velocities: two-dimensional hard disks tethered into dipoles, no cell system, the velocity
// of the chain rotated by a fixed angle at every end of chain (the shipped hard_disk_dipoles/hard_disk_dipoles.ini and
// single_hard_disk_dipole.ini).
//
// One warp per chain. Per event (= one iteration of single_process_mediator.py:91-156):
//   lanes  <- the hard-sphere factors between the active leaf and the leaves of every other object (factor type map
//             entries between different objects, factor_type_maps.py:333-347; TwoLeafUnitEventHandler.send_event_time,
//             two_leaf_unit_event_handler.py:105-138 with HardSpherePotential.displacement for a GENERAL velocity,
//             hard_sphere_potential.py:65-99), then the tether with the partner leaf (hard_dipole_potential.py:75-114);
//             every lane keeps the earliest of its share, the warp takes the argmin (heap order: time, then scan order)
//   the end-of-chain candidate persists in the chain state: SingleIndependentActiveSequentialDirectionEndOfChainEventHandler
//             (single_independent_active_sequential_direction_end_of_chain_event_handler.py:64-122) rotates the velocity
//   out-state: hard potentials always lift (two_leaf_unit_event_handler.py:140-154); the velocity moves from the old
//             active leaf to the new one, the root units follow with the leaf weight as the reference accumulates it
//             (_register_velocity_change_leaf_cnode / _commit_sub_tree_non_leaf_velocity_change, abstracts.py:165-227).
// Hard potentials draw no random numbers; only the end of chain does (the next active leaf).
#pragma once

#include "ecmc_kernels.cuh"

namespace ecmc {

struct DiskProgram {
    int n_inter;
    int inter[ECMC_MAX_INTER_FACTORS][2];
    int inter_kind, bond_kind;             // ECMC_POT_HARD_SPHERE / ECMC_POT_HARD_DIPOLE
    double inter_p0, inter_p1, bond_p0, bond_p1;
    double eoc_cos, eoc_sin;
    double weight;                          // of a leaf in its root unit: 1 / nodes_per_root
    int initial_direction, pad;
};

// order-preserving 64-bit key of any double that is not NaN (negative times sort before positive ones)
ECMC_D unsigned long long ordered_key(double x) {
    const long long bits = __double_as_longlong(x);
    return bits < 0 ? ~(unsigned long long)bits : (unsigned long long)bits | 0x8000000000000000ull;
}

ECMC_D double hard_time(int kind, double p0, double p1, double vv, double vs, double ss) {
    return kind == ECMC_POT_HARD_SPHERE ? hard_sphere_time(p0, vv, vs, ss) : hard_dipole_time(p0, p1, vv, vs, ss);
}

template <bool RECORD, int WARPS>
__global__ void __launch_bounds__(WARPS * 32)
disk_kernel(const __grid_constant__ DeviceProgram P, const __grid_constant__ DiskProgram K, const DeviceState S, const RunArgs A) {
    const int lane = threadIdx.x & 31;
    const int chain = S.first_chain + blockIdx.x * WARPS + (threadIdx.x >> 5);
    if (chain >= S.first_chain + S.n_chains) return;
    const int npr = P.nodes_per_root, n_roots = P.n_particles / npr;
    Particle *part = S.particles + (size_t)chain * P.n_particles;
    Particle *roots = S.roots + (size_t)chain * n_roots;
    EcmcChainState *stp = S.chains + chain;

    int active = stp->active;
    Time now = {stp->time_q, stp->time_r};
    Time eoc = {stp->eoc_q, stp->eoc_r};
    int eoc_next = stp->eoc_next_active;
    unsigned long long ev = stp->event_counter;
    const uint32_t stream = stp->stream;
    bool was_pending = stp->pending_kind != ECMC_EVENT_NONE;
    double vx = stp->velocity[0], vy = stp->velocity[1];
    double rvx = stp->root_velocity[0], rvy = stp->root_velocity[1];
    Particle a = part[active];
    Particle ar = roots[active / npr];

    const Time until = {A.until_q, A.until_r};
    const double L = P.length, half = P.half_length;
    const unsigned max_events = A.max_events > 0 ? (unsigned)min(A.max_events, 0x7fffffffLL) : 0x7fffffffu;
    const int n_factor_slots = n_roots * K.n_inter, n_slots = n_factor_slots + P.n_bonds;
    unsigned n_events = 0, n_bond = 0, n_factor = 0, n_eoc = 0;
    unsigned long long n_candidates = 0, n_targets = 0;
    bool stopped_by_time = false;

    while (n_events < max_events) {
        Time bt = time_inf();
        int bkind = ECMC_EVENT_NONE, btarget = -1, n_cand = 0;
        double kept_x = 0.0, kept_y = 0.0, kept_rx = 0.0, kept_ry = 0.0;
        Time kept_stamp = now;
        if (was_pending) {
            // a candidate that survived a host control event: nothing is recomputed
            bkind = stp->pending_kind;
            bt.q = stp->pending_q; bt.r = stp->pending_r;
            btarget = stp->pending_target;
            kept_x = stp->pending_position; kept_y = stp->pending_position_y;
            kept_rx = stp->pending_root_position; kept_ry = stp->pending_root_position_y;
            kept_stamp.q = stp->pending_stamp_q; kept_stamp.r = stp->pending_stamp_r;
        } else {
            const int active_root = active / npr, active_child = active - active_root * npr;
            const double vv = dot3(vx, vy, 0.0, vx, vy, 0.0);
            double best_x = INFINITY;
            int best_seq = kSeqNone, best_target = -1, best_kind = ECMC_EVENT_NONE, count = 0, evaluated = 0;
            for (int s = lane; s < n_slots; s += 32) {
                int target = -1, kind = ECMC_EVENT_NONE;
                if (s < n_factor_slots) {
                    const int root = s / K.n_inter, f = s - root * K.n_inter;
                    if (root != active_root && K.inter[f][0] == active_child) {
                        target = root * npr + K.inter[f][1];
                        kind = ECMC_EVENT_FACTOR_PAIR;
                    }
                } else {
                    const int b = s - n_factor_slots;
                    const int partner = P.bonds[b][0] == active_child ? P.bonds[b][1]
                                                                       : (P.bonds[b][1] == active_child ? P.bonds[b][0] : -1);
                    if (partner >= 0) {
                        target = active_root * npr + partner;
                        kind = ECMC_EVENT_BOND;
                    }
                }
                if (target < 0) continue;
                evaluated++;
                const Particle q = part[target];
                const double sx = correct_separation_in_box(q.x - a.x, L, half);
                const double sy = correct_separation_in_box(q.y - a.y, L, half);
                const double vs = dot3(vx, vy, 0.0, sx, sy, 0.0), ss = dot3(sx, sy, 0.0, sx, sy, 0.0);
                const double dt = kind == ECMC_EVENT_FACTOR_PAIR ? hard_time(K.inter_kind, K.inter_p0, K.inter_p1, vv, vs, ss)
                                                                 : hard_time(K.bond_kind, K.bond_p0, K.bond_p1, vv, vs, ss);
                if (!isinf(dt)) count++;  // heap_scheduler.py:139
                const double x = now.r + dt;  // Time.__add__: orders the candidates of one event (see time_key)
                if (x < best_x) { best_x = x; best_seq = s; best_target = target; best_kind = kind; }
            }
            n_cand = __reduce_add_sync(kFull, count);
            n_targets += (unsigned long long)__reduce_add_sync(kFull, evaluated);  // handler calls of the reference
            const bool have = best_seq != kSeqNone;
            const int owner = warp_argmin(have ? ordered_key(best_x) : ~0ull, best_seq, lane);
            best_seq = __shfl_sync(kFull, best_seq, owner);
            if (best_seq != kSeqNone) {
                best_x = __shfl_sync(kFull, best_x, owner);
                bkind = __shfl_sync(kFull, best_kind, owner);
                btarget = __shfl_sync(kFull, best_target, owner);
                const double fl = floor(best_x);
                bt.q = now.q + fl; bt.r = best_x - fl;
            }
        }

        n_cand++;  // the end of chain
        const bool eoc_first = time_lt(eoc, bt);
        const Time event_time = eoc_first ? eoc : bt;
        const int kind = eoc_first ? ECMC_EVENT_END_OF_CHAIN : bkind;
        if (!time_lt(event_time, until)) {
            if (lane == 0) {
                stp->pending_kind = bkind;
                stp->pending_q = bt.q; stp->pending_r = bt.r;
                stp->pending_rate = 0.0;
                stp->pending_target = btarget;
                if (!was_pending) {
                    stp->pending_position = a.x; stp->pending_position_y = a.y;
                    stp->pending_root_position = ar.x; stp->pending_root_position_y = ar.y;
                    stp->pending_stamp_q = now.q; stp->pending_stamp_r = now.r;
                }
            }
            stopped_by_time = true;
            break;
        }
        if (was_pending) {
            if (lane == 0) stp->pending_kind = ECMC_EVENT_NONE;
            if (kind != ECMC_EVENT_END_OF_CHAIN) {
                // the kept handler's in-state predates the control event's time slice
                a.x = kept_x; a.y = kept_y;
                ar.x = kept_rx; ar.y = kept_ry;
                now = kept_stamp;
            }
            was_pending = false;
        }

        // ---- time slice of the active leaf and of its root unit, each with its own velocity (abstracts.py:82-101)
        {
            const double dt = time_sub(event_time, now);
            a.x = correct_position_entry(__dadd_rn(a.x, __dmul_rn(vx, dt)), L);
            ar.x = correct_position_entry(__dadd_rn(ar.x, __dmul_rn(rvx, dt)), L);
            a.y = correct_position_entry(__dadd_rn(a.y, __dmul_rn(vy, dt)), L);
            ar.y = correct_position_entry(__dadd_rn(ar.y, __dmul_rn(rvy, dt)), L);
            now = event_time;
        }
        int new_active = active;
        double nvx = vx, nvy = vy;
        switch (kind) {
        case ECMC_EVENT_FACTOR_PAIR: new_active = btarget; n_factor++; break;
        case ECMC_EVENT_BOND: new_active = btarget; n_bond++; break;
        case ECMC_EVENT_END_OF_CHAIN:
            // _get_new_velocity (:101-122)
            nvx = __dsub_rn(__dmul_rn(vx, K.eoc_cos), __dmul_rn(vy, K.eoc_sin));
            nvy = __dadd_rn(__dmul_rn(vx, K.eoc_sin), __dmul_rn(vy, K.eoc_cos));
            new_active = eoc_next;
            n_eoc++;
            break;
        default: break;
        }
        if (RECORD && lane == 0 && (int)n_events < A.records_per_chain) {
            EcmcEventRecord rec;
            rec.kind = kind; rec.target = kind == ECMC_EVENT_END_OF_CHAIN ? new_active : btarget; rec.target_cell = -1;
            rec.accepted = 1; rec.n_candidates = n_cand;
            rec.new_active = new_active; rec.new_direction = 0;
            rec.mode = 0;
            rec.time_q = event_time.q; rec.time_r = event_time.r;
            rec.active_pos[0] = a.x; rec.active_pos[1] = a.y; rec.active_pos[2] = 0.0;
            A.records[(size_t)chain * A.records_per_chain + n_events] = rec;
        }
        // velocity of the root units: -old x weight for the old active leaf, +new x weight for the new one
        {
            const int old_root = active / npr, new_root = new_active / npr;
            const double w = K.weight;
            if (new_active == active) {
                if (kind == ECMC_EVENT_END_OF_CHAIN) {
                    rvx = __dadd_rn(rvx, __dmul_rn(__dadd_rn(-vx, nvx), w));
                    rvy = __dadd_rn(rvy, __dmul_rn(__dadd_rn(-vy, nvy), w));
                }
            } else if (new_root == old_root) {
                rvx = __dadd_rn(rvx, __dadd_rn(__dmul_rn(-vx, w), __dmul_rn(nvx, w)));
                rvy = __dadd_rn(rvy, __dadd_rn(__dmul_rn(-vy, w), __dmul_rn(nvy, w)));
            } else {
                rvx = __dmul_rn(nvx, w);
                rvy = __dmul_rn(nvy, w);
            }
            vx = nvx; vy = nvy;
            if (new_active != active) {
                if (lane == 0) {
                    part[active].x = a.x; part[active].y = a.y;
                    if (new_root != old_root) { roots[old_root].x = ar.x; roots[old_root].y = ar.y; }
                }
                __syncwarp();
                a = part[new_active];
                if (new_root != old_root) ar = roots[new_root];
                active = new_active;
            }
        }
        ev++;
        n_events++;
        n_candidates += (unsigned long long)n_cand;
        if (kind == ECMC_EVENT_END_OF_CHAIN) {
            eoc = time_add(now, time_sub(now, now) + P.chain_time);
            const StreamKey next_key = {P.seed, stream, ev};
            eoc_next = draw_end_of_chain_active(P, next_key);
        }
    }

    if (stopped_by_time) {
        // the sampling / end-of-run handler time-slices the active unit (fixed_interval_sampling_event_handler.py:96-109)
        const double dt = time_sub(until, now);
        a.x = correct_position_entry(__dadd_rn(a.x, __dmul_rn(vx, dt)), L);
        ar.x = correct_position_entry(__dadd_rn(ar.x, __dmul_rn(rvx, dt)), L);
        a.y = correct_position_entry(__dadd_rn(a.y, __dmul_rn(vy, dt)), L);
        ar.y = correct_position_entry(__dadd_rn(ar.y, __dmul_rn(rvy, dt)), L);
        now = until;
    }
    if (lane == 0) {
        part[active].x = a.x; part[active].y = a.y;
        roots[active / npr].x = ar.x; roots[active / npr].y = ar.y;
        stp->active = active;
        stp->time_q = now.q; stp->time_r = now.r;
        stp->eoc_q = eoc.q; stp->eoc_r = eoc.r;
        stp->eoc_next_active = eoc_next;
        stp->event_counter = ev;
        stp->velocity[0] = vx; stp->velocity[1] = vy;
        stp->root_velocity[0] = rvx; stp->root_velocity[1] = rvy;
        if (A.stats) {
            unsigned long long *st = reinterpret_cast<unsigned long long *>(A.stats);
            if (n_events) atomicAdd(st + 0, (unsigned long long)n_events);
            if (n_eoc) atomicAdd(st + 5, (unsigned long long)n_eoc);
            if (n_candidates) atomicAdd(st + 6, n_candidates);
            if (n_bond) atomicAdd(st + 9, (unsigned long long)n_bond);
            if (n_factor) atomicAdd(st + 10, (unsigned long long)n_factor);
            if (n_targets) atomicAdd(st + 11, n_targets);
        }
    }
}

// InitialChainStartOfRunEventHandler (initial_chain_start_of_run_event_handler.py:92-131) for a program with general
// velocities: the active leaf gets `speed` along the initial direction, its root unit that velocity times the leaf's
// weight (abstracts.py:165-190). Runs after start_kernel, one thread per chain.
__global__ void disk_start_kernel(const DeviceState S, double speed, double weight, int initial_direction) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= S.n_chains) return;
    EcmcChainState *st = S.chains + S.first_chain + c;
    for (int d = 0; d < 2; d++) {
        st->velocity[d] = d == initial_direction ? speed : 0.0;
        st->root_velocity[d] = __dmul_rn(st->velocity[d], weight);
    }
}

}  // namespace ecmc
