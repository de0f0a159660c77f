// ecmc_spec_cta.cuh -- the Lennard-Jones / cell-veto kernel for FEW chains (workload C5: one chain of 65536 particles):
// one CTA of four warps advances one Markov chain, 32 events at a time.
//
// lj_spec_kernel (ecmc_spec.cuh) gives every chain one warp and evaluates 8 events side by side under the assumption that
// they are rejected cell vetoes. With thousands of chains that fills the GPU; with one chain (or a few dozen) the rate is
// the latency of one warp's batch, and every batch commits only 5.9 of its 8 events before the first event that is not a
// plain rejected veto. Here four warps share the chain: warp w evaluates the events 8 w .. 8 w + 7 of a batch of 32 with
// exactly the per-event code of lj_spec_kernel (4 lanes per event), so a batch commits 10.6 events on average (9.7 %
// breaking events) in little more than the time of an 8-event batch. What is sequential stays sequential and exact: the
// time and position before event e are the fold of the veto increments of the events before it, performed one addition
// after the other in the order of the one-event loop (every lane runs the recurrence up to its own event from the 32
// increments in shared memory). The candidate list, the increments and the per-event results go through shared memory;
// everything chain-uniform (lifting state, counters, the general out-state of the breaking event) is computed by all
// 128 threads from identical inputs, thread 0 writes. The committed events are those of lj_spec_kernel and of the
// one-event kernel, bit for bit (tests/test_gpu_spec.py).
#pragma once

#include "ecmc_spec.cuh"

namespace ecmc {

constexpr int kChainWarps = 4;                 // warps per chain (= per CTA)
constexpr int kChainLanes = 4;                 // lanes per event (G)
constexpr int kChainBatch = kChainWarps * 32 / kChainLanes;  // events per batch

template <bool RECORD, bool PRUNE>
__global__ void __launch_bounds__(kChainWarps * 32)
lj_chain_kernel(const __grid_constant__ DeviceProgram P, const DeviceState S, const RunArgs A) {
    constexpr int G = kChainLanes, W = 32 / G, WT = kChainBatch;
    constexpr int kDoubles = PRUNE ? 3 : 2;
    extern __shared__ double spec_shared[];
    // per-event results of a batch (written by the leader lane of every event, read by everybody after the barrier)
    __shared__ double s_dt[WT], s_time_q[WT], s_time_r[WT], s_next_x[WT], s_best[WT], s_rate[WT], s_uconf[WT];
    __shared__ int s_kind[WT], s_target[WT], s_cell[WT], s_cand[WT];
    __shared__ unsigned s_break[kChainWarps];
    __shared__ int s_count, s_delta;
    const int tid = threadIdx.x;
    const int lane = tid & 31, warp = tid >> 5;
    const int chain = S.first_chain + blockIdx.x;
    if (chain >= S.first_chain + S.n_chains) return;
    const bool writer = tid == 0;
    const int cap = A.list_capacity;
    double *l_p0 = spec_shared;
    double *l_perp2 = l_p0 + cap;
    double *l_bound = l_p0 + 2 * cap;  // PRUNE only
    int *l_target = reinterpret_cast<int *>(l_p0 + kDoubles * cap);
    int *l_seq = l_target + cap;
    int *l_live = l_seq + cap;  // PRUNE only: the entries that can fire at all inside the window (ecmc_spec.cuh)
    __shared__ int s_live;
    int n_live = 0;
    int count = -1;  // entries of the valid list; -1: rebuild
    double x_build = 0.0, window = 0.0;

    const LennardJones &lj = P.cand_potential.lj;
    Particle *part = S.particles + (size_t)chain * P.n_particles;
    int *occ = S.occupants + (size_t)chain * P.n_cells;
    int *sur = S.surplus + (size_t)chain * P.max_surplus;
    EcmcChainState *stp = S.chains + chain;

    // chain state -> registers of every thread (uniform over the CTA)
    int active = stp->active, dir = stp->direction;
    Time now = {stp->time_q, stp->time_r};
    Time eoc = {stp->eoc_q, stp->eoc_r};
    int eoc_next = stp->eoc_next_active;
    int active_cell = stp->active_cell;
    unsigned long long ev = stp->event_counter;
    const uint32_t stream = __reduce_or_sync(kFull, stp->stream);
    bool was_pending = stp->pending_kind != ECMC_EVENT_NONE;
    int n_surplus = S.n_surplus[chain];
    Moving a = rotate_in(part[active], dir);
    int cid0 = (active_cell / P.cumulative[0]) % P.per_side[0];
    int cid1 = (active_cell / P.cumulative[1]) % P.per_side[1];
    int cid2 = (active_cell / P.cumulative[2]) % P.per_side[2];
    int next_cell = 0;
    auto next_boundary = [&]() {
        const int id = dir == 0 ? cid0 : (dir == 1 ? cid1 : cid2);
        const int nid = id + 1 == P.per_side[dir] ? 0 : id + 1;
        next_cell = active_cell + (nid - id) * P.cumulative[dir];
        return __ldg(P.cell_min_axis + dir * P.max_per_side + nid);
    };
    double boundary = next_boundary();
    // the kept candidate of a host control event is read before anybody may clear it
    const int pending_kind0 = stp->pending_kind, pending_target0 = stp->pending_target;
    const double pending_q0 = stp->pending_q, pending_r0 = stp->pending_r, pending_rate0 = stp->pending_rate,
                 pending_position0 = stp->pending_position, pending_stamp_q0 = stp->pending_stamp_q,
                 pending_stamp_r0 = stp->pending_stamp_r;
    __syncthreads();

    const Time until = {A.until_q, A.until_r};
    const double L = P.length, half = P.half_length, speed = P.speed;
    const unsigned max_events = A.max_events > 0 ? (unsigned)min(A.max_events, 0x7fffffffLL) : 0x7fffffffu;
    const int e_local = lane / G, g = lane % G, leader = lane - g;
    const int E = warp * W + e_local;  // this lane's event of the batch

    Counters n = {0, 0, 0ull, 0ull};
    bool stopped_by_time = false;

    while (n.events < max_events) {
        Time bt = time_inf();
        int bkind = ECMC_EVENT_NONE, btarget = -1, bcell = -1;
        double brate = 0.0;
        int n_cand = 0;
        double u_confirmation = 0.0;
        double kept_position = 0.0;
        Time kept_stamp = now;
        if (was_pending) {
            // a candidate that survived a host control event: nothing is recomputed, no draws are consumed
            bkind = pending_kind0;
            bt.q = pending_q0; bt.r = pending_r0;
            brate = pending_rate0;
            if (bkind == ECMC_EVENT_PAIR) btarget = pending_target0; else bcell = pending_target0;
            kept_position = pending_position0;
            kept_stamp.q = pending_stamp_q0; kept_stamp.r = pending_stamp_r0;
            u_confirmation = stream_double({P.seed, stream, ev}, ECMC_SLOT(ECMC_SLOT_CONFIRM, 0), 0);
        } else {
            if (PRUNE && count >= 0) {
                double travelled = a.p0 - x_build;
                if (travelled < 0.0) travelled += L;
                if (!(travelled <= kWindowUse * window)) count = -1;
            }
            if (count < 0) {
                // ---- warp 0 rebuilds the candidate list (the code of lj_spec_kernel), the others wait
                if (warp == 0) {
                    const int n_slots = P.n_nearby + n_surplus;
                    int found_so_far = 0;
#pragma unroll 1
                    for (int cursor = 0; cursor < n_slots; cursor += 32) {
                        const int s = cursor + lane;
                        int found = -1;
                        if (s < P.n_nearby) {
                            const int code = __ldg(P.nearby + s);
                            int x = cid0 + (code & 1023), y = cid1 + ((code >> 10) & 1023), z = cid2 + (code >> 20);
                            if (x >= P.per_side[0]) x -= P.per_side[0];
                            if (y >= P.per_side[1]) y -= P.per_side[1];
                            if (z >= P.per_side[2]) z -= P.per_side[2];
                            found = occ[x * P.cumulative[0] + y * P.cumulative[1] + z * P.cumulative[2]];
                        } else if (s < n_slots) {
                            found = sur[s - P.n_nearby];
                        }
                        const unsigned occupied = __ballot_sync(kFull, found >= 0);
                        if (found >= 0) {
                            const int rank = found_so_far + __popc(occupied & ((1u << lane) - 1u));
                            l_target[rank] = found;
                            l_seq[rank] = s;
                        }
                        found_so_far += __popc(occupied);
                    }
                    if (lane == 0) { s_count = found_so_far; s_live = 0; }
                }
                __syncthreads();
                count = s_count;
                if (PRUNE) {
                    x_build = a.p0;
                    window = fmin(kWindowSteps * speed * P.inv_beta * P.upper[dir].inv_total_rate_speed, 0.25 * L);
                }
                // positions of the targets: all four warps share the entries
#pragma unroll 1
                for (int i = tid; i < count; i += kChainWarps * 32) {
                    const Moving tp = rotate_in(part[l_target[i]], dir);
                    const double s1 = correct_separation_in_box(tp.p1 - a.p1, L, half);
                    const double s2 = correct_separation_in_box(tp.p2 - a.p2, L, half);
                    const double perp2 = fma(s1, s1, s2 * s2);
                    l_p0[i] = tp.p0;
                    l_perp2[i] = perp2;
                    if (PRUNE) {
                        const double ahead = correct_separation_in_box(tp.p0 - a.p0, L, half);
                        const double behind = ahead - window;
                        double nearest = (ahead >= 0.0 && behind <= 0.0) ? 0.0 : fmin(fabs(ahead), fabs(behind));
                        if (behind < -half) nearest = fmin(nearest, half - window);
                        l_bound[i] = lj_force_bound(lj, fma(nearest, nearest, perp2));
                        // entries whose energy cannot rise anywhere on the window need no random number (ecmc_spec.cuh);
                        // the order of the live entries is free: the winner is a minimum over (time, sequence number)
                        const double end2 = fma(behind, behind, perp2);
                        const bool falls = (behind > 1.0e-9 * L && end2 > lj.r0sq * (1.0 + 1.0e-9)) ||
                                           (ahead < -1.0e-9 * L && behind >= -half && end2 < lj.r0sq * (1.0 - 1.0e-9));
                        if (!falls) l_live[atomicAdd(&s_live, 1)] = i;
                    }
                }
                __syncthreads();
                n_live = s_live;
            }

            // ---- 32 events side by side ----------------------------------------------------------------------
            const int w_eff = (int)min((unsigned)WT, max_events - n.events);
            const StreamKey key = {P.seed, stream, ev + (unsigned long long)E};
            const uint32_t special_slot = g == 0 ? ECMC_SLOT(ECMC_SLOT_VETO_TIME, 0)
                                                 : (g == 1 ? ECMC_SLOT(ECMC_SLOT_VETO_CHOICE, 0) : ECMC_SLOT(ECMC_SLOT_CONFIRM, 0));
            const Philox4 b = stream_block(key, special_slot, 0);
            const double u_first = words_to_double(b.w[0], b.w[1]), u_second = words_to_double(b.w[2], b.w[3]);
            const DeviceWalker *w = &P.upper[dir];
            uint32_t choice = 0;
            {
                bool found = g != 1;
                Philox4 words = b;
                for (uint32_t block = 1;; block++) {
#pragma unroll
                    for (int j = 0; j < 4; j++) {
                        const uint32_t r = words.w[j] >> (32 - w->bits);
                        if (!found && r < (uint32_t)w->n_entries) { choice = r; found = true; }
                    }
                    if (__all_sync(kFull, found)) break;
                    words = stream_block(key, ECMC_SLOT(ECMC_SLOT_VETO_CHOICE, 0), block);
                }
            }
            choice = __shfl_sync(kFull, choice, leader + 1);
            const double u_conf = __shfl_sync(kFull, u_first, leader + 2);
            const WalkerEntry entry = w->entries[choice];
            const bool first_cell = 0.0 + (w->mean_rate - 0.0) * u_first <= entry.rate_a;
            const int relative = first_cell ? entry.cell_a : entry.cell_b;
            const double veto_rate = first_cell ? entry.bound_a : entry.bound_b;
            const double veto_dt = -log_unit_interval(1.0 - u_second) * P.inv_beta * w->inv_total_rate_speed;
            if (g == 0) s_dt[E] = veto_dt;
            __syncthreads();

            // time and position before event E if all earlier events of the batch are rejected vetoes: the additions of
            // Time.__add__ (time.py:115-133) and of the time slice (abstracts.py:82-95), one event after the other
            Time my_now = now;
            double my_x = a.p0;
            {
                // (first the events of the warps before this one -- the same trip count for all lanes of the warp --, then
                // the events of this warp before E; the time chain r -> r + dt -> floor -> r' is what the loop waits for)
                const int warp_first = warp * W;
#pragma unroll 4
                for (int k = 0; k < warp_first; k++) {
                    const double xr = my_now.r + s_dt[k];
                    const double fl = floor(xr);
                    Time t_next;
                    t_next.q = my_now.q + fl; t_next.r = xr - fl;
                    const double x_raw = __dadd_rn(my_x, __dmul_rn(speed, time_sub(t_next, my_now)));
                    my_now = t_next;
                    my_x = x_raw >= L ? x_raw - L : x_raw;
                }
#pragma unroll
                for (int j = 0; j < W - 1; j++) {
                    const double xr = my_now.r + s_dt[warp_first + j];
                    const double fl = floor(xr);
                    Time t_next;
                    t_next.q = my_now.q + fl; t_next.r = xr - fl;
                    const double x_raw = __dadd_rn(my_x, __dmul_rn(speed, time_sub(t_next, my_now)));
                    if (j < e_local) {
                        my_now = t_next;
                        my_x = x_raw >= L ? x_raw - L : x_raw;
                    }
                }
            }
            const double my_veto_dt = __shfl_sync(kFull, veto_dt, leader);
            const double xv = my_now.r + my_veto_dt;
            Time my_veto_time;
            {
                const double fl = floor(xv);
                my_veto_time.q = my_now.q + fl; my_veto_time.r = xv - fl;
            }
            const double my_raw_x = __dadd_rn(my_x, __dmul_rn(speed, time_sub(my_veto_time, my_now)));
            const double my_next_x = my_raw_x >= L ? my_raw_x - L : my_raw_x;
            const bool my_left = (boundary == 0.0 ? my_next_x < my_x : my_next_x >= boundary) || !(my_raw_x >= 0.0 && my_raw_x < 2.0 * L);
            double boundary_separation = boundary - my_x;
            if (boundary_separation < 0.0) boundary_separation = boundary_separation + L;
            const double xb = my_now.r + boundary_separation * P.inv_speed;
            const double reach = PRUNE ? fma(speed * (fmin(xv, xb) - my_now.r), 1.0 + 1.0e-9, 1.0e-12) : 0.0;
            bool beyond_window = false;
            if (PRUNE) {
                double travelled = my_x - x_build;
                if (travelled < 0.0) travelled += L;
                beyond_window = !(travelled + reach <= window);
            }

            // ---- pair candidates of event E: entries g, g + G, ... of the list (the loop of lj_spec_kernel)
            double best_x = INFINITY;
            int best_seq = kSeqNone, best_target = -1, n_finite = 0;
            auto candidate = [&](int i, double u) {
                const double s0 = correct_separation_in_box(l_p0[i] - my_x, L, half);
                const double du = -log_unit_interval(1.0 - u) * P.inv_beta;
                const double x = my_now.r + lj_displacement(lj, s0, l_perp2[i], du) * P.inv_speed;
                if (x < INFINITY) {
                    n_finite++;
                    const int seq = l_seq[i];
                    const double kx = time_order(x), kb = time_order(best_x);
                    if (kx < kb || (kx == kb && seq < best_seq)) { best_x = x; best_seq = seq; best_target = l_target[i]; }
                }
            };
            if (!PRUNE) {
#pragma unroll 1
                for (int base = 0; base < count; base += G) {
                    const int i = min(base + g, count - 1);
                    const Philox4 pb = stream_block(key, ECMC_SLOT(ECMC_SLOT_PAIR_TIME, l_target[i]), 0);
                    const int finite_before = n_finite;
                    candidate(i, words_to_double(pb.w[0], pb.w[1]));
                    if (base + g >= count) n_finite = finite_before;
                }
            } else {
                const double threshold = reach * (P.beta / (1.0 - 1.0e-9));
                int queued = -1;
                double queued_u = 0.0;
                const int my_count = beyond_window ? count : n_live;
                const int walk = __any_sync(kFull, beyond_window) ? count : n_live;
#pragma unroll 1
                for (int base = 0;; base += G) {
                    const bool last = base >= walk;
                    const int k = base + g;
                    int i = 0;
                    bool maybe = false;
                    double u = 0.0;
                    if (!last) {
                        const bool valid = k < my_count;
                        i = valid ? (beyond_window ? k : l_live[k]) : 0;
                        const Philox4 pb = stream_block(key, ECMC_SLOT(ECMC_SLOT_PAIR_TIME, valid ? l_target[i] : 0), 0);
                        u = words_to_double(pb.w[0], pb.w[1]);
                        maybe = valid && (beyond_window || !(l_bound[i] * threshold < u));
                    }
                    if (__any_sync(kFull, queued >= 0 && (last || maybe))) {
                        if (queued >= 0) candidate(queued, queued_u);
                        queued = -1;
                    }
                    if (last) break;
                    if (maybe) { queued = i; queued_u = u; }
                }
            }
#pragma unroll
            for (int offset = 1; offset < G; offset <<= 1) {
                const double ox = __shfl_xor_sync(kFull, best_x, offset);
                const int oseq = __shfl_xor_sync(kFull, best_seq, offset);
                const int otarget = __shfl_xor_sync(kFull, best_target, offset);
                n_finite += __shfl_xor_sync(kFull, n_finite, offset);
                const double ko = time_order(ox), kb = time_order(best_x);
                if (ko < kb || (ko == kb && oseq < best_seq)) { best_x = ox; best_seq = oseq; best_target = otarget; }
            }

            // ---- the winner of event E, on the leader lanes
            int my_kind = best_seq != kSeqNone ? ECMC_EVENT_PAIR : ECMC_EVENT_NONE;
            int my_cell = -1, my_target = best_target, my_occupant = -1;
            double my_best = best_x;
            bool plain = false, violation = false;
            int my_candidates = n_finite;
            if (g == 0 && E < w_eff) {
                if (xv < INFINITY) {
                    my_candidates++;
                    if (time_order(xv) < time_order(my_best) || my_kind == ECMC_EVENT_NONE) { my_best = xv; my_kind = ECMC_EVENT_CELL_VETO; }
                }
                if (xb < INFINITY) {
                    my_candidates++;
                    if (time_order(xb) < time_order(my_best) || my_kind == ECMC_EVENT_NONE) { my_best = xb; my_kind = ECMC_EVENT_CELL_BOUNDARY; }
                }
                const double fl = floor(my_best);
                Time t_event;
                t_event.q = my_now.q + fl; t_event.r = my_best - fl;
                const bool wins = my_kind == ECMC_EVENT_CELL_VETO && !time_lt(eoc, t_event) && time_lt(t_event, until);
                if (my_kind == ECMC_EVENT_CELL_VETO) {
                    int tx = cid0 + (relative & 1023), ty = cid1 + ((relative >> 10) & 1023), tz = cid2 + (relative >> 20);
                    if (tx >= P.per_side[0]) tx -= P.per_side[0];
                    if (ty >= P.per_side[1]) ty -= P.per_side[1];
                    if (tz >= P.per_side[2]) tz -= P.per_side[2];
                    my_cell = tx * P.cumulative[0] + ty * P.cumulative[1] + tz * P.cumulative[2];
                    my_target = -1;
                } else if (my_kind == ECMC_EVENT_CELL_BOUNDARY) {
                    my_cell = next_cell;
                    my_target = -1;
                }
                if (wins) {
                    bool accepted = false;
                    my_occupant = occ[my_cell];
                    if (my_occupant >= 0) {
                        const Moving tp = rotate_in(part[my_occupant], dir);
                        const double sx = correct_separation_in_box(tp.p0 - my_next_x, L, half);
                        const double sy = correct_separation_in_box(tp.p1 - a.p1, L, half);
                        const double sz = correct_separation_in_box(tp.p2 - a.p2, L, half);
                        const double real = lj_derivative(P.veto_potential.lj, sx, fma(sy, sy, sz * sz)) * speed;
                        if (real > 0.0) {
                            violation = veto_rate < real;
                            accepted = 0.0 + (veto_rate - 0.0) * u_conf < real;
                        }
                    }
                    plain = !accepted && !my_left;
                }
                // publish what the commit and the out-state of a breaking event need
                s_time_q[E] = my_veto_time.q; s_time_r[E] = my_veto_time.r; s_next_x[E] = my_next_x;
                s_kind[E] = my_kind; s_target[E] = my_target; s_cell[E] = my_cell; s_cand[E] = my_candidates;
                s_best[E] = my_best; s_rate[E] = veto_rate; s_uconf[E] = u_conf;
            }
            const unsigned breaking = __ballot_sync(kFull, g == 0 && !plain);
            if (lane == 0) s_break[warp] = breaking;
            __syncthreads();
            // events 0 .. e_star - 1 are plain rejected vetoes: the earliest breaking event over the four warps
            int e_star = w_eff;
#pragma unroll
            for (int v = kChainWarps - 1; v >= 0; v--) {
                const unsigned bits = s_break[v];
                if (bits) e_star = min(w_eff, v * W + (__ffs(bits) - 1) / G);
            }

            // ---- commit them
            if (e_star > 0) {
                const bool mine = g == 0 && E < e_star;
                if (RECORD && mine && (int)(n.events + E) < A.records_per_chain) {
                    EcmcEventRecord rec;
                    rec.kind = ECMC_EVENT_CELL_VETO; rec.target = my_occupant; rec.target_cell = my_cell;
                    rec.accepted = 0; rec.n_candidates = my_candidates + 1;
                    rec.new_active = active; rec.new_direction = dir; rec.mode = 0;
                    rec.time_q = my_veto_time.q; rec.time_r = my_veto_time.r;
                    Moving after = a;
                    after.p0 = my_next_x;
                    const Particle lab = rotate_out(after, dir);
                    rec.active_pos[0] = lab.x; rec.active_pos[1] = lab.y; rec.active_pos[2] = lab.z;
                    A.records[(size_t)chain * A.records_per_chain + n.events + E] = rec;
                }
                // candidates of the committed events: lane l sums event l (every warp does the same sum)
                n.candidates += (unsigned long long)__reduce_add_sync(kFull, lane < e_star ? s_cand[lane] + 1 : 0);
                const unsigned violations = __ballot_sync(kFull, mine && violation);
                if (violations && lane == 0 && A.stats)
                    atomicAdd(reinterpret_cast<unsigned long long *>(A.stats) + 7, (unsigned long long)__popc(violations));
                n.targets += (unsigned long long)e_star * (unsigned long long)count;
                n.events += (unsigned)e_star;
                n.veto += (unsigned)e_star;
                ev += (unsigned long long)e_star;
                now.q = s_time_q[e_star - 1];
                now.r = s_time_r[e_star - 1];
                a.p0 = s_next_x[e_star - 1];
            }
            if (e_star >= w_eff) continue;  // the event limit, or a whole batch of rejected vetoes
            // ---- event e_star goes through the general out-state code below
            bkind = s_kind[e_star];
            btarget = s_target[e_star];
            bcell = s_cell[e_star];
            brate = bkind == ECMC_EVENT_CELL_VETO ? s_rate[e_star] : 0.0;
            n_cand = s_cand[e_star];
            u_confirmation = s_uconf[e_star];
            const double x_star = s_best[e_star];
            if (bkind != ECMC_EVENT_NONE) {
                const double fl = floor(x_star);
                bt.q = now.q + fl; bt.r = x_star - fl;
            }
            n.targets += (unsigned long long)count;
        }

        // ================= one event, general: the tail of lj_spec_kernel, computed by all threads =================
        n_cand++;  // the end-of-chain candidate lives in the scheduler since the chain started
        const bool eoc_first = time_lt(eoc, bt);
        const Time event_time = eoc_first ? eoc : bt;
        const int kind = eoc_first ? ECMC_EVENT_END_OF_CHAIN : bkind;
        if (!time_lt(event_time, until)) {
            if (writer) {
                stp->pending_kind = bkind;
                stp->pending_q = bt.q; stp->pending_r = bt.r;
                stp->pending_rate = brate;
                stp->pending_target = bkind == ECMC_EVENT_PAIR ? btarget : bcell;
                if (!was_pending) {
                    stp->pending_position = a.p0;
                    stp->pending_stamp_q = now.q; stp->pending_stamp_r = now.r;
                }
            }
            stopped_by_time = true;
            break;
        }
        if (was_pending) {
            if (writer) stp->pending_kind = ECMC_EVENT_NONE;
            if (kind != ECMC_EVENT_END_OF_CHAIN) {
                a.p0 = kept_position;
                now = kept_stamp;
            }
            was_pending = false;
        }
        const double x_before = a.p0;
        a.p0 = correct_position_entry(__dadd_rn(x_before, __dmul_rn(speed, time_sub(event_time, now))), L);
        now = event_time;
        const bool left_cell = boundary == 0.0 ? a.p0 < x_before : a.p0 >= boundary;
        int new_active = active, accepted = 0, rec_target = -1;
        switch (kind) {
        case ECMC_EVENT_PAIR:
            rec_target = btarget;
            accepted = 1;
            new_active = btarget;
            if (writer) count_rare(A, 0, 1);
            break;
        case ECMC_EVENT_CELL_VETO: {
            const int t = occ[bcell];
            rec_target = t;
            if (t >= 0) {
                const Moving tp = rotate_in(part[t], dir);
                const double sx = correct_separation_in_box(tp.p0 - a.p0, L, half);
                const double sy = correct_separation_in_box(tp.p1 - a.p1, L, half);
                const double sz = correct_separation_in_box(tp.p2 - a.p2, L, half);
                const double real = lj_derivative(P.veto_potential.lj, sx, fma(sy, sy, sz * sz)) * speed;
                if (real > 0.0) {
                    if (brate < real && writer) count_rare(A, 0, 7);
                    if (0.0 + (brate - 0.0) * u_confirmation < real) { accepted = 1; new_active = t; }
                }
            }
            n.veto++;
            if (accepted && writer) count_rare(A, 0, 3);
            break;
        }
        case ECMC_EVENT_CELL_BOUNDARY:
            if (writer) count_rare(A, 0, 4);
            a.p0 = boundary;
            break;
        case ECMC_EVENT_END_OF_CHAIN:
            new_active = eoc_next;
            rec_target = new_active;
            accepted = 1;
            if (writer) count_rare(A, 0, 5);
            break;
        default: break;
        }
        if (RECORD && writer && (int)n.events < A.records_per_chain) {
            EcmcEventRecord rec;
            rec.kind = kind; rec.target = rec_target; rec.target_cell = kind == ECMC_EVENT_END_OF_CHAIN ? -1 : bcell;
            rec.accepted = accepted; rec.n_candidates = n_cand;
            rec.new_active = new_active;
            rec.new_direction = kind == ECMC_EVENT_END_OF_CHAIN ? (dir + 1) % 3 : dir;
            rec.mode = 0;
            rec.time_q = event_time.q; rec.time_r = event_time.r;
            const Particle lab = rotate_out(a, dir);
            rec.active_pos[0] = lab.x; rec.active_pos[1] = lab.y; rec.active_pos[2] = lab.z;
            A.records[(size_t)chain * A.records_per_chain + n.events] = rec;
        }
        ev++;
        n.events++;
        n.candidates += (unsigned long long)n_cand;
        const bool moved_on = new_active != active || kind == ECMC_EVENT_CELL_BOUNDARY || kind == ECMC_EVENT_END_OF_CHAIN || left_cell;
        if (moved_on) {
            count = -1;
            Particle lab = rotate_out(a, dir);
            if (kind == ECMC_EVENT_END_OF_CHAIN) dir = dir == 2 ? 0 : dir + 1;
            const bool handed_over = new_active != active;
            if (handed_over) {
                __syncthreads();  // everybody has read the old occupancy (the veto target above)
                if (writer) {
                    store_position(part + active, lab);
                    s_delta = occupancy_insert(occ, sur, n_surplus, 1, P.max_surplus, active_cell, active);
                }
                __syncthreads();
                const int delta = s_delta;
                if (delta == 2) { if (writer) count_rare(A, 0, 8); } else n_surplus += delta;
                active = new_active;
                lab = part[active];
            }
            a = rotate_in(lab, dir);
            cell_identifier_of(P, lab, cid0, cid1, cid2);
            active_cell = cid0 * P.cumulative[0] + cid1 * P.cumulative[1] + cid2 * P.cumulative[2];
            if (handed_over) {
                __syncthreads();  // s_delta has been read
                if (writer) s_delta = occupancy_remove(occ, sur, n_surplus, 1, active_cell, active);
                __syncthreads();
                const int delta = s_delta;
                if (delta == 2) { if (writer) count_rare(A, 0, 8); } else n_surplus += delta;
            }
            boundary = next_boundary();
        }
        if (kind == ECMC_EVENT_END_OF_CHAIN) {
            eoc = time_add(now, time_sub(now, now) + P.chain_time);
            eoc_next = (int)stream_randbelow({P.seed, stream, ev}, ECMC_SLOT(ECMC_SLOT_END_OF_CHAIN, 0), (uint32_t)P.n_particles);
        }
    }

    if (stopped_by_time) {
        a.p0 = correct_position_entry(__dadd_rn(a.p0, __dmul_rn(speed, time_sub(until, now))), L);
        now = until;
    }
    if (writer) {
        store_position(part + active, rotate_out(a, dir));
        stp->active = active; stp->direction = dir;
        stp->time_q = now.q; stp->time_r = now.r;
        stp->eoc_q = eoc.q; stp->eoc_r = eoc.r;
        stp->eoc_next_active = eoc_next; stp->active_cell = active_cell;
        stp->event_counter = ev;
        S.n_surplus[chain] = n_surplus;
        if (A.stats) {
            unsigned long long *st = reinterpret_cast<unsigned long long *>(A.stats);
            if (n.events) atomicAdd(st + 0, (unsigned long long)n.events);
            if (n.veto) atomicAdd(st + 2, (unsigned long long)n.veto);
            if (n.candidates) atomicAdd(st + 6, n.candidates);
            if (n.targets) atomicAdd(st + 11, n.targets);
        }
    }
}

}  // namespace ecmc
