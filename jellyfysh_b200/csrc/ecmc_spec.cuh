// ecmc_spec.cuh -- the Lennard-Jones / cell-veto event kernel (workloads C2 and C5): one warp advances one Markov
// chain, W = 32 / G events at a time.
//
// Nine events out of ten of this configuration are REJECTED CELL VETOES: the cell-veto candidate
// (CellVetoEventHandler.send_event_time, cell_veto_event_handler.py:200-238) is the earliest, its target cell is empty
// or its confirmation fails (leaf_unit_cell_veto_event_handler.py:117-149), and nothing changes but the position of the
// active particle along its line and the time. The veto candidate of event k (Walker cell, bound, time increment) is a
// function of the random stream, the event index and the direction only -- not of any position. So the kernel assumes
// that the next W events are rejected vetoes: under that assumption the time and the position of the active particle
// before each of them follow from the veto increments alone (the same additions the one-event-at-a-time loop performs,
// in the same order), and the W events can be evaluated side by side:
//
//   lane = e * G + g      event e of the batch, share g of its work
//   g = 0 draws the veto (uniform, time), g = 1 the Walker table index, g = 2 the confirmation number
//   every lane walks over its share of the candidate list (entries g, g + G, ...) for ITS event: separation along the
//   line of motion from the position before event e, potential change keyed by (event, target), Lennard-Jones
//   inversion, running minimum on (time, sequence number) -- the scheduler's order (heap.c:176-178) --, then G lanes
//   combine by log2(G) shuffle steps
//   the lane g = 0 compares with the veto, the cell boundary and the end of chain, and -- if the veto won -- looks up
//   the occupant of the target cell and confirms against the real derivative
//
// The first event that is NOT a plain rejected veto (a pair event, an accepted veto, a cell boundary, the end of the
// chain, a time limit, or a time slice that left the cell) ends the batch: the events before it are committed in one
// step, that event itself goes through the general out-state code (the same as event_kernel's), and everything after
// it was computed from a wrong assumption and is thrown away. With 9 % breaking events, W = 8 commits 6.1 events per
// batch on average; the work thrown away is the price for (a) no per-event warp argmin, gather, loop and commit
// overhead, (b) lanes that never diverge between candidate kinds, (c) W independent dependent-load chains (Walker
// entry -> occupant -> particle) in flight at once instead of one per event.
//
// The candidate list (occupants of the 27 nearby cells + surplus) lives in shared memory together with, per target, the
// coordinate along the line of motion and the squared distance from that line -- which only change when the list
// does -- and is rebuilt after every event that changes the active particle, its cell or the direction.
//
// PRUNE (ecmc_run; never with event records): a pair candidate is only inverted if it can fire before the earliest
// of (veto, boundary) of its event. The potential change it draws is at least u / beta (-log(1 - u) >= u), and along
// the next `reach` of the line of motion the energy cannot rise by more than reach x max |dU/dr| over the distances
// the pair can have there. That force bound is computed when the list is built, for a window of displacement ahead
// of the active particle (24 mean veto steps; the list is rebuilt when half of it is used up), and stored with the
// list: one multiplication and one comparison per (event, target) decide, everything within rounding distance of the
// threshold -- or beyond the window -- is computed in full. About half of the targets do not even need their random
// number: a pair whose energy cannot rise anywhere on the window (the target stays ahead of the active particle and
// outside the minimum sphere, or behind it and inside) has no event there whatever potential change it draws, so the
// events inside the window walk over the other, "live", entries only (l_live). The winner of every event is
// unchanged (tests/test_gpu_spec.py), only EcmcStats.candidates then counts the evaluated candidates.
#pragma once

#include "ecmc_kernels.cuh"

namespace ecmc {

// Upper bound of |dU/dx| of the Lennard-Jones pair energy on a straight line at squared distance perp2 from the
// target: |dU/dx| = |sd| / r |U'(r)| <= sup over r >= sqrt(perp2) of |U'(r)|. |U'| falls with r on the repulsive side
// and peaks at r = (26/7)^(1/6) sigma on the attractive side.
ECMC_D double lj_force_bound(const LennardJones &p, double perp2) {
    if (!(perp2 > 0.0)) return INFINITY;
    const double inv = 1.0 / perp2;
    const double x = p.sigma2 * inv;
    const double x3 = x * x * x;
    const double here = p.k * x3 * fabs(fma(12.0, x3, -6.0)) * sqrt(inv);
    const double bound = perp2 < p.r_inflection_sq ? fmax(here, p.force_max) : here;
    return bound * (1.0 + 1.0e-9);
}

// PRUNE: the window of displacement the force bounds of a candidate list cover, in mean cell-veto steps (the list is
// rebuilt when half of it is used up). Measured on C2 (profiles/README.md).
#ifndef ECMC_SPEC_WINDOW_STEPS
#define ECMC_SPEC_WINDOW_STEPS 24.0
#endif
constexpr double kWindowSteps = ECMC_SPEC_WINDOW_STEPS;
// ... and the fraction of the window after which the list is rebuilt
#ifndef ECMC_SPEC_WINDOW_USE
#define ECMC_SPEC_WINDOW_USE 0.5
#endif
constexpr double kWindowUse = ECMC_SPEC_WINDOW_USE;

// key of a candidate time in the scheduler's order: see time_key (a rounding-negative x sorts first)
ECMC_D double time_order(double x) { return x > 0.0 ? x : 0.0; }

// HOST: one launch is a whole step from host buffers (ecmc_submit_from_host_sparse): before the events the warp reads
// its chain's start configuration from RunArgs.host_in (the caller's pinned buffer, read over the link by the loads of
// the kernel itself -- no staging copy, no separate pack / start kernels waiting for a free SM), bins it into the cells
// and starts the run (start_chain); every position the events change is written through to RunArgs.host_out -- the
// position of a particle when it hands the activity over, and the last active particle at the end.
//
// MODEL: kSpecLennardJones -- chargeless Lennard-Jones pair factors, Lennard-Jones cell veto (C2, C5) --, or kSpecCoulomb --
// the Coulomb atoms of C3 (coulomb_atoms/cell_veto.ini): pair candidates from the inverse-power Coulomb bound
// (TwoLeafUnitBoundingPotentialEventHandler, two_leaf_unit_bounding_potential_event_handler.py:112-168) confirmed against
// the merged-image Coulomb derivative, the same derivative for the cell veto (leaf_unit_cell_veto_event_handler.py:
// 117-149), with charges. The batch is the same; what differs is the inversion of a pair candidate, the force bound and
// the rule for the live entries (a repulsive pair cannot fire while the target stays behind, an attractive one while it
// stays ahead), the Walker table by the sign of the active charge, and the confirmation of a veto, which is a
// warp-wide Ewald sum: the leaders only collect the separations, the warp then evaluates them one after the other, in
// event order, up to the first event that ends the batch.
constexpr int kSpecLennardJones = 0, kSpecCoulomb = 1;

template <bool RECORD, bool PRUNE, int G, int WARPS, bool HOST = false, int MODEL = kSpecLennardJones>
__global__ void __launch_bounds__(WARPS * 32, ECMC_RESIDENT_WARPS / WARPS)
lj_spec_kernel(const __grid_constant__ DeviceProgram P, const DeviceState S, const RunArgs A) {
    static_assert(G == 4 || G == 8, "lanes per event");
    static_assert(!(HOST && MODEL != kSpecLennardJones), "host steps in one launch: chargeless programs only");
    constexpr bool COULOMB = MODEL == kSpecCoulomb;
    constexpr int IPCB = ECMC_POT_INVERSE_POWER_COULOMB_BOUNDING, MIC = ECMC_POT_MERGED_IMAGE_COULOMB;
    constexpr int W = 32 / G;
    constexpr int kDoubles = (PRUNE ? 3 : 2) + (COULOMB ? 1 : 0);
    extern __shared__ double spec_shared[];
    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    const int chain = S.first_chain + blockIdx.x * WARPS + warp;
    if (chain >= S.first_chain + S.n_chains) return;
    // the candidate list of this chain: coordinate along the line of motion, squared distance from it, (force bound),
    // target particle, sequence number (= slot in the scan order of the reference's taggers)
    const int cap = A.list_capacity;
    // PRUNE: one more int per entry, the indices of the entries that can fire at all inside the window (l_live)
    constexpr int kEntryDoubles = PRUNE ? kDoubles + 2 : kDoubles + 1;
    double *l_p0 = spec_shared + (size_t)warp * cap * kEntryDoubles;
    double *l_perp2 = l_p0 + cap;
    double *l_bound = l_p0 + 2 * cap;  // PRUNE only
    int *l_target = reinterpret_cast<int *>(l_p0 + kDoubles * cap);
    int *l_seq = l_target + cap;
    int *l_live = l_seq + cap;  // PRUNE only
    double *l_kc = l_p0 + (PRUNE ? 3 : 2) * cap;  // COULOMB only: prefactor x charge product of the pair
    // COULOMB: per-warp scratch of the Ewald sum (mic_derivative_warp), behind the lists of all warps
    double *trig = spec_shared + (size_t)WARPS * cap * kEntryDoubles + (size_t)warp * kTrigDoubles;
    int n_live = 0;
    int count = -1;  // entries of the valid list; -1: rebuild
    double x_build = 0.0, window = 0.0;  // PRUNE: where the list was built, and how far its force bounds reach

    const LennardJones &lj = P.cand_potential.lj;
    const bool pair_use_charge = COULOMB && P.pair_use_charge != 0, veto_use_charge = COULOMB && P.veto_use_charge != 0;
    Particle *part = S.particles + (size_t)chain * P.n_particles;
    int *occ = S.occupants + (size_t)chain * P.n_cells;
    int *sur = S.surplus + (size_t)chain * P.max_surplus;
    EcmcChainState *stp = S.chains + chain;
    unsigned host_writes = 0;
    if (HOST) {
        // the start configuration: 3 n doubles, read as a flat array (coalesced, eight loads in flight per lane) and
        // scattered into the 32-byte particle records; chargeless programs only (charge = 1)
        const double *in = A.host_in + (size_t)chain * P.n_particles * 3;
        double *records = reinterpret_cast<double *>(part);
        const int n_values = P.n_particles * 3;
#pragma unroll 1
        for (int base = 0; base < n_values; base += 256) {
            double v[8];
#pragma unroll
            for (int j = 0; j < 8; j++) {
                const int idx = base + j * 32 + lane;
                v[j] = idx < n_values ? __ldcv(in + idx) : 0.0;
            }
#pragma unroll
            for (int j = 0; j < 8; j++) {
                const int idx = base + j * 32 + lane;
                if (idx < n_values) {
                    const int particle = idx / 3;
                    records[4 * particle + (idx - 3 * particle)] = v[j];
                }
            }
        }
        for (int i = lane; i < P.n_particles; i += 32) records[4 * i + 3] = 1.0;
        __syncwarp();
        start_chain(P, S, nullptr, A.first_stream, A.initial_active, A.initial_direction, A.stats, chain, lane,
                    A.keep_state != 0);
        __syncwarp();
    }

    // chain state -> registers (uniform over the warp)
    int active = stp->active, dir = stp->direction;
    Time now = {stp->time_q, stp->time_r};
    Time eoc = {stp->eoc_q, stp->eoc_r};
    int eoc_next = stp->eoc_next_active;
    int active_cell = stp->active_cell;
    unsigned long long ev = stp->event_counter;
    // (through REDUX: the key of the random stream is then provably warp-uniform and its round keys live in uniform registers)
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

    const Time until = {A.until_q, A.until_r};
    const double L = P.length, half = P.half_length, speed = P.speed;
    const unsigned max_events = A.max_events > 0 ? (unsigned)min(A.max_events, 0x7fffffffLL) : 0x7fffffffu;
    const int e = lane / G, g = lane % G, leader = lane - g;

    Counters n = {0, 0, 0ull, 0ull};
    bool stopped_by_time = false;

    // COULOMB: the kernel is several thousand instructions (the Ewald sum, two inversions), and warps that are all
    // somewhere else in it starve on instruction fetch (measured: 19 of 26 % of the stall samples on the Ewald sum alone):
    // the warps of a CTA meet at a barrier before every batch, like the chains of molecule_kernel before every event.
#ifdef ECMC_SPEC_ALIGN_LJ
    constexpr bool ALIGNED = true;   // (experiment: the barrier for the Lennard-Jones model as well)
#else
    constexpr bool ALIGNED = COULOMB;
#endif
    bool done = false;
    while (ALIGNED || n.events < max_events) {
        if constexpr (ALIGNED) {
            const bool finished = done || n.events >= max_events;
            if (__syncthreads_and(finished)) break;
            if (finished) continue;
        }
        // the interaction winner of the event that goes through the general out-state code
        Time bt = time_inf();
        int bkind = ECMC_EVENT_NONE, btarget = -1, bcell = -1;
        double brate = 0.0;
        int n_cand = 0;
        double u_confirmation = 0.0;
        double kept_position = 0.0;
        Time kept_stamp = now;
        if (was_pending) {
            // a candidate that survived a host control event: nothing is recomputed, no draws are consumed
            bkind = stp->pending_kind;
            bt.q = stp->pending_q; bt.r = stp->pending_r;
            brate = stp->pending_rate;
            if (bkind == ECMC_EVENT_PAIR) btarget = stp->pending_target; else bcell = stp->pending_target;
            kept_position = stp->pending_position;
            kept_stamp.q = stp->pending_stamp_q; kept_stamp.r = stp->pending_stamp_r;
            u_confirmation = stream_double({P.seed, stream, ev}, ECMC_SLOT(ECMC_SLOT_CONFIRM, 0), 0);
        } else {
            if (PRUNE && count >= 0) {
                // the force bounds of the list hold for a window of displacement from where it was built
                double travelled = a.p0 - x_build;
                if (travelled < 0.0) travelled += L;
                if (!(travelled <= kWindowUse * window)) count = -1;
            }
            if (count < 0) {
                // ---- rebuild the candidate list: occupants of the nearby cells (ExcludedCellsTagger,
                // excluded_cells_tagger.py:129-132), then the surplus (SurplusCellsTagger, :129-131)
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
                __syncwarp();
                if (PRUNE) {
                    x_build = a.p0;
                    if constexpr (COULOMB) {
                        // mean cell-veto step of this active charge: 1 / (beta total rate |charge factor|)
                        double factor = veto_use_charge ? a.charge * 1.0 : 1.0;
                        if (veto_use_charge && P.veto_target_charge != 1.0) factor = factor / P.veto_target_charge;
                        const DeviceWalker *wb = factor > 0.0 ? &P.upper[dir] : &P.lower[dir];
                        const double step = speed * P.inv_beta / (wb->total_rate * fabs(factor) * speed);
                        window = fmin(kWindowSteps * (step < INFINITY ? step : L), 0.25 * L);
                    } else {
                    window = fmin(kWindowSteps * speed * P.inv_beta * P.upper[dir].inv_total_rate_speed, 0.25 * L);
                    }
                }
                n_live = 0;
#pragma unroll 1
                for (int base = 0; base < found_so_far; base += 32) {
                    const int i = base + lane;
                    bool live = false;
                    if (i < found_so_far) {
                        const Moving tp = rotate_in(part[l_target[i]], dir);
                        const double s1 = correct_separation_in_box(tp.p1 - a.p1, L, half);
                        const double s2 = correct_separation_in_box(tp.p2 - a.p2, L, half);
                        const double perp2 = fma(s1, s1, s2 * s2);
                        l_p0[i] = tp.p0;
                        l_perp2[i] = perp2;
                        double kc = 0.0;
                        if constexpr (COULOMB) {
                            // prefactor c1 c2 as InversePowerCoulombBoundingPotential gets it (displacement_time<IPCB>)
                            kc = P.cand_potential.p0 * (pair_use_charge ? a.charge : 1.0) * (pair_use_charge ? tp.charge : 1.0);
                            l_kc[i] = kc;
                        }
                        if (PRUNE) {
                            // smallest distance of the pair while the active particle covers the window
                            const double ahead = correct_separation_in_box(tp.p0 - a.p0, L, half);
                            const double behind = ahead - window;
                            double nearest = (ahead >= 0.0 && behind <= 0.0) ? 0.0 : fmin(fabs(ahead), fabs(behind));
                            if (behind < -half) nearest = fmin(nearest, half - window);  // the separation wraps around
                            if constexpr (COULOMB) {
                                // |d/dx (kc / r)| = |kc| |sx| / r^3 <= |kc| / r^2 at the smallest distance on the window
                                l_bound[i] = fabs(kc) / fma(nearest, nearest, perp2) * (1.0 + 1.0e-9);
                                // the bound kc / r falls while a repulsive pair recedes (target behind) and while an
                                // attractive one approaches (target ahead); a neutral pair never fires
                                const bool falls = kc > 0.0 ? (ahead < -1.0e-9 * L && behind >= -half)
                                                            : (kc < 0.0 ? behind > 1.0e-9 * L : true);
                                live = !falls;
                            } else {
                            l_bound[i] = lj_force_bound(lj, fma(nearest, nearest, perp2));
                            // A pair whose energy cannot rise anywhere on the window -- the target stays ahead and outside
                            // the minimum sphere (attractive while approaching), or stays behind and inside it (repulsive
                            // while receding) -- has no event there whatever potential change it draws: it needs no
                            // random number at all. (Margins keep the rule on the safe side of rounding.)
                            const double end2 = fma(behind, behind, perp2);  // squared distance at the end of the window
                            const bool falls = (behind > 1.0e-9 * L && end2 > lj.r0sq * (1.0 + 1.0e-9)) ||
                                               (ahead < -1.0e-9 * L && behind >= -half && end2 < lj.r0sq * (1.0 - 1.0e-9));
                            live = !falls;
                            }
                        }
                    }
                    if (PRUNE) {
                        const unsigned alive = __ballot_sync(kFull, live);
                        if (live) l_live[n_live + __popc(alive & ((1u << lane) - 1u))] = i;
                        n_live += __popc(alive);
                    }
                }
                __syncwarp();
                count = found_so_far;
            }

            // ---- W events side by side ---------------------------------------------------------------------
            const int w_eff = (int)min((unsigned)W, max_events - n.events);
            const StreamKey key = {P.seed, stream, ev + (unsigned long long)e};
            const uint32_t special_slot = g == 0 ? ECMC_SLOT(ECMC_SLOT_VETO_TIME, 0)
                                                 : (g == 1 ? ECMC_SLOT(ECMC_SLOT_VETO_CHOICE, 0) : ECMC_SLOT(ECMC_SLOT_CONFIRM, 0));
            const Philox4 b = stream_block(key, special_slot, 0);
            const double u_first = words_to_double(b.w[0], b.w[1]), u_second = words_to_double(b.w[2], b.w[3]);
            const DeviceWalker *w = &P.upper[dir];  // chargeless handlers: the charge factor is 1 > 0
            // InnerPointEstimator.charge_correction_factor (inner_point_estimator.py:165-192), as event_kernel forms it
            double charge_factor = 1.0;
            if constexpr (COULOMB) {
                if (veto_use_charge) {
                    charge_factor = a.charge * 1.0;
                    if (P.veto_target_charge != 1.0) charge_factor = charge_factor / P.veto_target_charge;
                }
                if (!(charge_factor > 0.0)) { charge_factor *= -1.0; w = &P.lower[dir]; }
            }
            // random.choice(table) = table[_randbelow(n)]: rejection on the top bits of successive words (lane g = 1)
            uint32_t choice = 0;
            {
                bool found = g != 1;  // only the lanes g = 1 hold the table-index words
                Philox4 words = b;
                for (uint32_t block = 1;; block++) {
#pragma unroll
                    for (int j = 0; j < 4; j++) {
                        const uint32_t r = words.w[j] >> (32 - w->bits);
                        if (!found && r < (uint32_t)w->n_entries) { choice = r; found = true; }
                    }
                    if (__all_sync(kFull, found)) break;  // all four words rejected: 0.1 % of the draws for 1701 of 2048
                    words = stream_block(key, ECMC_SLOT(ECMC_SLOT_VETO_CHOICE, 0), block);
                }
            }
            choice = __shfl_sync(kFull, choice, leader + 1);
            const double u_conf = __shfl_sync(kFull, u_first, leader + 2);
            // the veto candidate of event e (meaningful on the leaders): Walker.sample_cell (walker.py:114-118),
            // expovariate(beta) / (total rate * speed)
            const WalkerEntry entry = w->entries[choice];
            const bool first_cell = 0.0 + (w->mean_rate - 0.0) * u_first <= entry.rate_a;
            const int relative = first_cell ? entry.cell_a : entry.cell_b;
            double veto_rate = first_cell ? entry.bound_a : entry.bound_b;
            double veto_dt = -log_unit_interval(1.0 - u_second) * P.inv_beta * w->inv_total_rate_speed;
            if constexpr (COULOMB) {
                veto_rate = veto_rate * charge_factor;
                // (a neutral active unit has no cell-veto events)
                if (veto_use_charge)
                    veto_dt = charge_factor > 0.0
                                  ? -log_unit_interval(1.0 - u_second) * P.inv_beta / (w->total_rate * charge_factor * speed)
                                  : INFINITY;
            }

            // time and position before every event of the batch if all earlier ones are rejected vetoes: the additions
            // of Time.__add__ (time.py:115-133) and of the time slice (abstracts.py:82-95), one event after the other
            // lane (e, .) applies the increments of the events 0 .. e - 1 and stops
            Time my_now = now;
            double my_x = a.p0;
#pragma unroll 1
            for (int k = 0; k < W - 1; k++) {
                const double dtv = __shfl_sync(kFull, veto_dt, k * G);
                const double xr = my_now.r + dtv;
                const double fl = floor(xr);
                Time t_next;
                t_next.q = my_now.q + fl; t_next.r = xr - fl;
                const double x_raw = __dadd_rn(my_x, __dmul_rn(speed, time_sub(t_next, my_now)));
                if (k < e) {
                    my_now = t_next;
                    my_x = x_raw >= L ? x_raw - L : x_raw;  // correct_position_entry for [0, 2 L); anything else breaks below
                }
            }
            // ... and its own event as a rejected veto
            const double my_veto_dt = __shfl_sync(kFull, veto_dt, leader);
            const double xv = my_now.r + my_veto_dt;
            Time my_veto_time;
            {
                const double fl = floor(xv);
                my_veto_time.q = my_now.q + fl; my_veto_time.r = xv - fl;
            }
            const double my_raw_x = __dadd_rn(my_x, __dmul_rn(speed, time_sub(my_veto_time, my_now)));
            const double my_next_x = my_raw_x >= L ? my_raw_x - L : my_raw_x;
            // a time slice that leaves the cell (without a boundary event) or the range of the short modulo ends the batch
            const bool my_left = (boundary == 0.0 ? my_next_x < my_x : my_next_x >= boundary) || !(my_raw_x >= 0.0 && my_raw_x < 2.0 * L);
            // special candidates of event e
            double boundary_separation = boundary - my_x;
            if (boundary_separation < 0.0) boundary_separation = boundary_separation + L;  // next_image, hypercubic_setting.py:191
            const double xb = my_now.r + boundary_separation * P.inv_speed;
            // PRUNE: no pair candidate beyond this displacement can be the interaction winner of event e
            const double reach = PRUNE ? fma(speed * (fmin(xv, xb) - my_now.r), 1.0 + 1.0e-9, 1.0e-12) : 0.0;
            bool beyond_window = false;  // PRUNE: the stored force bounds do not cover this event
            if (PRUNE) {
                double travelled = my_x - x_build;
                if (travelled < 0.0) travelled += L;
                beyond_window = !(travelled + reach <= window);
            }

            // ---- pair candidates of event e: entries g, g + G, ... of the list
            double best_x = INFINITY;
            int best_seq = kSeqNone, best_target = -1, n_finite = 0;
            // TwoLeafUnitEventHandler.send_event_time (two_leaf_unit_event_handler.py:127-138) for list entry i with the
            // uniform u of its potential change
            auto candidate = [&](int i, double u) {
                const double s0 = correct_separation_in_box(l_p0[i] - my_x, L, half);
                const double du = -log_unit_interval(1.0 - u) * P.inv_beta;
                double displacement;
                if constexpr (COULOMB) displacement = ipcb_displacement(l_kc[i], s0, l_perp2[i], du, L);
                else displacement = lj_displacement(lj, s0, l_perp2[i], du);
                const double x = my_now.r + displacement * P.inv_speed;
                if (x < INFINITY) {  // heap_scheduler.py:139; NaN never wins
                    n_finite++;
                    const int seq = l_seq[i];
                    const double kx = time_order(x), kb = time_order(best_x);
                    if (kx < kb || (kx == kb && seq < best_seq)) { best_x = x; best_seq = seq; best_target = l_target[i]; }
                }
            };
            if (!PRUNE) {
#pragma unroll 1
                for (int base = 0; base < count; base += G) {
                    const int i = min(base + g, count - 1);  // lanes beyond the end repeat the last entry
                    const Philox4 pb = stream_block(key, ECMC_SLOT(ECMC_SLOT_PAIR_TIME, l_target[i]), 0);
                    const int finite_before = n_finite;
                    candidate(i, words_to_double(pb.w[0], pb.w[1]));
                    if (base + g >= count) n_finite = finite_before;  // (the repeated entry changes nothing else)
                }
            } else {
                // Pruned: the loop only draws the uniforms and compares them with the force bound; the rare candidate
                // that may fire before the special candidates waits in a one-entry queue per lane and is inverted
                // when a lane needs its queue again, or after the loop -- once per batch instead of once per pass.
                const double threshold = reach * (P.beta / (1.0 - 1.0e-9));  // bound * reach < u / beta (1 - 1e-9)
                int queued = -1;
                double queued_u = 0.0;
                // events inside the window walk over the live entries only; an event beyond it over the whole list
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
            // combine the G shares of an event
#pragma unroll
            for (int offset = 1; offset < G; offset <<= 1) {
                const double ox = __shfl_xor_sync(kFull, best_x, offset);
                const int oseq = __shfl_xor_sync(kFull, best_seq, offset);
                const int otarget = __shfl_xor_sync(kFull, best_target, offset);
                n_finite += __shfl_xor_sync(kFull, n_finite, offset);
                const double ko = time_order(ox), kb = time_order(best_x);
                if (ko < kb || (ko == kb && oseq < best_seq)) { best_x = ox; best_seq = oseq; best_target = otarget; }
            }

            // ---- the winner of event e, on the leader lanes
            int my_kind = best_seq != kSeqNone ? ECMC_EVENT_PAIR : ECMC_EVENT_NONE;
            int my_cell = -1, my_target = best_target, my_occupant = -1;
            double my_best = best_x;
            bool plain = false, violation = false;
            int my_candidates = n_finite;
            // COULOMB: the veto of this event won and its target cell is occupied: separation and charge of the occupant
            bool confirm = false, confirm_left = false;
            double csx = 0.0, csy = 0.0, csz = 0.0, cc2 = 1.0;
            if (g == 0 && e < w_eff) {
                // sequence numbers: pair slots in scan order, then veto, then boundary
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
                    // translate(active cell, relative cell), per axis (cuboid_periodic_cells.py:182-207; modular)
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
                    // mediator.py:265-292 + leaf_unit_cell_veto_event_handler.py:117-149, at the position after the time slice
                    bool accepted = false;
                    my_occupant = occ[my_cell];
                    if (my_occupant >= 0) {
                        const Moving tp = rotate_in(part[my_occupant], dir);
                        const double sx = correct_separation_in_box(tp.p0 - my_next_x, L, half);
                        const double sy = correct_separation_in_box(tp.p1 - a.p1, L, half);
                        const double sz = correct_separation_in_box(tp.p2 - a.p2, L, half);
                        if constexpr (COULOMB) {
                            confirm = true;
                            csx = sx; csy = sy; csz = sz;
                            cc2 = veto_use_charge ? tp.charge : 1.0;
                        } else {
                        const double real = lj_derivative(P.veto_potential.lj, sx, fma(sy, sy, sz * sz)) * speed;
                        if (real > 0.0) {
                            violation = veto_rate < real;
                            accepted = 0.0 + (veto_rate - 0.0) * u_conf < real;
                        }
                        }
                    }
                    plain = !accepted && !my_left;
                    confirm_left = my_left;
                }
            }
            if constexpr (COULOMB) {
                // The merged-image Coulomb derivative is a sum over the 32 lanes (mic_derivative_warp): the confirmations of
                // the batch are evaluated by the whole warp one after the other, in event order, up to the first event that
                // ends the batch anyway (everything after it is discarded).
                const unsigned ends = __ballot_sync(kFull, g == 0 && !(plain || (confirm && !confirm_left)));
                const int first_end = ends ? __ffs(ends) - 1 : 32;
                unsigned todo = __ballot_sync(kFull, confirm && !confirm_left) & (first_end < 32 ? (1u << first_end) - 1u : kFull);
                const double c1 = veto_use_charge ? a.charge : 1.0;
                while (todo) {
                    const int src = __ffs(todo) - 1;
                    todo &= todo - 1;
                    const double real = derivative_warp<MIC>(P.veto_potential, 0, speed, __shfl_sync(kFull, csx, src),
                                                             __shfl_sync(kFull, csy, src), __shfl_sync(kFull, csz, src), c1,
                                                             __shfl_sync(kFull, cc2, src), trig, lane);
                    const double rate = __shfl_sync(kFull, veto_rate, src), u = __shfl_sync(kFull, u_conf, src);
                    const bool accepted = real > 0.0 && 0.0 + (rate - 0.0) * u < real;
                    if (lane == src) {
                        violation = real > 0.0 && rate < real;
                        plain = !accepted;  // (this leader has not left its cell: confirm_left is false)
                    }
                    if (accepted) break;  // the batch ends here
                }
            }
            const unsigned breaking = __ballot_sync(kFull, g == 0 && !plain);
            // events 0 .. e_star - 1 are plain rejected vetoes
            const int e_star = breaking ? min(w_eff, (__ffs(breaking) - 1) / G) : w_eff;

            // ---- commit them
            if (e_star > 0) {
                const bool mine = g == 0 && e < e_star;
                if (RECORD && mine && (int)(n.events + e) < A.records_per_chain) {
                    EcmcEventRecord rec;
                    rec.kind = ECMC_EVENT_CELL_VETO; rec.target = my_occupant; rec.target_cell = my_cell;
                    rec.accepted = 0; rec.n_candidates = my_candidates + 1;
                    rec.new_active = active; rec.new_direction = dir; rec.mode = 0;
                    rec.time_q = my_veto_time.q; rec.time_r = my_veto_time.r;
                    Moving after = a;
                    after.p0 = my_next_x;
                    const Particle lab = rotate_out(after, dir);
                    rec.active_pos[0] = lab.x; rec.active_pos[1] = lab.y; rec.active_pos[2] = lab.z;
                    A.records[(size_t)chain * A.records_per_chain + n.events + e] = rec;
                }
                n.candidates += (unsigned long long)(__reduce_add_sync(kFull, mine ? my_candidates + 1 : 0));
                const unsigned violations = __ballot_sync(kFull, mine && violation);
                if (violations && lane == 0 && A.stats)
                    atomicAdd(reinterpret_cast<unsigned long long *>(A.stats) + 7, (unsigned long long)__popc(violations));
                n.targets += (unsigned long long)e_star * (unsigned long long)count;
                n.events += (unsigned)e_star;
                n.veto += (unsigned)e_star;
                ev += (unsigned long long)e_star;
                // the state before event e_star = the state after event e_star - 1
                now.q = __shfl_sync(kFull, my_veto_time.q, (e_star - 1) * G);
                now.r = __shfl_sync(kFull, my_veto_time.r, (e_star - 1) * G);
                a.p0 = __shfl_sync(kFull, my_next_x, (e_star - 1) * G);
            }
            if (e_star >= w_eff) continue;  // the event limit, or a whole batch of rejected vetoes
            // ---- event e_star goes through the general out-state code below
            const int source = e_star * G;
            bkind = __shfl_sync(kFull, my_kind, source);
            btarget = __shfl_sync(kFull, my_target, source);
            bcell = __shfl_sync(kFull, my_cell, source);
            brate = bkind == ECMC_EVENT_CELL_VETO ? __shfl_sync(kFull, veto_rate, source) : 0.0;  // stored by the veto handler only
            n_cand = __shfl_sync(kFull, my_candidates, source);
            u_confirmation = __shfl_sync(kFull, u_conf, source);
            const double x_star = __shfl_sync(kFull, my_best, source);
            if (bkind != ECMC_EVENT_NONE) {
                const double fl = floor(x_star);
                bt.q = now.q + fl; bt.r = x_star - fl;
            }
            n.targets += (unsigned long long)count;
        }

        // ================= one event, general: the tail of event_kernel for this configuration =================
        n_cand++;  // the end-of-chain candidate lives in the scheduler since the chain started
        const bool eoc_first = time_lt(eoc, bt);
        const Time event_time = eoc_first ? eoc : bt;
        const int kind = eoc_first ? ECMC_EVENT_END_OF_CHAIN : bkind;
        if (!time_lt(event_time, until)) {
            // a host control event comes first: the interaction winner stays scheduled
            if (lane == 0) {
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
            if constexpr (ALIGNED) { done = true; continue; }
            break;
        }
        if (was_pending) {
            if (lane == 0) stp->pending_kind = ECMC_EVENT_NONE;
            if (kind != ECMC_EVENT_END_OF_CHAIN) {
                a.p0 = kept_position;  // the kept handler's in-state predates the control event's time slice
                now = kept_stamp;
            }
            was_pending = false;
        }
        // time slice of the active particle (event_handler/abstracts/abstracts.py:82-95)
        const double x_before = a.p0;
        a.p0 = correct_position_entry(__dadd_rn(x_before, __dmul_rn(speed, time_sub(event_time, now))), L);
        now = event_time;
        const bool left_cell = boundary == 0.0 ? a.p0 < x_before : a.p0 >= boundary;
        int new_active = active, accepted = 0, rec_target = -1;
        switch (kind) {
        case ECMC_EVENT_PAIR:  // two_leaf_unit_event_handler.py:140-154
            rec_target = btarget;
            if constexpr (COULOMB) {
                // two_leaf_unit_bounding_potential_event_handler.py:148-168 + event_handler_with_bounding_potential.py:75-101
                const Moving tp = rotate_in(part[btarget], dir);
                const double sx = correct_separation_in_box(tp.p0 - a.p0, L, half);
                const double sy = correct_separation_in_box(tp.p1 - a.p1, L, half);
                const double sz = correct_separation_in_box(tp.p2 - a.p2, L, half);
                const double c1 = pair_use_charge ? a.charge : 1.0, c2 = pair_use_charge ? tp.charge : 1.0;
                const double bounding_rate = derivative_warp<IPCB>(P.cand_potential, 0, speed, sx, sy, sz, c1, c2, trig, lane);
                const double real = derivative_warp<MIC>(P.real_potential, 0, speed, sx, sy, sz, c1, c2, trig, lane);
                if (real > 0.0) {
                    if (bounding_rate < real) count_rare(A, lane, 7);
                    if (0.0 + (bounding_rate - 0.0) * u_confirmation < real) accepted = 1;
                }
                if (accepted) new_active = btarget;
            } else {
            accepted = 1;
            new_active = btarget;
            }
            count_rare(A, lane, 1);
            break;
        case ECMC_EVENT_CELL_VETO: {
            const int t = occ[bcell];
            rec_target = t;
            if (t >= 0) {
                const Moving tp = rotate_in(part[t], dir);
                const double sx = correct_separation_in_box(tp.p0 - a.p0, L, half);
                const double sy = correct_separation_in_box(tp.p1 - a.p1, L, half);
                const double sz = correct_separation_in_box(tp.p2 - a.p2, L, half);
                double real;
                if constexpr (COULOMB)
                    real = derivative_warp<MIC>(P.veto_potential, 0, speed, sx, sy, sz, veto_use_charge ? a.charge : 1.0,
                                                veto_use_charge ? tp.charge : 1.0, trig, lane);
                else real = lj_derivative(P.veto_potential.lj, sx, fma(sy, sy, sz * sz)) * speed;
                if (real > 0.0) {
                    if (brate < real) count_rare(A, lane, 7);
                    if (0.0 + (brate - 0.0) * u_confirmation < real) { accepted = 1; new_active = t; }
                }
            }
            n.veto++;
            if (accepted) count_rare(A, lane, 3);
            break;
        }
        case ECMC_EVENT_CELL_BOUNDARY:  // lands exactly on the lower boundary of the new cell (cell_boundary_event_handler.py:158-173)
            count_rare(A, lane, 4);
            a.p0 = boundary;
            break;
        case ECMC_EVENT_END_OF_CHAIN:  // abstracts/end_of_chain_event_handler.py:107-187, new direction = (d + 1) mod D
            new_active = eoc_next;
            rec_target = new_active;
            accepted = 1;
            count_rare(A, lane, 5);
            break;
        default: break;
        }
        if (RECORD && lane == 0 && (int)n.events < A.records_per_chain) {
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
            count = -1;  // the candidate list belongs to the old active particle / cell / direction
            Particle lab = rotate_out(a, dir);
            if (kind == ECMC_EVENT_END_OF_CHAIN) dir = dir == 2 ? 0 : dir + 1;
            // SingleActiveCellOccupancy.update (single_active_cell_occupancy.py:149-203)
            const bool handed_over = new_active != active;
            if (handed_over) {
                int delta = 0;
                if (lane == 0) {
                    store_position(part + active, lab);
                    if (HOST) {
                        double *out = A.host_out + ((size_t)chain * P.n_particles + active) * 3;
                        out[0] = lab.x; out[1] = lab.y; out[2] = lab.z;
                        host_writes++;
                    }
                    delta = occupancy_insert(occ, sur, n_surplus, 1, P.max_surplus, active_cell, active);
                }
                delta = __shfl_sync(kFull, delta, 0);
                if (delta == 2) count_rare(A, lane, 8); else n_surplus += delta;
                __syncwarp();
                active = new_active;
                lab = part[active];
            }
            // the oracle recomputes the cell from the position after every event; only these events can change it
            a = rotate_in(lab, dir);
            cell_identifier_of(P, lab, cid0, cid1, cid2);
            active_cell = cid0 * P.cumulative[0] + cid1 * P.cumulative[1] + cid2 * P.cumulative[2];
            if (handed_over) {
                int delta = 0;
                if (lane == 0) delta = occupancy_remove(occ, sur, n_surplus, 1, active_cell, active);
                delta = __shfl_sync(kFull, delta, 0);
                if (delta == 2) count_rare(A, lane, 8); else n_surplus += delta;
                __syncwarp();
            }
            boundary = next_boundary();
        }
        if (kind == ECMC_EVENT_END_OF_CHAIN) {
            // the next end-of-chain candidate: chain_time after this one, new active by randint
            // (single_independent_active_periodic_direction_end_of_chain_event_handler.py:203-237)
            eoc = time_add(now, time_sub(now, now) + P.chain_time);
            eoc_next = (int)stream_randbelow({P.seed, stream, ev}, ECMC_SLOT(ECMC_SLOT_END_OF_CHAIN, 0), (uint32_t)P.n_particles);
        }
    }

    if (stopped_by_time) {
        // the sampling / end-of-run handler time-slices the active unit (fixed_interval_sampling_event_handler.py:96-109)
        a.p0 = correct_position_entry(__dadd_rn(a.p0, __dmul_rn(speed, time_sub(until, now))), L);
        now = until;
    }
    if (lane == 0) {
        store_position(part + active, rotate_out(a, dir));
        if (HOST) {
            const Particle lab = rotate_out(a, dir);
            double *out = A.host_out + ((size_t)chain * P.n_particles + active) * 3;
            out[0] = lab.x; out[1] = lab.y; out[2] = lab.z;
            if (A.host_writes) atomicAdd(A.host_writes, (unsigned long long)(host_writes + 1));
        }
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
