// ecmc_kernels.cuh -- the event kernel: one warp advances one Markov chain, event after event.
//
// Per event (= one iteration of the reference's mediator loop, jellyfysh/mediator/single_process_mediator.py:91-156):
//   lanes  <- candidate targets: occupants of the nearby cells of the active cell (ExcludedCellsTagger,
//             excluded_cells_tagger.py:129-132), then the surplus list (SurplusCellsTagger, :129-131);
//             more than 32 candidates are taken in passes
//   lane 0 <- the cell-veto candidate (CellVetoEventHandler.send_event_time, cell_veto_event_handler.py:200-238)
//   lane 1 <- the cell-boundary candidate (CellBoundaryEventHandler.send_event_time, :122-156)
//   argmin over (quotient, remainder, sequence) by warp shuffles  == HeapScheduler.get_succeeding_event
//   the end-of-chain candidate persists in the chain state and is compared last
//   out-state (lifting) + commit + occupancy update, all lanes redundantly, lane 0 writes.
// Chain-level state lives in registers between events; positions / occupancy stay in HBM (L2) and are
// gathered with one 32-byte load per target.
#pragma once

#include "ecmc_program.cuh"

namespace ecmc {

constexpr unsigned kFull = 0xffffffffu;
// compile-time kind of a displaced even power potential whose power is known to be 2 (harmonic bond): never appears in
// an EcmcPotential, only as a template argument chosen by the engine
constexpr int kPotHarmonic = 103;
// resident warps (= chains) per SM the event kernel is compiled for: sets the register budget (65536 / 32 / warps)
#ifndef ECMC_RESIDENT_WARPS
#define ECMC_RESIDENT_WARPS 28
#endif
constexpr int kSeqNone = 0x7fffffff;
constexpr int kListCapacity = 128;  // per-warp compact candidate list (targets + sequence numbers: 1 KB)

template <int KIND>
ECMC_D int resolve_kind(int runtime_kind) { return KIND >= 0 ? KIND : runtime_kind; }

ECMC_D bool needs_potential_change(int kind) { return !(kind == ECMC_POT_HARD_SPHERE || kind == ECMC_POT_HARD_DIPOLE); }

// component along the direction of motion and squared norm of the others
ECMC_D void split_separation(int dir, double sx, double sy, double sz, double &sd, double &perp2) {
    if (dir == 0) { sd = sx; perp2 = fma(sy, sy, sz * sz); }
    else if (dir == 1) { sd = sy; perp2 = fma(sx, sx, sz * sz); }
    else { sd = sz; perp2 = fma(sx, sx, sy * sy); }
}

// InvertiblePotential.displacement(velocity, separation, charges, potential_change): a TIME
// (jellyfysh/potential/potential.py:218-301, potential/abstracts.py:212-243)
// `inv_speed` = 1 / speed; the hard potentials take the velocity itself.
template <int KIND>
ECMC_D double displacement_time(const PotentialParams &p, int dir, double inv_speed, double length, double sx, double sy,
                                double sz, double c1, double c2, double du) {
    double sd, perp2;
    split_separation(dir, sx, sy, sz, sd, perp2);
    switch (resolve_kind<KIND>(p.kind)) {
    case ECMC_POT_LENNARD_JONES: return lj_displacement(p.lj, sd, perp2, du) * inv_speed;
    case ECMC_POT_INVERSE_POWER: {
        double r2, p2;
        ip_squares(dir, sx, sy, sz, r2, p2);
        return ip_displacement(p.ip, sd, p2, r2, c1, c2, du) * inv_speed;
    }
    case ECMC_POT_DISPLACED_EVEN_POWER: return dep_displacement<false>(p.dep, sd, perp2, du) * inv_speed;
    case kPotHarmonic: return dep_displacement<true>(p.dep, sd, perp2, du) * inv_speed;
    case ECMC_POT_HARD_SPHERE: {
        const double speed = 1.0 / inv_speed;
        return hard_sphere_time(p.p0, __dmul_rn(speed, speed), __dmul_rn(speed, sd), dot3(sx, sy, sz, sx, sy, sz));
    }
    case ECMC_POT_HARD_DIPOLE: {
        const double speed = 1.0 / inv_speed;
        return hard_dipole_time(p.p0, p.p1, __dmul_rn(speed, speed), __dmul_rn(speed, sd), dot3(sx, sy, sz, sx, sy, sz));
    }
    case ECMC_POT_INVERSE_POWER_COULOMB_BOUNDING:
        return ipcb_displacement(p.p0 * c1 * c2, sd, perp2, du, length) * inv_speed;
    default: return NAN;
    }
}

// Can this pair candidate fire at all? `du_lower` = u / beta is a lower bound of the potential change the candidate
// will draw (-log(1 - u) >= u). True only if the displacement is certainly infinite -- then the candidate never enters
// the scheduler (heap_scheduler.py:139) and its logarithm and inversion need not be computed. Separation in the frame
// of motion (s0 along the direction).
template <int KIND>
ECMC_D bool certainly_dead(const PotentialParams &p, double s0, double s1, double s2, double du_lower) {
    switch (resolve_kind<KIND>(p.kind)) {
    case ECMC_POT_LENNARD_JONES: return lj_certainly_dead(p.lj, s0, fma(s1, s1, s2 * s2), du_lower);
    default: return false;
    }
}

// Potential.derivative(velocity, separation, charges) (potential/abstracts.py:80-103). Warp-collective: all 32
// lanes call it with identical arguments (the merged-image Coulomb sum is spread over the lanes).
template <int KIND>
ECMC_D double derivative_warp(const PotentialParams &p, int dir, double speed, double sx, double sy, double sz,
                              double c1, double c2, double *trig, int lane) {
    double sd, perp2;
    split_separation(dir, sx, sy, sz, sd, perp2);
    switch (resolve_kind<KIND>(p.kind)) {
    case ECMC_POT_LENNARD_JONES: return lj_derivative(p.lj, sd, perp2) * speed;
    case ECMC_POT_INVERSE_POWER: {
        double r2, p2;
        ip_squares(dir, sx, sy, sz, r2, p2);
        return ip_derivative(p.ip, sd, r2, c1, c2) * speed;
    }
    case ECMC_POT_DISPLACED_EVEN_POWER: return dep_derivative(p.dep, sd, perp2) * speed;
    case ECMC_POT_INVERSE_POWER_COULOMB_BOUNDING: return ipcb_derivative(p.p0 * c1 * c2, sd, perp2) * speed;
    case ECMC_POT_MERGED_IMAGE_COULOMB: {
        // permutation_3d (base/vectors.py:217-239): x is the direction of motion
        const double px = sd;
        const double py = dir == 0 ? sy : (dir == 1 ? sz : sx);
        const double pz = dir == 0 ? sz : (dir == 1 ? sx : sy);
        return p.mic.prefactor * c1 * c2 * mic_derivative_warp(p.mic, px, py, pz, trig, lane) * speed;
    }
    default: return NAN;
    }
}

template <int REAL, int VETO>
struct UsesMic {
    static constexpr bool value = REAL < 0 || VETO < 0 || REAL == ECMC_POT_MERGED_IMAGE_COULOMB ||
                                  VETO == ECMC_POT_MERGED_IMAGE_COULOMB;
};

// position -> per-axis cell identifiers, int(x / cell side) (cuboid_cells.py:190-211)
ECMC_D void cell_identifier_of(const DeviceProgram &P, const Particle &a, int id[3]) {
    id[0] = (int)(a.x / P.side_length[0]);
    id[1] = P.dimension > 1 ? (int)(a.y / P.side_length[1]) : 0;
    id[2] = P.dimension > 2 ? (int)(a.z / P.side_length[2]) : 0;
}
ECMC_D void cell_identifier_of(const DeviceProgram &P, const Particle &a, int &id0, int &id1, int &id2) {
    id0 = (int)(a.x / P.side_length[0]);
    id1 = P.dimension > 1 ? (int)(a.y / P.side_length[1]) : 0;
    id2 = P.dimension > 2 ? (int)(a.z / P.side_length[2]) : 0;
}
ECMC_D int flat_cell(const DeviceProgram &P, const int id[3]) {
    return id[0] * P.cumulative[0] + id[1] * P.cumulative[1] + id[2] * P.cumulative[2];
}
ECMC_D double component(const Particle &a, int dir) { return dir == 0 ? a.x : (dir == 1 ? a.y : a.z); }
ECMC_D void set_component(Particle &a, int dir, double v) {
    if (dir == 0) a.x = v; else if (dir == 1) a.y = v; else a.z = v;
}

// Is `cell` (flat index) one of the nearby cells of the cell with identifiers (c0, c1, c2)? Per axis the periodic
// distance must be within the neighbour layers (cuboid_periodic_cells.py:74-100).
ECMC_D bool cell_is_nearby(const DeviceProgram &P, int cell, int c0, int c1, int c2) {
    const int ids[3] = {(cell / P.cumulative[0]) % P.per_side[0], (cell / P.cumulative[1]) % P.per_side[1],
                        (cell / P.cumulative[2]) % P.per_side[2]};
    const int ref[3] = {c0, c1, c2};
    bool nearby = true;
#pragma unroll
    for (int k = 0; k < 3; k++) {
        int delta = ids[k] - ref[k];
        if (delta < 0) delta += P.per_side[k];
        nearby = nearby && (delta <= P.neighbor_layers || delta >= P.per_side[k] - P.neighbor_layers);
    }
    return nearby;
}
// relative_cell(cell, active cell) as a flat index (cuboid_periodic_cells.py:155-180; modular per axis, checked on
// the host)
ECMC_D int relative_cell_of(const DeviceProgram &P, int cell, int c0, int c1, int c2) {
    const int ref[3] = {c0, c1, c2};
    int relative = 0;
#pragma unroll
    for (int k = 0; k < 3; k++) {
        int delta = (cell / P.cumulative[k]) % P.per_side[k] - ref[k];
        if (delta < 0) delta += P.per_side[k];
        relative += delta * P.cumulative[k];
    }
    return relative;
}

// SingleActiveCellOccupancy: append to the occupants of a cell, or to the surplus when the cell is full
// (single_active_cell_occupancy.py:117-121, :176-180). Serial, called by one lane. Returns the change of
// n_surplus, or 2 on overflow of the surplus capacity.
ECMC_D int occupancy_insert(int *occ, int *sur, int n_surplus, int m, int max_surplus, int cell, int id) {
    for (int s = 0; s < m; s++)
        if (occ[cell * m + s] < 0) { occ[cell * m + s] = id; return 0; }
    if (n_surplus >= max_surplus) return 2;
    sur[n_surplus] = id;
    return 1;
}
// Remove from the occupants (list.remove keeps the order of the rest; the surplus is not promoted, see
// single_active_cell_occupancy.py:186-193) or else from the surplus. Returns the change of n_surplus, 2 if absent.
ECMC_D int occupancy_remove(int *occ, int *sur, int n_surplus, int m, int cell, int id) {
    for (int s = 0; s < m; s++)
        if (occ[cell * m + s] == id) {
            for (int t = s; t + 1 < m; t++) occ[cell * m + t] = occ[cell * m + t + 1];
            occ[cell * m + m - 1] = -1;
            return 0;
        }
    for (int s = 0; s < n_surplus; s++)
        if (sur[s] == id) {
            for (int t = s; t + 1 < n_surplus; t++) sur[t] = sur[t + 1];
            return -1;
        }
    return 2;
}

// Order-preserving 64-bit key of a candidate time. All candidates of one event are `now + dt` with the same
// `now`, and Time.__add__ splits x = now.remainder + dt exactly into floor(x) and x - floor(x) (base/time.py:129-133),
// so the lexicographic comparison of (quotient, remainder) (heap.c:176-178) is the comparison of the doubles x,
// and for x >= 0 that is the comparison of their bit patterns as unsigned integers.
ECMC_D unsigned long long time_key(double x) {  // x finite; a (rounding-) negative x sorts first
    return x > 0.0 ? (unsigned long long)__double_as_longlong(x) : 0ull;
}
// argmin over the warp of (key, sequence): the lane that owns the minimum, by two 32-bit REDUX.MIN steps on the key
// and one on the sequence number (ties in time are broken like the oracle's scan: lowest sequence first)
ECMC_D int warp_argmin(unsigned long long key, int seq, int lane) {
    const unsigned hi = (unsigned)(key >> 32), lo = (unsigned)key;
    const unsigned min_hi = __reduce_min_sync(kFull, hi);
    const unsigned min_lo = __reduce_min_sync(kFull, hi == min_hi ? lo : 0xffffffffu);
    const bool tied = hi == min_hi && lo == min_lo;
    const int min_seq = __reduce_min_sync(kFull, tied ? seq : kSeqNone);
    return __ffs(__ballot_sync(kFull, tied && seq == min_seq)) - 1;
}

// Per-launch event counters of one chain, kept in registers; the rare ones (accepted vetoes, bound violations,
// capacity overflows) go straight to the global statistics, the boundary count is what remains of `events`.
struct Counters {
    unsigned events, veto;
    unsigned long long candidates, targets;
};
ECMC_D void count_rare(const RunArgs &A, int lane, int index) {
    if (lane == 0 && A.stats) atomicAdd(reinterpret_cast<unsigned long long *>(A.stats) + index, 1ull);
}

// SINGLE: the program stores one occupant per cell (the reference's default maximum_number_occupants = 1), which
// removes the slot arithmetic from the candidate gather.
// The active particle in the frame of its motion: p0 is the coordinate along the direction of motion, p1 and p2 the
// following ones cyclically (the order of permutation_3d, base/vectors.py:217-239). Targets are rotated into the same
// frame once per candidate; everything downstream is free of direction selects.
struct Moving {
    double p0, p1, p2, charge;
};
ECMC_D Moving rotate_in(const Particle &q, int dir) {
    Moving m;
    m.p0 = dir == 0 ? q.x : (dir == 1 ? q.y : q.z);
    m.p1 = dir == 0 ? q.y : (dir == 1 ? q.z : q.x);
    m.p2 = dir == 0 ? q.z : (dir == 1 ? q.x : q.y);
    m.charge = q.charge;
    return m;
}
// write the coordinates of a particle back; its charge never changes
ECMC_D void store_position(Particle *slot, const Particle &lab) {
    slot->x = lab.x; slot->y = lab.y; slot->z = lab.z;
}
ECMC_D Particle rotate_out(const Moving &m, int dir) {
    Particle q;
    q.x = dir == 0 ? m.p0 : (dir == 1 ? m.p2 : m.p1);
    q.y = dir == 0 ? m.p1 : (dir == 1 ? m.p0 : m.p2);
    q.z = dir == 0 ? m.p2 : (dir == 1 ? m.p1 : m.p0);
    q.charge = m.charge;
    return q;
}

// The unit that becomes active at the next end of chain: randint over the point masses, or for composite objects
// (randint over the roots, randint over the children), continuing in the same word stream
// (single_independent_active_periodic_direction_end_of_chain_event_handler.py:203-237).
ECMC_D int draw_end_of_chain_active(const DeviceProgram &P, const StreamKey &key) {
    if (P.nodes_per_root <= 1)
        return (int)stream_randbelow(key, ECMC_SLOT(ECMC_SLOT_END_OF_CHAIN, 0), (uint32_t)P.n_particles);
    uint32_t index = 0;
    const uint32_t root = stream_randbelow_from(key, ECMC_SLOT(ECMC_SLOT_END_OF_CHAIN, 0),
                                                (uint32_t)(P.n_particles / P.nodes_per_root), index);
    const uint32_t child = stream_randbelow_from(key, ECMC_SLOT(ECMC_SLOT_END_OF_CHAIN, 0), (uint32_t)P.nodes_per_root, index);
    return (int)(root * (uint32_t)P.nodes_per_root + child);
}

// COMPOSITE: the point masses are leaves of composite point objects (EcmcProgram.nodes_per_root > 1): the root unit of
// the active leaf is time-sliced with it, and the factor-type-map pair factors inside the active leaf's object
// (EcmcProgram.bonds) are extra candidates.
// FAR_PAIRS: the far field is ECMC_FAR_CELL_BOUNDING (one candidate per occupied non-nearby cell) instead of the
// cell veto. Both switches are compile-time so that the Lennard-Jones / cell-veto kernel carries none of their code.
template <int CAND, int REAL, int VETO, bool SINGLE, bool RECORD, int WARPS, bool COMPOSITE, bool FAR_PAIRS>
__global__ void __launch_bounds__(WARPS * 32, ECMC_RESIDENT_WARPS / WARPS)
event_kernel(const __grid_constant__ DeviceProgram P, const DeviceState S, const RunArgs A) {
    __shared__ double trig_all[UsesMic<REAL, VETO>::value ? WARPS * kTrigDoubles : 1];
    __shared__ int list_all[WARPS * 2 * kListCapacity];
    // Most events change nothing but the position of the active particle (a rejected cell veto: nine out of ten for
    // C2). The candidate list of such an event is the previous one, and so are the positions of its targets: the list
    // stays in shared memory and the (rotated) positions of the first pass are cached next to it, so that an unchanged
    // event costs neither the gather of the 27 nearby cells nor the two dependent L2 round trips behind it.
    __shared__ double cache_all[WARPS * 4 * 32];
    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    const int chain = S.first_chain + blockIdx.x * WARPS + warp;
    if (chain >= S.first_chain + S.n_chains) return;
    double *cache_p0 = cache_all + warp * 128, *cache_p1 = cache_p0 + 32, *cache_p2 = cache_p0 + 64, *cache_charge = cache_p0 + 96;
    int cached_count = -1;  // >= 0: entries of the still valid list of the previous event
    double *trig = UsesMic<REAL, VETO>::value ? trig_all + warp * kTrigDoubles : trig_all;
    int *list_target = list_all + warp * 2 * kListCapacity;  // compacted candidates of the current event
    int *list_seq = list_target + kListCapacity;

    Particle *part = S.particles + (size_t)chain * P.n_particles;
    const int m = SINGLE ? 1 : P.max_occupants;
    int *occ = S.occupants + (size_t)chain * P.n_cells * m;
    int *sur = S.surplus + (size_t)chain * P.max_surplus;
    EcmcChainState *stp = S.chains + chain;

    // chain state -> registers (uniform over the warp)
    int active = stp->active, dir = stp->direction;
    Time now = {stp->time_q, stp->time_r};
    Time eoc = {stp->eoc_q, stp->eoc_r};
    int eoc_next = stp->eoc_next_active;
    int active_cell = stp->active_cell;
    unsigned long long ev = stp->event_counter;
    const uint32_t stream = stp->stream;
    bool was_pending = stp->pending_kind != ECMC_EVENT_NONE;  // only the first iteration can start from a kept candidate
    int n_surplus = S.n_surplus[chain];
    Moving a = rotate_in(part[active], dir);
    // composite objects: the root unit of the active leaf, coordinate along the direction of motion
    Particle *roots = COMPOSITE ? S.roots + (size_t)chain * (P.n_particles / P.nodes_per_root) : nullptr;
    double root_p0 = 0.0;
    if (COMPOSITE) root_p0 = component(roots[active / P.nodes_per_root], dir);
    // per-axis identifiers of the active cell: scalars, never indexed by a runtime direction (that would put them
    // into local memory)
    int cid0 = (active_cell / P.cumulative[0]) % P.per_side[0];
    int cid1 = (active_cell / P.cumulative[1]) % P.per_side[1];
    int cid2 = (active_cell / P.cumulative[2]) % P.per_side[2];

    // lower boundary of the next cell in the direction of motion (cell_boundary_event_handler.py:135-156); 0.0 when
    // the active cell is the last one of its row. Reloaded whenever the active cell or the direction changes.
    int next_cell = 0;  // flat index of that cell
    auto next_boundary = [&]() {
        const int id = dir == 0 ? cid0 : (dir == 1 ? cid1 : cid2);
        const int nid = id + 1 == P.per_side[dir] ? 0 : id + 1;
        next_cell = active_cell + (nid - id) * P.cumulative[dir];
        return __ldg(P.cell_min_axis + dir * P.max_per_side + nid);
    };
    double boundary = next_boundary();

    const Time until = {A.until_q, A.until_r};
    const double L = P.length, half = P.half_length, speed = P.speed;
    // Measured on B200 (C2): the two-phase pass makes the kernel SLOWER (4.6e8 vs 6.0e8 events/s): the special lanes'
    // logarithm is no longer shared with the pair lanes and ptxas spills 770 instead of 470 bytes at 72 registers.
    // Kept behind ECMC_PRUNE for the next round (profiles/README.md); the default build does not contain it.
#ifdef ECMC_PRUNE
    constexpr bool PRUNE = !RECORD && !COMPOSITE && !FAR_PAIRS && CAND == ECMC_POT_LENNARD_JONES;
#else
    constexpr bool PRUNE = false;
#endif
    // The Lennard-Jones instantiations are picked only for chargeless handlers on modular cell translations
    // (pick_kernel): those configuration tests are compile-time constants there.
    constexpr bool FAST = CAND == ECMC_POT_LENNARD_JONES && REAL == 0 && (VETO == ECMC_POT_LENNARD_JONES || VETO == 0);
    const bool pair_use_charge = FAST ? false : P.pair_use_charge != 0;
    const bool veto_use_charge = FAST ? false : P.veto_use_charge != 0;
    const bool translate_modular = FAST ? true : P.translate_modular != 0;
    const bool has_pairs = FAST ? true : P.pair_handler != ECMC_PAIR_NONE;
    const int dimension = FAST ? 3 : P.dimension;
    const bool has_boundary = FAST ? true : P.no_cells == 0;  // without a cell system there is no cell boundary
    const bool cand_needs_du = needs_potential_change(resolve_kind<CAND>(P.cand_potential.kind));
    const bool has_veto = VETO != 0 && !FAR_PAIRS && P.veto_enabled == ECMC_FAR_CELL_VETO;
    const bool has_far_pairs = VETO != 0 && FAR_PAIRS;
    const unsigned max_events = A.max_events > 0 ? (unsigned)min(A.max_events, 0x7fffffffLL) : 0x7fffffffu;

    Counters n = {0, 0, 0ull, 0ull};
    unsigned n_bond_events = 0;
    bool stopped_by_time = false;

    while (n.events < max_events) {
        const StreamKey key = {P.seed, stream, ev};
        // the interaction winner: (time, kind, target particle, target cell, bounding rate)
        Time bt = time_inf();
        int bkind = ECMC_EVENT_NONE, btarget = -1, bcell = -1;
        double brate = 0.0;
        int n_cand = 0;
        bool have_confirmation = false;  // the confirmation draw of this event is already known
        double u_confirmation = 0.0;
        double kept_position = 0.0, kept_root_position = 0.0;
        Time kept_stamp = now;
        if (was_pending) {
            // a candidate that survived a host control event: nothing is recomputed, no draws are consumed
            bkind = stp->pending_kind;
            bt.q = stp->pending_q; bt.r = stp->pending_r;
            brate = stp->pending_rate;
            if (bkind == ECMC_EVENT_PAIR || (FAR_PAIRS && bkind == ECMC_EVENT_CELL_BOUNDING) ||
                (COMPOSITE && bkind == ECMC_EVENT_BOND))
                btarget = stp->pending_target;
            else bcell = stp->pending_target;
            kept_position = stp->pending_position;
            if (COMPOSITE) kept_root_position = stp->pending_root_position;
            kept_stamp.q = stp->pending_stamp_q; kept_stamp.r = stp->pending_stamp_r;
        } else {
            // Candidate gather: the occupant slots of the nearby cells, then the surplus list, are scanned 32 at a time
            // and the occupied ones are compacted (ballot + popc) into a per-warp list in shared memory. The list is
            // then worked off in passes of 32 lanes; the first pass also carries the two special candidates (lane 0
            // cell veto, lane 1 cell boundary), so 30 real pair candidates -- the ~16 occupied nearby cells of a dense
            // liquid plus ~14 surplus particles -- cost one pass. The sequence number that breaks ties follows the
            // oracle's scan: pair slots in scan order, veto, boundary.
            const int nearby_slots = has_pairs ? P.n_nearby * m : 0;
            const int n_pair_slots = has_pairs ? nearby_slots + n_surplus : 0;
            // with a cell-bounding far field every occupied cell that is not nearby is one more candidate
            // (cell_bounding_potential_tagger.py:150-155); these slots follow the pair slots, one per cell
            // composite objects: the factor-type-map factors of the active leaf's object come after the pair slots
            const int n_bond_slots = COMPOSITE ? P.n_bonds : 0;
            const int far_base = n_pair_slots + n_bond_slots;
            const int n_scan_slots = far_base + (has_far_pairs ? P.n_cells : 0);
            const int special_seq = FAR_PAIRS ? far_base + P.n_cells : far_base;
            const double c_active = pair_use_charge ? a.charge : 1.0;
            unsigned long long best_key = 0x7ff0000000000000ull;  // best of the passes so far (uniform)
            double best_x = INFINITY;
            int best_seq = kSeqNone;
            int cursor = 0;      // next slot to scan
            bool first = true;   // the special candidates have not been processed yet
            double d_star = INFINITY;  // PRUNE: no candidate beyond this displacement can win the event
            while (first || cursor < n_scan_slots) {
              int count = 0;  // entries in the compact list
              const bool from_cache = first && cached_count >= 0;
              if (from_cache) { count = cached_count; cursor = n_scan_slots; }
              while (cursor < n_scan_slots && count <= kListCapacity - 32) {
                const int s = cursor + lane;
                int found = -1;
                if (s < nearby_slots) {
                    const int ci = m == 1 ? s : s / m;
                    const int code = __ldg(P.nearby + ci);
                    int x = cid0 + (code & 1023), y = cid1 + ((code >> 10) & 1023), z = cid2 + (code >> 20);
                    if (x >= P.per_side[0]) x -= P.per_side[0];
                    if (y >= P.per_side[1]) y -= P.per_side[1];
                    if (z >= P.per_side[2]) z -= P.per_side[2];
                    const int cell = x * P.cumulative[0] + y * P.cumulative[1] + z * P.cumulative[2];
                    found = occ[cell * m + (s - ci * m)];
                } else if (s < n_pair_slots) {
                    found = sur[s - nearby_slots];
                } else if (COMPOSITE && s < far_base) {
                    const int b = s - n_pair_slots;
                    const int root = active / P.nodes_per_root, child = active - root * P.nodes_per_root;
                    const int partner = P.bonds[b][0] == child ? P.bonds[b][1] : (P.bonds[b][1] == child ? P.bonds[b][0] : -1);
                    if (partner >= 0) found = root * P.nodes_per_root + partner;
                } else if (FAR_PAIRS && s < n_scan_slots) {
                    const int cell = s - far_base;
                    if (!cell_is_nearby(P, cell, cid0, cid1, cid2)) found = occ[cell];
                }
                const unsigned occupied = __ballot_sync(kFull, found >= 0);
                if (found >= 0) {
                    const int rank = count + __popc(occupied & ((1u << lane) - 1u));
                    list_target[rank] = found;
                    list_seq[rank] = s;
                }
                count += __popc(occupied);
                cursor += 32;
              }
              __syncwarp();
              if (first) cached_count = cursor >= n_scan_slots ? count : -1;  // reusable if it is the whole list
              n.targets += (unsigned long long)count;
              const int shift = first ? 2 : 0;
              for (int base = 0; base < count + shift; base += 32) {
                const int entry = base + lane - shift;  // -2 = veto, -1 = boundary
                int target = -1, s = -1;
                if (entry >= 0 && entry < count) { target = list_target[entry]; s = list_seq[entry]; }
                const bool is_pair = target >= 0;
                const bool is_veto = entry == -2 && has_veto;
                const bool is_boundary = entry == -1 && has_boundary;
                Moving tp;
                if (is_pair) {
                    if (from_cache && base == 0) {
                        tp.p0 = cache_p0[lane]; tp.p1 = cache_p1[lane]; tp.p2 = cache_p2[lane]; tp.charge = cache_charge[lane];
                    } else {
                        tp = rotate_in(part[target], dir);
                        if (first && base == 0) {
                            cache_p0[lane] = tp.p0; cache_p1[lane] = tp.p1; cache_p2[lane] = tp.p2; cache_charge[lane] = tp.charge;
                        }
                    }
                }
                // One Philox block per lane, all lanes together: pair lanes draw their potential change (slot keyed
                // by the target), the veto lane its (Walker uniform, time) pair, and the boundary lane -- which needs
                // no random number -- computes the veto lane's table-index words.
                // If the last lane of the first pass is idle it draws the confirmation number of the out-state ahead
                // of time (slot CONFIRM): a third of all events need it, and here it costs nothing.
                const bool draws_confirmation = first && base == 0 && lane == 31 && !is_pair;
                const bool is_bond = COMPOSITE && is_pair && s >= n_pair_slots && s < far_base;
                const uint32_t slot = is_pair ? ECMC_SLOT(is_bond ? ECMC_SLOT_FACTOR_TIME : ECMC_SLOT_PAIR_TIME, target)
                                              : (is_boundary ? ECMC_SLOT(ECMC_SLOT_VETO_CHOICE, 0)
                                                             : (draws_confirmation ? ECMC_SLOT(ECMC_SLOT_CONFIRM, 0)
                                                                                   : ECMC_SLOT(ECMC_SLOT_VETO_TIME, 0)));
                Philox4 b = stream_block(key, slot, 0);
                const double u_first = words_to_double(b.w[0], b.w[1]), u_second = words_to_double(b.w[2], b.w[3]);
                if (first && base == 0) {
                    have_confirmation = __shfl_sync(kFull, (int)draws_confirmation, 31) != 0;
                    u_confirmation = __shfl_sync(kFull, u_first, 31);
                }
                // Passes after the first hold surplus particles only, scattered over the whole box: most of them
                // cannot fire (their potential change exceeds the depth of the attractive tail), which one division
                // decides. If that holds for every lane the pass ends here, without logarithm and inversion.
                double s0 = 0.0, s1 = 0.0, s2 = 0.0;
                if (is_pair) {
                    s0 = correct_separation_in_box(tp.p0 - a.p0, L, half);
                    s1 = correct_separation_in_box(tp.p1 - a.p1, L, half);
                    s2 = correct_separation_in_box(tp.p2 - a.p2, L, half);
                }
                // PRUNE (Lennard-Jones, no event records): the first pass runs in two phases. Phase 0 evaluates the
                // special lanes only; their earliest time, and the end of chain, bound how far the active particle can
                // move in this event. Phase 1 evaluates the pair lanes -- but only if one of them may fire within that
                // displacement (lj_may_fire_within: one division, no logarithm). In nine events out of ten none can,
                // and the event costs neither the pair lanes' logarithm nor their inversion. The winner is unchanged;
                // only the count of finite candidates (EcmcStats.candidates) then covers the evaluated lanes only.
                const bool prune_first = PRUNE && first && base == 0;
                bool alive = true;
                if (base > 0 || !first) {
                    alive = is_pair && (((COMPOSITE || FAR_PAIRS) && s >= n_pair_slots) || !certainly_dead<CAND>(P.cand_potential, s0, s1, s2,
                                                                        cand_needs_du ? u_first * P.inv_beta : 0.0));
                    if (PRUNE)
                        alive = alive && lj_may_fire_within(P.cand_potential.lj, s0, fma(s1, s1, s2 * s2), d_star, u_first * P.inv_beta);
                    if (!__any_sync(kFull, alive)) continue;
                }
                // the table-index words travel from the boundary lane (lane 1 of pass 0) to the veto lane (lane 0)
                uint32_t choice[4];
#pragma unroll
                for (int j = 0; j < 4; j++) choice[j] = __shfl_down_sync(kFull, b.w[j], 1);
                double dt = INFINITY;
                int kind = ECMC_EVENT_NONE, cell = -1, seq = kSeqNone;
                double rate = 0.0;
                const bool is_far = FAR_PAIRS && is_pair && s >= far_base;  // cell-bounding candidate
                for (int phase = 0; phase < (prune_first ? 2 : 1); phase++) {
                bool on = !PRUNE || alive;
                if (prune_first) {
                    if (phase == 0) {
                        on = !is_pair;
                    } else {
                        const double x0 = __shfl_sync(kFull, now.r + dt, 0), x1 = __shfl_sync(kFull, now.r + dt, 1);
                        const double eoc_x = (eoc.q - now.q) + eoc.r;
                        const double t_rel = fmin(fmin(x0, x1), eoc_x) - now.r;
                        d_star = fma(speed * t_rel, 1.0 + 1.0e-9, 1.0e-12);
                        on = is_pair && lj_may_fire_within(P.cand_potential.lj, s0, fma(s1, s1, s2 * s2), d_star, u_first * P.inv_beta);
                        if (!__any_sync(kFull, on)) break;
                    }
                }
                if (!on) continue;
                // random.expovariate(beta): pair lanes use their first double, the veto lane its second
                const double exponential = -log_unit_interval(1.0 - (is_veto ? u_second : u_first)) * P.inv_beta;
                if (is_bond) {
                    // TwoLeafUnitEventHandler.send_event_time with the factor's own potential
                    dt = displacement_time<-1>(P.bond_potential, 0, P.inv_speed, L, s0, s1, s2, 1.0, 1.0,
                                               needs_potential_change(P.bond_potential.kind) ? exponential : 0.0);
                    kind = ECMC_EVENT_BOND;
                    seq = s;
                } else if (is_far) {
                    // TwoLeafUnitCellBoundingPotentialEventHandler.send_event_time (:137-177): constant event rate =
                    // bound of the relative cell x charge correction factor (cell_bounding_potential.py:155-238)
                    const int relative = relative_cell_of(P, s - far_base, cid0, cid1, cid2);
                    double charge_product = 1.0;
                    if (veto_use_charge) charge_product = a.charge * tp.charge / P.veto_target_charge;
                    const double *bound = P.bounds + (relative * P.dimension + dir) * 2;
                    rate = charge_product > 0.0 ? __ldg(bound) * charge_product : -__ldg(bound + 1) * charge_product;
                    dt = rate > 0.0 ? exponential / rate * P.inv_speed : INFINITY;
                    cell = relative;
                    kind = ECMC_EVENT_CELL_BOUNDING;
                    seq = s;
                } else if (is_pair) {
                    dt = displacement_time<CAND>(P.cand_potential, 0, P.inv_speed, L, s0, s1, s2, c_active,
                                                 pair_use_charge ? tp.charge : 1.0, cand_needs_du ? exponential : 0.0);
                    kind = ECMC_EVENT_PAIR;
                    seq = s;
                } else if (is_veto) {
                    // CellVetoEventHandler.send_event_time (cell_veto_event_handler.py:200-238)
                    // InnerPointEstimator.charge_correction_factor (inner_point_estimator.py:165-192)
                    double charge_factor = 1.0;
                    if (veto_use_charge) {
                        charge_factor = a.charge * 1.0;
                        if (P.veto_target_charge != 1.0) charge_factor = charge_factor / P.veto_target_charge;
                    }
                    const DeviceWalker *w = &P.upper[dir];
                    if (!(charge_factor > 0.0)) { charge_factor *= -1.0; w = &P.lower[dir]; }
                    // random.choice(table) = table[_randbelow(n)]: rejection on the top bits of successive words
                    uint32_t e = 0;
                    bool found = false;
#pragma unroll
                    for (int j = 0; j < 4; j++) {
                        const uint32_t r = choice[j] >> (32 - w->bits);
                        if (!found && r < (uint32_t)w->n_entries) { e = r; found = true; }
                    }
                    for (uint32_t block = 1; !found; block++) {
                        const Philox4 more = stream_block(key, ECMC_SLOT(ECMC_SLOT_VETO_CHOICE, 0), block);
#pragma unroll
                        for (int j = 0; j < 4; j++) {
                            const uint32_t r = more.w[j] >> (32 - w->bits);
                            if (!found && r < (uint32_t)w->n_entries) { e = r; found = true; }
                        }
                    }
                    const WalkerEntry entry = w->entries[e];
                    // Walker.sample_cell (walker.py:114-118): random.uniform(0.0, mean) <= rate
                    const bool first = 0.0 + (w->mean_rate - 0.0) * u_first <= entry.rate_a;
                    const int relative = first ? entry.cell_a : entry.cell_b;
                    rate = (first ? entry.bound_a : entry.bound_b) * charge_factor;
                    // translate(active cell, relative cell), per axis (cuboid_periodic_cells.py:182-207)
                    int tx, ty, tz;
                    const int rx = relative & 1023, ry = (relative >> 10) & 1023, rz = relative >> 20;
                    if (translate_modular) {
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
                    // expovariate(beta) / (total rate * charge factor * speed): the reciprocal of the table's part is
                    // precomputed, chargeless handlers never divide
                    dt = exponential * w->inv_total_rate_speed;
                    // (a neutral active unit has no cell-veto events: a zero rate is an infinite time, not -inf)
                    if (veto_use_charge) dt = charge_factor > 0.0 ? exponential / (w->total_rate * charge_factor * speed) : INFINITY;
                    kind = ECMC_EVENT_CELL_VETO;
                    seq = special_seq;
                } else if (is_boundary) {
                    // CellBoundaryEventHandler.send_event_time (cell_boundary_event_handler.py:122-156)
                    double separation = boundary - a.p0;
                    if (separation < 0.0) separation = separation + L;  // next_image, hypercubic_setting.py:191
                    dt = separation * P.inv_speed;
                    cell = next_cell;
                    kind = ECMC_EVENT_CELL_BOUNDARY;
                    seq = special_seq + 1;
                }
                }
                // Time.__add__: event time = now + dt; x orders the candidates (see time_key)
                const double x = now.r + dt;
                const bool finite = kind != ECMC_EVENT_NONE && x < INFINITY;  // heap_scheduler.py:139; NaN never wins
                n_cand += __popc(__ballot_sync(kFull, finite));
                const unsigned long long k64 = finite ? time_key(x) : 0x7ff0000000000000ull;
                // argmin of this pass, then against the earlier passes (usually there is one pass)
                const int owner = warp_argmin(k64, finite ? seq : kSeqNone, lane);
                const unsigned long long pass_key = __shfl_sync(kFull, k64, owner);
                const int pass_seq = __shfl_sync(kFull, finite ? seq : kSeqNone, owner);
                if (pass_key < best_key || (pass_key == best_key && pass_seq < best_seq)) {
                    best_key = pass_key; best_seq = pass_seq;
                    best_x = __shfl_sync(kFull, x, owner);
                    bkind = __shfl_sync(kFull, kind, owner);
                    btarget = __shfl_sync(kFull, target, owner);
                    bcell = __shfl_sync(kFull, cell, owner);
                    brate = __shfl_sync(kFull, rate, owner);
                }
              }
              first = false;
              __syncwarp();
            }
            if (best_seq != kSeqNone) {
                const double fl = floor(best_x);
                bt.q = now.q + fl; bt.r = best_x - fl;
            } else {
                bkind = ECMC_EVENT_NONE;
            }
        }

        // the end-of-chain candidate lives in the scheduler since the chain started
        n_cand++;
        const bool eoc_first = time_lt(eoc, bt);
        const Time event_time = eoc_first ? eoc : bt;
        const int kind = eoc_first ? ECMC_EVENT_END_OF_CHAIN : bkind;
        if (!time_lt(event_time, until)) {
            // a host control event comes first: the interaction winner stays scheduled
            if (lane == 0) {
                stp->pending_kind = bkind;
                stp->pending_q = bt.q; stp->pending_r = bt.r;
                stp->pending_rate = brate;
                stp->pending_target = (bkind == ECMC_EVENT_PAIR || (FAR_PAIRS && bkind == ECMC_EVENT_CELL_BOUNDING) ||
                                       (COMPOSITE && bkind == ECMC_EVENT_BOND)) ? btarget : bcell;
                if (!was_pending) {
                    stp->pending_position = a.p0;
                    if (COMPOSITE) stp->pending_root_position = root_p0;
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
                a.p0 = kept_position;
                if (COMPOSITE) root_p0 = kept_root_position;
                now = kept_stamp;
            }
            was_pending = false;
        }

        // ---- out-state: time slice of the active particle (event_handler/abstracts/abstracts.py:82-95)
        const double x_before = a.p0;
        {
            const double dt = time_sub(event_time, now);
            a.p0 = correct_position_entry(__dadd_rn(x_before, __dmul_rn(speed, dt)), L);
            // the root unit moves with velocity * weight and carries the same time stamp (abstracts.py:82-101,165-190)
            if (COMPOSITE) root_p0 = correct_position_entry(__dadd_rn(root_p0, __dmul_rn(P.root_speed, dt)), L);
            now = event_time;
        }
        // Did the time slice itself carry the particle out of its cell (without a boundary event)? Cell `id` holds
        // exactly the doubles in [cell_min[id], cell_min[id + 1]) and the particle only moves forward.
        const bool left_cell = boundary == 0.0 ? a.p0 < x_before : a.p0 >= boundary;
        int new_active = active, accepted = 0, rec_target = -1;
        switch (kind) {
        case ECMC_EVENT_PAIR: {
            rec_target = btarget;
            if (P.pair_handler == ECMC_PAIR_TWO_LEAF_UNIT) {
                accepted = 1;  // two_leaf_unit_event_handler.py:140-154
            } else {
                // two_leaf_unit_bounding_potential_event_handler.py:148-168 +
                // event_handler_with_bounding_potential.py:75-101
                const Moving tp = rotate_in(part[btarget], dir);
                const double sx = correct_separation_in_box(tp.p0 - a.p0, L, half);
                const double sy = correct_separation_in_box(tp.p1 - a.p1, L, half);
                const double sz = correct_separation_in_box(tp.p2 - a.p2, L, half);
                const double c1 = pair_use_charge ? a.charge : 1.0, c2 = pair_use_charge ? tp.charge : 1.0;
                const double bounding_rate = derivative_warp<CAND>(P.cand_potential, 0, speed, sx, sy, sz, c1, c2, trig, lane);
                const double real = derivative_warp<REAL>(P.real_potential, 0, speed, sx, sy, sz, c1, c2, trig, lane);
                if (real > 0.0) {
                    if (bounding_rate < real) count_rare(A, lane, 7);
                    const double u = have_confirmation ? u_confirmation : stream_double(key, ECMC_SLOT(ECMC_SLOT_CONFIRM, 0), 0);
                    if (0.0 + (bounding_rate - 0.0) * u < real) accepted = 1;
                }
            }
            if (accepted) new_active = btarget;
            count_rare(A, lane, 1);
            break;
        }
        case ECMC_EVENT_CELL_VETO: {
            // mediator.py:265-292 + leaf_unit_cell_veto_event_handler.py:117-149
            const int t = occ[bcell * m];
            rec_target = t;
            if (t >= 0) {
                const Moving tp = rotate_in(part[t], dir);
                const double sx = correct_separation_in_box(tp.p0 - a.p0, L, half);
                const double sy = correct_separation_in_box(tp.p1 - a.p1, L, half);
                const double sz = correct_separation_in_box(tp.p2 - a.p2, L, half);
                const double c1 = veto_use_charge ? a.charge : 1.0, c2 = veto_use_charge ? tp.charge : 1.0;
                const double real = derivative_warp<VETO>(P.veto_potential, 0, speed, sx, sy, sz, c1, c2, trig, lane);
                if (real > 0.0) {
                    if (brate < real) count_rare(A, lane, 7);
                    const double u = have_confirmation ? u_confirmation : stream_double(key, ECMC_SLOT(ECMC_SLOT_CONFIRM, 0), 0);
                    if (0.0 + (brate - 0.0) * u < real) { accepted = 1; new_active = t; }
                }
            }
            n.veto++;
            if (accepted) count_rare(A, lane, 3);
            break;
        }
        case ECMC_EVENT_CELL_BOUNDING: {
            // TwoLeafUnitCellBoundingPotentialEventHandler.send_out_state (:179-211): the stored bounding rate (times the
            // speed) against the real derivative, event_handler_with_bounding_potential.py:75-101
            if (!FAR_PAIRS) break;
            rec_target = btarget;
            const Moving tp = rotate_in(part[btarget], dir);
            const double sx = correct_separation_in_box(tp.p0 - a.p0, L, half);
            const double sy = correct_separation_in_box(tp.p1 - a.p1, L, half);
            const double sz = correct_separation_in_box(tp.p2 - a.p2, L, half);
            const double c1 = veto_use_charge ? a.charge : 1.0, c2 = veto_use_charge ? tp.charge : 1.0;
            const double bounding_rate = brate * speed;
            const double real = derivative_warp<VETO>(P.veto_potential, 0, speed, sx, sy, sz, c1, c2, trig, lane);
            if (real > 0.0) {
                if (bounding_rate < real) count_rare(A, lane, 7);
                const double u = have_confirmation ? u_confirmation : stream_double(key, ECMC_SLOT(ECMC_SLOT_CONFIRM, 0), 0);
                if (0.0 + (bounding_rate - 0.0) * u < real) accepted = 1;
            }
            if (accepted) new_active = btarget;
            count_rare(A, lane, 1);
            break;
        }
        case ECMC_EVENT_BOND: {
            // TwoLeafUnitEventHandler.send_out_state (two_leaf_unit_event_handler.py:140-154)
            if (!COMPOSITE) break;
            rec_target = btarget;
            accepted = 1;
            new_active = btarget;
            n_bond_events++;
            break;
        }
        case ECMC_EVENT_CELL_BOUNDARY: {
            // lands exactly on the lower boundary of the new cell (cell_boundary_event_handler.py:158-173)
            count_rare(A, lane, 4);
            a.p0 = boundary;
            break;
        }
        case ECMC_EVENT_END_OF_CHAIN:
            // abstracts/end_of_chain_event_handler.py:107-187, new direction = (d + 1) mod D
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
            rec.new_direction = kind == ECMC_EVENT_END_OF_CHAIN ? (dir + 1) % dimension : dir;
            rec.mode = 0;
            rec.time_q = event_time.q; rec.time_r = event_time.r;
            const Particle lab = rotate_out(a, dir);
            rec.active_pos[0] = lab.x; rec.active_pos[1] = lab.y; rec.active_pos[2] = lab.z;
            A.records[(size_t)chain * A.records_per_chain + n.events] = rec;
        }
        ev++;
        n.events++;
        n.candidates += (unsigned long long)n_cand;
        Particle lab;  // the old active particle in the lab frame, where the rare paths below need it
        const bool needs_lab = new_active != active || kind == ECMC_EVENT_CELL_BOUNDARY ||
                               kind == ECMC_EVENT_END_OF_CHAIN || left_cell;
        if (needs_lab) lab = rotate_out(a, dir);
        if (COMPOSITE && (new_active != active || kind == ECMC_EVENT_END_OF_CHAIN)) {
            // the root of the old active leaf stops here (or turns with the chain); the new one is picked up below
            if (lane == 0) set_component(roots[active / P.nodes_per_root], dir, root_p0);
            __syncwarp();
        }
        if (kind == ECMC_EVENT_END_OF_CHAIN) dir = dir + 1 == dimension ? 0 : dir + 1;
        if (COMPOSITE && (new_active != active || kind == ECMC_EVENT_END_OF_CHAIN))
            root_p0 = component(roots[new_active / P.nodes_per_root], dir);

        // ---- SingleActiveCellOccupancy.update (single_active_cell_occupancy.py:149-203)
        if (new_active != active) {
            cached_count = -1;
            int delta = 0;
            if (lane == 0) {
                store_position(part + active, lab);
                delta = occupancy_insert(occ, sur, n_surplus, m, P.max_surplus, active_cell, active);
            }
            delta = __shfl_sync(kFull, delta, 0);
            if (delta == 2) count_rare(A, lane, 8); else n_surplus += delta;
            __syncwarp();
            active = new_active;
            lab = part[active];
            a = rotate_in(lab, dir);
            cell_identifier_of(P, lab, cid0, cid1, cid2);
            active_cell = cid0 * P.cumulative[0] + cid1 * P.cumulative[1] + cid2 * P.cumulative[2];
            delta = 0;
            if (lane == 0) delta = occupancy_remove(occ, sur, n_surplus, m, active_cell, active);
            delta = __shfl_sync(kFull, delta, 0);
            if (delta == 2) count_rare(A, lane, 8); else n_surplus += delta;
            __syncwarp();
            boundary = next_boundary();
        } else if (kind == ECMC_EVENT_CELL_BOUNDARY || kind == ECMC_EVENT_END_OF_CHAIN || left_cell) {
            cached_count = -1;
            // the oracle recomputes the cell from the position after every event; only these can change it
            a = rotate_in(lab, dir);
            cell_identifier_of(P, lab, cid0, cid1, cid2);
            active_cell = cid0 * P.cumulative[0] + cid1 * P.cumulative[1] + cid2 * P.cumulative[2];
            boundary = next_boundary();
        }
        if (kind == ECMC_EVENT_END_OF_CHAIN) {
            // the next end-of-chain candidate: chain_time after this one, new active by randint
            // (single_independent_active_periodic_direction_end_of_chain_event_handler.py:203-237)
            eoc = time_add(now, time_sub(now, now) + P.chain_time);
            const StreamKey next_key = {P.seed, stream, ev};
            eoc_next = COMPOSITE ? draw_end_of_chain_active(P, next_key)
                                 : (int)stream_randbelow(next_key, ECMC_SLOT(ECMC_SLOT_END_OF_CHAIN, 0), (uint32_t)P.n_particles);
        }
    }

    if (stopped_by_time) {
        // the sampling / end-of-run handler time-slices the active unit (fixed_interval_sampling_event_handler.py:96-109)
        const double dt = time_sub(until, now);
        a.p0 = correct_position_entry(__dadd_rn(a.p0, __dmul_rn(speed, dt)), L);
        if (COMPOSITE) root_p0 = correct_position_entry(__dadd_rn(root_p0, __dmul_rn(P.root_speed, dt)), L);
        now = until;
    }
    if (lane == 0) {
        store_position(part + active, rotate_out(a, dir));
        if (COMPOSITE) set_component(roots[active / P.nodes_per_root], dir, root_p0);
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
            if (COMPOSITE && n_bond_events) atomicAdd(st + 9, (unsigned long long)n_bond_events);
            if (n.candidates) atomicAdd(st + 6, n.candidates);
            if (n.targets) atomicAdd(st + 11, n.targets);
        }
    }
}

// ---------------------------------------------------------------------------------------------------------
// start of run: SingleActiveCellOccupancy.initialize (single_active_cell_occupancy.py:95-121) +
// InitialChainStartOfRunEventHandler (initial_chain_start_of_run_event_handler.py:92-131). One warp per chain.
// Particles enter in identifier order: the first max_occupants of a cell become occupants, the rest surplus.
// ---------------------------------------------------------------------------------------------------------
// start of run of ONE chain by its warp (all 32 lanes call it).
// keep_state (host steps that CONTINUE their chains, ECMC_OPTION_CONTINUE_HOST_STEPS): the configuration was replaced by
// the caller's, the lifting state of the chain -- active particle, direction, clock, end of chain, event counter, random
// stream -- stays what the last launch left; only the cell occupancy is rebuilt from the positions
// (SingleActiveCellOccupancy.initialize, :95-121) and a kept candidate is dropped.
ECMC_D void start_chain(const DeviceProgram &P, const DeviceState &S, const uint32_t *streams, uint32_t first_stream,
                        int initial_active, int initial_direction, EcmcStats *stats, int chain, int lane,
                        bool keep_state = false) {
    Particle *part = S.particles + (size_t)chain * P.n_particles;
    const int m = P.max_occupants;
    int *occ = S.occupants + (size_t)chain * P.n_cells * m;
    int *sur = S.surplus + (size_t)chain * P.max_surplus;
    for (int i = lane; i < P.n_cells * m; i += 32) occ[i] = -1;
    __syncwarp();
    int n_surplus = 0, overflow = 0;
    if (m == 1) {
        // lowest identifier wins the cell; everyone else goes to the surplus in identifier order
        for (int i = lane; i < P.n_particles; i += 32) {
            int id[3];
            const Particle p = part[i];
            cell_identifier_of(P, p, id);
            atomicMin(reinterpret_cast<unsigned int *>(occ + flat_cell(P, id)), (unsigned int)i);
        }
        __syncwarp();
        for (int base = 0; base < P.n_particles; base += 32) {
            const int i = base + lane;
            bool extra = false;
            if (i < P.n_particles) {
                int id[3];
                const Particle p = part[i];
                cell_identifier_of(P, p, id);
                extra = occ[flat_cell(P, id)] != i;
            }
            const unsigned mask = __ballot_sync(kFull, extra);
            if (extra) {
                const int slot = n_surplus + __popc(mask & ((1u << lane) - 1u));
                if (slot < P.max_surplus) sur[slot] = i;
            }
            n_surplus += __popc(mask);
        }
        if (n_surplus > P.max_surplus) { overflow = n_surplus - P.max_surplus; n_surplus = P.max_surplus; }
    } else {
        if (lane == 0) {
            for (int i = 0; i < P.n_particles; i++) {
                int id[3];
                const Particle p = part[i];
                cell_identifier_of(P, p, id);
                const int delta = occupancy_insert(occ, sur, n_surplus, m, P.max_surplus, flat_cell(P, id), i);
                if (delta == 2) overflow++; else n_surplus += delta;
            }
        }
        n_surplus = __shfl_sync(kFull, n_surplus, 0);
        overflow = __shfl_sync(kFull, overflow, 0);
    }
    __syncwarp();
    if (lane == 0 && keep_state) {
        EcmcChainState st = S.chains[chain];
        int id[3];
        const Particle a = part[st.active];
        cell_identifier_of(P, a, id);
        st.active_cell = flat_cell(P, id);
        const int delta = occupancy_remove(occ, sur, n_surplus, m, st.active_cell, st.active);
        if (delta == 2) overflow++; else n_surplus += delta;
        st.pending_kind = ECMC_EVENT_NONE;
        st.kept_kind = 0;
        S.chains[chain] = st;
        S.n_surplus[chain] = n_surplus;
        if (overflow && stats) atomicAdd(reinterpret_cast<unsigned long long *>(stats) + 8, (unsigned long long)overflow);
    } else if (lane == 0) {
        EcmcChainState st = {};  // every field defined: the state is downloaded, compared and checkpointed as bytes
        st.active = initial_active; st.direction = initial_direction;
        st.time_q = 0.0; st.time_r = 0.0;
        st.event_counter = 0;
        st.stream = streams ? streams[chain] : first_stream + (uint32_t)chain;
        int id[3];
        const Particle a = part[initial_active];
        cell_identifier_of(P, a, id);
        st.active_cell = flat_cell(P, id);
        const int delta = occupancy_remove(occ, sur, n_surplus, m, st.active_cell, initial_active);
        if (delta == 2) overflow++; else n_surplus += delta;
        const Time now = {0.0, 0.0};
        const Time eoc = time_add(now, time_sub(now, now) + P.chain_time);
        st.eoc_q = eoc.q; st.eoc_r = eoc.r;
        const StreamKey key = {P.seed, st.stream, 0ull};
        st.eoc_next_active = draw_end_of_chain_active(P, key);
        st.pending_kind = ECMC_EVENT_NONE; st.pending_target = 0; st.mode = 0;
        st.pending_q = 0.0; st.pending_r = 0.0; st.pending_rate = 0.0; st.pending_position = 0.0;
        st.pending_root_position = 0.0;
        st.pending_stamp_q = 0.0; st.pending_stamp_r = 0.0;
        S.chains[chain] = st;
        S.n_surplus[chain] = n_surplus;
        if (overflow && stats) atomicAdd(reinterpret_cast<unsigned long long *>(stats) + 8, (unsigned long long)overflow);
    }
}

template <int WARPS>
__global__ void __launch_bounds__(WARPS * 32)
start_kernel(const __grid_constant__ DeviceProgram P, const DeviceState S, const uint32_t *streams, uint32_t first_stream,
             int initial_active, int initial_direction, EcmcStats *stats, bool keep_state) {
    const int lane = threadIdx.x & 31;
    const int chain = S.first_chain + blockIdx.x * WARPS + (threadIdx.x >> 5);
    if (chain >= S.first_chain + S.n_chains) return;
    start_chain(P, S, streams, first_stream, initial_active, initial_direction, stats, chain, lane, keep_state);
}

// host layout [n_chains][n_particles][dimension] (+ charges [n_chains][n_particles]) <-> 32-byte particle records
__global__ void pack_particles_kernel(const double *positions, const double *charges, Particle *out, size_t n, int dimension) {
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        Particle p;
        p.x = positions[i * dimension];
        p.y = dimension > 1 ? positions[i * dimension + 1] : 0.0;
        p.z = dimension > 2 ? positions[i * dimension + 2] : 0.0;
        p.charge = charges ? charges[i] : 1.0;
        out[i] = p;
    }
}
__global__ void unpack_particles_kernel(const Particle *in, double *positions, size_t n, int dimension) {
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        const Particle p = in[i];
        positions[i * dimension] = p.x;
        if (dimension > 1) positions[i * dimension + 1] = p.y;
        if (dimension > 2) positions[i * dimension + 2] = p.z;
    }
}

// ---------------------------------------------------------------------------------------------------------
// Pair-separation histogram (SeparationOutputHandler, separation_output_handler.py:75-97): one block per
// (chain, tile of 1024 particles j). The tile is staged in shared memory, every thread walks over particles i and
// bins |r_ij| for the j > i of the tile into a shared-memory histogram, flushed to the global one at the end.
// Streaming, coalesced: each position is read once per tile from L2/HBM; the histogram never leaves the SM.
// ---------------------------------------------------------------------------------------------------------
constexpr int kHistogramTile = 1024;
constexpr int kHistogramMaxBins = 4096;

// `first`, `stride`: only the particles first, first + stride, ... take part (n_particles counts those): stride =
// nodes_per_root and first = child index select one leaf of every composite object, e.g. the oxygens of water
// (OxygenOxygenSeparationOutputHandler, oxygen_oxygen_separation_output_handler.py).
__global__ void __launch_bounds__(256)
separation_histogram_kernel(const Particle *particles, int n_particles, int n_tiles, double length, int n_bins,
                            double r_min, double inv_bin_width, unsigned long long *histogram, int first, int stride,
                            int particles_per_chain) {
    __shared__ double tile_x[kHistogramTile], tile_y[kHistogramTile], tile_z[kHistogramTile];
    __shared__ unsigned int bins[kHistogramMaxBins];
    const int chain = blockIdx.x / n_tiles, tile = blockIdx.x % n_tiles;
    const Particle *part = particles + (size_t)chain * particles_per_chain + first;
    const int j0 = tile * kHistogramTile, j1 = min(n_particles, j0 + kHistogramTile);
    for (int b = threadIdx.x; b < n_bins; b += blockDim.x) bins[b] = 0;
    for (int j = j0 + threadIdx.x; j < j1; j += blockDim.x) {
        const Particle p = part[(size_t)j * stride];
        tile_x[j - j0] = p.x; tile_y[j - j0] = p.y; tile_z[j - j0] = p.z;
    }
    __syncthreads();
    const double half = 0.5 * length;
    for (int i = threadIdx.x; i < j1 - 1; i += blockDim.x) {
        const Particle p = part[(size_t)i * stride];
        for (int j = max(i + 1, j0); j < j1; j++) {
            const double sx = correct_separation_in_box(tile_x[j - j0] - p.x, length, half);
            const double sy = correct_separation_in_box(tile_y[j - j0] - p.y, length, half);
            const double sz = correct_separation_in_box(tile_z[j - j0] - p.z, length, half);
            const double r = sqrt(fma(sx, sx, fma(sy, sy, sz * sz)));
            const double scaled = (r - r_min) * inv_bin_width;
            if (scaled >= 0.0 && scaled <= (double)n_bins) {
                const int bin = min((int)scaled, n_bins - 1);  // r == r_max belongs to the last bin (numpy.histogram)
                atomicAdd(&bins[bin], 1u);
            }
        }
    }
    __syncthreads();
    for (int b = threadIdx.x; b < n_bins; b += blockDim.x)
        if (bins[b]) atomicAdd(histogram + b, (unsigned long long)bins[b]);
}

// PolarizationOutputHandler.write (polarization_output_handler.py:76-101): per chain the sum over all leaf units of
// charge x (leaf position closest to its root unit, base/node.py:164-188), in the reference's order of additions (root
// by root, leaf by leaf, component by component) -- one thread per chain, the observable is sampled rarely.
__global__ void __launch_bounds__(128)
polarization_kernel(const Particle *particles, const Particle *roots, const double *charges, int n_chains, int n_roots,
                    int nodes_per_root, int dimension, double length, double *out) {
    const int chain = blockIdx.x * blockDim.x + threadIdx.x;
    if (chain >= n_chains) return;
    const Particle *part = particles + (size_t)chain * n_roots * nodes_per_root;
    const Particle *root = roots + (size_t)chain * n_roots;
    const double half = 0.5 * length;
    double px = 0.0, py = 0.0, pz = 0.0;
    for (int r = 0; r < n_roots; r++) {
        const Particle c = root[r];
        for (int k = 0; k < nodes_per_root; k++) {
            const Particle leaf = part[r * nodes_per_root + k];
            const double q = charges ? charges[r * nodes_per_root + k] : leaf.charge;
            px = __dadd_rn(px, __dmul_rn(q, __dadd_rn(c.x, correct_separation_in_box(leaf.x - c.x, length, half))));
            py = __dadd_rn(py, __dmul_rn(q, __dadd_rn(c.y, correct_separation_in_box(leaf.y - c.y, length, half))));
            if (dimension > 2)
                pz = __dadd_rn(pz, __dmul_rn(q, __dadd_rn(c.z, correct_separation_in_box(leaf.z - c.z, length, half))));
        }
    }
    out[(size_t)chain * dimension] = px;
    out[(size_t)chain * dimension + 1] = py;
    if (dimension > 2) out[(size_t)chain * dimension + 2] = pz;
}

// BondLengthAndAngleOutputHandler.write (bond_length_and_angle_output_handler.py:77-103) for objects of three leaves
// (hydrogen, oxygen, hydrogen): histograms of the two bond lengths |r_H - r_O| and of the angle between the two bonds
// (base/vectors.py: angle_between_two_vectors = acos of the normalised dot product). One thread per object.
__global__ void __launch_bounds__(256)
bond_histogram_kernel(const Particle *particles, size_t n_objects, double length, int n_bins, double length_min,
                      double inv_length_width, double angle_min, double inv_angle_width, unsigned long long *length_histogram,
                      unsigned long long *angle_histogram) {
    const double half = 0.5 * length;
    for (size_t o = blockIdx.x * (size_t)blockDim.x + threadIdx.x; o < n_objects; o += (size_t)gridDim.x * blockDim.x) {
        const Particle h1 = particles[3 * o], ox = particles[3 * o + 1], h2 = particles[3 * o + 2];
        const double ax = correct_separation_in_box(h1.x - ox.x, length, half), ay = correct_separation_in_box(h1.y - ox.y, length, half),
                     az = correct_separation_in_box(h1.z - ox.z, length, half);
        const double bx = correct_separation_in_box(h2.x - ox.x, length, half), by = correct_separation_in_box(h2.y - ox.y, length, half),
                     bz = correct_separation_in_box(h2.z - ox.z, length, half);
        const double na = sqrt(fma(ax, ax, fma(ay, ay, az * az))), nb = sqrt(fma(bx, bx, fma(by, by, bz * bz)));
        const double values[3] = {na, nb, acos(fma(ax, bx, fma(ay, by, az * bz)) / (na * nb))};
        for (int k = 0; k < 3; k++) {
            const double scaled = k < 2 ? (values[k] - length_min) * inv_length_width : (values[k] - angle_min) * inv_angle_width;
            if (scaled >= 0.0 && scaled <= (double)n_bins)
                atomicAdd((k < 2 ? length_histogram : angle_histogram) + min((int)scaled, n_bins - 1), 1ull);
        }
    }
}

// ---------------------------------------------------------------------------------------------------------
// batched potential arithmetic (ecmc_potential_derivative / ecmc_potential_displacement): one warp per element
// for the merged-image Coulomb sum, one thread per element otherwise.
// ---------------------------------------------------------------------------------------------------------
struct BatchArgs {
    int dimension, dir;
    double speed, length;
    double velocity[3];
    size_t n;
    const double *separations, *charges, *potential_changes;
    double *out;
};

__global__ void __launch_bounds__(256)
derivative_batch_kernel(const __grid_constant__ PotentialParams p, const BatchArgs b) {
    __shared__ double trig_all[8 * kTrigDoubles];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    double *trig = trig_all + warp * kTrigDoubles;
    if (p.kind == ECMC_POT_MERGED_IMAGE_COULOMB) {
        for (size_t i = blockIdx.x * 8ull + warp; i < b.n; i += (size_t)gridDim.x * 8ull) {
            const double *s = b.separations + i * b.dimension;
            const double c1 = b.charges ? b.charges[2 * i] : 1.0, c2 = b.charges ? b.charges[2 * i + 1] : 1.0;
            const double v = derivative_warp<-1>(p, b.dir, b.speed, s[0], s[1], b.dimension > 2 ? s[2] : 0.0, c1, c2, trig, lane);
            if (lane == 0) b.out[i] = v;
        }
    } else {
        for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < b.n; i += (size_t)gridDim.x * blockDim.x) {
            const double *s = b.separations + i * b.dimension;
            const double c1 = b.charges ? b.charges[2 * i] : 1.0, c2 = b.charges ? b.charges[2 * i + 1] : 1.0;
            b.out[i] = derivative_warp<-1>(p, b.dir, b.speed, s[0], b.dimension > 1 ? s[1] : 0.0,
                                           b.dimension > 2 ? s[2] : 0.0, c1, c2, trig, lane);
        }
    }
}

__global__ void __launch_bounds__(256)
displacement_batch_kernel(const __grid_constant__ PotentialParams p, const BatchArgs b) {
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < b.n; i += (size_t)gridDim.x * blockDim.x) {
        const double *s = b.separations + i * b.dimension;
        const double sx = s[0], sy = b.dimension > 1 ? s[1] : 0.0, sz = b.dimension > 2 ? s[2] : 0.0;
        const double c1 = b.charges ? b.charges[2 * i] : 1.0, c2 = b.charges ? b.charges[2 * i + 1] : 1.0;
        const double du = b.potential_changes ? b.potential_changes[i] : 0.0;
        double out;
        if (p.kind == ECMC_POT_HARD_SPHERE || p.kind == ECMC_POT_HARD_DIPOLE) {
            // general velocity (hard_sphere_potential.py:65-99, hard_dipole_potential.py:75-114)
            const double vx = b.velocity[0], vy = b.velocity[1], vz = b.velocity[2];
            const double vv = dot3(vx, vy, vz, vx, vy, vz);
            const double vs = dot3(vx, vy, vz, sx, sy, sz);
            const double ss = dot3(sx, sy, sz, sx, sy, sz);
            out = p.kind == ECMC_POT_HARD_SPHERE ? hard_sphere_time(p.p0, vv, vs, ss) : hard_dipole_time(p.p0, p.p1, vv, vs, ss);
        } else {
            out = displacement_time<-1>(p, b.dir, 1.0 / b.speed, b.length, sx, sy, sz, c1, c2, du);
        }
        b.out[i] = out;
    }
}

}  // namespace ecmc
