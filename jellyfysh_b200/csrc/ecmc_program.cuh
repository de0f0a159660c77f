// ecmc_program.cuh -- the device-side description of one ECMC configuration ("program") and of the state
// of the chains in HBM. Plain data, passed to the kernels by value (__grid_constant__).
//
// HBM layout (DESIGN.md "Data layout"):
//   particles  [n_chains][n_particles]   32-byte records (x, y, z, charge): a gather of one target particle
//                                        costs exactly one 32-byte DRAM sector and one vector load
//   occupants  [n_chains][n_cells][max_occupants] int32, -1 = empty (SingleActiveCellOccupancy._occupied_cells)
//   surplus    [n_chains][max_surplus] int32 + n_surplus[n_chains]   (SingleActiveCellOccupancy._surplus)
//   chains     [n_chains] EcmcChainState (include/ecmc.h), read once / written once per launch
// Tables shared by all chains (cell geometry, nearby offsets, Walker tables, Ewald term lists) are read-only.
#pragma once

#include "ecmc_math.cuh"

namespace ecmc {

struct __align__(32) Particle {
    double x, y, z, charge;
};

// Walker alias-table entry (jellyfysh/event_handler/walker.py:69-103): cell_a if uniform(0, mean) <= rate_a, else
// cell_b. One 32-byte record per entry = one sector per cell-veto draw: the relative cells come as packed per-axis
// identifiers (x | y << 10 | z << 20, no divisions on the device) together with the derivative bound of each cell
// for this direction and sign (CellVetoEventHandler._derivative_bounds), which the confirmation step needs.
struct __align__(32) WalkerEntry {
    double rate_a;
    int cell_a, cell_b;
    double bound_a, bound_b;
};

struct DeviceWalker {
    const WalkerEntry *entries;
    int n_entries;
    int bits;  // n_entries.bit_length(), for CPython's _randbelow
    double total_rate, mean_rate;
    double inv_total_rate_speed;  // 1 / (total_rate * speed)
};

struct PotentialParams {
    int kind;  // EcmcPotentialKind
    LennardJones lj;
    InversePower ip;
    DisplacedEvenPower dep;
    double p0, p1;  // hard sphere radius | hard dipole min, max | Coulomb bound prefactor
    MergedImageCoulomb mic;
};

constexpr int kMaxFourierCutoff = 15;
constexpr int kTrigDoubles = 3 * 2 * (kMaxFourierCutoff + 1);
// the same for kernels that also evaluate three sums side by side (mic_derivative_warp3: 3 x 3 x 2 x (cutoff + 1) doubles,
// up to cutoff 6 -- the shipped potentials; larger cutoffs take the sums one at a time)
constexpr int kTrigDoubles3 = 128;
static_assert(kTrigDoubles3 >= kTrigDoubles, "one sum with the largest Fourier cutoff must fit");
constexpr int kMaxNearby = 343;  // (2 * 3 + 1)^3

struct DeviceProgram {
    int dimension, n_particles, n_cells, n_nearby;
    int per_side[3], cumulative[3];
    int max_occupants, max_surplus;
    int pair_handler, pair_use_charge;
    int veto_enabled, veto_use_charge;
    uint32_t seed;
    int no_cells;  // EcmcProgram.no_cells: every other unit is a candidate, no cell-boundary events
    double length, half_length, beta, speed, chain_time, veto_target_charge;
    double side_length[3];
    // cand: the invertible potential the pair candidates are drawn from (the pair potential of a
    // TwoLeafUnitEventHandler, the bounding potential of a TwoLeafUnitBoundingPotentialEventHandler);
    // real: the potential a bounded pair event is confirmed against; veto: the cell-veto potential
    PotentialParams cand_potential, real_potential, veto_potential;
    // tables
    const int *nearby;            // [n_nearby] relative cell identifiers packed x | y << 10 | z << 20 (cuboid_periodic_cells.py:74-100)
    const double *cell_min_axis;  // [3][max_per_side] lower cell boundary per axis index (cuboid_cells.py:119-131)
    const int *translate_axis;    // [3][max_per_side][max_per_side] (cuboid_periodic_cells.py:182-207)
    int max_per_side;
    int translate_modular;        // 1 if translate_axis[d][a][r] == (a + r) mod cells_per_side[d] everywhere
    DeviceWalker upper[3], lower[3];
    const double *bounds;         // [n_cells][dimension][2] (upper, -lower) per relative cell: ECMC_FAR_CELL_BOUNDING only
    int neighbor_layers, pad3;
    double inv_beta, inv_speed;
    // composite point objects: leaves grouped into roots (EcmcProgram.nodes_per_root), intramolecular pair factors
    int nodes_per_root, n_bonds;
    int bonds[ECMC_MAX_BONDS][2];
    double root_speed;            // speed * weight, weight = 1 / nodes_per_root
    PotentialParams bond_potential;
};

struct DeviceState {
    Particle *particles;
    Particle *roots;   // [n_chains][n_particles / nodes_per_root] root-unit positions (composite objects only)
    int *occupants;
    int *surplus;
    int *n_surplus;
    EcmcChainState *chains;
    int n_chains;      // chains of this launch ...
    int first_chain;   // ... starting at this one (launches on chain slices overlap copies with compute)
};

struct RunArgs {
    double until_q, until_r;
    long long max_events;   // <= 0: unlimited
    EcmcEventRecord *records;
    int records_per_chain;
    EcmcStats *stats;
    int list_capacity;      // lj_spec_kernel: entries of the per-chain candidate list in shared memory
    // lj_spec_kernel<..., HOST = true> (a whole host step in one launch, ecmc_submit_from_host_sparse): the chain's warp
    // reads its start configuration from the caller's pinned buffer (device-visible mapping), bins it into the cells,
    // starts the run, and writes every position it changes through to the same kind of buffer
    const double *host_in;  // [n_chains][n_particles][3]
    double *host_out;
    uint32_t first_stream;
    int initial_active, initial_direction;
    unsigned long long *host_writes;  // counts the particles written to host_out
    int keep_state;         // the step continues the chains (start_chain)
};

}  // namespace ecmc
