// ecmc_engine.cu -- libecmc_b200.so: the C ABI of include/ecmc.h over the sm_100a kernels of ecmc_kernels.cuh.
// Host side: turns an EcmcProgram into the device program (cell geometry, nearby-cell list, translate tables,
// Walker tables, Ewald term lists), owns the HBM state of the chains and launches the kernels on one stream.
// There is no CPU implementation of the hot path in this library.
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <new>
#include <string>
#include <vector>

#include "ecmc_kernels.cuh"
#include "ecmc_molecules.cuh"
#include "ecmc_spec.cuh"
#include "ecmc_spec_cta.cuh"
#include "ecmc_disks.cuh"

using namespace ecmc;

#define ECMC_API extern "C" __attribute__((visibility("default")))

namespace {

thread_local std::string g_create_error;

// Chains (warps) per CTA; must divide ECMC_RESIDENT_WARPS (28). Measured on C2 (profiles/README.md): 1 / 2 / 4 / 7 / 14
// warps -> 8.07 / 7.96 / 7.92 / 7.97 / 8.23e8 events/s; 28 does not fit the static shared memory (2 KB per warp).
#ifndef ECMC_WARPS_PER_BLOCK
#define ECMC_WARPS_PER_BLOCK 14
#endif
constexpr int kWarpsPerBlock = ECMC_WARPS_PER_BLOCK;
static_assert(ECMC_RESIDENT_WARPS % ECMC_WARPS_PER_BLOCK == 0, "whole CTAs must fill the resident warps of an SM");
// molecule_kernel without the per-event CTA barrier (composite objects other than water, ECMC_MOLECULE_ALIGNED=0): small
// CTAs leave the compiler its 168 registers (a CTA of 14 warps would cap them at 128)
constexpr int kMoleculeWarps = 4;
// molecule_kernel with the per-event CTA barrier (the shipped water potentials): warps per CTA
#ifndef ECMC_ALIGNED_WARPS
#define ECMC_ALIGNED_WARPS 8
#endif
constexpr int kAlignedWarps = ECMC_ALIGNED_WARPS;
// More chains than kAlignedWarps per SM: CTAs of sixteen aligned warps. The kernel's time is the latency of each chain's
// dependent instructions, so warps per SM are throughput -- and ONE large CTA per SM, whose warps fetch the event's
// instructions together, beats several small ones (measured on B200, 32 water molecules per chain, one wave of chains
// each: 8 warps per SM 4.5e7, 12: 6.2e7, 16: 7.0e7 events/s at 128 registers and 200 B of spills, 20: 7.2e7 at 96
// registers and 658 B; 2 CTAs x 8 warps 5.2e7, 4 x 4: 4.0e7, 3 x 6: 3.4e7).
constexpr int kWideWarps = 16;

struct EventPair {
    cudaEvent_t start, stop;
};

}  // namespace

struct EcmcHandle {
    int device = 0;
    int n_chains = 0;
    EcmcProgram program{};
    DeviceProgram dprog{};
    DeviceState state{};
    cudaStream_t stream = nullptr;
    std::vector<cudaStream_t> slice_streams;  // ecmc_run_from_host pipelines chain slices over these
    std::vector<void *> allocations;
    EcmcStats *d_stats = nullptr;
    EcmcStats *h_stats = nullptr;     // pinned
    bool roots_uploaded = false;
    bool molecules = false;           // composite objects in root-level cells: molecule_kernel
    MoleculeProgram mprog;
    bool disks = false;               // general velocities (EcmcProgram.eoc_sequential): disk_kernel
    DiskProgram kprog{};
    double *d_staging = nullptr;      // [n_chains][n_particles][dimension] + charges
    double *d_staging_charges = nullptr;
    uint32_t *d_streams = nullptr;
    std::vector<EventPair> timed;     // launches not yet accounted
    std::vector<EventPair> free_events;
    double kernel_seconds = 0.0;
    uint64_t kernel_launches = 0;
    bool started = false;
    bool spec = true;        // Lennard-Jones / cell-veto programs: lj_spec_kernel (ecmc_set_option)
    bool spec_prune = true;  // ... with the force-bound pruning of pair candidates in ecmc_run / ecmc_run_from_host
    int spec_lanes = 4;      // lanes per speculated event (4: 8 events per batch, 8: 4 events per batch)
    bool chain_blocks = true; // few chains: lj_chain_kernel, one CTA of four warps per chain (ecmc_spec_cta.cuh)
    bool host_fused = true;   // sparse host steps of Lennard-Jones / cell-veto programs as ONE launch per chain slice
    bool host_continue = false; // host steps continue the chains instead of starting a new run each (ecmc_set_option)
    bool spec_coulomb = true;   // Coulomb atoms (bound / merged-image Coulomb / cell veto): the batched kernel too
    int coulomb_warps = 0;      // ... with this many chains per CTA (7, 14 or 28; 0: the largest that fits)
    std::string kernel_name; // ecmc_kernel_name
    bool slices_busy = false; // ecmc_submit_from_host work in flight on the slice streams (until ecmc_wait)
    int slices_layout = 0;    // ... and how its chains were cut into slices (steps are ordered slice by slice)
    unsigned long long *d_changed = nullptr;  // particles written back by ecmc_submit_from_host_sparse since the last wait
    uint64_t host_bytes_written = 0;          // ... as bytes, accumulated by ecmc_wait
    std::string error;
};

namespace {

// Which molecule_kernel a handle runs: the warps per CTA of the aligned kernel for the shipped water potentials (Coulomb
// bound / merged-image Coulomb / harmonic bonds / Lennard-Jones; kAlignedWarps, or kWideWarps when there are more chains
// than that per SM), kMoleculeWarps for the generic instantiation, -1 for the water potentials without the CTA barrier
// (ECMC_MOLECULE_ALIGNED=0).
int molecule_cta_warps(const EcmcHandle *h) {
    const DeviceProgram &d = h->dprog;
    const bool water = d.cand_potential.kind == ECMC_POT_INVERSE_POWER_COULOMB_BOUNDING &&
                       d.real_potential.kind == ECMC_POT_MERGED_IMAGE_COULOMB &&
                       (!d.veto_enabled || d.veto_potential.kind == ECMC_POT_MERGED_IMAGE_COULOMB) &&
                       (d.n_bonds == 0 || (d.bond_potential.kind == ECMC_POT_DISPLACED_EVEN_POWER &&
                                           d.bond_potential.dep.power == 2.0)) &&
                       (h->mprog.n_inter == 0 || h->mprog.inter_potential.kind == ECMC_POT_LENNARD_JONES);
    if (!water || h->mprog.root_mode || h->mprog.cell_child) return kMoleculeWarps;
    if (const char *env = std::getenv("ECMC_MOLECULE_ALIGNED"))
        if (std::atoi(env) == 0) return -1;
    int sm_count = 148;
    if (cudaDeviceGetAttribute(&sm_count, cudaDevAttrMultiProcessorCount, h->device) != cudaSuccess) sm_count = 148;
    return h->n_chains > kAlignedWarps * sm_count ? kWideWarps : kAlignedWarps;
}

int fail(EcmcHandle *h, int code, const std::string &message) {
    if (h) h->error = message; else g_create_error = message;
    return code;
}
#define CUDA_TRY(h, expr)                                                                                      \
    do {                                                                                                       \
        cudaError_t err_ = (expr);                                                                             \
        if (err_ != cudaSuccess)                                                                               \
            return fail(h, ECMC_ERR_CUDA, std::string(#expr) + ": " + cudaGetErrorString(err_));               \
    } while (0)

template <typename T>
int device_alloc(EcmcHandle *h, T **out, size_t count) {
    void *p = nullptr;
    CUDA_TRY(h, cudaMalloc(&p, std::max<size_t>(count, 1) * sizeof(T)));
    h->allocations.push_back(p);
    *out = static_cast<T *>(p);
    return ECMC_OK;
}
template <typename T>
int device_upload(EcmcHandle *h, const T **out, const std::vector<T> &host) {
    T *p = nullptr;
    int rc = device_alloc(h, &p, host.size());
    if (rc) return rc;
    if (!host.empty()) CUDA_TRY(h, cudaMemcpy(p, host.data(), host.size() * sizeof(T), cudaMemcpyHostToDevice));
    *out = p;
    return ECMC_OK;
}

// Python's float % for a positive divisor (setting/hypercubic_setting.py:117)
double host_py_mod(double x, double L) {
    double m = std::fmod(x, L);
    if (m != 0.0) { if (m < 0.0) m += L; } else m = 0.0;
    return m;
}

// CuboidCells geometry per axis (cuboid_cells.py:98-146): cell `id` covers exactly the doubles x with
// int(x / side) == id; cell_min / cell_max are the smallest / largest of them.
void axis_geometry(int n, double side, std::vector<double> &cell_min, std::vector<double> &cell_max) {
    cell_min.assign(n, 0.0);
    cell_max.assign(n, 0.0);
    const double inf = INFINITY;
    for (int id = 0; id < n; id++) {
        double lo = id * side, hi = (id + 1) * side;
        if (lo > 0.0) {
            while ((int)(lo / side) >= id) lo = std::nextafter(lo, -inf);
            while ((int)(lo / side) < id) lo = std::nextafter(lo, inf);
        }
        while ((int)(hi / side) <= id) hi = std::nextafter(hi, inf);
        while ((int)(hi / side) > id) hi = std::nextafter(hi, -inf);
        cell_min[id] = lo;
        cell_max[id] = hi;
    }
}

int bit_length(uint32_t n) {
    int bits = 0;
    while ((n >> bits) != 0) bits++;
    return bits;
}

int make_potential(EcmcHandle *h, const EcmcPotential &in, double L, PotentialParams *out) {
    PotentialParams p;
    std::memset(&p, 0, sizeof(p));
    p.kind = in.kind;
    switch (in.kind) {
    case ECMC_POT_NONE: break;
    case ECMC_POT_INVERSE_POWER:
        p.ip = make_inverse_power(in.params[0], in.params[1]);
        if (!(p.ip.power > 0.0)) return fail(h, ECMC_ERR_INVALID, "inverse power potential: power must be > 0");
        break;
    case ECMC_POT_LENNARD_JONES:
        if (!(in.params[0] > 0.0) || !(in.params[1] > 0.0))
            return fail(h, ECMC_ERR_INVALID, "Lennard-Jones potential: prefactor and characteristic_length must be > 0");
        p.lj = make_lennard_jones(in.params[0], in.params[1]);
        break;
    case ECMC_POT_DISPLACED_EVEN_POWER:
        p.dep.k = in.params[0];
        p.dep.r0 = in.params[1];
        p.dep.power = in.params[2];
        if (!(p.dep.k > 0.0) || !(p.dep.power > 0.0) || std::fmod(p.dep.power, 2.0) != 0.0)
            return fail(h, ECMC_ERR_INVALID, "displaced even power potential: prefactor > 0 and an even power required");
        break;
    case ECMC_POT_HARD_SPHERE: p.p0 = in.params[0]; break;
    case ECMC_POT_HARD_DIPOLE: p.p0 = in.params[0]; p.p1 = in.params[1]; break;
    case ECMC_POT_INVERSE_POWER_COULOMB_BOUNDING: p.p0 = in.params[0]; break;
    case ECMC_POT_MERGED_IMAGE_COULOMB: {
        // constants and Fourier coefficients of merged_image_coulomb_potential.c:77-119; the loops of
        // derivative() (:205-274) become flat term lists
        const double prefactor = in.params[0], alpha = in.params[1];
        const int fc = (int)in.params[2], pc = (int)in.params[3];
        if (fc < 0 || fc > kMaxFourierCutoff || pc < 0 || pc > 100 || !(alpha > 0.0))
            return fail(h, ECMC_ERR_INVALID, "merged image Coulomb potential: cutoffs / alpha out of range");
        p.mic.prefactor = prefactor;
        p.mic.alpha_over_length = alpha / L;
        p.mic.alpha_over_length_sq = alpha * alpha / (L * L);
        p.mic.two_alpha_root_pi = 2.0 * alpha / (L * std::sqrt(M_PI));
        p.mic.length = L;
        p.mic.two_pi_over_length = 2.0 * M_PI / L;
        p.mic.fourier_cutoff = fc;
        std::vector<int> images, modes;
        std::vector<double> coefficients;
        for (int k = -pc; k <= pc; k++) {
            const int cy = (int)std::sqrt((double)(pc * pc - k * k));
            for (int j = -cy; j <= cy; j++) {
                const int cx = (int)std::sqrt((double)(pc * pc - j * j - k * k));
                for (int i = -cx; i <= cx; i++) images.push_back((i & 0xff) | ((j & 0xff) << 8) | ((k & 0xff) << 16));
            }
        }
        for (int i = 1; i <= fc; i++) {
            const int cy = (int)std::sqrt((double)(fc * fc - i * i));
            for (int j = 0; j <= cy; j++) {
                const int cz = (int)std::sqrt((double)(fc * fc - i * i - j * j));
                for (int k = 0; k <= cz; k++) {
                    const double multiplicity = (j == 0 && k == 0) ? 1.0 : ((j == 0 || k == 0) ? 2.0 : 4.0);
                    const double norm_sq = (double)(i * i + j * j + k * k);
                    modes.push_back(i | (j << 8) | (k << 16));
                    coefficients.push_back(4.0 * i * multiplicity / (norm_sq * L * L) * std::exp(-M_PI * M_PI * norm_sq / (alpha * alpha)));
                }
            }
        }
        p.mic.n_images = (int)images.size();
        p.mic.n_modes = (int)modes.size();
        int rc = device_upload(h, &p.mic.images, images);
        if (!rc) rc = device_upload(h, &p.mic.modes, modes);
        if (!rc) rc = device_upload(h, &p.mic.coefficients, coefficients);
        if (rc) return rc;
        break;
    }
    default: return fail(h, ECMC_ERR_INVALID, "unknown potential kind " + std::to_string(in.kind));
    }
    *out = p;
    return ECMC_OK;
}

bool is_invertible(int kind) {
    return kind == ECMC_POT_INVERSE_POWER || kind == ECMC_POT_LENNARD_JONES || kind == ECMC_POT_DISPLACED_EVEN_POWER ||
           kind == ECMC_POT_HARD_SPHERE || kind == ECMC_POT_HARD_DIPOLE || kind == ECMC_POT_INVERSE_POWER_COULOMB_BOUNDING;
}
bool has_derivative(int kind) {
    return kind == ECMC_POT_INVERSE_POWER || kind == ECMC_POT_LENNARD_JONES || kind == ECMC_POT_DISPLACED_EVEN_POWER ||
           kind == ECMC_POT_MERGED_IMAGE_COULOMB || kind == ECMC_POT_INVERSE_POWER_COULOMB_BOUNDING;
}

int upload_walker(EcmcHandle *h, const EcmcWalkerTable &in, DeviceWalker *out, const DeviceProgram &d, const double *bounds,
                  int direction, int rate_index) {
    std::memset(out, 0, sizeof(*out));
    if (in.n_entries <= 0) return ECMC_OK;
    if (!in.cell_a || !in.cell_b || !in.rate_a) return fail(h, ECMC_ERR_INVALID, "Walker table with null arrays");
    std::vector<WalkerEntry> entries(in.n_entries);
    auto pack = [&d](int cell) {
        return ((cell / d.cumulative[0]) % d.per_side[0]) | (((cell / d.cumulative[1]) % d.per_side[1]) << 10) |
               (((cell / d.cumulative[2]) % d.per_side[2]) << 20);
    };
    auto bound = [&](int cell) { return bounds[((size_t)cell * d.dimension + direction) * 2 + rate_index]; };
    for (int e = 0; e < in.n_entries; e++) {
        if (in.cell_a[e] < 0 || in.cell_a[e] >= d.n_cells || in.cell_b[e] >= d.n_cells)
            return fail(h, ECMC_ERR_INVALID, "Walker table entry refers to a cell outside the cell system");
        entries[e].rate_a = in.rate_a[e];
        entries[e].cell_a = pack(in.cell_a[e]);
        entries[e].bound_a = bound(in.cell_a[e]);
        entries[e].cell_b = in.cell_b[e] >= 0 ? pack(in.cell_b[e]) : -1;
        entries[e].bound_b = in.cell_b[e] >= 0 ? bound(in.cell_b[e]) : 0.0;
    }
    out->n_entries = in.n_entries;
    out->bits = bit_length((uint32_t)in.n_entries);
    out->total_rate = in.total_rate;
    out->mean_rate = in.mean_rate;
    out->inv_total_rate_speed = 1.0 / (in.total_rate * d.speed);
    return device_upload(h, &out->entries, entries);
}

int build_device_program(EcmcHandle *h) {
    const EcmcProgram &p = h->program;
    DeviceProgram &d = h->dprog;
    std::memset(&d, 0, sizeof(d));
    if (p.abi_version != ECMC_ABI_VERSION) return fail(h, ECMC_ERR_INVALID, "ABI version mismatch");
    if (p.dimension < 1 || p.dimension > 3) return fail(h, ECMC_ERR_INVALID, "dimension must be 1, 2 or 3");
    if (p.n_particles < 1 || p.n_particles >= (1 << 24))
        return fail(h, ECMC_ERR_INVALID, "n_particles must be in [1, 2^24)");
    if (!(p.system_length > 0.0) || !(p.beta > 0.0) || !(p.speed > 0.0) || !(p.chain_time > 0.0))
        return fail(h, ECMC_ERR_INVALID, "system_length, beta, speed and chain_time must be > 0");
    if (p.max_occupants < 1 || p.max_surplus < 0 || p.neighbor_layers < 0 || p.neighbor_layers > 3)
        return fail(h, ECMC_ERR_INVALID, "max_occupants >= 1, max_surplus >= 0, 0 <= neighbor_layers <= 3 required");
    if (p.initial_active < 0 || p.initial_active >= p.n_particles || p.initial_direction < 0 ||
        p.initial_direction >= p.dimension)
        return fail(h, ECMC_ERR_INVALID, "initial_active / initial_direction out of range");
    if (p.no_cells) {
        for (int k = 0; k < p.dimension; k++)
            if (p.cells_per_side[k] != 1) return fail(h, ECMC_ERR_INVALID, "no_cells needs cells_per_side = 1");
        const int units = p.cell_level == 1 && p.nodes_per_root > 1 ? p.n_particles / p.nodes_per_root : p.n_particles;
        if (p.neighbor_layers != 0 || p.max_occupants != 1 || p.max_surplus < units - 1 || p.veto_enabled != ECMC_FAR_NONE)
            return fail(h, ECMC_ERR_INVALID, "no_cells needs neighbor_layers = 0, max_occupants = 1, max_surplus >= units - 1 "
                                             "and no far field");
    }
    d.no_cells = p.no_cells ? 1 : 0;
    d.dimension = p.dimension;
    d.n_particles = p.n_particles;
    d.n_cells = 1;
    d.max_per_side = 1;
    for (int k = 0; k < 3; k++) {
        const int n = k < p.dimension ? p.cells_per_side[k] : 1;
        if (n < 1 || n > 1023) return fail(h, ECMC_ERR_INVALID, "cells_per_side must be in [1, 1023]");
        d.per_side[k] = n;
        d.cumulative[k] = d.n_cells;
        d.n_cells *= n;
        d.side_length[k] = p.system_length / n;
        d.max_per_side = std::max(d.max_per_side, n);
    }
    d.max_occupants = p.max_occupants;
    d.max_surplus = std::max(p.max_surplus, 1);
    d.pair_handler = p.pair_handler;
    d.pair_use_charge = p.pair_use_charge;
    d.veto_enabled = p.veto_enabled;
    d.veto_use_charge = p.veto_use_charge;
    d.seed = p.seed;
    d.length = p.system_length;
    d.half_length = p.system_length / 2.0;
    d.beta = p.beta;
    d.speed = p.speed;
    d.chain_time = p.chain_time;
    d.veto_target_charge = p.veto_target_charge;
    d.inv_beta = 1.0 / p.beta;
    d.inv_speed = 1.0 / p.speed;
    // composite point objects
    d.nodes_per_root = p.nodes_per_root > 1 ? p.nodes_per_root : 1;
    if (p.n_particles % d.nodes_per_root) return fail(h, ECMC_ERR_INVALID, "n_particles must be a multiple of nodes_per_root");
    if (p.n_bonds < 0 || p.n_bonds > ECMC_MAX_BONDS || (p.n_bonds > 0 && d.nodes_per_root == 1))
        return fail(h, ECMC_ERR_INVALID, "bonds need composite objects and n_bonds <= ECMC_MAX_BONDS");
    d.n_bonds = p.n_bonds;
    for (int b = 0; b < p.n_bonds; b++)
        for (int k = 0; k < 2; k++) {
            if (p.bonds[b][k] < 0 || p.bonds[b][k] >= d.nodes_per_root || p.bonds[b][0] == p.bonds[b][1])
                return fail(h, ECMC_ERR_INVALID, "bond child indices out of range");
            d.bonds[b][k] = p.bonds[b][k];
        }
    d.root_speed = p.speed * (1.0 / d.nodes_per_root);  // velocity component * weight (abstracts.py:181)
    // molecules: composite objects in root-level cells (water)
    h->molecules = p.cell_level == 1 && d.nodes_per_root > 1;
    std::memset(&h->mprog, 0, sizeof(h->mprog));
    if ((p.pair_handler == ECMC_PAIR_TWO_COMPOSITE_SUMMED_BOUNDING || p.bending_enabled ||
         (p.n_inter_factors > 0 && !p.eoc_sequential)) && !h->molecules)
        return fail(h, ECMC_ERR_INVALID, "composite-object handlers need cell_level = 1 and nodes_per_root > 1");
    // general velocities: two-dimensional composite point objects without a cell system, hard potentials only
    h->disks = p.eoc_sequential != 0;
    std::memset(&h->kprog, 0, sizeof(h->kprog));
    if (h->disks) {
        DiskProgram &k = h->kprog;
        if (p.dimension != 2 || !p.no_cells || d.nodes_per_root < 2 || h->molecules || p.pair_handler != ECMC_PAIR_NONE ||
            p.veto_enabled != ECMC_FAR_NONE)
            return fail(h, ECMC_ERR_INVALID, "general velocities need two dimensions, composite point objects, no cell system "
                                             "and factor-type-map pair factors only");
        if (p.n_inter_factors < 0 || p.n_inter_factors > ECMC_MAX_INTER_FACTORS)
            return fail(h, ECMC_ERR_INVALID, "n_inter_factors out of range");
        auto hard = [](int kind) { return kind == ECMC_POT_HARD_SPHERE || kind == ECMC_POT_HARD_DIPOLE; };
        if ((p.n_inter_factors > 0 && !hard(p.inter_potential.kind)) || (p.n_bonds > 0 && !hard(p.bond_potential.kind)))
            return fail(h, ECMC_ERR_INVALID, "general velocities need hard potentials");
        k.n_inter = p.n_inter_factors;
        for (int f = 0; f < p.n_inter_factors; f++)
            for (int j = 0; j < 2; j++) {
                if (p.inter_factors[f][j] < 0 || p.inter_factors[f][j] >= d.nodes_per_root)
                    return fail(h, ECMC_ERR_INVALID, "inter-object factor child index out of range");
                k.inter[f][j] = p.inter_factors[f][j];
            }
        k.inter_kind = p.inter_potential.kind; k.inter_p0 = p.inter_potential.params[0]; k.inter_p1 = p.inter_potential.params[1];
        k.bond_kind = p.bond_potential.kind; k.bond_p0 = p.bond_potential.params[0]; k.bond_p1 = p.bond_potential.params[1];
        k.eoc_cos = p.eoc_cos; k.eoc_sin = p.eoc_sin;
        k.weight = 1.0 / d.nodes_per_root;
        k.initial_direction = p.initial_direction;
    }
    if (h->molecules) {
        MoleculeProgram &m = h->mprog;
        if (p.dimension != 3 || d.nodes_per_root > 3 || p.max_occupants != 1)
            return fail(h, ECMC_ERR_INVALID, "molecules need dimension 3, nodes_per_root <= 3 and max_occupants = 1");
        // ECMC_PAIR_TWO_LEAF_UNIT_BOUNDING here: one bounded handler per pair of leaves of different objects
        if (p.pair_handler != ECMC_PAIR_NONE && p.pair_handler != ECMC_PAIR_TWO_COMPOSITE_SUMMED_BOUNDING &&
            p.pair_handler != ECMC_PAIR_TWO_LEAF_UNIT_BOUNDING)
            return fail(h, ECMC_ERR_INVALID, "molecules need the composite-object pair handler or bounded leaf-to-leaf factors");
        // ECMC_FAR_CELL_BOUNDING here: TwoCompositeObjectCellBoundingPotentialEventHandler, one candidate per object in a
        // cell that is not nearby
        if (p.composite_lifting < ECMC_LIFTING_INSIDE_FIRST || p.composite_lifting > ECMC_LIFTING_RATIO)
            return fail(h, ECMC_ERR_INVALID, "unknown lifting scheme");
        if (p.n_inter_factors < 0 || p.n_inter_factors > ECMC_MAX_INTER_FACTORS)
            return fail(h, ECMC_ERR_INVALID, "n_inter_factors out of range");
        m.composite_lifting = p.composite_lifting;
        m.n_inter = p.n_inter_factors;
        for (int f = 0; f < p.n_inter_factors; f++)
            for (int k = 0; k < 2; k++) {
                if (p.inter_factors[f][k] < 0 || p.inter_factors[f][k] >= d.nodes_per_root)
                    return fail(h, ECMC_ERR_INVALID, "inter-object factor child index out of range");
                m.inter[f][k] = p.inter_factors[f][k];
            }
        if (p.n_inter_factors > 0) {
            if (!is_invertible(p.inter_potential.kind)) return fail(h, ECMC_ERR_INVALID, "inter-object potential is not invertible");
            int rc_inter = make_potential(h, p.inter_potential, p.system_length, &m.inter_potential);
            if (rc_inter) return rc_inter;
        }
        m.bending_enabled = p.bending_enabled ? 1 : 0;
        m.boundary_keeps_factors = p.boundary_keeps_factors ? 1 : 0;
        // a cell system for one kind of leaf (water/coulomb_power_bounded_lj_cell_bounded.ini)
        m.cell_child = p.cell_child;
        if (p.cell_child) {
            if (p.no_cells || p.cell_child < 1 || p.cell_child > d.nodes_per_root || p.root_mode ||
                p.pair_handler != ECMC_PAIR_TWO_COMPOSITE_SUMMED_BOUNDING || p.n_inter_factors != 1 ||
                p.inter_factors[0][0] != p.cell_child - 1 || p.inter_factors[0][1] != p.cell_child - 1 ||
                !p.boundary_keeps_factors || !(p.inter_bound_max_displacement > 0.0) ||
                (p.veto_enabled != ECMC_FAR_NONE && p.veto_enabled != ECMC_FAR_CELL_BOUNDING) || p.veto_use_charge ||
                !has_derivative(p.inter_potential.kind))
                return fail(h, ECMC_ERR_INVALID, "leaf cells need composite-object pair factors, one chargeless two-leaf factor "
                                                 "between the stored leaves, a cell system and boundary_keeps_factors");
            m.inter_bound_offset = p.inter_bound_offset;
            m.inter_bound_max_displacement = p.inter_bound_max_displacement;
        }
        // root-unit-active mode (dipoles/dipole_motion.ini)
        m.root_mode = p.root_mode ? 1 : 0;
        if (p.root_mode) {
            if (d.nodes_per_root != 2 || !p.no_cells || p.bending_enabled || p.veto_enabled != ECMC_FAR_NONE ||
                p.pair_handler != ECMC_PAIR_TWO_COMPOSITE_SUMMED_BOUNDING)
                return fail(h, ECMC_ERR_INVALID, "the root-unit-active mode needs objects of two leaves with the composite-object "
                                                 "pair handler and no cell system");
            if (!(p.switch_chain_length[0] > 0.0) || !(p.switch_chain_length[1] > 0.0))
                return fail(h, ECMC_ERR_INVALID, "switch_chain_length must be > 0");
            m.switch_length[0] = p.switch_chain_length[0];
            m.switch_length[1] = p.switch_chain_length[1];
        }
        if (p.bending_enabled) {
            if (p.bending_potential.kind != ECMC_POT_BENDING || d.nodes_per_root != 3 || !(p.bending_max_displacement > 0.0) ||
                p.bending_lifting < ECMC_LIFTING_INSIDE_FIRST || p.bending_lifting > ECMC_LIFTING_RATIO)
                return fail(h, ECMC_ERR_INVALID, "bending needs three leaves per object, a bending potential, a lifting scheme and max_displacement > 0");
            m.bending_lifting = p.bending_lifting;
            for (int i = 0; i < 3; i++) {
                if (p.bending_children[i] < 0 || p.bending_children[i] >= 3) return fail(h, ECMC_ERR_INVALID, "bending child out of range");
                m.bending_children[i] = p.bending_children[i];
            }
            for (int i = 0; i < 4; i++) {
                if (p.bending_separations[i] < 0 || p.bending_separations[i] >= 3) return fail(h, ECMC_ERR_INVALID, "bending separation index out of range");
                m.bending_separations[i] = p.bending_separations[i];
            }
            m.bending_prefactor = p.bending_potential.params[0];
            m.bending_angle = p.bending_potential.params[1];
            m.bending_offset = p.bending_offset;
            m.bending_max_displacement = p.bending_max_displacement;
        }
    }

    // potentials
    int rc;
    bool relative_modular = true;
    if (p.n_bonds > 0) {
        if (!is_invertible(p.bond_potential.kind)) return fail(h, ECMC_ERR_INVALID, "bond potential is not invertible");
        if ((rc = make_potential(h, p.bond_potential, p.system_length, &d.bond_potential))) return rc;
    }
    if (p.pair_handler == ECMC_PAIR_TWO_LEAF_UNIT) {
        if (!is_invertible(p.pair_potential.kind)) return fail(h, ECMC_ERR_INVALID, "pair potential is not invertible");
        if ((rc = make_potential(h, p.pair_potential, p.system_length, &d.cand_potential))) return rc;
    } else if (p.pair_handler == ECMC_PAIR_TWO_LEAF_UNIT_BOUNDING || p.pair_handler == ECMC_PAIR_TWO_COMPOSITE_SUMMED_BOUNDING) {
        if (!is_invertible(p.pair_bounding_potential.kind) || !has_derivative(p.pair_bounding_potential.kind))
            return fail(h, ECMC_ERR_INVALID, "pair bounding potential must be invertible with a derivative");
        if (!has_derivative(p.pair_potential.kind)) return fail(h, ECMC_ERR_INVALID, "pair potential has no derivative");
        if ((rc = make_potential(h, p.pair_bounding_potential, p.system_length, &d.cand_potential))) return rc;
        if ((rc = make_potential(h, p.pair_potential, p.system_length, &d.real_potential))) return rc;
    } else if (p.pair_handler != ECMC_PAIR_NONE) {
        return fail(h, ECMC_ERR_INVALID, "unknown pair handler kind");
    }
    if ((d.cand_potential.kind == ECMC_POT_MERGED_IMAGE_COULOMB || d.real_potential.kind == ECMC_POT_MERGED_IMAGE_COULOMB ||
         d.cand_potential.kind == ECMC_POT_INVERSE_POWER_COULOMB_BOUNDING) && p.dimension != 3)
        return fail(h, ECMC_ERR_INVALID, "Coulomb potentials need dimension 3");

    // nearby cells of cell zero, duplicates removed, in the order of cuboid_periodic_cells.py:74-100
    {
        std::vector<int> nearby;
        const int nl = p.neighbor_layers, w = 2 * nl + 1;
        int total = 1;
        for (int k = 0; k < p.dimension; k++) total *= w;
        for (int t = 0; t < total; t++) {
            int rem = t, id[3] = {0, 0, 0};
            for (int k = 0; k < p.dimension; k++) {
                const int off = rem % w - nl;
                rem /= w;
                id[k] = ((off % d.per_side[k]) + d.per_side[k]) % d.per_side[k];
            }
            const int code = id[0] | (id[1] << 10) | (id[2] << 20);
            bool dup = false;
            for (int c : nearby) dup = dup || c == code;
            if (!dup) nearby.push_back(code);
        }
        d.n_nearby = (int)nearby.size();
        if ((rc = device_upload(h, &d.nearby, nearby))) return rc;
    }
    // cell boundaries and the translate table, per axis
    {
        const int mps = d.max_per_side;
        std::vector<double> cmin_all(3 * mps, 0.0);
        std::vector<int> translate(3 * mps * mps, 0);
        d.translate_modular = 1;
        relative_modular = true;
        for (int k = 0; k < 3; k++) {
            std::vector<double> cmin, cmax;
            axis_geometry(d.per_side[k], d.side_length[k], cmin, cmax);
            for (int a = 0; a < d.per_side[k]; a++) {
                cmin_all[k * mps + a] = cmin[a];
                for (int r = 0; r < d.per_side[k]; r++) {
                    const double x = host_py_mod((cmax[a] + cmin[a]) / 2.0 + cmin[r], p.system_length);
                    translate[(k * mps + a) * mps + r] = (int)(x / d.side_length[k]);
                    if (translate[(k * mps + a) * mps + r] != (a + r) % d.per_side[k]) d.translate_modular = 0;
                    // relative_cell(cell a, reference r), cuboid_periodic_cells.py:155-180
                    const double y = host_py_mod((cmax[a] + cmin[a]) / 2.0 - cmin[r], p.system_length);
                    if ((int)(y / d.side_length[k]) != ((a - r) % d.per_side[k] + d.per_side[k]) % d.per_side[k])
                        relative_modular = false;
                }
            }
        }
        if ((rc = device_upload(h, &d.cell_min_axis, cmin_all))) return rc;
        if ((rc = device_upload(h, &d.translate_axis, translate))) return rc;
    }
    // far field: cell veto (Walker tables) or cell bounding (one candidate per occupied far cell)
    d.neighbor_layers = p.neighbor_layers;
    if (p.veto_enabled != ECMC_FAR_NONE && p.veto_enabled != ECMC_FAR_CELL_VETO && p.veto_enabled != ECMC_FAR_CELL_BOUNDING)
        return fail(h, ECMC_ERR_INVALID, "unknown far-field kind");
    if (p.veto_enabled) {
        if (!p.veto_tables || !p.veto_tables->bounds) return fail(h, ECMC_ERR_INVALID, "far field enabled without tables");
        if (!has_derivative(p.veto_potential.kind)) return fail(h, ECMC_ERR_INVALID, "far-field potential has no derivative");
        if (p.veto_potential.kind == ECMC_POT_MERGED_IMAGE_COULOMB && p.dimension != 3)
            return fail(h, ECMC_ERR_INVALID, "Coulomb potentials need dimension 3");
        if (p.veto_use_charge && !(p.veto_target_charge != 0.0))
            return fail(h, ECMC_ERR_INVALID, "veto_target_charge must not be 0");
        if ((rc = make_potential(h, p.veto_potential, p.system_length, &d.veto_potential))) return rc;
    }
    if (p.veto_enabled == ECMC_FAR_CELL_VETO) {
        for (int k = 0; k < p.dimension; k++) {
            if ((rc = upload_walker(h, p.veto_tables->upper[k], &d.upper[k], d, p.veto_tables->bounds, k, 0))) return rc;
            if ((rc = upload_walker(h, p.veto_tables->lower[k], &d.lower[k], d, p.veto_tables->bounds, k, 1))) return rc;
            if (d.upper[k].n_entries <= 0 && d.lower[k].n_entries <= 0)
                return fail(h, ECMC_ERR_INVALID, "veto enabled but a direction has no Walker table");
        }
    } else if (p.veto_enabled == ECMC_FAR_CELL_BOUNDING) {
        // the reference's handler takes exactly two units (two_leaf_unit_cell_bounding_potential_event_handler.py:158)
        if (p.max_occupants != 1) return fail(h, ECMC_ERR_INVALID, "cell bounding needs max_occupants == 1");
        if (!relative_modular) return fail(h, ECMC_ERR_INVALID, "cell geometry with non-modular relative cells");
        std::vector<double> bounds(p.veto_tables->bounds, p.veto_tables->bounds + (size_t)d.n_cells * p.dimension * 2);
        if ((rc = device_upload(h, &d.bounds, bounds))) return rc;
    }
    return ECMC_OK;
}

// ---- kernel dispatch: specialised instantiations for the named configurations, one generic fallback ----------
typedef void (*EventKernel)(const DeviceProgram, const DeviceState, const RunArgs);

template <int CAND, int REAL, int VETO>
EventKernel pick_record(bool record, bool single) {
    if (single)
        return record ? event_kernel<CAND, REAL, VETO, true, true, kWarpsPerBlock, false, false>
                      : event_kernel<CAND, REAL, VETO, true, false, kWarpsPerBlock, false, false>;
    return record ? event_kernel<CAND, REAL, VETO, false, true, kWarpsPerBlock, false, false>
                  : event_kernel<CAND, REAL, VETO, false, false, kWarpsPerBlock, false, false>;
}
template <int CAND, int REAL, int VETO>
EventKernel pick_composite(bool record) {
    return record ? event_kernel<CAND, REAL, VETO, false, true, kWarpsPerBlock, true, false>
                  : event_kernel<CAND, REAL, VETO, false, false, kWarpsPerBlock, true, false>;
}
// cell-bounding far field: the reference's handler needs one occupant per cell (SINGLE)
template <int CAND, int REAL, int VETO>
EventKernel pick_far_pairs(bool record) {
    return record ? event_kernel<CAND, REAL, VETO, true, true, kWarpsPerBlock, false, true>
                  : event_kernel<CAND, REAL, VETO, true, false, kWarpsPerBlock, false, true>;
}

EventKernel pick_kernel(const DeviceProgram &d, bool record) {
    const int cand = d.pair_handler == ECMC_PAIR_NONE ? 0 : d.cand_potential.kind;
    const int real = d.pair_handler == ECMC_PAIR_TWO_LEAF_UNIT_BOUNDING ? d.real_potential.kind : 0;
    const int veto = d.veto_enabled ? d.veto_potential.kind : 0;
    const int LJ = ECMC_POT_LENNARD_JONES, HS = ECMC_POT_HARD_SPHERE, MIC = ECMC_POT_MERGED_IMAGE_COULOMB,
              IPCB = ECMC_POT_INVERSE_POWER_COULOMB_BOUNDING;
    const bool single = d.max_occupants == 1;
    if (d.nodes_per_root > 1) {
        // composite point objects: hard disks / spheres tethered into dipoles (C1), anything else generic
        if (cand == HS && real == 0 && veto == 0) return pick_composite<ECMC_POT_HARD_SPHERE, 0, 0>(record);
        return pick_composite<-1, -1, -1>(record);
    }
    if (d.veto_enabled == ECMC_FAR_CELL_BOUNDING) {
        if (cand == IPCB && real == MIC && veto == MIC)
            return pick_far_pairs<ECMC_POT_INVERSE_POWER_COULOMB_BOUNDING, ECMC_POT_MERGED_IMAGE_COULOMB, ECMC_POT_MERGED_IMAGE_COULOMB>(record);
        return pick_far_pairs<-1, -1, -1>(record);
    }
    // the Lennard-Jones kernels assume chargeless handlers and modular cell translations (FAST in event_kernel)
    const bool fast = !d.pair_use_charge && !d.veto_use_charge && d.translate_modular && !d.no_cells && d.dimension == 3;
    if (fast && cand == LJ && real == 0 && veto == LJ) return pick_record<ECMC_POT_LENNARD_JONES, 0, ECMC_POT_LENNARD_JONES>(record, single);
    if (fast && cand == LJ && real == 0 && veto == 0) return pick_record<ECMC_POT_LENNARD_JONES, 0, 0>(record, single);
    if (cand == HS && real == 0 && veto == 0) return pick_record<ECMC_POT_HARD_SPHERE, 0, 0>(record, false);
    if (cand == IPCB && real == MIC && veto == MIC && !d.no_cells)
        return pick_record<ECMC_POT_INVERSE_POWER_COULOMB_BOUNDING, ECMC_POT_MERGED_IMAGE_COULOMB, ECMC_POT_MERGED_IMAGE_COULOMB>(record, single);
    return pick_record<-1, -1, -1>(record, false);
}

// ---- the Lennard-Jones / cell-veto kernel that evaluates several events side by side (ecmc_spec.cuh) -------------
struct SpecLaunch {
    EventKernel kernel = nullptr;
    size_t shared_bytes = 0;
    int capacity = 0;
    bool chain_blocks = false;  // lj_chain_kernel: one CTA of kChainWarps warps per chain
    bool coulomb = false;       // the Coulomb model of lj_spec_kernel
    int warps = kWarpsPerBlock; // chains per CTA
};
// lj_chain_kernel while every chain can have an SM of its own
constexpr int kChainKernelMaxChains = 148;

template <bool RECORD, bool PRUNE>
EventKernel pick_spec_lanes(int lanes) {
    return lanes == 8 ? lj_spec_kernel<RECORD, PRUNE, 8, kWarpsPerBlock> : lj_spec_kernel<RECORD, PRUNE, 4, kWarpsPerBlock>;
}
// the Coulomb atoms (C3): the same batch with the inverse-power Coulomb bound, merged-image Coulomb and charges
template <bool RECORD, bool PRUNE, int WARPS>
EventKernel pick_spec_coulomb(int lanes) {
    return lanes == 8 ? lj_spec_kernel<RECORD, PRUNE, 8, WARPS, false, kSpecCoulomb>
                      : lj_spec_kernel<RECORD, PRUNE, 4, WARPS, false, kSpecCoulomb>;
}
template <int WARPS>
EventKernel pick_spec_coulomb_warps(bool record, bool prune, int lanes) {
    if (record) return pick_spec_coulomb<true, false, WARPS>(lanes);
    return prune ? pick_spec_coulomb<false, true, WARPS>(lanes) : pick_spec_coulomb<false, false, WARPS>(lanes);
}
// the whole-host-step form of the same kernel (RunArgs.host_in / host_out)
EventKernel pick_spec_host(bool prune, int lanes) {
    if (prune) return lanes == 8 ? lj_spec_kernel<false, true, 8, kWarpsPerBlock, true> : lj_spec_kernel<false, true, 4, kWarpsPerBlock, true>;
    return lanes == 8 ? lj_spec_kernel<false, false, 8, kWarpsPerBlock, true> : lj_spec_kernel<false, false, 4, kWarpsPerBlock, true>;
}

// Chargeless 3D Lennard-Jones pair factors with a Lennard-Jones cell veto, one occupant per cell, modular cell
// translations, and a candidate list (nearby cells + the longest possible surplus list) that fits the shared memory
// of a CTA next to its sibling on the SM. Everything else runs event_kernel.
bool pick_spec(const EcmcHandle *h, bool record, SpecLaunch *out) {
    const DeviceProgram &d = h->dprog;
    if (!h->spec || h->molecules || d.nodes_per_root > 1) return false;
    if (d.dimension != 3 || d.no_cells || !d.translate_modular) return false;
    if (d.max_occupants != 1 || d.veto_enabled != ECMC_FAR_CELL_VETO) return false;
    const bool lennard_jones = d.pair_handler == ECMC_PAIR_TWO_LEAF_UNIT && !d.pair_use_charge && !d.veto_use_charge &&
                               d.cand_potential.kind == ECMC_POT_LENNARD_JONES && d.veto_potential.kind == ECMC_POT_LENNARD_JONES;
    // Coulomb atoms: pair candidates from the inverse-power Coulomb bound confirmed against merged-image Coulomb, the same
    // potential for the cell veto
    const bool coulomb = d.pair_handler == ECMC_PAIR_TWO_LEAF_UNIT_BOUNDING && h->spec_coulomb &&
                         d.cand_potential.kind == ECMC_POT_INVERSE_POWER_COULOMB_BOUNDING &&
                         d.real_potential.kind == ECMC_POT_MERGED_IMAGE_COULOMB &&
                         d.veto_potential.kind == ECMC_POT_MERGED_IMAGE_COULOMB;
    if (!lennard_jones && !coulomb) return false;
    for (int k = 0; k < 3; k++)
        if (d.upper[k].n_entries <= 0 || (coulomb && d.veto_use_charge && d.lower[k].n_entries <= 0)) return false;
    const bool prune = h->spec_prune && !record;
    const int capacity = (d.n_nearby + d.max_surplus + 31) / 32 * 32;
    // per entry: coordinate, squared distance from the line, (force bound), (charge product), target + sequence number,
    // (live index); Coulomb: plus the scratch of the Ewald sum per warp
    // warps (chains) per CTA. The Coulomb model aligns the warps of a CTA at a barrier per batch, and the more warps fetch
    // the same instructions together the better (measured on C3, N = 64: 7 / 14 / 28 warps -> 5.2 / 6.3 / 8.0e8 events/s):
    // the largest CTA whose lists fit the shared memory of an SM, even if fewer warps than the register file allows are
    // then resident
    int warps = kWarpsPerBlock;
    auto bytes_of = [&](int w) {
        return (size_t)w * capacity * ((prune ? 5 : 3) + (coulomb ? 1 : 0)) * sizeof(double) +
               (coulomb ? (size_t)w * kTrigDoubles * sizeof(double) : 0);
    };
    if (coulomb) {
        warps = h->coulomb_warps;
        if (warps == 0)
            for (int w : {28, 14, 7})
                if (warps == 0 && bytes_of(w) <= 200 * 1024) warps = w;
        if (warps == 0) return false;
    }
    const size_t bytes = bytes_of(warps);
    if (coulomb ? bytes > 200 * 1024 : bytes > 100 * 1024) return false;  // Lennard-Jones: two CTAs per SM
    out->capacity = capacity;
    out->shared_bytes = bytes;
    out->coulomb = coulomb;
    out->warps = warps;
    if (coulomb) {
        out->kernel = warps == 7 ? pick_spec_coulomb_warps<7>(record, prune, h->spec_lanes)
                                 : (warps == 28 ? pick_spec_coulomb_warps<28>(record, prune, h->spec_lanes)
                                                : pick_spec_coulomb_warps<14>(record, prune, h->spec_lanes));
        return true;
    }
    // few chains (the single large chain C5): one CTA of four warps per chain, 32 events per batch
    out->chain_blocks = h->chain_blocks && h->n_chains <= kChainKernelMaxChains;
    if (out->chain_blocks) {
        out->shared_bytes = (size_t)capacity * (prune ? 5 : 3) * sizeof(double);
        if (record) out->kernel = lj_chain_kernel<true, false>;
        else out->kernel = prune ? lj_chain_kernel<false, true> : lj_chain_kernel<false, false>;
        return true;
    }
    if (record) out->kernel = pick_spec_lanes<true, false>(h->spec_lanes);
    else out->kernel = prune ? pick_spec_lanes<false, true>(h->spec_lanes) : pick_spec_lanes<false, false>(h->spec_lanes);
    return true;
}

int acquire_events(EcmcHandle *h, EventPair *out) {
    if (!h->free_events.empty()) {
        *out = h->free_events.back();
        h->free_events.pop_back();
        return ECMC_OK;
    }
    CUDA_TRY(h, cudaEventCreate(&out->start));
    CUDA_TRY(h, cudaEventCreate(&out->stop));
    return ECMC_OK;
}

int collect_timings(EcmcHandle *h) {
    for (EventPair &e : h->timed) {
        float ms = 0.0f;
        CUDA_TRY(h, cudaEventElapsedTime(&ms, e.start, e.stop));
        h->kernel_seconds += 1.0e-3 * ms;
        h->free_events.push_back(e);
    }
    h->timed.clear();
    return ECMC_OK;
}

int launch_events(EcmcHandle *h, double until_q, double until_r, int64_t max_events, EcmcEventRecord *d_records,
                  int records_per_chain) {
    if (!h->started) return fail(h, ECMC_ERR_STATE, "ecmc_run before ecmc_start / ecmc_upload_chain_states");
    if (std::isnan(until_q) || std::isnan(until_r)) return fail(h, ECMC_ERR_INVALID, "until time is NaN");
    if (max_events <= 0 && std::isinf(until_q))
        return fail(h, ECMC_ERR_INVALID, "neither a time limit nor an event limit: the run would not end");
    CUDA_TRY(h, cudaSetDevice(h->device));
    RunArgs args{};
    args.until_q = until_q;
    args.until_r = until_r;
    args.max_events = max_events;
    args.records = d_records;
    args.records_per_chain = records_per_chain;
    args.stats = h->d_stats;
    args.list_capacity = 0;
    EventPair ev;
    int rc = acquire_events(h, &ev);
    if (rc) return rc;
    const int blocks = (h->n_chains + kWarpsPerBlock - 1) / kWarpsPerBlock;
    CUDA_TRY(h, cudaEventRecord(ev.start, h->stream));
    if (h->disks) {
        constexpr int kDiskWarps = 4;
        const int disk_blocks = (h->n_chains + kDiskWarps - 1) / kDiskWarps;
        if (d_records) disk_kernel<true, kDiskWarps><<<disk_blocks, kDiskWarps * 32, 0, h->stream>>>(h->dprog, h->kprog, h->state, args);
        else disk_kernel<false, kDiskWarps><<<disk_blocks, kDiskWarps * 32, 0, h->stream>>>(h->dprog, h->kprog, h->state, args);
    } else if (h->molecules) {
        typedef void (*MoleculeKernel)(const DeviceProgram, const MoleculeProgram, const DeviceState, const RunArgs);
        const int IPCB = ECMC_POT_INVERSE_POWER_COULOMB_BOUNDING, MIC = ECMC_POT_MERGED_IMAGE_COULOMB,
                  DEP = kPotHarmonic, LJ = ECMC_POT_LENNARD_JONES;
        MoleculeKernel kernel;
        const int warps = molecule_cta_warps(h);
        const bool water = warps < 0 || warps > kMoleculeWarps, aligned = warps > kMoleculeWarps, wide = warps == kWideWarps;
        if (h->mprog.root_mode)
            kernel = d_records ? molecule_kernel<-1, -1, -1, -1, true, kMoleculeWarps, false, true>
                               : molecule_kernel<-1, -1, -1, -1, false, kMoleculeWarps, false, true>;
        else if (h->mprog.cell_child)
            kernel = d_records ? molecule_kernel<-1, -1, -1, -1, true, kMoleculeWarps, false, false, true>
                               : molecule_kernel<-1, -1, -1, -1, false, kMoleculeWarps, false, false, true>;
        else if (water && aligned && wide)
            kernel = d_records ? molecule_kernel<IPCB, MIC, DEP, LJ, true, kWideWarps, true>
                               : molecule_kernel<IPCB, MIC, DEP, LJ, false, kWideWarps, true>;
        else if (water && aligned)
            kernel = d_records ? molecule_kernel<IPCB, MIC, DEP, LJ, true, kAlignedWarps, true>
                               : molecule_kernel<IPCB, MIC, DEP, LJ, false, kAlignedWarps, true>;
        else if (water)
            kernel = d_records ? molecule_kernel<IPCB, MIC, DEP, LJ, true, kMoleculeWarps, false>
                               : molecule_kernel<IPCB, MIC, DEP, LJ, false, kMoleculeWarps, false>;
        else
            kernel = d_records ? molecule_kernel<-1, -1, -1, -1, true, kMoleculeWarps, false>
                               : molecule_kernel<-1, -1, -1, -1, false, kMoleculeWarps, false>;
        const int cta_warps = warps < 0 ? kMoleculeWarps : warps;
        kernel<<<(h->n_chains + cta_warps - 1) / cta_warps, cta_warps * 32, 0, h->stream>>>(h->dprog, h->mprog, h->state, args);
    } else {
        SpecLaunch spec;
        if (pick_spec(h, d_records != nullptr, &spec)) {
            args.list_capacity = spec.capacity;
            CUDA_TRY(h, cudaFuncSetAttribute(spec.kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)spec.shared_bytes));
            if (spec.chain_blocks)
                spec.kernel<<<h->n_chains, kChainWarps * 32, spec.shared_bytes, h->stream>>>(h->dprog, h->state, args);
            else
                spec.kernel<<<(h->n_chains + spec.warps - 1) / spec.warps, spec.warps * 32, spec.shared_bytes, h->stream>>>(
                    h->dprog, h->state, args);
        } else {
            const EventKernel kernel = pick_kernel(h->dprog, d_records != nullptr);
            kernel<<<blocks, kWarpsPerBlock * 32, 0, h->stream>>>(h->dprog, h->state, args);
        }
    }
    CUDA_TRY(h, cudaGetLastError());
    CUDA_TRY(h, cudaEventRecord(ev.stop, h->stream));
    h->timed.push_back(ev);
    h->kernel_launches++;
    return ECMC_OK;
}

}  // namespace

// ==========================================================================================================
// lifecycle
// ==========================================================================================================
ECMC_API int ecmc_abi_version(void) { return ECMC_ABI_VERSION; }

ECMC_API const char *ecmc_last_error(const EcmcHandle *h) { return h ? h->error.c_str() : g_create_error.c_str(); }

ECMC_API void ecmc_destroy(EcmcHandle *h) {
    if (!h) return;
    cudaSetDevice(h->device);
    if (h->stream) cudaStreamSynchronize(h->stream);
    for (EventPair &e : h->timed) { cudaEventDestroy(e.start); cudaEventDestroy(e.stop); }
    for (EventPair &e : h->free_events) { cudaEventDestroy(e.start); cudaEventDestroy(e.stop); }
    for (void *p : h->allocations) cudaFree(p);
    if (h->h_stats) cudaFreeHost(h->h_stats);
    for (cudaStream_t s : h->slice_streams) cudaStreamDestroy(s);
    if (h->stream) cudaStreamDestroy(h->stream);
    delete h;
}

ECMC_API int ecmc_create(const EcmcProgram *program, int device, int n_chains, EcmcHandle **out) {
    if (!out) return fail(nullptr, ECMC_ERR_INVALID, "out is NULL");
    *out = nullptr;
    if (!program) return fail(nullptr, ECMC_ERR_INVALID, "program is NULL");
    if (n_chains < 1) return fail(nullptr, ECMC_ERR_INVALID, "n_chains must be >= 1");
    int count = 0;
    cudaError_t err = cudaGetDeviceCount(&count);
    if (err != cudaSuccess || count == 0)
        return fail(nullptr, ECMC_ERR_CUDA, std::string("no CUDA device: ") + cudaGetErrorString(err) +
                                                " (libecmc_b200 has no CPU fallback)");
    if (device < 0 || device >= count) return fail(nullptr, ECMC_ERR_INVALID, "device index out of range");
    EcmcHandle *h = new (std::nothrow) EcmcHandle();
    if (!h) return fail(nullptr, ECMC_ERR_INVALID, "out of host memory");
    h->device = device;
    h->n_chains = n_chains;
    h->program = *program;
    // development switches (A/B measurements); the supported interface is ecmc_set_option
    if (const char *env = std::getenv("ECMC_SPEC")) h->spec = std::atoi(env) != 0;
    if (const char *env = std::getenv("ECMC_SPEC_PRUNE")) h->spec_prune = std::atoi(env) != 0;
    if (const char *env = std::getenv("ECMC_SPEC_LANES")) h->spec_lanes = std::atoi(env) == 8 ? 8 : 4;
    if (const char *env = std::getenv("ECMC_CHAIN_BLOCKS")) h->chain_blocks = std::atoi(env) != 0;
    if (const char *env = std::getenv("ECMC_COULOMB_WARPS")) {
        const int warps = std::atoi(env);
        if (warps == 7 || warps == 14 || warps == 28) h->coulomb_warps = warps;
    }
    int rc = ECMC_OK;
    do {
        if ((err = cudaSetDevice(device)) != cudaSuccess) { rc = fail(h, ECMC_ERR_CUDA, cudaGetErrorString(err)); break; }
        if ((err = cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking)) != cudaSuccess) {
            rc = fail(h, ECMC_ERR_CUDA, cudaGetErrorString(err));
            break;
        }
        if ((rc = build_device_program(h))) break;
        const DeviceProgram &d = h->dprog;
        const size_t n = (size_t)n_chains * d.n_particles;
        h->state.n_chains = n_chains;
        h->state.first_chain = 0;
        if ((rc = device_alloc(h, &h->state.particles, n))) break;
        if (d.nodes_per_root > 1 && (rc = device_alloc(h, &h->state.roots, n / d.nodes_per_root))) break;
        if ((rc = device_alloc(h, &h->state.occupants, (size_t)n_chains * d.n_cells * d.max_occupants))) break;
        if ((rc = device_alloc(h, &h->state.surplus, (size_t)n_chains * d.max_surplus))) break;
        if ((rc = device_alloc(h, &h->state.n_surplus, (size_t)n_chains))) break;
        if ((rc = device_alloc(h, &h->state.chains, (size_t)n_chains))) break;
        if ((rc = device_alloc(h, &h->d_stats, 1))) break;
        if ((rc = device_alloc(h, &h->d_staging, n * d.dimension))) break;
        if ((rc = device_alloc(h, &h->d_staging_charges, n))) break;
        if ((rc = device_alloc(h, &h->d_streams, (size_t)n_chains))) break;
        if ((err = cudaMallocHost(&h->h_stats, sizeof(EcmcStats))) != cudaSuccess) {
            rc = fail(h, ECMC_ERR_CUDA, cudaGetErrorString(err));
            break;
        }
        if ((err = cudaMemsetAsync(h->d_stats, 0, sizeof(EcmcStats), h->stream)) != cudaSuccess ||
            (err = cudaMemsetAsync(h->state.n_surplus, 0, sizeof(int) * n_chains, h->stream)) != cudaSuccess ||
            (err = cudaStreamSynchronize(h->stream)) != cudaSuccess) {
            rc = fail(h, ECMC_ERR_CUDA, cudaGetErrorString(err));
            break;
        }
    } while (0);
    if (rc) {
        g_create_error = h->error;
        ecmc_destroy(h);
        return rc;
    }
    h->program.veto_tables = nullptr;  // copied; the caller's arrays are not referenced after create
    *out = h;
    return ECMC_OK;
}

// ==========================================================================================================
// state
// ==========================================================================================================
ECMC_API int ecmc_upload_positions(EcmcHandle *h, const double *positions, const double *charges) {
    if (!h || !positions) return fail(h, ECMC_ERR_INVALID, "null argument");
    CUDA_TRY(h, cudaSetDevice(h->device));
    const size_t n = (size_t)h->n_chains * h->dprog.n_particles;
    CUDA_TRY(h, cudaMemcpyAsync(h->d_staging, positions, n * h->dprog.dimension * sizeof(double), cudaMemcpyHostToDevice, h->stream));
    if (charges)
        CUDA_TRY(h, cudaMemcpyAsync(h->d_staging_charges, charges, n * sizeof(double), cudaMemcpyHostToDevice, h->stream));
    const int blocks = (int)std::min<size_t>((n + 255) / 256, 148 * 16);
    pack_particles_kernel<<<blocks, 256, 0, h->stream>>>(h->d_staging, charges ? h->d_staging_charges : nullptr,
                                                         h->state.particles, n, h->dprog.dimension);
    CUDA_TRY(h, cudaGetLastError());
    return ECMC_OK;
}

ECMC_API int ecmc_download_positions(EcmcHandle *h, double *positions) {
    if (!h || !positions) return fail(h, ECMC_ERR_INVALID, "null argument");
    CUDA_TRY(h, cudaSetDevice(h->device));
    const size_t n = (size_t)h->n_chains * h->dprog.n_particles;
    const int blocks = (int)std::min<size_t>((n + 255) / 256, 148 * 16);
    unpack_particles_kernel<<<blocks, 256, 0, h->stream>>>(h->state.particles, h->d_staging, n, h->dprog.dimension);
    CUDA_TRY(h, cudaGetLastError());
    CUDA_TRY(h, cudaMemcpyAsync(positions, h->d_staging, n * h->dprog.dimension * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
    CUDA_TRY(h, cudaStreamSynchronize(h->stream));
    return collect_timings(h);
}

ECMC_API int ecmc_download_chain(EcmcHandle *h, int chain, double *positions, double *roots, EcmcChainState *state) {
    if (!h || !positions) return fail(h, ECMC_ERR_INVALID, "null argument");
    if (chain < 0 || chain >= h->n_chains) return fail(h, ECMC_ERR_INVALID, "chain index out of range");
    const DeviceProgram &d = h->dprog;
    if (roots && d.nodes_per_root <= 1) return fail(h, ECMC_ERR_INVALID, "the program has no composite objects");
    CUDA_TRY(h, cudaSetDevice(h->device));
    const size_t n = (size_t)d.n_particles, n_roots = roots ? n / d.nodes_per_root : 0;
    unpack_particles_kernel<<<(int)((n + 255) / 256), 256, 0, h->stream>>>(h->state.particles + (size_t)chain * n, h->d_staging,
                                                                        n, d.dimension);
    CUDA_TRY(h, cudaGetLastError());
    CUDA_TRY(h, cudaMemcpyAsync(positions, h->d_staging, n * d.dimension * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
    if (roots) {  // (the staging buffer is reused: stream order keeps the two copies apart)
        unpack_particles_kernel<<<(int)((n_roots + 255) / 256), 256, 0, h->stream>>>(h->state.roots + (size_t)chain * n_roots,
                                                                                  h->d_staging, n_roots, d.dimension);
        CUDA_TRY(h, cudaGetLastError());
        CUDA_TRY(h, cudaMemcpyAsync(roots, h->d_staging, n_roots * d.dimension * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
    }
    if (state)
        CUDA_TRY(h, cudaMemcpyAsync(state, h->state.chains + chain, sizeof(EcmcChainState), cudaMemcpyDeviceToHost, h->stream));
    CUDA_TRY(h, cudaStreamSynchronize(h->stream));
    return collect_timings(h);
}

ECMC_API int ecmc_upload_roots(EcmcHandle *h, const double *roots) {
    if (!h || !roots) return fail(h, ECMC_ERR_INVALID, "null argument");
    if (h->dprog.nodes_per_root <= 1) return fail(h, ECMC_ERR_INVALID, "the program has no composite objects");
    CUDA_TRY(h, cudaSetDevice(h->device));
    const size_t n = (size_t)h->n_chains * (h->dprog.n_particles / h->dprog.nodes_per_root);
    CUDA_TRY(h, cudaMemcpyAsync(h->d_staging, roots, n * h->dprog.dimension * sizeof(double), cudaMemcpyHostToDevice, h->stream));
    const int blocks = (int)std::min<size_t>((n + 255) / 256, 148 * 16);
    pack_particles_kernel<<<blocks, 256, 0, h->stream>>>(h->d_staging, nullptr, h->state.roots, n, h->dprog.dimension);
    CUDA_TRY(h, cudaGetLastError());
    CUDA_TRY(h, cudaStreamSynchronize(h->stream));
    h->roots_uploaded = true;
    return ECMC_OK;
}

ECMC_API int ecmc_download_roots(EcmcHandle *h, double *roots) {
    if (!h || !roots) return fail(h, ECMC_ERR_INVALID, "null argument");
    if (h->dprog.nodes_per_root <= 1) return fail(h, ECMC_ERR_INVALID, "the program has no composite objects");
    CUDA_TRY(h, cudaSetDevice(h->device));
    const size_t n = (size_t)h->n_chains * (h->dprog.n_particles / h->dprog.nodes_per_root);
    const int blocks = (int)std::min<size_t>((n + 255) / 256, 148 * 16);
    unpack_particles_kernel<<<blocks, 256, 0, h->stream>>>(h->state.roots, h->d_staging, n, h->dprog.dimension);
    CUDA_TRY(h, cudaGetLastError());
    CUDA_TRY(h, cudaMemcpyAsync(roots, h->d_staging, n * h->dprog.dimension * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
    CUDA_TRY(h, cudaStreamSynchronize(h->stream));
    return ECMC_OK;
}

ECMC_API int ecmc_start(EcmcHandle *h, const uint32_t *streams, uint32_t first_stream) {
    if (!h) return fail(h, ECMC_ERR_INVALID, "null handle");
    if (h->dprog.nodes_per_root > 1 && !h->roots_uploaded)
        return fail(h, ECMC_ERR_STATE, "composite objects: ecmc_upload_roots before ecmc_start");
    CUDA_TRY(h, cudaSetDevice(h->device));
    if (streams)
        CUDA_TRY(h, cudaMemcpyAsync(h->d_streams, streams, sizeof(uint32_t) * h->n_chains, cudaMemcpyHostToDevice, h->stream));
    const int blocks = (h->n_chains + kWarpsPerBlock - 1) / kWarpsPerBlock;
    if (h->molecules)
        molecule_start_kernel<kMoleculeWarps><<<(h->n_chains + kMoleculeWarps - 1) / kMoleculeWarps, kMoleculeWarps * 32, 0, h->stream>>>(
            h->dprog, h->state, streams ? h->d_streams : nullptr, first_stream, h->program.initial_active,
            h->program.initial_direction, h->d_stats, h->mprog.root_mode ? h->mprog.switch_length[0] : 0.0,
            h->mprog.cell_child);
    else
        start_kernel<kWarpsPerBlock><<<blocks, kWarpsPerBlock * 32, 0, h->stream>>>(
            h->dprog, h->state, streams ? h->d_streams : nullptr, first_stream, h->program.initial_active,
            h->program.initial_direction, h->d_stats, false);
    if (h->disks)
        disk_start_kernel<<<(h->n_chains + 127) / 128, 128, 0, h->stream>>>(h->state, h->dprog.speed, h->kprog.weight,
                                                                           h->kprog.initial_direction);
    CUDA_TRY(h, cudaGetLastError());
    h->started = true;
    return ECMC_OK;
}

ECMC_API int ecmc_upload_chain_states(EcmcHandle *h, const EcmcChainState *states) {
    if (!h || !states) return fail(h, ECMC_ERR_INVALID, "null argument");
    for (int c = 0; c < h->n_chains; c++) {
        const EcmcChainState &s = states[c];
        if (s.active < 0 || s.active >= h->dprog.n_particles || s.direction < 0 || s.direction >= h->dprog.dimension ||
            s.active_cell < 0 || s.active_cell >= h->dprog.n_cells || s.eoc_next_active < 0 ||
            s.eoc_next_active >= h->dprog.n_particles)
            return fail(h, ECMC_ERR_INVALID, "chain state " + std::to_string(c) + " out of range");
        // kept candidates index particles / cells on the device: a corrupt checkpoint must not read out of bounds
        const int index_limit = std::max(h->dprog.n_particles, h->dprog.n_cells);
        if (s.pending_kind < ECMC_EVENT_NONE || s.pending_kind > 16 ||
            (s.pending_kind != ECMC_EVENT_NONE && (s.pending_target < -1 || s.pending_target >= index_limit)) ||
            s.kept_kind < -1 || s.kept_kind > 16 || (s.kept_kind > 0 && (s.kept_target < -1 || s.kept_target >= index_limit)))
            return fail(h, ECMC_ERR_INVALID, "chain state " + std::to_string(c) + ": kept candidate out of range");
        // root-unit-active mode: `active` is the first leaf of the object whose root unit is active
        if (s.mode != 0 && (s.mode != 1 || !h->mprog.root_mode || s.active % 2 != 0 || s.eoc_next_active % 2 != 0))
            return fail(h, ECMC_ERR_INVALID, "chain state " + std::to_string(c) + ": mode out of range");
    }
    CUDA_TRY(h, cudaSetDevice(h->device));
    CUDA_TRY(h, cudaMemcpyAsync(h->state.chains, states, sizeof(EcmcChainState) * h->n_chains, cudaMemcpyHostToDevice, h->stream));
    CUDA_TRY(h, cudaStreamSynchronize(h->stream));
    h->started = true;
    return ECMC_OK;
}

ECMC_API int ecmc_download_chain_states(EcmcHandle *h, EcmcChainState *states) {
    if (!h || !states) return fail(h, ECMC_ERR_INVALID, "null argument");
    CUDA_TRY(h, cudaSetDevice(h->device));
    CUDA_TRY(h, cudaMemcpyAsync(states, h->state.chains, sizeof(EcmcChainState) * h->n_chains, cudaMemcpyDeviceToHost, h->stream));
    CUDA_TRY(h, cudaStreamSynchronize(h->stream));
    return collect_timings(h);
}

ECMC_API int ecmc_upload_cells(EcmcHandle *h, const int32_t *occupants, const int32_t *surplus, const int32_t *n_surplus) {
    if (!h || !occupants || !surplus || !n_surplus) return fail(h, ECMC_ERR_INVALID, "null argument");
    const DeviceProgram &d = h->dprog;
    for (int c = 0; c < h->n_chains; c++)
        if (n_surplus[c] < 0 || n_surplus[c] > d.max_surplus) return fail(h, ECMC_ERR_INVALID, "n_surplus out of range");
    CUDA_TRY(h, cudaSetDevice(h->device));
    CUDA_TRY(h, cudaMemcpyAsync(h->state.occupants, occupants, sizeof(int) * (size_t)h->n_chains * d.n_cells * d.max_occupants,
                                cudaMemcpyHostToDevice, h->stream));
    CUDA_TRY(h, cudaMemcpyAsync(h->state.surplus, surplus, sizeof(int) * (size_t)h->n_chains * d.max_surplus,
                                cudaMemcpyHostToDevice, h->stream));
    CUDA_TRY(h, cudaMemcpyAsync(h->state.n_surplus, n_surplus, sizeof(int) * (size_t)h->n_chains, cudaMemcpyHostToDevice, h->stream));
    CUDA_TRY(h, cudaStreamSynchronize(h->stream));
    return ECMC_OK;
}

ECMC_API int ecmc_download_cells(EcmcHandle *h, int32_t *occupants, int32_t *surplus, int32_t *n_surplus) {
    if (!h || !occupants || !surplus || !n_surplus) return fail(h, ECMC_ERR_INVALID, "null argument");
    const DeviceProgram &d = h->dprog;
    CUDA_TRY(h, cudaSetDevice(h->device));
    CUDA_TRY(h, cudaMemcpyAsync(occupants, h->state.occupants, sizeof(int) * (size_t)h->n_chains * d.n_cells * d.max_occupants,
                                cudaMemcpyDeviceToHost, h->stream));
    CUDA_TRY(h, cudaMemcpyAsync(surplus, h->state.surplus, sizeof(int) * (size_t)h->n_chains * d.max_surplus,
                                cudaMemcpyDeviceToHost, h->stream));
    CUDA_TRY(h, cudaMemcpyAsync(n_surplus, h->state.n_surplus, sizeof(int) * (size_t)h->n_chains, cudaMemcpyDeviceToHost, h->stream));
    CUDA_TRY(h, cudaStreamSynchronize(h->stream));
    return collect_timings(h);
}

// ==========================================================================================================
// the hot path
// ==========================================================================================================
ECMC_API int ecmc_run(EcmcHandle *h, double until_q, double until_r, int64_t max_events_per_chain) {
    if (!h) return fail(h, ECMC_ERR_INVALID, "null handle");
    return launch_events(h, until_q, until_r, max_events_per_chain, nullptr, 0);
}

ECMC_API int ecmc_sync(EcmcHandle *h, EcmcStats *stats) {
    if (!h) return fail(h, ECMC_ERR_INVALID, "null handle");
    CUDA_TRY(h, cudaSetDevice(h->device));
    CUDA_TRY(h, cudaMemcpyAsync(h->h_stats, h->d_stats, sizeof(EcmcStats), cudaMemcpyDeviceToHost, h->stream));
    CUDA_TRY(h, cudaMemsetAsync(h->d_stats, 0, sizeof(EcmcStats), h->stream));
    CUDA_TRY(h, cudaStreamSynchronize(h->stream));
    if (stats) *stats = *h->h_stats;
    int rc = collect_timings(h);
    if (rc) return rc;
    if (h->h_stats->capacity_errors)
        return fail(h, ECMC_ERR_CAPACITY, "surplus list / occupant capacity overflowed in " +
                                              std::to_string(h->h_stats->capacity_errors) + " place(s): raise max_surplus");
    return ECMC_OK;
}

ECMC_API int ecmc_run_recorded(EcmcHandle *h, double until_q, double until_r, int64_t max_events_per_chain,
                               EcmcEventRecord *records, int32_t records_per_chain, EcmcStats *stats) {
    if (!h || !records || records_per_chain < 1) return fail(h, ECMC_ERR_INVALID, "null / empty record buffer");
    CUDA_TRY(h, cudaSetDevice(h->device));
    const size_t n = (size_t)h->n_chains * records_per_chain;
    EcmcEventRecord *d_records = nullptr;
    CUDA_TRY(h, cudaMalloc(&d_records, n * sizeof(EcmcEventRecord)));
    cudaMemsetAsync(d_records, 0, n * sizeof(EcmcEventRecord), h->stream);
    int rc = launch_events(h, until_q, until_r, max_events_per_chain, d_records, records_per_chain);
    if (!rc) {
        cudaError_t err = cudaMemcpyAsync(records, d_records, n * sizeof(EcmcEventRecord), cudaMemcpyDeviceToHost, h->stream);
        if (err != cudaSuccess) rc = fail(h, ECMC_ERR_CUDA, cudaGetErrorString(err));
    }
    const int rc_sync = ecmc_sync(h, stats);
    cudaFree(d_records);
    return rc ? rc : rc_sync;
}

// Host buffers in, host buffers out. The chains are cut into slices; every slice runs its own
// H2D -> pack -> start -> events -> unpack -> D2H sequence on its own stream, so the copies of one slice overlap the
// event kernels of the others (the kernels of different slices run concurrently: a slice fills only part of the GPU).
// ecmc_submit_from_host only enqueues: successive steps are ordered slice by slice by their streams, so a step may read
// the host buffer the step before it writes, and the copies of a slice overlap the events of the other slices across
// steps as well. ecmc_wait synchronises the slices and returns the counters of all steps submitted since the last wait.
namespace {

// Sparse write-back of a step: every particle whose coordinates differ from the staged input of the step is written
// straight into the caller's pinned host buffer (a device-visible mapping of it): only the changed coordinates cross the
// link -- the particles that were active during the step, about one in ten for a step of 1024 events on 1024 particles.
__global__ void __launch_bounds__(256)
write_changed_particles_kernel(const Particle *particles, const double *staged, double *host_out, size_t n, int dimension,
                               unsigned long long *changed) {
    unsigned mine = 0;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        const Particle p = particles[i];
        const double *in = staged + i * dimension;
        bool differs = __double_as_longlong(p.x) != __double_as_longlong(in[0]);
        if (dimension > 1) differs = differs || __double_as_longlong(p.y) != __double_as_longlong(in[1]);
        if (dimension > 2) differs = differs || __double_as_longlong(p.z) != __double_as_longlong(in[2]);
        if (differs) {
            double *out = host_out + i * dimension;
            out[0] = p.x;
            if (dimension > 1) out[1] = p.y;
            if (dimension > 2) out[2] = p.z;
            mine++;
        }
    }
    mine = __reduce_add_sync(0xffffffffu, mine);
    if ((threadIdx.x & 31) == 0 && mine) atomicAdd(changed, (unsigned long long)mine);
}

int submit_from_host(EcmcHandle *h, const double *positions_in, const double *charges, uint32_t first_stream,
                     double until_q, double until_r, int64_t max_events_per_chain, double *positions_out, bool sparse) {
    if (!h || !positions_in) return fail(h, ECMC_ERR_INVALID, "null argument");
    double *mapped_out = nullptr;
    if (sparse) {
        if (!positions_out) return fail(h, ECMC_ERR_INVALID, "sparse write-back needs positions_out");
        CUDA_TRY(h, cudaSetDevice(h->device));
        cudaPointerAttributes attributes{};
        if (cudaPointerGetAttributes(&attributes, positions_out) != cudaSuccess || attributes.type != cudaMemoryTypeHost ||
            !attributes.devicePointer) {
            cudaGetLastError();
            return fail(h, ECMC_ERR_INVALID, "sparse write-back needs positions_out in pinned host memory the device can "
                                             "address (ecmc_host_alloc, cudaHostAlloc, cudaHostRegister)");
        }
        mapped_out = static_cast<double *>(attributes.devicePointer);
        if (!h->d_changed) {
            int rc_alloc = device_alloc(h, &h->d_changed, 1);
            if (rc_alloc) return rc_alloc;
            CUDA_TRY(h, cudaMemset(h->d_changed, 0, sizeof(unsigned long long)));
        }
    }
    if (h->dprog.nodes_per_root > 1 && !h->roots_uploaded)
        return fail(h, ECMC_ERR_STATE, "composite objects: ecmc_upload_roots before ecmc_run_from_host");
    if (h->molecules || h->disks)
        return fail(h, ECMC_ERR_INVALID, "molecules / general velocities: use ecmc_upload_positions / _roots, ecmc_start, ecmc_run");
    if (std::isnan(until_q) || std::isnan(until_r)) return fail(h, ECMC_ERR_INVALID, "until time is NaN");
    if (max_events_per_chain <= 0 && std::isinf(until_q))
        return fail(h, ECMC_ERR_INVALID, "neither a time limit nor an event limit: the run would not end");
    CUDA_TRY(h, cudaSetDevice(h->device));
    const DeviceProgram &d = h->dprog;
    int max_slices = 8;
    if (const char *env = std::getenv("ECMC_HOST_SLICES")) max_slices = std::max(1, std::atoi(env));
    const int slices = std::max(1, std::min(max_slices, h->n_chains / 256));
    while ((int)h->slice_streams.size() < slices) {
        cudaStream_t s;
        CUDA_TRY(h, cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking));
        h->slice_streams.push_back(s);
    }
    if (!h->slices_busy) CUDA_TRY(h, cudaStreamSynchronize(h->stream));  // earlier work on the handle's stream comes first
    RunArgs args{};
    args.until_q = until_q;
    args.until_r = until_r;
    args.max_events = max_events_per_chain;
    args.records = nullptr;
    args.records_per_chain = 0;
    args.stats = h->d_stats;
    args.list_capacity = 0;
    EventKernel kernel = pick_kernel(d, false);
    SpecLaunch spec;
    if (pick_spec(h, false, &spec)) {
        kernel = spec.kernel;
        args.list_capacity = spec.capacity;
        CUDA_TRY(h, cudaFuncSetAttribute(spec.kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)spec.shared_bytes));
    }
    const size_t per_chain = (size_t)d.n_particles;
    // Lennard-Jones / cell-veto programs with a sparse write-back: one launch per chain slice does the whole step -- the
    // kernel reads the configuration from the pinned buffer itself, bins it, runs the events and writes the positions it
    // changes through to the buffer (lj_spec_kernel<..., HOST>); nothing is staged, no copy engine is involved
    // ECMC_OPTION_CONTINUE_HOST_STEPS: from the second step on the chains continue (lifting state kept on the device)
    const bool keep_state = h->host_continue && h->started;
    args.keep_state = keep_state ? 1 : 0;
    bool fused = sparse && !charges && spec.kernel && !spec.chain_blocks && !spec.coulomb && h->host_fused;
    bool fused_copy = true;
    if (const char *env = std::getenv("ECMC_FUSED_ZEROCOPY")) fused_copy = std::atoi(env) == 0;
    if (fused) {
        cudaPointerAttributes attributes{};
        if (cudaPointerGetAttributes(&attributes, positions_in) != cudaSuccess || attributes.type != cudaMemoryTypeHost ||
            !attributes.devicePointer) {
            cudaGetLastError();
            fused = false;  // pageable input: the staged path copies it
        } else {
            args.host_in = static_cast<const double *>(attributes.devicePointer);
            if (fused_copy) args.host_in = h->d_staging;  // the copy engine brings the slice, the kernel reads the staging buffer
            args.host_out = mapped_out;
            args.first_stream = first_stream;
            args.initial_active = h->program.initial_active;
            args.initial_direction = h->program.initial_direction;
            args.host_writes = h->d_changed;
            kernel = pick_spec_host(h->spec_prune, h->spec_lanes);
            CUDA_TRY(h, cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)spec.shared_bytes));
        }
    }
    int fused_slices = 32;
    if (const char *env = std::getenv("ECMC_FUSED_SLICES")) fused_slices = std::max(1, std::atoi(env));
    const int base_slices = slices;
    if (fused) {
        const int wanted = std::max(1, std::min(fused_slices, h->n_chains / (2 * kWarpsPerBlock)));
        while ((int)h->slice_streams.size() < wanted) {
            cudaStream_t s;
            CUDA_TRY(h, cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking));
            h->slice_streams.push_back(s);
        }
    }
    const int n_slices = fused ? std::max(1, std::min(fused_slices, h->n_chains / (2 * kWarpsPerBlock))) : base_slices;
    // Steps in flight are ordered stream by stream: a step that cuts the chains differently waits for them
    const int layout = n_slices * 2 + (fused ? 1 : 0);
    if (h->slices_busy && layout != h->slices_layout)
        for (cudaStream_t s : h->slice_streams) CUDA_TRY(h, cudaStreamSynchronize(s));
    h->slices_layout = layout;
    // (whole CTAs per slice: a slice boundary inside a CTA would leave warps idle)
    const int ctas = (h->n_chains + kWarpsPerBlock - 1) / kWarpsPerBlock;
    const int base = fused ? 0 : h->n_chains / n_slices, extra = fused ? 0 : h->n_chains % n_slices;
    int first = 0;
    for (int k = 0; k < n_slices; k++) {
        int count = base + (k < extra ? 1 : 0);
        if (fused) {
            const int cta_first = (int)((long long)ctas * k / n_slices), cta_last = (int)((long long)ctas * (k + 1) / n_slices);
            count = std::min(h->n_chains, cta_last * kWarpsPerBlock) - cta_first * kWarpsPerBlock;
            if (count <= 0) continue;
        }
        cudaStream_t s = h->slice_streams[k];
        if (fused) {
            if (fused_copy) {
                const size_t offset = (size_t)first * per_chain * 3, n = (size_t)count * per_chain * 3;
                CUDA_TRY(h, cudaMemcpyAsync(h->d_staging + offset, positions_in + offset, n * sizeof(double),
                                            cudaMemcpyHostToDevice, s));
            }
            DeviceState slice = h->state;
            slice.first_chain = first;
            slice.n_chains = count;
            const int blocks = (count + kWarpsPerBlock - 1) / kWarpsPerBlock;
            kernel<<<blocks, kWarpsPerBlock * 32, spec.shared_bytes, s>>>(d, slice, args);
            CUDA_TRY(h, cudaGetLastError());
            h->kernel_launches++;
            first += count;
            continue;
        }
        const size_t offset = (size_t)first * per_chain, n = (size_t)count * per_chain;
        CUDA_TRY(h, cudaMemcpyAsync(h->d_staging + offset * d.dimension, positions_in + offset * d.dimension,
                                    n * d.dimension * sizeof(double), cudaMemcpyHostToDevice, s));
        if (charges)
            CUDA_TRY(h, cudaMemcpyAsync(h->d_staging_charges + offset, charges + offset, n * sizeof(double),
                                        cudaMemcpyHostToDevice, s));
        const int copy_blocks = (int)std::min<size_t>((n + 255) / 256, 148 * 4);
        pack_particles_kernel<<<copy_blocks, 256, 0, s>>>(h->d_staging + offset * d.dimension,
                                                          charges ? h->d_staging_charges + offset : nullptr,
                                                          h->state.particles + offset, n, d.dimension);
        DeviceState slice = h->state;
        slice.first_chain = first;
        slice.n_chains = count;
        const int blocks = (count + kWarpsPerBlock - 1) / kWarpsPerBlock;
        start_kernel<kWarpsPerBlock><<<blocks, kWarpsPerBlock * 32, 0, s>>>(
            d, slice, nullptr, first_stream, h->program.initial_active, h->program.initial_direction, h->d_stats, keep_state);
        if (spec.chain_blocks) kernel<<<count, kChainWarps * 32, spec.shared_bytes, s>>>(d, slice, args);
        else if (spec.kernel) kernel<<<(count + spec.warps - 1) / spec.warps, spec.warps * 32, spec.shared_bytes, s>>>(d, slice, args);
        else kernel<<<blocks, kWarpsPerBlock * 32, 0, s>>>(d, slice, args);
        CUDA_TRY(h, cudaGetLastError());
        h->kernel_launches++;
        if (sparse) {
            write_changed_particles_kernel<<<copy_blocks, 256, 0, s>>>(h->state.particles + offset,
                                                                       h->d_staging + offset * d.dimension,
                                                                       mapped_out + offset * d.dimension, n, d.dimension,
                                                                       h->d_changed);
            CUDA_TRY(h, cudaGetLastError());
        } else if (positions_out) {
            unpack_particles_kernel<<<copy_blocks, 256, 0, s>>>(h->state.particles + offset,
                                                                h->d_staging + offset * d.dimension, n, d.dimension);
            CUDA_TRY(h, cudaMemcpyAsync(positions_out + offset * d.dimension, h->d_staging + offset * d.dimension,
                                        n * d.dimension * sizeof(double), cudaMemcpyDeviceToHost, s));
        }
        first += count;
    }
    h->slices_busy = true;
    h->started = true;
    return ECMC_OK;
}

}  // namespace

ECMC_API int ecmc_submit_from_host(EcmcHandle *h, const double *positions_in, const double *charges, uint32_t first_stream,
                                   double until_q, double until_r, int64_t max_events_per_chain, double *positions_out) {
    return submit_from_host(h, positions_in, charges, first_stream, until_q, until_r, max_events_per_chain, positions_out, false);
}

ECMC_API int ecmc_submit_from_host_sparse(EcmcHandle *h, const double *positions_in, const double *charges,
                                          uint32_t first_stream, double until_q, double until_r,
                                          int64_t max_events_per_chain, double *positions_out) {
    return submit_from_host(h, positions_in, charges, first_stream, until_q, until_r, max_events_per_chain, positions_out, true);
}

ECMC_API int ecmc_wait(EcmcHandle *h, EcmcStats *stats) {
    if (!h) return fail(h, ECMC_ERR_INVALID, "null handle");
    CUDA_TRY(h, cudaSetDevice(h->device));
    for (cudaStream_t s : h->slice_streams) CUDA_TRY(h, cudaStreamSynchronize(s));
    h->slices_busy = false;
    if (h->d_changed) {
        unsigned long long changed = 0;
        CUDA_TRY(h, cudaMemcpy(&changed, h->d_changed, sizeof(changed), cudaMemcpyDeviceToHost));
        CUDA_TRY(h, cudaMemset(h->d_changed, 0, sizeof(changed)));
        h->host_bytes_written += changed * (uint64_t)h->dprog.dimension * sizeof(double);
    }
    return ecmc_sync(h, stats);
}

ECMC_API uint64_t ecmc_host_bytes_written(EcmcHandle *h) { return h ? h->host_bytes_written : 0; }

ECMC_API int ecmc_host_alloc(size_t bytes, void **out) {
    if (!out || bytes == 0) return ECMC_ERR_INVALID;
    void *p = nullptr;
    if (cudaHostAlloc(&p, bytes, cudaHostAllocPortable | cudaHostAllocMapped) != cudaSuccess) {
        cudaGetLastError();
        return ECMC_ERR_CUDA;
    }
    *out = p;
    return ECMC_OK;
}

ECMC_API int ecmc_host_free(void *p) { return cudaFreeHost(p) == cudaSuccess ? ECMC_OK : ECMC_ERR_CUDA; }

ECMC_API int ecmc_run_from_host(EcmcHandle *h, const double *positions_in, const double *charges, uint32_t first_stream,
                                double until_q, double until_r, int64_t max_events_per_chain, double *positions_out,
                                EcmcStats *stats) {
    const int rc = ecmc_submit_from_host(h, positions_in, charges, first_stream, until_q, until_r, max_events_per_chain,
                                         positions_out);
    return rc ? rc : ecmc_wait(h, stats);
}

ECMC_API int ecmc_separation_histogram(EcmcHandle *h, int32_t n_bins, double r_min, double r_max, uint64_t *histogram) {
    return ecmc_separation_histogram_subset(h, 0, 1, n_bins, r_min, r_max, histogram);
}

ECMC_API int ecmc_separation_histogram_subset(EcmcHandle *h, int32_t first, int32_t stride, int32_t n_bins, double r_min,
                                              double r_max, uint64_t *histogram) {
    if (!h || !histogram) return fail(h, ECMC_ERR_INVALID, "null argument");
    if (stride < 1 || first < 0 || first >= stride || first >= h->dprog.n_particles)
        return fail(h, ECMC_ERR_INVALID, "histogram subset needs 0 <= first < stride and first < n_particles");
    if (n_bins < 1 || n_bins > kHistogramMaxBins || !(r_max > r_min) || r_min < 0.0)
        return fail(h, ECMC_ERR_INVALID, "histogram needs 1 <= n_bins <= 4096 and 0 <= r_min < r_max");
    CUDA_TRY(h, cudaSetDevice(h->device));
    unsigned long long *d_histogram = nullptr;
    CUDA_TRY(h, cudaMalloc(&d_histogram, sizeof(unsigned long long) * n_bins));
    int rc = ECMC_OK;
    do {
        cudaError_t err = cudaMemsetAsync(d_histogram, 0, sizeof(unsigned long long) * n_bins, h->stream);
        if (err != cudaSuccess) { rc = fail(h, ECMC_ERR_CUDA, cudaGetErrorString(err)); break; }
        const int n_subset = (h->dprog.n_particles - first + stride - 1) / stride;
        const int n_tiles = (n_subset + kHistogramTile - 1) / kHistogramTile;
        separation_histogram_kernel<<<h->n_chains * n_tiles, 256, 0, h->stream>>>(
            h->state.particles, n_subset, n_tiles, h->dprog.length, n_bins, r_min,
            (double)n_bins / (r_max - r_min), d_histogram, first, stride, h->dprog.n_particles);
        std::vector<unsigned long long> counts(n_bins);
        if ((err = cudaGetLastError()) != cudaSuccess ||
            (err = cudaMemcpyAsync(counts.data(), d_histogram, sizeof(unsigned long long) * n_bins, cudaMemcpyDeviceToHost,
                                   h->stream)) != cudaSuccess ||
            (err = cudaStreamSynchronize(h->stream)) != cudaSuccess) {
            rc = fail(h, ECMC_ERR_CUDA, cudaGetErrorString(err));
            break;
        }
        for (int b = 0; b < n_bins; b++) histogram[b] += counts[b];
    } while (0);
    cudaFree(d_histogram);
    return rc;
}

ECMC_API int ecmc_polarization(EcmcHandle *h, const double *charges, double *polarization) {
    if (!h || !polarization) return fail(h, ECMC_ERR_INVALID, "null argument");
    const DeviceProgram &d = h->dprog;
    if (d.nodes_per_root < 2 || !h->state.roots) return fail(h, ECMC_ERR_INVALID, "polarization needs composite point objects");
    CUDA_TRY(h, cudaSetDevice(h->device));
    double *d_out = nullptr;
    const size_t bytes = sizeof(double) * (size_t)h->n_chains * d.dimension;
    const size_t charge_bytes = charges ? sizeof(double) * (size_t)d.n_particles : 0;
    CUDA_TRY(h, cudaMalloc(&d_out, bytes + charge_bytes));
    double *d_charges = charges ? d_out + (size_t)h->n_chains * d.dimension : nullptr;
    cudaError_t err = charges ? cudaMemcpyAsync(d_charges, charges, charge_bytes, cudaMemcpyHostToDevice, h->stream) : cudaSuccess;
    if (err == cudaSuccess) {
        polarization_kernel<<<(h->n_chains + 127) / 128, 128, 0, h->stream>>>(
            h->state.particles, h->state.roots, d_charges, h->n_chains, d.n_particles / d.nodes_per_root, d.nodes_per_root,
            d.dimension, d.length, d_out);
        err = cudaGetLastError();
    }
    if (err == cudaSuccess) err = cudaMemcpyAsync(polarization, d_out, bytes, cudaMemcpyDeviceToHost, h->stream);
    if (err == cudaSuccess) err = cudaStreamSynchronize(h->stream);
    cudaFree(d_out);
    if (err != cudaSuccess) return fail(h, ECMC_ERR_CUDA, cudaGetErrorString(err));
    return ECMC_OK;
}

ECMC_API int ecmc_bond_histograms(EcmcHandle *h, int32_t n_bins, double length_min, double length_max, double angle_min,
                                  double angle_max, uint64_t *length_histogram, uint64_t *angle_histogram) {
    if (!h || !length_histogram || !angle_histogram) return fail(h, ECMC_ERR_INVALID, "null argument");
    const DeviceProgram &d = h->dprog;
    if (d.nodes_per_root != 3 || d.dimension != 3) return fail(h, ECMC_ERR_INVALID, "bond histograms need objects of three leaves in 3D");
    if (n_bins < 1 || n_bins > kHistogramMaxBins || !(length_max > length_min) || !(angle_max > angle_min))
        return fail(h, ECMC_ERR_INVALID, "histograms need 1 <= n_bins <= 4096 and min < max");
    CUDA_TRY(h, cudaSetDevice(h->device));
    unsigned long long *d_histograms = nullptr;
    const size_t bytes = sizeof(unsigned long long) * 2 * (size_t)n_bins;
    CUDA_TRY(h, cudaMalloc(&d_histograms, bytes));
    std::vector<unsigned long long> counts(2 * (size_t)n_bins);
    const size_t n_objects = (size_t)h->n_chains * (d.n_particles / 3);
    cudaError_t err = cudaMemsetAsync(d_histograms, 0, bytes, h->stream);
    if (err == cudaSuccess) {
        bond_histogram_kernel<<<(int)std::min<size_t>((n_objects + 255) / 256, 148 * 8), 256, 0, h->stream>>>(
            h->state.particles, n_objects, d.length, n_bins, length_min, (double)n_bins / (length_max - length_min), angle_min,
            (double)n_bins / (angle_max - angle_min), d_histograms, d_histograms + n_bins);
        err = cudaGetLastError();
    }
    if (err == cudaSuccess) err = cudaMemcpyAsync(counts.data(), d_histograms, bytes, cudaMemcpyDeviceToHost, h->stream);
    if (err == cudaSuccess) err = cudaStreamSynchronize(h->stream);
    cudaFree(d_histograms);
    if (err != cudaSuccess) return fail(h, ECMC_ERR_CUDA, cudaGetErrorString(err));
    for (int b = 0; b < n_bins; b++) {
        length_histogram[b] += counts[b];
        angle_histogram[b] += counts[n_bins + b];
    }
    return ECMC_OK;
}

ECMC_API int ecmc_set_option(EcmcHandle *h, int option, int value) {
    if (!h) return fail(h, ECMC_ERR_INVALID, "null handle");
    switch (option) {
    case ECMC_OPTION_BATCHED_EVENTS: h->spec = value != 0; return ECMC_OK;
    case ECMC_OPTION_PRUNE_CANDIDATES: h->spec_prune = value != 0; return ECMC_OK;
    case ECMC_OPTION_CHAIN_BLOCKS: h->chain_blocks = value != 0; return ECMC_OK;
    case ECMC_OPTION_FUSED_HOST_STEPS: h->host_fused = value != 0; return ECMC_OK;
    case ECMC_OPTION_CONTINUE_HOST_STEPS: h->host_continue = value != 0; return ECMC_OK;
    case ECMC_OPTION_LANES_PER_EVENT:
        if (value != 4 && value != 8) return fail(h, ECMC_ERR_INVALID, "lanes per event: 4 or 8");
        h->spec_lanes = value;
        return ECMC_OK;
    default: return fail(h, ECMC_ERR_INVALID, "unknown option " + std::to_string(option));
    }
}

ECMC_API void *ecmc_stream(EcmcHandle *h) { return h ? (void *)h->stream : nullptr; }
ECMC_API double ecmc_kernel_seconds(EcmcHandle *h) { return h ? h->kernel_seconds : 0.0; }
ECMC_API uint64_t ecmc_kernel_launches(EcmcHandle *h) { return h ? h->kernel_launches : 0; }

ECMC_API const char *ecmc_kernel_name(EcmcHandle *h, int record) {
    if (!h) return "";
    const DeviceProgram &d = h->dprog;
    auto potential = [](int kind) -> std::string {
        switch (kind) {
        case 0: return "none";
        case ECMC_POT_LENNARD_JONES: return "LJ";
        case ECMC_POT_HARD_SPHERE: return "HS";
        case ECMC_POT_MERGED_IMAGE_COULOMB: return "MIC";
        case ECMC_POT_INVERSE_POWER_COULOMB_BOUNDING: return "IPCB";
        case ECMC_POT_INVERSE_POWER: return "IP";
        default: return "pot" + std::to_string(kind);
        }
    };
    const std::string cand = d.pair_handler == ECMC_PAIR_NONE ? "none" : potential(d.cand_potential.kind);
    const bool bounded = d.pair_handler == ECMC_PAIR_TWO_LEAF_UNIT_BOUNDING || d.pair_handler == ECMC_PAIR_TWO_COMPOSITE_SUMMED_BOUNDING;
    const std::string real = bounded ? potential(d.real_potential.kind) : "none";
    const std::string veto = d.veto_enabled ? potential(d.veto_potential.kind) : "none";
    SpecLaunch spec;
    if (h->disks) {
        h->kernel_name = "disk_kernel<record=" + std::to_string(record != 0) + ">";
    } else if (h->molecules) {
        h->kernel_name = "molecule_kernel<cand=" + cand + ", real=" + real + ", veto=" + veto + ", record=" + std::to_string(record != 0) +
                         (h->mprog.root_mode ? ", root mode>"
                          : (h->mprog.cell_child ? ", leaf cells>" : (molecule_cta_warps(h) == kWideWarps ? ", warps=16>" : ">")));
    } else if (pick_spec(h, record != 0, &spec) && spec.chain_blocks) {
        h->kernel_name = "lj_chain_kernel<record=" + std::to_string(record != 0) + ", prune=" +
                         std::to_string(h->spec_prune && !record) + ", warps per chain=" + std::to_string(kChainWarps) + ">";
    } else if (pick_spec(h, record != 0, &spec) && spec.coulomb) {
        h->kernel_name = "lj_spec_kernel<coulomb, record=" + std::to_string(record != 0) + ", prune=" +
                         std::to_string(h->spec_prune && !record) + ", lanes=" + std::to_string(h->spec_lanes) + ", warps=" +
                         std::to_string(spec.warps) + ">";
    } else if (pick_spec(h, record != 0, &spec)) {
        h->kernel_name = "lj_spec_kernel<record=" + std::to_string(record != 0) + ", prune=" +
                         std::to_string(h->spec_prune && !record) + ", lanes=" + std::to_string(h->spec_lanes) + ", warps=" +
                         std::to_string(kWarpsPerBlock) + ">";
    } else {
        h->kernel_name = "event_kernel<cand=" + cand + ", real=" + real + ", veto=" + veto +
                         (d.nodes_per_root > 1 ? ", composite" : "") +
                         (d.veto_enabled == ECMC_FAR_CELL_BOUNDING ? ", far_pairs" : "") + ", record=" +
                         std::to_string(record != 0) + ", warps=" + std::to_string(kWarpsPerBlock) + ">";
    }
    return h->kernel_name.c_str();
}

// ==========================================================================================================
// batched potential arithmetic
// ==========================================================================================================
namespace {

int potential_batch(bool displacement, const EcmcPotential *potential, int dimension, double system_length,
                    const double *velocity, size_t n, const double *separations, const double *charges,
                    const double *potential_changes, double *out, int device) {
    if (!potential || !velocity || !out || (n && !separations)) return fail(nullptr, ECMC_ERR_INVALID, "null argument");
    if (dimension < 1 || dimension > 3) return fail(nullptr, ECMC_ERR_INVALID, "dimension must be 1, 2 or 3");
    const bool hard = potential->kind == ECMC_POT_HARD_SPHERE || potential->kind == ECMC_POT_HARD_DIPOLE;
    if (displacement ? !is_invertible(potential->kind) : !has_derivative(potential->kind))
        return fail(nullptr, ECMC_ERR_INVALID, displacement ? "potential is not invertible" : "potential has no derivative");
    if ((potential->kind == ECMC_POT_MERGED_IMAGE_COULOMB || potential->kind == ECMC_POT_INVERSE_POWER_COULOMB_BOUNDING) &&
        dimension != 3)
        return fail(nullptr, ECMC_ERR_INVALID, "Coulomb potentials need dimension 3");
    // StandardVelocityPotential._analyse_velocity (potential/abstracts.py:105-140)
    int dir = -1;
    bool standard = true;
    for (int k = 0; k < dimension; k++)
        if (velocity[k] != 0.0) {
            if (dir >= 0) standard = false;
            dir = k;
        }
    if (dir < 0 || !(velocity[dir] > 0.0)) standard = false;
    if (!standard && !(hard && displacement))
        return fail(nullptr, ECMC_ERR_INVALID, "velocity must have exactly one non-zero, positive component");
    int count = 0;
    if (cudaGetDeviceCount(&count) != cudaSuccess || count == 0)
        return fail(nullptr, ECMC_ERR_CUDA, "no CUDA device (libecmc_b200 has no CPU fallback)");
    if (device < 0 || device >= count) return fail(nullptr, ECMC_ERR_INVALID, "device index out of range");
    if (n == 0) return ECMC_OK;
    EcmcHandle scratch;  // owns the temporary device allocations
    scratch.device = device;
    int rc = ECMC_OK;
    cudaError_t err;
    do {
        if ((err = cudaSetDevice(device)) != cudaSuccess) { rc = fail(nullptr, ECMC_ERR_CUDA, cudaGetErrorString(err)); break; }
        PotentialParams params;
        if ((rc = make_potential(&scratch, *potential, system_length, &params))) { g_create_error = scratch.error; break; }
        BatchArgs b;
        std::memset(&b, 0, sizeof(b));
        b.dimension = dimension;
        b.dir = standard ? dir : 0;
        b.speed = standard ? velocity[dir] : 1.0;
        b.length = system_length;
        for (int k = 0; k < 3; k++) b.velocity[k] = k < dimension ? velocity[k] : 0.0;
        b.n = n;
        double *d_sep = nullptr, *d_charges = nullptr, *d_du = nullptr, *d_out = nullptr;
        if ((rc = device_alloc(&scratch, &d_sep, n * dimension)) || (rc = device_alloc(&scratch, &d_out, n))) break;
        if (charges && (rc = device_alloc(&scratch, &d_charges, 2 * n))) break;
        if (potential_changes && displacement && (rc = device_alloc(&scratch, &d_du, n))) break;
        err = cudaMemcpy(d_sep, separations, n * dimension * sizeof(double), cudaMemcpyHostToDevice);
        if (err == cudaSuccess && d_charges) err = cudaMemcpy(d_charges, charges, 2 * n * sizeof(double), cudaMemcpyHostToDevice);
        if (err == cudaSuccess && d_du) err = cudaMemcpy(d_du, potential_changes, n * sizeof(double), cudaMemcpyHostToDevice);
        if (err != cudaSuccess) { rc = fail(nullptr, ECMC_ERR_CUDA, cudaGetErrorString(err)); break; }
        b.separations = d_sep;
        b.charges = d_charges;
        b.potential_changes = d_du;
        b.out = d_out;
        if (displacement) {
            const int blocks = (int)std::min<size_t>((n + 255) / 256, 148 * 8);
            displacement_batch_kernel<<<blocks, 256>>>(params, b);
        } else {
            const size_t work = params.kind == ECMC_POT_MERGED_IMAGE_COULOMB ? (n + 7) / 8 : (n + 255) / 256;
            const int blocks = (int)std::min<size_t>(work, 148 * 8);
            derivative_batch_kernel<<<blocks, 256>>>(params, b);
        }
        if ((err = cudaGetLastError()) != cudaSuccess || (err = cudaDeviceSynchronize()) != cudaSuccess ||
            (err = cudaMemcpy(out, d_out, n * sizeof(double), cudaMemcpyDeviceToHost)) != cudaSuccess) {
            rc = fail(nullptr, ECMC_ERR_CUDA, cudaGetErrorString(err));
            break;
        }
    } while (0);
    for (void *p : scratch.allocations) cudaFree(p);
    return rc;
}

}  // namespace

ECMC_API int ecmc_potential_derivative(const EcmcPotential *potential, int dimension, double system_length,
                                       const double *velocity, size_t n, const double *separations,
                                       const double *charges, double *out, int device) {
    return potential_batch(false, potential, dimension, system_length, velocity, n, separations, charges, nullptr, out, device);
}

ECMC_API int ecmc_potential_displacement(const EcmcPotential *potential, int dimension, double system_length,
                                         const double *velocity, size_t n, const double *separations,
                                         const double *charges, const double *potential_changes, double *out, int device) {
    return potential_batch(true, potential, dimension, system_length, velocity, n, separations, charges, potential_changes,
                           out, device);
}

// ==========================================================================================================
// the random stream on the host (same functions the kernels use)
// ==========================================================================================================
ECMC_API void ecmc_random_doubles(uint32_t seed, uint32_t stream, uint64_t event, uint32_t slot, uint32_t first, size_t n,
                                  double *out) {
    const StreamKey key = {seed, stream, event};
    for (size_t i = 0; i < n; i++) out[i] = stream_double(key, slot, first + (uint32_t)i);
}

ECMC_API void ecmc_random_words(uint32_t seed, uint32_t stream, uint64_t event, uint32_t slot, uint32_t first, size_t n,
                                uint32_t *out) {
    const StreamKey key = {seed, stream, event};
    for (size_t i = 0; i < n; i++) out[i] = stream_word(key, slot, first + (uint32_t)i);
}
