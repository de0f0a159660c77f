/* ecmc.h -- C ABI of libecmc_b200.so, the B200-native batched event-chain Monte Carlo engine.
 *
 * This is the drop-in boundary for ONE hot path of JeLLyFysh (reference paths are relative to the
 * reference checkout): the body of the mediator loop, jellyfysh/mediator/single_process_mediator.py:91-156,
 * i.e. for the active particle of a Markov chain: compute every candidate event time, take the argmin that
 * the scheduler would pop, apply the out-state (lifting) and commit it. The engine advances thousands of
 * independent chains at once on one GPU; everything else of the reference (factory, taggers' tag graph,
 * sampling / output handlers, end of run) stays on the host and calls in here.
 *
 * Conventions (same spirit as the reference's cffi modules, e.g.
 * jellyfysh/potential/merged_image_coulomb_potential/merged_image_coulomb_potential_build.py:36-44 and
 * jellyfysh/scheduler/heap_scheduler/heap_build.py:37-55): plain C, opaque handle owned by the caller,
 * caller-allocated output buffers, int status returns (0 = ok, <0 = error, message via ecmc_last_error),
 * no callbacks into the host language, no torch types, one handle = one GPU, one host thread per handle.
 * There is NO CPU fallback: every entry point that computes needs a CUDA device and fails otherwise.
 */
#ifndef ECMC_B200_H
#define ECMC_B200_H

#include <stdint.h>
#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ECMC_ABI_VERSION 6
#define ECMC_MAX_BONDS 4
#define ECMC_MAX_DIM 3

/* ---- status codes ---------------------------------------------------------------------------------- */
#define ECMC_OK 0
#define ECMC_ERR_INVALID (-1)   /* bad argument / unsupported program          */
#define ECMC_ERR_CUDA (-2)      /* CUDA runtime error (no device, OOM, launch) */
#define ECMC_ERR_STATE (-3)     /* call order violated (e.g. run before start) */
#define ECMC_ERR_CAPACITY (-4)  /* a device-side capacity (surplus list, occupants) overflowed */

/* ---- potentials -----------------------------------------------------------------------------------
 * Parameter carriers for the reference's potential classes. params[] meaning per kind:
 *  INVERSE_POWER            [0]=power [1]=prefactor          jellyfysh/potential/inverse_power_potential.py:32-179
 *  LENNARD_JONES            [0]=prefactor [1]=characteristic_length
 *                                                            jellyfysh/potential/lennard_jones_potential.py:31-135
 *  DISPLACED_EVEN_POWER     [0]=prefactor [1]=equilibrium_separation [2]=power
 *                                                            jellyfysh/potential/displaced_even_power_potential.py:73-142
 *  HARD_SPHERE              [0]=radius                       jellyfysh/potential/hard_sphere_potential.py:33-126
 *  HARD_DIPOLE              [0]=minimum_separation [1]=maximum_separation
 *                                                            jellyfysh/potential/hard_dipole_potential.py:33-141
 *  MERGED_IMAGE_COULOMB     [0]=prefactor [1]=alpha [2]=fourier_cutoff [3]=position_cutoff
 *                              jellyfysh/potential/merged_image_coulomb_potential/merged_image_coulomb_potential.c:77-274
 *  INVERSE_POWER_COULOMB_BOUNDING [0]=prefactor
 *                jellyfysh/potential/inverse_power_coulomb_bounding_potential/inverse_power_coulomb_bounding_potential.c:53-139
 *  BENDING                  [0]=prefactor [1]=equilibrium_angle (three-unit potential, derivative only)
 *                                                            jellyfysh/potential/bending_potential.py:60-138
 */
enum EcmcPotentialKind {
    ECMC_POT_NONE = 0,
    ECMC_POT_INVERSE_POWER = 1,
    ECMC_POT_LENNARD_JONES = 2,
    ECMC_POT_DISPLACED_EVEN_POWER = 3,
    ECMC_POT_HARD_SPHERE = 4,
    ECMC_POT_HARD_DIPOLE = 5,
    ECMC_POT_MERGED_IMAGE_COULOMB = 6,
    ECMC_POT_INVERSE_POWER_COULOMB_BOUNDING = 7,
    ECMC_POT_BENDING = 8
};

typedef struct EcmcPotential {
    int32_t kind;      /* EcmcPotentialKind */
    int32_t reserved;
    double params[6];
} EcmcPotential;

/* ---- pair event handlers --------------------------------------------------------------------------- */
enum EcmcPairHandlerKind {
    ECMC_PAIR_NONE = 0,
    /* TwoLeafUnitEventHandler: invertible potential, event always accepted.
     * jellyfysh/event_handler/two_leaf_unit_event_handler.py:105-154 */
    ECMC_PAIR_TWO_LEAF_UNIT = 1,
    /* TwoLeafUnitBoundingPotentialEventHandler: candidate from an invertible bounding potential, confirmed
     * against the real potential's derivative.
     * jellyfysh/event_handler/two_leaf_unit_bounding_potential_event_handler.py:112-168.
     * With cell_level = 1 (composite objects as the units of the cells / of a configuration without cells): one such
     * handler per pair (active leaf, leaf of ANOTHER object) -- the non-local factor type map entries "[0, 2], Coulomb"
     * ... of factor_set_dipoles_atomic.txt / factor_set_water_atomic.txt (dipoles/atom_factors.ini). */
    ECMC_PAIR_TWO_LEAF_UNIT_BOUNDING = 2,
    /* TwoCompositeObjectSummedBoundingPotentialEventHandler: the active leaf against every leaf of a composite object
     * in a nearby cell / the surplus; candidate = minimum over the target leaves of the bounding potential's
     * displacement, confirmed against the summed derivative, new active leaf from the lifting scheme.
     * jellyfysh/event_handler/two_composite_object_summed_bounding_potential_event_handler.py:119-202.
     * Needs cell_level = 1 (cells hold root units). With it the far field ECMC_FAR_CELL_VETO is the
     * CompositeObjectCellVetoEventHandler (composite_object_cell_veto_event_handler.py:110-162). */
    ECMC_PAIR_TWO_COMPOSITE_SUMMED_BOUNDING = 3
};

/* ---- lifting schemes (jellyfysh/lifting/{inside_first,outside_first,ratio}_lifting.py) ---------------------------- */
enum EcmcLiftingKind {
    ECMC_LIFTING_NONE = 0,
    ECMC_LIFTING_INSIDE_FIRST = 1,
    ECMC_LIFTING_OUTSIDE_FIRST = 2,
    ECMC_LIFTING_RATIO = 3
};
#define ECMC_MAX_INTER_FACTORS 4

/* ---- far-field handlers ---------------------------------------------------------------------------------------- */
enum EcmcFarFieldKind {
    ECMC_FAR_NONE = 0,
    /* CellVetoTagger + LeafUnitCellVetoEventHandler, jellyfysh/event_handler/leaf_unit_cell_veto_event_handler.py:117-149 */
    ECMC_FAR_CELL_VETO = 1,
    /* CellBoundingPotentialTagger (jellyfysh/activator/tagger/cell_bounding_potential_tagger.py:127-155) +
     * TwoLeafUnitCellBoundingPotentialEventHandler (two_leaf_unit_cell_bounding_potential_event_handler.py:137-211) with
     * CellBoundingPotential (jellyfysh/potential/cell_bounding_potential.py:155-238) */
    ECMC_FAR_CELL_BOUNDING = 2
};

/* ---- cell-veto tables -------------------------------------------------------------------------------
 * Walker alias table for one direction of motion and one sign, as built by the reference at init
 * (jellyfysh/event_handler/walker.py:69-103; jellyfysh/event_handler/abstracts/cell_veto_event_handler.py:134-159).
 * Entry e is (cell_a[e] with rate_a[e], cell_b[e]); cell_b[e] < 0 for single-item entries. Cells are flat
 * indices sum_d id_d * prod_{d'<d} cells_per_side[d'] of the cell *relative to cell zero*. */
typedef struct EcmcWalkerTable {
    int32_t n_entries;
    int32_t reserved;
    const int32_t *cell_a;
    const int32_t *cell_b;
    const double *rate_a;
    double total_rate;
    double mean_rate;
} EcmcWalkerTable;

typedef struct EcmcVetoTables {
    EcmcWalkerTable upper[ECMC_MAX_DIM]; /* per direction of motion; used when the charge factor is > 0 */
    EcmcWalkerTable lower[ECMC_MAX_DIM]; /* used when the charge factor is < 0 (n_entries may be 0 if unused) */
    /* derivative bounds [n_cells][dimension][2] = (upper_bound, -lower_bound) per relative cell; entries of
     * nearby (excluded) cells are ignored. */
    const double *bounds;
} EcmcVetoTables;

/* ---- the program: one configuration of the hot path, shared by all chains of a handle ---------------- */
typedef struct EcmcProgram {
    int32_t abi_version;      /* must be ECMC_ABI_VERSION */
    int32_t dimension;        /* 2 or 3; HypercubicSetting, jellyfysh/setting/hypercubic_setting.py:200-223 */
    int32_t n_particles;      /* point masses (leaf units) per chain */
    int32_t no_cells;         /* 1: the configuration has no cell system (no internal state in the TagActivator; the pair
                               * factors come from a FactorTypeMapInStateTagger, factor_type_map_in_state_tagger.py:83-107):
                               * every other unit is a candidate of every event and there are no cell-boundary events.
                               * Requires cells_per_side = 1, neighbor_layers = 0, max_surplus >= units - 1, no far field. */
    double system_length;
    double beta;
    /* CuboidPeriodicCells + SingleActiveCellOccupancy */
    int32_t cells_per_side[ECMC_MAX_DIM];
    int32_t neighbor_layers;
    int32_t max_occupants;    /* occupants stored per cell (reference maximum_number_occupants; >= 1) */
    int32_t max_surplus;      /* capacity of the per-chain surplus list */
    /* nearby + surplus pair factor */
    int32_t pair_handler;     /* EcmcPairHandlerKind */
    int32_t pair_use_charge;  /* 1: potentials get (charge_active, charge_target); 0: (1.0, 1.0) / none */
    EcmcPotential pair_potential;
    EcmcPotential pair_bounding_potential; /* only for ECMC_PAIR_TWO_LEAF_UNIT_BOUNDING */
    /* The far field, all non-nearby cells: EcmcFarFieldKind. ECMC_FAR_CELL_VETO = LeafUnitCellVetoEventHandler (one
     * candidate per event from the Walker tables); ECMC_FAR_CELL_BOUNDING = TwoLeafUnitCellBoundingPotentialEventHandler
     * (one candidate per occupied non-nearby cell from the constant bound of its relative cell; needs
     * max_occupants == 1 like the reference, uses only veto_tables->bounds). Both confirm against veto_potential. */
    int32_t veto_enabled;
    int32_t veto_use_charge;
    EcmcPotential veto_potential;
    double veto_target_charge; /* InnerPointEstimator target charge (charge_correction_factor) */
    const EcmcVetoTables *veto_tables;
    /* chain control: SingleIndependentActivePeriodicDirectionEndOfChainEventHandler +
     * InitialChainStartOfRunEventHandler */
    double chain_time;
    double speed;
    int32_t initial_direction;
    int32_t initial_active;
    /* counter-based random stream, see DESIGN.md "Random stream" */
    uint32_t seed;
    uint32_t reserved1;
    /* ---- composite point objects (dipoles, molecules): TreeStateHandler with two node levels,
     * jellyfysh/state_handler/tree_state_handler.py:104-230. The n_particles leaf units are grouped into
     * n_particles / nodes_per_root root units; leaf (root r, child k) has the flat identifier r * nodes_per_root + k.
     * The root unit of the active leaf moves with velocity * weight, weight = 1 / nodes_per_root
     * (event_handler/abstracts/abstracts.py:165-227), its position is kept alongside (ecmc_upload_roots).
     * nodes_per_root <= 1: point masses only (single node level). Cells hold leaf identifiers (cell_level = 2). */
    int32_t nodes_per_root;
    /* Intramolecular two-leaf factors from a factor type map (FactorTypeMapInStateTagger,
     * jellyfysh/activator/tagger/factor_type_map_in_state_tagger.py:83-107, e.g. "[0, 1], Dipole" of
     * config_files/factor_set_files/factor_set_hard_disk_dipoles.txt): pairs of child indices (a, b) of the same root,
     * each handled by a TwoLeafUnitEventHandler with bond_potential (hard dipole tether, harmonic bond). */
    int32_t n_bonds;
    int32_t bonds[ECMC_MAX_BONDS][2];
    EcmcPotential bond_potential;
    /* ---- molecules in root-level cells (water, C4) ------------------------------------------------------------
     * cell_level: 1 = the cells hold root units (SingleActiveCellOccupancy cell_level = 1 with two node levels: the
     * cell-boundary handler follows the root unit, which moves with speed / nodes_per_root), 0 or 2 = leaf units. */
    int32_t cell_level;
    int32_t composite_lifting;   /* EcmcLiftingKind of the composite-object handlers (pair + cell veto) */
    /* Two-leaf factors between DIFFERENT composite objects from a non-local factor type map entry, e.g.
     * "[1, 4], LennardJones" of factor_set_water.txt: the active leaf with child index a against child b of every
     * other object (factor_type_maps.py:333-347), each by a TwoLeafUnitEventHandler with inter_potential. */
    int32_t n_inter_factors;
    int32_t inter_factors[ECMC_MAX_INTER_FACTORS][2];
    EcmcPotential inter_potential;
    /* Three-leaf bending factor inside an object (FixedSeparationsEventHandlerWithPiecewiseConstantBoundingPotential,
     * jellyfysh/event_handler/fixed_separations_event_handler_with_piecewise_constant_bounding_potential.py:113-187):
     * children bending_children[0..2] in factor order, separations r[s[1]] - r[s[0]] and r[s[3]] - r[s[2]] (indices into
     * the three units), bounding rate = max(derivative now, derivative after max_displacement) + offset. */
    int32_t bending_enabled;
    int32_t bending_lifting;     /* EcmcLiftingKind */
    int32_t bending_children[3];
    int32_t bending_separations[4];
    /* 1: a cell-boundary event of the root unit does not trash the leaf-level factor handlers (bonds, inter-object
     * factors, bending), which then keep their candidates -- the tag lists of the shipped water configurations */
    int32_t boundary_keeps_factors;
    EcmcPotential bending_potential;
    double bending_offset;
    double bending_max_displacement;
    /* ---- general velocities: two-dimensional composite point objects without a cell system whose end-of-chain handler
     * is SingleIndependentActiveSequentialDirectionEndOfChainEventHandler
     * (single_independent_active_sequential_direction_end_of_chain_event_handler.py:64-122, the shipped
     * hard_disk_dipoles/hard_disk_dipoles.ini): every end of chain rotates the velocity by delta_phi,
     * (vx, vy) -> (vx cos - vy sin, vx sin + vy cos) with eoc_cos = cos(delta_phi), eoc_sin = sin(delta_phi) as the
     * handler computes them. The pair factors of such a program are the bonds (inside an object) and the inter_factors
     * (leaf a of the active object against leaf b of every other object), all with hard potentials; the velocity of
     * the active leaf and of its root unit live in EcmcChainState.velocity / root_velocity. */
    int32_t eoc_sequential;
    /* ---- root-unit-active mode (the shipped dipoles/dipole_motion.ini): composite point objects of TWO leaves without a
     * cell system whose independent active unit alternates between a leaf unit and the ROOT unit of an object -- the whole
     * object then moves, root and leaves with the full velocity. RootLeafUnitActiveSwitcher
     * (root_leaf_unit_active_switcher.py:102-228) switches: switch_chain_length[0] after the root unit handed over to a
     * leaf (or after the start of the run) the root unit of the active leaf takes over (aim_mode root_unit_active),
     * switch_chain_length[1] later one of its leaves, drawn by random.choice, takes over again (aim_mode
     * leaf_unit_active); every switch re-creates the end-of-chain candidate. While a root unit is active the factors are
     * those of the root-unit-active handlers with the potentials of the leaf mode: per other object one
     * RootUnitActiveTwoCompositeObjectSummedBoundingPotentialEventHandler
     * (root_unit_active_two_composite_object_summed_bounding_potential_event_handler.py:116-190: minimum over ALL pairs
     * (local leaf, target leaf) of the bounding potential's displacement, confirmed against the summed derivatives, the
     * target object takes over as a whole) with pair_potential / pair_bounding_potential, and per inter_factors entry
     * (a, b) and other object one RootUnitActiveTwoLeafUnitEventHandler (root_unit_active_two_leaf_unit_event_handler.py:
     * 72-125: local leaf a against leaf b of the other object, which then takes over as a whole) with inter_potential.
     * Requires nodes_per_root = 2, no_cells, pair_handler = ECMC_PAIR_TWO_COMPOSITE_SUMMED_BOUNDING. */
    int32_t root_mode;
    double eoc_cos, eoc_sin;
    double switch_chain_length[2];
    /* ---- a cell system for ONE kind of leaf only (the shipped water/coulomb_power_bounded_lj_cell_bounded.ini):
     * SingleActiveCellOccupancy(cell_level = 2, charge = <indicator>) stores only the leaves with child index
     * cell_child - 1 of every object (the oxygens; single_active_cell_occupancy.py:62-121) in cells_per_side /
     * neighbor_layers / max_occupants cells, and has an active cell only while such a leaf is active. 0: no such system.
     * With it (requires cell_level = 1, pair_handler = ECMC_PAIR_TWO_COMPOSITE_SUMMED_BOUNDING, no_cells = 0):
     *  - the composite-object pair factors come from the factor type map: every other object is a candidate of every event;
     *  - while a leaf of that kind is active, the two-leaf factor inter_factors[0] = (cell_child - 1, cell_child - 1) with the
     *    same leaf of the objects in nearby cells and of the surplus is handled by
     *    TwoLeafUnitEventHandlerWithPiecewiseConstantBoundingPotential
     *    (two_leaf_unit_event_handler_with_piecewise_constant_bounding_potential.py:103-162: constant bound = max(derivative
     *    now, derivative after inter_bound_max_displacement) + inter_bound_offset, confirmed against inter_potential), with
     *    those in all other cells by TwoLeafUnitCellBoundingPotentialEventHandler (veto_enabled = ECMC_FAR_CELL_BOUNDING,
     *    veto_potential, veto_tables->bounds), and the cell boundary is that of the active LEAF;
     *  - a cell-boundary event re-creates only the candidates found through the cells; the composite pairs, bonds and the
     *    bending factor keep theirs (boundary_keeps_factors = 1, EcmcChainState.kept_*).
     * Occupants, surplus and pending / kept targets of such a program are LEAF identifiers. */
    int32_t cell_child;
    int32_t reserved3;
    double inter_bound_offset;
    double inter_bound_max_displacement;
} EcmcProgram;

/* ---- per-chain lifting state ("who is active, where is the clock") ------------------------------------
 * Replaces TreeLiftingState (jellyfysh/state_handler/lifting_state/tree_lifting_state.py:56-147) plus the
 * persistent end-of-chain candidate and a candidate that survived a host control event (sampling). */
typedef struct EcmcChainState {
    int32_t active;           /* identifier of the active particle */
    int32_t direction;        /* direction of motion */
    double time_q, time_r;    /* time stamp of the active particle as (quotient, remainder), base/time.py */
    double eoc_q, eoc_r;      /* scheduled end-of-chain event time */
    int32_t eoc_next_active;  /* leaf drawn (randint; root then child for composite objects) when the end-of-chain
                               * candidate was created */
    int32_t active_cell;      /* flat cell index the occupancy system attributes to the active particle */
    uint64_t event_counter;   /* events committed so far: the random-stream event index */
    uint32_t stream;          /* random-stream id of this chain (Philox key word 0) */
    int32_t pending_kind;     /* EcmcEventKind of a candidate kept across a host control event, or 0 */
    int32_t pending_target;   /* pair / bond: target id; veto, boundary: target cell */
    int32_t mode;             /* EcmcProgram.root_mode: 1 while the ROOT unit of `active`'s object is the independent active
                               * unit (`active` is then the first leaf of that object), 0 while the leaf `active` is */
    double pending_q, pending_r;
    double pending_rate;      /* bounding event rate stored by the handler for the confirmation step */
    /* The kept handler computes its out-state from the in-state it was given BEFORE the control event
     * time-sliced the active particle: coordinate along the direction of motion and time stamp at that moment. */
    double pending_position;
    double pending_stamp_q, pending_stamp_r;
    double pending_root_position; /* the same for the root unit of the active leaf (composite objects) */
    /* Molecules (cell_level = 1) with EcmcProgram.boundary_keeps_factors: the earliest candidate of the leaf-level
     * factors (bonds, inter-object factors, bending) when a cell-boundary event of the root unit left them running
     * -- their handlers are not trashed by that event (e.g. [MoleculeCellBoundary] of the shipped water
     * configurations), so they keep their event times and in-states, and consume no new draws. kept_kind = 0: nothing
     * kept (all factors are recomputed); -1: kept, but no factor candidate was finite. */
    int32_t kept_kind;
    int32_t kept_target;
    double kept_q, kept_r;
    double kept_rate;
    double kept_position, kept_root_position;
    double kept_stamp_q, kept_stamp_r;
    /* Programs with general velocities (EcmcProgram.eoc_sequential): velocity of the active leaf unit and of its root
     * unit -- the root's is kept as the reference accumulates it (velocity changes times the weight,
     * event_handler/abstracts/abstracts.py:165-227), not recomputed from the leaf's --, and the second coordinates of
     * the in-state of a kept candidate (pending_position / pending_root_position hold the first ones). */
    double velocity[2];
    double root_velocity[2];
    double pending_position_y, pending_root_position_y;
    /* EcmcProgram.root_mode: the candidate event time of the RootLeafUnitActiveSwitcher that is running (time stamp of the
     * root unit when it was created + its chain length) and the time of the last committed end of chain (the
     * _last_committed_event_time that a re-created end-of-chain candidate continues from,
     * single_independent_active_periodic_direction_end_of_chain_event_handler.py:203-215). While a root unit is active,
     * pending_position / pending_position_y hold the in-state coordinates of its first / second leaf. */
    double switch_q, switch_r;
    double eoc_last_q, eoc_last_r;
} EcmcChainState;

enum EcmcEventKind {
    ECMC_EVENT_NONE = 0,
    ECMC_EVENT_PAIR = 1,          /* nearby-cell or surplus pair factor */
    ECMC_EVENT_CELL_VETO = 2,
    ECMC_EVENT_CELL_BOUNDARY = 3,
    ECMC_EVENT_END_OF_CHAIN = 4,
    ECMC_EVENT_CELL_BOUNDING = 5, /* pair factor of a non-nearby cell, found through the cell-bounding potential */
    ECMC_EVENT_BOND = 6,          /* intramolecular two-leaf factor of the factor type map (EcmcProgram.bonds) */
    ECMC_EVENT_FACTOR_PAIR = 7,   /* two-leaf factor between different objects (EcmcProgram.inter_factors) */
    ECMC_EVENT_BENDING = 8,       /* three-leaf bending factor */
    ECMC_EVENT_SWITCH = 9         /* RootLeafUnitActiveSwitcher: the root unit / a leaf unit of the same object takes over */
};

/* One committed event, as the scheduler + winning handler of the reference would report it. */
typedef struct EcmcEventRecord {
    int32_t kind;             /* EcmcEventKind of the winner (argmin) */
    int32_t target;           /* pair / cell bounding / accepted or rejected veto: target particle (-1: empty veto cell);
                               * composite-object pair and veto: the target ROOT unit */
    int32_t target_cell;      /* veto: sampled target cell; boundary: new cell; cell bounding: relative cell; else -1 */
    int32_t accepted;         /* 1 if the velocity was handed over (lifting happened) */
    int32_t n_candidates;     /* finite candidate times that entered the argmin */
    int32_t new_active;       /* active particle after the event */
    int32_t new_direction;
    int32_t mode;             /* EcmcChainState.mode after the event (0 unless the program has a root-unit-active mode) */
    double time_q, time_r;    /* event time */
    double active_pos[ECMC_MAX_DIM]; /* position of the (old) active particle after the event */
} EcmcEventRecord;

typedef struct EcmcStats {
    uint64_t events;            /* events committed by this call, all chains */
    uint64_t pair_events;       /* nearby / surplus pair events and cell-bounding pair events */
    uint64_t veto_events;
    uint64_t veto_accepted;
    uint64_t boundary_events;
    uint64_t end_of_chain_events;
    uint64_t candidates;        /* finite candidates that entered an argmin */
    uint64_t bound_violations;  /* real derivative exceeded its bound (reference: bounding_potential_warning) */
    uint64_t capacity_errors;   /* surplus / occupant overflow (fatal: results invalid) */
    uint64_t bond_events;       /* events of the intramolecular factor-type-map factors (bonds + bending) */
    uint64_t factor_pair_events; /* events of the inter-object two-leaf factors */
    uint64_t pair_targets;      /* pair targets gathered from nearby cells / surplus / far cells (event_kernel): the
                                 * handler calls the reference would make, finite or not */
} EcmcStats;

typedef struct EcmcHandle EcmcHandle;

/* ---- lifecycle ---------------------------------------------------------------------------------------- */
/* Build an engine for n_chains independent chains on CUDA device `device`. Copies the program and its tables. */
int ecmc_create(const EcmcProgram *program, int device, int n_chains, EcmcHandle **out);
void ecmc_destroy(EcmcHandle *h);
const char *ecmc_last_error(const EcmcHandle *h); /* h may be NULL: error of the last failed ecmc_create */
int ecmc_abi_version(void);

/* ---- state (replaces TreeStateHandler extract/insert, jellyfysh/state_handler/tree_state_handler.py:104-230) */
/* positions: [n_chains][n_particles][dimension] doubles; charges: [n_chains][n_particles] or NULL (all 1.0). */
int ecmc_upload_positions(EcmcHandle *h, const double *positions, const double *charges);
int ecmc_download_positions(EcmcHandle *h, double *positions);
/* One chain only -- what an output handler or a dump of a single chain needs (the four methods of the reference's state
 * handler contract, jellyfysh/state_handler/state_handler.py:63-165, read one global state): positions
 * [n_particles][dimension]; roots [n_particles / nodes_per_root][dimension] or NULL; state (lifting state) or NULL. */
int ecmc_download_chain(EcmcHandle *h, int chain, double *positions, double *roots, EcmcChainState *state);
/* Root-unit positions of composite objects, [n_chains][n_particles / nodes_per_root][dimension] (programs with
 * nodes_per_root > 1 only; ECMC_ERR_INVALID otherwise). Must be uploaded before ecmc_start. */
int ecmc_upload_roots(EcmcHandle *h, const double *roots);
int ecmc_download_roots(EcmcHandle *h, double *roots);
/* InitialChainStartOfRunEventHandler (initial_chain_start_of_run_event_handler.py:92-131) +
 * SingleActiveCellOccupancy.initialize (single_active_cell_occupancy.py:95-121): bins all particles,
 * activates program.initial_active at time 0 and schedules the first end-of-chain event.
 * streams: n_chains random-stream ids, or NULL for first_stream + chain index. */
int ecmc_start(EcmcHandle *h, const uint32_t *streams, uint32_t first_stream);
/* explicit lifting / cell state, for resuming and for event-by-event parity checks */
int ecmc_upload_chain_states(EcmcHandle *h, const EcmcChainState *states);
int ecmc_download_chain_states(EcmcHandle *h, EcmcChainState *states);
/* occupants: [n_chains][n_cells][max_occupants] (-1 = empty); surplus: [n_chains][max_surplus];
 * n_surplus: [n_chains] */
int ecmc_upload_cells(EcmcHandle *h, const int32_t *occupants, const int32_t *surplus, const int32_t *n_surplus);
int ecmc_download_cells(EcmcHandle *h, int32_t *occupants, int32_t *surplus, int32_t *n_surplus);

/* ---- the hot path -------------------------------------------------------------------------------------- */
/* Advance every chain until its next event time would reach (until_q, until_r) or it has committed
 * max_events_per_chain events in this call, whichever is first (max_events_per_chain <= 0: no event limit;
 * until_q = +inf: no time limit). Chains stopped by the time limit are time-sliced to it, as a sampling
 * handler does (jellyfysh/event_handler/fixed_interval_sampling_event_handler.py:96-109).
 * Asynchronous on the handle's stream; stats (may be NULL) are valid after ecmc_sync. */
int ecmc_run(EcmcHandle *h, double until_q, double until_r, int64_t max_events_per_chain);
int ecmc_sync(EcmcHandle *h, EcmcStats *stats);
/* Same, but additionally writes the first `records_per_chain` events of every chain of this call to
 * records[n_chains][records_per_chain] (host buffer, filled at return; kind == ECMC_EVENT_NONE marks unused
 * slots). Synchronous. Used by the parity tests. */
int ecmc_run_recorded(EcmcHandle *h, double until_q, double until_r, int64_t max_events_per_chain,
                      EcmcEventRecord *records, int32_t records_per_chain, EcmcStats *stats);
/* Host-buffer convenience for the plugin layer: upload positions -> start -> run -> download, one call. */
int ecmc_run_from_host(EcmcHandle *h, const double *positions_in, const double *charges, uint32_t first_stream,
                       double until_q, double until_r, int64_t max_events_per_chain, double *positions_out,
                       EcmcStats *stats);
/* The same step without blocking: ecmc_submit_from_host enqueues the copies and kernels of one step and returns,
 * ecmc_wait blocks until every submitted step has completed and returns the counters summed over them. Steps submitted
 * back to back are ordered chain slice by chain slice (stream order), so step k + 1 may read the buffer step k writes
 * (positions_in of k + 1 == positions_out of k), and the host copies of one slice overlap the events of the others across
 * steps. The host buffers must be page-locked and stay valid until ecmc_wait; no other call on the handle in between. */
int ecmc_submit_from_host(EcmcHandle *h, const double *positions_in, const double *charges, uint32_t first_stream,
                          double until_q, double until_r, int64_t max_events_per_chain, double *positions_out);
int ecmc_wait(EcmcHandle *h, EcmcStats *stats);
/* ecmc_submit_from_host with a SPARSE write-back: positions_out must hold the input configuration of the step wherever it
 * does not change (the intended use is positions_out == positions_in, the buffer the next step reads) and must be
 * page-locked host memory the device can address (ecmc_host_alloc, cudaHostAlloc, cudaHostRegister). Only the
 * coordinates of the particles that moved during the step -- the units that were active, tree_state_handler.py:187-211
 * writes back exactly those -- are written, by the device, straight into the buffer: about a tenth of the bytes of the
 * full copy for a step of 1024 events on 1024 particles, and no device-to-host copy of the rest. After ecmc_wait the
 * buffer is the complete configuration. ecmc_host_bytes_written: bytes written this way since ecmc_create. */
int ecmc_submit_from_host_sparse(EcmcHandle *h, const double *positions_in, const double *charges, uint32_t first_stream,
                                 double until_q, double until_r, int64_t max_events_per_chain, double *positions_out);
uint64_t ecmc_host_bytes_written(EcmcHandle *h);
/* Page-locked, device-addressable host memory for the buffers of the *_from_host calls (cudaHostAlloc). */
int ecmc_host_alloc(size_t bytes, void **out);
int ecmc_host_free(void *p);
/* ---- observables --------------------------------------------------------------------------------------------
 * Histogram of the shortest pair separations |r_ij| (all pairs i < j of every chain) into n_bins equal bins on
 * [r_min, r_max]: what SeparationOutputHandler.write prints sample by sample
 * (jellyfysh/input_output_handler/output_handler/separation_output_handler.py:75-97) and
 * jellyfysh/output/plotting_functions.py histograms afterwards, accumulated on the device over all chains at once.
 * Counts are ADDED to histogram[n_bins] (host buffer), so successive sampling times accumulate; separations outside
 * the range are not counted. Ranks of a multi-GPU run sum their histograms with one all-reduce (sharding.py). */
int ecmc_separation_histogram(EcmcHandle *h, int32_t n_bins, double r_min, double r_max, uint64_t *histogram);
/* The same over the particles first, first + stride, first + 2 stride, ... only: with stride = nodes_per_root and
 * first = a child index, the separations between one kind of leaf of every composite object -- the oxygen-oxygen
 * separations of water that OxygenOxygenSeparationOutputHandler samples
 * (jellyfysh/input_output_handler/output_handler/oxygen_oxygen_separation_output_handler.py). */
int ecmc_separation_histogram_subset(EcmcHandle *h, int32_t first, int32_t stride, int32_t n_bins, double r_min,
                                     double r_max, uint64_t *histogram);

/* Polarization of every chain: sum over all leaf units of charge x (leaf position closest to its root unit), what
 * PolarizationOutputHandler.write prints per sample (jellyfysh/input_output_handler/output_handler/
 * polarization_output_handler.py:76-101, base/node.py:164-188), same order of additions. charges: [n_particles] (the
 * output handler's own charge of every leaf, the same in all chains) or NULL for the charges uploaded with the positions.
 * polarization: [n_chains][dimension] (host buffer, overwritten). Composite point objects only. */
int ecmc_polarization(EcmcHandle *h, const double *charges, double *polarization);
/* Bond lengths |r_H - r_O| (both bonds) and bond angles of every three-leaf object (child 1 = the centre) of every chain,
 * what BondLengthAndAngleOutputHandler.write prints (bond_length_and_angle_output_handler.py:77-103), ADDED to two
 * histograms of n_bins equal bins on [length_min, length_max] / [angle_min, angle_max] (host buffers). */
int ecmc_bond_histograms(EcmcHandle *h, int32_t n_bins, double length_min, double length_max, double angle_min,
                         double angle_max, uint64_t *length_histogram, uint64_t *angle_histogram);

/* ---- execution options (no reference counterpart: how the device schedules the same events) ------------------
 * The Lennard-Jones / cell-veto configurations (chargeless 3D Lennard-Jones pair factors in the nearby cells and the
 * surplus, Lennard-Jones cell veto, one occupant per cell) are advanced by a kernel that evaluates several successive
 * events of a chain side by side under the assumption that they are rejected cell vetoes, and commits them up to the
 * first event that is not (csrc/ecmc_spec.cuh). The committed events -- winner, target, lifting, times, positions --
 * are those of the one-event-at-a-time loop (single_process_mediator.py:91-156), event for event. The Coulomb atoms
 * (pair candidates from the inverse-power Coulomb bound confirmed against merged-image Coulomb, merged-image Coulomb cell
 * veto, one occupant per cell, with charges) run the same batch with their potentials (csrc/ecmc_spec.cuh, kSpecCoulomb).
 *   ECMC_OPTION_BATCHED_EVENTS    1 (default) / 0: fall back to the one-event-at-a-time kernel
 *   ECMC_OPTION_PRUNE_CANDIDATES  1 (default) / 0: in ecmc_run / ecmc_run_from_host a pair candidate that provably cannot
 *                                 precede the cell-veto / cell-boundary candidate of its event is not inverted (its
 *                                 potential change is bounded from below by the uniform it is drawn from, its energy
 *                                 rise by the largest force on its line). Winners are unchanged; EcmcStats.candidates
 *                                 then counts the evaluated finite candidates only. ecmc_run_recorded never prunes.
 *   ECMC_OPTION_LANES_PER_EVENT   4 (default: 8 events per batch) or 8 (4 events per batch)
 *   ECMC_OPTION_CHAIN_BLOCKS      1 (default) / 0: engines of at most 148 chains (the single large chain of the metric's
 *                                 "ns/event single chain") give every chain a CTA of four warps that evaluates 32 events
 *                                 per batch (csrc/ecmc_spec_cta.cuh) instead of one warp with 8; same committed events
 *   ECMC_OPTION_FUSED_HOST_STEPS  1 (default) / 0: ecmc_submit_from_host_sparse of these programs runs a host step as ONE
 *                                 launch per chain slice -- the kernel reads the start configuration from the pinned
 *                                 buffer itself, bins it into the cells, runs the events and writes every position it
 *                                 changes through to the buffer -- instead of copy, pack, start, events, write-back
 *   ECMC_OPTION_CONTINUE_HOST_STEPS 0 (default) / 1: a step of ecmc_run_from_host / ecmc_submit_from_host[_sparse] on a
 *                                 handle that has run before CONTINUES its chains -- active unit, direction, clock, end of
 *                                 chain, event counter and random stream stay what the last launch left on the device,
 *                                 only the configuration comes from the host and the cell occupancy is rebuilt from it
 *                                 (first_stream is then ignored) -- instead of a start of run per step
 *                                 (initial_chain_start_of_run_event_handler.py:92-131). While no cell holds two
 *                                 particles the events are those of an uninterrupted ecmc_run, bit for bit. */
enum EcmcOption { ECMC_OPTION_BATCHED_EVENTS = 1, ECMC_OPTION_PRUNE_CANDIDATES = 2, ECMC_OPTION_LANES_PER_EVENT = 3,
                  ECMC_OPTION_CHAIN_BLOCKS = 4, ECMC_OPTION_FUSED_HOST_STEPS = 5, ECMC_OPTION_CONTINUE_HOST_STEPS = 6 };
int ecmc_set_option(EcmcHandle *h, int option, int value);

/* The CUDA stream the handle launches on (a cudaStream_t), so callers can time with events on it. */
void *ecmc_stream(EcmcHandle *h);
/* Seconds of device time (CUDA events on the handle's stream) spent in event kernels since create. */
double ecmc_kernel_seconds(EcmcHandle *h);
uint64_t ecmc_kernel_launches(EcmcHandle *h);
/* Name of the event kernel that ecmc_run (record = 0) / ecmc_run_recorded (record = 1) launches for this handle's program
 * and options, e.g. "lj_spec_kernel<record=0, prune=1, lanes=4, warps=14>": the string a profile of the run is matched
 * against (bench.py, profiles/). Valid until the next call on the handle. */
const char *ecmc_kernel_name(EcmcHandle *h, int record);

/* ---- batched potential arithmetic on the device -------------------------------------------------------
 * Back the host-side Potential classes -- derivative(velocity, separation, charges) and
 * displacement(velocity, separation, charges, potential_change) of jellyfysh/potential/potential.py:154-301 --
 * and the init-time estimators. velocity: [dimension]; standard-velocity potentials need exactly one positive
 * component (jellyfysh/potential/abstracts.py:105-140), the hard potentials take any velocity.
 * separations: [n][dimension]; charges: [n][2] or NULL (1.0); potential_changes: [n] or NULL; out: [n] (seconds
 * for displacement, as in the reference). */
int ecmc_potential_derivative(const EcmcPotential *potential, int dimension, double system_length,
                              const double *velocity, size_t n, const double *separations, const double *charges,
                              double *out, int device);
int ecmc_potential_displacement(const EcmcPotential *potential, int dimension, double system_length,
                                const double *velocity, size_t n, const double *separations, const double *charges,
                                const double *potential_changes, double *out, int device);
/* Draws of the counter-based random stream, for host-side reproduction: out[n] uniform doubles in [0,1)
 * of (seed, stream, event, slot), starting at draw index `first`. Pure host function (no device needed). */
void ecmc_random_doubles(uint32_t seed, uint32_t stream, uint64_t event, uint32_t slot, uint32_t first, size_t n,
                         double *out);
void ecmc_random_words(uint32_t seed, uint32_t stream, uint64_t event, uint32_t slot, uint32_t first, size_t n,
                       uint32_t *out);

/* Random-stream slots: (kind << 24) | index. See DESIGN.md. */
#define ECMC_SLOT_PAIR_TIME 1u       /* index = target particle; double 0 -> expovariate(beta) */
#define ECMC_SLOT_VETO_TIME 2u       /* double 0 -> Walker uniform, double 1 -> expovariate(beta) */
#define ECMC_SLOT_VETO_CHOICE 3u     /* words -> random.choice over the Walker table (rejection loop) */
#define ECMC_SLOT_CONFIRM 4u         /* double 0 -> uniform(0, bounding rate) in the out-state */
#define ECMC_SLOT_END_OF_CHAIN 5u    /* words -> randint(0, n_particles - 1) (rejection loop) */
#define ECMC_SLOT_LIFTING 6u         /* doubles -> Lifting.insert / RatioLifting draws */
#define ECMC_SLOT_FACTOR_TIME 7u     /* index = target leaf; factor-type-map pair factors: double 0 -> expovariate(beta) */
#define ECMC_SLOT_BENDING_TIME 8u    /* double 0 -> expovariate(beta) of the bending factor */
#define ECMC_SLOT_SWITCH 9u          /* words -> random.choice over the leaves of the object whose root unit hands over */
/* Composite-object pair candidates draw double k of (ECMC_SLOT_PAIR_TIME, target root) for target leaf k. All draws of an
 * out-state come from ECMC_SLOT_CONFIRM in call order: 0 = confirmation, 1 = Lifting.insert of the active unit,
 * 2 = RatioLifting.get_active_identifier. */
#define ECMC_SLOT(kind, index) (((uint32_t)(kind) << 24) | ((uint32_t)(index) & 0xFFFFFFu))

#ifdef __cplusplus
}
#endif
#endif /* ECMC_B200_H */
