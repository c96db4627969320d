/* mcx.h — C ABI of libmcx, the B200-native replacement for MCell4's per-timestep
 * diffuse-and-react hot path (DiffuseReactEvent).
 *
 * Every entry point names the reference interface it replaces (paths relative to the
 * mcellteam/mcell tree).  Signatures are plain C: pointers, sizes, POD structs; no torch,
 * no C++ types.  All lengths are MCell internal length units (1/sqrt(surface_grid_density)
 * um, libmcell/api/mcell4_converter.cpp:243-247), all times are iterations
 * (mcell4_converter.cpp:237).
 *
 * Threading: a handle is NOT re-entrant; one host thread per handle (the reference is
 * single threaded, src4/diffuse_react_event.cpp:54).  Errors never exit()/throw across the
 * ABI (cf. World::fatal_error, src4/world.cpp:561-565): functions return MCX_OK or a
 * negative code and mcx_last_error() returns the text.
 */
#ifndef MCX_H
#define MCX_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MCX_ABI_VERSION 3

/* ---- error codes ------------------------------------------------------------------- */
#define MCX_OK 0
#define MCX_ERR_INVALID_ARG (-1)
#define MCX_ERR_CUDA (-2)        /* CUDA runtime failure or no CUDA device: there is NO CPU fallback */
#define MCX_ERR_CAPACITY (-3)    /* molecule / product / pending buffers exhausted */
#define MCX_ERR_ESCAPED (-4)     /* molecule left the partition (diffuse_react_event.cpp:592-612) */
#define MCX_ERR_STATE (-5)       /* call order violated (e.g. step before upload) */
#define MCX_ERR_OVERFLOW (-6)    /* per-molecule subpartition set overflowed */
#define MCX_ERR_COMM (-7)        /* NCCL failure */

#define MCX_NONE 0xFFFFFFFFu
#define MCX_MAX_PRODUCTS 4
#define MCX_TRACE_K 4

/* Pseudo species for surface-class rules (libmcell/api/model.cpp:95, rxn_utils.inl:192-203) */
#define MCX_ALL_MOLECULES 0xFFFFFFF0u
#define MCX_ALL_VOLUME_MOLECULES 0xFFFFFFF1u
#define MCX_ALL_SURFACE_MOLECULES 0xFFFFFFF2u

/* time sentinels (src4/defines.h:178-179) */
#define MCX_TIME_INVALID (-256.0)
#define MCX_TIME_FOREVER (1e20)

typedef struct mcx_handle mcx_handle;

/* ---- configuration: SimulationConfig (src4/simulation_config.h:28-135) + BNGConfig scalars */
enum { MCX_RNG_PHILOX = 0, MCX_RNG_TAPE = 1 };

typedef struct mcx_config {
  uint32_t abi_version;            /* MCX_ABI_VERSION */
  int32_t  device;                 /* CUDA device ordinal */
  uint64_t seed;                   /* Config.seed; Philox key */
  double   origin[3];              /* Partition::origin_corner (partition.h:248-251) */
  double   partition_edge_length;  /* SimulationConfig::partition_edge_length */
  uint32_t num_subparts_per_edge;  /* num_subparts_per_partition_edge (<= 300, defines.h:159) */
  uint32_t use_expanded_list;      /* config.use_expanded_list (mcell4_converter.cpp:84-87) */
  double   rxn_radius_3d;          /* BNGConfig::rxn_radius_3d, length units */
  double   cell_edge;              /* neighbour-cell edge of the device grid; 0 = auto */
  double   active_llf[3];          /* box that bounds every molecule position (e.g. geometry bbox) */
  double   active_urb[3];          /*   llf == urb == 0 -> whole partition */
  uint64_t max_molecules;          /* slot capacity per device (incl. products and ghosts) */
  uint32_t max_resolve_rounds;     /* conflict-resolution rounds per iteration; 0 = default 8 */
  uint32_t rng_mode;               /* MCX_RNG_PHILOX | MCX_RNG_TAPE (replay) */
  int32_t  rank;                   /* slab decomposition: this process' rank ...           */
  int32_t  world_size;             /* ... of world_size (1 = single GPU)                    */
  uint64_t initial_iteration;      /* Config.initial_iteration (checkpoint resume) */
  double   halo_width;             /* multi-GPU: width of the redundantly evaluated halo, length units; 0 = auto */
} mcx_config;

/* ---- species: BNG::Species subset (src/mcell_species.c:207-300) --------- */
enum {
  MCX_SP_VOL = 1u << 0,            /* volume molecule; clear = surface molecule living on a wall tile */
  MCX_SP_CAN_DIFFUSE = 1u << 1,    /* D != 0 */
  MCX_SP_CANT_INITIATE = 1u << 2   /* TARGET_ONLY */
};
typedef struct mcx_species {
  double   space_step;             /* sqrt(4*1e8*D*time_unit)/length_unit */
  double   time_step;              /* in iterations (1.0 unless custom time step) */
  uint32_t flags;                  /* MCX_SP_* */
  uint32_t reserved;
} mcx_species;

/* ---- reactions: RxnClass / pathway tables ------------------------------ */
enum { MCX_RXN_UNIMOL = 1, MCX_RXN_BIMOL_VOLVOL = 2,
       MCX_RXN_BIMOL_VOLSURF = 3, /* reactants[0] = volume species, reactants[1] = surface species */
       MCX_RXN_BIMOL_VOLWALL = 4  /* reactants[0] = volume species (or MCX_ALL_*), reactants[1] = surface class: a
                                     Standard reaction of a volume molecule with a reactive surface
                                     (collide_and_react_with_walls -> test_intersect -> outcome_intersect,
                                     diffuse_react_event.cpp:991-1067, 1916-1988); reached through a
                                     mcx_surf_class_rxn of type MCX_SURF_STANDARD.  Products are volume species; a kept
                                     reactant 0 reflects, or crosses the wall when kept_info gives it another
                                     orientation (RX_FLIP) */,
       MCX_RXN_BIMOL_SURFSURF = 5 /* both reactants are surface species (either order).  After its move a surface
                                     molecule looks at the molecules on the tiles around its own (find_neighbor_tiles,
                                     grid_utils.inl:1754-1801), tests the matching classes once with the local
                                     probability factor 3 / (number of neighbour tiles) (react_2D_all_neighbors,
                                     diffuse_react_event.cpp:1250-1393; test_bimolecular / test_many_bimolecular,
                                     rxn_utils.inl:336-414, 475-580) and reacts with at most one of them.  Surface
                                     products take the tiles of the consumed reactants (find_surf_product_positions,
                                     :1993-2288, recycled positions) and, beyond those, vacant tiles around the
                                     initiator (the general branch, kept_info required); pathways that free two tiles
                                     but fill one next to a volume product, or keep one reactant and consume the
                                     other while needing vacant tiles, are refused.  max_fixed_p / cum_prob hold the plain pathway probabilities
                                     (rate / grid density scaling done by the table builder, as for the other
                                     kinds) */ };
typedef struct mcx_rxn_class {
  uint32_t kind;                   /* MCX_RXN_* */
  uint32_t reactants[2];           /* species ids in rule order; [1] = MCX_NONE for unimol */
  uint32_t first_pathway;          /* index into the pathway array */
  uint32_t n_pathways;
  uint32_t reserved;
  double   max_fixed_p;            /* RxnClass::get_max_fixed_p() = cum_probs[last] */
  int32_t  reactant_orientation[2];/* RxnClass::get_reactant_orientation (rxn_utils.inl:58-84): +1 ', -1 , and 0 = none */
} mcx_rxn_class;
typedef struct mcx_pathway {
  double   cum_prob;               /* cumulative probability (src/mcell_reactions.c:2932-2933) */
  uint32_t n_products;             /* newly created products (kept reactants are NOT listed) */
  uint32_t products[MCX_MAX_PRODUCTS];
  uint32_t keep_reactant_mask;     /* bit r: rule reactant r appears on both sides (is_simple_cplx_reactant_on_both_
                                      sides_of_rxn_w_identical_compartments, diffuse_react_event.cpp:2558-2565) */
  uint32_t rxn_rule_id;            /* slot of the per-reaction occurrence counter */
  uint32_t kept_info;              /* kept reactants among the rule's products (diffuse_react_event.cpp:2618-2716):
                                      MCX_KEPT_VALID | rule product order, one nibble per rule product (0-3 =
                                      products[k], 8 + r = kept reactant r, 0xF = end) | product-side orientation of
                                      kept reactant r in bits 24 + 2 r (0 none, 1 = up ', 2 = down ,).  The order is the
                                      order of the orientation draws; a kept volume reactant of a volume-surface
                                      reaction whose product-side orientation differs from its reactant-side one
                                      passes through the wall (RX_FLIP, :945-970), a kept surface reactant takes its
                                      product-side orientation.  0 (not valid): kept reactants stay as they are */
  int32_t  product_orientation[MCX_MAX_PRODUCTS]; /* rule orientation of products[k]; 0 = none: one random bit
                                      when a surface is involved (diffuse_react_event.cpp:2622-2627) */
} mcx_pathway;

#define MCX_KEPT_VALID (1u << 31)
#define MCX_KEPT_ORDER_END 0xFu
#define MCX_KEPT_ORDER_REACTANT 8u   /* nibble value 8 + r: kept reactant r */

/* ---- surface classes (mcell4_converter.cpp:515-622; rxn_utils.inl:263-287) ----------- */
enum { MCX_SURF_REFLECTIVE = 0, MCX_SURF_TRANSPARENT = 1, MCX_SURF_ABSORPTIVE = 2,
       MCX_SURF_STANDARD = 3 /* a finite-rate reaction: rxn_class names a MCX_RXN_BIMOL_VOLWALL class */ };
typedef struct mcx_surf_class_rxn {
  uint32_t species;                /* species id, MCX_ALL_MOLECULES, MCX_ALL_VOLUME_MOLECULES or MCX_ALL_SURFACE_MOLECULES */
  uint32_t surf_class;             /* value used in wall_surf_class[] */
  int32_t  orientation;            /* 0: both sides; +1: hits on the FRONT only; -1: BACK only */
  uint32_t type;                   /* MCX_SURF_* */
  uint32_t rxn_class;              /* MCX_SURF_STANDARD: index of the reaction class */
} mcx_surf_class_rxn;

/* ---- molecules: SoA view of Partition::molecules (src4/molecule.h:52-260) ------------ */
enum {
  MCX_MOL_DEFUNCT = 1u << 0,          /* MOLECULE_FLAG_DEFUNCT */
  MCX_MOL_SCHEDULE_UNIMOL = 1u << 1,  /* MOLECULE_FLAG_SCHEDULE_UNIMOL_RXN: lifetime not drawn yet */
  MCX_MOL_PARTIAL = 1u << 2,          /* diffusion_time is fractional (born mid-iteration) */
  MCX_MOL_CVI_PENDING = 1u << 3       /* v.counted_volume_index is a guess: it is recomputed by a ray cast when the molecule
                                         is next evaluated (Partition::add_volume_molecule does that for
                                         COUNTED_VOLUME_INDEX_INVALID, partition.h:572-576); needs
                                         mcx_set_counted_volume_objects */
};
typedef struct mcx_mol_soa {
  uint64_t  n;
  double   *x, *y, *z;             /* Molecule::v.pos */
  uint32_t *id;                    /* Molecule::id (unique, persistent) */
  uint32_t *species;               /* Molecule::species_id */
  uint32_t *flags;                 /* MCX_MOL_* */
  double   *diffusion_time;        /* Molecule::diffusion_time; NULL = all at current iteration */
  double   *unimol_rxn_time;       /* Molecule::unimol_rxn_time; NULL = MCX_TIME_INVALID */
  /* Molecule::s for surface molecules (src4/molecule.h); all NULL = volume molecules only.
   * For a surface molecule x,y,z are ignored on upload and returned as uv2xyz(u,v) on download. */
  uint32_t *wall;                  /* s.wall_index; MCX_NONE for a volume molecule */
  uint32_t *tile;                  /* s.grid_tile_index on that wall's grid */
  int32_t  *orientation;           /* s.orientation: +1 up / -1 down */
  double   *u, *v;                 /* s.pos in the wall's uv frame */
  uint32_t *counted_volume;        /* v.counted_volume_index (< 256); NULL = 0 = outside every counted object */
} mcx_mol_soa;

/* ---- per-call statistics: SimulationStats mirror (src4/simulation_stats.h:46-83) ----- */
typedef struct mcx_step_stats {
  uint64_t iterations;
  uint64_t molecule_steps;         /* sum over iterations of molecules diffused (the bench metric) */
  uint64_t n_live;                 /* live molecules after the call */
  uint64_t ray_polygon_tests;
  uint64_t ray_polygon_colls;
  uint64_t mol_wall_reflections;
  uint64_t mol_wall_transparent;
  uint64_t mol_wall_absorptions;
  uint64_t vol_mol_vol_mol_collisions;
  uint64_t bimol_rxns;
  uint64_t unimol_rxns;
  uint64_t wall_redos;
  uint64_t resolve_retries;        /* molecules re-evaluated after losing a reaction conflict */
  uint64_t unresolved_conflicts;   /* proposals dropped after max_resolve_rounds */
  uint64_t products_created;
  uint64_t kernel_launches;        /* libmcx kernels launched by this call */
  double   device_ms;              /* CUDA-event time of the iteration loop */
  double   ms_diffuse;             /* with mcx_set_profiling: summed CUDA-event time of k_diffuse_fast */
  double   ms_resolve;             /*   ... of the conflict-resolution rounds */
  double   ms_sort;                /*   ... of histogram scan + scatter */
  uint64_t profiled_iterations;    /*   iterations covered by the sums above */
  double   ms_diffuse_slow;        /*   ... of k_diffuse_slow (deferred molecules, generic evaluation) */
  uint64_t deferred_molecules;     /* molecules the fast diffuse pass handed to the generic evaluation */
  uint64_t deferred_by_reason[8];  /* ... split by MCX_DEFER_* */
} mcx_step_stats;

/* why k_diffuse_fast handed a molecule to the generic evaluation (diagnostics) */
enum {
  MCX_DEFER_TIMING = 0,       /* partial step, newborn, unimolecular event inside the iteration, non-diffusing */
  MCX_DEFER_GEOMETRY = 1,     /* leaves the partition or crosses several subpartition faces near walls */
  MCX_DEFER_WALL = 2,         /* a wall of the start/end subpartition is not rejected by the plane test */
  MCX_DEFER_PROBE_SHAPE = 3,  /* swept box overlaps more than 2x2 cell rows */
  MCX_DEFER_MULTI_HIT = 4,    /* more than one collision partner */
  MCX_DEFER_FOREIGN_HIT = 5,  /* single partner outside the molecule's own subpartition */
  MCX_DEFER_DISK = 6          /* collision next to walls: the exact_disk occlusion factor is needed */
};

/* ---- replay trace (kernel-level parity; mirrors the reference's DEBUG_* dumps,
 *      include/debug_config.h:153-245) ------------------------------------------------- */
enum {
  MCX_OUT_NONE = 0, MCX_OUT_MOVED = 1, MCX_OUT_REACTED = 2, MCX_OUT_ABSORBED = 3,
  MCX_OUT_UNIMOL = 4, MCX_OUT_CONSUMED = 5, MCX_OUT_STATIC = 6,
  MCX_OUT_SURFMOVE = 7, /* surface molecule took a new tile (move_sm_on_same_triangle / move_sm_to_new_triangle) */
  MCX_OUT_WALLRXN = 8   /* reacted with the surface class of a wall (MCX_RXN_BIMOL_VOLWALL) */
};
typedef struct mcx_trace_rec {
  uint32_t id;
  uint32_t outcome;                /* MCX_OUT_* */
  uint32_t n_words;                /* random words consumed by this molecule this iteration */
  uint32_t n_wall_hits;            /* walls processed (reflect + transparent + absorb) */
  uint32_t n_collisions;           /* vol-vol collisions evaluated in time order */
  uint32_t n_redo;
  uint32_t wall[MCX_TRACE_K];      /* first K wall indices hit, in order */
  uint32_t wall_side[MCX_TRACE_K]; /* 1 = FRONT, 2 = BACK */
  uint32_t partner[MCX_TRACE_K];   /* first K collision partner ids, in evaluation order */
  uint32_t rxn_class;              /* MCX_NONE if no reaction */
  uint32_t rxn_pathway;
  uint32_t rxn_partner;            /* partner id for bimolecular, MCX_NONE otherwise */
  uint32_t rounds;                 /* evaluation passes (1 + conflict retries) */
  uint64_t event_hash;             /* hash over the complete ordered event sequence */
  double   pos[3];                 /* final position (reaction position if consumed as initiator) */
  double   t_event;                /* absolute time of the reaction / absorption, else 0 */
} mcx_trace_rec;

/* ---- lifecycle ----------------------------------------------------------------------- */
/* Replaces: DiffuseReactEvent construction in World::init_simulation (src4/world.cpp:279-282)
 * plus the Partition constructor's grid set-up (src4/partition.cpp:36-77). */
int mcx_create(const mcx_config* cfg, mcx_handle** out);
/* Replaces: Scheduler deleting the event (src4/scheduler.h:124-126). */
void mcx_destroy(mcx_handle* h);
/* Replaces: World::fatal_error message path (src4/world.cpp:561-565). h may be NULL (create errors). */
const char* mcx_last_error(const mcx_handle* h);
int mcx_abi_version(void);

/* ---- one-time immutable tables -------------------------------------------------------- */
/* Replaces: Partition::add_geometry_vertex / add_uninitialized_wall,
 * Wall::initialize_wall_constants (src4/wall.cpp:281-342) and Partition::finalize_walls
 * (src4/partition.cpp:91-118 -> geometry_utils.inl:110-207 -> wall_utils.inl:326-504).
 * vertices: 3*n_vertices doubles; tri: 3*n_walls vertex indices;
 * wall_surf_class: n_walls entries or NULL (MCX_NONE = default reflective). */
int mcx_set_geometry(mcx_handle* h, const double* vertices, uint64_t n_vertices,
                     const uint32_t* tri, uint64_t n_walls,
                     const uint32_t* wall_surf_class, const uint32_t* wall_object);
/* Replaces: p.get_species(id) lookups (src4/partition.h:989-997; libbng Species). */
int mcx_set_species(mcx_handle* h, const mcx_species* species, uint32_t n_species);
/* Replaces: RxnContainer::get_bimol_rxn_class / get_unimol_rxn_class and
 * RxnClass::{get_max_fixed_p,get_pathway_index_for_probability} (libbng; call sites
 * src4/collision_utils.inl:537-538, src4/rxn_utils.inl:353-412,712-782). */
int mcx_set_reactions(mcx_handle* h, const mcx_rxn_class* classes, uint32_t n_classes,
                      const mcx_pathway* pathways, uint32_t n_pathways);
/* Replaces: RxnUtils::trigger_intersect table walk (src4/rxn_utils.inl:263-287). */
int mcx_set_surface_classes(mcx_handle* h, const mcx_surf_class_rxn* rules, uint32_t n_rules);

/* ---- molecule state ------------------------------------------------------------------- */
/* Replaces: Partition::add_volume_molecule for a batch (src4/partition.h:555-611);
 * HOST pointers. Resets the device population. */
int mcx_upload_molecules(mcx_handle* h, const mcx_mol_soa* mols);
/* Replaces: Partition::get_molecules() read access (src4/partition.h:648-666) for viz,
 * checkpoint and API introspection.  capacity = arrays' length; out->n = live count. */
int mcx_download_molecules(mcx_handle* h, mcx_mol_soa* out, uint64_t capacity);
uint64_t mcx_num_molecules(mcx_handle* h);

/* ---- the hot path --------------------------------------------------------------------- */
/* Replaces: DiffuseReactEvent::step() (src4/diffuse_react_event.cpp:53-64) called
 * n_iterations times in a row; the scheduler's barrier contract allows up to
 * time_up_to_next_barrier iterations per call (src4/diffuse_react_event.h:37,142-150).
 * Also folds SortMolsBySubpartEvent::step (sort_mols_by_subpart_event.cpp:62-98) and
 * DefragmentationEvent::step (defragmentation_event.cpp:31-114) into every iteration. */
int mcx_step(mcx_handle* h, uint32_t n_iterations, mcx_step_stats* stats_out);
/* One iteration in which molecule `id` consumes words[offset_by_id[id] ...] instead of its
 * Philox stream (requires rng_mode == MCX_RNG_TAPE); trace_out[id] receives the per-molecule
 * event trace.  Replaces the reference's lock-step DEBUG_DIFFUSION/DEBUG_COLLISIONS/DEBUG_RXNS
 * trace method (include/debug_config.h:153-245).  n_ids = number of trace slots. */
int mcx_replay_step(mcx_handle* h, const uint32_t* words, uint64_t n_words,
                    const uint64_t* offset_by_id, uint64_t n_ids,
                    mcx_trace_rec* trace_out, mcx_step_stats* stats_out);
/* Like mcx_step for one iteration with the Philox streams, but also returns the trace. */
int mcx_trace_step(mcx_handle* h, uint64_t n_ids, mcx_trace_rec* trace_out,
                   mcx_step_stats* stats_out);

/* Per-kernel CUDA-event timing inside mcx_step (events on the launching stream). Off by default. */
int mcx_set_profiling(mcx_handle* h, int enabled);

/* ---- observables ---------------------------------------------------------------------- */
/* Replaces: MolOrRxnCountEvent::compute_counts world-count fast path
 * (src4/mol_or_rxn_count_event.cpp:622-653) and rxn occurrence counters
 * (src4/partition.h:1036-1077).  Arrays may be NULL.  Multi-GPU: summed over ranks. */
int mcx_counts(mcx_handle* h, uint64_t* per_species, uint32_t n_species,
               uint64_t* per_rxn_rule, uint32_t n_rxn_rules);

/* Counted volumes (Partition::counted_volumes, World::init_counted_volumes src4/world.cpp:146-155; the host owns
 * the geometry analysis, as the reference's VtkUtils does).  A counted volume is one distinct set of enclosing
 * counted objects; index 0 = outside all (defines.h:290).  Per wall: the counted volume on its front (normal) side
 * and on its back side — for a wall of a non-counted object both are the volume that object lies in.  A molecule
 * crossing a transparent wall takes the index of the side it arrives on
 * (CollisionUtils::update_counted_volume_id_when_crossing_wall, collision_utils.inl:1637-1694, non-intersecting
 * objects); products inherit it.  Call after mcx_set_geometry; n_counted_volumes <= 256. */
int mcx_set_counted_volumes(mcx_handle* h, uint32_t n_counted_volumes, const uint8_t* wall_cv_front,
                            const uint8_t* wall_cv_back);
/* Replaces: MolOrRxnCountEvent::compute_counts terms restricted to a volume (mol_or_rxn_count_event.cpp:607-716)
 * and Partition::inc_rxn_in_volume_occured_count (partition.h:1036-1077).
 * mol_counts[species * n_counted_volumes + cv], rxn_counts[rxn_rule * n_counted_volumes + cv]; either may be NULL. */
int mcx_counts_by_volume(mcx_handle* h, uint64_t* mol_counts, uint64_t* rxn_counts);
/* Counted objects that intersect each other (the reference recomputes the counted volume from waypoints then,
 * update_counted_volume_id_when_crossing_wall / compute_counted_volume_using_waypoints, collision_utils.inl:1568-1694):
 * a wall of such an object has no single volume in front of and behind it.  cv_object_mask[cv] = the set of counted
 * objects that enclose counted volume cv (bit k = geometry object k, k < 32; every set that can occur must be listed);
 * a molecule that crosses a wall of an object named in intersecting_objects toggles that object's bit in its set
 * instead of reading the wall's pair; volume products of unimolecular surface reactions on such walls and molecules
 * flagged MCX_MOL_CVI_PENDING get their set from a ray cast at their next evaluation; releases inside regions read it off
 * their ray.  Call after mcx_set_counted_volumes; NULL switches it off. */
int mcx_set_counted_volume_objects(mcx_handle* h, const uint32_t* cv_object_mask, uint32_t intersecting_objects);

/* Region borders for surface molecules (ray_trace_surf, diffuse_react_event.cpp:1627-1665; reflect_absorb_inside_out /
 * outside_in, diffusion_utils.inl:598-700): wall_edge_border[w] bit e = edge e of wall w (0: v0-v1, 1: v1-v2, 2: v2-v0) is
 * a border of a reactive region the wall belongs to (WallUtils::is_wall_edge_region_border with
 * region_must_be_reactive).  A surface molecule that reaches such an edge — leaving the region or entering it — turns
 * back when the wall's surface class is MCX_SURF_REFLECTIVE for its species and orientation, is destroyed when it is
 * MCX_SURF_ABSORPTIVE (absorptive region border), passes otherwise.  NULL removes the borders. */
int mcx_set_region_borders(mcx_handle* h, const uint8_t* wall_edge_border);

/* Counted surface regions (MolOrRxnCountEvent terms CountType::PresentOnSurfaceRegion and RxnCountOnSurfaceRegion,
 * src4/mol_or_rxn_count_event.cpp:528-534, 588-600; the reference keeps reaction counts per wall,
 * Partition::inc_rxn_on_surface_occured_count, diffuse_react_event.cpp:2513-2521).  The host numbers the distinct sets of
 * counted regions a wall can belong to (set 0 = none) and hands over one index per wall; a region expression is then
 * evaluated per set on the host (wall_matches_region_expr_recursively).  Call after mcx_set_geometry;
 * n_region_sets <= 256.  Reactions are counted from this call on. */
int mcx_set_surface_regions(mcx_handle* h, uint32_t n_region_sets, const uint8_t* wall_region_set);
/* mol_counts[species * n_region_sets + set]: surface molecules on the walls of each set;
 * rxn_counts[rxn_rule * n_region_sets + set]: reactions whose initiator was a surface molecule there; either may be NULL. */
int mcx_counts_by_surface_region(mcx_handle* h, uint64_t* mol_counts, uint64_t* rxn_counts);

/* ---- release on the device (new; ReleaseEvent::release_ellipsoid_or_rectcuboid, src4/release_event.cpp:953-1003) --- */
/* The reference releases molecules on the host, sequentially from its one random stream; 1e8 molecules cannot be
 * pushed through that (or through PCIe as a 52-byte SoA) quickly.  mcx_release_volume_molecules creates `number`
 * volume molecules of one species on the device with the reference's arithmetic per molecule — CUBIC: pos = rng_dbl
 * - 0.5 per axis; SPHERICAL: the same, redrawn while |pos|^2 >= 0.25; SPHERICAL_SHELL: then pos /= 2 |pos| ((0, 0,
 * 0.5) when |pos| == 0); location = pos * diameter + location — each molecule drawing from its OWN Philox stream
 * (seed, molecule id, iteration | 2^63: a domain disjoint from the diffusion streams).  New molecules get the ids
 * first_id .. first_id + number - 1, MOLECULE_FLAG_SCHEDULE_UNIMOL_RXN, diffusion_time = release_time (which must lie
 * in [iteration, iteration + 1)) and the given counted-volume index.  Needs rng_mode == MCX_RNG_PHILOX and a previous
 * mcx_upload_molecules (which may be empty).  With several ranks every rank makes the same call and keeps the molecules
 * of its own slab. */
#define MCX_RELEASE_CUBIC 0
#define MCX_RELEASE_SPHERICAL 1
#define MCX_RELEASE_SPHERICAL_SHELL 2
/* MCX_RELEASE_REGION: ReleaseEvent::release_inside_regions (release_event.cpp:904-951) — uniform in the box (location =
 * centre, diameter = edges of the region's bounding box, region_llf / region_urb), a point is kept when it lies inside
 * every closed object of region_in and outside every object of region_out (bit k = geometry object k of
 * mcx_set_geometry's wall_object, k < 32: the INTERSECT / DIFFERENCE operators of the release's region expression;
 * a UNION is released as separate calls over disjoint pieces), redrawn from the molecule's own stream otherwise.  The
 * test is a ray cast through the subpartition wall lists (Region::is_point_inside, geometry.cpp:1048-1086); the
 * counted volume of each molecule comes from the same ray (compute_counted_volume_for_pos, collision_utils.inl:
 * 1515-1566) and counted_volume_index is ignored. */
#define MCX_RELEASE_REGION 3
typedef struct mcx_release {
  uint32_t species;
  uint32_t shape;                  /* MCX_RELEASE_* */
  uint64_t number;
  double   location[3];            /* length units */
  double   diameter[3];            /* length units */
  double   release_time;           /* iterations; 0 = start of the current iteration */
  uint32_t counted_volume_index;   /* Molecule::v.counted_volume_index of the released molecules (0 = outside all) */
  uint32_t reserved;
  uint32_t region_in, region_out;  /* MCX_RELEASE_REGION: object masks (see above) */
  /* MCX_RELEASE_REGION, general region expressions (RegionExprNode trees, is_point_inside_region_expr_recursively,
   * release_event.cpp:787-813): region_expr_len > 0 replaces the two masks by a postfix program of region_expr_len
   * bytes — k < 32 pushes "inside geometry object k", MCX_REGION_UNION / _INTERSECT / _DIFFERENCE pop two values
   * (left below right) and push the result; the point is kept when the one value left is true */
  uint32_t region_expr_len;
  uint8_t  region_expr[28];
} mcx_release;
#define MCX_REGION_UNION 0x80
#define MCX_REGION_INTERSECT 0x81
#define MCX_REGION_DIFFERENCE 0x82
int mcx_release_volume_molecules(mcx_handle* h, const mcx_release* r, uint32_t* first_id_out);
/* Partition::next_molecule_id (partition.h): the id the next new molecule takes.  mcx_upload_molecules raises it above
 * every uploaded id; a run resumed from a checkpoint restores the saved value with mcx_set_next_molecule_id (ids of
 * molecules that no longer exist must not be handed out again, and the per-molecule random streams are keyed by id), which
 * can only raise it.  Same call on every rank. */
int mcx_get_next_molecule_id(mcx_handle* h, uint32_t* next_id_out);
int mcx_set_next_molecule_id(mcx_handle* h, uint32_t next_id);
/* Replaces: Wall::has_initialized_grid of every wall (src4/wall.h:339-346) as part of a checkpoint.  A wall gets its grid
 * with its first surface molecule and keeps it; the neighbour search of the surface-surface reactions skips walls without
 * one (src4/grid_utils.inl:1243, 783-790), so the flags belong to the state of a model with MCX_RXN_BIMOL_SURFSURF
 * classes.  mcx_upload_molecules never takes a grid away (the host may add and remove molecules between steps, the
 * reference's walls keep their grids); mcx_set_geometry starts over.  A run resumed on a new handle restores the saved
 * flags with mcx_set_wall_grids after the upload (they are OR-ed in).  One byte per wall, 0 / 1. */
int mcx_get_wall_grids(mcx_handle* h, uint8_t* has_grid_out, uint64_t n_walls);
int mcx_set_wall_grids(mcx_handle* h, const uint8_t* has_grid, uint64_t n_walls);
/* ReleaseEvent::release_list (release_event.cpp:1008-1040), volume molecules: one molecule of species[k] at
 * (x[k], y[k], z[k]) (length units) with counted_volume[k] (may be NULL: 0), ids first_id .. first_id + n - 1 in list
 * order, added to the resident population without a download / upload round trip.  Same call on every rank. */
int mcx_release_list(mcx_handle* h, uint64_t n, const uint32_t* species, const double* x, const double* y, const double* z,
                     const uint32_t* counted_volume, double release_time, uint32_t* first_id_out);

/* Surface molecules onto regions (ReleaseEvent::release_onto_regions, release_event.cpp:640-760): `number` molecules of a
 * surface species on vacant tiles of the listed walls (the walls of the release's surface region(s), in the order of
 * cumm_area_and_pwall_index_pairs), each tile chosen like the reference does — A = rng_dbl * total_area, the wall by
 * bisection of the cumulative areas, tile = num_tiles * (A - area before the wall) / wall.area — and taken only if it
 * is vacant.  The reference places one molecule after the other from its one random stream; here every molecule draws
 * from its OWN Philox stream (release domain) and the placement runs in rounds: a tile goes to the lowest id that
 * picked it, the others (and those that picked an occupied tile) draw again next round; whoever is still without a tile
 * after MCX_SURFACE_RELEASE_ROUNDS rounds takes the first vacant tiles in wall-list order, lowest id first (the
 * reference's fall-back, :702-744).  MCX_ERR_CAPACITY when the walls have fewer vacant tiles than `number`.
 * Position on the tile: random (GridUtils::grid2uv_random, config.randomize_smol_pos) or the tile centre (grid2uv);
 * orientation 0 draws one bit per molecule (place_single_molecule_onto_grid, grid_utils.inl:2097-2166).  New molecules
 * get ids first_id .. first_id + number - 1 and MOLECULE_FLAG_SCHEDULE_UNIMOL_RXN.  One device only (a rank knows
 * the occupancy of its own slab alone). */
#define MCX_SURFACE_RELEASE_ROUNDS 64
typedef struct mcx_surface_release {
  uint32_t species;
  int32_t  orientation;            /* +1, -1, or 0 = random */
  uint64_t number;
  double   release_time;           /* iterations; 0 = start of the current iteration */
  const uint32_t* walls;           /* wall indices */
  uint64_t n_walls;
  uint32_t randomize_pos;          /* config.randomize_smol_pos */
  uint32_t reserved;
} mcx_surface_release;
int mcx_release_surface_molecules(mcx_handle* h, const mcx_surface_release* r, uint32_t* first_id_out);

/* ---- multi-GPU (new: the reference has a single partition, world.cpp:147,277) --------- */
/* nccl_unique_id: the 128-byte ncclUniqueId created by rank 0 and broadcast by the host
 * plumbing (torch.distributed).  Call order on every rank: mcx_create (rank/world_size in the config),
 * mcx_set_species/..., mcx_comm_init, mcx_upload_molecules (each rank may upload any superset of its own slab:
 * molecules of other slabs are dropped), mcx_step.  Slabs are z-layers of the device cell grid. */
int mcx_comm_init(mcx_handle* h, const void* nccl_unique_id, uint32_t id_bytes);
/* Slab layout of this rank after mcx_comm_init: z-layers of the global device cell grid.  A position z belongs to
 * layer clamp(floor((z - grid_origin_z) * layer_rcp), 0, n_layers - 1); rank r owns layers [b(r), b(r + 1)) with
 * b(0) = 0, b(world) = n_layers and b(k) = halo_layers + (n_layers - 2 * halo_layers) * k / world in between (integer
 * division): the two outermost ranks have one halo only and own halo_layers more, so that every rank evaluates the
 * same number of layers.  Hosts that distribute molecules use exactly this arithmetic (IEEE double for the layer)
 * so that host and device agree on every molecule. */
typedef struct mcx_slab_info {
  double   grid_origin_z;          /* length units */
  double   layer_rcp;              /* 1 / layer thickness */
  uint32_t n_layers;               /* global */
  uint32_t layer_lo, layer_hi;     /* owned by this rank: [layer_lo, layer_hi) */
  uint32_t halo_layers;
  int32_t  rank, world_size;
} mcx_slab_info;
int mcx_slab_info_get(mcx_handle* h, mcx_slab_info* out);
/* How the halo refresh of this handle moves its records: 0 = single device (no exchange), 1 = NCCL send/recv,
 * 2 = stores into the neighbours' memory over NVLink (peer memory).  Negative on error. */
int mcx_comm_halo_path(mcx_handle* h);
/* Which kernel evaluated the whole-step molecules in the last mcx_step / mcx_replay_step / mcx_trace_step call:
 * 0 = the gather walk over the cell-sorted snapshot (k_diffuse_fast), 1 = shared-memory tiles staged with bulk copies
 * (k_diffuse_tile; selected with the environment variable MCX_TILE=1 when the cell grid allows it).  Both compute
 * DiffuseReactEvent::diffuse_molecules (src4/diffuse_react_event.cpp:67-161) with identical results. */
int mcx_fast_pass_kind(mcx_handle* h);
/* Rank 0 creates the id (ncclGetUniqueId) that every rank passes to mcx_comm_init; returns the number of bytes
 * written (<= bytes) or a negative error. */
int mcx_comm_unique_id(void* out, uint32_t bytes);

/* ---- host helpers that need no device (also exported for the CPU test tier) ----------- */
/* Surface grid of one triangle for host-side placement of surface molecules (release stays on the host):
 * v9 = three vertices; returns num_tiles = ceil(sqrt(area))^2 (Grid::initialize, src4/wall.cpp:38-74). */
uint32_t mcx_grid_num_tiles(const double* v9);
/* The subpartition wall lists the device walks, built exactly as mcx_set_geometry builds them (host only): CSR over the
 * n^3 subpartitions, start_out[n^3 + 1], ascending wall indices in list_out (up to cap entries); returns the number of
 * entries.  Replaces Partition::finalize_walls -> GeometryUtils::wall_subparts_collision_test -> WallUtils::wall_in_box
 * (src4/partition.cpp:91-118, geometry_utils.inl:110-207, wall_utils.inl:326-504). */
uint64_t mcx_walls_per_subpart(const double* origin3, double partition_edge_length, uint32_t n_subparts_per_edge,
                               double rxn_radius_3d, uint32_t use_expanded_list, const double* vertices, uint64_t n_vertices,
                               const uint32_t* tri, uint64_t n_walls, uint32_t* start_out, uint32_t* list_out, uint64_t cap);
/* The neighbour tiles of every tile of every wall, built exactly as the device table of the surface-surface partner search
 * is (host only): CSR over all tiles in wall order (tile_start of a wall = sum of the tiles of the walls before it),
 * start_out[total tiles + 1], (wall, tile) pairs in list_out (up to cap pairs) in the order the reference walks its list;
 * returns the number of pairs.  Every wall is taken to have a grid; at run time the entries of walls without one are
 * skipped.  Replaces GridUtils::find_neighbor_tiles and everything under it (src4/grid_utils.inl:296-1801), which the
 * reference runs per molecule and step from react_2D_all_neighbors (src4/diffuse_react_event.cpp:1267). */
uint64_t mcx_tile_neighbor_table(const double* vertices, uint64_t n_vertices, const uint32_t* tri, uint64_t n_walls,
                                 const uint32_t* wall_object, uint32_t* start_out, uint32_t* list_out, uint64_t cap);
/* Centre of a tile in the wall's uv frame (GridUtils::grid2uv, src4/grid_utils.inl:233-253). */
void mcx_grid2uv(const double* v9, uint32_t tile, double* uv2);
/* Tile under a point of the wall (GridUtils::xyz2grid_tile_index, src4/grid_utils.inl:48-118). */
uint32_t mcx_xyz2grid(const double* v9, const double* xyz3);
/* Philox4x32-10 block for (seed, molecule id, iteration, block index); the device stream
 * of molecule `id` is the concatenation of blocks 0,1,2,... */
void mcx_philox_block(uint64_t seed, uint32_t mol_id, uint64_t iteration, uint32_t block,
                      uint32_t out[4]);

#ifdef __cplusplus
}
#endif
#endif /* MCX_H */
