"""ctypes mirror of include/mcx.h (the C ABI of libmcx).  Struct layouts must match the header
field for field; tests/test_abi.py checks sizes against the library's own sizeof table."""
import ctypes as C

MCX_ABI_VERSION = 3
MCX_OK = 0
MCX_ERR_INVALID_ARG, MCX_ERR_CUDA, MCX_ERR_CAPACITY, MCX_ERR_ESCAPED = -1, -2, -3, -4
MCX_ERR_STATE, MCX_ERR_OVERFLOW, MCX_ERR_COMM = -5, -6, -7
MCX_NONE = 0xFFFFFFFF
MCX_MAX_PRODUCTS = 4
MCX_TRACE_K = 4
MCX_ALL_MOLECULES = 0xFFFFFFF0
MCX_ALL_VOLUME_MOLECULES = 0xFFFFFFF1
MCX_ALL_SURFACE_MOLECULES = 0xFFFFFFF2
MCX_TIME_INVALID = -256.0
MCX_TIME_FOREVER = 1e20
MCX_RNG_PHILOX, MCX_RNG_TAPE = 0, 1
MCX_SP_VOL, MCX_SP_CAN_DIFFUSE, MCX_SP_CANT_INITIATE = 1, 2, 4
MCX_RXN_UNIMOL, MCX_RXN_BIMOL_VOLVOL, MCX_RXN_BIMOL_VOLSURF, MCX_RXN_BIMOL_VOLWALL, MCX_RXN_BIMOL_SURFSURF = 1, 2, 3, 4, 5
MCX_SURF_REFLECTIVE, MCX_SURF_TRANSPARENT, MCX_SURF_ABSORPTIVE, MCX_SURF_STANDARD = 0, 1, 2, 3
MCX_MOL_DEFUNCT, MCX_MOL_SCHEDULE_UNIMOL, MCX_MOL_PARTIAL, MCX_MOL_CVI_PENDING = 1, 2, 4, 8
MCX_KEPT_VALID, MCX_KEPT_ORDER_END, MCX_KEPT_ORDER_REACTANT = 1 << 31, 0xF, 8
MCX_OUT_NONE, MCX_OUT_MOVED, MCX_OUT_REACTED, MCX_OUT_ABSORBED = 0, 1, 2, 3
MCX_OUT_UNIMOL, MCX_OUT_CONSUMED, MCX_OUT_STATIC, MCX_OUT_SURFMOVE, MCX_OUT_WALLRXN = 4, 5, 6, 7, 8

c_u32, c_u64, c_i32, c_f64 = C.c_uint32, C.c_uint64, C.c_int32, C.c_double
P = C.POINTER


class mcx_config(C.Structure):
    _fields_ = [
        ("abi_version", c_u32), ("device", c_i32), ("seed", c_u64),
        ("origin", c_f64 * 3), ("partition_edge_length", c_f64),
        ("num_subparts_per_edge", c_u32), ("use_expanded_list", c_u32),
        ("rxn_radius_3d", c_f64), ("cell_edge", c_f64),
        ("active_llf", c_f64 * 3), ("active_urb", c_f64 * 3),
        ("max_molecules", c_u64), ("max_resolve_rounds", c_u32), ("rng_mode", c_u32),
        ("rank", c_i32), ("world_size", c_i32), ("initial_iteration", c_u64), ("halo_width", c_f64),
    ]


MCX_RELEASE_CUBIC, MCX_RELEASE_SPHERICAL, MCX_RELEASE_SPHERICAL_SHELL, MCX_RELEASE_REGION = 0, 1, 2, 3
MCX_REGION_UNION, MCX_REGION_INTERSECT, MCX_REGION_DIFFERENCE = 0x80, 0x81, 0x82


class mcx_release(C.Structure):
    _fields_ = [("species", c_u32), ("shape", c_u32), ("number", c_u64), ("location", c_f64 * 3), ("diameter", c_f64 * 3),
                ("release_time", c_f64), ("counted_volume_index", c_u32), ("reserved", c_u32),
                ("region_in", c_u32), ("region_out", c_u32), ("region_expr_len", c_u32), ("region_expr", C.c_uint8 * 28)]


class mcx_surface_release(C.Structure):
    _fields_ = [("species", c_u32), ("orientation", c_i32), ("number", c_u64), ("release_time", c_f64),
                ("walls", C.c_void_p), ("n_walls", c_u64), ("randomize_pos", c_u32), ("reserved", c_u32)]


class mcx_slab_info(C.Structure):
    _fields_ = [("grid_origin_z", c_f64), ("layer_rcp", c_f64), ("n_layers", c_u32), ("layer_lo", c_u32),
                ("layer_hi", c_u32), ("halo_layers", c_u32), ("rank", c_i32), ("world_size", c_i32)]


class mcx_species(C.Structure):
    _fields_ = [("space_step", c_f64), ("time_step", c_f64), ("flags", c_u32), ("reserved", c_u32)]


class mcx_rxn_class(C.Structure):
    _fields_ = [("kind", c_u32), ("reactants", c_u32 * 2), ("first_pathway", c_u32),
                ("n_pathways", c_u32), ("reserved", c_u32), ("max_fixed_p", c_f64),
                ("reactant_orientation", c_i32 * 2)]


class mcx_pathway(C.Structure):
    _fields_ = [("cum_prob", c_f64), ("n_products", c_u32), ("products", c_u32 * MCX_MAX_PRODUCTS),
                ("keep_reactant_mask", c_u32), ("rxn_rule_id", c_u32), ("kept_info", c_u32),
                ("product_orientation", c_i32 * MCX_MAX_PRODUCTS)]


class mcx_surf_class_rxn(C.Structure):
    _fields_ = [("species", c_u32), ("surf_class", c_u32), ("orientation", c_i32), ("type", c_u32), ("rxn_class", c_u32)]


class mcx_mol_soa(C.Structure):
    _fields_ = [("n", c_u64), ("x", P(c_f64)), ("y", P(c_f64)), ("z", P(c_f64)),
                ("id", P(c_u32)), ("species", P(c_u32)), ("flags", P(c_u32)),
                ("diffusion_time", P(c_f64)), ("unimol_rxn_time", P(c_f64)),
                ("wall", P(c_u32)), ("tile", P(c_u32)), ("orientation", P(c_i32)), ("u", P(c_f64)), ("v", P(c_f64)),
                ("counted_volume", P(c_u32))]


class mcx_step_stats(C.Structure):
    _fields_ = [(n, c_u64) for n in (
        "iterations", "molecule_steps", "n_live", "ray_polygon_tests", "ray_polygon_colls",
        "mol_wall_reflections", "mol_wall_transparent", "mol_wall_absorptions",
        "vol_mol_vol_mol_collisions", "bimol_rxns", "unimol_rxns", "wall_redos",
        "resolve_retries", "unresolved_conflicts", "products_created", "kernel_launches")] + [("device_ms", c_f64), ("ms_diffuse", c_f64),
                                                    ("ms_resolve", c_f64), ("ms_sort", c_f64),
                                                    ("profiled_iterations", c_u64), ("ms_diffuse_slow", c_f64),
                                                    ("deferred_molecules", c_u64), ("deferred_by_reason", c_u64 * 8)]

    def as_dict(self):
        return {n: (list(getattr(self, n)) if n == "deferred_by_reason" else getattr(self, n)) for n, _ in self._fields_}


class mcx_trace_rec(C.Structure):
    _fields_ = [("id", c_u32), ("outcome", c_u32), ("n_words", c_u32), ("n_wall_hits", c_u32),
                ("n_collisions", c_u32), ("n_redo", c_u32),
                ("wall", c_u32 * MCX_TRACE_K), ("wall_side", c_u32 * MCX_TRACE_K),
                ("partner", c_u32 * MCX_TRACE_K),
                ("rxn_class", c_u32), ("rxn_pathway", c_u32), ("rxn_partner", c_u32), ("rounds", c_u32),
                ("event_hash", c_u64), ("pos", c_f64 * 3), ("t_event", c_f64)]


import numpy as _np

TRACE_DTYPE = _np.dtype([
    ("id", "<u4"), ("outcome", "<u4"), ("n_words", "<u4"), ("n_wall_hits", "<u4"),
    ("n_collisions", "<u4"), ("n_redo", "<u4"),
    ("wall", "<u4", (MCX_TRACE_K,)), ("wall_side", "<u4", (MCX_TRACE_K,)), ("partner", "<u4", (MCX_TRACE_K,)),
    ("rxn_class", "<u4"), ("rxn_pathway", "<u4"), ("rxn_partner", "<u4"), ("rounds", "<u4"),
    ("event_hash", "<u8"), ("pos", "<f8", (3,)), ("t_event", "<f8")])
assert TRACE_DTYPE.itemsize == C.sizeof(mcx_trace_rec), (TRACE_DTYPE.itemsize, C.sizeof(mcx_trace_rec))

# every symbol include/mcx.h declares (tests check the built library exports each of them)
EXPORTED_SYMBOLS = [
    "mcx_create", "mcx_destroy", "mcx_last_error", "mcx_abi_version", "mcx_set_geometry",
    "mcx_set_species", "mcx_set_reactions", "mcx_set_surface_classes", "mcx_upload_molecules",
    "mcx_download_molecules", "mcx_num_molecules", "mcx_step", "mcx_replay_step", "mcx_trace_step",
    "mcx_counts", "mcx_comm_init", "mcx_comm_unique_id", "mcx_slab_info_get", "mcx_comm_halo_path", "mcx_philox_block", "mcx_set_profiling",
    "mcx_grid_num_tiles", "mcx_grid2uv", "mcx_xyz2grid", "mcx_set_counted_volumes", "mcx_counts_by_volume", "mcx_release_volume_molecules", "mcx_fast_pass_kind", "mcx_walls_per_subpart",
    "mcx_set_surface_regions", "mcx_counts_by_surface_region", "mcx_release_list", "mcx_release_surface_molecules", "mcx_set_region_borders", "mcx_set_counted_volume_objects", "mcx_get_next_molecule_id", "mcx_set_next_molecule_id", "mcx_tile_neighbor_table", "mcx_get_wall_grids", "mcx_set_wall_grids",
]


def ptr(a, ctype):
    """numpy array -> typed pointer (None passes NULL)."""
    if a is None:
        return C.cast(None, P(ctype))
    return a.ctypes.data_as(P(ctype))
