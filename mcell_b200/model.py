"""Host-side model description and table builder.

Mirrors the part of pymcell4's Model/Config/Species/ReactionRule vocabulary that feeds the
diffuse-and-react hot path, and derives the flat tables the C ABI takes, with the arithmetic the
reference keeps in libbng / MCell3:

* units            libmcell/api/mcell4_converter.cpp:237-256  (length_unit = 1/sqrt(grid_density),
                   time_unit = time_step, default interaction radius 1/sqrt(pi*density) um)
* partition        mcell4_converter.cpp:280-390 (centred origin aligned to subpartition length)
* space_step       src/mcell_species.c:270-273   sqrt(4*1e8*D*time_unit)/length_unit
* bimol pb_factor  src/react_util.c:163-181       1e15/N_AV / (2*sqrt(pi)*R_um^2*eff_vel)
* unimol           src/react_util.c:74-78         k * time_unit
* vol-surf         src/react_util.c:111-157       1e11*grid_density/(2*N_AV)*sqrt(pi*t_step/D), x2 when both
                   reactants carry orientations of the same class
* cum_probs        src/mcell_reactions.c:2932-2933
* create_box / create_icosphere  libmcell/api/geometry_utils.cpp:32-91,124-262
"""
import ctypes as C
import math
from dataclasses import dataclass, field

import numpy as np

from . import abi

N_AV = 6.0221417930e23  # src/mcell_structs_shared.h:18
MY_PI = 3.14159265358979323846
MAX_SUBPARTS_PER_PARTITION = 300  # src4/defines.h:159


@dataclass
class Config:
    """libmcell/definition/simulation_setup.yaml defaults."""
    seed: int = 1
    time_step: float = 1e-6
    surface_grid_density: float = 10000.0
    interaction_radius: float = None
    partition_dimension: float = 10.0
    subpartition_dimension: float = 0.5
    initial_partition_origin: tuple = None


@dataclass
class Species:
    name: str
    diffusion_constant_3d: float = 0.0
    target_only: bool = False
    surface: bool = False          # surface molecule; diffusion_constant_3d then holds diffusion_constant_2d


def _parse_oriented(name):
    """MCell3 orientation marks on a species name: "L'" = up (+1), "R," = down (-1), none = 0."""
    if name.endswith("'"):
        return name[:-1], 1
    if name.endswith(","):
        return name[:-1], -1
    return name, 0


@dataclass
class ReactionRule:
    """reactants/products are species names; fwd_rate in 1/s (unimol) or 1/(M*s) (bimol)."""
    reactants: list
    products: list
    fwd_rate: float
    name: str = ""


@dataclass
class SurfaceProperty:
    surf_class: int
    type: int                      # abi.MCX_SURF_*
    species: str = None            # None = ALL_MOLECULES
    orientation: int = 0           # 0 any, +1 front, -1 back


def create_box(edge_um):
    """Vertex/face tables of geometry_utils.create_box (same order as CellBlender)."""
    h = edge_um / 2
    v = np.array([[-h, -h, -h], [-h, -h, h], [-h, h, -h], [-h, h, h],
                  [h, -h, -h], [h, -h, h], [h, h, -h], [h, h, h]], dtype=np.float64)
    f = np.array([[1, 2, 0], [3, 6, 2], [7, 4, 6], [5, 0, 4], [6, 0, 2], [3, 5, 7],
                  [1, 3, 2], [3, 7, 6], [7, 5, 4], [5, 1, 0], [6, 4, 0], [3, 1, 5]], dtype=np.uint32)
    return v, f


def create_icosphere(radius_um, subdivisions):
    """geometry_utils.create_icosphere: 20*4^(subdivisions-1) faces, watertight."""
    if not 1 <= subdivisions <= 8:
        raise ValueError("subdivisions must be in [1, 8]")
    t = (1.0 + math.sqrt(5.0)) / 2.0

    def norm(a):
        l = 1.0 / math.sqrt(a[0] * a[0] + a[1] * a[1] + a[2] * a[2])
        return (a[0] * l, a[1] * l, a[2] * l)

    verts = [norm(p) for p in [(-1, t, 0), (1, t, 0), (-1, -t, 0), (1, -t, 0), (0, -1, t), (0, 1, t),
                               (0, -1, -t), (0, 1, -t), (t, 0, -1), (t, 0, 1), (-t, 0, -1), (-t, 0, 1)]]
    faces = [(0, 11, 5), (0, 5, 1), (0, 1, 7), (0, 7, 10), (0, 10, 11), (1, 5, 9), (5, 11, 4), (11, 10, 2),
             (10, 7, 6), (7, 1, 8), (3, 9, 4), (3, 4, 2), (3, 2, 6), (3, 6, 8), (3, 8, 9), (4, 9, 5),
             (2, 4, 11), (6, 2, 10), (8, 6, 7), (9, 8, 1)]
    for _ in range(subdivisions - 1):
        div = {}
        out = []

        def mid(a, b):
            key = (a, b) if a < b else (b, a)
            if key in div:
                return div[key]
            va, vb = verts[a], verts[b]
            verts.append(norm((0.5 * (va[0] + vb[0]), 0.5 * (va[1] + vb[1]), 0.5 * (va[2] + vb[2]))))
            div[key] = len(verts) - 1
            return div[key]

        for f0, f1, f2 in faces:
            f3, f4, f5 = mid(f0, f1), mid(f1, f2), mid(f2, f0)
            out += [(f0, f3, f5), (f3, f1, f4), (f4, f2, f5), (f3, f4, f5)]
        faces = out
    v = np.array(verts, dtype=np.float64) * radius_um
    return v, np.array(faces, dtype=np.uint32)


@dataclass
class Tables:
    """Flat tables in ABI form; the same object drives libmcx and (in tests) the oracle."""
    cfg: abi.mcx_config
    species: object
    classes: object
    pathways: object
    surf_rules: object
    vertices: np.ndarray       # (nv,3) length units
    tri: np.ndarray            # (nw,3)
    wall_surf_class: np.ndarray
    species_names: list = field(default_factory=list)
    rule_names: list = field(default_factory=list)
    length_unit: float = 0.01
    time_unit: float = 1e-6
    # counted volumes (World::init_counted_volumes): index 0 = outside every counted object
    n_counted_volumes: int = 1
    wall_cv_front: np.ndarray = None
    wall_cv_back: np.ndarray = None
    wall_object: np.ndarray = None
    counted_volume_sets: list = field(default_factory=lambda: [frozenset()])
    cv_object_mask: np.ndarray = None      # per counted volume: bit k = counted object k encloses it
    cv_intersecting: int = 0               # counted objects that intersect another counted object
    # counted surface regions (MolOrRxnCountTerm region expressions over wall regions): set 0 = no counted region
    n_region_sets: int = 1
    wall_region_set: np.ndarray = None
    region_sets: list = field(default_factory=lambda: [frozenset()])
    region_names: list = field(default_factory=list)
    wall_edge_border: np.ndarray = None   # per wall: bit e = edge e is a border of a reactive region


class Model:
    def __init__(self, config=None):
        self.config = config or Config()
        self.species = []
        self.rules = []
        self.surface_properties = []
        self._verts = []
        self._tris = []
        self._wall_class = []
        self._counted = []
        self._regions = []   # (name, object index, face indices of that object)
        self._region_class = []
        self.wall_rules = []  # (surf class, reactant, products, rate, name): finite-rate surface-class reactions

    # -- subsystem ------------------------------------------------------------------------
    def add_species(self, name, D, target_only=False, surface=False):
        self.species.append(Species(name, D, target_only, surface))
        return len(self.species) - 1

    def add_reaction_rule(self, reactants, products, fwd_rate, name=""):
        self.rules.append(ReactionRule(list(reactants), list(products), fwd_rate, name or
                                       ("+".join(reactants) + "->" + "+".join(products))))

    def add_surface_property(self, surf_class, type_, species=None, orientation=0):
        self.surface_properties.append(SurfaceProperty(surf_class, type_, species, orientation))

    def add_surface_class_reaction(self, surf_class, reactant, products, fwd_rate, name=""):
        """A finite-rate reaction of a volume species with a surface class (MCell's "A' @ sc -> ...": a Standard reaction
        with a reactive surface, collide_and_react_with_walls / test_intersect / outcome_intersect).  reactant: "A'"
        (hits on the front), "A," (back) or "A" (both sides); products: volume species with orientation marks — the
        side of the wall they appear on; the reactant itself among the products is kept: with its own mark it
        reflects, with the other one it crosses the wall."""
        self.wall_rules.append((int(surf_class), reactant, list(products), float(fwd_rate),
                                name or (reactant + "@sc%d->" % surf_class + "+".join(products))))

    # -- instantiation --------------------------------------------------------------------
    def add_geometry_object(self, vertices_um, faces, surf_class=abi.MCX_NONE, counted=False):
        """surf_class: scalar or per-face array.  counted: the (closed) object is a counted volume
        (GeometryObject::is_counted_volume_or_compartment)."""
        self._counted.append(bool(counted))
        base = sum(len(v) for v in self._verts)
        self._verts.append(np.asarray(vertices_um, dtype=np.float64))
        self._tris.append(np.asarray(faces, dtype=np.uint32) + np.uint32(base))
        sc = np.broadcast_to(np.asarray(surf_class, dtype=np.uint32), (len(faces),)).copy()
        self._wall_class.append(sc)

    def add_surface_region(self, name, object_index, faces, surf_class=None):
        """A named surface region = some faces of one geometry object (Region, src4/region.h); used by surface-region
        counts (CountType::PresentOnSurfaceRegion / RxnCountOnSurfaceRegion).  surf_class: the region is reactive — its
        faces get that surface class and its outline becomes a region border for surface molecules
        (Region::is_edge, WallUtils::is_wall_edge_region_border).  Returns the region's index."""
        faces = np.asarray(faces, dtype=np.int64)
        self._regions.append((name, int(object_index), faces))
        self._region_class.append(surf_class)
        if surf_class is not None:
            self._wall_class[int(object_index)][faces] = np.uint32(surf_class)
        return len(self._regions) - 1

    # -- derived units ----------------------------------------------------------------------
    @property
    def length_unit(self):
        return 1.0 / math.sqrt(self.config.surface_grid_density)

    @property
    def rxn_radius_um(self):
        c = self.config
        return c.interaction_radius if c.interaction_radius is not None else 1.0 / math.sqrt(MY_PI * c.surface_grid_density)

    def space_step(self, D):
        return math.sqrt(4.0 * 1.0e8 * D * self.config.time_step) / self.length_unit

    def _partition(self):
        """mcell4_converter.cpp:280-390 (no explicit origin; geometry bbox + 0.01 um margin)."""
        c, lu = self.config, self.length_unit
        edge = c.partition_dimension / lu
        origin = np.array([-edge / 2] * 3)
        if self._verts:
            allv = np.concatenate(self._verts)
            llf, urb = allv.min(0) - 0.01, allv.max(0) + 0.01
            auto_dim = float(urb.max() - llf.min())
            if c.initial_partition_origin is None and auto_dim > c.partition_dimension:
                edge = auto_dim / lu
                origin = llf / lu
        if c.initial_partition_origin is not None:
            origin = np.array(c.initial_partition_origin, dtype=np.float64) / lu
        sp_len = c.subpartition_dimension / lu
        if int(edge / sp_len) > MAX_SUBPARTS_PER_PARTITION:
            sp_len = edge / MAX_SUBPARTS_PER_PARTITION
        eps = 1e-12

        def floor_mult(v):
            if v >= 0:
                return float(int((v + eps) / sp_len)) * sp_len
            return float(int((v + eps - sp_len) / sp_len)) * sp_len

        new_origin = np.array([floor_mult(v) for v in origin])
        edge_enlarged = edge + float((origin - new_origin).max())
        res = float(int((edge_enlarged + eps) / sp_len)) * sp_len
        if not abs(edge_enlarged - res) < eps:
            res += sp_len
        n = int(round(res / sp_len))
        return new_origin, res, n

    def build(self, max_molecules=0, device=0, rng_mode=abi.MCX_RNG_PHILOX, cell_edge=0.0,
              rank=0, world_size=1, max_resolve_rounds=0):
        c, lu = self.config, self.length_unit
        names = [s.name for s in self.species]
        idx = {n: i for i, n in enumerate(names)}
        origin, edge, n_sub = self._partition()

        sp = (abi.mcx_species * max(1, len(self.species)))()
        for i, s in enumerate(self.species):
            sp[i].space_step = self.space_step(s.diffusion_constant_3d)
            sp[i].time_step = 1.0
            sp[i].flags = (0 if s.surface else abi.MCX_SP_VOL) | (abi.MCX_SP_CAN_DIFFUSE if s.diffusion_constant_3d > 0 else 0) | \
                (abi.MCX_SP_CANT_INITIATE if s.target_only else 0)

        # group rules into reaction classes (same reactant set), cumulative probabilities
        groups = {}
        for r_id, r in enumerate(self.rules):
            key = tuple(sorted(idx[_parse_oriented(x)[0]] for x in r.reactants))
            groups.setdefault(key, []).append((r_id, r))
        # the reference keeps one reaction class per reactant geometry; the device tables hold one class per species
        # pair, so rules of one pair must agree in their reactant orientations
        for key, rules in groups.items():
            geoms = {tuple(sorted((idx[n], o) for n, o in (_parse_oriented(x) for x in r.reactants))) for _, r in rules}
            if len(geoms) > 1:
                raise ValueError("rules on the species pair %s differ in reactant orientation (separate reaction classes "
                                 "per geometry are not built)" % (tuple(self.species[k].name for k in key),))
        # finite-rate surface-class reactions: one class per (species, side, surface class), behind the ordinary classes
        wall_groups = {}
        for w_id, (sc, reactant, prods, rate, _) in enumerate(self.wall_rules):
            rn, ro = _parse_oriented(reactant)
            wall_groups.setdefault((idx[rn], ro, sc), []).append((len(self.rules) + w_id, prods, rate))
        classes = (abi.mcx_rxn_class * max(1, len(groups) + len(wall_groups)))()
        n_path = sum(len(v) for v in groups.values()) + sum(len(v) for v in wall_groups.values())
        pathways = (abi.mcx_pathway * max(1, n_path))()
        pi = 0
        has_bimol = False
        for ci, (key, rules) in enumerate(groups.items()):
            rc = classes[ci]
            first = rules[0][1]
            parsed = [_parse_oriented(x) for x in first.reactants]
            r_idx = [idx[n] for n, _ in parsed]
            r_orient = [o for _, o in parsed]
            n_surf_reactants = sum(1 for k in r_idx if self.species[k].surface)
            if len(key) == 2 and n_surf_reactants == 1:
                # class order is (volume, surface) whatever order the rule was written in
                if self.species[r_idx[0]].surface:
                    r_idx, r_orient = r_idx[::-1], r_orient[::-1]
                rc.kind = abi.MCX_RXN_BIMOL_VOLSURF
                rc.reactants[0], rc.reactants[1] = r_idx[0], r_idx[1]
                rc.reactant_orientation[0], rc.reactant_orientation[1] = r_orient[0], r_orient[1]
                D_tot = self.species[r_idx[0]].diffusion_constant_3d
                t_step = 1.0 * c.time_step
                pb_factor = 0.0 if D_tot <= 0 else 1.0e11 * c.surface_grid_density / (2.0 * N_AV) * math.sqrt(MY_PI * t_step / D_tot)
                if (r_orient[0] + r_orient[1]) * (r_orient[0] - r_orient[1]) == 0 and r_orient[0] * r_orient[1] != 0:
                    pb_factor *= 2.0
            elif len(key) == 2 and n_surf_reactants == 2:
                # two surface molecules (src/react_util.c:84-100): 3 neighbours, both molecules may initiate
                rc.kind = abi.MCX_RXN_BIMOL_SURFSURF
                rc.reactants[0], rc.reactants[1] = r_idx[0], r_idx[1]
                rc.reactant_orientation[0], rc.reactant_orientation[1] = r_orient[0], r_orient[1]
                sa, sb = self.species[r_idx[0]], self.species[r_idx[1]]
                if sa.target_only and sb.target_only:
                    raise ValueError("both reactants TARGET_ONLY")
                pb_factor = c.time_step * c.surface_grid_density / (3.0 if (sa.target_only or sb.target_only) else 6.0)
            elif len(key) == 1:
                rc.kind = abi.MCX_RXN_UNIMOL
                rc.reactants[0], rc.reactants[1] = r_idx[0], abi.MCX_NONE
                pb_factor = c.time_step
            else:
                has_bimol = True
                rc.kind = abi.MCX_RXN_BIMOL_VOLVOL
                rc.reactants[0], rc.reactants[1] = r_idx[0], r_idx[1]
                sa, sb = self.species[r_idx[0]], self.species[r_idx[1]]
                eff_a = self.space_step(sa.diffusion_constant_3d) / 1.0
                eff_b = self.space_step(sb.diffusion_constant_3d) / 1.0
                if sa.target_only and sb.target_only:
                    raise ValueError("both reactants TARGET_ONLY")
                if sa.target_only:
                    eff_a = 0
                elif sb.target_only:
                    eff_b = 0
                if eff_a + eff_b > 0:
                    eff_vel = (eff_a + eff_b) * lu / c.time_step
                    R = self.rxn_radius_um
                    pb_factor = 1.0 / (2.0 * math.sqrt(MY_PI) * R * R * eff_vel)
                    pb_factor *= 1.0e15 / N_AV
                else:
                    pb_factor = 0.0
            rc.first_pathway, rc.n_pathways = pi, len(rules)
            cum = 0.0
            for r_id, r in rules:
                pw = pathways[pi]
                cum += pb_factor * r.fwd_rate
                pw.cum_prob = cum
                # kept reactants: rule reactant that re-appears unchanged among the products
                pparsed = [_parse_oriented(x) for x in r.products]
                prods = [idx[n] for n, _ in pparsed]
                porient = [o for _, o in pparsed]
                keep = 0
                kept_at = {}   # position among the rule's products -> kept reactant (class order)
                for k, rs in enumerate(r_idx):
                    for q, ps in enumerate(prods):
                        if ps == rs and q not in kept_at:
                            kept_at[q] = k
                            keep |= 1 << k
                            break
                new_at = [q for q in range(len(prods)) if q not in kept_at]
                if len(new_at) > abi.MCX_MAX_PRODUCTS:
                    raise ValueError("too many products")
                pw.n_products = len(new_at)
                for k, q in enumerate(new_at):
                    pw.products[k] = prods[q]
                    pw.product_orientation[k] = porient[q]
                pw.keep_reactant_mask = keep
                # rule order of the products (= order of the orientation draws, diffuse_react_event.cpp:2618-2627) and the
                # product-side orientation of the kept reactants (:2689-2716)
                info = abi.MCX_KEPT_VALID
                for q in range(6):
                    if q >= len(prods):
                        nib = abi.MCX_KEPT_ORDER_END
                    elif q in kept_at:
                        nib = abi.MCX_KEPT_ORDER_REACTANT + kept_at[q]
                        info |= {0: 0, 1: 1, -1: 2}[porient[q]] << (24 + 2 * kept_at[q])
                    else:
                        nib = new_at.index(q)
                    info |= nib << (4 * q)
                pw.kept_info = info
                pw.rxn_rule_id = r_id
                pi += 1
            rc.max_fixed_p = cum

        wall_class_rules = []
        for wi, ((sp_i, ro, sc), wrules) in enumerate(wall_groups.items()):
            ci = len(groups) + wi
            rc = classes[ci]
            rc.kind = abi.MCX_RXN_BIMOL_VOLWALL
            rc.reactants[0], rc.reactants[1] = sp_i, sc
            rc.reactant_orientation[0], rc.reactant_orientation[1] = ro, 1
            # src/react_util.c:110-157: the vol-wall factor; doubled when the molecule carries the mark of the surface class
            D_tot = self.species[sp_i].diffusion_constant_3d
            pb_factor = 0.0 if D_tot <= 0 else 1.0e11 * c.surface_grid_density / (2.0 * N_AV) * math.sqrt(MY_PI * c.time_step / D_tot)
            if ro != 0:
                pb_factor *= 2.0
            rc.first_pathway, rc.n_pathways = pi, len(wrules)
            cum = 0.0
            for r_id, prods_s, rate in wrules:
                pw = pathways[pi]
                cum += pb_factor * rate
                pw.cum_prob = cum
                pparsed = [_parse_oriented(x) for x in prods_s]
                prods = [idx[n] for n, _ in pparsed]
                porient = [o for _, o in pparsed]
                if any(self.species[q].surface for q in prods):
                    raise ValueError("surface products of a surface-class reaction are not built")
                kept_q = prods.index(sp_i) if sp_i in prods else None
                new_at = [q for q in range(len(prods)) if q != kept_q]
                if len(new_at) > abi.MCX_MAX_PRODUCTS:
                    raise ValueError("too many products")
                pw.n_products = len(new_at)
                for k, q in enumerate(new_at):
                    pw.products[k] = prods[q]
                    pw.product_orientation[k] = porient[q]
                pw.keep_reactant_mask = 0 if kept_q is None else 1
                info = abi.MCX_KEPT_VALID
                for q in range(6):
                    if q >= len(prods):
                        nib = abi.MCX_KEPT_ORDER_END
                    elif q == kept_q:
                        nib = abi.MCX_KEPT_ORDER_REACTANT
                        info |= {0: 0, 1: 1, -1: 2}[porient[q]] << 24
                    else:
                        nib = new_at.index(q)
                    info |= nib << (4 * q)
                pw.kept_info = info
                pw.rxn_rule_id = r_id
                pi += 1
            rc.max_fixed_p = cum
            wall_class_rules.append((sp_i, sc, ro, ci))

        rules = (abi.mcx_surf_class_rxn * max(1, len(self.surface_properties) + len(wall_class_rules)))()
        for i, s in enumerate(self.surface_properties):
            rules[i].species = abi.MCX_ALL_MOLECULES if s.species is None else idx[s.species]
            rules[i].surf_class, rules[i].orientation, rules[i].type = s.surf_class, s.orientation, s.type
        for k, (sp_i, sc, ro, ci) in enumerate(wall_class_rules):
            r_ = rules[len(self.surface_properties) + k]
            r_.species, r_.surf_class, r_.orientation, r_.type, r_.rxn_class = sp_i, sc, ro, abi.MCX_SURF_STANDARD, ci

        if self._verts:
            verts = np.concatenate(self._verts) / lu   # mcell4_converter.cpp:921-923
            tri = np.concatenate(self._tris)
            wsc = np.concatenate(self._wall_class)
        else:
            verts, tri, wsc = np.zeros((0, 3)), np.zeros((0, 3), np.uint32), np.zeros((0,), np.uint32)

        cfg = abi.mcx_config()
        cfg.abi_version = abi.MCX_ABI_VERSION
        cfg.device = device
        cfg.seed = c.seed
        for k in range(3):
            cfg.origin[k] = origin[k]
        cfg.partition_edge_length = edge
        cfg.num_subparts_per_edge = n_sub
        cfg.use_expanded_list = 1 if has_bimol else 0      # mcell4_converter.cpp:84-87
        cfg.rxn_radius_3d = self.rxn_radius_um / lu
        cfg.cell_edge = cell_edge
        if len(verts):
            llf, urb = verts.min(0), verts.max(0)
            for k in range(3):
                cfg.active_llf[k], cfg.active_urb[k] = llf[k], urb[k]
        cfg.max_molecules = max_molecules
        cfg.max_resolve_rounds = max_resolve_rounds
        cfg.rng_mode = rng_mode
        cfg.rank, cfg.world_size = rank, world_size
        # guard of mcell4_converter.cpp:121-127
        if has_bimol and cfg.rxn_radius_3d * math.sqrt(2.0) >= (edge / n_sub) / 2:
            raise ValueError("reaction radius too large for the subpartition size")
        t = Tables(cfg, sp, classes, pathways, rules, np.ascontiguousarray(verts, np.float64),
                   np.ascontiguousarray(tri, np.uint32), np.ascontiguousarray(wsc, np.uint32),
                   names, [r.name for r in self.rules] + [w_[4] for w_ in self.wall_rules], lu, c.time_step)
        t.wall_object = np.concatenate([np.full(len(f), k, np.uint32) for k, f in enumerate(self._tris)]) if self._tris else np.zeros(0, np.uint32)
        if any(self._counted):
            _assign_counted_volumes(t, self._counted)
        if self._regions:
            # every distinct set of regions a wall belongs to gets one index (set 0 = none), as Partition keeps the
            # regions of a wall (wall_matches_region_expr_recursively, mol_or_rxn_count_event.cpp)
            first_wall = np.cumsum([0] + [len(f) for f in self._tris])
            member = [set() for _ in range(len(tri))]
            for k, (_, obj, faces) in enumerate(self._regions):
                for f in faces:
                    member[first_wall[obj] + int(f)].add(k)
            sets = [frozenset()]
            wrs = np.zeros(len(tri), np.uint8)
            for wi, ms in enumerate(member):
                fs = frozenset(ms)
                if fs not in sets:
                    sets.append(fs)
                wrs[wi] = sets.index(fs)
            if len(sets) > 256:
                raise ValueError("more than 256 distinct sets of surface regions")
            t.region_sets, t.n_region_sets, t.wall_region_set = sets, len(sets), wrs
            t.region_names = [r[0] for r in self._regions]
            # borders of the reactive regions: an edge (e: v_e - v_{e+1}) of a region wall whose neighbour across it is not
            # in the region (or that has none)
            if any(c_ is not None for c_ in self._region_class):
                border = np.zeros(len(tri), np.uint8)
                for k, (_, obj, faces) in enumerate(self._regions):
                    if self._region_class[k] is None:
                        continue
                    walls = [int(first_wall[obj] + f) for f in faces]
                    count = {}
                    for wi in walls:
                        for e in range(3):
                            key = tuple(sorted((int(tri[wi][e]), int(tri[wi][(e + 1) % 3]))))
                            count[key] = count.get(key, 0) + 1
                    for wi in walls:
                        for e in range(3):
                            key = tuple(sorted((int(tri[wi][e]), int(tri[wi][(e + 1) % 3]))))
                            if count[key] == 1:
                                border[wi] |= 1 << e
                t.wall_edge_border = border
        t.n_species, t.n_classes, t.n_pathways = len(self.species), len(groups) + len(wall_groups), n_path
        t.n_surf_rules = len(self.surface_properties) + len(wall_class_rules)
        t.n_rules = len(self.rules) + len(self.wall_rules)
        return t


def points_inside_mesh(points, tri_xyz):
    """Parity ray cast along +x with a fixed irrational skew (closed meshes): boolean per point.  Host-side
    stand-in for VtkUtils::is_point_inside_counted_volume (vtk_utils.cpp:285-570)."""
    p = np.asarray(points, np.float64)
    d = np.array([1.0, 0.0137131, 0.0071393])
    v0, v1, v2 = tri_xyz[:, 0], tri_xyz[:, 1], tri_xyz[:, 2]
    e1, e2 = v1 - v0, v2 - v0
    h = np.cross(d, e2)
    a = (e1 * h).sum(1)
    inside = np.zeros(len(p), bool)
    ok = np.abs(a) > 1e-300
    f = np.where(ok, 1.0 / np.where(ok, a, 1.0), 0.0)
    for s0 in range(0, len(p), 4096):
        q = p[s0:s0 + 4096]
        sv = q[:, None, :] - v0[None]
        u = f[None] * (sv * h[None]).sum(2)
        qq = np.cross(sv, e1[None])
        v = f[None] * (qq * d[None, None]).sum(2)
        tt = f[None] * (qq * e2[None]).sum(2)
        hit = ok[None] & (u >= 0) & (v >= 0) & (u + v <= 1) & (tt > 0)
        inside[s0:s0 + 4096] = hit.sum(1) % 2 == 1
    return inside


def _assign_counted_volumes(t, counted):
    """Counted volumes of non-intersecting closed objects: every distinct set of enclosing counted objects is one
    volume (index 0 = none).  Per wall: the volume in front of it and behind it, probed a small step off the
    centroid along the normal."""
    tri = t.vertices[t.tri]
    cen = tri.mean(1)
    nrm = np.cross(tri[:, 1] - tri[:, 0], tri[:, 2] - tri[:, 0])
    ln = np.linalg.norm(nrm, axis=1)
    nrm = nrm / np.where(ln > 0, ln, 1.0)[:, None]
    step = 1e-3 * np.sqrt(np.maximum(ln, 1e-30))[:, None]

    def sets_of(points):
        member = np.zeros((len(points), len(counted)), bool)
        for k, c in enumerate(counted):
            if c:
                member[:, k] = points_inside_mesh(points, tri[t.wall_object == k])
        return [frozenset(np.flatnonzero(row).tolist()) for row in member]

    fs, bs = sets_of(cen + nrm * step), sets_of(cen - nrm * step)
    sets = [frozenset()]
    for s_ in fs + bs:
        if s_ not in sets:
            sets.append(s_)
    if len(sets) > 256:
        raise ValueError("more than 256 counted volumes")
    index = {s_: i for i, s_ in enumerate(sets)}
    t.counted_volume_sets = sets
    t.n_counted_volumes = len(sets)
    t.wall_cv_front = np.array([index[s_] for s_ in fs], np.uint8)
    t.wall_cv_back = np.array([index[s_] for s_ in bs], np.uint8)
    # counted objects that intersect another one: some of their vertices lie inside, some outside of it; their walls have no
    # single volume in front of / behind them (mcx_set_counted_volume_objects: membership toggles instead)
    t.cv_object_mask = np.array([sum(1 << k for k in s_) for s_ in sets], np.uint32) if len(counted) <= 32 else None
    inter = 0
    if t.cv_object_mask is not None:
        for a in range(len(counted)):
            if not counted[a]:
                continue
            va = np.unique(t.tri[t.wall_object == a])
            for b in range(len(counted)):
                if a == b or not counted[b]:
                    continue
                ins = points_inside_mesh(t.vertices[va], tri[t.wall_object == b])
                if ins.any() and not ins.all():
                    inter |= (1 << a) | (1 << b)
    t.cv_intersecting = inter


def counted_volume_of(t, positions):
    """Counted volume index of each position (Partition::add_volume_molecule computes it by a ray cast,
    collision_utils.inl:1515-1566)."""
    if t.n_counted_volumes <= 1:
        return np.zeros(len(positions), np.uint32)
    tri = t.vertices[t.tri]
    n_obj = int(t.wall_object.max()) + 1
    member = np.zeros((len(positions), n_obj), bool)
    used = set().union(*t.counted_volume_sets)
    for k in used:
        member[:, k] = points_inside_mesh(positions, tri[t.wall_object == k])
    index = {s_: i for i, s_ in enumerate(t.counted_volume_sets)}
    return np.array([index[frozenset(np.flatnonzero(row).tolist())] for row in member], np.uint32)


def release_uniform_box(rng, n, edge_um, length_unit, margin=0.0):
    """Uniform release inside a centred cube (release_event.cpp:904-1003 semantics: three uniform
    draws per molecule).  rng: numpy Generator.  Returns (n,3) positions in length units."""
    h = (edge_um / 2 - margin) / length_unit
    return rng.uniform(-h, h, size=(n, 3))


def release_on_walls(rng, tables, walls, n, species, orientation=1, first_id=0, schedule_unimol=True):
    """Host-side placement of n surface molecules on distinct random tiles of the given walls, one per tile, at the
    tile centres (release_event.cpp:1063-1197 places by density on vacant tiles; the device path is "next" row f3).
    Uses libmcx's host helpers for the grid arithmetic (Grid::initialize, GridUtils::grid2uv).  rng: numpy Generator."""
    from . import engine
    L = engine.load_library()
    L.mcx_grid_num_tiles.argtypes = [C.c_void_p]
    L.mcx_grid_num_tiles.restype = C.c_uint32
    L.mcx_grid2uv.argtypes = [C.c_void_p, C.c_uint32, C.c_void_p]
    L.mcx_grid2uv.restype = None
    walls = np.asarray(walls, np.uint32)
    v9 = np.ascontiguousarray(tables.vertices[tables.tri[walls]].reshape(len(walls), 9))
    counts = np.array([L.mcx_grid_num_tiles(C.c_void_p(v9[k].ctypes.data)) for k in range(len(walls))], np.int64)
    total = int(counts.sum())
    if n > total:
        raise ValueError("more surface molecules than tiles")
    pick = np.sort(rng.choice(total, size=n, replace=False))
    start = np.concatenate([[0], np.cumsum(counts)])
    wi = np.searchsorted(start, pick, side="right") - 1
    m = MolArrays(n)
    uv = np.zeros(2)
    for k in range(n):
        t = int(pick[k] - start[wi[k]])
        L.mcx_grid2uv(C.c_void_p(v9[wi[k]].ctypes.data), t, C.c_void_p(uv.ctypes.data))
        m.wall[k], m.tile[k], m.u[k], m.v[k] = walls[wi[k]], t, uv[0], uv[1]
    m.orientation[:] = orientation
    m.species[:] = species
    m.id[:] = np.arange(first_id, first_id + n, dtype=np.uint32)
    if schedule_unimol:
        m.flags[:] = abi.MCX_MOL_SCHEDULE_UNIMOL
    return m


class MolArrays:
    """Owning numpy SoA + the ctypes view (mcx_mol_soa)."""
    FIELDS = ("x", "y", "z", "id", "species", "flags", "diffusion_time", "unimol_rxn_time",
              "wall", "tile", "orientation", "u", "v", "counted_volume")

    def __init__(self, n, with_times=True):
        self.x = np.zeros(n); self.y = np.zeros(n); self.z = np.zeros(n)
        self.id = np.zeros(n, np.uint32); self.species = np.zeros(n, np.uint32); self.flags = np.zeros(n, np.uint32)
        self.diffusion_time = np.zeros(n) if with_times else None
        self.unimol_rxn_time = np.full(n, abi.MCX_TIME_INVALID) if with_times else None
        # Molecule::s of surface molecules (wall == MCX_NONE: volume molecule)
        self.wall = np.full(n, abi.MCX_NONE, np.uint32); self.tile = np.full(n, abi.MCX_NONE, np.uint32)
        self.orientation = np.zeros(n, np.int32); self.u = np.zeros(n); self.v = np.zeros(n)
        self.counted_volume = np.zeros(n, np.uint32)
        self.n = n

    @classmethod
    def from_positions(cls, pos, species, first_id=0, schedule_unimol=False, iteration=0):
        n = len(pos)
        m = cls(n)
        m.x[:], m.y[:], m.z[:] = pos[:, 0], pos[:, 1], pos[:, 2]
        m.id[:] = np.arange(first_id, first_id + n, dtype=np.uint32)
        m.species[:] = species
        m.diffusion_time[:] = iteration
        if schedule_unimol:
            m.flags[:] = abi.MCX_MOL_SCHEDULE_UNIMOL
        return m

    def view(self, n=None):
        s = abi.mcx_mol_soa()
        s.n = self.n if n is None else n
        s.x, s.y, s.z = abi.ptr(self.x, C.c_double), abi.ptr(self.y, C.c_double), abi.ptr(self.z, C.c_double)
        s.id, s.species, s.flags = abi.ptr(self.id, C.c_uint32), abi.ptr(self.species, C.c_uint32), abi.ptr(self.flags, C.c_uint32)
        s.diffusion_time = abi.ptr(self.diffusion_time, C.c_double)
        s.unimol_rxn_time = abi.ptr(self.unimol_rxn_time, C.c_double)
        # surface arrays are optional in the ABI: views assembled by hand (bench.py) may not carry them
        if self.wall is not None and len(self.wall) >= len(self.x) and len(self.wall) > 0:
            s.wall, s.tile = abi.ptr(self.wall, C.c_uint32), abi.ptr(self.tile, C.c_uint32)
            s.orientation = abi.ptr(self.orientation, C.c_int32)
            s.u, s.v = abi.ptr(self.u, C.c_double), abi.ptr(self.v, C.c_double)
            s.counted_volume = abi.ptr(self.counted_volume, C.c_uint32)
        return s

    def truncated(self, n):
        m = MolArrays(0)
        for k in self.FIELDS:
            setattr(m, k, getattr(self, k)[:n].copy())
        m.n = n
        return m

    def sorted_by_id(self):
        o = np.argsort(self.id[:self.n], kind="stable")
        m = MolArrays(0)
        for k in self.FIELDS:
            setattr(m, k, getattr(self, k)[:self.n][o].copy())
        m.n = self.n
        return m

    @classmethod
    def concat(cls, parts):
        """One population from several (ids must already be distinct)."""
        m = cls(0)
        for k in cls.FIELDS:
            setattr(m, k, np.concatenate([getattr(q, k)[:q.n] for q in parts]))
        m.n = sum(q.n for q in parts)
        return m
