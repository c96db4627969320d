"""CPU tier: PINS the geometry of surface diffusion against the REFERENCE'S OWN compiled code.

tests/golden/mcell3_surface_vectors.npz holds outputs of oracle/_ref/libmcell3ref.so — surface_net,
init_edge_transform, find_edge_point and traverse_surface of the reference's src/wall_util.c and ray_trace_2D of its
src/diffuse.c, compiled unmodified (the
MCell3 originals of src4/geometry.cpp:258-356, wall.cpp:134-235, geometry_utils.inl:222-342; oracle/Makefile: ref) —
on the cases of tests/golden/mcell3_surface_cases.py: closed and open meshes (free edges), regular and irregular
triangles, moves that stay inside, leave through each edge, start on an edge or a vertex, run along an edge or through
a corner.  The oracle must reproduce neighbour walls, forward/backward roles, every transform and every crossing point
BIT FOR BIT.  Where the compiled reference is present it is additionally driven live on fresh cases."""
import os
import sys

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "golden"))
import mcell3_surface_cases as sc  # noqa: E402
from oracle import oracle_py as O  # noqa: E402

G = np.load(os.path.join(HERE, "golden", "mcell3_surface_vectors.npz"))


def test_edge_pairing_and_transforms_bit_exact():
    L = O.lib()
    for name, (v, f) in sc.meshes().items():
        nb, fw, tr = O.mesh_edges(L.orc_unit_mesh_edges, v, f)
        assert (nb == G[name + "_nb"]).all(), name
        assert (fw == G[name + "_fw"]).all(), name
        assert (tr == G[name + "_tr"]).all(), name
    assert (G["open_box_nb"] < 0).sum() == 7 and (G["icosphere4_nb"] >= 0).all()


def test_traverse_surface_bit_exact():
    L = O.lib()
    for name, (v, f) in sc.meshes().items():
        qw, qs, quv = sc.traverse_queries(len(f))
        tw, tuv = O.traverse_surface(L.orc_unit_traverse_surface, v, f, qw, qs, quv)
        assert (tw == G[name + "_tw"]).all(), name
        assert (tuv == G[name + "_tuv"]).all(), name


def test_ray_trace_surf_bit_exact_against_compiled_ray_trace_2d():
    """The whole walk of a 2-D move — find_edge_point, traverse_surface of position and displacement, reflection at free
    edges, the loop around them — against the reference's ray_trace_2D (src/diffuse.c; species without region-border
    reactions, no periodic box)."""
    L = O.lib()
    crossed = reflected_mesh = 0
    for name, (v, f) in sc.meshes().items():
        qw, quv, qdisp = sc.ray_queries(v, f)
        w, uv = O.ray_trace_surf(L.orc_unit_ray_trace_surf, v, f, qw, quv, qdisp)
        assert (w == G[name + "_rayw"]).all(), name
        assert (uv == G[name + "_rayuv"]).all(), name
        crossed += int((w != qw.astype(np.int32)).sum())
        reflected_mesh += name == "open_box"
    assert crossed > 5000 and reflected_mesh == 1


def test_uv2grid_exact():
    """Tile under a uv point (GridUtils::uv2grid_tile_index): where a surface molecule lands after a 2-D move."""
    import ctypes as C
    L = O.lib()
    big = sc.triangles(seed=14, n=40) * 5.0
    pts = sc.uv_points(big)
    assert len(pts) == len(G["uv2grid"])
    for i, (ti, uv) in enumerate(pts):
        assert L.orc_unit_uv2grid(C.c_void_p(big[ti].ctypes.data), C.c_void_p(uv.ctypes.data)) == int(G["uv2grid"][i]), i
    assert G["uv2grid"].max() > 200


def test_find_edge_point_bit_exact_all_outcomes():
    L = O.lib()
    tris = sc.triangles()
    moves = sc.edge_moves(tris)
    assert len(moves) == len(G["fep_code"])
    for i, (ti, loc, disp) in enumerate(moves):
        code, pt = O.find_edge_point(L.orc_unit_find_edge_point, tris[ti], loc, disp)
        assert code == int(G["fep_code"][i]), i
        if code >= 0:
            assert (pt == G["fep_pt"][i]).all(), i
    assert set(np.unique(G["fep_code"]).tolist()) == {-2, -1, 0, 1, 2}


def test_surface_geometry_live_against_compiled_reference():
    R3 = O.ref_mcell3_lib()
    if R3 is None or not hasattr(R3, "ref3_mesh_edges"):
        pytest.skip("oracle/_ref/libmcell3ref.so (with the mesh entry points) not built here")
    L = O.lib()
    rng = np.random.default_rng(77)
    from mcell_b200.model import create_icosphere
    for trial in range(3):
        v, f = create_icosphere(0.4, 3)
        v = np.ascontiguousarray(v * 100.0 * (1.0 + 0.3 * rng.uniform(-1, 1, (len(v), 1))) + rng.uniform(-20, 20, 3))
        f = np.ascontiguousarray(f[rng.permutation(len(f))[: len(f) - 5 * trial]], np.uint32)   # permuted, some faces missing
        a = O.mesh_edges(R3.ref3_mesh_edges, v, f)
        b = O.mesh_edges(L.orc_unit_mesh_edges, v, f)
        assert all((x == y).all() for x, y in zip(a, b)), trial
        qw, qs, quv = sc.traverse_queries(len(f), seed=100 + trial)
        ta = O.traverse_surface(R3.ref3_traverse_surface, v, f, qw, qs, quv)
        tb = O.traverse_surface(L.orc_unit_traverse_surface, v, f, qw, qs, quv)
        assert (ta[0] == tb[0]).all() and (ta[1] == tb[1]).all(), trial
        rw, ruv, rdisp = sc.ray_queries(v, f, seed=200 + trial, n=2000)
        ra = O.ray_trace_surf(R3.ref3_ray_trace_2d, v, f, rw, ruv, rdisp)
        rb = O.ray_trace_surf(L.orc_unit_ray_trace_surf, v, f, rw, ruv, rdisp)
        assert (ra[0] == rb[0]).all() and (ra[1] == rb[1]).all(), trial
    tris = sc.triangles(seed=31, n=40)
    for ti, loc, disp in sc.edge_moves(tris, seed=32, per_tri=30):
        ca, pa = O.find_edge_point(R3.ref3_find_edge_point, tris[ti], loc, disp)
        cb, pb = O.find_edge_point(L.orc_unit_find_edge_point, tris[ti], loc, disp)
        assert ca == cb and (ca < 0 or (pa == pb).all()), (ti, loc, disp)
