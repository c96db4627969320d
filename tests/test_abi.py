"""CPU tier: the C-ABI library loads, exports every symbol include/mcx.h declares, its struct layouts match the
ctypes mirror, and compute entry points fail loudly (no CPU fallback) when no CUDA device is present."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from mcell_b200 import abi, engine

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _header_functions():
    src = open(os.path.join(ROOT, "include", "mcx.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(mcx_[a-z0-9_]+)\s*\(", src)))


def test_header_and_binding_list_agree():
    assert _header_functions() == sorted(abi.EXPORTED_SYMBOLS)


def test_library_exports_every_declared_symbol():
    L = engine.load_library()
    for name in _header_functions():
        assert hasattr(L, name), name
    assert L.mcx_abi_version() == abi.MCX_ABI_VERSION


def test_struct_layouts_match_library():
    L = engine.load_library()
    L.mcx_sizeof.restype = C.c_int
    structs = [abi.mcx_config, abi.mcx_species, abi.mcx_rxn_class, abi.mcx_pathway, abi.mcx_surf_class_rxn,
               abi.mcx_mol_soa, abi.mcx_step_stats, abi.mcx_trace_rec, abi.mcx_slab_info, abi.mcx_release,
               abi.mcx_surface_release]
    for i, s in enumerate(structs):
        assert L.mcx_sizeof(i) == C.sizeof(s), s.__name__
    assert L.mcx_sizeof(99) == -1


def test_header_constants_match_binding():
    src = open(os.path.join(ROOT, "include", "mcx.h")).read()
    for name in ("MCX_ABI_VERSION", "MCX_MAX_PRODUCTS", "MCX_TRACE_K"):
        m = re.search(r"#define\s+%s\s+(\d+)" % name, src)
        assert int(m.group(1)) == getattr(abi, name)
    for name in ("MCX_ERR_INVALID_ARG", "MCX_ERR_CUDA", "MCX_ERR_CAPACITY", "MCX_ERR_ESCAPED", "MCX_ERR_STATE",
                 "MCX_ERR_OVERFLOW", "MCX_ERR_COMM"):
        m = re.search(r"#define\s+%s\s+\((-\d+)\)" % name, src)
        assert int(m.group(1)) == getattr(abi, name)


def test_philox_host_helper_matches_oracle_restatement():
    from oracle import oracle_py as O
    L = O.lib()
    for seed, mid, it, blk in [(0, 0, 0, 0), (1, 7, 3, 2), (2**40 + 5, 2**32 - 1, 2**33, 9)]:
        a = engine.philox_block(seed, mid, it, blk)
        b = np.zeros(4, np.uint32)
        L.orc_philox_block(C.c_uint64(seed), C.c_uint32(mid), C.c_uint64(it), C.c_uint32(blk), C.c_void_p(b.ctypes.data))
        assert (a == b).all()


def test_philox_published_known_answers():
    """Random123 kat_vectors for philox4x32-10 (Salmon et al., SC'11)."""
    from oracle import oracle_py as O
    L = O.lib()

    def blk(ctr, key):
        # mcx layout: ctr = (block, it_lo, it_hi, mol_id), key = (seed_lo, seed_hi)
        seed = key[0] | (key[1] << 32)
        it = ctr[1] | (ctr[2] << 32)
        a = engine.philox_block(seed, ctr[3], it, ctr[0])
        b = np.zeros(4, np.uint32)
        L.orc_philox_block(C.c_uint64(seed), C.c_uint32(ctr[3]), C.c_uint64(it), C.c_uint32(ctr[0]), C.c_void_p(b.ctypes.data))
        assert (a == b).all()
        return [int(x) for x in a]

    assert blk((0, 0, 0, 0), (0, 0)) == [0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8]
    f = 0xffffffff
    assert blk((f, f, f, f), (f, f)) == [0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd]
    assert blk((0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344), (0xa4093822, 0x299f31d0)) == \
        [0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1]


def _no_gpu():
    try:
        import torch
        return not torch.cuda.is_available()
    except Exception:
        return True


@pytest.mark.skipif(not _no_gpu(), reason="only meaningful on a box without a CUDA device")
def test_no_cpu_fallback_create_fails_loudly():
    import common as cm
    t, mols = cm.free_diffusion_box(n=10)
    with pytest.raises(engine.McxError) as ei:
        engine.Engine(t)
    assert ei.value.code == abi.MCX_ERR_CUDA
    assert "no CPU fallback" in str(ei.value)


def test_create_rejects_bad_config_before_touching_the_device():
    L = engine.load_library()
    cfg = abi.mcx_config()
    h = C.c_void_p()
    assert L.mcx_create(C.byref(cfg), C.byref(h)) == abi.MCX_ERR_INVALID_ARG  # abi_version 0
    assert b"ABI" in L.mcx_last_error(None)
    cfg.abi_version = abi.MCX_ABI_VERSION
    assert L.mcx_create(C.byref(cfg), C.byref(h)) == abi.MCX_ERR_INVALID_ARG  # zero subpartitions
    assert L.mcx_step(None, 1, None) == abi.MCX_ERR_INVALID_ARG
    assert L.mcx_num_molecules(None) == 0


def test_product_never_references_the_oracle():
    """The product path (mcell_b200/) must not import, link or load anything under oracle/."""
    pkg = os.path.join(ROOT, "mcell_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp", ".inc")):
                txt = open(os.path.join(dirpath, f), errors="ignore").read()
                assert "liboracle" not in txt and "oracle_py" not in txt and "from oracle" not in txt, f
