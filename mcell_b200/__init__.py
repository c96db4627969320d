"""mcell_b200 — B200-native replacement for MCell4's diffuse-and-react hot path.

csrc/      hand-written sm_100a kernels + the C ABI (include/mcx.h) -> libmcx.so
engine.py  ctypes binding of the C ABI (fails loudly without the CUDA extension / a GPU)
model.py   host-side table builder (units, space steps, reaction probabilities, geometry generators)
"""
from . import abi  # noqa: F401
from .model import Model, Config, MolArrays, create_box, create_icosphere, release_uniform_box  # noqa: F401
from .engine import Engine, McxError, load_library, philox_block  # noqa: F401
