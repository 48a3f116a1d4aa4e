"""CPU: the C-ABI library loads and exports every symbol include/dfprior_b200.h declares
(no compute calls -- there is no GPU here), and the binding's table matches the headers."""
import ctypes as C
import os
import re

import pytest

from conftest import REPO
from attentive_dfprior_b200 import _lib as L

HEADERS = [os.path.join(REPO, "include", h) for h in ("dfprior_b200.h", "dfprior_mesh.h")]


def declared_symbols():
    names = set()
    for header in HEADERS:
        txt = open(header).read()
        txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
        names |= set(re.findall(r"\b(dfp_[a-z0-9_]+)\s*\(", txt))
    return sorted(names)


@pytest.fixture(scope="module")
def lib():
    if not os.path.exists(L.LIB_PATH):
        from attentive_dfprior_b200 import build
        build.build(verbose=False)
    return L.lib()


def test_header_and_binding_agree():
    assert declared_symbols() == sorted(L.SYMBOLS)


def test_library_exports_every_declared_symbol(lib):
    for name in declared_symbols():
        assert hasattr(lib, name), name


def test_version_and_sizes(lib):
    assert lib.dfp_abi_version() == L.ABI_VERSION
    assert lib.dfp_packed_param_floats() >= L.N_PARAMS
    assert lib.dfp_render_workspace_bytes(1000, 48, 0, 2, 0) > 1000 * 48 * 8
    assert lib.dfp_render_workspace_bytes(1000, 48, 0, 2, 1) > lib.dfp_render_workspace_bytes(1000, 48, 0, 2, 0)
    assert lib.dfp_points_workspace_bytes(500000, 1, 0) > 0
    assert lib.dfp_launch_count() == 0


def test_invalid_arguments_return_error_codes_not_crashes(lib):
    """Argument validation happens before any CUDA call, so it can be exercised without a GPU."""
    rc = lib.dfp_render_fwd(None, None, None, None, None, None, 0, None)
    assert rc == -1 and b"null" in lib.dfp_last_error()
    sc, rays, cfg, out = L.DfpScene(), L.DfpRays(), L.DfpCfg(), L.DfpRenderOut()
    cfg.stage = 7
    rc = lib.dfp_render_fwd(C.byref(sc), 1, C.byref(rays), C.byref(cfg), C.byref(out), 1, 0, None)
    assert rc == -1 and b"stage" in lib.dfp_last_error()
    cfg.stage = 0
    rc = lib.dfp_render_fwd(C.byref(sc), 1, C.byref(rays), C.byref(cfg), C.byref(out), 1, 0, None)
    assert rc == -1 and b"grid" in lib.dfp_last_error()
    assert lib.dfp_pack_params(None, None, None) == -1


def test_struct_sizes_match_header_layout():
    assert C.sizeof(L.DfpGrid) == 24 and C.sizeof(L.DfpTsdf) == 48
    assert C.sizeof(L.DfpScene) == 3 * 24 + 48 + 2 * 48
    assert C.sizeof(L.DfpCfg) == 24 + 64 * 4 * 2 + 64 * 8 * 2
    assert C.sizeof(L.DfpRays) == 48 and C.sizeof(L.DfpPoints) == 32
