"""The C-ABI library builds for sm_100a, loads, and exports exactly what include/foho_b200.h
declares (no compute calls: this runs on the GPU-less build box)."""
import ctypes
import os
import re
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _header_functions():
    src = open(os.path.join(ROOT, "include", "foho_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    names = re.findall(r"\b(foho_[a-z0-9_]+)\s*\(", src)
    return sorted(set(names))


def test_library_exports_every_declared_symbol():
    from followmyhold_b200 import _lib
    _lib.build()
    lib = _lib.load()
    declared = _header_functions()
    assert declared, "no functions parsed from the header"
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in include/foho_b200.h but not exported"
    assert sorted(_lib.EXPORTED_SYMBOLS) == declared
    out = subprocess.run(["nm", "-D", "--defined-only", str(_lib.LIB_PATH)], capture_output=True, text=True).stdout
    exported = sorted(set(re.findall(r"\b(foho_[a-z0-9_]+)\b", out)))
    assert exported == declared, (set(exported) ^ set(declared))


def test_struct_layouts_match_header_sizes():
    """ctypes mirrors of the POD structs must have the C sizes (checked with a tiny C program)."""
    from followmyhold_b200 import _lib
    prog = r'''
#include <stdio.h>
#include "foho_b200.h"
int main(void){printf("%zu %zu %zu %zu %zu %zu %zu %zu %zu\n", sizeof(foho_weights), sizeof(foho_guidance_desc), sizeof(foho_update_desc),
  sizeof(foho_gemm_desc), sizeof(foho_attn_desc), sizeof(foho_attn_bwd_desc), sizeof(foho_raster_desc), sizeof(foho_dmc_desc),
  sizeof(foho_icp_problem));return 0;}
'''
    import tempfile
    with tempfile.TemporaryDirectory() as td:
        c = os.path.join(td, "s.c"); exe = os.path.join(td, "s")
        open(c, "w").write(prog)
        subprocess.run(["gcc", "-I", os.path.join(ROOT, "include"), c, "-o", exe], check=True)
        sizes = [int(x) for x in subprocess.run([exe], capture_output=True, text=True).stdout.split()]
    mirrors = (_lib.Weights, _lib.GuidanceDesc, _lib.UpdateDesc, _lib.GemmDesc, _lib.AttnDesc, _lib.AttnBwdDesc, _lib.RasterDesc,
               _lib.DmcDesc, _lib.IcpProblem)
    for m, n in zip(mirrors, sizes):
        assert ctypes.sizeof(m) == n, (m.__name__, ctypes.sizeof(m), n)


def test_argument_validation_without_gpu():
    """Invalid descriptors are rejected before any CUDA call."""
    from followmyhold_b200 import _lib
    lib = _lib.load()
    assert lib.foho_abi_version() == _lib.ABI_VERSION
    assert lib.foho_guidance_energy_fwd_bwd(None, None) == -1
    d = _lib.GuidanceDesc()
    assert lib.foho_guidance_energy_fwd_bwd(ctypes.byref(d), None) == -1          # NULL pointers
    assert lib.foho_guidance_workspace_bytes(0, 64, 778, 1538, 0, 0) == 0
    assert lib.foho_guidance_workspace_bytes(8, 256, 778, 1538, 65536, 0) > 0
    assert lib.foho_icp_workspace_bytes(0, 10) == 0 and lib.foho_icp_workspace_bytes(5000, 10000) > 0
    assert lib.foho_icp_run(None, 1, None, 1, 1, 0, 0, 0.5, 2.0, None, None, None, None, None, 0, None) == -1
    assert lib.foho_status_string(-3).decode().startswith("workspace")
    assert "driver" in lib.foho_status_string(-5).decode()
    assert lib.foho_icp_run_batch(None, 1, None) == -1 and lib.foho_icp_run_batch((_lib.IcpProblem * 1)(), 0, None) == 0
    assert lib.foho_remove_close_workspace_bytes(0) == 0 and lib.foho_remove_close_workspace_bytes(30000) > 0
    assert lib.foho_remove_close(None, 10, 0.1, None, None, 0, None) == -1
    assert lib.foho_tc_attention_bwd(None, None) == -1
    assert lib.foho_tc_attention_bwd(ctypes.byref(_lib.AttnBwdDesc()), None) == -1
    w = _lib.default_weights()
    assert abs(w.w_dist - 10.0) < 1e-9 and abs(w.w_int_lo - 1e-9) < 1e-15 and abs(w.dist_margin - 0.01) < 1e-9


def test_sass_uses_tma_bulk_copies():
    """The dense stream kernel is TMA-staged: UBLKCP (cp.async.bulk) must be in the SASS."""
    from followmyhold_b200 import _lib
    _lib.build()
    r = subprocess.run(["cuobjdump", "-sass", str(_lib.LIB_PATH)], capture_output=True, text=True)
    if r.returncode != 0:
        pytest.skip("cuobjdump unavailable")
    assert "UBLKCP" in r.stdout and "SYNCS" in r.stdout
    assert "sm_100a" in r.stdout or "sm_100" in r.stdout
