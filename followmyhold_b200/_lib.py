"""ctypes binding of ``libfoho_b200.so`` (the C-ABI declared in ``include/foho_b200.h``).

There is deliberately no CPU fallback: if the shared library is missing or a symbol is
absent, importing callers get a loud ``FohoLibraryError``.
"""
from __future__ import annotations

import ctypes as C
import os
import shutil
import subprocess
from pathlib import Path
from typing import List, Optional

PKG_DIR = Path(__file__).resolve().parent
REPO_ROOT = PKG_DIR.parent
CSRC = PKG_DIR / "csrc"
LIB_PATH = PKG_DIR / "libfoho_b200.so"
SOURCES = ["guidance_stream.cu", "guidance_sparse.cu", "guidance_objmesh.cu", "guidance_chamfer.cu", "guidance_voxdist.cu", "guidance_update.cu", "icp.cu", "mesh_sample.cu",
           "mesh_sdf.cu", "mesh_decimate.cpp", "decoder_gemm.cu", "decoder_attn.cu", "decoder_attn_bwd.cu", "decoder_ops.cu", "guidance_raster.cu", "guidance_dmc.cu"]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
              "-Xcompiler", "-fPIC", "-Xcompiler", "-O3"]

FOHO_NUM_TERMS = 16
FOHO_E_WORKSPACE = -3
ABI_VERSION = 7
TERM_NAMES = ["total", "pen", "con", "int", "count", "mom", "ch", "kp", "treg_h", "treg_o", "dist", "vreg",
              "edge", "mean_d2", "ncand", "flags"]

# every symbol include/foho_b200.h declares (tests/test_capi_symbols.py checks both directions)
EXPORTED_SYMBOLS = [
    "foho_abi_version", "foho_status_string", "foho_default_weights",
    "foho_guidance_workspace_bytes", "foho_guidance_energy_fwd_bwd", "foho_guidance_update", "foho_guidance_update_f16",
    "foho_guidance_accel_bytes", "foho_guidance_prepare_statics",
    "foho_scheduler_step", "foho_scheduler_step_f16", "foho_mock_decoder_forward", "foho_mock_decoder_backward", "foho_mock_decoder_forward_f16", "foho_mock_decoder_backward_f16",
    "foho_icp_workspace_bytes", "foho_icp_run", "foho_icp_run_batch",
    "foho_remove_close_workspace_bytes", "foho_remove_close",
    "foho_mesh2sdf_workspace_bytes", "foho_mesh2sdf_lattice", "foho_intersection_count", "foho_mesh_decimate",
    "foho_tc_gemm", "foho_tc_attention", "foho_tc_attention_bwd",
    "foho_dec_layernorm", "foho_dec_layernorm_bwd", "foho_dec_softmax", "foho_dec_softmax_bwd", "foho_dec_fourier_embed",
    "foho_dec_head", "foho_dec_head_bwd", "foho_dec_gather_rows", "foho_dec_cast",
    "foho_dec_rowdot", "foho_dec_gather_f32", "foho_dec_compact_workspace_bytes", "foho_dec_compact_grad",
    "foho_raster_workspace_bytes", "foho_raster_losses_fwd_bwd",
    "foho_dmc_workspace_bytes", "foho_dmc_extract", "foho_dmc_backward",
]


class FohoLibraryError(RuntimeError):
    pass


class FohoStatusError(RuntimeError):
    def __init__(self, fn: str, status: int, msg: str):
        super().__init__(f"{fn} failed with status {status}: {msg}")
        self.status = status


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise FohoLibraryError("nvcc not found; cannot build libfoho_b200.so")


def _stale() -> bool:
    if not LIB_PATH.exists():
        return True
    t = LIB_PATH.stat().st_mtime
    deps = list(CSRC.glob("*.cu")) + list(CSRC.glob("*.cuh")) + list(CSRC.glob("*.cpp")) + [REPO_ROOT / "include" / "foho_b200.h"]
    return any(d.stat().st_mtime > t for d in deps)


def build(force: bool = False, verbose: bool = False) -> Path:
    """Compile every CUDA source for sm_100a into the in-tree shared library."""
    if not force and not _stale():
        return LIB_PATH
    nvcc = _nvcc()
    objs: List[str] = []
    build_dir = PKG_DIR / "build"
    build_dir.mkdir(exist_ok=True)
    procs = []
    for src in SOURCES:
        obj = build_dir / (Path(src).stem + ".o")
        cmd = [nvcc, *NVCC_FLAGS, *os.environ.get("FOHO_B200_EXTRA_NVCC_FLAGS", "").split(), "-c", str(CSRC / src), "-o", str(obj)]
        if verbose:
            print(" ".join(cmd))
        procs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
        objs.append(str(obj))
    for src, p in procs:
        out, _ = p.communicate()
        if p.returncode != 0:
            raise FohoLibraryError(f"nvcc failed on {src}:\n{out}")
        if verbose and out:
            print(out)
    tmp = str(LIB_PATH) + ".tmp"
    cmd = [nvcc, "-shared", "-o", tmp, *objs, "-gencode", "arch=compute_100a,code=sm_100a", "-lcudart"]
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if r.returncode != 0:
        raise FohoLibraryError(f"link failed:\n{r.stdout}")
    os.replace(tmp, LIB_PATH)
    return LIB_PATH


class Weights(C.Structure):
    _fields_ = [(n, C.c_float) for n in (
        "w_dist", "w_vreg", "w_edge", "w_treg_o", "w_hand", "w_kp", "w_treg_h", "w_int_lo", "w_int_hi",
        "dist_margin", "w_pen", "w_con", "w_ivol", "w_ch", "w_mom", "con_margin")]


class GuidanceDesc(C.Structure):
    _fields_ = [
        ("B", C.c_int32), ("D", C.c_int32), ("Vh", C.c_int32), ("Fh", C.c_int32), ("P", C.c_int32),
        ("n_joints", C.c_int32), ("image_h", C.c_int32), ("image_w", C.c_int32), ("late_step", C.c_int32),
        ("stream_variant", C.c_int32), ("stage_mask", C.c_int32), ("serial", C.c_int32), ("stream_stages", C.c_int32), ("stream_prefetch", C.c_int32), ("stream_ctas", C.c_int32), ("lane", C.c_int32), ("fov_deg", C.c_float), ("bound", C.c_float), ("w", Weights),
        ("sdf", C.c_void_p), ("grad_sdf", C.c_void_p), ("hand_rest", C.c_void_p), ("hand_faces", C.c_void_p),
        ("cloud", C.c_void_p), ("T_h2m", C.c_void_p), ("obj_center", C.c_void_p), ("theta", C.c_void_p),
        ("j_regressor", C.c_void_p), ("kps_2d", C.c_void_p), ("grad_hand_ext", C.c_void_p),
        ("grad_theta", C.c_void_p), ("terms", C.c_void_p), ("hand_moge", C.c_void_p), ("hand_grid", C.c_void_p),
        ("Vo_total", C.c_int32), ("Eo_total", C.c_int32), ("obj_verts", C.c_void_p),
        ("obj_vert_offsets", C.c_void_p), ("obj_edges", C.c_void_p), ("obj_edge_offsets", C.c_void_p),
        ("grad_obj_verts", C.c_void_p), ("workspace", C.c_void_p), ("workspace_bytes", C.c_size_t),
        ("accel", C.c_void_p), ("accel_bytes", C.c_size_t),
        ("hand_nbr_off", C.c_void_p), ("hand_nbr", C.c_void_p), ("nbr_stride", C.c_int32), ("reserved1", C.c_int32),
        ("trace", C.c_void_p), ("sticky_flags", C.c_void_p), ("obj_moge", C.c_void_p), ("grad_obj_ext", C.c_void_p),
    ]


class UpdateDesc(C.Structure):
    _fields_ = [
        ("B", C.c_int32), ("L", C.c_int32), ("step", C.c_int32),
        ("beta1", C.c_float), ("beta2", C.c_float), ("eps", C.c_float), ("weight_decay", C.c_float),
        ("lr_theta", C.c_float * 6), ("lr_velocity", C.c_float), ("sigma", C.c_float), ("theta_mask", C.c_uint32),
        ("theta", C.c_void_p), ("grad_theta", C.c_void_p), ("theta_m", C.c_void_p), ("theta_v", C.c_void_p),
        ("velocity", C.c_void_p), ("grad_velocity", C.c_void_p), ("vel_m", C.c_void_p), ("vel_v", C.c_void_p),
        ("x_t", C.c_void_p), ("x1", C.c_void_p),
        ("terms", C.c_void_p), ("nan_flag", C.c_void_p),
    ]


class GemmDesc(C.Structure):
    _fields_ = [
        ("M", C.c_int32), ("N", C.c_int32), ("K", C.c_int32), ("batch", C.c_int32),
        ("a_mn_major", C.c_int32), ("b_mn_major", C.c_int32), ("c_f32", C.c_int32), ("res_f32", C.c_int32),
        ("act", C.c_int32), ("block_n", C.c_int32), ("max_ctas", C.c_int32), ("alpha", C.c_float),
        ("A", C.c_void_p), ("lda", C.c_int64), ("bsa", C.c_int64),
        ("B", C.c_void_p), ("ldb", C.c_int64), ("bsb", C.c_int64),
        ("C", C.c_void_p), ("ldc", C.c_int64), ("bsc", C.c_int64),
        ("bias", C.c_void_p),
        ("res", C.c_void_p), ("ldr", C.c_int64), ("bsr", C.c_int64),
        ("aux_in", C.c_void_p), ("aux_out", C.c_void_p), ("ldaux", C.c_int64), ("bsaux", C.c_int64),
        ("row_vec", C.c_void_p), ("bs_rowvec", C.c_int64),
    ]


class AttnDesc(C.Structure):
    _fields_ = [
        ("n_img", C.c_int32), ("heads", C.c_int32), ("n_q", C.c_int32), ("n_k", C.c_int32),
        ("q_shared", C.c_int32), ("max_ctas", C.c_int32), ("scale", C.c_float), ("variant", C.c_int32),
        ("q", C.c_void_p), ("ldq", C.c_int64), ("hsq", C.c_int64),
        ("k", C.c_void_p), ("ldk", C.c_int64), ("hsk", C.c_int64),
        ("v", C.c_void_p), ("ldv", C.c_int64), ("hsv", C.c_int64),
        ("out", C.c_void_p), ("ldo", C.c_int64), ("out_img_stride", C.c_int64), ("lse2", C.c_void_p), ("lse2_stride", C.c_int64),
    ]


class AttnBwdDesc(C.Structure):
    _fields_ = [
        ("n_img", C.c_int32), ("heads", C.c_int32), ("n_q", C.c_int32), ("n_k", C.c_int32),
        ("max_ctas", C.c_int32), ("scale", C.c_float),
        ("q", C.c_void_p), ("ldq", C.c_int64), ("hsq", C.c_int64),
        ("k", C.c_void_p), ("ldk", C.c_int64), ("hsk", C.c_int64),
        ("v", C.c_void_p), ("ldv", C.c_int64), ("hsv", C.c_int64),
        ("d_out", C.c_void_p), ("lddo", C.c_int64), ("hsdo", C.c_int64),
        ("lse2", C.c_void_p), ("lse2_stride", C.c_int64), ("delta", C.c_void_p), ("delta_stride", C.c_int64),
        ("dq", C.c_void_p), ("lddq", C.c_int64), ("hsdq", C.c_int64),
        ("dk", C.c_void_p), ("lddk", C.c_int64), ("hsdk", C.c_int64),
        ("dv", C.c_void_p), ("lddv", C.c_int64), ("hsdv", C.c_int64),
    ]


class RasterDesc(C.Structure):
    _fields_ = [
        ("B", C.c_int32), ("V_total", C.c_int32), ("F_total", C.c_int32), ("H", C.c_int32), ("W", C.c_int32),
        ("tile_cap", C.c_int32), ("w_normal", C.c_float), ("w_disp", C.c_float), ("w_sil", C.c_float), ("accumulate_grad", C.c_int32),
        ("V1", C.c_int32), ("F1", C.c_int32), ("skip_set1", C.c_int32), ("reserved", C.c_int32),
        ("vert_offsets2", C.c_void_p), ("face_offsets2", C.c_void_p),
        ("verts", C.c_void_p), ("faces", C.c_void_p), ("vert_offsets", C.c_void_p), ("face_offsets", C.c_void_p),
        ("fov_deg", C.c_void_p), ("gt_normals", C.c_void_p), ("gt_mask", C.c_void_p), ("n_valid", C.c_void_p),
        ("gt_disp", C.c_void_p), ("gt_sil", C.c_void_p), ("losses", C.c_void_p), ("grad_verts", C.c_void_p),
        ("out_p2f", C.c_void_p), ("out_zbuf", C.c_void_p), ("out_nraw", C.c_void_p),
        ("workspace", C.c_void_p), ("workspace_bytes", C.c_size_t),
    ]


class DmcDesc(C.Structure):
    _fields_ = [
        ("B", C.c_int32), ("D", C.c_int32), ("bound", C.c_float), ("cap_verts", C.c_int32), ("cap_faces", C.c_int32),
        ("cap_edges", C.c_int32), ("index_base", C.c_int32), ("reserved", C.c_int32),
        ("sdf", C.c_void_p), ("verts", C.c_void_p), ("faces", C.c_void_p), ("edges", C.c_void_p),
        ("vert_offsets", C.c_void_p), ("face_offsets", C.c_void_p), ("edge_offsets", C.c_void_p),
        ("cube_of_vert", C.c_void_p), ("flags", C.c_void_p), ("workspace", C.c_void_p), ("workspace_bytes", C.c_size_t),
    ]


class IcpProblem(C.Structure):
    _fields_ = [
        ("source", C.c_void_p), ("target", C.c_void_p),
        ("Ns", C.c_int32), ("Nt", C.c_int32), ("n_iter", C.c_int32), ("n_outliers", C.c_int32), ("fixed_scale", C.c_int32),
        ("reserved", C.c_int32),
        ("min_scale", C.c_double), ("max_scale", C.c_double),
        ("transform_out", C.c_void_p), ("cost_out", C.c_void_p), ("cost_history", C.c_void_p),
        ("nn_index_last", C.c_void_p),
        ("workspace", C.c_void_p), ("workspace_bytes", C.c_size_t),
    ]


_lib: Optional[C.CDLL] = None


def load(build_if_missing: bool = False) -> C.CDLL:
    """Load the shared library (never falls back to anything else)."""
    global _lib
    if _lib is not None:
        return _lib
    if not LIB_PATH.exists():
        if build_if_missing:
            build()
        else:
            raise FohoLibraryError(
                f"{LIB_PATH} is missing. Build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "(needs nvcc). There is no CPU fallback for the guidance / alignment kernels.")
    try:
        lib = C.CDLL(str(LIB_PATH))
    except OSError as e:  # pragma: no cover - depends on the box
        raise FohoLibraryError(f"cannot load {LIB_PATH}: {e}") from e
    missing = [s for s in EXPORTED_SYMBOLS if not hasattr(lib, s)]
    if missing:
        raise FohoLibraryError(f"{LIB_PATH} lacks symbols {missing}; rebuild it")
    lib.foho_abi_version.restype = C.c_int
    lib.foho_status_string.restype = C.c_char_p
    lib.foho_status_string.argtypes = [C.c_int]
    lib.foho_default_weights.argtypes = [C.POINTER(Weights)]
    lib.foho_default_weights.restype = None
    lib.foho_guidance_workspace_bytes.restype = C.c_size_t
    lib.foho_guidance_workspace_bytes.argtypes = [C.c_int32] * 6
    lib.foho_guidance_energy_fwd_bwd.restype = C.c_int
    lib.foho_guidance_energy_fwd_bwd.argtypes = [C.POINTER(GuidanceDesc), C.c_void_p]
    lib.foho_guidance_accel_bytes.restype = C.c_size_t
    lib.foho_guidance_accel_bytes.argtypes = [C.c_int32] * 3
    lib.foho_guidance_prepare_statics.restype = C.c_int
    lib.foho_guidance_prepare_statics.argtypes = [C.POINTER(GuidanceDesc), C.c_void_p]
    lib.foho_guidance_update.restype = C.c_int
    lib.foho_guidance_update.argtypes = [C.POINTER(UpdateDesc), C.c_void_p]
    lib.foho_guidance_update_f16.restype = C.c_int
    lib.foho_guidance_update_f16.argtypes = [C.POINTER(UpdateDesc), C.c_void_p]
    lib.foho_scheduler_step.restype = C.c_int
    lib.foho_scheduler_step.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_float,
                                        C.c_float, C.c_void_p]
    lib.foho_scheduler_step_f16.restype = C.c_int
    lib.foho_scheduler_step_f16.argtypes = list(lib.foho_scheduler_step.argtypes)
    lib.foho_mock_decoder_forward.restype = C.c_int
    lib.foho_mock_decoder_forward.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_int64,
                                              C.c_int32, C.c_float, C.c_void_p]
    lib.foho_mock_decoder_backward.restype = C.c_int
    lib.foho_mock_decoder_backward.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_int64, C.c_int32,
                                               C.c_float, C.c_void_p]
    lib.foho_mock_decoder_forward_f16.restype = C.c_int
    lib.foho_mock_decoder_forward_f16.argtypes = lib.foho_mock_decoder_forward.argtypes
    lib.foho_mock_decoder_backward_f16.restype = C.c_int
    lib.foho_mock_decoder_backward_f16.argtypes = lib.foho_mock_decoder_backward.argtypes
    lib.foho_icp_workspace_bytes.restype = C.c_size_t
    lib.foho_icp_workspace_bytes.argtypes = [C.c_int32, C.c_int32]
    lib.foho_icp_run.restype = C.c_int
    lib.foho_icp_run.argtypes = [C.c_void_p, C.c_int32, C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_int32,
                                 C.c_double, C.c_double, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                 C.c_void_p, C.c_size_t, C.c_void_p]
    lib.foho_icp_run_batch.restype = C.c_int
    lib.foho_icp_run_batch.argtypes = [C.POINTER(IcpProblem), C.c_int32, C.c_void_p]
    lib.foho_remove_close_workspace_bytes.restype = C.c_size_t
    lib.foho_remove_close_workspace_bytes.argtypes = [C.c_int32]
    lib.foho_remove_close.restype = C.c_int
    lib.foho_remove_close.argtypes = [C.c_void_p, C.c_int32, C.c_double, C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p]
    lib.foho_mesh2sdf_workspace_bytes.restype = C.c_size_t
    lib.foho_mesh2sdf_workspace_bytes.argtypes = [C.c_int32] * 5
    lib.foho_mesh2sdf_lattice.restype = C.c_int
    lib.foho_mesh2sdf_lattice.argtypes = [C.c_void_p, C.c_int32, C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p,
                                          C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p,
                                          C.c_size_t, C.c_void_p]
    lib.foho_intersection_count.restype = C.c_int
    lib.foho_intersection_count.argtypes = [C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p]
    lib.foho_mesh_decimate.restype = C.c_int
    lib.foho_mesh_decimate.argtypes = [C.c_void_p, C.c_int32, C.c_void_p, C.c_int32, C.c_int32, C.c_double, C.c_void_p,
                                       C.POINTER(C.c_int32), C.c_void_p, C.POINTER(C.c_int32)]
    lib.foho_tc_gemm.restype = C.c_int
    lib.foho_tc_gemm.argtypes = [C.POINTER(GemmDesc), C.c_void_p]
    lib.foho_tc_attention.restype = C.c_int
    lib.foho_tc_attention.argtypes = [C.POINTER(AttnDesc), C.c_void_p]
    lib.foho_tc_attention_bwd.restype = C.c_int
    lib.foho_tc_attention_bwd.argtypes = [C.POINTER(AttnBwdDesc), C.c_void_p]
    vp, i32, i64, f32 = C.c_void_p, C.c_int32, C.c_int64, C.c_float
    lib.foho_dec_layernorm.argtypes = [vp, i32, i64, i64, vp, vp, f32, vp, i32, i64, i64, i64, i32, vp]
    lib.foho_dec_layernorm_bwd.argtypes = [vp, i32, i64, i64, vp, f32, vp, i32, i64, i64, vp, vp, i32, i64, i64, i64, i32, vp]
    lib.foho_dec_softmax.argtypes = [vp, vp, i64, i32, vp]
    lib.foho_dec_softmax_bwd.argtypes = [vp, vp, vp, i64, i32, f32, vp]
    lib.foho_dec_fourier_embed.argtypes = [vp, vp, i64, i32, i32, i32, vp]
    lib.foho_dec_head.argtypes = [vp, i64, vp, vp, f32, vp, f32, vp, vp, i64, vp]
    lib.foho_dec_head_bwd.argtypes = [vp, i64, vp, f32, vp, vp, vp, f32, vp, i64, i64, vp]
    lib.foho_dec_gather_rows.argtypes = [vp, i64, vp, vp, i64, i64, i32, vp]
    lib.foho_dec_cast.argtypes = [vp, i64, vp, i64, i64, i32, f32, i32, vp]
    lib.foho_dmc_workspace_bytes.argtypes = [i32, i32]
    lib.foho_dmc_workspace_bytes.restype = C.c_size_t
    lib.foho_dmc_extract.argtypes = [C.POINTER(DmcDesc), vp]
    lib.foho_dmc_extract.restype = C.c_int
    lib.foho_dmc_backward.argtypes = [C.POINTER(DmcDesc), vp, vp, vp]
    lib.foho_dmc_backward.restype = C.c_int
    lib.foho_raster_workspace_bytes.argtypes = [C.POINTER(RasterDesc)]
    lib.foho_raster_workspace_bytes.restype = C.c_size_t
    lib.foho_raster_losses_fwd_bwd.argtypes = [C.POINTER(RasterDesc), vp]
    lib.foho_raster_losses_fwd_bwd.restype = C.c_int
    lib.foho_dec_rowdot.argtypes = [vp, i64, vp, i64, vp, i64, i64, i32, vp]
    lib.foho_dec_rowdot.restype = C.c_int
    lib.foho_dec_gather_f32.argtypes = [vp, i64, vp, vp, i64, i32, vp]
    lib.foho_dec_gather_f32.restype = C.c_int
    lib.foho_dec_compact_workspace_bytes.argtypes = [i32, i64]
    lib.foho_dec_compact_workspace_bytes.restype = C.c_size_t
    lib.foho_dec_compact_grad.argtypes = [vp, i32, i64, i32, vp, vp, vp, vp, vp, C.c_size_t, vp]
    lib.foho_dec_compact_grad.restype = C.c_int
    for _n in ("layernorm", "layernorm_bwd", "softmax", "softmax_bwd", "fourier_embed", "head", "head_bwd", "gather_rows", "cast"):
        getattr(lib, "foho_dec_" + _n).restype = C.c_int
    if lib.foho_abi_version() != ABI_VERSION:
        raise FohoLibraryError("libfoho_b200.so ABI version mismatch; rebuild it")
    _lib = lib
    return lib


def check(fn: str, status: int) -> None:
    if status != 0:
        msg = load().foho_status_string(status)
        raise FohoStatusError(fn, status, msg.decode() if msg else "?")


def default_weights() -> Weights:
    w = Weights()
    load().foho_default_weights(C.byref(w))
    return w
