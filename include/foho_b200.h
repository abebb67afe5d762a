/*
 * foho_b200.h -- C-ABI of the B200-native FollowMyHold guidance / alignment hot path.
 *
 * The reference (aidilayce/FollowMyHold) has no FFI of its own: its hot path is Python
 * calling pytorch3d / kaolin / scipy / trimesh library kernels.  Each entry point below
 * therefore names the *Python seam* it replaces (file:line under the reference tree);
 * INTEGRATION.md shows the ctypes stub a maintainer would drop into the reference.
 *
 * Contract:
 *   - plain C: raw pointers + sizes, no torch / C++ types in any signature;
 *   - every pointer marked "device" is CUDA device memory owned by the caller;
 *   - functions enqueue work on the given stream and return without synchronising
 *     (except the *_host variants, which are synchronous by definition);
 *   - no device allocation: the caller passes a workspace sized by the matching
 *     *_workspace_bytes() query.  The only resources the library creates are host side: per device and
 *     lane, three CUDA streams and five events for the fork/join of foho_guidance_energy_fwd_bwd,
 *     created on its first call (make that call outside a stream capture);
 *   - return value: 0 ok, <0 invalid argument (FOHO_E_*), >0 a cudaError_t value;
 *   - thread-safe for distinct (stream, workspace) pairs; calls that share a lane's side streams
 *     are serialised by an internal mutex while they enqueue.
 *   - there is NO CPU fallback: every compute entry point needs a CUDA device.
 */
#ifndef FOHO_B200_H
#define FOHO_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define FOHO_ABI_VERSION 7

#define FOHO_OK 0
#define FOHO_E_NULL (-1)      /* required pointer is NULL            */
#define FOHO_E_SHAPE (-2)     /* size out of the supported range     */
#define FOHO_E_WORKSPACE (-3) /* workspace too small / misaligned    */
#define FOHO_E_ARG (-4)       /* other invalid scalar argument       */
#define FOHO_E_DRIVER (-5)    /* cuTensorMapEncodeTiled unavailable / rejected the tensor map */

/* indices into the per-sample `terms` output of foho_guidance_energy_fwd_bwd */
enum {
  FOHO_T_TOTAL = 0,
  FOHO_T_PEN = 1,     /* NS a13  mean relu(-s_i)                                   */
  FOHO_T_CON = 2,     /* NS a13  mean clamp(|s_i|-m,0)                             */
  FOHO_T_INT = 3,     /* NS a14  mean_vox relu(-SDF_o) relu(-SDF_h)                */
  FOHO_T_COUNT = 4,   /* REF a9  count(sdf_o<0 & sdf_h<0)/1000 (pipelines.py:231)  */
  FOHO_T_MOM = 5,     /* NS a10v mean_vox relu(-SDF_o)|y|^2                        */
  FOHO_T_CH = 6,      /* NS a15  symmetric chamfer hand <-> cloud                  */
  FOHO_T_KP = 7,      /* REF a11 2-D key-point MSE (pipelines.py:1490-1495)        */
  FOHO_T_TREG_H = 8,  /* REF a10 mean(t_h^2) (pipelines.py:1498)                   */
  FOHO_T_TREG_O = 9,  /* REF a10 mean(t_o^2) (pipelines.py:1571)                   */
  FOHO_T_DIST = 10,   /* REF a7  mean clamp(d2-0.01,0) (pipelines.py:1529-1541)    */
  FOHO_T_VREG = 11,   /* REF a10 mean(obj verts^2) (pipelines.py:1570)             */
  FOHO_T_EDGE = 12,   /* REF a10 mesh_edge_loss (pipelines.py:1575)                */
  FOHO_T_MEAN_D2 = 13,/* REF     mean hand->object d2 (weight switch, :1561)       */
  FOHO_T_NCAND = 14,  /* diagnostics: voxels inside hand & object                  */
  FOHO_T_FLAGS = 15,  /* diagnostics: bit0 = candidate list overflow, bit1 = a cloud point farther
                         than 63 units from its nearest hand vertex (gradient clamped)       */
  FOHO_NUM_TERMS = 16
};

/* Loss weights; the REF defaults are the literals of
 * third_party_patches/hy3dgen/shapegen/pipelines.py:1499-1504,1561-1564,1578-1588. */
typedef struct foho_weights {
  float w_dist, w_vreg, w_edge, w_treg_o, w_hand, w_kp, w_treg_h;
  float w_int_lo, w_int_hi, dist_margin;
  float w_pen, w_con, w_ivol, w_ch, w_mom, con_margin;
} foho_weights;

/* One batched guidance evaluation.  All arrays are contiguous, float32 unless noted.
 * theta layout per sample: [s_h, t_h(3), q_h(4, wxyz), s_o, t_o(3), q_o(4)] = 16 floats
 * (leaves of third_party/utilz/code_utils.py:57-78). */
typedef struct foho_guidance_desc {
  int32_t B;            /* samples in the batch (>=1)                                   */
  int32_t D;            /* lattice points per axis (reference 65; 2..1024)              */
  int32_t Vh;           /* hand vertices (MANO: 778), <= 4096                           */
  int32_t Fh;           /* hand faces (MANO: 1538), <= 8192                             */
  int32_t P;            /* cloud points per sample (0 disables the chamfer term)        */
  int32_t n_joints;     /* rows of j_regressor (MANO: 16), 0 disables key-points        */
  int32_t image_h, image_w;
  int32_t late_step;    /* 1 when i >= num_inference_steps-3 (pipelines.py:1561)        */
  int32_t stream_variant; /* 0 = default kernel choice, 1 = LDG/STG, 2 = TMA bulk       */
  int32_t stage_mask;   /* profiling hook: 0 = whole evaluation; else bit0 prep, bit1 dense
                           stream, bit2 chamfer, bit3 hand voxels, bit4 finalize, bit5 object mesh */
  int32_t serial;       /* 0 = the sparse kernels run beside the dense stream on library-owned side
                           streams (fork/join by events; legal under stream capture); 1 = every kernel
                           in series on the caller's stream                                    */
  int32_t stream_stages;   /* TMA stream: ring depth of 16 KB tiles (0 = default 6) */
  int32_t stream_prefetch; /* TMA stream: bulk loads kept in flight per CTA (0 = default 4)      */
  int32_t stream_ctas;     /* TMA stream: persistent CTAs per SM (0 = default 1)                */
  int32_t lane;            /* 0..3: which set of library side streams this call forks onto; callers
                              that keep several evaluations in flight (micro-batches on different
                              streams) give each its own lane so their side chains do not order
                              behind one another                                                */
  float fov_deg;        /* MoGe fov_x in degrees (guidance/run.py:228-230)              */
  float bound;          /* lattice half extent, 1.10 (pipelines.py:1127)                */
  foho_weights w;

  const float *sdf;          /* device [B,D,D,D]  negative inside, z fastest            */
  float *grad_sdf;           /* device [B,D,D,D]  OUT dE/dSDF (fully overwritten)       */
  const float *hand_rest;    /* device [B,Vh,3]   aligned MANO verts, MoGe space        */
  const int32_t *hand_faces; /* device [Fh,3]     shared topology                       */
  const float *cloud;        /* device [B,P,3]    MoGe partial cloud (may be NULL if P=0)*/
  const float *T_h2m;        /* device [B,16]     row-major 4x4 Hunyuan->MoGe similarity */
  const float *obj_center;   /* device [B,3]      centre of the object similarity       */
  const float *theta;        /* device [B,16]                                            */
  const float *j_regressor;  /* device [n_joints,Vh] or NULL                             */
  const float *kps_2d;       /* device [B,n_joints+5,2] or NULL                          */
  const float *grad_hand_ext;/* device [B,Vh,3] or NULL: dE_ext/d(transformed hand verts)
                                added before the chain rule (renderer losses live upstream) */
  float *grad_theta;         /* device [B,16]  OUT                                       */
  float *terms;              /* device [B,FOHO_NUM_TERMS] OUT                            */
  float *hand_moge;          /* device [B,Vh,3] OUT transformed hand verts (may be NULL) */
  float *hand_grid;          /* device [B,Vh,3] OUT same verts in lattice units (may be NULL) */

  /* optional explicit object mesh (FlexiCubes output, Hunyuan space; pipelines.py:1509): enables the
   * REF terms a7 distance_loss (:1529-1541), a10 obj_verts_loss / mesh_edge_loss (:1570,1575) and the
   * w_intersection switch (:1561-1564).  The similarity theta_o acts about the bbox centre of the
   * T_h2m-transformed vertices (:108-118); sample b owns packed vertices
   * [obj_vert_offsets[b], obj_vert_offsets[b+1]) and edges [obj_edge_offsets[b], obj_edge_offsets[b+1]). */
  int32_t Vo_total;          /* total packed object vertices over the batch (0 = none)   */
  int32_t Eo_total;          /* total packed unique edges                                */
  const float *obj_verts;    /* device [Vo_total,3]                                      */
  const int32_t *obj_vert_offsets; /* device [B+1]                                       */
  const int32_t *obj_edges;  /* device [Eo_total,2] indices into the packed vertex array */
  const int32_t *obj_edge_offsets; /* device [B+1]                                       */
  float *grad_obj_verts;     /* device [Vo_total,3] OUT dE/d(obj_verts) (may be NULL)    */

  void *workspace;           /* device, >= foho_guidance_workspace_bytes(...)            */
  size_t workspace_bytes;

  /* optional per-image search structures for the chamfer term, filled once per image set by
   * foho_guidance_prepare_statics (hand_rest and cloud must not change afterwards).  NULL selects the
   * brute-force search; results are identical up to ties between equidistant neighbours. */
  void *accel;               /* device, 256-byte aligned, >= foho_guidance_accel_bytes(...) */
  size_t accel_bytes;
  /* optional (with accel): Delaunay neighbour graph of the REST hand vertices of every sample, CSR:
   * sample b's vertex i has neighbours hand_nbr[b*nbr_stride + off[i] .. off[i+1]) with
   * off = hand_nbr_off + b*(Vh+1).  Enables the greedy-walk cloud->hand search (exact on a Delaunay
   * graph or any super-graph of it); NULL selects the box search. */
  const int32_t *hand_nbr_off;   /* device [B,Vh+1] */
  const uint16_t *hand_nbr;      /* device [B,nbr_stride], 16-byte aligned; nbr_stride a multiple of 8 */
  int32_t nbr_stride;
  int32_t reserved1;

  /* optional timeline trace (profiling hook, NULL = off): device [FOHO_TRACE_KERNELS][2] uint64, the
   * caller initialises every pair to {UINT64_MAX, 0}; each kernel folds in the %globaltimer (ns) of
   * its first CTA start and last CTA end.  Order: prep, stream, chamfer_h2c, chamfer_c2h,
   * chamfer_brute, raster, compact, voxdist, finalize_verts, assemble, keypoints. */
  void *trace;
  /* optional: device [B] int32, OR-accumulated with every evaluation's FOHO_T_FLAGS so that an overflow in ANY
   * evaluation of a captured graph is still visible afterwards (the host layer turns bit0 into a hard error) */
  int32_t *sticky_flags;
  /* explicit object mesh, renderer hooks (row f2): the transformed object vertices leave in absolute MoGe coordinates so
   * that the rasteriser can draw them, and an external gradient with respect to those vertices (the joined hand + object
   * image terms, pipelines.py:1544-1569) joins the chain rule to theta_o and to the Hunyuan-space vertices */
  float *obj_moge;             /* device [Vo_total,3] OUT, or NULL                                      */
  const float *grad_obj_ext;   /* device [Vo_total,3] dE_ext/d(transformed object verts), or NULL       */
} foho_guidance_desc;
#define FOHO_TRACE_KERNELS 11

int foho_abi_version(void);
const char *foho_status_string(int status);
void foho_default_weights(foho_weights *w);

/* Replaces the body of the phase-2 inner iteration between `scheduler.step_final(...)`
 * and `total_loss.backward()` of
 * third_party_patches/hy3dgen/shapegen/pipelines.py:1480-1600 (volume formulation of
 * BASELINE.json north_star; term-by-term map in DESIGN.md). */
size_t foho_guidance_workspace_bytes(int32_t B, int32_t D, int32_t Vh, int32_t Fh, int32_t P,
                                     int32_t Vo_total);
int foho_guidance_energy_fwd_bwd(const foho_guidance_desc *desc, void *cuda_stream);

/* Build the search structures of desc->accel from desc->hand_rest [B,Vh,3] (Vh <= 1024),
 * desc->hand_faces [Fh,3] and desc->cloud [B,P,3]: the per-image setup the reference does once before
 * its loop (mesh / target loading, pipelines.py:1218-1256); only B, Vh, Fh, P, hand_rest, hand_faces,
 * cloud, accel, accel_bytes are read.  Contents: the cloud in Morton order with a two-level box
 * hierarchy, the rest hand as Morton-ordered leaves of 8 vertices, the faces in Morton order of their
 * rest centroids (Fh <= 2048; required then), and warm-start hints that later evaluations update
 * (they only speed the searches up; results never depend on them). */
size_t foho_guidance_accel_bytes(int32_t B, int32_t Vh, int32_t P);
int foho_guidance_prepare_statics(const foho_guidance_desc *desc, void *cuda_stream);

/* Replaces `joint_optimizer.step()` (torch.optim.AdamW(eps=1e-4), pipelines.py:1478,1601;
 * Adam of :1318 with weight_decay=0) fused with `scheduler.step_final`
 * (schedulers.py:411-493): one launch updates the 16 scalar leaves of every sample and
 * the velocity tensor, and emits x1 = x_t + (1-sigma) v_new for the next decode.
 * Arithmetic: the op sequence of torch's multi-tensor CUDA path with its rounding (mul, lerp, mul, addcmul,
 * sqrt, div, add, addcdiv; csrc/foho_adamw.cuh) -- bit-equal to torch.optim.Adam/AdamW stepping CUDA
 * parameters.  The hyper-parameters travel as float; derived scalars (1-beta, 1-lr*wd, lr/(1-beta1^t)) are
 * formed in double from the shortest decimal that round-trips each float, i.e. from the value the caller
 * wrote (0.9, 1e-4, ...), as torch forms them from python floats. */
typedef struct foho_update_desc {
  int32_t B;
  int32_t L;                 /* velocity elements per sample (3072*64)                   */
  int32_t step;              /* 1-based optimiser step (state is reset every outer step) */
  float beta1, beta2, eps, weight_decay;
  float lr_theta[6];         /* scale_h, trans_h, rot_h, scale_o, trans_o, rot_o
                                (src/foho/configs/guid_config.py:20-25)                  */
  float lr_velocity;         /* noise_obj_lr2 = 1e-2 (guid_config.py:26)                 */
  float sigma;               /* sigma_k of the current outer step                        */
  uint32_t theta_mask;       /* bit g set = leaf group g is optimised in this phase      */
  float *theta;              /* device [B,16] in/out                                     */
  const float *grad_theta;   /* device [B,16]                                            */
  float *theta_m, *theta_v;  /* device [B,16] Adam moments in/out                        */
  float *velocity;           /* device [B,L] in/out (may be NULL: scalars only)          */
  const float *grad_velocity;/* device [B,L]                                             */
  float *vel_m, *vel_v;      /* device [B,L]                                             */
  const float *x_t;          /* device [B,L] current latents (may be NULL)               */
  float *x1;                 /* device [B,L] OUT x_t + (1-sigma) v_new (may be NULL)     */
  /* NaN guard of the inner loop, evaluated on the device (no host round trip).  The reference tests
   * `torch.isnan(total_loss)` BEFORE `backward()` / `optimizer.step()` and leaves the inner loop
   * (`break`, pipelines.py:1590-1592; `return None` in the object-only phase, :1442-1444).  With both
   * pointers set, sample b is left untouched (leaves, moments, velocity, x1) by this and every later
   * update of the outer step once `terms[b*FOHO_NUM_TERMS + FOHO_T_TOTAL]` is NaN; `nan_flag[b]`
   * receives the 1-based optimiser step at which that first happened (0 = never; the caller zeroes
   * it when a new outer step starts).  Both NULL: no guard.                                         */
  const float *terms;        /* device [B,FOHO_NUM_TERMS] of the evaluation just done     */
  int32_t *nan_flag;         /* device [B] in/out                                         */
} foho_update_desc;
int foho_guidance_update(const foho_update_desc *desc, void *cuda_stream);

/* Same update with the velocity / latent arrays in fp16 -- the reference's dtype (the leaf is a clone of
 * the DiT's half output, code_utils.py:43-78): `velocity`, `grad_velocity`, `vel_m`, `vel_v`, `x_t`, `x1`
 * of the descriptor then point to __half [B,L]; the 16 scalar leaves and their moments stay float32.
 * Reproduces torch.optim.AdamW on a half parameter op by op (moments in half, a rounding after every
 * elementwise op) and the half `step_final` (schedulers.py:470-484). */
int foho_guidance_update_f16(const foho_update_desc *desc, void *cuda_stream);

/* Replaces `scheduler.step(noise_pred_obj, t, obj_latents).prev_sample`
 * (schedulers.py:235-319, call site pipelines.py:1612): prev = x + (sigma_next-sigma) v. */
int foho_scheduler_step(const float *x_t, const float *velocity, float *prev_sample, float *pred_x1,
                        int64_t n, float sigma, float sigma_next, void *cuda_stream);

/* Same step for fp16 latents / model output (the reference's dtype: latents are created fp16 at
 * pipelines.py:1204-1205).  Reproduces torch's mixed-precision evaluation of schedulers.py:294-309 bit
 * for bit: factor and product rounded to half, sum in fp32, result cast back to half.
 * x_t, velocity, prev_sample, pred_x1: device __half [n]. */
int foho_scheduler_step_f16(const void *x_t, const void *velocity, void *prev_sample, void *pred_x1, int64_t n,
                            float sigma, float sigma_next, void *cuda_stream);

/* Stand-in for `latent2sdf` (pipelines.py:292-338, the VAE decoder -- SURVEY.md 8f rank 1,
 * not built yet) so the loop can be driven end to end with mock latents: a fixed sparse
 * linear decoder  SDF[b, tap[j]] = SDF0[b, tap[j]] + alpha * x1[b, j]  and its exact adjoint
 * g_x1[b, j] = alpha * G[b, tap[j]]  (then dE/dv = (1-sigma) g_x1).  tap: device int64 [L]
 * distinct voxel indices shared by the batch. */
int foho_mock_decoder_forward(float *sdf, const float *sdf0, const float *x1, const int64_t *tap, int32_t B,
                              int64_t vol, int32_t L, float alpha, void *cuda_stream);
int foho_mock_decoder_backward(const float *grad_sdf, const int64_t *tap, float *grad_velocity, int32_t B,
                               int64_t vol, int32_t L, float alpha_times_one_minus_sigma, void *cuda_stream);
/* the same with half latents / latent gradients (the reference's dtype, pipelines.py:1204); the volume stays float */
int foho_mock_decoder_forward_f16(float *sdf, const float *sdf0, const void *x1, const int64_t *tap, int32_t B,
                                  int64_t vol, int32_t L, float alpha, void *cuda_stream);
int foho_mock_decoder_backward_f16(const float *grad_sdf, const int64_t *tap, void *grad_velocity, int32_t B,
                                   int64_t vol, int32_t L, float alpha_times_one_minus_sigma, void *cuda_stream);

/* Replaces `icp(...)` of src/foho/alignment/mesh_align.py:56-175 for the configuration
 * both callers use (on_surface=False, no rotation/reflection search): trimmed
 * similarity ICP, float64, Euclidean 1-NN (scipy cKDTree.query semantics), trim of
 * int(outliers*count_source) worst pairs, trimesh.registration.procrustes
 * (reflection=False), scale renormalise + clip, best-by-pre-update-cost.
 * source [Ns,3], target [Nt,3] device float64.  OUT transform [16] row-major device
 * float64, OUT cost [1] device float64, OUT (optional) cost_history [n_iter]. */
size_t foho_icp_workspace_bytes(int32_t Ns, int32_t Nt);
int foho_icp_run(const double *source, int32_t Ns, const double *target, int32_t Nt, int32_t n_iter,
                 int32_t n_outliers, int32_t fixed_scale, double min_scale, double max_scale,
                 double *transform_out, double *cost_out, double *cost_history, int32_t *nn_index_last,
                 void *workspace, size_t workspace_bytes, void *cuda_stream);

/* Several independent loops in ONE persistent launch (the images of a batch; `align_meshes_impl` is called once per
 * image by h2m.py:35-54 / mano.py:24-43): the SMs are divided between the problems, each problem's CTAs synchronise
 * among themselves only.  Every result is identical to foho_icp_run's on a grid of the same size; the fields are
 * foho_icp_run's arguments.  `problems` is a HOST array. */
typedef struct foho_icp_problem {
  const double *source; const double *target;
  int32_t Ns, Nt, n_iter, n_outliers, fixed_scale, reserved;
  double min_scale, max_scale;
  double *transform_out, *cost_out, *cost_history;
  int32_t *nn_index_last;
  void *workspace; size_t workspace_bytes;
} foho_icp_problem;
int foho_icp_run_batch(const foho_icp_problem *problems, int32_t n_problems, void *cuda_stream);

/* Replaces trimesh.points.remove_close inside trimesh.sample.sample_surface_even, which `icp` calls for both point
 * sets (src/foho/alignment/mesh_align.py:79,85): keep_mask[i] = 0 for every point that some pair (a < b) with
 * |p_a - p_b| <= radius drops -- the member that appears in more pairs, a when the counts are equal (cKDTree.query_pairs
 * + bincount + argmax).  points [N,3] device float64, OUT keep_mask [N] device uint8. */
size_t foho_remove_close_workspace_bytes(int32_t N);
int foho_remove_close(const double *points, int32_t N, double radius, uint8_t *keep_mask, void *workspace,
                      size_t workspace_bytes, void *cuda_stream);

/* Exact mesh -> signed distance on a rectilinear lattice: replaces `mesh2sdf`
 * (third_party/utilz/kaolin_sdf_ops.py:88-109: kaolin point_to_mesh_distance +
 * check_sign) for the grid of get_sdf_of_meshes (:131-160), whose points are the
 * product xs x ys x zs of three np.linspace arrays ("ij" order, z fastest).
 * verts device [V,3] fp32, faces device [F,3] int32, xs/ys/zs device fp32 ascending,
 * OUT sdf device [nx,ny,nz] fp32 = sqrt(d2) * (inside ? -1 : +1). */
size_t foho_mesh2sdf_workspace_bytes(int32_t V, int32_t F, int32_t nx, int32_t ny, int32_t nz);
int foho_mesh2sdf_lattice(const float *verts, int32_t V, const int32_t *faces, int32_t F, const float *xs,
                          const float *ys, const float *zs, int32_t nx, int32_t ny, int32_t nz, float *sdf_out,
                          void *workspace, size_t workspace_bytes, void *cuda_stream);

/* REF a9 on two lattice SDFs: count(sdf_obj<0 & sdf_hand<0) (pipelines.py:231-239).
 * OUT count device int64[1]. */
int foho_intersection_count(const float *sdf_hand, const float *sdf_obj, int64_t n, long long *count_out,
                            void *cuda_stream);

/* HOST function (CPU memory, no stream): quadric edge-collapse decimation of a triangle mesh down to
 * `target_faces`.  Replaces `FaceReducer()(mesh)` of the guidance stage (src/foho/guidance/run.py:161 ->
 * hy3dgen reduce_face -> MeshLab meshing_decimation_quadric_edge_collapse with preserveboundary,
 * boundaryweight=3, preservenormal, preservetopology); runs once per image after the loop.
 * IN  verts double [V,3], faces int32 [F,3];  OUT out_verts double [<=V,3], out_faces int32 [<=F,3] (caller
 * allocates V and F rows), *out_V / *out_F = rows written.  A mesh with F <= target_faces comes back
 * unchanged apart from dropped index-degenerate faces and unreferenced vertices. */
int foho_mesh_decimate(const double *verts, int32_t V, const int32_t *faces, int32_t F, int32_t target_faces,
                       double boundary_weight, double *out_verts, int32_t *out_V, int32_t *out_faces,
                       int32_t *out_F);

/* ---------------------------------------------------------------------------------------------
 * Row f1: latent -> SDF decode (third_party_patches/hy3dgen/shapegen/pipelines.py:292-312 `latent2sdf`,
 * call sites :1392,1508,1641).  The dense contractions run on the 5th-generation tensor cores
 * (tcgen05.mma, TMEM accumulators, TMA operand tiles); everything is fp16 operands / fp32 accumulation,
 * the reference's own dtype (`vae` is fp16, pipelines.py:302-306).
 * --------------------------------------------------------------------------------------------- */

/* One (batched) tensor-core GEMM with a fused epilogue -- the building block every nn.Linear of the
 * decoder and every attention product of its adjoint goes through:
 *     C[b][m][n] = res[b][m][n] + act( alpha * sum_k A[b](m,k) B[b](n,k) + bias[n] )
 * A is [M,K] and B is [N,K] (the nn.Linear weight layout) when K-major; an MN-major operand is stored
 * [K,M] / [K,N] instead (row = k).  Leading dimensions and batch strides in ELEMENTS; operands fp16,
 * 16-byte aligned, lda/ldb/K multiples of 8.  act: 0 none, 1 GELU (erf), 2 multiply by GELU'(aux_in[m][n])
 * (the backward of act 1), 3 exp2(x - row_vec[m]) (softmax probabilities from saved log2-sum-exps: the attention
 * adjoint never materialises float32 scores), 4 aux_in[m][n] * (x - alpha * row_vec[m]) (softmax backward, row_vec =
 * rowsum(dO o O)).  aux_out (optional, fp16) receives the pre-activation value. */
typedef struct foho_gemm_desc {
  int32_t M, N, K, batch;
  int32_t a_mn_major, b_mn_major;
  int32_t c_f32;             /* C is float32 (else fp16)                                    */
  int32_t res_f32;           /* res is float32 (else fp16)                                  */
  int32_t act;
  int32_t block_n;           /* 0 = choose; 64 / 128 / 256 columns per CTA tile             */
  int32_t max_ctas;          /* 0 = one persistent CTA per SM                               */
  float alpha;
  const void *A; int64_t lda, bsa;
  const void *B; int64_t ldb, bsb;
  void *C; int64_t ldc, bsc;
  const float *bias;         /* device [N] float32 or NULL                                   */
  const void *res; int64_t ldr, bsr;     /* NULL = none; may alias C (in-place accumulate)  */
  const void *aux_in; void *aux_out; int64_t ldaux, bsaux;
  const float *row_vec; int64_t bs_rowvec;   /* act 3 / 4: one float per output row (batch stride in elements) */
} foho_gemm_desc;
int foho_tc_gemm(const foho_gemm_desc *desc, void *cuda_stream);

/* Attention forward on the tensor cores (head dimension 64, no mask):
 *     O[i][q][h*64 + :] = softmax_k( scale * Q[q][h] . K[i][k][h] ) V[i][k][h]
 * the cross attention of the lattice queries onto the latent tokens (hy3dgen geo_decoder, called at
 * pipelines.py:304) and the self attention of the ShapeVAE transformer (pipelines.py:299).  Q, K, V are fp16
 * row-major with leading dimensions ld* (elements) and head h at column offset h*hs* -- so the fused
 * [tokens, heads, (q|k|v)] projections are read in place.  q holds n_q rows when q_shared (one lattice for all
 * images), else n_img*n_q rows; k, v hold n_img*n_k rows; n_k must be a multiple of 128. */
typedef struct foho_attn_desc {
  int32_t n_img, heads, n_q, n_k;
  int32_t q_shared;
  int32_t max_ctas;          /* 0 = one persistent CTA per SM */
  float scale;               /* 1/sqrt(64) = 0.125 */
  int32_t variant;           /* 0 = two query tiles per CTA in ping-pong, P kept in TMEM (default); 1 = one tile per CTA;
                                2 = two tiles, P through shared memory (A/B measurements) */
  const void *q; int64_t ldq, hsq;
  const void *k; int64_t ldk, hsk;
  const void *v; int64_t ldv, hsv;
  void *out; int64_t ldo, out_img_stride;   /* fp16 [n_img][n_q][heads*64], 16-byte aligned */
  float *lse2;               /* optional OUT float32 [n_img][heads][lse2_stride]: log2-sum-exp of the scaled scores of query
                                row q at [..][q] (variants 0 and 2) */
  int64_t lse2_stride;       /* 0 = n_q; larger when a chunk of queries writes into the rows of a longer table */
} foho_attn_desc;
int foho_tc_attention(const foho_attn_desc *desc, void *cuda_stream);

/* The adjoint of the attention above, fused (no score matrix in memory): what autograd runs for `loss.backward()`
 * through the transformer's self attention and the geo_decoder's cross attention (pipelines.py:1590-1600 -> :299,304).
 *     P = exp2(scale log2(e) Q K^T - lse2)    dS = P o (dO V^T - delta)
 *     dV = P^T dO    dK = scale dS^T Q    dQ = scale dS K
 * q, d_out hold n_img*n_q rows, k, v hold n_img*n_k rows (fp16 row-major, leading dimension ld*, head h at column
 * h*hs*); lse2 as written by foho_tc_attention, delta[i][h][q] = d_out[q][h] . out[q][h] (foho_dec_rowdot); dq, dk,
 * dv are fp16 views of the same form (16-byte aligned); dq may be NULL (the lattice queries of the cross attention do
 * not depend on the latents: their work items are skipped).  n_k must be a multiple of 128.  Deterministic (no atomics). */
typedef struct foho_attn_bwd_desc {
  int32_t n_img, heads, n_q, n_k;
  int32_t max_ctas;          /* 0 = one persistent CTA per SM */
  float scale;
  const void *q; int64_t ldq, hsq;
  const void *k; int64_t ldk, hsk;
  const void *v; int64_t ldv, hsv;
  const void *d_out; int64_t lddo, hsdo;
  const float *lse2; int64_t lse2_stride;     /* [n_img][heads][lse2_stride], 0 = n_q */
  const float *delta; int64_t delta_stride;   /* [n_img][heads][delta_stride], 0 = n_q */
  void *dq; int64_t lddq, hsdq;
  void *dk; int64_t lddk, hsdk;
  void *dv; int64_t lddv, hsdv;
} foho_attn_bwd_desc;
int foho_tc_attention_bwd(const foho_attn_bwd_desc *desc, void *cuda_stream);

/* Row-wise kernels of the decoder (HBM-bound; fp16 activations, fp32 arithmetic; every pointer device memory).
 * A "row view" addresses row r of a [outer, inner, width] tensor at base + (r / inner)*ldo + (r % inner)*ldi
 * (elements), so the per-head q_norm / k_norm (width 64) read the fused projections in place. */
int foho_dec_layernorm(const void *x, int32_t inner_x, int64_t ldo_x, int64_t ldi_x, const float *w, const float *b, float eps,
                       void *y, int32_t inner_y, int64_t ldo_y, int64_t ldi_y, int64_t rows, int32_t width /* 64 | 1024 */,
                       void *cuda_stream);
/* dx = dLN/dx^T dy (+ add): weights are frozen, only the input gradient exists */
int foho_dec_layernorm_bwd(const void *x, int32_t inner_x, int64_t ldo_x, int64_t ldi_x, const float *w, float eps, const void *dy,
                           int32_t inner_dy, int64_t ldo_dy, int64_t ldi_dy, const void *add, void *dx, int32_t inner_dx,
                           int64_t ldo_dx, int64_t ldi_dx, int64_t rows, int32_t width, void *cuda_stream);
/* materialised attention of the adjoint: P = softmax(S) (S float32, P fp16); dS = scale * P o (dP - rowsum(P o dP)) */
int foho_dec_softmax(const float *S, void *P, int64_t rows, int32_t T, void *cuda_stream);
int foho_dec_softmax_bwd(const void *P, const float *dP, void *dS, int64_t rows, int32_t T, float scale, void *cuda_stream);
/* hy3dgen FourierEmbedder: [x, sin(x 2^k), cos(x 2^k)] of fp16-rounded coordinates -> fp16 [n, ld], padding zeroed */
int foho_dec_fourier_embed(const float *xyz, void *out, int64_t n, int32_t ld, int32_t num_freqs, int32_t include_pi, void *cuda_stream);
/* sdf[idx ? idx[r] : r] = -(w_out . ln_post(x[r]) + b_out): the last two modules of geo_decoder, the float cast and the
 * sign flip of pipelines.py:309-312 */
int foho_dec_head(const void *x, int64_t ldx, const float *ln_w, const float *ln_b, float eps, const float *w_out, float b_out,
                  const int32_t *idx, float *out, int64_t rows, void *cuda_stream);
/* its backward for the rows that carry a gradient: dx[r] = d(-logit)/dx * g_scale * dS[idx ? idx[r] : r] */
int foho_dec_head_bwd(const void *x, int64_t ldx, const float *ln_w, float eps, const float *w_out, const int32_t *idx, const float *dS,
                      float g_scale, void *dx, int64_t lddx, int64_t rows, void *cuda_stream);
int foho_dec_gather_rows(const void *in, int64_t ld_in, const int32_t *idx, void *out, int64_t ld_out, int64_t rows, int32_t width,
                         void *cuda_stream);
/* row-major views [rows, cols], cols contiguous, leading dimensions in elements.
 * mode 0: f32 -> f16 (x scale); 1: f16 -> f32 (x scale); 2: f16 -> f32 accumulate; 3: out += in (fp16, contiguous);
 * 4: f16 -> f16 (x scale) */
int foho_dec_cast(const void *in, int64_t ld_in, void *out, int64_t ld_out, int64_t rows, int32_t cols, float scale, int32_t mode,
                  void *cuda_stream);

/* delta[h][r] = sum_d a[r][h][d] * b[r][h][d] (64 channels per head; a, b fp16 [rows, heads*64] with leading dimensions lda,
 * ldb): the rowsum(dO o O) of the attention backward.  out float32, head h at out + h*out_head_stride (0 = rows: a dense
 * [heads, rows] table; larger when a chunk of rows writes into the columns of a longer table). */
int foho_dec_rowdot(const void *a, int64_t lda, const void *b, int64_t ldb, float *out, int64_t out_head_stride, int64_t rows,
                    int32_t heads, void *cuda_stream);
/* out[h][i] = src[h * src_head_stride + idx[i]] (float32): per-head gather of saved log-sum-exps for the active rows */
int foho_dec_gather_f32(const float *src, int64_t src_head_stride, const int32_t *idx, float *out, int64_t n, int32_t heads,
                        void *cuda_stream);

/* Sparse view of a dense dE/dSDF [B, V]: the non-zero entries of image b, in a
 * deterministic order, go to idx / val [b, 0..count[b]) (capacity cap per image, the rest zeroed: index 0, gradient 0);
 * count [B] receives the true number, bit 0 of *flags (optional) is set when it exceeds cap.  This is what the
 * decoder's adjoint consumes: the energy touches a few thousand voxels, not the lattice. */
size_t foho_dec_compact_workspace_bytes(int32_t B, int64_t V);
int foho_dec_compact_grad(const float *g, int32_t B, int64_t V, int32_t cap, int32_t *idx, float *val, int32_t *count,
                          int32_t *flags, void *workspace, size_t workspace_bytes, void *cuda_stream);

/* ---------------------------------------------------------------------------------------------
 * Row f2, first part: the renderer of the guidance loop and its image-space losses, forward and backward to the mesh
 * vertices.  Replaces, per inner iteration, `render_normal_and_disparity(renderer, mesh)`
 * (third_party_patches/hy3dgen/shapegen/pipelines.py:272-289: pytorch3d naive rasteriser with faces_per_pixel = 1 +
 * the PhongNormalShader of :74-92 + min/max normalisation), `normal_alignment_loss` (:178-187), the disparity L1 and the
 * silhouette BCE (:1567-1569) with their weights (:1580-1583), and the autograd pass back to the vertices; camera and
 * raster settings of src/foho/guidance/run.py:84-116 (R = diag(-1,1,-1), T = 0, fov_x per image).
 * Meshes of the B images are packed: image b owns vertices [vert_offsets[b], vert_offsets[b+1]) and faces
 * [face_offsets[b], face_offsets[b+1]); face indices address the packed vertex array.  The silhouette is hard
 * (sigma = 1e-8): its BCE is reported, no gradient flows through it.  pytorch3d semantics are restated from memory
 * (oracle/raster_oracle.py): PARITY UNPINNED.
 * --------------------------------------------------------------------------------------------- */
typedef struct foho_raster_desc {
  int32_t B, V_total, F_total, H, W;
  int32_t tile_cap;            /* faces per 16x16-pixel tile the workspace holds (0 = 1024); overflow -> losses[b][7] = 1 */
  float w_normal, w_disp, w_sil;   /* 10, 10, 10 (pipelines.py:1580-1583) */
  int32_t accumulate_grad;     /* grad_verts += instead of = (several renders of one vertex buffer)              */
  /* optional second set per image (the extracted object mesh, whose counts live on the device): vertices [V1, V_total)
   * and faces [F1, F_total) of the same arrays, image b owning [V1 + vert_offsets2[b], V1 + vert_offsets2[b+1]) and
   * [F1 + face_offsets2[b], ..); entries beyond offsets2[B] are ignored.  NULL: everything is set 1.  V_total / F_total
   * are then CAPACITIES; skip_set1 renders set 2 alone (the object-only step). */
  int32_t V1, F1, skip_set1, reserved;
  const int32_t *vert_offsets2, *face_offsets2;
  const float *verts;          /* device [V_total,3] world (MoGe) space                       */
  const int32_t *faces;        /* device [F_total,3]                                          */
  const int32_t *vert_offsets; /* device [B+1]                                                */
  const int32_t *face_offsets; /* device [B+1]                                                */
  const float *fov_deg;        /* device [B] MoGe fov_x in degrees                            */
  const float *gt_normals;     /* device [B,H,W,3] target normal map                          */
  const uint8_t *gt_mask;      /* device [B,H,W]   valid mask of the normal loss              */
  const int32_t *n_valid;      /* device [B]       number of set pixels of gt_mask            */
  const float *gt_disp;        /* device [B,H,W]   target (normalised) disparity              */
  const float *gt_sil;         /* device [B,H,W]   target silhouette                          */
  float *losses;               /* device [B,8] OUT: l_normal, l_disp, l_sil, weighted total, min / max of the raw normal
                                  colours, max disparity, tile overflow flag                  */
  float *grad_verts;           /* device [V_total,3] OUT dE/d(verts), or NULL (forward only)  */
  int32_t *out_p2f;            /* optional device [B,H,W] OUT packed face index per pixel (-1 = background) */
  float *out_zbuf;             /* optional device [B,H,W] OUT view depth (-1 = background)    */
  float *out_nraw;             /* optional device [B,H,W,3] OUT blended normal colour before normalisation */
  void *workspace; size_t workspace_bytes;   /* >= foho_raster_workspace_bytes(desc), 256-byte aligned */
} foho_raster_desc;
size_t foho_raster_workspace_bytes(const foho_raster_desc *desc);
int foho_raster_losses_fwd_bwd(const foho_raster_desc *desc, void *cuda_stream);

/* ---------------------------------------------------------------------------------------------
 * Row f2, second part: surface extraction from the decoded volume and its backward -- where the reference calls kaolin's
 * FlexiCubes without weights (third_party_patches/hy3dgen/shapegen/pipelines.py:1142-1143,1393,1509,1642).  NOT a
 * restatement of FlexiCubes (its tables are not reproducible offline): Dual Marching Cubes as defined in
 * oracle/surface_oracle.py -- one dual vertex per sign-changing cube at the mean of its edge zero crossings, one quad per
 * interior sign-changing lattice edge, wound inside -> outside, split along (0, 2).  Meshes of the B images come out
 * packed: image b owns vertices [vert_offsets[b], vert_offsets[b+1]), triangles [face_offsets[b], ..) and unique edges
 * [edge_offsets[b], ..); counts stay on the device.  Orders are deterministic.  *flags: bit0 vertex capacity, bit1 face
 * capacity, bit2 edge capacity exceeded (the mesh is then truncated; the offsets never point beyond the capacities).
 * --------------------------------------------------------------------------------------------- */
typedef struct foho_dmc_desc {
  int32_t B, D;                /* volumes [B,D,D,D], negative inside, lattice linspace(-bound, bound, D)          */
  float bound;                 /* 1.10                                                                           */
  int32_t cap_verts, cap_faces, cap_edges;   /* capacities of the packed outputs (totals over the batch)         */
  int32_t index_base;          /* added to every vertex index written to `faces` (the renderer's joint vertex array) */
  int32_t reserved;
  const float *sdf;            /* device [B,D,D,D]                                                               */
  float *verts;                /* device [cap_verts,3] OUT, Hunyuan space                                        */
  int32_t *faces;              /* device [cap_faces,3] OUT                                                       */
  int32_t *edges;              /* device [cap_edges,2] OUT packed vertex indices (no base), or NULL              */
  int32_t *vert_offsets, *face_offsets, *edge_offsets;   /* device [B+1] OUT                                     */
  int32_t *cube_of_vert;       /* device [cap_verts] OUT: lattice cube of each vertex (needed by the backward)   */
  int32_t *flags;              /* device [1], OR-accumulated                                                     */
  void *workspace; size_t workspace_bytes;   /* >= foho_dmc_workspace_bytes(B, D), 256-byte aligned; shared by both calls */
} foho_dmc_desc;
size_t foho_dmc_workspace_bytes(int32_t B, int32_t D);
int foho_dmc_extract(const foho_dmc_desc *desc, void *cuda_stream);
/* grad_sdf [B,D,D,D] += d(verts)/d(sdf)^T grad_verts (of the extraction this descriptor last produced) */
int foho_dmc_backward(const foho_dmc_desc *desc, const float *grad_verts, float *grad_sdf, void *cuda_stream);

#ifdef __cplusplus
}
#endif
#endif /* FOHO_B200_H */
