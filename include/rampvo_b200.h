/*
 * rampvo_b200.h — C-ABI of librampvo_b200.so
 *
 * B200 (sm_100a) implementation of the RAMP-VO per-frame recurrent-update hot path
 * (uzh-rpg/rampvo): altcorr patch gather + patch<->frame correlation lookup, the
 * projective transform, the fastba Gauss-Newton / Schur bundle adjustment over the
 * patch graph, the patch-graph bookkeeping (neighbors, group plans) and the update
 * operator's non-GEMM stages.
 *
 * Conventions
 *   - every pointer is a DEVICE pointer unless the function name ends in `_host`;
 *   - every call is asynchronous on `stream` (a cudaStream_t passed as void*), except
 *     the `_host` variants which copy in, run, copy out and synchronise `stream`;
 *   - index arrays (ii, jj, kk, ix, jx) are int64, like the reference's `long` accessors;
 *   - return value 0 = success; non-zero = error, message via rvo_last_error();
 *     nothing ever calls exit()/abort() (the reference exit(1)s at block_e.cu:20-26,
 *     ba.cpp:151-152 — SURVEY.md §5);
 *   - poses are rows [tx,ty,tz,qx,qy,qz,qw] (ramp/lietorch/groups.py:273), world->camera.
 *
 * Each entry point cites the reference interface it replaces (paths relative to the
 * reference checkout, uzh-rpg/rampvo @ 9353aaa).
 */
#ifndef RAMPVO_B200_H
#define RAMPVO_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define RVO_ABI_VERSION 1

/* element types of feature tensors */
enum { RVO_F16 = 0, RVO_F32 = 1 };

/* error codes */
enum {
  RVO_OK = 0,
  RVO_ERR_ARG = 1,      /* bad argument (shape / dtype / alignment / null pointer) */
  RVO_ERR_CUDA = 2,     /* a CUDA runtime call failed */
  RVO_ERR_WORKSPACE = 3 /* workspace too small */
};

int rvo_abi_version(void);
const char* rvo_last_error(void);
/* compute capability of the current device as major*10+minor (100 on B200); <0 on error */
int rvo_device_cc(void);
/* number of kernels this library has launched so far in this process (library-internal cub sort /
 * scan passes are not counted) */
uint64_t rvo_launch_count(void);
/* SMs the persistent grids launched by the CALLING host thread may fill (1..148; 0 = all, the default).  Every heavy
 * kernel of this library keeps one CTA per SM, so two streams whose kernels are launched (or captured into CUDA
 * graphs) under disjoint budgets share the GPU spatially: Ramp_vo gives the encoder stream and the update stream
 * their own budgets (config SM_SPLIT).  Results do not depend on it. */
int rvo_set_sm_budget(int n_sms);
int rvo_get_sm_budget(void);

/* A strided view of a 4-D feature tensor, logical dims [N, C, H, W], strides in ELEMENTS.
 * Channels-last storage (sC == 1) selects the tensor-core fast path of rvo_corr_*. */
typedef struct {
  const void* data;
  int32_t dtype;           /* RVO_F16 / RVO_F32 */
  int32_t N, C, H, W;
  int64_t sN, sC, sH, sW;
} rvo_fmap_t;

/* ------------------------------------------------------------------ altcorr ---- */

/* cuda_corr.patchify_forward (ramp/altcorr/correlation.cpp:47, kernel
 * correlation_kernel.cu:17-47): patches[b,m,c,a,b'] = net[b,c,floor(y)+a-R,floor(x)+b'-R],
 * D = 2R+2, zero outside the map.  `net` is the [B,C,H,W] view (N == B); coords [B,M,2] f32
 * (x,y); out is a dense [B,M,C,D,D] tensor of net's dtype. Bit-exact gather. */
int rvo_patchify_forward(const rvo_fmap_t* net, const float* coords, int M, int radius,
                         void* out, void* stream);

/* altcorr.patchify(..., mode='bilinear') (ramp/altcorr/correlation.py:51-68): gather + 4-corner
 * blend fused, evaluated in fp32 with the reference's operation order
 * ((1-dy)*(1-dx))*P[a,b] + ((1-dy)*dx)*P[a,b+1] + (dy*(1-dx))*P[a+1,b] + (dy*dx)*P[a+1,b+1].
 * out dims [B,M,C,d,d], d = 2R+1, with caller-given element strides (so the gmap ring can be
 * written patch-pixel-major / channels-last); out_dtype RVO_F16 (rounded once) or RVO_F32. */
int rvo_patchify_bilinear(const rvo_fmap_t* net, const float* coords, int M, int radius,
                          void* out, int out_dtype, int64_t oB, int64_t oM, int64_t oC,
                          int64_t oH, int64_t oW, void* stream);

/* cuda_corr.forward (ramp/altcorr/correlation.cpp:27, kernel correlation_kernel.cu:83-136 plus the
 * host-side bilinear blend and permute at :221-232).
 *   fmap1: patch features, logical [Np, C, P, P] (N=Np, H=W=P)
 *   fmap2: frame features, logical [Nf, C, H2, W2]
 *   coords [E,2,P,P] f32 (x plane then y plane), ii[e] -> patch, jj[e] -> frame
 *   out    [E, 2R+1 (x offset), 2R+1 (y offset), P, P] of fmap1's dtype
 * fp32 accumulation and fp32 blend (the reference accumulates in the feature dtype). */
int rvo_corr_forward(const rvo_fmap_t* fmap1, const rvo_fmap_t* fmap2, const float* coords,
                     const int64_t* ii, const int64_t* jj, int E, int radius, void* out,
                     void* stream);

/* Ramp_vo.corr (ramp/Ramp_vo.py:175-182): both pyramid levels in ONE launch.
 *   level l reads pyr[l] at coords * scale[l]  (scale = 1, 0.25 on the hot path)
 *   patch index = kk[e] % pmod, frame index = jj[e] % fmod (ring buffers; pass 0 for no modulo)
 *   out [E, 7,7,P,P, nlevels] flattened to [E, 49*P*P*nlevels] (= 882) of fmap1's dtype, rows
 *   out_ld elements apart (0 = dense; the update operator uses 896 so that its first GEMM sees a
 *   K that is a multiple of 8 — the pad columns are never written). */
int rvo_corr_pyramid(const rvo_fmap_t* fmap1, const rvo_fmap_t* pyr, const float* scale,
                     int nlevels, const float* coords, const int64_t* kk, const int64_t* jj,
                     int64_t pmod, int64_t fmod, int E, int radius, void* out, int64_t out_ld,
                     void* stream);

/* The same lookup as rvo_corr_pyramid (radius 3, P = 3, C = 128, channels-last fp16) computed as
 * tile GEMMs on the tcgen05 tensor cores with TMEM accumulators: rows (edge, patch pixel, level) are
 * binned by the 16x16-position tile of the target frame that holds their 8x8 window, one CTA runs
 * D[128 rows, 256 positions] = A * B^T per tile with tcgen05.mma and extracts / blends each row's
 * window in the epilogue.  Output "tile layout":
 *   out[e*out_ld + ((lvl*9 + pix)*7 + a)*8 + b]   a = y offset, b = x offset in 0..6; b = 7 is a
 *   zero pad (one aligned 16-byte store per row); out_ld >= 504*nlevels, % 8 == 0.  (The reference layout of
 *   ramp/Ramp_vo.py:182 is index ((b*7 + a)*9 + pix)*nlevels + lvl.)
 * ws: rvo_corr_tiles_ws_bytes(pyr, nlevels, E) bytes of device scratch.  Limits: E * 9 * nlevels < 2^31 rows and
 * at most 24 576 tiles over all levels and ring frames (32 frames of up to ~1000 x 800 input pixels at 1/4 scale);
 * larger problems return RVO_ERR_ARG — use rvo_corr_pyramid. */
int64_t rvo_corr_tiles_ws_bytes(const rvo_fmap_t* pyr, int nlevels, int E);
int rvo_corr_tiles(const rvo_fmap_t* fmap1, const rvo_fmap_t* pyr, const float* scale, int nlevels,
                   const float* coords, const int64_t* kk, const int64_t* jj, int64_t pmod,
                   int64_t fmod, int E, void* out, int64_t out_ld, void* ws, int64_t ws_bytes,
                   void* stream);

/* Host-buffer variant of rvo_corr_pyramid: every pointer (also inside the rvo_fmap_t views) is
 * HOST memory; copies in, launches, copies the [E,882] result back, synchronises. */
int rvo_corr_pyramid_host(const rvo_fmap_t* fmap1, const rvo_fmap_t* pyr, const float* scale,
                          int nlevels, const float* coords, const int64_t* kk, const int64_t* jj,
                          int64_t pmod, int64_t fmod, int E, int radius, void* out, void* stream);

/* cuda_corr.backward (ramp/altcorr/correlation.cpp:50-55; correlation_kernel.cu:140-190,236-290): gradients of
 * rvo_corr_forward w.r.t. fmap1 [Np,C,P,P] and fmap2 [Nf,C,H2,W2] (strided views, fp16 / fp32) given corr_grad
 * [E,d,d,P,P] fp32 (x-offset dim first, d = 2R+1; the 4-corner blend adjoint is applied inside).  Outputs are DENSE
 * fp32 tensors of the logical shapes (zeroed here); the reference returns the inputs' dtype — the Python layer casts. */
int rvo_corr_backward(const rvo_fmap_t* fmap1, const rvo_fmap_t* fmap2, const float* coords, const int64_t* ii,
                      const int64_t* jj, const float* corr_grad, int E, int radius, float* fmap1_grad,
                      float* fmap2_grad, void* stream);

/* cuda_corr.patchify_backward (correlation_kernel.cu:50-80,306-331) for ONE map: patch_grad [M,C,D,D] (D = 2R+2,
 * fp16 / fp32), coords [M,2] -> net_grad [C,H,W] dense fp32 (zeroed here). */
int rvo_patchify_backward(const void* patch_grad, int dtype, const float* coords, int M, int C, int H, int W,
                          int radius, float* net_grad, void* stream);

/* ------------------------------------------------------- projective_ops -------- */

/* flags for rvo_transform */
enum { RVO_TF_TONLY = 1, RVO_TF_NOCLAMP = 2 };

/* pops.transform (ramp/projective_ops.py:50-101) fused into one kernel: iproj (:16-26),
 * Gij = T_j * T_i^-1 (lietorch se3.h:36-47), act4 (se3.h:53-56), proj with Z clamped at 0.1
 * (:29-47).
 *   poses [*,7], patches [*,3,P,P], intrinsics [*,4]; ii/jj/kk [E]
 *   coords_pp  [E,P,P,2]  (pops.transform layout) or NULL
 *   coords_cf  [E,2,P,P]  (Ramp_vo.reproject layout, Ramp_vo.py:192) or NULL
 *   depth_out  [E,P,P]    (the d of proj(depth=True)) or NULL
 *   valid_out  [E,P,P]    ((Z > 0.2) as float, `valid=True`) or NULL
 * RVO_TF_TONLY zeroes the rotation of Gij (flow_mag, :113); RVO_TF_NOCLAMP gives the un-clamped
 * projection of cuda_ba.reproject (ramp/fastba/ba_cuda.cu:379-429) with intrinsics[0] for
 * every frame. */
int rvo_transform(const float* poses, const float* patches, const float* intrinsics,
                  const int64_t* ii, const int64_t* jj, const int64_t* kk, int E, int P,
                  int flags, float* coords_pp, float* coords_cf, float* depth_out,
                  float* valid_out, void* stream);

/* pops.transform(jacobian=True) (ramp/projective_ops.py:68-96): additionally the centre-pixel
 * Jacobians Ji [E,2,6], Jj [E,2,6], Jz [E,2,1] and valid [E] = (Z > 0.2). */
int rvo_transform_jac(const float* poses, const float* patches, const float* intrinsics,
                      const int64_t* ii, const int64_t* jj, const int64_t* kk, int E, int P,
                      float* coords_pp, float* valid, float* Ji, float* Jj, float* Jz,
                      void* stream);

/* cuda_ba.reproject (ramp/fastba/ba.cpp:49-57, kernel ba_cuda.cu:379-429): coords [E,2,P,P]. */
int rvo_reproject(const float* poses, const float* patches, const float* intrinsics,
                  const int64_t* ii, const int64_t* jj, const int64_t* kk, int E, int P,
                  float* coords, void* stream);

/* pops.point_cloud (ramp/projective_ops.py:103-105) restricted to what Ramp_vo.update uses
 * (Ramp_vo.py:308-310): centre pixel, X/W normalised: points [m,3]. ix[k] = frame of patch k. */
int rvo_point_cloud(const float* poses, const float* patches, const float* intrinsics,
                    const int64_t* ix, int m, int P, float* points, void* stream);

/* pops.flow_mag (ramp/projective_ops.py:108-118): per edge, per patch pixel
 * beta*|x(i->j) - x(i->i)| + (1-beta)*|x_tonly(i->j) - x(i->i)|; out [E,P,P]. */
int rvo_flow_mag(const float* poses, const float* patches, const float* intrinsics,
                 const int64_t* ii, const int64_t* jj, const int64_t* kk, int E, int P,
                 float beta, float* out, void* stream);

/* ------------------------------------------------------------------ encoder ------ */

/* Normalisation / activation / residual glue of the encoder CNNs (ramp/extractor.py:8-57,
 * 288-311) on channels-last fp16 activations [npix, C] (C % 8 == 0):
 *   rvo_in_stats  sums[0..C) = per-channel sum, sums[C..2C) = sum of squares over the npix pixels
 *                 (InstanceNorm2d statistics, biased variance, extractor.py:30-34);
 *   rvo_in_apply  out = relu( R + relu(T) ) with T = IN(t) when sums_t != NULL else t, and
 *                 R = IN(res) / res / absent — the tail of ResidualBlock.forward (extractor.py:47-57)
 *                 and of conv1 -> norm1 -> relu1 (:295-297) in one pass.  eps = 1e-5. */
int rvo_in_stats(const void* x16, int64_t npix, int C, float* sums, void* stream);
int rvo_in_apply(const void* t16, const float* sums_t, const void* res16, const float* sums_res,
                 int64_t npix, int C, float eps, void* out16, void* stream);

/* One scale of the recurrent multi-scale stem (MultiScaleMergerDoubleNet.forward,
 * ramp/extractor.py:540-560) for one event stack [Ce,H,W] f32 and one image [Ci,H,W] f32: strided
 * conv_1 (k, stride, pad; :326-345) -> per-pixel LSTM cell for one step from a zero state
 * (:351-381) for both modalities -> super state ss <- W_ev [ss ; h_ev] + b and, when use_image,
 * ss <- W_im [ss ; h_im] + b (:404-411,446-452).  ss_prev16 (NULL = zeros) and ss_out16 are
 * channels-last fp16 [Ho,Wo,h], h in {16,32,64}.  `params` is one packed f32 buffer whose layout
 * rvo_stem_params_layout reports: offsets of conv_1 weight/bias (events), conv_1 weight/bias
 * (image), LSTM [i|g|o] rows of weight_ih and (bias_ih + bias_hh) (events), the same (image),
 * TRANSPOSED [2h,h] super-state weight and bias (events), the same (image). */
int rvo_stem_params_layout(int Ce, int Ci, int k, int h, int* offsets12, int* total);
int rvo_stem_forward(const float* params, int Ce, int Ci, int k, int stride, int pad, int h,
                     const float* events, const float* image, int H, int W, const void* ss_prev16,
                     int use_image, void* ss_out16, void* stream);

/* ------------------------------------------------------- VO state-machine helpers */

/* The damped-linear motion model of Ramp_vo.__call__ (ramp/Ramp_vo.py:356-363):
 * poses[n] = Exp(damping * Log(P1 * P2^-1)) * P1, P1 = poses[n-1], P2 = poses[n-2], with the
 * lietorch formulas (so3.h:115-215, se3.h:36-47,124-142).  One launch instead of ~150 tensor ops. */
int rvo_motion_model(float* poses, int n, float damping, void* stream);

/* Ramp_vo.keyframe's two motionmag calls (ramp/Ramp_vo.py:227-241) in one launch:
 * out4 = [sum flow_mag over edges (fi->fj), pixel count, sum over edges (fj->fi), pixel count];
 * mean = sum / count on the host side (flow_mag: ramp/projective_ops.py:108-118). */
int rvo_pair_flow(const float* poses, const float* patches, const float* intrinsics,
                  const int64_t* ii, const int64_t* jj, const int64_t* kk, int E, int P, int64_t fi,
                  int64_t fj, float beta, float* out4, void* stream);

/* ------------------------------------------------------------- patch-graph plan -- */

/* The bookkeeping the reference redoes with torch::_unique + host loops in every call
 * (ramp/fastba/ba.cpp:62-95 for neighbors, ramp/fastba/ba_cuda.cu:447-449 for the compact patch
 * numbering kx/ku, ramp/blocks.py:43 for the SoftAgg groups) is computed ONCE per graph, on the
 * device, with no host synchronisation: edges are sorted by (kk, jj, edge id) so that every
 * patch owns one contiguous segment.
 *   kmax / jmax: exclusive upper bounds on the values in kk / jj (sizes of the patch and pose
 *   tables); 0 = unknown (kk < 2^42, jj < 2^21 assumed).  Values must be non-negative.
 *   plan: caller-owned device buffer of rvo_plan_bytes(E) bytes. */
int64_t rvo_plan_bytes(int E);
int rvo_graph_plan(const int64_t* kk, const int64_t* jj, int E, int64_t kmax, int64_t jmax,
                   void* plan, int64_t plan_bytes, void* stream);
/* device pointers into a plan: count[0] = number of distinct patches U; perm[E] = edge ids in
 * sorted order; seg_of[E] = compact patch id of each sorted position (the reference's `ku`,
 * permuted); seg_start[U+1]; kx[U] = sorted distinct patch ids.  Any out pointer may be NULL. */
int rvo_plan_groups(const void* plan, int E, const int32_t** count, const int32_t** perm,
                    const int32_t** seg_of, const int32_t** seg_start, const int64_t** kx);

/* ------------------------------------------------------------------ fastba ------ */

/* cuda_ba.neighbors (ramp/fastba/ba.cpp:59-97) entirely on the device (the reference round-trips
 * through the CPU): group edges by kk, order each group by jj ascending (stable in edge index);
 * ix[e] = previous edge in that order or -1, jx[e] = next or -1.  Bit-exact.
 * rvo_neighbors builds a plan in `ws` (rvo_neighbors_ws_bytes(E) bytes) and links it;
 * rvo_plan_neighbors reuses an existing plan. */
int64_t rvo_neighbors_ws_bytes(int E);
int rvo_neighbors(const int64_t* kk, const int64_t* jj, int E, int64_t kmax, int64_t jmax,
                  int64_t* ix, int64_t* jx, void* ws, int64_t ws_bytes, void* stream);
int rvo_plan_neighbors(const void* plan, int E, int64_t* ix, int64_t* jx, void* stream);

/* cuda_ba.forward (ramp/fastba/ba.cpp:32-46 -> cuda_ba, ba_cuda.cu:433-582): `iterations`
 * Gauss-Newton steps with the patch (inverse-depth) block eliminated by a Schur complement;
 * updates poses[t0..t1) and the depths of every patch referenced by kk IN PLACE.
 *   poses [n_poses,7], patches [n_patches,3,P,P], intrinsics [*,4] (row 0 used, ba_cuda.cu:254-258)
 *   target [E,2], weight [E,2], lmbda [1] (device scalar), ii/jj/kk [E]
 *   PPF / eff_impl are accepted for signature parity: E is always kept block-sparse (one dense
 *   6N-row per patch, never the 6N x M matrix), so both settings give the same result.
 *   Frames outside [t0,t1) are fixed (the reference only guards the lower end, ba_cuda.cu:338-345).
 * ws: workspace of rvo_ba_ws_bytes(E, n_patches, t1-t0) bytes; it begins with the graph plan. */
int64_t rvo_ba_ws_bytes(int E, int64_t n_patches, int n_free);
int rvo_ba_forward(float* poses, float* patches, const float* intrinsics, const float* target,
                   const float* weight, const float* lmbda, const int64_t* ii, const int64_t* jj,
                   const int64_t* kk, int E, int64_t n_poses, int64_t n_patches, int P, int PPF,
                   int t0, int t1, int iterations, int eff_impl, void* ws, int64_t ws_bytes,
                   void* stream);

/* rvo_ba_forward with the edge grouping taken from an existing rvo_graph_plan(kk, jj, ...) — the
 * update operator of the same frame already built it (neighbors / SoftAgg groups), so the sort is
 * not repeated. */
int rvo_ba_forward_planned(float* poses, float* patches, const float* intrinsics,
                           const float* target, const float* weight, const float* lmbda,
                           const int64_t* ii, const int64_t* jj, const void* plan, int E,
                           int64_t n_poses, int64_t n_patches, int P, int t0, int t1, int iterations,
                           void* ws, int64_t ws_bytes, void* stream);

/* rvo_ba_forward_planned whose window start t0 lives in DEVICE memory (t0_dev[0]; the window is
 * [t0, t0 + n_free)): every host-side scalar of the call is then constant from frame to frame, so a
 * CUDA graph that contains the whole recurrent update can be replayed while the window slides. */
int rvo_ba_forward_dyn(float* poses, float* patches, const float* intrinsics, const float* target,
                       const float* weight, const float* lmbda, const int64_t* ii, const int64_t* jj,
                       const void* plan, int E, int64_t n_poses, int64_t n_patches, int P, int n_free,
                       const int32_t* t0_dev, int iterations, void* ws, int64_t ws_bytes,
                       void* stream);

/* rvo_ba_forward_dyn with the target formation of ramp/Ramp_vo.py:288-296 folded into the edge load:
 *   target[e] = coords[e, :, P/2, P/2] + delta[e]        (coords [E,2,P,P]: the reprojected patches)
 *   weight[e] = 0 where target lies outside [0,wd] x [0,ht]   (filter_features, ramp/utils.py:557-570)
 * weight_out (optional, [E,2]) receives the filtered confidences (Ramp_vo.last_weight). */
int rvo_ba_forward_fused(float* poses, float* patches, const float* intrinsics, const float* coords,
                         const float* delta, const float* weight, float ht, float wd, float* weight_out,
                         const float* lmbda, const int64_t* ii, const int64_t* jj, const void* plan, int E,
                         int64_t n_patches, int P, int n_free, const int32_t* t0_dev, int iterations,
                         void* ws, int64_t ws_bytes, void* stream);

/* rvo_ba_assemble with the fused target formation of rvo_ba_forward_fused (sharded graphs: local edges ->
 * [S | y] before the all-reduce). */
int rvo_ba_assemble_fused(const float* poses, const float* patches, const float* intrinsics, const float* coords,
                          const float* delta, const float* weight, float ht, float wd, float* weight_out,
                          const float* lmbda, const int64_t* ii, const int64_t* jj, int E, int64_t n_patches, int P,
                          int t0, int t1, float* Sy, void* ws, int64_t ws_bytes, void* stream);

/* One Gauss-Newton iteration in three steps, split where a patch graph sharded by source frame
 * needs its all-reduce:
 *   rvo_ba_plan      once per graph: sort / group the edges into `ws`;
 *   rvo_ba_assemble  local edges -> reduced camera system [S | y], row-major [6n, 6n+1], n = t1-t0,
 *                    BEFORE damping, full symmetric S.  Sy == NULL keeps it inside ws.  Per-patch
 *                    Q, u and E rows stay in ws for the back substitution;
 *   (the caller sums Sy over ranks — NCCL all-reduce)
 *   rvo_ba_solve     damping S += I*(1e-4*S+1), Cholesky, pose retraction T <- Exp(dX) T, local
 *                    depth back substitution dZ = Q (u - E^T dX) and depth retraction. */
int rvo_ba_plan(const int64_t* kk, const int64_t* jj, int E, int64_t n_poses, int64_t n_patches,
                int n_free, void* ws, int64_t ws_bytes, void* stream);
int rvo_ba_assemble(const float* poses, const float* patches, const float* intrinsics,
                    const float* target, const float* weight, const float* lmbda,
                    const int64_t* ii, const int64_t* jj, int E, int64_t n_patches, int P, int t0,
                    int t1, float* Sy, void* ws, int64_t ws_bytes, void* stream);
int rvo_ba_solve(float* poses, float* patches, const float* Sy, int E, int64_t n_patches, int P,
                 int t0, int t1, void* ws, int64_t ws_bytes, void* stream);

/* The replicated half of rvo_ba_solve alone (damping, Cholesky, pose retraction) for a rank that
 * owns no edge of the current window; ws: rvo_ba_ws_bytes(1, 1, t1-t0) bytes or more. */
int rvo_ba_solve_poses(float* poses, const float* Sy, int t0, int t1, void* ws, int64_t ws_bytes,
                       void* stream);

/* Host-buffer variant of rvo_ba_forward (all pointers HOST; poses/patches copied back). */
int rvo_ba_forward_host(float* poses, float* patches, const float* intrinsics, const float* target,
                        const float* weight, const float* lmbda, const int64_t* ii,
                        const int64_t* jj, const int64_t* kk, int E, int64_t n_poses,
                        int64_t n_patches, int P, int PPF, int t0, int t1, int iterations,
                        int eff_impl, void* stream);

/* ------------------------------------------------------------ update operator --- */

/* The non-GEMM stages of Update.forward (ramp/net.py:69-90).
 *
 * rvo_softagg: SoftAgg.forward's torch.unique + scatter_softmax + scatter_sum
 * (ramp/blocks.py:42-45) in one pass over a graph plan whose first sort key is the group id
 * (kk for agg_kk; ii*12345+jj for agg_ij, net.py:84-85):
 *   y[g, c] = sum_{e in g} fx[e,c] * softmax_{e in g}(gx[e,c])
 * fx, gx [E,C] (dtype RVO_F16/RVO_F32), y [U,C] (y_dtype), groups numbered in ascending key order
 * like torch.unique's inverse.  max_groups bounds U (0 = E).
 * rvo_expand_add: net[e,:] += hy[group(e),:]  (the `[:, jx]` expand of blocks.py:47-48 fused with
 * the residual add of net.py:84-85); net fp32 [E,C], C % 4 == 0.
 * rvo_gather_rows: out[e,:] = idx[e] >= 0 ? src[idx[e],:] : 0  (mask_ix * net[:, ix], net.py:78-82). */
int rvo_softagg(const void* fx, const void* gx, int dtype, const void* plan, int E, int C,
                int64_t max_groups, void* y, int y_dtype, void* stream);
int rvo_expand_add(const void* hy, int dtype, const void* plan, int E, int C, float* net,
                   void* stream);
int rvo_gather_rows(const float* src, const int64_t* idx, int E, int C, void* out, int out_dtype,
                    void* stream);

/* nn.Linear (+ nn.ReLU) of the update operator (ramp/net.py:36-67: c1, c2, corr[2], corr[5], gru gates /
 * residual MLPs; ramp/blocks.py:36-38: SoftAgg f, g, h) as ONE hand-written tcgen05 GEMM:
 *   y16[M, N] = act(x16[M, K] * w16[N, K]^T + bias16[N]),  fp16 operands, fp32 accumulation (TMEM), fp16 out
 * (torch.nn.functional.linear under autocast).  K must be 384 (the update operator's width: the weight slice
 * stays resident in shared memory), N a multiple of 192 (384 or 768 on the hot path); ldx / ldy = row pitches
 * in elements (multiples of 8, 16-byte aligned bases); bias16 may be NULL; relu != 0 applies max(., 0). */
int rvo_up_linear(const void* x16, int64_t ldx, const void* w16, const void* bias16, int M, int K, int N,
                  int relu, void* y16, int64_t ldy, void* stream);
/* rvo_up_linear whose input row r is x16[gather[r]] (gather[r] < 0: a zero row) — `mask_ix * net[:, ix]` of
 * ramp/net.py:78-82 folded into the operand load (the gathered copy is never materialised); gather == NULL is
 * rvo_up_linear. */
int rvo_up_linear_gather(const void* x16, int64_t ldx, const int64_t* gather, const void* w16, const void* bias16,
                         int M, int K, int N, int relu, void* y16, int64_t ldy, void* stream);

/* Fused row kernels of the mixed-precision update operator (C = 384; fp16 GEMM operands, fp32
 * hidden state and LayerNorm, eps 1e-3 — the dtypes Update.forward has under autocast,
 * ramp/Ramp_vo.py:280, SURVEY.md appendix "dtype drift").  Each replaces 3-8 elementwise / cast /
 * LayerNorm launches of the reference:
 *   rvo_up_ln_relu        y16 = relu(LN(x16))                               net.py:54-56
 *   rvo_up_add3_ln        net = LN(net + imap16[idx % mod] + h16) [; net16 = half(net)]  net.py:74-75, Ramp_vo.py:282
 *   rvo_up_add_cast       net += t16 [; net16 = half(net)]                  net.py:81-82
 *   rvo_up_softagg_fg     rvo_softagg on a fused [E,2C] = [f(x) | g(x)] GEMM output; rows of y past the
 *                         last group are zeroed                             blocks.py:42-45
 *   rvo_up_expand_add_ln  net[e] += hy16[group(e)]; then either net16 = half(net) (x32 == NULL) or
 *                         x32 = LN(net), x16 = half(x32)                    blocks.py:47-48, net.py:84-87
 *   rvo_up_gated_tail     y = x32 + sigmoid(a16) * r16 (blocks.py:30-31); mode 0: out32 = LN(y),
 *                         out16 = half(out32); mode 1: out32 = y, delta = Wd relu(y) + bd,
 *                         weight = sigmoid(Ww relu(y) + bw) with Wd, Ww [2,C] (net.py:63-67,90) */
int rvo_up_ln_relu(const void* x16, const float* gamma, const float* beta, int E, int C, void* y16,
                   void* stream);
int rvo_up_add3_ln(const float* net_in, const void* imap16, const int64_t* idx, int64_t mod,
                   const void* h16, const float* gamma, const float* beta, int E, int C,
                   float* net_out, void* net16_out, void* stream);
int rvo_up_add_cast(float* net, const void* t16, int E, int C, void* net16, void* stream);
int rvo_up_softagg_fg(const void* fg16, const void* plan, int E, int C, int64_t max_groups, void* y16,
                      void* stream);
int rvo_up_expand_add_ln(const void* hy16, const void* plan, int E, int C, float* net, void* net16,
                         const float* gamma, const float* beta, float* x32, void* x16, void* stream);
int rvo_up_gated_tail(const float* x32, const void* a16, const void* r16, int E, int C, int mode,
                      const float* gamma, const float* beta, float* out32, void* out16,
                      const float* Wd, const float* bd, const float* Ww, const float* bw, float* delta,
                      float* weight, void* stream);

/* The patch-graph step of a new frame when no keyframe was dropped, in two launches instead of a dozen tensor ops:
 * remove_factors of the edges whose source frame is < lim (ramp/Ramp_vo.py:203-208, order preserved) followed by the
 * forward edges (every patch of frames [max(n-r,0), n-1) -> frame n-1, Ramp_vo.py:312-318) and the backward edges
 * (every patch of frame n-1 -> frames [max(n-r,0), n), patch-major, Ramp_vo.py:320-325) of the new frame n-1
 * (n = number of frames AFTER the new one was added, M patches per frame, r = PATCH_LIFETIME).
 *   ii/jj/kk [E0] -> ii_out/jj_out/kk_out [E_new]; src_row [E_new] int32 = old row of a surviving edge, -1 for a new
 *   one; net_in [E0,C] -> net_out [E_new,C] (rows follow their edges, new edges start at zero; may be NULL).
 *   drop_k >= 0: keyframe drop_k was dropped first (ramp/Ramp_vo.py:249-262): its edges (ii == k or jj == k) go, and
 *   frames / patches behind it are renumbered (ii, jj > k: -1; kk of ii > k: -M) before the lim test; -1 = no drop.
 *   E_new is the caller's (host-side) edge count: status[0] (device float) is 0 when the device agrees, else its count + 1.
 *   tile_state: rvo_edges_step_tiles(E0) device words kept by the caller between calls (zeroed once); epoch: a value
 *   that differs from every earlier call on the same tile_state (a frame counter) — the CTAs of one launch exchange
 *   their counts through it without a reset pass. */
int64_t rvo_edges_step_tiles(int E0);
int rvo_edges_step(const int64_t* ii, const int64_t* jj, const int64_t* kk, int E0, int drop_k, int lim, int n, int M,
                   int r, int64_t* ii_out, int64_t* jj_out, int64_t* kk_out, int E_new, int32_t* src_row, float* status,
                   uint64_t* tile_state, uint32_t epoch, const float* net_in, int C, float* net_out, void* stream);

/* the hidden-state half of rvo_edges_step alone (net_out[e] = net_in[src_row[e]], 0 for src_row[e] < 0), so that it
 * can run on a side stream beside the head of the next update */
int rvo_net_rows(const float* net_in, const int32_t* src_row, int E, int C, float* net_out, void* stream);

/* ---- fused Linear chains of the update operator (csrc/up_chain.cu) ------------------------------- */

/* Update.forward (ramp/net.py:69-90) is five row-local stretches separated by four cross-row exchanges (the two
 * neighbour gathers of net.py:78-82 and the two SoftAgg reductions of net.py:84-85).  rvo_up_chain runs ONE such
 * stretch — up to 6 nn.Linear layers with everything between them — as one persistent tcgen05 kernel: a tile of
 * 128 edge rows stays in shared memory (fp16, the dtype Linear sees under autocast) from layer to layer, the
 * weights stream through a TMA ring, accumulators live in TMEM, and the element-wise / LayerNorm / residual /
 * gate / head arithmetic of the reference runs in the epilogues with its rounding points (fp16 Linear outputs,
 * fp32 residual stream and LayerNorm, eps 1e-3).
 *
 * prologue (how the A tile of layer 0 is formed)
 *   RVO_CHAIN_PRO_ROWS       rows of a16 [*, K0] fp16 (row pitch lda), row e or gather[e] (< 0: a zero row — the
 *                            mask_ix * net[:, ix] of net.py:78-82); K0 = layer[0].K may exceed 384 (streamed)
 *   RVO_CHAIN_PRO_EXPAND     half(x32[e] + hy_a[grp_a[e]])                      (blocks.py:47-48, net.py:84)
 *   RVO_CHAIN_PRO_EXPAND_LN  n = LN((x32[e] + hy_a[grp_a[e]]) + hy_b[grp_b[e]]); A = half(n); n is the residual of
 *                            the first GATED epilogue                           (net.py:84-85, gru[0] net.py:47)
 * epilogue of a layer, t = half(acc + bias)
 *   RVO_CHAIN_EPI_RELU        next A = relu(t)                                   (Linear + ReLU)
 *   RVO_CHAIN_EPI_LN_RELU     next A = half(relu(LN(t)))                         (corr[2:5], net.py:54-56)
 *   RVO_CHAIN_EPI_ADD3_LN     out32[e] = LN((net_in[e] + imap16[imap_idx[e] % imap_mod]) + t); out16 = half(out32)
 *                                                                                (net.py:74-75, Ramp_vo.py:282)
 *   RVO_CHAIN_EPI_RES         v = res32[e] + t; out32[e] = v; out16[e] = half(v) (if given); next A = half(v)
 *                                                                                (net.py:81-82)
 *   RVO_CHAIN_EPI_STORE16     layer.y16[e * ldy + c] = t; A unchanged            (SoftAgg f, g, h: blocks.py:36-38)
 *   RVO_CHAIN_EPI_GATE        gate = half(sigmoid(t)) kept for the next GATED epilogue; A unchanged (blocks.py:23)
 *   RVO_CHAIN_EPI_GATED_LN    y = residual + half(gate * t); m = LN(y); residual = m; next A = half(m)
 *                                                                                (blocks.py:30-31, net.py:49)
 *   RVO_CHAIN_EPI_GATED_HEADS y as above -> out32[e]; delta[e] = Wd relu(y) + bd; weight[e] = sigmoid(Ww relu(y) + bw)
 *                                                                                (net.py:63-67,90)
 * Every layer has N = 384 outputs; layers after the first have K = 384.  scratch32 / scratch16: per-CTA residual
 * and gate rows, rvo_up_chain_scratch_rows() rows of 384 floats / halves each (only for the GATED epilogues and
 * PRO_EXPAND_LN).
 * Layouts.  Row-major [M, 384]: a16, net_in, imap16, hy_a / hy_b, out16, y16 and the out32 of GATED_HEADS (the
 * arrays other code reads).  TILE-BLOCKED fp32, private to the chains: x32, res32 and the out32 of RES / ADD3_LN —
 * ceil(M / 128) * 128 rows, element (e, c) at ((e / 128 * 96 + c / 4) * 128 + e % 128) * 4 + c % 4 floats, so that
 * the 32 rows a warp owns in an epilogue are one contiguous 512-byte access.  out32 may alias res32. */
enum { RVO_CHAIN_PRO_ROWS = 0, RVO_CHAIN_PRO_EXPAND = 1, RVO_CHAIN_PRO_EXPAND_LN = 2 };
enum {
  RVO_CHAIN_EPI_RELU = 0, RVO_CHAIN_EPI_LN_RELU = 1, RVO_CHAIN_EPI_ADD3_LN = 2, RVO_CHAIN_EPI_RES = 3,
  RVO_CHAIN_EPI_STORE16 = 4, RVO_CHAIN_EPI_GATE = 5, RVO_CHAIN_EPI_GATED_LN = 6, RVO_CHAIN_EPI_GATED_HEADS = 7
};
#define RVO_CHAIN_MAX_LAYERS 6
typedef struct {
  const void* w16;      /* [384, K] fp16, nn.Linear layout (row pitch K halves, K % 8 == 0) */
  const void* bias16;   /* [384] fp16 */
  const float* gamma;   /* LayerNorm weight / bias of the epilogue (LN_RELU, ADD3_LN, GATED_LN) */
  const float* beta;
  void* y16;            /* STORE16 destination, already offset to its first column */
  int64_t ldy;          /* row pitch of y16 in halves (multiple of 8) */
  int32_t K;
  int32_t epilogue;
} rvo_chain_layer_t;
typedef struct {
  int32_t M, n_layers, prologue, reserved;
  const void* a16; int64_t lda; const int64_t* gather;                                   /* PRO_ROWS */
  const float* x32; const void* hy_a; const int32_t* grp_a; const void* hy_b; const int32_t* grp_b;
  const float* pro_gamma; const float* pro_beta;                                         /* PRO_EXPAND(_LN) */
  const float* net_in; const void* imap16; const int64_t* imap_idx; int64_t imap_mod;    /* EPI_ADD3_LN */
  const float* res32;                                                                    /* EPI_RES */
  float* out32; void* out16;
  const float* Wd; const float* bd; const float* Ww; const float* bw; float* delta; float* weight;  /* heads */
  float* scratch32; void* scratch16;
  rvo_chain_layer_t layer[RVO_CHAIN_MAX_LAYERS];
} rvo_chain_t;
int64_t rvo_up_chain_scratch_rows(void);
int rvo_up_chain(const rvo_chain_t* chain, void* stream);
/* The chain kernel runs as thread-block clusters whose CTAs share every weight stage by TMA multicast (the L2 is read
 * once per cluster).  rvo_up_chain_info reports the cluster size picked for this device (the largest of 4 / 2 / 1
 * whose first wave covers the GPU) and the CTAs of one wave; rvo_up_chain_set_cluster forces 1, 2 or 4 (0 = automatic
 * again) — a benchmarking knob, results do not depend on it. */
int rvo_up_chain_info(int* cluster_size, int* ctas);
int rvo_up_chain_set_cluster(int cluster_size);

/* group id of every EDGE (not sorted position) of a plan: grp[e] = seg_of[s] with perm[s] = e — the index the
 * expand prologues of rvo_up_chain take. */
int rvo_plan_edge_groups(const void* plan, int E, const int32_t** grp);

/* ---- encoder convolutions on the tensor cores (csrc/conv_tc.cu) --------------------------------- */

/* nn.Conv2d of the RAMP encoder CNNs (ramp/extractor.py:12-13,47,79,88,286: 7x7 s2, 3x3 s1/s2, 1x1 s1/s2) as a
 * tcgen05 implicit GEMM, channels-last fp16 in / fp32 accumulate / fp16 out:
 *   src0 [H,W,C0] (+ src1 [H,W,C1]: the input is their channel concatenation, extractor.py:302,309 torch.cat),
 *   w_packed [Cout, Kpad] fp16 with k = (ky*ks + kx)*(C0+C1) + ci, zero padded to Kpad = rvo_conv2d_kpad(ks, C0+C1),
 *   bias [Cout] fp32 (or NULL), out [Ho,Wo,Cout] fp16 with Ho = (H + 2 pad - ks)/stride + 1.
 * stats (optional, [2*Cout] fp32): per-channel sum and sum of squares of the rounded outputs = the statistics pass
 * of nn.InstanceNorm2d (extractor.py:30-34), in the layout rvo_in_apply consumes; ACCUMULATED into the buffer,
 * which the caller zeroes (one memset for all the layers of a frame instead of one per layer).
 * Limits: C0, C1 multiples of 8; Cout a multiple of 16; ks*ks*(C0+C1) <= 832. */
int rvo_conv2d_kpad(int ks, int Cin);
int rvo_conv2d_nhwc(const void* src0, int C0, const void* src1, int C1, int H, int W, int ks, int stride, int pad,
                    const void* w_packed, const float* bias, int Cout, void* out, float* stats, void* stream);

/* ---- SingleScale encoder front end (csrc/scene_lstm.cu) ----------------------------------------- */

/* ramp/extractor.py:233-261 (MergerLSTMsceneEncoder.forward) for one event stack [Ce,H,W] + one image [Ci,H,W]
 * (fp32, planar): per-pixel nn.LSTM step for both modalities with the (h, c) state carried across calls, the
 * events/image presence tests (any(x != 0)) evaluated on the device, the shared 1x1 super-state convolution.
 *   params       packed fp32 block of rvo_scene_lstm_params_floats(Ce, Ci) floats:
 *                events  W_ih [4h,Ce] | W_hh [4h,h] | b_ih + b_hh [4h];  image likewise;  superstate W [h,2h] | b [h]
 *   state_ev/im  [2,h,H*W] fp32 (h then c), super_state [h,H*W] fp32: read unless `first`, always written
 *   flags        2 x int32 device scratch (presence flags of this call)
 *   out16        fp16 channels-last [H,W,16]: the new super state, channel 15 zero (h = 15) */
int rvo_scene_lstm_params_floats(int Ce, int Ci);
int rvo_scene_lstm_forward(const float* params, int Ce, int Ci, const float* events, const float* image, int H,
                           int W, float* state_ev, float* state_im, float* super_state, int32_t* flags, int first,
                           void* out16, void* stream);

/* ---- per-frame glue (csrc/frame_ops.cu) ------------------------------------------------------- */

/* Event-biased patch selection: ramp/utils.py:186-226 (get_coords_from_topk_events) with nms_image
 * (:157-183).  events [C,H,W] fp32 (one event stack), H and W multiples of 4; the top-M cells of the transposed,
 * NMS-filtered mean |event| map.  border: border_suppression_size; nms: odd window (non_max_supp_rad, 0 = off).
 *   gather_order == 0: coords [M,2] fp32 = (idx * (1/H'), idx % H') — torch's CUDA true division by a scalar
 *     multiplies by the reciprocal — ordered by value descending, ties by ascending flat index: what torch.topk
 *     returns on CUDA whenever its final key/value sort is stable (k > 32); idx_out / val_out optional.
 *   gather_order != 0: idx_out [M] int64 and val_out [M] fp32 in torch.topk's PRE-sort order (values above the
 *     k-th value by ascending index, then ties with it by ascending index); for k <= 32 torch finishes with an
 *     unstable bitonic sort, which the caller reproduces by applying the same sort to val_out.  coords unused.
 * ws: rvo_select_ws_bytes(H, W) bytes of device scratch. */
int64_t rvo_select_ws_bytes(int H, int W);
int rvo_select_patches(const float* events, int C, int H, int W, int M, int border, int nms, int gather_order,
                       float* coords, int64_t* idx_out, float* val_out, void* ws, int64_t ws_bytes, void* stream);

/* Pyramid level 2 (ramp/Ramp_vo.py:381, F.avg_pool2d(fmap, 4, 4)) on a channels-last map [H,W,C] ->
 * [H/4,W/4,C]; fp32 accumulation, one rounding. */
int rvo_pyramid_level2(const void* fmap, int dtype, int H, int W, int C, void* out, void* stream);

/* Up to 8 device-to-device copies in one launch: the ring-buffer slot writes of a new frame
 * (ramp/Ramp_vo.py:374-381).  src / dst / bytes are HOST arrays of n entries. */
int rvo_copy_segments(const void* const* src, void* const* dst, const int64_t* bytes, int n, void* stream);

/* The state writes of a new frame (ramp/Ramp_vo.py:345-372) in one launch: tstamps[n] = counter,
 * intrinsics[n] = intr4 (HOST array, already / RES), index[n+1,:] = n+1, index_map[n+1] = m_next,
 * colors[n] = uint8((clr[:, [2,1,0]] + 0.5) * 127.5), patches[n] = patches_new with the inverse depth set to
 * depth_rand[m] (median_frames == 0; null keeps the staged value) or to the lower median of the depths of the
 * previous `median_frames` frames (torch.median semantics, :370-371; evaluated on the patch-centre depths, which
 * every pixel of a patch shares). */
int rvo_frame_commit(const float* patches_new, const float* clr, const float* depth_rand, float* patches,
                     int64_t* tstamps, float* intrinsics, int64_t* index, int64_t* index_map, uint8_t* colors,
                     const float* intr4, int n, int M, int P, int N, int64_t counter, int64_t m_next,
                     int median_frames, void* stream);

/* utils/transformers.py:128-161 (EventToStack_Numpy): events in arrival order (x, y uint16 pixel, p polarity
 * value) -> [bins,H,W] stack; bin = int32(float32(bins * i) / N), out-of-image events dropped, the sum cast to
 * int8.  stack_f32 (required) is the tensor the encoder consumes; stack_i8 (optional) the int8 stack itself. */
int rvo_event_stack(const uint16_t* x, const uint16_t* y, const float* p, int64_t n_events, int bins, int H,
                    int W, float* stack_f32, int8_t* stack_i8, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* RAMPVO_B200_H */
