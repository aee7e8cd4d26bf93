// geom.cu — fused projective transform kernels (ramp/projective_ops.py, cuda_ba.reproject).
//
// One thread per (edge, patch pixel): the 7-float poses and 4-float intrinsics are tiny and
// L1/L2 resident, so recomputing the relative pose per pixel is cheaper than a second launch or a
// shuffle, and it makes every output store fully coalesced.  Bound: HBM, ~276 B/edge.
#include "common.cuh"

namespace rvo {

struct EdgeGeom {
  float tij[3], qij[4];
};

__device__ __forceinline__ void normalize_q(float* q) {
  const float s = rsqrtf(q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3]);
  q[0] *= s; q[1] *= s; q[2] *= s; q[3] *= s;
}

// unit == true: lietorch semantics (quaternions re-normalised on load, so3.h:31-37);
// unit == false: the raw arithmetic of cuda_ba (ba_cuda.cu:74-85).
__device__ __forceinline__ void edge_rel(const float* __restrict__ poses, int64_t i, int64_t j,
                                         bool tonly, EdgeGeom& g, bool unit = true) {
  float ti[3], qi[4], tj[3], qj[4];
  load_pose(poses, i, ti, qi);
  load_pose(poses, j, tj, qj);
  if (unit) { normalize_q(qi); normalize_q(qj); }
  rel_se3(ti, qi, tj, qj, g.tij, g.qij);
  if (tonly) {  // projective_ops.py:59-60: rotation part of Gij <- identity
    g.qij[0] = 0.f; g.qij[1] = 0.f; g.qij[2] = 0.f; g.qij[3] = 1.f;
  }
}

__device__ __forceinline__ void act4(const EdgeGeom& g, const float* X0, float* X1) {
  rot_q(g.qij, X0, X1);
  X1[0] += X0[3] * g.tij[0];
  X1[1] += X0[3] * g.tij[1];
  X1[2] += X0[3] * g.tij[2];
  X1[3] = X0[3];
}

template <bool NOCLAMP>
__global__ void __launch_bounds__(256)
transform_kernel(const float* __restrict__ poses, const float* __restrict__ patches,
                 const float* __restrict__ intr, const int64_t* __restrict__ ii,
                 const int64_t* __restrict__ jj, const int64_t* __restrict__ kk, int E, int P,
                 int tonly, float* __restrict__ coords_pp, float* __restrict__ coords_cf,
                 float* __restrict__ depth_out, float* __restrict__ valid_out) {
  const int PP = P * P;
  const int64_t total = (int64_t)E * PP;
  for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < total;
       t += (int64_t)gridDim.x * blockDim.x) {
    const int e = (int)(t / PP);
    const int p = (int)(t - (int64_t)e * PP);
    const int64_t i = ii[e], j = jj[e], k = kk[e];
    EdgeGeom g;
    edge_rel(poses, i, j, tonly != 0, g, !NOCLAMP);
    const float* Ki = intr + (NOCLAMP ? 0 : i * 4);
    const float* Kj = intr + (NOCLAMP ? 0 : j * 4);
    const float* pk = patches + k * 3 * PP;
    float X0[4], X1[4];
    X0[0] = (pk[p] - Ki[2]) / Ki[0];
    X0[1] = (pk[PP + p] - Ki[3]) / Ki[1];
    X0[2] = 1.0f;
    X0[3] = pk[2 * PP + p];
    act4(g, X0, X1);
    float x, y, d;
    if (NOCLAMP) {  // ba_cuda.cu:422-423
      x = Kj[0] * (X1[0] / X1[2]) + Kj[2];
      y = Kj[1] * (X1[1] / X1[2]) + Kj[3];
      d = 1.0f / X1[2];
    } else {  // projective_ops.py:40-42
      d = 1.0f / fmaxf(X1[2], 0.1f);
      x = Kj[0] * (d * X1[0]) + Kj[2];
      y = Kj[1] * (d * X1[1]) + Kj[3];
    }
    if (coords_pp) {
      float2 v = make_float2(x, y);
      reinterpret_cast<float2*>(coords_pp)[t] = v;
    }
    if (coords_cf) {
      coords_cf[(int64_t)e * 2 * PP + p] = x;
      coords_cf[(int64_t)e * 2 * PP + PP + p] = y;
    }
    if (depth_out) depth_out[t] = d;
    if (valid_out) valid_out[t] = (X1[2] > 0.2f) ? 1.0f : 0.0f;
  }
}

// Centre-pixel Jacobians, one thread per edge (projective_ops.py:68-96).
__global__ void __launch_bounds__(256)
transform_jac_kernel(const float* __restrict__ poses, const float* __restrict__ patches,
                     const float* __restrict__ intr, const int64_t* __restrict__ ii,
                     const int64_t* __restrict__ jj, const int64_t* __restrict__ kk, int E, int P,
                     float* __restrict__ valid, float* __restrict__ Ji, float* __restrict__ Jj,
                     float* __restrict__ Jz) {
  const int PP = P * P;
  const int c = (P / 2) * P + (P / 2);
  for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < E; e += gridDim.x * blockDim.x) {
    const int64_t i = ii[e], j = jj[e], k = kk[e];
    EdgeGeom g;
    edge_rel(poses, i, j, false, g);
    const float* Ki = intr + i * 4;
    const float* Kj = intr + j * 4;
    const float* pk = patches + k * 3 * PP;
    float X0[4], X1[4];
    X0[0] = (pk[c] - Ki[2]) / Ki[0];
    X0[1] = (pk[PP + c] - Ki[3]) / Ki[1];
    X0[2] = 1.0f;
    X0[3] = pk[2 * PP + c];
    act4(g, X0, X1);
    const float X = X1[0], Y = X1[1], Z = X1[2], H = X1[3];
    const float fx = Kj[0], fy = Kj[1];
    const float d = (fabsf(Z) > 0.2f) ? 1.0f / Z : 0.0f;
    // Jj = Jp * Ja
    float J0[6] = {fx * d * H, 0.f, -fx * X * d * d * H, -fx * X * d * d * Y,
                   fx * d * Z + fx * X * d * d * X, -fx * d * Y};
    float J1[6] = {0.f, fy * d * H, -fy * Y * d * d * H, -fy * d * Z - fy * Y * d * d * Y,
                   fy * Y * d * d * X, fy * d * X};
    float A0[6], A1[6];
    adjT_se3(g.tij, g.qij, J0, A0);
    adjT_se3(g.tij, g.qij, J1, A1);
#pragma unroll
    for (int q = 0; q < 6; q++) {
      Jj[(int64_t)e * 12 + q] = J0[q];
      Jj[(int64_t)e * 12 + 6 + q] = J1[q];
      Ji[(int64_t)e * 12 + q] = -A0[q];
      Ji[(int64_t)e * 12 + 6 + q] = -A1[q];
    }
    Jz[(int64_t)e * 2 + 0] = fx * d * g.tij[0] - fx * X * d * d * g.tij[2];
    Jz[(int64_t)e * 2 + 1] = fy * d * g.tij[1] - fy * Y * d * d * g.tij[2];
    valid[e] = (Z > 0.2f) ? 1.0f : 0.0f;
  }
}

__global__ void __launch_bounds__(256)
point_cloud_kernel(const float* __restrict__ poses, const float* __restrict__ patches,
                   const float* __restrict__ intr, const int64_t* __restrict__ ix, int m, int P,
                   float* __restrict__ points) {
  const int PP = P * P;
  const int c = (P / 2) * P + (P / 2);
  for (int k = blockIdx.x * blockDim.x + threadIdx.x; k < m; k += gridDim.x * blockDim.x) {
    const int64_t i = ix[k];
    float t[3], q[4];
    load_pose(poses, i, t, q);
    const float* K = intr + i * 4;
    const float* pk = patches + (int64_t)k * 3 * PP;
    float X0[3] = {(pk[c] - K[2]) / K[0], (pk[PP + c] - K[3]) / K[1], 1.0f};
    const float d = pk[2 * PP + c];
    // T^-1 = (q^-1, -(q^-1 t))  (se3.h:36-38); act4: R^-1 X + t_inv * d
    float qinv[4] = {-q[0], -q[1], -q[2], q[3]};
    float rt[3], rx[3];
    rot_q(qinv, t, rt);
    rot_q(qinv, X0, rx);
    points[(int64_t)k * 3 + 0] = (rx[0] + (-rt[0]) * d) / d;
    points[(int64_t)k * 3 + 1] = (rx[1] + (-rt[1]) * d) / d;
    points[(int64_t)k * 3 + 2] = (rx[2] + (-rt[2]) * d) / d;
  }
}

__global__ void __launch_bounds__(256)
flow_mag_kernel(const float* __restrict__ poses, const float* __restrict__ patches,
                const float* __restrict__ intr, const int64_t* __restrict__ ii,
                const int64_t* __restrict__ jj, const int64_t* __restrict__ kk, int E, int P,
                float beta, float* __restrict__ out) {
  const int PP = P * P;
  const int64_t total = (int64_t)E * PP;
  for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < total;
       t += (int64_t)gridDim.x * blockDim.x) {
    const int e = (int)(t / PP);
    const int p = (int)(t - (int64_t)e * PP);
    const int64_t i = ii[e], j = jj[e], k = kk[e];
    const float* Ki = intr + i * 4;
    const float* Kj = intr + j * 4;
    const float* pk = patches + k * 3 * PP;
    float X0[4], X1[4];
    X0[0] = (pk[p] - Ki[2]) / Ki[0];
    X0[1] = (pk[PP + p] - Ki[3]) / Ki[1];
    X0[2] = 1.0f;
    X0[3] = pk[2 * PP + p];
    float c[3][2];
#pragma unroll
    for (int v = 0; v < 3; v++) {
      EdgeGeom g;
      const int64_t jv = (v == 0) ? i : j;
      edge_rel(poses, i, jv, v == 2, g);
      const float* K = (v == 0) ? Ki : Kj;
      act4(g, X0, X1);
      const float d = 1.0f / fmaxf(X1[2], 0.1f);
      c[v][0] = K[0] * (d * X1[0]) + K[2];
      c[v][1] = K[1] * (d * X1[1]) + K[3];
    }
    const float f1 = sqrtf((c[1][0] - c[0][0]) * (c[1][0] - c[0][0]) +
                           (c[1][1] - c[0][1]) * (c[1][1] - c[0][1]));
    const float f2 = sqrtf((c[2][0] - c[0][0]) * (c[2][0] - c[0][0]) +
                           (c[2][1] - c[0][1]) * (c[2][1] - c[0][1]));
    out[t] = beta * f1 + (1.0f - beta) * f2;
  }
}

static inline int grid_for(int64_t n) {
  int64_t b = (n + 255) / 256;
  int64_t cap = (int64_t)sm_budget() * 16;
  return (int)(b < 1 ? 1 : (b > cap ? cap : b));
}

}  // namespace rvo

using namespace rvo;

extern "C" int rvo_transform(const float* poses, const float* patches, const float* intrinsics,
                             const int64_t* ii, const int64_t* jj, const int64_t* kk, int E, int P,
                             int flags, float* coords_pp, float* coords_cf, float* depth_out,
                             float* valid_out, void* stream) {
  RVO_CHECK_ARG(E >= 0 && P >= 1 && P <= 9, "rvo_transform: bad E=%d P=%d", E, P);
  if (E == 0) return RVO_OK;
  RVO_CHECK_ARG(poses && patches && intrinsics && ii && jj && kk, "rvo_transform: null input");
  RVO_CHECK_ARG(coords_pp || coords_cf, "rvo_transform: no output requested");
  cudaStream_t st = (cudaStream_t)stream;
  const int grid = grid_for((int64_t)E * P * P);
  if (flags & RVO_TF_NOCLAMP)
    transform_kernel<true><<<grid, 256, 0, st>>>(poses, patches, intrinsics, ii, jj, kk, E, P,
                                                 flags & RVO_TF_TONLY, coords_pp, coords_cf,
                                                 depth_out, valid_out);
  else
    transform_kernel<false><<<grid, 256, 0, st>>>(poses, patches, intrinsics, ii, jj, kk, E, P,
                                                  flags & RVO_TF_TONLY, coords_pp, coords_cf,
                                                  depth_out, valid_out);
  RVO_LAUNCH_CHECK("transform_kernel");
  return RVO_OK;
}

extern "C" int rvo_transform_jac(const float* poses, const float* patches, const float* intrinsics,
                                 const int64_t* ii, const int64_t* jj, const int64_t* kk, int E,
                                 int P, float* coords_pp, float* valid, float* Ji, float* Jj,
                                 float* Jz, void* stream) {
  RVO_CHECK_ARG(E >= 0 && P >= 1 && P <= 9, "rvo_transform_jac: bad E=%d P=%d", E, P);
  if (E == 0) return RVO_OK;
  RVO_CHECK_ARG(poses && patches && intrinsics && ii && jj && kk && valid && Ji && Jj && Jz,
                "rvo_transform_jac: null pointer");
  cudaStream_t st = (cudaStream_t)stream;
  if (coords_pp) {
    transform_kernel<false><<<grid_for((int64_t)E * P * P), 256, 0, st>>>(
        poses, patches, intrinsics, ii, jj, kk, E, P, 0, coords_pp, nullptr, nullptr, nullptr);
    RVO_LAUNCH_CHECK("transform_kernel");
  }
  transform_jac_kernel<<<grid_for(E), 256, 0, st>>>(poses, patches, intrinsics, ii, jj, kk, E, P,
                                                    valid, Ji, Jj, Jz);
  RVO_LAUNCH_CHECK("transform_jac_kernel");
  return RVO_OK;
}

extern "C" int rvo_reproject(const float* poses, const float* patches, const float* intrinsics,
                             const int64_t* ii, const int64_t* jj, const int64_t* kk, int E, int P,
                             float* coords, void* stream) {
  return rvo_transform(poses, patches, intrinsics, ii, jj, kk, E, P, RVO_TF_NOCLAMP, nullptr,
                       coords, nullptr, nullptr, stream);
}

extern "C" int rvo_point_cloud(const float* poses, const float* patches, const float* intrinsics,
                               const int64_t* ix, int m, int P, float* points, void* stream) {
  RVO_CHECK_ARG(m >= 0 && P >= 1 && P <= 9, "rvo_point_cloud: bad m=%d P=%d", m, P);
  if (m == 0) return RVO_OK;
  RVO_CHECK_ARG(poses && patches && intrinsics && ix && points, "rvo_point_cloud: null pointer");
  point_cloud_kernel<<<grid_for(m), 256, 0, (cudaStream_t)stream>>>(poses, patches, intrinsics, ix,
                                                                    m, P, points);
  RVO_LAUNCH_CHECK("point_cloud_kernel");
  return RVO_OK;
}

extern "C" int rvo_flow_mag(const float* poses, const float* patches, const float* intrinsics,
                            const int64_t* ii, const int64_t* jj, const int64_t* kk, int E, int P,
                            float beta, float* out, void* stream) {
  RVO_CHECK_ARG(E >= 0 && P >= 1 && P <= 9, "rvo_flow_mag: bad E=%d P=%d", E, P);
  if (E == 0) return RVO_OK;
  RVO_CHECK_ARG(poses && patches && intrinsics && ii && jj && kk && out,
                "rvo_flow_mag: null pointer");
  flow_mag_kernel<<<grid_for((int64_t)E * P * P), 256, 0, (cudaStream_t)stream>>>(
      poses, patches, intrinsics, ii, jj, kk, E, P, beta, out);
  RVO_LAUNCH_CHECK("flow_mag_kernel");
  return RVO_OK;
}

// ------------------------------------------------------------------ VO state-machine helpers ----
//
// The reference evaluates these with dozens of tiny tensor ops per frame (lietorch log/exp/mul on two
// poses, boolean-mask edge selection + flow_mag + mean + .item()); each is one small kernel here.

namespace rvo {

__device__ __forceinline__ void cross3(const float* a, const float* b, float* c) {
  c[0] = a[1] * b[2] - a[2] * b[1];
  c[1] = a[2] * b[0] - a[0] * b[2];
  c[2] = a[0] * b[1] - a[1] * b[0];
}

// quaternion product (xyzw)
__device__ __forceinline__ void qmul(const float* a, const float* b, float* c) {
  c[0] = a[3] * b[0] + a[0] * b[3] + a[1] * b[2] - a[2] * b[1];
  c[1] = a[3] * b[1] - a[0] * b[2] + a[1] * b[3] + a[2] * b[0];
  c[2] = a[3] * b[2] + a[0] * b[1] - a[1] * b[0] + a[2] * b[3];
  c[3] = a[3] * b[3] - a[0] * b[0] - a[1] * b[1] - a[2] * b[2];
}

constexpr float kLieEps = 1e-6f;  // ramp/lietorch/include/common.h:7

// poses[n] = Exp(damping * Log(P1 * P2^-1)) * P1 with P1 = poses[n-1], P2 = poses[n-2]
// (ramp/Ramp_vo.py:356-363; so3.h:115-215, se3.h:36-47,124-142)
__global__ void motion_model_kernel(float* __restrict__ poses, int n, float damping) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  float t1[3], q1[4], t2[3], q2[4];
  load_pose(poses, n - 1, t1, q1);
  load_pose(poses, n - 2, t2, q2);
  normalize_q(q1);
  normalize_q(q2);
  // P2^-1
  float q2i[4] = {-q2[0], -q2[1], -q2[2], q2[3]}, t2i[3], r[3];
  rot_q(q2i, t2, r);
  t2i[0] = -r[0]; t2i[1] = -r[1]; t2i[2] = -r[2];
  // D = P1 * P2^-1
  float qd[4], td[3];
  qmul(q1, q2i, qd);
  normalize_q(qd);
  rot_q(q1, t2i, r);
  td[0] = t1[0] + r[0]; td[1] = t1[1] + r[1]; td[2] = t1[2] + r[2];
  // phi = Log(qd)  (so3.h:115-151)
  const float sq = qd[0] * qd[0] + qd[1] * qd[1] + qd[2] * qd[2], w = qd[3];
  float f;
  if (sq < kLieEps * kLieEps) {
    f = 2.0f / w - (2.0f / 3.0f) * sq / (w * w * w);
  } else {
    const float nn = sqrtf(sq);
    if (fabsf(w) < kLieEps) f = (w > 0 ? 3.14159265358979323846f : -3.14159265358979323846f) / nn;
    else f = 2.0f * atanf(nn / w) / nn;
  }
  float phi[3] = {f * qd[0], f * qd[1], f * qd[2]};
  // tau = Vinv(phi) * td  (so3.h:193-208): Vinv = I - 0.5 Phi + c2 Phi^2
  const float th = sqrtf(phi[0] * phi[0] + phi[1] * phi[1] + phi[2] * phi[2]);
  const float half = 0.5f * th;
  const float c2 = (th < kLieEps) ? (1.0f / 12.0f)
                                  : (1.0f - th * cosf(half) / (2.0f * sinf(half))) / (th * th);
  float pt[3], ppt[3], tau[3];
  cross3(phi, td, pt);
  cross3(phi, pt, ppt);
  for (int a = 0; a < 3; a++) tau[a] = td[a] - 0.5f * pt[a] + c2 * ppt[a];
  // xi = damping * [tau, phi]
  for (int a = 0; a < 3; a++) { tau[a] *= damping; phi[a] *= damping; }
  // Exp(xi)  (so3.h:153-191, se3.h:134-142)
  const float th2 = phi[0] * phi[0] + phi[1] * phi[1] + phi[2] * phi[2];
  const float thn = sqrtf(th2);
  float imag, real, a1, b1;
  if (thn < kLieEps) {
    const float th4 = th2 * th2;
    imag = 0.5f - (1.0f / 48.0f) * th2 + (1.0f / 3840.0f) * th4;
    real = 1.0f - (1.0f / 8.0f) * th2 + (1.0f / 384.0f) * th4;
    a1 = 0.5f - (1.0f / 24.0f) * th2;
    b1 = (1.0f / 6.0f) - (1.0f / 120.0f) * th2;
  } else {
    imag = sinf(0.5f * thn) / thn;
    real = cosf(0.5f * thn);
    a1 = (1.0f - cosf(thn)) / th2;
    b1 = (thn - sinf(thn)) / (th2 * thn);
  }
  float qe[4] = {imag * phi[0], imag * phi[1], imag * phi[2], real};
  normalize_q(qe);
  float te[3];
  cross3(phi, tau, pt);
  cross3(phi, pt, ppt);
  for (int a = 0; a < 3; a++) te[a] = tau[a] + a1 * pt[a] + b1 * ppt[a];
  // out = Exp(xi) * P1
  float qo[4], to[3];
  qmul(qe, q1, qo);
  normalize_q(qo);
  rot_q(qe, t1, r);
  to[0] = te[0] + r[0]; to[1] = te[1] + r[1]; to[2] = te[2] + r[2];
  float* o = poses + (size_t)n * 7;
  o[0] = to[0]; o[1] = to[1]; o[2] = to[2];
  o[3] = qo[0]; o[4] = qo[1]; o[5] = qo[2]; o[6] = qo[3];
}

// out[0] = sum of flow_mag over edges with (ii,jj) == (fi,fj), out[1] = their pixel count,
// out[2], out[3] = the same for (fj,fi)   (Ramp_vo.motionmag x2 in keyframe(), Ramp_vo.py:227-241)
__global__ void __launch_bounds__(256)
pair_flow_kernel(const float* __restrict__ poses, const float* __restrict__ patches,
                 const float* __restrict__ intr, const int64_t* __restrict__ ii,
                 const int64_t* __restrict__ jj, const int64_t* __restrict__ kk, int E, int P,
                 int64_t fi, int64_t fj, float beta, float* __restrict__ out) {
  const int PP = P * P;
  float s[2] = {0.f, 0.f}, c[2] = {0.f, 0.f};
  for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < E; e += gridDim.x * blockDim.x) {
    const int64_t i = ii[e], j = jj[e];
    int which = -1;
    if (i == fi && j == fj) which = 0;
    else if (i == fj && j == fi) which = 1;
    if (which < 0) continue;
    const float* Ki = intr + i * 4;
    const float* Kj = intr + j * 4;
    const float* pk = patches + kk[e] * 3 * PP;
    EdgeGeom g0, g1, g2;
    edge_rel(poses, i, i, false, g0);
    edge_rel(poses, i, j, false, g1);
    edge_rel(poses, i, j, true, g2);
    float acc = 0.f;
    for (int p = 0; p < PP; p++) {
      float X0[4] = {(pk[p] - Ki[2]) / Ki[0], (pk[PP + p] - Ki[3]) / Ki[1], 1.0f, pk[2 * PP + p]};
      float X1[4], cc[3][2];
      const EdgeGeom* gs[3] = {&g0, &g1, &g2};
#pragma unroll
      for (int v = 0; v < 3; v++) {
        act4(*gs[v], X0, X1);
        const float* K = (v == 0) ? Ki : Kj;
        const float d = 1.0f / fmaxf(X1[2], 0.1f);
        cc[v][0] = K[0] * (d * X1[0]) + K[2];
        cc[v][1] = K[1] * (d * X1[1]) + K[3];
      }
      const float f1 = sqrtf((cc[1][0] - cc[0][0]) * (cc[1][0] - cc[0][0]) + (cc[1][1] - cc[0][1]) * (cc[1][1] - cc[0][1]));
      const float f2 = sqrtf((cc[2][0] - cc[0][0]) * (cc[2][0] - cc[0][0]) + (cc[2][1] - cc[0][1]) * (cc[2][1] - cc[0][1]));
      acc += beta * f1 + (1.0f - beta) * f2;
    }
    s[which] += acc;
    c[which] += (float)PP;
  }
#pragma unroll
  for (int w = 0; w < 2; w++) {
    float a = s[w], b = c[w];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      a += __shfl_xor_sync(0xffffffffu, a, o);
      b += __shfl_xor_sync(0xffffffffu, b, o);
    }
    if ((threadIdx.x & 31) == 0 && b > 0.f) {
      atomicAdd(&out[2 * w], a);
      atomicAdd(&out[2 * w + 1], b);
    }
  }
}

}  // namespace rvo

extern "C" int rvo_motion_model(float* poses, int n, float damping, void* stream) {
  RVO_CHECK_ARG(poses && n >= 2, "rvo_motion_model: needs two previous poses (n=%d)", n);
  motion_model_kernel<<<1, 32, 0, (cudaStream_t)stream>>>(poses, n, damping);
  RVO_LAUNCH_CHECK("motion_model_kernel");
  return RVO_OK;
}

extern "C" int rvo_pair_flow(const float* poses, const float* patches, const float* intrinsics,
                             const int64_t* ii, const int64_t* jj, const int64_t* kk, int E, int P,
                             int64_t fi, int64_t fj, float beta, float* out4, void* stream) {
  RVO_CHECK_ARG(E >= 0 && P >= 1 && P <= 9, "rvo_pair_flow: bad E=%d P=%d", E, P);
  RVO_CHECK_ARG(out4, "rvo_pair_flow: null output");
  RVO_CUDA(cudaMemsetAsync(out4, 0, 4 * sizeof(float), (cudaStream_t)stream));
  if (E == 0) return RVO_OK;
  RVO_CHECK_ARG(poses && patches && intrinsics && ii && jj && kk, "rvo_pair_flow: null pointer");
  pair_flow_kernel<<<grid_for(E), 256, 0, (cudaStream_t)stream>>>(poses, patches, intrinsics, ii, jj, kk,
                                                                  E, P, fi, fj, beta, out4);
  RVO_LAUNCH_CHECK("pair_flow_kernel");
  return RVO_OK;
}
