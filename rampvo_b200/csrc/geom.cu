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
  int64_t cap = (int64_t)kNumSMs * 16;
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
