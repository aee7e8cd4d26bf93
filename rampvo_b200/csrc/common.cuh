// common.cuh — shared helpers for librampvo_b200 (sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include "../../include/rampvo_b200.h"

namespace rvo {

// thread-local error message behind rvo_last_error()
void set_error(const char* fmt, ...);
int cuda_fail(cudaError_t e, const char* what);

#define RVO_CHECK_ARG(cond, ...)                  \
  do {                                            \
    if (!(cond)) {                                \
      ::rvo::set_error(__VA_ARGS__);              \
      return RVO_ERR_ARG;                         \
    }                                             \
  } while (0)

#define RVO_CUDA(call)                                            \
  do {                                                            \
    cudaError_t e__ = (call);                                     \
    if (e__ != cudaSuccess) return ::rvo::cuda_fail(e__, #call);  \
  } while (0)

// every kernel launch of this library goes through RVO_LAUNCH_CHECK: it also feeds the launch
// counter behind rvo_launch_count() (bench.py's "gpu_launches").
extern unsigned long long g_launches;

#define RVO_LAUNCH_CHECK(name)                                          \
  do {                                                                  \
    ::rvo::g_launches++;                                                \
    cudaError_t e__ = cudaGetLastError();                               \
    if (e__ != cudaSuccess) return ::rvo::cuda_fail(e__, name);         \
  } while (0)

static inline int cdiv(int64_t a, int64_t b) { return (int)((a + b - 1) / b); }

constexpr int kNumSMs = 148;  // B200

// SMs the persistent grids of the calling host thread may fill (default: all).  The two-stream frame loop gives the
// encoder stream and the update stream disjoint budgets: every heavy kernel here holds one CTA per SM (shared
// memory), so a kernel launched with a grid of `budget` CTAs leaves the other SMs to the other stream's kernels
// — spatial partitioning without MPS / green contexts.  rvo_set_sm_budget() sets it (CUDA graphs bake the grid
// sizes in at capture).
extern thread_local int g_sm_budget;
static inline int sm_budget() { return g_sm_budget; }

// ---- SE3 device helpers (quaternion xyzw, as ramp/fastba/ba_cuda.cu:36-174 and
// ramp/lietorch/include/so3.h:55-60, se3.h:36-56 define the algebra) ----

__device__ __forceinline__ void rot_q(const float* q, const float* X, float* Y) {
  // Y = R(q) X :  uv = 2 (q_v x X);  Y = X + q_w uv + q_v x uv
  float uv0 = 2.0f * (q[1] * X[2] - q[2] * X[1]);
  float uv1 = 2.0f * (q[2] * X[0] - q[0] * X[2]);
  float uv2 = 2.0f * (q[0] * X[1] - q[1] * X[0]);
  Y[0] = X[0] + q[3] * uv0 + (q[1] * uv2 - q[2] * uv1);
  Y[1] = X[1] + q[3] * uv1 + (q[2] * uv0 - q[0] * uv2);
  Y[2] = X[2] + q[3] * uv2 + (q[0] * uv1 - q[1] * uv0);
}

// Gij = Tj * Ti^-1 :  q_ij = q_j (x) conj(q_i),  t_ij = t_j - R(q_ij) t_i
__device__ __forceinline__ void rel_se3(const float* ti, const float* qi, const float* tj,
                                        const float* qj, float* tij, float* qij) {
  qij[0] = -qj[3] * qi[0] + qj[0] * qi[3] - qj[1] * qi[2] + qj[2] * qi[1];
  qij[1] = -qj[3] * qi[1] + qj[1] * qi[3] - qj[2] * qi[0] + qj[0] * qi[2];
  qij[2] = -qj[3] * qi[2] + qj[2] * qi[3] - qj[0] * qi[1] + qj[1] * qi[0];
  qij[3] = qj[3] * qi[3] + qj[0] * qi[0] + qj[1] * qi[1] + qj[2] * qi[2];
  float r[3];
  rot_q(qij, ti, r);
  tij[0] = tj[0] - r[0];
  tij[1] = tj[1] - r[1];
  tij[2] = tj[2] - r[2];
}

// Y = Ad(G)^T X for G = (t, q):  [R^T a, R^T b + R^T (a x t)],  X = [a, b]
__device__ __forceinline__ void adjT_se3(const float* t, const float* q, const float* X, float* Y) {
  float qinv[4] = {-q[0], -q[1], -q[2], q[3]};
  rot_q(qinv, X, Y);
  rot_q(qinv, X + 3, Y + 3);
  float u[3], v[3];
  u[0] = t[2] * X[1] - t[1] * X[2];
  u[1] = t[0] * X[2] - t[2] * X[0];
  u[2] = t[1] * X[0] - t[0] * X[1];
  rot_q(qinv, u, v);
  Y[3] += v[0];
  Y[4] += v[1];
  Y[5] += v[2];
}

__device__ __forceinline__ void load_pose(const float* __restrict__ poses, int64_t i, float* t,
                                          float* q) {
  const float* p = poses + i * 7;
  t[0] = p[0]; t[1] = p[1]; t[2] = p[2];
  q[0] = p[3]; q[1] = p[4]; q[2] = p[5]; q[3] = p[6];
}

}  // namespace rvo
