// update_ops.cu — the non-GEMM stages of the recurrent update operator (ramp/net.py:69-90,
// ramp/blocks.py:33-50) for sm_100a.
//
//   softagg   SoftAgg.forward's scatter_softmax + scatter_sum (blocks.py:42-45) as ONE pass over
//             pre-grouped edges: a CTA owns one group (a patch, or an (i,j) frame pair), its threads
//             own channels, and walk the group's edges with an online (running-max) softmax.  f and g
//             are each read exactly once, fully coalesced; nothing is scattered, no atomics, no
//             torch.unique.  Bound: HBM, 2 * E * C * sizeof(half) bytes in + U * C out.
//   expand    h(y)[:, group_of_edge] added to the hidden state (blocks.py:47-48 + net.py:84-85).
//   gather    mask * net[:, ix] (net.py:78-82) for the neighbour MLPs.
#include "common.cuh"

namespace rvo {

template <typename T>
__device__ __forceinline__ float ldf(const T* p);
template <>
__device__ __forceinline__ float ldf<float>(const float* p) { return *p; }
template <>
__device__ __forceinline__ float ldf<__half>(const __half* p) { return __half2float(*p); }

template <typename T>
__device__ __forceinline__ void stf(T* p, float v);
template <>
__device__ __forceinline__ void stf<float>(float* p, float v) { *p = v; }
template <>
__device__ __forceinline__ void stf<__half>(__half* p, float v) { *p = __float2half_rn(v); }

// y[g, c] = sum_e f[e,c] * softmax_e(g[e,c])  over the edges e of group g (sorted positions
// seg_start[g] .. seg_start[g+1]; edge id = perm[s]).
template <typename T, typename TO>
__global__ void __launch_bounds__(128)
softagg_kernel(const int32_t* __restrict__ count, const int32_t* __restrict__ perm,
               const int32_t* __restrict__ seg_start, const T* __restrict__ fx,
               const T* __restrict__ gx, int C, int cap, TO* __restrict__ y) {
  int U = count[0];
  if (U > cap) U = cap;
  for (int grp = blockIdx.x; grp < U; grp += gridDim.x) {
    const int s0 = seg_start[grp], s1 = seg_start[grp + 1];
    for (int c = threadIdx.x; c < C; c += blockDim.x) {
      float m = -INFINITY, den = 0.0f, num = 0.0f;
      for (int s = s0; s < s1; s++) {
        const size_t off = (size_t)perm[s] * C + c;
        const float gv = ldf<T>(gx + off), fv = ldf<T>(fx + off);
        const float mn = fmaxf(m, gv);
        const float sc = __expf(m - mn), w = __expf(gv - mn);
        den = den * sc + w;
        num = num * sc + w * fv;
        m = mn;
      }
      stf<TO>(y + (size_t)grp * C + c, s1 > s0 ? num / den : 0.0f);
    }
  }
}

// net[e, c] += hy[group_of_edge[e], c];  group_of_edge[perm[s]] = seg_of[s]
template <typename T>
__global__ void __launch_bounds__(256)
expand_add_kernel(const int32_t* __restrict__ perm, const int32_t* __restrict__ seg_of,
                  const T* __restrict__ hy, int E, int C, float* __restrict__ net) {
  const int64_t total = (int64_t)E * (C / 4);
  for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < total;
       t += (int64_t)gridDim.x * blockDim.x) {
    const int s = (int)(t / (C / 4)), c4 = (int)(t - (int64_t)s * (C / 4));
    const int e = perm[s], grp = seg_of[s];
    float4 v = reinterpret_cast<float4*>(net + (size_t)e * C)[c4];
    const T* h = hy + (size_t)grp * C + c4 * 4;
    v.x += ldf<T>(h); v.y += ldf<T>(h + 1); v.z += ldf<T>(h + 2); v.w += ldf<T>(h + 3);
    reinterpret_cast<float4*>(net + (size_t)e * C)[c4] = v;
  }
}

// out[e, :] = (idx[e] >= 0) ? src[idx[e], :] : 0   (fp32 in, fp16/fp32 out)
template <typename TO>
__global__ void __launch_bounds__(256)
gather_rows_kernel(const float* __restrict__ src, const int64_t* __restrict__ idx, int E, int C,
                   TO* __restrict__ out) {
  const int64_t total = (int64_t)E * C;
  for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < total;
       t += (int64_t)gridDim.x * blockDim.x) {
    const int e = (int)(t / C), c = (int)(t - (int64_t)e * C);
    const int64_t j = idx[e];
    stf<TO>(out + t, j >= 0 ? src[(size_t)j * C + c] : 0.0f);
  }
}

static inline int grid_cap(int64_t n, int per_sm) {
  const int64_t cap = (int64_t)kNumSMs * per_sm;
  return (int)(n < 1 ? 1 : (n > cap ? cap : n));
}

}  // namespace rvo

using namespace rvo;

extern "C" int rvo_softagg(const void* fx, const void* gx, int dtype, const void* plan, int E, int C,
                           int64_t max_groups, void* y, int y_dtype, void* stream) {
  RVO_CHECK_ARG(E >= 0 && C >= 1, "rvo_softagg: E=%d C=%d", E, C);
  if (E == 0) return RVO_OK;
  RVO_CHECK_ARG(fx && gx && plan && y, "rvo_softagg: null pointer");
  const int32_t *count, *perm, *seg_start;
  int rc = rvo_plan_groups(plan, E, &count, &perm, nullptr, &seg_start, nullptr);
  if (rc != RVO_OK) return rc;
  const int cap = (int)((max_groups > 0 && max_groups < E) ? max_groups : E);
  const int grid = grid_cap(cap, 16);
  cudaStream_t st = (cudaStream_t)stream;
#define RVO_SA(TI, TO) \
  softagg_kernel<TI, TO><<<grid, 128, 0, st>>>(count, perm, seg_start, (const TI*)fx, (const TI*)gx, C, cap, (TO*)y)
  if (dtype == RVO_F16 && y_dtype == RVO_F16) RVO_SA(__half, __half);
  else if (dtype == RVO_F16 && y_dtype == RVO_F32) RVO_SA(__half, float);
  else if (dtype == RVO_F32 && y_dtype == RVO_F32) RVO_SA(float, float);
  else if (dtype == RVO_F32 && y_dtype == RVO_F16) RVO_SA(float, __half);
  else RVO_CHECK_ARG(false, "rvo_softagg: unsupported dtypes %d -> %d", dtype, y_dtype);
#undef RVO_SA
  RVO_LAUNCH_CHECK("softagg_kernel");
  return RVO_OK;
}

extern "C" int rvo_expand_add(const void* hy, int dtype, const void* plan, int E, int C, float* net,
                              void* stream) {
  RVO_CHECK_ARG(E >= 0 && C >= 4 && C % 4 == 0, "rvo_expand_add: E=%d C=%d", E, C);
  if (E == 0) return RVO_OK;
  RVO_CHECK_ARG(hy && plan && net, "rvo_expand_add: null pointer");
  const int32_t *perm, *seg_of;
  int rc = rvo_plan_groups(plan, E, nullptr, &perm, &seg_of, nullptr, nullptr);
  if (rc != RVO_OK) return rc;
  const int grid = grid_cap(((int64_t)E * (C / 4) + 255) / 256, 16);
  cudaStream_t st = (cudaStream_t)stream;
  if (dtype == RVO_F16) expand_add_kernel<__half><<<grid, 256, 0, st>>>(perm, seg_of, (const __half*)hy, E, C, net);
  else if (dtype == RVO_F32) expand_add_kernel<float><<<grid, 256, 0, st>>>(perm, seg_of, (const float*)hy, E, C, net);
  else RVO_CHECK_ARG(false, "rvo_expand_add: dtype %d", dtype);
  RVO_LAUNCH_CHECK("expand_add_kernel");
  return RVO_OK;
}

extern "C" int rvo_gather_rows(const float* src, const int64_t* idx, int E, int C, void* out,
                               int out_dtype, void* stream) {
  RVO_CHECK_ARG(E >= 0 && C >= 1, "rvo_gather_rows: E=%d C=%d", E, C);
  if (E == 0) return RVO_OK;
  RVO_CHECK_ARG(src && idx && out, "rvo_gather_rows: null pointer");
  const int grid = grid_cap(((int64_t)E * C + 255) / 256, 16);
  cudaStream_t st = (cudaStream_t)stream;
  if (out_dtype == RVO_F16) gather_rows_kernel<__half><<<grid, 256, 0, st>>>(src, idx, E, C, (__half*)out);
  else if (out_dtype == RVO_F32) gather_rows_kernel<float><<<grid, 256, 0, st>>>(src, idx, E, C, (float*)out);
  else RVO_CHECK_ARG(false, "rvo_gather_rows: dtype %d", out_dtype);
  RVO_LAUNCH_CHECK("gather_rows_kernel");
  return RVO_OK;
}
