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
#include <stdlib.h>

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

__global__ void gather_rows384_kernel(const float* __restrict__ src, const int64_t* __restrict__ idx,
                                      int E, __half* __restrict__ out);

static inline int grid_cap(int64_t n, int per_sm) {
  const int64_t cap = (int64_t)sm_budget() * per_sm;
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
  if (out_dtype == RVO_F16 && C == 384) {
    const int64_t b = ((int64_t)E * 32 + 255) / 256;
    gather_rows384_kernel<<<(int)(b > sm_budget() * 8 ? sm_budget() * 8 : b), 256, 0, st>>>(src, idx, E, (__half*)out);
    RVO_LAUNCH_CHECK("gather_rows384_kernel");
    return RVO_OK;
  }
  if (out_dtype == RVO_F16) gather_rows_kernel<__half><<<grid, 256, 0, st>>>(src, idx, E, C, (__half*)out);
  else if (out_dtype == RVO_F32) gather_rows_kernel<float><<<grid, 256, 0, st>>>(src, idx, E, C, (float*)out);
  else RVO_CHECK_ARG(false, "rvo_gather_rows: dtype %d", out_dtype);
  RVO_LAUNCH_CHECK("gather_rows_kernel");
  return RVO_OK;
}

// ------------------------------------------------------------------ fused row kernels (C = 384) ----
//
// One warp per edge row; lane l owns elements [128 j + 4 l, 128 j + 4 l + 4), j = 0..2, so every
// load / store is a fully coalesced 16-byte (fp32) or 8-byte (fp16) vector access.  LayerNorm is
// evaluated in fp32 with eps = 1e-3 like the reference modules under autocast (ramp/net.py:42-58).

namespace rvo {

constexpr int kC = 384;
constexpr int kChunks = kC / 128;

struct Row {
  float v[kChunks][4];
};

__device__ __forceinline__ void row_load_f32(const float* __restrict__ p, int lane, Row& r) {
#pragma unroll
  for (int j = 0; j < kChunks; j++) {
    const float4 t = reinterpret_cast<const float4*>(p + 128 * j)[lane];
    r.v[j][0] = t.x; r.v[j][1] = t.y; r.v[j][2] = t.z; r.v[j][3] = t.w;
  }
}
__device__ __forceinline__ void row_load_f16(const __half* __restrict__ p, int lane, Row& r) {
#pragma unroll
  for (int j = 0; j < kChunks; j++) {
    const uint2 t = reinterpret_cast<const uint2*>(p + 128 * j)[lane];
    const float2 a = __half22float2(*reinterpret_cast<const __half2*>(&t.x));
    const float2 b = __half22float2(*reinterpret_cast<const __half2*>(&t.y));
    r.v[j][0] = a.x; r.v[j][1] = a.y; r.v[j][2] = b.x; r.v[j][3] = b.y;
  }
}
__device__ __forceinline__ void row_store_f32(float* __restrict__ p, int lane, const Row& r) {
#pragma unroll
  for (int j = 0; j < kChunks; j++)
    reinterpret_cast<float4*>(p + 128 * j)[lane] = make_float4(r.v[j][0], r.v[j][1], r.v[j][2], r.v[j][3]);
}
__device__ __forceinline__ void row_store_f16(__half* __restrict__ p, int lane, const Row& r) {
#pragma unroll
  for (int j = 0; j < kChunks; j++) {
    const __half2 a = __floats2half2_rn(r.v[j][0], r.v[j][1]);
    const __half2 b = __floats2half2_rn(r.v[j][2], r.v[j][3]);
    uint2 t;
    t.x = *reinterpret_cast<const uint32_t*>(&a);
    t.y = *reinterpret_cast<const uint32_t*>(&b);
    reinterpret_cast<uint2*>(p + 128 * j)[lane] = t;
  }
}

__device__ __forceinline__ float warp_sum_f(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// in-place LayerNorm of a row held in registers (two-pass: mean, then centred variance)
__device__ __forceinline__ void row_layer_norm(Row& r, const float* __restrict__ gamma,
                                               const float* __restrict__ beta, int lane, float eps) {
  float s = 0.f;
#pragma unroll
  for (int j = 0; j < kChunks; j++)
#pragma unroll
    for (int q = 0; q < 4; q++) s += r.v[j][q];
  const float mean = warp_sum_f(s) * (1.0f / kC);
  float ss = 0.f;
#pragma unroll
  for (int j = 0; j < kChunks; j++)
#pragma unroll
    for (int q = 0; q < 4; q++) { const float d = r.v[j][q] - mean; ss += d * d; }
  const float rstd = rsqrtf(warp_sum_f(ss) * (1.0f / kC) + eps);
#pragma unroll
  for (int j = 0; j < kChunks; j++) {
    const float4 g = reinterpret_cast<const float4*>(gamma + 128 * j)[lane];
    const float4 b = reinterpret_cast<const float4*>(beta + 128 * j)[lane];
    r.v[j][0] = (r.v[j][0] - mean) * rstd * g.x + b.x;
    r.v[j][1] = (r.v[j][1] - mean) * rstd * g.y + b.y;
    r.v[j][2] = (r.v[j][2] - mean) * rstd * g.z + b.z;
    r.v[j][3] = (r.v[j][3] - mean) * rstd * g.w + b.w;
  }
}

#define RVO_ROW_LOOP(e, E) \
  for (int e = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; e < (E); e += (gridDim.x * blockDim.x) >> 5)

// y16 = relu(LN(x16))                                   (Update.corr[3:5], net.py:54-56)
__global__ void __launch_bounds__(256)
ln_relu_kernel(const __half* __restrict__ x, const float* __restrict__ gamma,
               const float* __restrict__ beta, int E, __half* __restrict__ y) {
  const int lane = threadIdx.x & 31;
  RVO_ROW_LOOP(e, E) {
    Row r;
    row_load_f16(x + (size_t)e * kC, lane, r);
    row_layer_norm(r, gamma, beta, lane, 1e-3f);
#pragma unroll
    for (int j = 0; j < kChunks; j++)
#pragma unroll
      for (int q = 0; q < 4; q++) r.v[j][q] = fmaxf(r.v[j][q], 0.f);
    row_store_f16(y + (size_t)e * kC, lane, r);
  }
}

// net32 = LN(net32 + imap16[idx[e] % mod] + h16)        (net.py:74-75, Ramp_vo.py:282)
__global__ void __launch_bounds__(256)
add3_ln_kernel(const float* __restrict__ net_in, const __half* __restrict__ imap,
               const int64_t* __restrict__ idx, int64_t mod, const __half* __restrict__ h,
               const float* __restrict__ gamma, const float* __restrict__ beta, int E,
               float* __restrict__ net_out, __half* __restrict__ net16_out) {
  const int lane = threadIdx.x & 31;
  RVO_ROW_LOOP(e, E) {
    Row a, b, c;
    row_load_f32(net_in + (size_t)e * kC, lane, a);
    int64_t k = idx[e];
    if (mod > 0) k %= mod;
    row_load_f16(imap + (size_t)k * kC, lane, b);
    row_load_f16(h + (size_t)e * kC, lane, c);
#pragma unroll
    for (int j = 0; j < kChunks; j++)
#pragma unroll
      for (int q = 0; q < 4; q++) a.v[j][q] = (a.v[j][q] + b.v[j][q]) + c.v[j][q];
    row_layer_norm(a, gamma, beta, lane, 1e-3f);
    row_store_f32(net_out + (size_t)e * kC, lane, a);
    if (net16_out) row_store_f16(net16_out + (size_t)e * kC, lane, a);
  }
}

// net32 += t16 ; optionally net16 = half(net32)         (net.py:81-82)
__global__ void __launch_bounds__(256)
add_cast_kernel(float* __restrict__ net, const __half* __restrict__ t, int E,
                __half* __restrict__ net16) {
  const int lane = threadIdx.x & 31;
  RVO_ROW_LOOP(e, E) {
    Row a, b;
    row_load_f32(net + (size_t)e * kC, lane, a);
    row_load_f16(t + (size_t)e * kC, lane, b);
#pragma unroll
    for (int j = 0; j < kChunks; j++)
#pragma unroll
      for (int q = 0; q < 4; q++) a.v[j][q] += b.v[j][q];
    row_store_f32(net + (size_t)e * kC, lane, a);
    if (net16) row_store_f16(net16 + (size_t)e * kC, lane, a);
  }
}

// net32[e] += hy16[group(e)] ; optionally net16 = half(net32); x32 = LN(net32), x16 = half(x32)
// (blocks.py:47-48 expand + net.py:84-85 residual [+ the first LayerNorm of Update.gru, net.py:47-52])
__global__ void __launch_bounds__(256)
expand_add_ln_kernel(const int32_t* __restrict__ perm, const int32_t* __restrict__ seg_of,
                     const __half* __restrict__ hy, int E, float* __restrict__ net,
                     __half* __restrict__ net16, const float* __restrict__ gamma,
                     const float* __restrict__ beta, float* __restrict__ x32,
                     __half* __restrict__ x16) {
  const int lane = threadIdx.x & 31;
  RVO_ROW_LOOP(s, E) {
    const int e = perm[s], grp = seg_of[s];
    Row a, b;
    row_load_f32(net + (size_t)e * kC, lane, a);
    row_load_f16(hy + (size_t)grp * kC, lane, b);
#pragma unroll
    for (int j = 0; j < kChunks; j++)
#pragma unroll
      for (int q = 0; q < 4; q++) a.v[j][q] += b.v[j][q];
    if (x32) {
      row_layer_norm(a, gamma, beta, lane, 1e-3f);
      row_store_f32(x32 + (size_t)e * kC, lane, a);
      row_store_f16(x16 + (size_t)e * kC, lane, a);
    } else {
      row_store_f32(net + (size_t)e * kC, lane, a);
      if (net16) row_store_f16(net16 + (size_t)e * kC, lane, a);
    }
  }
}

// GatedResidual tail: y = x32 + sigmoid(a16) * r16 (blocks.py:30-31), then either
//   mode 0: x32_out = LN(y), x16_out = half(LN(y))               (next LayerNorm of Update.gru)
//   mode 1: net_out = y; delta = Wd relu(y) + bd; weight = sigmoid(Ww relu(y) + bw)  (net.py:63-67,90)
__global__ void __launch_bounds__(256)
gated_tail_kernel(const float* __restrict__ x32, const __half* __restrict__ a16,
                  const __half* __restrict__ r16, int E, int mode, const float* __restrict__ gamma,
                  const float* __restrict__ beta, float* __restrict__ out32,
                  __half* __restrict__ out16, const float* __restrict__ Wd,
                  const float* __restrict__ bd, const float* __restrict__ Ww,
                  const float* __restrict__ bw, float* __restrict__ delta,
                  float* __restrict__ weight) {
  const int lane = threadIdx.x & 31;
  RVO_ROW_LOOP(e, E) {
    Row x, a, r;
    row_load_f32(x32 + (size_t)e * kC, lane, x);
    row_load_f16(a16 + (size_t)e * kC, lane, a);
    row_load_f16(r16 + (size_t)e * kC, lane, r);
#pragma unroll
    for (int j = 0; j < kChunks; j++)
#pragma unroll
      for (int q = 0; q < 4; q++) {
        // the reference multiplies two fp16 tensors (gate * res) before the fp32 add
        const float gate = __half2float(__float2half_rn(1.0f / (1.0f + __expf(-a.v[j][q]))));
        const float prod = __half2float(__float2half_rn(gate * r.v[j][q]));
        x.v[j][q] += prod;
      }
    if (mode == 0) {
      row_layer_norm(x, gamma, beta, lane, 1e-3f);
      row_store_f32(out32 + (size_t)e * kC, lane, x);
      row_store_f16(out16 + (size_t)e * kC, lane, x);
    } else {
      row_store_f32(out32 + (size_t)e * kC, lane, x);
      float acc[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
      for (int j = 0; j < kChunks; j++) {
#pragma unroll
        for (int q = 0; q < 4; q++) {
          // heads run in fp16 under autocast: relu(net) is rounded to fp16 before the Linear
          const float v = __half2float(__float2half_rn(fmaxf(x.v[j][q], 0.f)));
          const int c = 128 * j + 4 * lane + q;
          acc[0] += v * Wd[c];
          acc[1] += v * Wd[kC + c];
          acc[2] += v * Ww[c];
          acc[3] += v * Ww[kC + c];
        }
      }
#pragma unroll
      for (int q = 0; q < 4; q++) acc[q] = warp_sum_f(acc[q]);
      if (lane == 0) {
        const float d0 = __half2float(__float2half_rn(acc[0] + bd[0]));
        const float d1 = __half2float(__float2half_rn(acc[1] + bd[1]));
        const float w0 = __half2float(__float2half_rn(acc[2] + bw[0]));
        const float w1 = __half2float(__float2half_rn(acc[3] + bw[1]));
        delta[(size_t)e * 2 + 0] = d0;
        delta[(size_t)e * 2 + 1] = d1;
        weight[(size_t)e * 2 + 0] = __half2float(__float2half_rn(1.0f / (1.0f + __expf(-w0))));
        weight[(size_t)e * 2 + 1] = __half2float(__float2half_rn(1.0f / (1.0f + __expf(-w1))));
      }
    }
  }
}

// softagg on a fused [E, 2C] buffer: columns [0,C) = f(x), [C,2C) = g(x).  One CTA per group, one
// thread per channel PAIR (half2 loads), the group's edge ids staged in shared memory and the edge
// loop unrolled by 4 so that eight independent loads are in flight per thread (the serial
// running-max recurrence is cheap; the L2 latency of the row gathers is what has to be hidden).
constexpr int kSaThreads = kC / 2;   // 192
constexpr int kSaStage = 128;

__global__ void __launch_bounds__(kSaThreads)
softagg_fg_kernel(const int32_t* __restrict__ count, const int32_t* __restrict__ perm,
                  const int32_t* __restrict__ seg_start, const __half* __restrict__ fg, int cap,
                  __half* __restrict__ y) {
  __shared__ int32_t eid[kSaStage];
  int U = count[0];
  if (U > cap) U = cap;
  const int t = threadIdx.x;
  for (int grp = blockIdx.x; grp < cap; grp += gridDim.x) {
    __half2* yo = reinterpret_cast<__half2*>(y + (size_t)grp * kC) + t;
    if (grp >= U) {   // rows past the last group stay zero (they feed a dense GEMM)
      *yo = __floats2half2_rn(0.f, 0.f);
      continue;
    }
    const int s0 = seg_start[grp], s1 = seg_start[grp + 1];
    float m0 = -INFINITY, m1 = -INFINITY, den0 = 0.f, den1 = 0.f, num0 = 0.f, num1 = 0.f;
    for (int sb = s0; sb < s1; sb += kSaStage) {
      const int n = min(kSaStage, s1 - sb);
      __syncthreads();
      if (t < n) eid[t] = perm[sb + t];
      __syncthreads();
      for (int q = 0; q < n; q += 4) {
        float2 fv[4], gv[4];
#pragma unroll
        for (int u = 0; u < 4; u++) {
          const int qq = min(q + u, n - 1);
          const __half2* row = reinterpret_cast<const __half2*>(fg + (size_t)eid[qq] * (2 * kC));
          fv[u] = __half22float2(row[t]);
          gv[u] = __half22float2(row[kSaThreads + t]);
        }
#pragma unroll
        for (int u = 0; u < 4; u++) {
          if (q + u < n) {
            float mn = fmaxf(m0, gv[u].x);
            float sc = __expf(m0 - mn), w = __expf(gv[u].x - mn);
            den0 = den0 * sc + w; num0 = num0 * sc + w * fv[u].x; m0 = mn;
            mn = fmaxf(m1, gv[u].y);
            sc = __expf(m1 - mn); w = __expf(gv[u].y - mn);
            den1 = den1 * sc + w; num1 = num1 * sc + w * fv[u].y; m1 = mn;
          }
        }
      }
    }
    *yo = (s1 > s0) ? __floats2half2_rn(num0 / den0, num1 / den1) : __floats2half2_rn(0.f, 0.f);
  }
}

// out16[e] = idx[e] >= 0 ? half(src32[idx[e]]) : 0, one warp per row (C = 384)
__global__ void __launch_bounds__(256)
gather_rows384_kernel(const float* __restrict__ src, const int64_t* __restrict__ idx, int E,
                      __half* __restrict__ out) {
  const int lane = threadIdx.x & 31;
  RVO_ROW_LOOP(e, E) {
    const int64_t j = idx[e];
    Row r;
    if (j >= 0) {
      row_load_f32(src + (size_t)j * kC, lane, r);
    } else {
#pragma unroll
      for (int c = 0; c < kChunks; c++)
#pragma unroll
        for (int q = 0; q < 4; q++) r.v[c][q] = 0.f;
    }
    row_store_f16(out + (size_t)e * kC, lane, r);
  }
}

// CTAs of 256 threads per SM for the row kernels.  8 fill every thread slot of the SM; in the two-stream frame
// (encoder of frame t+1 beside the update of frame t) that keeps the encoder's latency-bound CTAs out until the
// row kernel has drained.  RVO_ROW_CTAS_PER_SM (debug builds) sweeps it.
static inline int row_ctas_per_sm() {
#ifdef RVO_DEBUG
  static const int v = getenv("RVO_ROW_CTAS_PER_SM") ? atoi(getenv("RVO_ROW_CTAS_PER_SM")) : 8;
  return v < 1 ? 1 : (v > 8 ? 8 : v);
#else
  return 8;
#endif
}

static inline int row_grid(int E) {
  const int64_t b = ((int64_t)E * 32 + 255) / 256;
  const int64_t cap = (int64_t)sm_budget() * row_ctas_per_sm();
  return (int)(b < 1 ? 1 : (b > cap ? cap : b));
}

}  // namespace rvo

extern "C" int rvo_up_ln_relu(const void* x16, const float* gamma, const float* beta, int E, int C,
                              void* y16, void* stream) {
  RVO_CHECK_ARG(C == kC, "rvo_up_ln_relu: C=%d (384 expected)", C);
  if (E <= 0) return RVO_OK;
  RVO_CHECK_ARG(x16 && gamma && beta && y16, "rvo_up_ln_relu: null pointer");
  ln_relu_kernel<<<row_grid(E), 256, 0, (cudaStream_t)stream>>>((const __half*)x16, gamma, beta, E,
                                                                (__half*)y16);
  RVO_LAUNCH_CHECK("ln_relu_kernel");
  return RVO_OK;
}

extern "C" int rvo_up_add3_ln(const float* net_in, const void* imap16, const int64_t* idx,
                              int64_t mod, const void* h16, const float* gamma, const float* beta,
                              int E, int C, float* net_out, void* net16_out, void* stream) {
  RVO_CHECK_ARG(C == kC, "rvo_up_add3_ln: C=%d (384 expected)", C);
  if (E <= 0) return RVO_OK;
  RVO_CHECK_ARG(net_in && imap16 && idx && h16 && gamma && beta && net_out, "rvo_up_add3_ln: null pointer");
  add3_ln_kernel<<<row_grid(E), 256, 0, (cudaStream_t)stream>>>(
      net_in, (const __half*)imap16, idx, mod, (const __half*)h16, gamma, beta, E, net_out, (__half*)net16_out);
  RVO_LAUNCH_CHECK("add3_ln_kernel");
  return RVO_OK;
}

extern "C" int rvo_up_add_cast(float* net, const void* t16, int E, int C, void* net16, void* stream) {
  RVO_CHECK_ARG(C == kC, "rvo_up_add_cast: C=%d (384 expected)", C);
  if (E <= 0) return RVO_OK;
  RVO_CHECK_ARG(net && t16, "rvo_up_add_cast: null pointer");
  add_cast_kernel<<<row_grid(E), 256, 0, (cudaStream_t)stream>>>(net, (const __half*)t16, E,
                                                                 (__half*)net16);
  RVO_LAUNCH_CHECK("add_cast_kernel");
  return RVO_OK;
}

extern "C" int rvo_up_softagg_fg(const void* fg16, const void* plan, int E, int C, int64_t max_groups,
                                 void* y16, void* stream) {
  RVO_CHECK_ARG(E >= 0 && C >= 1, "rvo_up_softagg_fg: E=%d C=%d", E, C);
  if (E == 0) return RVO_OK;
  RVO_CHECK_ARG(fg16 && plan && y16, "rvo_up_softagg_fg: null pointer");
  const int32_t *count, *perm, *seg_start;
  int rc = rvo_plan_groups(plan, E, &count, &perm, nullptr, &seg_start, nullptr);
  if (rc != RVO_OK) return rc;
  const int cap = (int)((max_groups > 0 && max_groups < E) ? max_groups : E);
  int grid = cap < sm_budget() * 2 * row_ctas_per_sm() ? cap : sm_budget() * 2 * row_ctas_per_sm();
  if (grid < 1) grid = 1;
  RVO_CHECK_ARG(C == kC, "rvo_up_softagg_fg: C=%d (384 expected)", C);
  softagg_fg_kernel<<<grid, kSaThreads, 0, (cudaStream_t)stream>>>(count, perm, seg_start,
                                                                   (const __half*)fg16, cap, (__half*)y16);
  RVO_LAUNCH_CHECK("softagg_fg_kernel");
  return RVO_OK;
}

extern "C" int rvo_up_expand_add_ln(const void* hy16, const void* plan, int E, int C, float* net,
                                    void* net16, const float* gamma, const float* beta, float* x32,
                                    void* x16, void* stream) {
  RVO_CHECK_ARG(C == kC, "rvo_up_expand_add_ln: C=%d (384 expected)", C);
  if (E <= 0) return RVO_OK;
  RVO_CHECK_ARG(hy16 && plan && net, "rvo_up_expand_add_ln: null pointer");
  RVO_CHECK_ARG(!x32 || (gamma && beta && x16), "rvo_up_expand_add_ln: LayerNorm outputs need gamma/beta/x16");
  const int32_t *perm, *seg_of;
  int rc = rvo_plan_groups(plan, E, nullptr, &perm, &seg_of, nullptr, nullptr);
  if (rc != RVO_OK) return rc;
  expand_add_ln_kernel<<<row_grid(E), 256, 0, (cudaStream_t)stream>>>(
      perm, seg_of, (const __half*)hy16, E, net, (__half*)net16, gamma, beta, x32, (__half*)x16);
  RVO_LAUNCH_CHECK("expand_add_ln_kernel");
  return RVO_OK;
}

extern "C" int rvo_up_gated_tail(const float* x32, const void* a16, const void* r16, int E, int C,
                                 int mode, const float* gamma, const float* beta, float* out32,
                                 void* out16, const float* Wd, const float* bd, const float* Ww,
                                 const float* bw, float* delta, float* weight, void* stream) {
  RVO_CHECK_ARG(C == kC, "rvo_up_gated_tail: C=%d (384 expected)", C);
  if (E <= 0) return RVO_OK;
  RVO_CHECK_ARG(x32 && a16 && r16 && out32, "rvo_up_gated_tail: null pointer");
  if (mode == 0) RVO_CHECK_ARG(gamma && beta && out16, "rvo_up_gated_tail: mode 0 needs gamma/beta/out16");
  else RVO_CHECK_ARG(Wd && bd && Ww && bw && delta && weight, "rvo_up_gated_tail: mode 1 needs the heads");
  gated_tail_kernel<<<row_grid(E), 256, 0, (cudaStream_t)stream>>>(
      x32, (const __half*)a16, (const __half*)r16, E, mode, gamma, beta, out32, (__half*)out16, Wd, bd,
      Ww, bw, delta, weight);
  RVO_LAUNCH_CHECK("gated_tail_kernel");
  return RVO_OK;
}
