// fastba.cu — patch-graph plan, neighbors and the Gauss-Newton / Schur bundle adjustment (sm_100a).
//
// Reference behaviour: ramp/fastba/ba.cpp:59-97 (neighbors), ramp/fastba/ba_cuda.cu:232-376
// (residuals + Hessian), :433-582 (Schur complement, Cholesky, retractions).
//
// Design (DESIGN.md "fastba"):
//   plan      edges are radix-sorted ONCE by (kk, jj, edge id); every patch then owns one contiguous
//             segment of the sorted list.  The same plan answers `neighbors` (previous / next entry
//             of the segment) and gives the compact patch numbering the reference gets from
//             torch::_unique (ba_cuda.cu:447-449) — no host round trip, no device sync.
//   assemble  ONE WARP PER PATCH walks its segment: every lane linearises one edge in registers, the
//             patch's C, u and dense E-row (6N floats) are reduced inside the warp, the pose blocks
//             go into a per-CTA shared-memory copy of the (upper-triangular) reduced camera system,
//             and the Schur product  S -= Q_k E_k E_k^T,  y -= Q_k u_k E_k  is applied from shared
//             memory before the CTA flushes once to global memory.  The reference instead issues 342
//             global atomicAdds per edge onto a 6Nx6N matrix and materialises the dense 6N x M E.
//   solve     one CTA: damping, in-shared-memory Cholesky of [S | y] (y carried as an extra row so
//             the forward substitution is free), back substitution, pose retraction.
//   depth     one warp per patch: dZ = Q (u - E_k . dX), depth retraction with the reference clamps.
// The split between `assemble` and `solve` is where a patch graph sharded by source frame
// all-reduces [S | y] (NCCL) — everything before it is local to the shard, everything after it is
// replicated.
#include <cub/cub.cuh>

#include "common.cuh"

namespace rvo {

// ------------------------------------------------------------------ plan ----

struct PlanView {
  int32_t* count;      // [4]   count[0] = number of distinct patches U
  int32_t* perm;       // [E]   edge ids sorted by (kk, jj, e)
  int32_t* seg_of;     // [E]   sorted position -> compact patch id (the reference's `ku`)
  int32_t* seg_start;  // [E+1] compact patch id -> first sorted position; seg_start[U] = E
  int64_t* kx;         // [E]   compact patch id -> patch id (sorted unique kk)
  uint64_t* keys_a;    // [E]   sort input; afterwards (as int32[E]) edge id -> compact patch id
  uint64_t* keys_b;    // [E]
  int32_t* vals_a;     // [E]   iota, later the head flags
  void* cub_tmp;
  size_t cub_bytes;
  size_t total;
};

static inline size_t align_up(size_t v, size_t a = 256) { return (v + a - 1) / a * a; }

static size_t cub_temp_bytes(int E) {
  size_t a = 0, b = 0;
  cub::DeviceRadixSort::SortPairs(nullptr, a, (const uint64_t*)nullptr, (uint64_t*)nullptr,
                                  (const int32_t*)nullptr, (int32_t*)nullptr, E, 0, 64);
  cub::DeviceScan::InclusiveSum(nullptr, b, (const int32_t*)nullptr, (int32_t*)nullptr, E);
  return a > b ? a : b;
}

static PlanView plan_layout(void* base, int E) {
  PlanView p;
  char* c = reinterpret_cast<char*>(base);
  size_t off = 0;
  const size_t n = (size_t)(E > 0 ? E : 1);
  auto take = [&](size_t bytes) { char* r = c + off; off += align_up(bytes); return r; };
  p.count = (int32_t*)take(4 * sizeof(int32_t));
  p.perm = (int32_t*)take(n * sizeof(int32_t));
  p.seg_of = (int32_t*)take(n * sizeof(int32_t));
  p.seg_start = (int32_t*)take((n + 1) * sizeof(int32_t));
  p.kx = (int64_t*)take(n * sizeof(int64_t));
  p.keys_a = (uint64_t*)take(n * sizeof(uint64_t));
  p.keys_b = (uint64_t*)take(n * sizeof(uint64_t));
  p.vals_a = (int32_t*)take(n * sizeof(int32_t));
  p.cub_bytes = cub_temp_bytes((int)n);
  p.cub_tmp = take(p.cub_bytes);
  p.total = off;
  return p;
}

__global__ void __launch_bounds__(256)
plan_keys_kernel(const int64_t* __restrict__ kk, const int64_t* __restrict__ jj, int E, int jbits,
                 uint64_t* __restrict__ keys, int32_t* __restrict__ vals) {
  const uint64_t jmask = (1ull << jbits) - 1ull;
  for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < E; e += gridDim.x * blockDim.x) {
    keys[e] = ((uint64_t)kk[e] << jbits) | ((uint64_t)jj[e] & jmask);
    vals[e] = e;
  }
}

__global__ void __launch_bounds__(256)
plan_heads_kernel(const uint64_t* __restrict__ keys, int E, int jbits, int32_t* __restrict__ flags) {
  for (int s = blockIdx.x * blockDim.x + threadIdx.x; s < E; s += gridDim.x * blockDim.x)
    flags[s] = (s == 0 || (keys[s] >> jbits) != (keys[s - 1] >> jbits)) ? 1 : 0;
}

// seg_of holds the inclusive scan of the head flags on entry, the 0-based segment id on exit.
__global__ void __launch_bounds__(256)
plan_segments_kernel(const uint64_t* __restrict__ keys, const int32_t* __restrict__ flags, int E,
                     int jbits, int32_t* __restrict__ seg_of, int32_t* __restrict__ seg_start,
                     int64_t* __restrict__ kx, int32_t* __restrict__ count,
                     const int32_t* __restrict__ perm, int32_t* __restrict__ grp_of_edge) {
  for (int s = blockIdx.x * blockDim.x + threadIdx.x; s < E; s += gridDim.x * blockDim.x) {
    const int seg = seg_of[s] - 1;
    seg_of[s] = seg;
    grp_of_edge[perm[s]] = seg;
    if (flags[s]) {
      seg_start[seg] = s;
      kx[seg] = (int64_t)(keys[s] >> jbits);
    }
    if (s == E - 1) {
      seg_start[seg + 1] = E;
      count[0] = seg + 1;
    }
  }
}

__global__ void __launch_bounds__(256)
neighbors_kernel(const int32_t* __restrict__ perm, const int32_t* __restrict__ seg_of, int E,
                 int64_t* __restrict__ ix, int64_t* __restrict__ jx) {
  for (int s = blockIdx.x * blockDim.x + threadIdx.x; s < E; s += gridDim.x * blockDim.x) {
    const int seg = seg_of[s];
    const int e = perm[s];
    ix[e] = (s > 0 && seg_of[s - 1] == seg) ? (int64_t)perm[s - 1] : -1;
    jx[e] = (s + 1 < E && seg_of[s + 1] == seg) ? (int64_t)perm[s + 1] : -1;
  }
}

static inline int ceil_log2(int64_t v) {
  int b = 0;
  while (b < 62 && ((int64_t)1 << b) < v) b++;
  return b;
}

static inline int grid1d(int64_t n, int per_sm = 8) {
  int64_t b = (n + 255) / 256;
  const int64_t cap = (int64_t)sm_budget() * per_sm;
  return (int)(b < 1 ? 1 : (b > cap ? cap : b));
}

static int build_plan(const int64_t* kk, const int64_t* jj, int E, int64_t kmax, int64_t jmax,
                      void* plan, int64_t plan_bytes, cudaStream_t st, PlanView* out) {
  RVO_CHECK_ARG(E >= 0, "rvo_graph_plan: E=%d", E);
  RVO_CHECK_ARG(plan, "rvo_graph_plan: null plan buffer");
  PlanView p = plan_layout(plan, E);
  RVO_CHECK_ARG((int64_t)p.total <= plan_bytes, "rvo_graph_plan: plan buffer %lld < %lld bytes",
                (long long)plan_bytes, (long long)p.total);
  if (out) *out = p;
  if (E == 0) {
    RVO_CUDA(cudaMemsetAsync(p.count, 0, 4 * sizeof(int32_t), st));
    return RVO_OK;
  }
  RVO_CHECK_ARG(kk && jj, "rvo_graph_plan: null index array");
  const int jbits = jmax > 0 ? (ceil_log2(jmax) < 1 ? 1 : ceil_log2(jmax)) : 21;
  const int kbits = kmax > 0 ? (ceil_log2(kmax) < 1 ? 1 : ceil_log2(kmax)) : 42;
  RVO_CHECK_ARG(jbits + kbits <= 63, "rvo_graph_plan: kmax/jmax too large");
  const int g = grid1d(E);
  plan_keys_kernel<<<g, 256, 0, st>>>(kk, jj, E, jbits, p.keys_a, p.vals_a);
  RVO_LAUNCH_CHECK("plan_keys_kernel");
  size_t tb = p.cub_bytes;
  RVO_CUDA(cub::DeviceRadixSort::SortPairs(p.cub_tmp, tb, p.keys_a, p.keys_b, p.vals_a, p.perm, E,
                                           0, jbits + kbits, st));
  plan_heads_kernel<<<g, 256, 0, st>>>(p.keys_b, E, jbits, p.vals_a);
  RVO_LAUNCH_CHECK("plan_heads_kernel");
  tb = p.cub_bytes;
  RVO_CUDA(cub::DeviceScan::InclusiveSum(p.cub_tmp, tb, p.vals_a, p.seg_of, E, st));
  // keys_a is dead once the sort has consumed it: its storage becomes the group id of every EDGE
  plan_segments_kernel<<<g, 256, 0, st>>>(p.keys_b, p.vals_a, E, jbits, p.seg_of, p.seg_start, p.kx,
                                          p.count, p.perm, reinterpret_cast<int32_t*>(p.keys_a));
  RVO_LAUNCH_CHECK("plan_segments_kernel");
  return RVO_OK;
}

// ------------------------------------------------------------------ BA: per-edge linearisation ----

struct EdgeLin {
  float Ji[2][6], Jj[2][6], Jz[2], r[2], w[2];
};

// ba_cuda.cu:265-326, same expressions (double literals included) so the gates agree exactly.
__device__ __forceinline__ void linearise_edge(const float* __restrict__ poses,
                                               const float* __restrict__ patches, float fx, float fy,
                                               float cx, float cy, int64_t i, int64_t j, int64_t k,
                                               int PP, int ctr, float tx, float ty, float wx,
                                               float wy, EdgeLin& L) {
  float ti[3], qi[4], tj[3], qj[4], tij[3], qij[4];
  load_pose(poses, i, ti, qi);
  load_pose(poses, j, tj, qj);
  rel_se3(ti, qi, tj, qj, tij, qij);
  const float* pk = patches + k * 3 * PP;
  float Xi[4], Xj[4];
  Xi[0] = (pk[ctr] - cx) / fx;
  Xi[1] = (pk[PP + ctr] - cy) / fy;
  Xi[2] = 1.0f;
  Xi[3] = pk[2 * PP + ctr];
  rot_q(qij, Xi, Xj);
  Xj[0] += Xi[3] * tij[0];
  Xj[1] += Xi[3] * tij[1];
  Xj[2] += Xi[3] * tij[2];
  const float X = Xj[0], Y = Xj[1], Z = Xj[2], W = Xi[3];
  const float d = (Z >= 0.2) ? (float)(1.0 / Z) : 0.0f;
  const float d2 = d * d;
  const float x1 = fx * (X / Z) + cx;
  const float y1 = fy * (Y / Z) + cy;
  const float rx = tx - x1, ry = ty - y1;
  const bool in_bounds = (sqrtf(rx * rx + ry * ry) < 128) && (Z > 0.2) && (x1 > -64) &&
                         (y1 > -64) && (x1 < 2 * cx + 64) && (y1 < 2 * cy + 64);
  const float mask = in_bounds ? 1.0f : 0.0f;
  L.r[0] = rx; L.r[1] = ry;
  L.w[0] = mask * wx; L.w[1] = mask * wy;
  L.Jz[0] = fx * (tij[0] * d - tij[2] * (X * d2));
  L.Jz[1] = fy * (tij[1] * d - tij[2] * (Y * d2));
  L.Jj[0][0] = fx * W * d;  L.Jj[0][1] = 0.0f;         L.Jj[0][2] = fx * -X * W * d2;
  L.Jj[0][3] = fx * -X * Y * d2; L.Jj[0][4] = fx * (1 + X * X * d2); L.Jj[0][5] = fx * -Y * d;
  L.Jj[1][0] = 0.0f;        L.Jj[1][1] = fy * W * d;   L.Jj[1][2] = fy * -Y * W * d2;
  L.Jj[1][3] = fy * (-1 - Y * Y * d2); L.Jj[1][4] = fy * (X * Y * d2); L.Jj[1][5] = fy * X * d;
  adjT_se3(tij, qij, L.Jj[0], L.Ji[0]);
  adjT_se3(tij, qij, L.Jj[1], L.Ji[1]);
}

// ------------------------------------------------------------------ BA: assemble ----

constexpr int kBaWarps = 8;
constexpr int kBaThreads = kBaWarps * 32;

// optional fused target formation (rvo_ba_forward_fused): coords [E,2,P,P] of the reprojected patches
struct BaFuse {
  const float* coords;
  float ht, wd;
  float* weight_out;    // filtered confidences [E,2] (Ramp_vo.last_weight), may be null
};

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// upper-triangle index helper: block (bi,bj) with bi<=bj, entry (a,b)
#define S_AT(r, c) S[(size_t)(r) * ld + (c)]

template <bool SMEM_S>
__global__ void __launch_bounds__(kBaThreads)
ba_assemble_kernel(const int32_t* __restrict__ count, const int32_t* __restrict__ perm,
                   const int32_t* __restrict__ seg_start, const int64_t* __restrict__ kx,
                   const float* __restrict__ poses, const float* __restrict__ patches,
                   const float* __restrict__ intr, const float* __restrict__ target,
                   const float* __restrict__ weight, const float* __restrict__ lmbda,
                   const int64_t* __restrict__ ii, const int64_t* __restrict__ jj, int P, int t0_arg,
                   const int32_t* __restrict__ t0_dev, int N, int cap, float* __restrict__ Sy_g,
                   float* __restrict__ Qg, float* __restrict__ ug, float* __restrict__ Eg, const BaFuse fz) {
  extern __shared__ float sm[];
  const int t0 = t0_dev ? t0_dev[0] : t0_arg;   // device-side window start keeps CUDA graphs replayable
  const int n6 = 6 * N, ld = n6 + 1;
  float* S = SMEM_S ? sm : Sy_g;                       // [n6][ld] upper triangle + y column
  float* Ew = sm + (SMEM_S ? (size_t)n6 * ld : 0);     // [W][n6]  E_k rows of this round
  float* EQw = Ew + kBaWarps * n6;                     // [W][n6]  Q_k E_k
  float* Qw = EQw + kBaWarps * n6;                     // [W]      Q_k u_k per warp
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int PP = P * P, ctr = (P >= 2) ? (P + 1) : 0;  // reference reads pixel [1][1]
  const float fx = intr[0], fy = intr[1], cx = intr[2], cy = intr[3];
  const float lam = lmbda[0];
  int U = count[0];
  if (U > cap) U = cap;

  if (SMEM_S) {
    for (int t = threadIdx.x; t < n6 * ld; t += blockDim.x) S[t] = 0.0f;
  }
  __syncthreads();

  for (int base = blockIdx.x * kBaWarps; base < U; base += gridDim.x * kBaWarps) {
    const int p = base + warp;
    float* E_k = Ew + warp * n6;
    for (int t = lane; t < n6; t += 32) E_k[t] = 0.0f;
    __syncwarp();
    float Ck = 0.0f, uk = 0.0f;
    if (p < U) {
      const int s0 = seg_start[p], s1 = seg_start[p + 1];
      const int64_t k = kx[p];
      for (int sb = s0; sb < s1; sb += 32) {
        const int s = sb + lane;
        const bool act = s < s1;
        int ip = -1, jp = -1;
        EdgeLin L;
        if (act) {
          const int e = perm[s];
          const int64_t i = ii[e], j = jj[e];
          float2 tg = reinterpret_cast<const float2*>(target)[e];
          float2 wg = reinterpret_cast<const float2*>(weight)[e];
          if (fz.coords) {
            // `target` holds the update operator's delta: target = reprojected patch centre + delta
            // (Ramp_vo.py:289) and filter_features (utils.py:557-570) zeroes the confidence of targets outside
            // [0, wd] x [0, ht] — both folded into this load instead of ~10 elementwise launches
            const float* c = fz.coords + (size_t)e * 2 * PP;
            tg.x += c[ctr];
            tg.y += c[PP + ctr];
            const bool bad = (tg.x < 0.0f) | (tg.x > fz.wd) | (tg.y < 0.0f) | (tg.y > fz.ht);
            if (bad) { wg.x = 0.0f; wg.y = 0.0f; }
            if (fz.weight_out) reinterpret_cast<float2*>(fz.weight_out)[e] = wg;
          }
          linearise_edge(poses, patches, fx, fy, cx, cy, i, j, k, PP, ctr, tg.x, tg.y, wg.x, wg.y, L);
          ip = (int)(i - t0);
          jp = (int)(j - t0);
          if (ip >= N) ip = -1;  // outside the free window: treated as fixed
          if (jp >= N) jp = -1;
        } else {
#pragma unroll
          for (int r = 0; r < 2; r++) {
            L.w[r] = 0.f; L.r[r] = 0.f; L.Jz[r] = 0.f;
#pragma unroll
            for (int a = 0; a < 6; a++) { L.Ji[r][a] = 0.f; L.Jj[r][a] = 0.f; }
          }
        }
        const bool fi = act && ip >= 0, fj = act && jp >= 0;
        // patch block (ba_cuda.cu:372-373)
#pragma unroll
        for (int r = 0; r < 2; r++) {
          Ck += L.w[r] * L.Jz[r] * L.Jz[r];
          uk += L.w[r] * L.r[r] * L.Jz[r];
        }
        // (j,j) block, v_j, E_j: the lanes of a segment have distinct j
        if (fj) {
#pragma unroll
          for (int a = 0; a < 6; a++) {
            const float wa0 = L.w[0] * L.Jj[0][a], wa1 = L.w[1] * L.Jj[1][a];
#pragma unroll
            for (int b = a; b < 6; b++)
              atomicAdd(&S_AT(6 * jp + a, 6 * jp + b), wa0 * L.Jj[0][b] + wa1 * L.Jj[1][b]);
            atomicAdd(&S_AT(6 * jp + a, n6),
                      L.w[0] * L.r[0] * L.Jj[0][a] + L.w[1] * L.r[1] * L.Jj[1][a]);
            atomicAdd(&E_k[6 * jp + a], L.w[0] * L.Jz[0] * L.Jj[0][a] + L.w[1] * L.Jz[1] * L.Jj[1][a]);
          }
        }
        // (i,j) block: stored in the upper triangle only
        if (fi && fj) {
          if (ip < jp) {
#pragma unroll
            for (int a = 0; a < 6; a++)
#pragma unroll
              for (int b = 0; b < 6; b++)
                atomicAdd(&S_AT(6 * ip + a, 6 * jp + b),
                          -L.w[0] * L.Ji[0][a] * L.Jj[0][b] - L.w[1] * L.Ji[1][a] * L.Jj[1][b]);
          } else if (jp < ip) {
#pragma unroll
            for (int a = 0; a < 6; a++)
#pragma unroll
              for (int b = 0; b < 6; b++)
                atomicAdd(&S_AT(6 * jp + a, 6 * ip + b),
                          -L.w[0] * L.Jj[0][a] * L.Ji[0][b] - L.w[1] * L.Jj[1][a] * L.Ji[1][b]);
          } else {  // self edge: both cross terms land on the diagonal block
#pragma unroll
            for (int a = 0; a < 6; a++)
#pragma unroll
              for (int b = a; b < 6; b++) {
                float v = -L.w[0] * L.Ji[0][a] * L.Jj[0][b] - L.w[1] * L.Ji[1][a] * L.Jj[1][b];
                v += -L.w[0] * L.Jj[0][a] * L.Ji[0][b] - L.w[1] * L.Jj[1][a] * L.Ji[1][b];
                atomicAdd(&S_AT(6 * ip + a, 6 * ip + b), v);
              }
          }
        }
        // (i,i) block, v_i, E_i: normally every edge of a patch has the same source frame, so
        // reduce across the warp first (33 values) and let lanes 0..32 issue one add each.
        float aii[21], vi[6], ei[6];
        {
          int q = 0;
#pragma unroll
          for (int a = 0; a < 6; a++) {
            const float wa0 = fi ? L.w[0] * L.Ji[0][a] : 0.f, wa1 = fi ? L.w[1] * L.Ji[1][a] : 0.f;
#pragma unroll
            for (int b = a; b < 6; b++) aii[q++] = wa0 * L.Ji[0][b] + wa1 * L.Ji[1][b];
            vi[a] = fi ? -(L.w[0] * L.r[0] * L.Ji[0][a] + L.w[1] * L.r[1] * L.Ji[1][a]) : 0.f;
            ei[a] = fi ? -(L.w[0] * L.Jz[0] * L.Ji[0][a] + L.w[1] * L.Jz[1] * L.Ji[1][a]) : 0.f;
          }
        }
        const int ip0 = __shfl_sync(0xffffffffu, ip, 0);
        const bool uniform = __all_sync(0xffffffffu, !act || ip == ip0);
        if (uniform) {
          if (ip0 >= 0) {  // warp-uniform branch
            float mine = 0.f, mine2 = 0.f;  // lane l keeps value l (and l+32 for l == 0)
            int q = 0;
#pragma unroll
            for (int a = 0; a < 6; a++)
#pragma unroll
              for (int b = a; b < 6; b++) {
                const float t = warp_sum(aii[q]);
                if (lane == q) mine = t;
                q++;
              }
#pragma unroll
            for (int a = 0; a < 6; a++) {
              const float t = warp_sum(vi[a]);
              if (lane == 21 + a) mine = t;
            }
#pragma unroll
            for (int a = 0; a < 5; a++) {
              const float t = warp_sum(ei[a]);
              if (lane == 27 + a) mine = t;
            }
            mine2 = warp_sum(ei[5]);
            if (lane < 21) {
              // invert q -> (a,b) of the upper triangle
              int a = 0, rem = lane;
              while (rem >= 6 - a) { rem -= 6 - a; a++; }
              atomicAdd(&S_AT(6 * ip0 + a, 6 * ip0 + a + rem), mine);
            } else if (lane < 27) {
              atomicAdd(&S_AT(6 * ip0 + (lane - 21), n6), mine);
            } else {
              atomicAdd(&E_k[6 * ip0 + (lane - 27)], mine);
            }
            if (lane == 0) atomicAdd(&E_k[6 * ip0 + 5], mine2);
          }
        } else if (fi) {
          int q = 0;
#pragma unroll
          for (int a = 0; a < 6; a++) {
#pragma unroll
            for (int b = a; b < 6; b++) atomicAdd(&S_AT(6 * ip + a, 6 * ip + b), aii[q++]);
            atomicAdd(&S_AT(6 * ip + a, n6), vi[a]);
            atomicAdd(&E_k[6 * ip + a], ei[a]);
          }
        }
      }
      Ck = warp_sum(Ck);
      uk = warp_sum(uk);
    }
    __syncwarp();
    // Q = 1/(C + lambda) (ba_cuda.cu:519); keep E_k, Q_k, u_k for the back substitution
    const float Qk = (p < U) ? 1.0f / (Ck + lam) : 0.0f;
    if (p < U) {
      if (lane == 0) { Qg[p] = Qk; ug[p] = uk; }
      for (int t = lane; t < n6; t += 32) {
        const float ev = E_k[t];
        Eg[(size_t)p * n6 + t] = ev;
        EQw[warp * n6 + t] = Qk * ev;
      }
    } else {
      for (int t = lane; t < n6; t += 32) EQw[warp * n6 + t] = 0.0f;
    }
    if (lane == 0) Qw[warp] = (p < U) ? uk : 0.0f;
    __syncthreads();
    // Schur product for the patches of this round: S -= (Q E)(E)^T, y -= (Q E) u  (ba_cuda.cu:555-556)
    if (n6 > 0) {
      for (int t = threadIdx.x; t < n6 * ld; t += blockDim.x) {
        const int r = t / ld, c = t - r * ld;
        if (c < r) continue;
        float acc = 0.0f;
        if (c < n6) {
#pragma unroll
          for (int w = 0; w < kBaWarps; w++) acc += EQw[w * n6 + r] * Ew[w * n6 + c];
        } else {
#pragma unroll
          for (int w = 0; w < kBaWarps; w++) acc += EQw[w * n6 + r] * Qw[w];
        }
        if (SMEM_S) S[t] -= acc;
        else if (acc != 0.0f) atomicAdd(&S[t], -acc);
      }
    }
    __syncthreads();
  }

  if (SMEM_S) {
    for (int t = threadIdx.x; t < n6 * ld; t += blockDim.x) {
      const int r = t / ld, c = t - r * ld;
      if (c < r) continue;
      const float v = S[t];
      if (v != 0.0f) atomicAdd(&Sy_g[t], v);
    }
  }
}

// mirror the upper triangle so callers (and the all-reduce) see the full symmetric [S | y]
__global__ void __launch_bounds__(256) ba_mirror_kernel(float* __restrict__ Sy, int n6) {
  const int ld = n6 + 1;
  for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < n6 * n6; t += gridDim.x * blockDim.x) {
    const int r = t / n6, c = t - r * n6;
    if (c < r) Sy[(size_t)r * ld + c] = Sy[(size_t)c * ld + r];
  }
}

// ------------------------------------------------------------------ BA: solve ----

// ba_cuda.cu:88-174 (expSO3 / expSE3 / retrSE3), same branches and constants.
__device__ __forceinline__ void retr_se3(const float* xi, float* t, float* q) {
  const float* phi = xi + 3;
  const float theta_sq = phi[0] * phi[0] + phi[1] * phi[1] + phi[2] * phi[2];
  const float theta_p4 = theta_sq * theta_sq;
  const float theta = sqrtf(theta_sq);
  float imag, real;
  if (theta_sq < 1e-8) {
    imag = (float)(0.5 - (1.0 / 48.0) * theta_sq + (1.0 / 3840.0) * theta_p4);
    real = (float)(1.0 - (1.0 / 8.0) * theta_sq + (1.0 / 384.0) * theta_p4);
  } else {
    imag = sinf(0.5f * theta) / theta;
    real = cosf(0.5f * theta);
  }
  const float dq[4] = {imag * phi[0], imag * phi[1], imag * phi[2], real};
  float dt[3] = {xi[0], xi[1], xi[2]};
  if (theta > 1e-4) {
    float tau[3] = {xi[0], xi[1], xi[2]};
    const float a = (1 - cosf(theta)) / theta_sq;
    float c1[3] = {phi[1] * tau[2] - phi[2] * tau[1], phi[2] * tau[0] - phi[0] * tau[2],
                   phi[0] * tau[1] - phi[1] * tau[0]};
    dt[0] += a * c1[0]; dt[1] += a * c1[1]; dt[2] += a * c1[2];
    const float b = (theta - sinf(theta)) / (theta * theta_sq);
    float c2[3] = {phi[1] * c1[2] - phi[2] * c1[1], phi[2] * c1[0] - phi[0] * c1[2],
                   phi[0] * c1[1] - phi[1] * c1[0]};
    dt[0] += b * c2[0]; dt[1] += b * c2[1]; dt[2] += b * c2[2];
  }
  float q1[4], t1[3];
  q1[0] = dq[3] * q[0] + dq[0] * q[3] + dq[1] * q[2] - dq[2] * q[1];
  q1[1] = dq[3] * q[1] + dq[1] * q[3] + dq[2] * q[0] - dq[0] * q[2];
  q1[2] = dq[3] * q[2] + dq[2] * q[3] + dq[0] * q[1] - dq[1] * q[0];
  q1[3] = dq[3] * q[3] - dq[0] * q[0] - dq[1] * q[1] - dq[2] * q[2];
  rot_q(dq, t, t1);
  t[0] = t1[0] + dt[0]; t[1] = t1[1] + dt[1]; t[2] = t1[2] + dt[2];
  q[0] = q1[0]; q[1] = q1[1]; q[2] = q1[2]; q[3] = q1[3];
}

// One CTA of (32 x TY) threads.  A = (n+1) x (n+1) lower-triangular workspace (shared memory, or
// global scratch when the window is too large): rows 0..n-1 hold S, row n holds y^T.
// Right-looking Cholesky with a 2-D thread mapping (x: column, y: row — no integer division in the
// trailing update); row n comes out as z = L^-1 y, then one warp solves L^T dX = z
// (ba_cuda.cu:558-562) with warp-level synchronisation only.
template <int SETS>
__global__ void __launch_bounds__(1024)
ba_solve_kernel(const float* __restrict__ Sy, int N, int t0_arg, const int32_t* __restrict__ t0_dev,
                float* __restrict__ poses, float* __restrict__ dX_g, float* __restrict__ A_g) {
  extern __shared__ float sm[];
  const int t0 = t0_dev ? t0_dev[0] : t0_arg;
  const int n = 6 * N, ld = n + 1;
  const int la = (n + 1) | 1;                              // odd row stride: conflict-free columns
  float* A = A_g ? A_g : sm;                              // [(n+1)][la]
  float* diag = A_g ? sm : sm + (size_t)(n + 1) * la;     // [n]  1 / L[j][j]
  float* x = diag + n;                                    // [n]
  const int tx = threadIdx.x, ty = threadIdx.y, TY = blockDim.y;
  const int tid = ty * 32 + tx;
  for (int r = ty; r <= n; r += TY)
    for (int c = tx; c <= r && c < n; c += 32) {
      float v;
      if (r < n) {
        v = Sy[(size_t)r * ld + c];
        if (r == c) v = v + (1e-4f * v + 1.0f);           // S += I * (1e-4 * S + 1)
      } else {
        v = Sy[(size_t)c * ld + n];
      }
      A[(size_t)r * la + c] = v;
    }
  __syncthreads();
  // Blocked by pose (panels of 6 columns): one warp factors the panel with warp-level synchronisation only, then
  // the whole CTA applies the panel's six rank-1 updates to the trailing matrix in ONE pass — 2 block barriers per
  // pose instead of 2 per column (the column-by-column version spent 120 / 360 barriers on the 60 / 180 unknowns of
  // default.yaml / precise.yaml: 46 / 269 us).  Every element still receives its updates in increasing column
  // order, so the factor is bit-identical to the right-looking algorithm.
  constexpr int PB = 6;
  for (int j0 = 0; j0 < n; j0 += PB) {
    if constexpr (SETS == 0) {
      // large windows (6N + 1 > 64 rows): the panel stays in shared memory, one warp, warp-level synchronisation only
      // (the register panel below needs 6 x ceil(rows / 32) live values per lane and measured slower at 181 rows)
      if (ty == 0) {
#pragma unroll 1
        for (int jj = 0; jj < PB; jj++) {
          const int j = j0 + jj;
          const float d = sqrtf(A[(size_t)j * la + j]);
          const float inv = 1.0f / d;
          for (int i = j + 1 + tx; i <= n; i += 32) A[(size_t)i * la + j] *= inv;
          if (tx == 0) diag[j] = inv;
          __syncwarp();
          for (int i = j + 1 + tx; i <= n; i += 32) {
            const float lij = A[(size_t)i * la + j];
            const int kmax = min(min(i, n - 1), j0 + PB - 1);
            for (int k = j + 1; k <= kmax; k++) A[(size_t)i * la + k] -= lij * A[(size_t)k * la + j];
          }
          __syncwarp();
        }
      }
    } else {
    if (ty == 0) {
        // the panel lives in REGISTERS: lane l owns rows l, l + 32, ... (SETS of them cover rows 0..n), six panel
        // columns each; pivots and the multipliers L[k][j] travel by warp shuffle.  (Through shared memory every
        // in-panel update was a dependent load -> FMA -> store round trip: ~900 cycles per column, 27 of the 46 us.)
        float a[SETS > 0 ? SETS : 1][PB];
  #pragma unroll
        for (int s = 0; s < SETS; s++) {
          const int r = tx + 32 * s;
  #pragma unroll
          for (int c = 0; c < PB; c++) a[s][c] = (r >= j0 && r <= n) ? A[(size_t)r * la + j0 + c] : 0.f;
        }
  #pragma unroll
        for (int jj = 0; jj < PB; jj++) {
          const int j = j0 + jj;                       // j < n: n is a multiple of 6
          float v = 0.f;
  #pragma unroll
          for (int s = 0; s < SETS; s++)
            if (s == (j >> 5)) v = a[s][jj];
          const float d = sqrtf(__shfl_sync(0xffffffffu, v, j & 31));
          const float inv = 1.0f / d;
  #pragma unroll
          for (int s = 0; s < SETS; s++)
            if (tx + 32 * s > j) a[s][jj] *= inv;
          if (tx == 0) diag[j] = inv;
  #pragma unroll
          for (int kk = jj + 1; kk < PB; kk++) {
            const int k = j0 + kk;
            float lk = 0.f;
  #pragma unroll
            for (int s = 0; s < SETS; s++)
              if (s == (k >> 5)) lk = a[s][jj];
            lk = __shfl_sync(0xffffffffu, lk, k & 31);
  #pragma unroll
            for (int s = 0; s < SETS; s++)
              if (tx + 32 * s >= k) a[s][kk] -= a[s][jj] * lk;
          }
        }
  #pragma unroll
        for (int s = 0; s < SETS; s++) {
          const int r = tx + 32 * s;
          if (r >= j0 && r <= n) {
  #pragma unroll
            for (int c = 0; c < PB; c++)
              if (r >= j0 + c) A[(size_t)r * la + j0 + c] = a[s][c];
          }
        }
      }
    }
    __syncthreads();
    for (int i = j0 + PB + ty; i <= n; i += TY) {
      float l[PB];
#pragma unroll
      for (int jj = 0; jj < PB; jj++) l[jj] = A[(size_t)i * la + j0 + jj];
      const int kmax = i < n - 1 ? i : n - 1;
      for (int k = j0 + PB + tx; k <= kmax; k += 32) {
        float a = A[(size_t)i * la + k];
#pragma unroll
        for (int jj = 0; jj < PB; jj++) a -= l[jj] * A[(size_t)k * la + j0 + jj];
        A[(size_t)i * la + k] = a;
      }
    }
    __syncthreads();
  }
  // back substitution on z = A[n][:], one warp
  if (ty == 0) {
    float* z = A + (size_t)n * la;
    for (int j = n - 1; j >= 0; j--) {
      const float xj = z[j] * diag[j];
      for (int i = tx; i < j; i += 32) z[i] -= A[(size_t)j * la + i] * xj;
      if (tx == 0) x[j] = xj;
      __syncwarp();
    }
  }
  __syncthreads();
  for (int t = tid; t < n; t += 32 * TY) dX_g[t] = x[t];
  // pose retraction T <- Exp(dX) T  (ba_cuda.cu:178-206)
  for (int i = tid; i < N; i += 32 * TY) {
    float* pp = poses + (size_t)(t0 + i) * 7;
    float t[3] = {pp[0], pp[1], pp[2]}, q[4] = {pp[3], pp[4], pp[5], pp[6]};
    float xi[6];
#pragma unroll
    for (int a = 0; a < 6; a++) xi[a] = x[6 * i + a];
    retr_se3(xi, t, q);
    pp[0] = t[0]; pp[1] = t[1]; pp[2] = t[2];
    pp[3] = q[0]; pp[4] = q[1]; pp[5] = q[2]; pp[6] = q[3];
  }
}

// dZ = Q (u - E^T dX)  (ba_cuda.cu:562; structure-only :521-531) and the depth retraction (:209-229)
__global__ void __launch_bounds__(256)
ba_depth_kernel(const int32_t* __restrict__ count, const int64_t* __restrict__ kx, int cap, int n6,
                const float* __restrict__ Qg, const float* __restrict__ ug,
                const float* __restrict__ Eg, const float* __restrict__ dX, int P,
                float* __restrict__ patches) {
  int U = count[0];
  if (U > cap) U = cap;
  const int lane = threadIdx.x & 31;
  const int PP = P * P;
  for (int p = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; p < U;
       p += (gridDim.x * blockDim.x) >> 5) {
    float dot = 0.0f;
    for (int t = lane; t < n6; t += 32) dot += Eg[(size_t)p * n6 + t] * dX[t];
    dot = warp_sum(dot);
    const float dZ = Qg[p] * (ug[p] - dot);
    float* pd = patches + (size_t)kx[p] * 3 * PP + 2 * PP;
    float d = pd[0];
    d = d + dZ;
    d = (d > 20) ? 1.0f : d;
    d = fmaxf(d, (float)1e-4);
    __syncwarp();
    for (int t = lane; t < PP; t += 32) pd[t] = d;
  }
}

// ------------------------------------------------------------------ BA: host orchestration ----

struct BaWs {
  PlanView plan;
  float* Sy;    // [n6][n6+1]
  float* dX;    // [n6]
  float* Qg;    // [cap]
  float* ug;    // [cap]
  float* Eg;    // [cap][n6]
  float* A;     // [(n6+1)^2] Cholesky scratch, only used when it does not fit shared memory
  size_t total;
};

constexpr size_t kSmemMax = 227 * 1024;

static inline size_t solve_smem_bytes(int n6) {
  return ((size_t)(n6 + 1) * ((n6 + 1) | 1) + 2 * (size_t)n6) * sizeof(float);
}
static inline size_t assemble_smem_bytes(int n6, bool smem_s) {
  return ((smem_s ? (size_t)n6 * (n6 + 1) : 0) + 2 * (size_t)kBaWarps * n6 + kBaWarps) *
         sizeof(float);
}

static BaWs ba_layout(void* base, int E, int64_t cap, int n_free) {
  BaWs w;
  w.plan = plan_layout(base, E);
  char* c = reinterpret_cast<char*>(base);
  size_t off = w.plan.total;
  const size_t n6 = (size_t)6 * (n_free > 0 ? n_free : 0);
  const size_t cp = (size_t)(cap > 0 ? cap : 1);
  auto take = [&](size_t bytes) { char* r = c + off; off += align_up(bytes ? bytes : 4); return r; };
  w.Sy = (float*)take(n6 * (n6 + 1) * sizeof(float));
  w.dX = (float*)take(n6 * sizeof(float));
  w.Qg = (float*)take(cp * sizeof(float));
  w.ug = (float*)take(cp * sizeof(float));
  w.Eg = (float*)take(cp * n6 * sizeof(float));
  w.A = (float*)take(solve_smem_bytes((int)n6) > kSmemMax ? (n6 + 1) * ((n6 + 1) | 1) * sizeof(float) : 0);
  w.total = off;
  return w;
}

static inline int64_t patch_cap(int E, int64_t n_patches) {
  return (n_patches > 0 && n_patches < E) ? n_patches : E;
}

static int ba_assemble(const BaWs& w, const float* poses, const float* patches,
                       const float* intrinsics, const float* target, const float* weight,
                       const float* lmbda, const int64_t* ii, const int64_t* jj, int E, int64_t cap,
                       int P, int t0, int t1, float* Sy_out, cudaStream_t st,
                       const int32_t* t0_dev = nullptr, BaFuse fz = BaFuse{nullptr, 0.f, 0.f, nullptr}) {
  const int N = t1 - t0, n6 = 6 * N;
  if (n6 > 0) RVO_CUDA(cudaMemsetAsync(Sy_out, 0, (size_t)n6 * (n6 + 1) * sizeof(float), st));
  const bool smem_s = assemble_smem_bytes(n6, true) <= kSmemMax;
  const size_t smem = assemble_smem_bytes(n6, smem_s);
  RVO_CHECK_ARG(smem <= kSmemMax, "rvo_ba: optimisation window of %d poses is too large", N);
  int64_t want = (cap + kBaWarps - 1) / kBaWarps;
  const int ctas_per_sm = smem_s ? (int)(kSmemMax / (smem + 1024) < 1 ? 1 : (kSmemMax / (smem + 1024) > 4 ? 4 : kSmemMax / (smem + 1024))) : 4;
  int grid = (int)(want < 1 ? 1 : (want > (int64_t)sm_budget() * ctas_per_sm ? (int64_t)sm_budget() * ctas_per_sm : want));
  if (smem_s) {
    RVO_CUDA(cudaFuncSetAttribute(ba_assemble_kernel<true>,
                                  cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemMax));
    ba_assemble_kernel<true><<<grid, kBaThreads, smem, st>>>(
        w.plan.count, w.plan.perm, w.plan.seg_start, w.plan.kx, poses, patches, intrinsics, target,
        weight, lmbda, ii, jj, P, t0, t0_dev, N, (int)cap, Sy_out, w.Qg, w.ug, w.Eg, fz);
  } else {
    ba_assemble_kernel<false><<<grid, kBaThreads, smem, st>>>(
        w.plan.count, w.plan.perm, w.plan.seg_start, w.plan.kx, poses, patches, intrinsics, target,
        weight, lmbda, ii, jj, P, t0, t0_dev, N, (int)cap, Sy_out, w.Qg, w.ug, w.Eg, fz);
  }
  RVO_LAUNCH_CHECK("ba_assemble_kernel");
  if (n6 > 0) {
    ba_mirror_kernel<<<grid1d((int64_t)n6 * n6), 256, 0, st>>>(Sy_out, n6);
    RVO_LAUNCH_CHECK("ba_mirror_kernel");
  }
  (void)E;
  return RVO_OK;
}

// SETS = 2: the 6-column panel of the Cholesky factorisation is register-resident (rows spread over the lanes of one
// warp, two per lane: windows of up to 10 poses); SETS = 0: the panel stays in shared memory
static int launch_ba_solve(dim3 threads, size_t smem, cudaStream_t st, const float* Sy, int N, int t0,
                           const int32_t* t0_dev, float* poses, float* dX, float* A_g) {
  const int sets = (6 * N + 1 + 31) / 32;
#define RVO_SOLVE(S)                                                                                         \
  do {                                                                                                       \
    RVO_CUDA(cudaFuncSetAttribute(ba_solve_kernel<S>, cudaFuncAttributeMaxDynamicSharedMemorySize,           \
                                  (int)kSmemMax));                                                           \
    ba_solve_kernel<S><<<1, threads, smem, st>>>(Sy, N, t0, t0_dev, poses, dX, A_g);                         \
  } while (0)
  if (sets <= 2) RVO_SOLVE(2);          // up to 10 poses (default.yaml): register-resident panel
  else RVO_SOLVE(0);                    // shared-memory panel
#undef RVO_SOLVE
  RVO_LAUNCH_CHECK("ba_solve_kernel");
  return RVO_OK;
}

static int ba_solve(const BaWs& w, float* poses, float* patches, const float* Sy, int64_t cap, int P,
                    int t0, int t1, cudaStream_t st, const int32_t* t0_dev = nullptr) {
  const int N = t1 - t0, n6 = 6 * N;
  if (N > 0) {
    const bool in_smem = solve_smem_bytes(n6) <= kSmemMax;
    const size_t smem = in_smem ? solve_smem_bytes(n6) : 2 * (size_t)n6 * sizeof(float);
    // (a fully unrolled single-warp register Cholesky of the WHOLE matrix was tried for 6N <= 64: 60 us vs 45 us —
    // ~160 KB of straight-line code bound by instruction fetch; only the 6-column panel is register-resident now)
    const dim3 threads(32, n6 <= 24 ? 8 : 32);
    int rc = launch_ba_solve(threads, smem, st, Sy, N, t0, t0_dev, poses, w.dX, in_smem ? nullptr : w.A);
    if (rc != RVO_OK) return rc;
  }
  int64_t warps = cap;
  int grid = (int)((warps * 32 + 255) / 256);
  if (grid > sm_budget() * 8) grid = sm_budget() * 8;
  if (grid < 1) grid = 1;
  ba_depth_kernel<<<grid, 256, 0, st>>>(w.plan.count, w.plan.kx, (int)cap, n6, w.Qg, w.ug, w.Eg,
                                        w.dX, P, patches);
  RVO_LAUNCH_CHECK("ba_depth_kernel");
  return RVO_OK;
}

static int ba_check(const char* who, const void* poses, const void* patches, const void* intr,
                    int E, int P, int t0, int t1) {
  RVO_CHECK_ARG(E >= 0, "%s: E=%d", who, E);
  RVO_CHECK_ARG(P >= 1 && P <= 9, "%s: patch size %d", who, P);
  RVO_CHECK_ARG(t0 >= 0 && t1 >= t0, "%s: bad window [%d,%d)", who, t0, t1);
  RVO_CHECK_ARG(poses && patches && intr, "%s: null pointer", who);
  return RVO_OK;
}

}  // namespace rvo

using namespace rvo;

extern "C" int64_t rvo_plan_bytes(int E) {
  if (E < 0) return -1;
  return (int64_t)plan_layout(nullptr, E).total;
}

extern "C" int rvo_graph_plan(const int64_t* kk, const int64_t* jj, int E, int64_t kmax,
                              int64_t jmax, void* plan, int64_t plan_bytes, void* stream) {
  return build_plan(kk, jj, E, kmax, jmax, plan, plan_bytes, (cudaStream_t)stream, nullptr);
}

extern "C" int rvo_plan_neighbors(const void* plan, int E, int64_t* ix, int64_t* jx, void* stream) {
  RVO_CHECK_ARG(E >= 0, "rvo_plan_neighbors: E=%d", E);
  if (E == 0) return RVO_OK;
  RVO_CHECK_ARG(plan && ix && jx, "rvo_plan_neighbors: null pointer");
  PlanView p = plan_layout(const_cast<void*>(plan), E);
  neighbors_kernel<<<grid1d(E), 256, 0, (cudaStream_t)stream>>>(p.perm, p.seg_of, E, ix, jx);
  RVO_LAUNCH_CHECK("neighbors_kernel");
  return RVO_OK;
}

extern "C" int rvo_plan_groups(const void* plan, int E, const int32_t** count,
                               const int32_t** perm, const int32_t** seg_of,
                               const int32_t** seg_start, const int64_t** kx) {
  RVO_CHECK_ARG(E >= 0 && plan, "rvo_plan_groups: bad arguments");
  PlanView p = plan_layout(const_cast<void*>(plan), E);
  if (count) *count = p.count;
  if (perm) *perm = p.perm;
  if (seg_of) *seg_of = p.seg_of;
  if (seg_start) *seg_start = p.seg_start;
  if (kx) *kx = p.kx;
  return RVO_OK;
}

extern "C" int rvo_plan_edge_groups(const void* plan, int E, const int32_t** grp) {
  RVO_CHECK_ARG(E >= 0 && plan && grp, "rvo_plan_edge_groups: bad arguments");
  PlanView p = plan_layout(const_cast<void*>(plan), E);
  *grp = reinterpret_cast<const int32_t*>(p.keys_a);
  return RVO_OK;
}

extern "C" int64_t rvo_neighbors_ws_bytes(int E) { return rvo_plan_bytes(E); }

extern "C" int rvo_neighbors(const int64_t* kk, const int64_t* jj, int E, int64_t kmax,
                             int64_t jmax, int64_t* ix, int64_t* jx, void* ws, int64_t ws_bytes,
                             void* stream) {
  if (E == 0) return RVO_OK;
  int rc = build_plan(kk, jj, E, kmax, jmax, ws, ws_bytes, (cudaStream_t)stream, nullptr);
  if (rc != RVO_OK) return rc;
  return rvo_plan_neighbors(ws, E, ix, jx, stream);
}

extern "C" int64_t rvo_ba_ws_bytes(int E, int64_t n_patches, int n_free) {
  if (E < 0 || n_free < 0) return -1;
  return (int64_t)ba_layout(nullptr, E, patch_cap(E, n_patches), n_free).total;
}

extern "C" int rvo_ba_plan(const int64_t* kk, const int64_t* jj, int E, int64_t n_poses,
                           int64_t n_patches, int n_free, void* ws, int64_t ws_bytes,
                           void* stream) {
  RVO_CHECK_ARG(ws, "rvo_ba_plan: null workspace");
  BaWs w = ba_layout(ws, E, patch_cap(E, n_patches), n_free);
  RVO_CHECK_ARG((int64_t)w.total <= ws_bytes, "rvo_ba_plan: workspace %lld < %lld bytes",
                (long long)ws_bytes, (long long)w.total);
  return build_plan(kk, jj, E, n_patches, n_poses, ws, (int64_t)w.plan.total, (cudaStream_t)stream,
                    nullptr);
}

extern "C" int rvo_ba_assemble(const float* poses, const float* patches, const float* intrinsics,
                               const float* target, const float* weight, const float* lmbda,
                               const int64_t* ii, const int64_t* jj, int E, int64_t n_patches, int P,
                               int t0, int t1, float* Sy, void* ws, int64_t ws_bytes,
                               void* stream) {
  int rc = ba_check("rvo_ba_assemble", poses, patches, intrinsics, E, P, t0, t1);
  if (rc != RVO_OK) return rc;
  if (E == 0) {
    const size_t n6 = 6 * (size_t)(t1 - t0);
    if (n6 && Sy) RVO_CUDA(cudaMemsetAsync(Sy, 0, n6 * (n6 + 1) * sizeof(float), (cudaStream_t)stream));
    return RVO_OK;
  }
  RVO_CHECK_ARG(target && weight && lmbda && ii && jj && ws, "rvo_ba_assemble: null pointer");
  const int64_t cap = patch_cap(E, n_patches);
  BaWs w = ba_layout(ws, E, cap, t1 - t0);
  RVO_CHECK_ARG((int64_t)w.total <= ws_bytes, "rvo_ba_assemble: workspace too small");
  return ba_assemble(w, poses, patches, intrinsics, target, weight, lmbda, ii, jj, E, cap, P, t0,
                     t1, Sy ? Sy : w.Sy, (cudaStream_t)stream);
}

extern "C" int rvo_ba_assemble_fused(const float* poses, const float* patches, const float* intrinsics,
                                     const float* coords, const float* delta, const float* weight, float ht,
                                     float wd, float* weight_out, const float* lmbda, const int64_t* ii,
                                     const int64_t* jj, int E, int64_t n_patches, int P, int t0, int t1, float* Sy,
                                     void* ws, int64_t ws_bytes, void* stream) {
  int rc = ba_check("rvo_ba_assemble_fused", poses, patches, intrinsics, E, P, t0, t1);
  if (rc != RVO_OK) return rc;
  if (E == 0) {
    const size_t n6 = 6 * (size_t)(t1 - t0);
    if (n6 && Sy) RVO_CUDA(cudaMemsetAsync(Sy, 0, n6 * (n6 + 1) * sizeof(float), (cudaStream_t)stream));
    return RVO_OK;
  }
  RVO_CHECK_ARG(coords && delta && weight && lmbda && ii && jj && ws, "rvo_ba_assemble_fused: null pointer");
  const int64_t cap = patch_cap(E, n_patches);
  BaWs w = ba_layout(ws, E, cap, t1 - t0);
  RVO_CHECK_ARG((int64_t)w.total <= ws_bytes, "rvo_ba_assemble_fused: workspace too small");
  BaFuse fz{coords, ht, wd, weight_out};
  return ba_assemble(w, poses, patches, intrinsics, delta, weight, lmbda, ii, jj, E, cap, P, t0, t1,
                     Sy ? Sy : w.Sy, (cudaStream_t)stream, nullptr, fz);
}

extern "C" int rvo_ba_solve(float* poses, float* patches, const float* Sy, int E, int64_t n_patches,
                            int P, int t0, int t1, void* ws, int64_t ws_bytes, void* stream) {
  int rc = ba_check("rvo_ba_solve", poses, patches, poses, E, P, t0, t1);
  if (rc != RVO_OK) return rc;
  if (E == 0) return RVO_OK;
  RVO_CHECK_ARG(ws, "rvo_ba_solve: null workspace");
  const int64_t cap = patch_cap(E, n_patches);
  BaWs w = ba_layout(ws, E, cap, t1 - t0);
  RVO_CHECK_ARG((int64_t)w.total <= ws_bytes, "rvo_ba_solve: workspace too small");
  return ba_solve(w, poses, patches, Sy ? Sy : w.Sy, cap, P, t0, t1, (cudaStream_t)stream);
}

static int ba_forward_impl(float* poses, float* patches, const float* intrinsics,
                           const float* target, const float* weight, const float* lmbda,
                           const int64_t* ii, const int64_t* jj, const int64_t* kk, int E,
                           int64_t n_poses, int64_t n_patches, int P, int PPF, int t0, int t1,
                           int iterations, int eff_impl, const void* ext_plan, void* ws,
                           int64_t ws_bytes, void* stream, const int32_t* t0_dev = nullptr,
                           BaFuse fz = BaFuse{nullptr, 0.f, 0.f, nullptr}) {
  (void)PPF; (void)eff_impl;  // block-sparse E is the only implementation; results do not depend on it
  int rc = ba_check("rvo_ba_forward", poses, patches, intrinsics, E, P, t0, t1);
  if (rc != RVO_OK) return rc;
  RVO_CHECK_ARG(iterations >= 0, "rvo_ba_forward: iterations=%d", iterations);
  RVO_CHECK_ARG(n_poses <= 0 || t1 <= n_poses, "rvo_ba_forward: t1=%d exceeds %lld poses", t1,
                (long long)n_poses);
  if (E == 0 || iterations == 0) return RVO_OK;
  RVO_CHECK_ARG(target && weight && lmbda && ii && jj && (kk || ext_plan) && ws, "rvo_ba_forward: null pointer");
  cudaStream_t st = (cudaStream_t)stream;
  const int64_t cap = patch_cap(E, n_patches);
  BaWs w = ba_layout(ws, E, cap, t1 - t0);
  RVO_CHECK_ARG((int64_t)w.total <= ws_bytes, "rvo_ba_forward: workspace %lld < %lld bytes",
                (long long)ws_bytes, (long long)w.total);
  if (ext_plan) {
    w.plan = plan_layout(const_cast<void*>(ext_plan), E);   // reuse the caller's (kk, jj) plan
  } else {
    rc = build_plan(kk, jj, E, n_patches, n_poses, ws, (int64_t)w.plan.total, st, nullptr);
    if (rc != RVO_OK) return rc;
  }
  for (int it = 0; it < iterations; it++) {
    BaFuse f = fz;
    if (it > 0) f.weight_out = nullptr;      // the filtered confidences do not change between iterations
    rc = ba_assemble(w, poses, patches, intrinsics, target, weight, lmbda, ii, jj, E, cap, P, t0, t1,
                     w.Sy, st, t0_dev, f);
    if (rc != RVO_OK) return rc;
    rc = ba_solve(w, poses, patches, w.Sy, cap, P, t0, t1, st, t0_dev);
    if (rc != RVO_OK) return rc;
  }
  return RVO_OK;
}

extern "C" int rvo_ba_solve_poses(float* poses, const float* Sy, int t0, int t1, void* ws,
                                  int64_t ws_bytes, void* stream) {
  RVO_CHECK_ARG(poses && Sy && ws && t0 >= 0 && t1 >= t0, "rvo_ba_solve_poses: bad arguments");
  const int N = t1 - t0, n6 = 6 * N;
  if (N == 0) return RVO_OK;
  BaWs w = ba_layout(ws, 1, 1, N);
  RVO_CHECK_ARG((int64_t)w.total <= ws_bytes, "rvo_ba_solve_poses: workspace too small");
  const bool in_smem = solve_smem_bytes(n6) <= kSmemMax;
  const size_t smem = in_smem ? solve_smem_bytes(n6) : 2 * (size_t)n6 * sizeof(float);
  const dim3 threads(32, n6 <= 24 ? 8 : 32);
  return launch_ba_solve(threads, smem, (cudaStream_t)stream, Sy, N, t0, nullptr, poses, w.dX,
                         in_smem ? nullptr : w.A);
}

extern "C" int rvo_ba_forward(float* poses, float* patches, const float* intrinsics,
                              const float* target, const float* weight, const float* lmbda,
                              const int64_t* ii, const int64_t* jj, const int64_t* kk, int E,
                              int64_t n_poses, int64_t n_patches, int P, int PPF, int t0, int t1,
                              int iterations, int eff_impl, void* ws, int64_t ws_bytes,
                              void* stream) {
  return ba_forward_impl(poses, patches, intrinsics, target, weight, lmbda, ii, jj, kk, E, n_poses,
                         n_patches, P, PPF, t0, t1, iterations, eff_impl, nullptr, ws, ws_bytes, stream);
}

extern "C" int rvo_ba_forward_planned(float* poses, float* patches, const float* intrinsics,
                                      const float* target, const float* weight, const float* lmbda,
                                      const int64_t* ii, const int64_t* jj, const void* plan, int E,
                                      int64_t n_poses, int64_t n_patches, int P, int t0, int t1,
                                      int iterations, void* ws, int64_t ws_bytes, void* stream) {
  RVO_CHECK_ARG(plan || E == 0, "rvo_ba_forward_planned: null plan");
  return ba_forward_impl(poses, patches, intrinsics, target, weight, lmbda, ii, jj, nullptr, E, n_poses,
                         n_patches, P, 0, t0, t1, iterations, 0, plan, ws, ws_bytes, stream);
}

extern "C" int rvo_ba_forward_dyn(float* poses, float* patches, const float* intrinsics,
                                  const float* target, const float* weight, const float* lmbda,
                                  const int64_t* ii, const int64_t* jj, const void* plan, int E,
                                  int64_t n_poses, int64_t n_patches, int P, int n_free,
                                  const int32_t* t0_dev, int iterations, void* ws, int64_t ws_bytes,
                                  void* stream) {
  RVO_CHECK_ARG(plan || E == 0, "rvo_ba_forward_dyn: null plan");
  RVO_CHECK_ARG(t0_dev && n_free >= 0, "rvo_ba_forward_dyn: needs the device-side window start");
  return ba_forward_impl(poses, patches, intrinsics, target, weight, lmbda, ii, jj, nullptr, E, 0,
                         n_patches, P, 0, 0, n_free, iterations, 0, plan, ws, ws_bytes, stream, t0_dev);
}

extern "C" int rvo_ba_forward_fused(float* poses, float* patches, const float* intrinsics,
                                    const float* coords, const float* delta, const float* weight,
                                    float ht, float wd, float* weight_out, const float* lmbda,
                                    const int64_t* ii, const int64_t* jj, const void* plan, int E,
                                    int64_t n_patches, int P, int n_free, const int32_t* t0_dev,
                                    int iterations, void* ws, int64_t ws_bytes, void* stream) {
  RVO_CHECK_ARG(plan || E == 0, "rvo_ba_forward_fused: null plan");
  RVO_CHECK_ARG(t0_dev && n_free >= 0, "rvo_ba_forward_fused: needs the device-side window start");
  RVO_CHECK_ARG(coords || E == 0, "rvo_ba_forward_fused: null coords");
  BaFuse fz{coords, ht, wd, weight_out};
  return ba_forward_impl(poses, patches, intrinsics, delta, weight, lmbda, ii, jj, nullptr, E, 0,
                         n_patches, P, 0, 0, n_free, iterations, 0, plan, ws, ws_bytes, stream, t0_dev, fz);
}

extern "C" int rvo_ba_forward_host(float* poses, float* patches, const float* intrinsics,
                                   const float* target, const float* weight, const float* lmbda,
                                   const int64_t* ii, const int64_t* jj, const int64_t* kk, int E,
                                   int64_t n_poses, int64_t n_patches, int P, int PPF, int t0,
                                   int t1, int iterations, int eff_impl, void* stream) {
  RVO_CHECK_ARG(E >= 0 && n_poses > 0 && n_patches > 0 && P >= 1, "rvo_ba_forward_host: bad sizes");
  cudaStream_t st = (cudaStream_t)stream;
  const size_t b_poses = (size_t)n_poses * 7 * 4, b_patch = (size_t)n_patches * 3 * P * P * 4;
  const size_t b_intr = (size_t)n_poses * 4 * 4, b_e2 = (size_t)E * 2 * 4, b_idx = (size_t)E * 8;
  const int64_t wsb = rvo_ba_ws_bytes(E, n_patches, t1 - t0);
  char* blk = nullptr;
  size_t off = 0;
  auto sub = [&](size_t bytes) { size_t o = off; off += align_up(bytes ? bytes : 4); return o; };
  const size_t o_poses = sub(b_poses), o_patch = sub(b_patch), o_intr = sub(b_intr),
               o_tgt = sub(b_e2), o_wgt = sub(b_e2), o_lm = sub(4), o_ii = sub(b_idx),
               o_jj = sub(b_idx), o_kk = sub(b_idx), o_ws = sub((size_t)wsb);
  RVO_CUDA(cudaMalloc((void**)&blk, off));
  int rc = RVO_OK;
  cudaError_t ce = cudaSuccess;
  auto h2d = [&](size_t o, const void* src, size_t bytes) {
    if (ce == cudaSuccess && bytes) ce = cudaMemcpyAsync(blk + o, src, bytes, cudaMemcpyHostToDevice, st);
  };
  h2d(o_poses, poses, b_poses); h2d(o_patch, patches, b_patch); h2d(o_intr, intrinsics, b_intr);
  h2d(o_tgt, target, b_e2); h2d(o_wgt, weight, b_e2); h2d(o_lm, lmbda, 4);
  h2d(o_ii, ii, b_idx); h2d(o_jj, jj, b_idx); h2d(o_kk, kk, b_idx);
  if (ce != cudaSuccess) rc = cuda_fail(ce, "rvo_ba_forward_host: H2D");
  if (rc == RVO_OK)
    rc = rvo_ba_forward((float*)(blk + o_poses), (float*)(blk + o_patch), (float*)(blk + o_intr),
                        (float*)(blk + o_tgt), (float*)(blk + o_wgt), (float*)(blk + o_lm),
                        (int64_t*)(blk + o_ii), (int64_t*)(blk + o_jj), (int64_t*)(blk + o_kk), E,
                        n_poses, n_patches, P, PPF, t0, t1, iterations, eff_impl, blk + o_ws, wsb, st);
  if (rc == RVO_OK) {
    ce = cudaMemcpyAsync(poses, blk + o_poses, b_poses, cudaMemcpyDeviceToHost, st);
    if (ce == cudaSuccess) ce = cudaMemcpyAsync(patches, blk + o_patch, b_patch, cudaMemcpyDeviceToHost, st);
    if (ce != cudaSuccess) rc = cuda_fail(ce, "rvo_ba_forward_host: D2H");
  }
  ce = cudaStreamSynchronize(st);
  if (rc == RVO_OK && ce != cudaSuccess) rc = cuda_fail(ce, "rvo_ba_forward_host: sync");
  cudaFree(blk);
  return rc;
}
