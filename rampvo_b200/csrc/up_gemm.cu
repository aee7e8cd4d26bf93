// up_gemm.cu — the Linear layers of the update operator (ramp/net.py:36-67, ramp/blocks.py:15-50) as a
// hand-written tcgen05 GEMM, sm_100a:  Y[M, N] = act(X[M, 384] W[N, 384]^T + b), fp16 in / fp32 accumulate /
// fp16 out, act = identity or ReLU (the nn.Linear + ReLU pairs of the reference under autocast).
//
// Shape of the problem: M = number of edges (45 312 at default.yaml, 660 600 at precise.yaml), K = 384,
// N = 384 or 768 — a tall-skinny GEMM whose weights (288 KB) are tiny and whose activations stream.  So:
//   * a CTA keeps ONE 192-row slice of W resident in shared memory for its whole life (6 K blocks of
//     [192 x 64] fp16 = 144 KB, loaded once by TMA) and walks M tiles of 128 rows persistently;
//   * X tiles stream through a 5-stage ring of [128 x 64] K blocks in the UMMA canonical K-major
//     SWIZZLE_128B layout, gathered by 4 producer warps with 16-byte cp.async (rows past M are zero-filled),
//     four K blocks in flight.  (Measured: one TMA box per K block — 128 rows of 128 bytes — tops out near
//     12 B/clk per SM here, half of what the tensor core consumes; the LSU path keeps up.)
//   * one elected thread issues 4 x tcgen05.mma (M128 N192 K16) per K block into one of two 192-column
//     TMEM accumulators; tcgen05.commit hands the smem stage back to the TMA producer and, after the 6th
//     K block, the accumulator to the epilogue;
//   * epilogue: 2 accumulators x 4 TMEM lane quadrants x 2 warps (alternate 32-column chunks), thread = output
//     row: tcgen05.ld 32 columns at a time, + bias (fp32, from shared memory), optional ReLU, fp16 pack, 16-byte stores — while the tensor
//     core already works on the next M tile in the other accumulator.
// Grid: 148 CTAs = (N / 192) column slices x 148 / (N / 192) row walkers.
#include <stdlib.h>

#include "common.cuh"
#include "tcgen05.cuh"

namespace rvo {

constexpr int kUgK = 384;                    // input features (resident-weight kernel)
constexpr int kUgKB = kUgK / 64;             // 6 K blocks of 64 halves (128 bytes)
constexpr int kUgM = 128;                    // rows per M tile = MMA M
constexpr int kUgN = 192;                    // output columns per CTA = MMA N
constexpr int kUgStages = 5;
constexpr int kUgWBytes = kUgKB * kUgN * 128;            // 147 456
constexpr int kUgXStage = kUgM * 128;                    // 16 384
constexpr int kUgSmemBytes = kUgWBytes + kUgStages * kUgXStage + 1024;
constexpr int kUgThreads = 768;              // warp 0 TMA (W), warp 1 MMA, warps 4-19 epilogue, 20-23 X producers

// debugging aid (RVO_UP_TRACE=1): clock64 stamps of CTA 0, read back with rvo_up_trace
__device__ long long g_ug_trace[8 * 64];
#define UG_TRACE(slot, i) do { if (trace && blockIdx.x == 0 && (i) < 64) g_ug_trace[(slot) * 64 + (i)] = clock64(); } while (0)

template <bool kRelu>
__global__ void __launch_bounds__(kUgThreads, 1)
up_linear_kernel(const __half* __restrict__ X, int64_t ldx, const int64_t* __restrict__ gather,
                 const __grid_constant__ TcTmap tmw, const __half* __restrict__ bias, int M, int n_slices,
                 __half* __restrict__ Y, int64_t ldy, int trace_arg) {
#ifdef RVO_DEBUG
  const int trace = trace_arg;      // -DRVO_DEBUG builds only
#else
  constexpr int trace = 0;          // release library: trace stamps and the store-skipping mode are compiled out
#endif
  extern __shared__ uint8_t smem_raw[];
  __shared__ uint64_t wfull[kUgKB], xfull[kUgStages], xempty[kUgStages], tfull[2], tempty[2];
  __shared__ uint32_t tmem_base_s;
  __shared__ float bias_s[kUgN];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int slice = blockIdx.x % n_slices;                 // which 192 output columns
  const int walker = blockIdx.x / n_slices, n_walkers = gridDim.x / n_slices;
  const int n_tiles = (M + kUgM - 1) / kUgM;

  if (tid == 0) {
    for (int kb = 0; kb < kUgKB; kb++) mbar_init(&wfull[kb], 1);
    for (int s = 0; s < kUgStages; s++) {
      mbar_init(&xfull[s], 128);
      mbar_init(&xempty[s], 1);
    }
    for (int s = 0; s < 2; s++) {
      mbar_init(&tfull[s], 1);
      mbar_init(&tempty[s], 256);
    }
    asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
  }
  if (tid < kUgN) bias_s[tid] = bias ? __half2float(bias[slice * kUgN + tid]) : 0.f;
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;\n" ::"r"(
                     smem_u32(&tmem_base_s))
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
  const uint32_t tmem_base = tmem_base_s;
  const uint32_t W_u = smem_u32(smem), X_u = W_u + kUgWBytes;

  if (warp == 0) {
    // ===== TMA producer =====
    if (lane == 0) {
      UG_TRACE(0, 0);
      for (int kb = 0; kb < kUgKB; kb++) {                 // one barrier per K block: the first MMA needs 24 KB, not 144
        mbar_expect_tx(&wfull[kb], (uint32_t)(kUgN * 128));
        tma_load_2d(W_u + kb * (kUgN * 128), &tmw, kb * 64, slice * kUgN, &wfull[kb]);
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer: UTCHMMA issue blocks while the tensor queue is full, so whatever this thread does
    // between two issues is tensor idle time — stage / phase counters are incremental, no division =====
    int s = 0, ph = 0, lt = 0;
    bool first = true;
    for (int t = walker; t < n_tiles; t += n_walkers, lt++) {
      const int acc = lt & 1;
      mbar_wait_spin(&tempty[acc], ((lt >> 1) & 1) ^ 1);
#pragma unroll
      for (int kb = 0; kb < kUgKB; kb++) {
        if (first) mbar_wait_spin(&wfull[kb], 0);
        mbar_wait_spin(&xfull[s], ph);
        // X block: written by cp.async (generic proxy), read by the tensor core (async proxy)
        asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
        asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
        if (lane == 0) {
          if (kb == 0) UG_TRACE(2, lt);
          const uint64_t da0 = umma_desc(X_u + s * kUgXStage), db0 = umma_desc(W_u + kb * (kUgN * 128));
#pragma unroll
          for (int k = 0; k < 4; k++)
            umma_f16(tmem_base + acc * 256, da0 + (uint64_t)((k * 32) >> 4), db0 + (uint64_t)((k * 32) >> 4),
                     umma_idesc_f16(kUgM, kUgN), (kb | k) ? 1u : 0u);
          umma_commit(&xempty[s]);                         // stage reusable once these MMAs have read it
          if (kb == kUgKB - 1) {
            umma_commit(&tfull[acc]);                      // accumulator complete
            UG_TRACE(3, lt);
          }
        }
        __syncwarp();
        if (++s == kUgStages) { s = 0; ph ^= 1; }
      }
      first = false;
    }
    if (first)                                             // no tile for this CTA: still wait for the W loads
      for (int kb = 0; kb < kUgKB; kb++) mbar_wait_spin(&wfull[kb], 0);
  } else if (warp >= 20) {
    // ===== X producers: thread owns 16-byte chunk `ch` of rows r0 + 16 j of every K block =====
    const int ptid = tid - 20 * 32, ch = ptid & 7, r0 = ptid >> 3;
    int it = 0;                                            // K-block counter over all tiles of this CTA
    for (int t = walker; t < n_tiles; t += n_walkers) {
      // source row of each of this thread's 8 rows: the row itself, or gather[row] (< 0: a zero row — the
      // mask_ix * net[:, ix] of ramp/net.py:78-82 folded into the operand load)
      int64_t src[8];
#pragma unroll
      for (int j = 0; j < 8; j++) {
        const int row = t * kUgM + r0 + 16 * j;
        src[j] = row < M ? (gather ? gather[row] : (int64_t)row) : -1;
      }
      for (int kb = 0; kb < kUgKB; kb++, it++) {
        const int s = it % kUgStages;
        mbar_wait(&xempty[s], ((it / kUgStages) & 1) ^ 1);
        if (ptid == 0) UG_TRACE(1, it);
        const uint32_t dst0 = X_u + s * kUgXStage + r0 * 128 + (uint32_t)((ch ^ (r0 & 7)) << 4);
        const __half* src0 = X + kb * 64 + ch * 8;
#pragma unroll
        for (int j = 0; j < 8; j++) {                      // (r0 + 16 j) & 7 == r0 & 7: one swizzle per thread
          const bool ok = src[j] >= 0;
          cp_async16(dst0 + j * (16 * 128), src0 + (ok ? src[j] : 0) * ldx, ok ? 16u : 0u);
        }
        // hardware-triggered arrival: the barrier is signalled when THIS thread's copies of the K block have
        // landed.  (The first version waited with cp.async.wait_group + fence.proxy.async before a software
        // arrive; that fence drains every copy still in flight, so the 5-stage ring delivered one K block per
        // memory round trip — ~630 ns per K block, three times the 196 ns the four MMAs need.)
        asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];\n" ::"r"(smem_u32(&xfull[s])) : "memory");
      }
    }
    asm volatile("cp.async.wait_all;\n" ::: "memory");
  } else if (warp >= 4) {
    // ===== epilogue: group g drains accumulator g (tiles lt == g mod 2); quadrant q, thread = row; the two
    // warps of a (group, quadrant) take alternate 32-column chunks (a single warp per scheduler issues at
    // ~0.35 IPC on this dependent LDTM -> add -> pack -> store chain) =====
    const int g = (warp - 4) >> 3, hf = ((warp - 4) >> 2) & 1, q = warp & 3;
    const uint32_t tlane = tmem_base + g * 256 + ((uint32_t)(q * 32) << 16);
    int lt = 0;
    for (int t = walker; t < n_tiles; t += n_walkers, lt++) {
      if ((lt & 1) != g) continue;
      const int row = t * kUgM + q * 32 + lane;
      mbar_wait(&tfull[g], (lt >> 1) & 1);
      asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
      if (q == 0 && hf == 0 && lane == 0) UG_TRACE(4, lt);
#pragma unroll 1
      for (int c0 = 32 * hf; c0 < kUgN; c0 += 64) {
        float v[32];
        tmem_ld32(tlane + c0, v);
        uint32_t pk[16];
#pragma unroll
        for (int j = 0; j < 16; j++) {
          float a = v[2 * j] + bias_s[c0 + 2 * j], b = v[2 * j + 1] + bias_s[c0 + 2 * j + 1];
          if (kRelu) {
            a = fmaxf(a, 0.f);
            b = fmaxf(b, 0.f);
          }
          const __half2 h = __floats2half2_rn(a, b);
          pk[j] = *reinterpret_cast<const uint32_t*>(&h);
        }
        // lanes 2i / 2i+1 trade 32-byte halves: each store then covers 16 rows x 64 contiguous bytes instead
        // of 32 rows x 32 bytes (half the cache lines per request)
        uint32_t rx[8];
        const bool odd = lane & 1;
#pragma unroll
        for (int j = 0; j < 8; j++) rx[j] = __shfl_xor_sync(0xffffffffu, odd ? pk[j] : pk[8 + j], 1);
        if (trace != 2) {
          const int rowA = row & ~1, rowB = row | 1;       // this pair's even / odd row
          __half* dA = Y + (int64_t)rowA * ldy + slice * kUgN + c0 + (odd ? 16 : 0);
          __half* dB = Y + (int64_t)rowB * ldy + slice * kUgN + c0 + (odd ? 16 : 0);
          // even lane: A = own low half, B = partner's low half; odd lane: A = partner's high half, B = own high
          if (rowA < M)
            asm volatile("st.global.v8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};\n" ::"l"(dA),
                         "r"(odd ? rx[0] : pk[0]), "r"(odd ? rx[1] : pk[1]), "r"(odd ? rx[2] : pk[2]),
                         "r"(odd ? rx[3] : pk[3]), "r"(odd ? rx[4] : pk[4]), "r"(odd ? rx[5] : pk[5]),
                         "r"(odd ? rx[6] : pk[6]), "r"(odd ? rx[7] : pk[7])
                         : "memory");
          if (rowB < M)
            asm volatile("st.global.v8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};\n" ::"l"(dB),
                         "r"(odd ? pk[8] : rx[0]), "r"(odd ? pk[9] : rx[1]), "r"(odd ? pk[10] : rx[2]),
                         "r"(odd ? pk[11] : rx[3]), "r"(odd ? pk[12] : rx[4]), "r"(odd ? pk[13] : rx[5]),
                         "r"(odd ? pk[14] : rx[6]), "r"(odd ? pk[15] : rx[7])
                         : "memory");
        }
      }
      asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
      if (q == 0 && hf == 0 && lane == 0) UG_TRACE(5, lt);
      mbar_arrive(&tempty[g]);
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
  __syncthreads();
  if (tid == 0) UG_TRACE(0, 2);
  if (warp == 1) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;\n" ::"r"(tmem_base) : "memory");
  }
}

}  // namespace rvo

using namespace rvo;

extern "C" int rvo_up_linear_gather(const void* x16, int64_t ldx, const int64_t* gather, const void* w16,
                                    const void* bias16, int M, int K, int N, int relu, void* y16, int64_t ldy,
                                    void* stream);

extern "C" int rvo_up_linear(const void* x16, int64_t ldx, const void* w16, const void* bias16, int M, int K,
                             int N, int relu, void* y16, int64_t ldy, void* stream) {
  return rvo_up_linear_gather(x16, ldx, nullptr, w16, bias16, M, K, N, relu, y16, ldy, stream);
}

extern "C" int rvo_up_linear_gather(const void* x16, int64_t ldx, const int64_t* gather, const void* w16,
                                    const void* bias16, int M, int K, int N, int relu, void* y16, int64_t ldy,
                                    void* stream) {
  RVO_CHECK_ARG(M >= 0 && K == kUgK && N > 0 && N % kUgN == 0 && N / kUgN <= 16,
                "rvo_up_linear: M=%d K=%d N=%d (K must be %d, N a multiple of %d, at most 16 column slices)", M, K, N,
                kUgK, kUgN);
  if (M == 0) return RVO_OK;
  RVO_CHECK_ARG(x16 && w16 && y16, "rvo_up_linear: null pointer");
  RVO_CHECK_ARG(ldx >= K && ldy >= N && ldy % 16 == 0 && (reinterpret_cast<uintptr_t>(y16) & 31u) == 0,
                "rvo_up_linear: row pitches / output alignment");
  RVO_CHECK_ARG((reinterpret_cast<uintptr_t>(x16) & 15u) == 0 && ldx % 8 == 0, "rvo_up_linear: x alignment");
  TcTmap tmw;
  int rc = make_tmap_2d_f16(w16, N, K, K, kUgN, &tmw, "rvo_up_linear(w)");
  if (rc != RVO_OK) return rc;
  cudaStream_t st = (cudaStream_t)stream;
  const int ug_slices = N / kUgN;
  int ug_grid = sm_budget() - sm_budget() % ug_slices;        // (column slices) x (row walkers)
  if (ug_grid < ug_slices) ug_grid = ug_slices;
#ifdef RVO_DEBUG
  const int trace = getenv("RVO_UP_TRACE") ? atoi(getenv("RVO_UP_TRACE")) : 0;
#else
  const int trace = 0;
#endif
  if (relu) {
    RVO_CUDA(cudaFuncSetAttribute(up_linear_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kUgSmemBytes));
    up_linear_kernel<true><<<ug_grid, kUgThreads, kUgSmemBytes, st>>>((const __half*)x16, ldx, gather, tmw,
                                                                      (const __half*)bias16, M, N / kUgN,
                                                                      (__half*)y16, ldy, trace);
  } else {
    RVO_CUDA(cudaFuncSetAttribute(up_linear_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kUgSmemBytes));
    up_linear_kernel<false><<<ug_grid, kUgThreads, kUgSmemBytes, st>>>((const __half*)x16, ldx, gather, tmw,
                                                                       (const __half*)bias16, M, N / kUgN,
                                                                       (__half*)y16, ldy, trace);
  }
  RVO_LAUNCH_CHECK("up_linear_kernel");
  return RVO_OK;
}

extern "C" int rvo_up_trace(long long* host_out) {
  RVO_CUDA(cudaDeviceSynchronize());
  RVO_CUDA(cudaMemcpyFromSymbol(host_out, g_ug_trace, sizeof(long long) * 8 * 64));
  return RVO_OK;
}
