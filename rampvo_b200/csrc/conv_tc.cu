// conv_tc.cu — the dense convolutions of the RAMP encoders (ramp/extractor.py:8-57 ResidualBlock, :60-130
// BasicEncoder4, :272-311 MultiScaleBasicEncoder4: 7x7 s2, 3x3 s1/s2, 1x1 s1/s2) as ONE hand-written implicit-GEMM
// kernel on the Blackwell tensor cores (tcgen05.mma, TMEM accumulators, TMA weight loads), channels-last fp16.
//
//   Y[p, co] = bias[co] + sum_{tap, ci} X[pixel(p) * stride + tap - pad, ci] * W[co, tap, ci]        (fp32 accumulate)
//
// GEMM view: M = output pixels (tiles of 128), N = output channels (one slice of <= 192 per CTA), K = taps x input
// channels, walked in K blocks of 64 halves (one 128-byte row of the UMMA canonical K-major SWIZZLE_128B layout).
//   * the (zero-padded) weight slice [N x K] stays RESIDENT in shared memory for the CTA's life (<= 104 KB, loaded
//     once by TMA, one mbarrier per K block so the first MMA starts after 1/KB of it has landed);
//   * A tiles are im2col'ed on the fly: 4 producer warps gather 16-byte channel runs with cp.async (zero fill for
//     the padding ring, for taps past the kernel and for pixels past the image) into a 5-stage ring; the input may
//     be the channel CONCATENATION of two tensors (extractor.py:302,309 torch.cat((x, x_down)) never materialises);
//   * one elected thread issues 4 x tcgen05.mma (M128 x N x K16) per K block into one of two TMEM accumulators;
//   * 8 epilogue warps (2 accumulators x 4 lane quadrants, thread = output pixel): tcgen05.ld, + bias, fp16 pack,
//     64-byte-per-row stores, and — for the InstanceNorm layers — the per-channel sum / sum of squares of the
//     ROUNDED outputs reduced with a 31-shuffle warp transpose and accumulated per CTA (extractor.py:30-34: the
//     statistics pass of nn.InstanceNorm2d costs nothing extra; rvo_in_apply consumes them).
// Persistent grid: CTA b owns column slice b % n_slices and walks pixel tiles b / n_slices, + n_walkers, ...
#include <stdlib.h>

#include "common.cuh"
#include "tcgen05.cuh"

namespace rvo {

constexpr int kCvM = 128;
constexpr int kCvMaxStages = 8;
constexpr int kCvAStage = kCvM * 128;          // 16 KB: 128 rows x 64 halves
constexpr int kCvMaxKB = 13;                   // 7x7x16 = 784 -> 13 K blocks
constexpr int kCvMaxN = 192;
constexpr size_t kCvMaxDynSmem = 222 * 1024;   // + ~4 KB of static shared memory <= 227 KB per CTA
constexpr int kCvThreads = 640;                // warp 0 TMA (W), warp 1 MMA, warps 4-11 epilogue, 12-19 A producers

struct ConvArgs {
  const __half* src0;
  const __half* src1;      // second half of a channel concat (or null)
  int C0, C1;              // channels of the two sources (multiples of 8)
  int H, W;                // input size
  int Ho, Wo;              // output size
  int ks, stride, pad;
  int KB;                  // K blocks of 64 (K = ks*ks*(C0+C1), zero padded)
  int N;                   // output channels per slice (16..192, multiple of 16)
  int n_slices;            // Cout = N * n_slices
  const float* bias;       // [Cout] fp32 or null
  __half* out;             // [Ho*Wo, Cout]
  float* stats;            // [2*Cout] sum / sum of squares, accumulated atomically, or null
  uint32_t tmem_cols;      // power of two >= 2N
  int stages;              // A ring depth (3..8), as many as fit beside the resident weights
  uint32_t wo_magic;       // ceil(2^32 / Wo): p / Wo == umulhi(p, wo_magic) for p < 2^32 / Wo
};

// debugging aid (-DRVO_DEBUG builds): globaltimer stamps of CTA `trace_cta`, read back with rvo_conv_trace
#ifdef RVO_DEBUG
__device__ unsigned long long g_cv_trace[8 * 64];
__device__ int g_cv_trace_cta = 0;
__device__ __forceinline__ unsigned long long gtime() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
#define CV_TRACE(slot, i) do { if (blockIdx.x == g_cv_trace_cta && (i) < 64) g_cv_trace[(slot) * 64 + (i)] = gtime(); } while (0)
#else
#define CV_TRACE(slot, i) do { } while (0)
#endif

__global__ void __launch_bounds__(kCvThreads, 1)
conv_tc_kernel(const ConvArgs a, const __grid_constant__ TcTmap tmw) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ uint64_t wfull[kCvMaxKB], xfull[kCvMaxStages], xempty[kCvMaxStages], tfull[2], tempty[2];
  __shared__ uint32_t tmem_base_s;
  __shared__ float bias_s[kCvMaxN];
  __shared__ float stat_s[2 * kCvMaxN];
  __shared__ int4 kinfo[kCvMaxKB * 8];       // per (K block, 16-byte chunk): where the im2col gather reads from
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int slice = blockIdx.x % a.n_slices;
  const int walker = blockIdx.x / a.n_slices, n_walkers = gridDim.x / a.n_slices;
  const int P = a.Ho * a.Wo;
  const int n_tiles = (P + kCvM - 1) / kCvM;
  const int N = a.N, KB = a.KB, Cout = a.N * a.n_slices;

  if (tid == 0) CV_TRACE(0, 0);
  if (tid == 0) {
    for (int kb = 0; kb < KB; kb++) mbar_init(&wfull[kb], 1);
    for (int s = 0; s < a.stages; s++) {
      mbar_init(&xfull[s], 256);
      mbar_init(&xempty[s], 1);
    }
    for (int s = 0; s < 2; s++) {
      mbar_init(&tfull[s], 1);
      mbar_init(&tempty[s], 256);
    }
    asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
  }
  if (tid < N) bias_s[tid] = a.bias ? a.bias[slice * N + tid] : 0.f;
  if (tid < 2 * N) stat_s[tid] = 0.f;
  if (tid >= 512 && tid - 512 < KB * 8) {
    const int e = tid - 512, Cin = a.C0 + a.C1;
    const int k0 = e * 8, tap = k0 / Cin, c = k0 - tap * Cin;
    const int ky = tap / a.ks, kx = tap - ky * a.ks;
    int4 ki;
    ki.x = ky * a.W + kx;
    ki.y = tap < a.ks * a.ks ? (int)((1u << ky) | (1u << (8 + kx))) : -1;    // -1: no row ever qualifies
    ki.z = c >= a.C0 ? c - a.C0 : c;
    ki.w = c >= a.C0 ? -a.C1 : a.C0;
    kinfo[e] = ki;
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;\n" ::"r"(smem_u32(&tmem_base_s)),
                 "r"(a.tmem_cols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
  const uint32_t tmem_base = tmem_base_s;
  if (tid == 0) CV_TRACE(0, 1);
  const uint32_t wblk = (uint32_t)N * 128;                  // bytes of one resident K block of the weight slice
  const uint32_t W_u = smem_u32(smem);
  const uint32_t X_u = W_u + ((KB * wblk + 1023u) & ~1023u);

  if (warp == 0) {
    // ===== weight slice by TMA, once =====
    if (lane == 0)
      for (int kb = 0; kb < KB; kb++) {
        mbar_expect_tx(&wfull[kb], wblk);
        tma_load_2d(W_u + kb * wblk, &tmw, kb * 64, slice * N, &wfull[kb]);
      }
  } else if (warp == 1) {
    // ===== MMA issuer =====
    int s = 0, ph = 0, lt = 0;
    bool first = true;
    const uint32_t idesc = umma_idesc_f16(kCvM, N);
    for (int t = walker; t < n_tiles; t += n_walkers, lt++) {
      const int acc = lt & 1;
      mbar_wait_spin(&tempty[acc], ((lt >> 1) & 1) ^ 1);
      for (int kb = 0; kb < KB; kb++) {
        if (first) mbar_wait_spin(&wfull[kb], 0);
        if (lane == 0 && kb == 0) CV_TRACE(1, lt);            // weights (first tile) / accumulator available
        mbar_wait_spin(&xfull[s], ph);
        // the A block was written by the producers' cp.async (generic proxy) and is read by the tensor core
        // (async proxy): one proxy fence by the issuing thread after the acquire
        asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
        asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
        if (lane == 0) {
          if (kb == 0) CV_TRACE(2, lt);                       // first A block of the tile has landed
          const uint64_t da0 = umma_desc(X_u + s * kCvAStage), db0 = umma_desc(W_u + kb * wblk);
#pragma unroll
          for (int k = 0; k < 4; k++)
            umma_f16(tmem_base + acc * N, da0 + (uint64_t)((k * 32) >> 4), db0 + (uint64_t)((k * 32) >> 4), idesc,
                     (kb | k) ? 1u : 0u);
          umma_commit(&xempty[s]);
          if (kb == KB - 1) {
            umma_commit(&tfull[acc]);
            CV_TRACE(3, lt);                                  // last MMA of the tile issued
          }
        }
        __syncwarp();
        if (++s == a.stages) { s = 0; ph ^= 1; }
      }
      first = false;
    }
    if (first)
      for (int kb = 0; kb < KB; kb++) mbar_wait_spin(&wfull[kb], 0);
  } else if (warp >= 12) {
    // ===== A producers (im2col gather): 8 warps; thread owns 16-byte chunk `ch` of rows r0 + 32 j.
    // A lone warp runs its dependent integer chain at ~5 cycles per instruction, so the loop body is kept to a
    // shared-memory table lookup (what tap / channel run this thread's chunk of K block kb is) plus, per row, a bit
    // test, an add and one multiply-add; per tile one magic-number division.  (Measured with the straightforward
    // index arithmetic: 520 ns of producer instructions per K block, 1.9 us of set-up per tile.) =====
    const int ptid = tid - 12 * 32, ch = ptid & 7, r0 = ptid >> 3;
    const int stages = a.stages;
    int it = 0, s = 0, ph = 1;
    for (int t = walker; t < n_tiles; t += n_walkers) {
      int pix0[4];
      uint32_t okm[4];
      const uint32_t p0 = (uint32_t)(t * kCvM + r0);
      int oy = (int)__umulhi(p0, a.wo_magic), ox = (int)p0 - oy * a.Wo;
#pragma unroll
      for (int j = 0; j < 4; j++) {
        const int iy0 = oy * a.stride - a.pad, ix0 = ox * a.stride - a.pad;
        pix0[j] = iy0 * a.W + ix0;
        // bit k of the low byte: tap row k lies inside the image; bit 8 + k: tap column k does
        const int ylo = max(0, -iy0), yhi = min(a.ks, a.H - iy0), xlo = max(0, -ix0), xhi = min(a.ks, a.W - ix0);
        const uint32_t my = yhi > ylo ? (((1u << yhi) - 1u) & ~((1u << ylo) - 1u)) : 0u;
        const uint32_t mx = xhi > xlo ? (((1u << xhi) - 1u) & ~((1u << xlo) - 1u)) : 0u;
        okm[j] = ((int)p0 + 32 * j < P) ? (my | (mx << 8)) : 0u;
        ox += 32;
        while (ox >= a.Wo) { ox -= a.Wo; oy++; }
      }
      if (ptid == 0) CV_TRACE(6, 63 - (it < 15 ? it : 15));
      for (int kb = 0; kb < KB; kb++, it++) {
        mbar_wait(&xempty[s], ph);
        if (ptid == 0) CV_TRACE(6, it);                       // stage free
        const int4 ki = kinfo[kb * 8 + ch];                   // {tap offset in pixels, needed mask bits, element offset, pixel stride}
        const __half* base = ki.w < 0 ? a.src1 : a.src0;
        const int cs = ki.w < 0 ? -ki.w : ki.w;
        const uint32_t dst0 = X_u + s * kCvAStage + r0 * 128 + (uint32_t)((ch ^ (r0 & 7)) << 4);
#pragma unroll
        for (int j = 0; j < 4; j++) {
          const bool ok = (okm[j] & (uint32_t)ki.y) == (uint32_t)ki.y;
          const int off = ok ? (pix0[j] + ki.x) * cs + ki.z : 0;
          cp_async16(dst0 + j * (32 * 128), base + off, ok ? 16u : 0u);
        }
        // the barrier arrival is triggered BY THE HARDWARE when this thread's copies have landed: the producer
        // never waits for its own loads (a wait_group + fence.proxy.async per K block drains every copy in
        // flight and degrades the ring to one K block per memory round trip: measured 590 ns per K block)
        asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];\n" ::"r"(smem_u32(&xfull[s])) : "memory");
        if (ptid == 0) CV_TRACE(7, it);                       // K block issued
        if (++s == stages) { s = 0; ph ^= 1; }
      }
    }
    asm volatile("cp.async.wait_all;\n" ::: "memory");
  } else if (warp >= 4) {
    // ===== epilogue: all 8 warps drain every tile: lane quadrant q (thread = output pixel), the two warps of a
    // quadrant take alternate 32-column chunks (one warp runs this dependent LDTM -> add -> pack -> store chain at
    // ~1.2 us per chunk; a CTA only sees 1-5 tiles, so finishing a tile sooner beats overlapping two of them) =====
    const int hf = (warp - 4) >> 2, q = warp & 3;
    int lt = 0;
    for (int t = walker; t < n_tiles; t += n_walkers, lt++) {
      const int g = lt & 1;
      const uint32_t tlane = tmem_base + g * N + ((uint32_t)(q * 32) << 16);
      const int p = t * kCvM + q * 32 + lane;
      const bool live = p < P;
      mbar_wait(&tfull[g], (lt >> 1) & 1);
      asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
      if (q == 0 && hf == 0 && lane == 0) CV_TRACE(4, lt);    // accumulator complete, epilogue starts
      for (int c0 = 32 * hf; c0 < N; c0 += 64) {
        float v[32];
        if (c0 + 32 <= N) {
          tmem_ld32(tlane + c0, v);
        } else {                                              // N = 16 mod 32: last half chunk
          tmem_ld16(tlane + c0, v);
#pragma unroll
          for (int j = 16; j < 32; j++) v[j] = 0.f;
        }
        const int nc = (N - c0) < 32 ? (N - c0) : 32;
        uint32_t pk[16];
#pragma unroll
        for (int j = 0; j < 16; j++) {
          const float x0 = v[2 * j] + bias_s[(c0 + 2 * j) < N ? c0 + 2 * j : 0];
          const float x1 = v[2 * j + 1] + bias_s[(c0 + 2 * j + 1) < N ? c0 + 2 * j + 1 : 0];
          const __half2 h = __floats2half2_rn(x0, x1);
          pk[j] = *reinterpret_cast<const uint32_t*>(&h);
          if (a.stats) {                                      // statistics of the ROUNDED values, zero for dead rows
            const float2 f = __half22float2(h);
            v[2 * j] = live ? f.x : 0.f;
            v[2 * j + 1] = live ? f.y : 0.f;
          }
        }
        if (live) {
          __half* dst = a.out + (int64_t)p * Cout + slice * N + c0;
          if (nc == 32) {
            uint4* d4 = reinterpret_cast<uint4*>(dst);
            d4[0] = make_uint4(pk[0], pk[1], pk[2], pk[3]);
            d4[1] = make_uint4(pk[4], pk[5], pk[6], pk[7]);
            d4[2] = make_uint4(pk[8], pk[9], pk[10], pk[11]);
            d4[3] = make_uint4(pk[12], pk[13], pk[14], pk[15]);
          } else {
            uint4* d4 = reinterpret_cast<uint4*>(dst);
            d4[0] = make_uint4(pk[0], pk[1], pk[2], pk[3]);
            d4[1] = make_uint4(pk[4], pk[5], pk[6], pk[7]);
          }
        }
        if (a.stats) {
          // column sums over the 32 rows of this warp: halving exchange, lane l ends up with column l
          float sq[32];
#pragma unroll
          for (int j = 0; j < 32; j++) sq[j] = v[j] * v[j];
#pragma unroll
          for (int off = 16, n = 32; off >= 1; off >>= 1, n >>= 1) {
            const bool upper = (lane & off) != 0;
#pragma unroll
            for (int i = 0; i < n / 2; i++) {
              const float send = upper ? v[i] : v[i + n / 2];
              const float keep = upper ? v[i + n / 2] : v[i];
              v[i] = keep + __shfl_xor_sync(0xffffffffu, send, off);
              const float send2 = upper ? sq[i] : sq[i + n / 2];
              const float keep2 = upper ? sq[i + n / 2] : sq[i];
              sq[i] = keep2 + __shfl_xor_sync(0xffffffffu, send2, off);
            }
          }
          if (lane < nc) {
            atomicAdd(&stat_s[c0 + lane], v[0]);
            atomicAdd(&stat_s[N + c0 + lane], sq[0]);
          }
        }
      }
      asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
      if (q == 0 && hf == 0 && lane == 0) CV_TRACE(5, lt);    // epilogue of the tile done
      mbar_arrive(&tempty[g]);
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
  __syncthreads();
  if (a.stats && tid < 2 * N) {
    const int c = tid < N ? tid : tid - N;
    atomicAdd(&a.stats[(tid < N ? 0 : Cout) + slice * N + c], stat_s[tid]);
  }
  if (warp == 1) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;\n" ::"r"(tmem_base), "r"(a.tmem_cols)
                 : "memory");
  }
  if (tid == 0) CV_TRACE(0, 2);
}

}  // namespace rvo

using namespace rvo;

#ifdef RVO_DEBUG
extern "C" int rvo_conv_trace(unsigned long long* host_out, int next_cta) {
  RVO_CUDA(cudaDeviceSynchronize());
  RVO_CUDA(cudaMemcpyFromSymbol(host_out, g_cv_trace, sizeof(unsigned long long) * 8 * 64));
  RVO_CUDA(cudaMemcpyToSymbol(g_cv_trace_cta, &next_cta, sizeof(int)));
  return RVO_OK;
}
#endif

extern "C" int rvo_conv2d_kpad(int ks, int Cin) {
  if (ks < 1 || Cin < 8 || Cin % 8) return -1;
  return ((ks * ks * Cin + 63) / 64) * 64;
}

extern "C" int rvo_conv2d_nhwc(const void* src0, int C0, const void* src1, int C1, int H, int W, int ks, int stride,
                               int pad, const void* w_packed, const float* bias, int Cout, void* out, float* stats,
                               void* stream) {
  RVO_CHECK_ARG(src0 && w_packed && out, "rvo_conv2d_nhwc: null pointer");
  RVO_CHECK_ARG(C0 >= 8 && C0 % 8 == 0 && C1 >= 0 && C1 % 8 == 0 && (C1 == 0 || src1),
                "rvo_conv2d_nhwc: channels %d + %d (multiples of 8)", C0, C1);
  RVO_CHECK_ARG(ks >= 1 && ks <= 7 && stride >= 1 && pad >= 0 && H >= 1 && W >= 1, "rvo_conv2d_nhwc: geometry");
  RVO_CHECK_ARG(((reinterpret_cast<uintptr_t>(src0) | reinterpret_cast<uintptr_t>(src1) |
                  reinterpret_cast<uintptr_t>(out)) & 15u) == 0, "rvo_conv2d_nhwc: 16-byte alignment");
  const int Cin = C0 + C1, Kpad = rvo_conv2d_kpad(ks, Cin), KB = Kpad / 64;
  RVO_CHECK_ARG(KB >= 1 && KB <= kCvMaxKB, "rvo_conv2d_nhwc: K = %d does not fit (max %d)", ks * ks * Cin, kCvMaxKB * 64);
  RVO_CHECK_ARG(Cout >= 16 && Cout % 16 == 0, "rvo_conv2d_nhwc: Cout = %d (multiple of 16)", Cout);
  const int Ho = (H + 2 * pad - ks) / stride + 1, Wo = (W + 2 * pad - ks) / stride + 1;
  RVO_CHECK_ARG(Ho >= 1 && Wo >= 1, "rvo_conv2d_nhwc: empty output");
  const int64_t P = (int64_t)Ho * Wo;
  const int n_tiles = (int)((P + kCvM - 1) / kCvM);
  // column slices: the fewest slices of <= 192 channels whose resident weights + the A ring fit shared memory,
  // then split further (down to 32 columns) while that buys fuller waves on the 148 SMs
  int n_slices = 1;
  auto fits = [&](int ns) {
    const int N = Cout / ns;
    return Cout % ns == 0 && N % 16 == 0 && N <= kCvMaxN &&
           (size_t)KB * N * 128 + 1024 + (size_t)4 * kCvAStage + 1024 <= kCvMaxDynSmem;
  };
  while (n_slices <= 16 && !fits(n_slices)) n_slices++;
  RVO_CHECK_ARG(n_slices <= 16, "rvo_conv2d_nhwc: Cout = %d with K = %d does not fit", Cout, Kpad);
  while (fits(n_slices * 2) && Cout / (n_slices * 2) >= 32 && (int64_t)n_tiles * n_slices < sm_budget()) n_slices *= 2;
  const int N = Cout / n_slices;
  int grid = sm_budget() - sm_budget() % n_slices;
  if (grid < n_slices) grid = n_slices;
  if ((int64_t)n_tiles * n_slices < grid) grid = n_tiles * n_slices;
  TcTmap tmw;
  int rc = make_tmap_2d_f16(w_packed, Cout, Kpad, Kpad, N, &tmw, "rvo_conv2d_nhwc(w)");
  if (rc != RVO_OK) return rc;
  ConvArgs a;
  a.src0 = (const __half*)src0; a.src1 = (const __half*)src1; a.C0 = C0; a.C1 = C1; a.H = H; a.W = W;
  a.Ho = Ho; a.Wo = Wo; a.ks = ks; a.stride = stride; a.pad = pad; a.KB = KB; a.N = N; a.n_slices = n_slices;
  a.bias = bias; a.out = (__half*)out; a.stats = stats;
  uint32_t cols = 32;
  while (cols < 2u * N) cols <<= 1;
  a.tmem_cols = cols;
  const size_t wbytes = (((size_t)KB * N * 128 + 1023) & ~(size_t)1023);
  int stages = (int)((kCvMaxDynSmem - 1024 - wbytes) / kCvAStage);
  a.stages = stages > kCvMaxStages ? kCvMaxStages : stages;
  a.wo_magic = (uint32_t)((0x100000000ull + (uint64_t)Wo - 1) / (uint64_t)Wo);
  RVO_CHECK_ARG(P < (int64_t)(0x100000000ull / (uint64_t)Wo) && (int64_t)H * W * (C0 > C1 ? C0 : C1) < 0x7fffffffll,
                "rvo_conv2d_nhwc: tensor too large for 32-bit index arithmetic");
  cudaStream_t st = (cudaStream_t)stream;
  const size_t smem = wbytes + (size_t)a.stages * kCvAStage + 1024;
  // one fixed opt-in for every layer shape: the attribute is per-function state, and a captured graph node must stay
  // launchable after a later call for a smaller layer (ncu replays graph nodes with the CURRENT attribute)
  RVO_CUDA(cudaFuncSetAttribute(conv_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kCvMaxDynSmem));
  conv_tc_kernel<<<grid, kCvThreads, smem, st>>>(a, tmw);
  RVO_LAUNCH_CHECK("conv_tc_kernel");
  return RVO_OK;
}
