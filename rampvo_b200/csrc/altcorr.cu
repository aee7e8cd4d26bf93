// altcorr.cu — patch gather (patchify) and patch<->frame correlation lookup (corr) for sm_100a.
//
// corr fast path (channels-last fp16, P=3, C%32==0): ONE WARP PER EDGE.  The 9 patch pixels of an
// edge reproject within ~1 px of each other, so their 8x8 lookup windows share one (8+2)x(8+2)
// union window per pyramid level.  The warp computes  D[pixel, position] = g[pixel,:] . f[position,:]
// for all 9 x 100 pairs with mma.sync.m16n8k16 (fp16 in, fp32 accumulate): the A operand (the patch,
// 16x128 with rows 9..15 zero) lives in 32 registers for the whole edge, the B operand (window
// positions x channels) is loaded straight from the channels-last map with 16-byte vector loads
// (a K permutation shared by A and B makes every fragment a contiguous 16 B run).  D goes to shared
// memory, the 4-corner bilinear blend, the (x,y) transpose and the pyramid-level interleave are
// fused in the epilogue, and the [E, 882] row the update operator consumes is written with
// coalesced 4-byte (level0, level1) stores.  Nothing but coords, indices and the output touches HBM
// besides the feature windows themselves.  Bound: HBM/L2 (DESIGN.md "corr").
//
// A strided CUDA-core kernel covers every other layout / dtype / patch size (same arithmetic,
// fp32 accumulate), so the reference's NCHW call signature keeps working.
#include <stdlib.h>

#include "common.cuh"

namespace rvo {

// ------------------------------------------------------------------ patchify ----

template <typename T>
__device__ __forceinline__ float to_f32(T v);
template <>
__device__ __forceinline__ float to_f32<float>(float v) { return v; }
template <>
__device__ __forceinline__ float to_f32<__half>(__half v) { return __half2float(v); }

template <typename T>
__device__ __forceinline__ T from_f32(float v);
template <>
__device__ __forceinline__ float from_f32<float>(float v) { return v; }
template <>
__device__ __forceinline__ __half from_f32<__half>(float v) { return __float2half_rn(v); }

__device__ __forceinline__ int floor_i(float v) {
  // floor with a clamp that keeps the int conversion defined; anything this far out is
  // outside every feature map anyway.  NaN maps to the low clamp (out of bounds -> zeros).
  float f = floorf(v);
  if (!(f > -1.0e6f)) f = -1.0e6f;
  if (f > 1.0e6f) f = 1.0e6f;
  return (int)f;
}

struct FmapView {
  const void* data;
  int N, C, H, W;
  int64_t sN, sC, sH, sW;
};

static inline FmapView view_of(const rvo_fmap_t* f) {
  FmapView v;
  v.data = f->data; v.N = f->N; v.C = f->C; v.H = f->H; v.W = f->W;
  v.sN = f->sN; v.sC = f->sC; v.sH = f->sH; v.sW = f->sW;
  return v;
}

// raw gather, out dense [B,M,C,D,D].  c_fast selects the thread->element order so that the
// global loads coalesce for channels-last (c fastest) or NCHW (window column fastest) maps.
template <typename T>
__global__ void __launch_bounds__(256)
patchify_raw_kernel(FmapView net, const float* __restrict__ coords, int M, int R, bool c_fast,
                    T* __restrict__ out) {
  const int D = 2 * R + 2;
  const int64_t total = (int64_t)net.N * M * net.C * D * D;
  const T* src = reinterpret_cast<const T*>(net.data);
  for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < total;
       t += (int64_t)gridDim.x * blockDim.x) {
    int64_t r = t;
    int c, a, b;
    if (c_fast) {
      c = (int)(r % net.C); r /= net.C;
      b = (int)(r % D); r /= D;
      a = (int)(r % D); r /= D;
    } else {
      b = (int)(r % D); r /= D;
      a = (int)(r % D); r /= D;
      c = (int)(r % net.C); r /= net.C;
    }
    const int m = (int)(r % M);
    const int n = (int)(r / M);
    const float x = coords[((int64_t)n * M + m) * 2 + 0];
    const float y = coords[((int64_t)n * M + m) * 2 + 1];
    const int i = floor_i(y) + (a - R);
    const int j = floor_i(x) + (b - R);
    T v = from_f32<T>(0.0f);
    if (i >= 0 && i < net.H && j >= 0 && j < net.W)
      v = src[n * net.sN + c * net.sC + i * net.sH + j * net.sW];
    out[((((int64_t)n * M + m) * net.C + c) * D + a) * D + b] = v;
  }
}

template <typename T, typename TO>
__global__ void __launch_bounds__(256)
patchify_bilinear_kernel(FmapView net, const float* __restrict__ coords, int M, int R, bool c_fast,
                         TO* __restrict__ out, int64_t oB, int64_t oM, int64_t oC, int64_t oH,
                         int64_t oW) {
  const int d = 2 * R + 1;
  const int64_t total = (int64_t)net.N * M * net.C * d * d;
  const T* src = reinterpret_cast<const T*>(net.data);
  for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < total;
       t += (int64_t)gridDim.x * blockDim.x) {
    int64_t r = t;
    int c, a, b;
    if (c_fast) {
      c = (int)(r % net.C); r /= net.C;
      b = (int)(r % d); r /= d;
      a = (int)(r % d); r /= d;
    } else {
      b = (int)(r % d); r /= d;
      a = (int)(r % d); r /= d;
      c = (int)(r % net.C); r /= net.C;
    }
    const int m = (int)(r % M);
    const int n = (int)(r / M);
    const float x = coords[((int64_t)n * M + m) * 2 + 0];
    const float y = coords[((int64_t)n * M + m) * 2 + 1];
    // correlation.py:57: offset = coords - coords.floor()
    const float dx = __fsub_rn(x, floorf(x));
    const float dy = __fsub_rn(y, floorf(y));
    const int i0 = floor_i(y) + (a - R);
    const int j0 = floor_i(x) + (b - R);
    float p[2][2];
#pragma unroll
    for (int u = 0; u < 2; u++)
#pragma unroll
      for (int v = 0; v < 2; v++) {
        const int i = i0 + u, j = j0 + v;
        p[u][v] = 0.0f;
        if (i >= 0 && i < net.H && j >= 0 && j < net.W)
          p[u][v] = to_f32<T>(src[n * net.sN + c * net.sC + i * net.sH + j * net.sW]);
      }
    // correlation.py:61-66, same association and no FMA contraction => bit-exact in fp32
    const float omx = __fsub_rn(1.0f, dx), omy = __fsub_rn(1.0f, dy);
    const float x00 = __fmul_rn(__fmul_rn(omy, omx), p[0][0]);
    const float x01 = __fmul_rn(__fmul_rn(omy, dx), p[0][1]);
    const float x10 = __fmul_rn(__fmul_rn(dy, omx), p[1][0]);
    const float x11 = __fmul_rn(__fmul_rn(dy, dx), p[1][1]);
    const float s = __fadd_rn(__fadd_rn(__fadd_rn(x00, x01), x10), x11);
    out[n * oB + m * oM + c * oC + a * oH + b * oW] = from_f32<TO>(s);
  }
}

// ------------------------------------------------------------------ corr: generic ----

constexpr int kMaxLevels = 2;

struct CorrLevels {
  FmapView f[kMaxLevels];
  float scale[kMaxLevels];
  int n;
};

// One warp per (edge, patch pixel, level): raw D x D window in shared memory, then the blend.
// Any strides / dtype / C / P / R <= 7.
template <typename T>
__global__ void __launch_bounds__(256)
corr_generic_kernel(FmapView g, CorrLevels L, const float* __restrict__ coords,
                    const int64_t* __restrict__ ii, const int64_t* __restrict__ jj, int64_t pmod,
                    int64_t fmod, int E, int R, T* __restrict__ out, int64_t out_ld) {
  extern __shared__ float smem[];
  const int D = 2 * R + 2, d = D - 1;
  const int P = g.H, PP = P * P;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float* win = smem + warp * (D * D);
  const int64_t nwork = (int64_t)E * PP * L.n;
  const T* g1 = reinterpret_cast<const T*>(g.data);
  for (int64_t w = (int64_t)blockIdx.x * (blockDim.x >> 5) + warp; w < nwork;
       w += (int64_t)gridDim.x * (blockDim.x >> 5)) {
    const int lvl = (int)(w % L.n);
    const int64_t ep = w / L.n;
    const int pix = (int)(ep % PP);
    const int e = (int)(ep / PP);
    const int i0 = pix / P, j0 = pix % P;
    int64_t ip = ii[e], jf = jj[e];
    if (pmod > 0) ip %= pmod;
    if (fmod > 0) jf %= fmod;
    const FmapView& f = L.f[lvl];
    const T* f2 = reinterpret_cast<const T*>(f.data);
    const float x = coords[((int64_t)e * 2 + 0) * PP + pix] * L.scale[lvl];
    const float y = coords[((int64_t)e * 2 + 1) * PP + pix] * L.scale[lvl];
    const int fx = floor_i(x), fy = floor_i(y);
    const bool idx_ok = ip >= 0 && ip < g.N && jf >= 0 && jf < f.N;
    for (int q = lane; q < D * D; q += 32) {
      const int a = q / D, b = q % D;
      const int i1 = fy + a - R, j1 = fx + b - R;
      float s = 0.0f;
      if (idx_ok && i1 >= 0 && i1 < f.H && j1 >= 0 && j1 < f.W) {
        const T* pa = g1 + ip * g.sN + i0 * g.sH + j0 * g.sW;
        const T* pb = f2 + jf * f.sN + i1 * f.sH + j1 * f.sW;
        for (int c = 0; c < g.C; c++) s = fmaf(to_f32<T>(pa[c * g.sC]), to_f32<T>(pb[c * f.sC]), s);
      }
      win[q] = s;
    }
    __syncwarp();
    const float dx = x - floorf(x), dy = y - floorf(y);
    for (int q = lane; q < d * d; q += 32) {
      const int b = q / d, a = q % d;  // output is [x offset][y offset]
      const float v = ((1.0f - dx) * (1.0f - dy)) * win[a * D + b] +
                      (dx * (1.0f - dy)) * win[a * D + b + 1] +
                      ((1.0f - dx) * dy) * win[(a + 1) * D + b] +
                      (dx * dy) * win[(a + 1) * D + b + 1];
      out[(int64_t)e * out_ld + ((int64_t)q * PP + pix) * L.n + lvl] = from_f32<T>(v);
    }
    __syncwarp();
  }
}

// ------------------------------------------------------------------ corr: tensor-core path ----

constexpr int kCorrR = 3;           // lookup radius of the fast path
constexpr int kWin = 2 * kCorrR + 2;  // 8: raw window per pixel
constexpr int kUni = kWin + 2;        // 10: union window (pixel floors differ by <= 2)
constexpr int kPosPad = 104;          // 13 n-tiles of 8 positions
constexpr int kWarpsPerCta = 8;

__device__ __forceinline__ void mma16816(float* c, const uint32_t* a, uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, "
      "{%0,%1,%2,%3};\n"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

__device__ __forceinline__ uint4 ldg_nc_v4(const void* p) {
  uint4 r;
  asm volatile("ld.global.nc.v4.u32 {%0,%1,%2,%3}, [%4];\n"
               : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
               : "l"(p));
  return r;
}

// D rows [row_lo, row_hi] of  A(16 x C) * window(C x ntiles*8)  -> Ds[row][wy*kUni + wx]
// window = WU x WU positions at (y0.., x0..) of frame map `fbase` ([H,W,C] channels-last).
// WU is a compile-time constant so that pos -> (wy, wx) is a multiply-shift, not a division.
template <int C, int WU, int NTB>
__device__ __forceinline__ void window_mma(const uint32_t (&afrag)[C / 16][4],
                                           const __half* __restrict__ fbase, int H, int W,
                                           int sH, int sW, int y0, int x0, int row_lo,
                                           int row_hi, float* __restrict__ Ds, int lane) {
  const int g = lane >> 2, t = lane & 3;
  constexpr int npos = WU * WU;
  constexpr int ntiles = (npos + 7) >> 3;
#pragma unroll 1
  for (int nt = 0; nt < ntiles; nt += NTB) {
    uint4 bq[NTB][C / 32];
#pragma unroll
    for (int u = 0; u < NTB; u++) {
      const int pos = (nt + u) * 8 + g;
      const int wy = pos / WU, wx = pos - wy * WU;
      const int y = y0 + wy, x = x0 + wx;
      const bool ok = pos < npos && (unsigned)y < (unsigned)H && (unsigned)x < (unsigned)W;
      const __half* p = fbase + (ok ? (y * sH + x * sW) : 0) + t * 8;
#pragma unroll
      for (int q = 0; q < C / 32; q++) {
        bq[u][q] = ok ? ldg_nc_v4(p + q * 32) : make_uint4(0u, 0u, 0u, 0u);
      }
    }
#pragma unroll
    for (int u = 0; u < NTB; u++) {
      if (nt + u >= ntiles) break;
      float c[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
      for (int q = 0; q < C / 32; q++) {
        mma16816(c, afrag[2 * q], bq[u][q].x, bq[u][q].y);
        mma16816(c, afrag[2 * q + 1], bq[u][q].z, bq[u][q].w);
      }
      // c0,c1: row g, positions 2t,2t+1 of this tile; c2,c3: row g+8
      const int pos = (nt + u) * 8 + 2 * t;
      if (WU == kUni) {
        // row-major 10-wide window: the smem column IS the position index; npos is even
        if (pos < npos) {
          if (g >= row_lo && g <= row_hi)
            *reinterpret_cast<float2*>(Ds + g * kPosPad + pos) = make_float2(c[0], c[1]);
          if (g == 0 && 8 >= row_lo && 8 <= row_hi)
            *reinterpret_cast<float2*>(Ds + 8 * kPosPad + pos) = make_float2(c[2], c[3]);
        }
      } else {
#pragma unroll
        for (int h = 0; h < 2; h++) {
          const int pp = pos + h;
          if (pp < npos) {
            const int wy = pp / WU, wx = pp - wy * WU;
            if (g >= row_lo && g <= row_hi) Ds[g * kPosPad + wy * kUni + wx] = c[h];
            if (g + 8 >= row_lo && g + 8 <= row_hi && g + 8 < 9)
              Ds[(g + 8) * kPosPad + wy * kUni + wx] = c[2 + h];
          }
        }
      }
    }
  }
}

template <int C, int NL, int NTB, int OCC>
__global__ void __launch_bounds__(kWarpsPerCta * 32, OCC)
corr_mma_kernel(const __half* __restrict__ gmap, int64_t gN, int64_t g_sN, int64_t g_sH,
                int64_t g_sW, CorrLevels L, const float* __restrict__ coords,
                const int64_t* __restrict__ kk, const int64_t* __restrict__ jj, int64_t pmod,
                int64_t fmod, int E, __half* __restrict__ out, int64_t out_ld) {
  __shared__ __align__(16) float Dsm[kWarpsPerCta][9 * kPosPad];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int g = lane >> 2, t = lane & 3;
  float* Ds = Dsm[warp];
  constexpr int NOUT = 49 * 9;                     // 441 outputs per level
  constexpr int NIT = (NOUT + 31) / 32;            // 14

  // output o = (b*7 + a)*9 + pix (b: x offset, a: y offset; correlation_kernel.cu:227-232).
  // The lane's 14 (pix, a*kUni + b) pairs never change: pack them once for the persistent loop.
  uint32_t tab[NIT / 2];
#pragma unroll
  for (int it = 0; it < NIT; it++) {
    const int o = it * 32 + lane;
    const int oo = o < NOUT ? o : NOUT - 1;
    const int pix = oo % 9, ab = oo / 9;
    const int b = ab / 7, a = ab - b * 7;
    const uint32_t v = (uint32_t)pix | ((uint32_t)(a * kUni + b) << 4);   // 4 + 7 bits
    if (it & 1) tab[it >> 1] |= v << 16; else tab[it >> 1] = v;
  }

  for (int e = blockIdx.x * kWarpsPerCta + warp; e < E; e += gridDim.x * kWarpsPerCta) {
    int64_t ip = kk[e], jf = jj[e];
    if (pmod > 0) ip %= pmod;
    if (fmod > 0) jf %= fmod;
    const bool p_ok = ip >= 0 && ip < gN;

    // A fragments: patch pixel rows (i0*3+j0), K permuted so that thread t of a quad owns the
    // 8 contiguous channels [32q + 8t, 32q + 8t + 8) of every 32-channel chunk q.
    uint32_t afrag[C / 16][4];
    {
      const __half* prow = gmap + (p_ok ? ip * g_sN : 0) + (g / 3) * g_sH + (g % 3) * g_sW + t * 8;
      const __half* prow8 = gmap + (p_ok ? ip * g_sN : 0) + 2 * g_sH + 2 * g_sW + t * 8;
#pragma unroll
      for (int q = 0; q < C / 32; q++) {
        uint4 lo = p_ok ? ldg_nc_v4(prow + q * 32) : make_uint4(0u, 0u, 0u, 0u);
        uint4 hi = (p_ok && g == 0) ? ldg_nc_v4(prow8 + q * 32) : make_uint4(0u, 0u, 0u, 0u);
        afrag[2 * q][0] = lo.x; afrag[2 * q][2] = lo.y;
        afrag[2 * q + 1][0] = lo.z; afrag[2 * q + 1][2] = lo.w;
        afrag[2 * q][1] = hi.x; afrag[2 * q][3] = hi.y;
        afrag[2 * q + 1][1] = hi.z; afrag[2 * q + 1][3] = hi.w;
      }
    }

    // lane p < 9 owns patch pixel p
    float cx = 0.f, cy = 0.f;
    if (lane < 9) {
      cx = coords[(int64_t)e * 18 + lane];
      cy = coords[(int64_t)e * 18 + 9 + lane];
    }

    __half2 acc0[NIT / 2];   // level-0 results wait here (packed fp16) for their level-1 partner
#pragma unroll
    for (int lvl = 0; lvl < NL; lvl++) {
      const FmapView& f = L.f[lvl];
      const bool f_ok = jf >= 0 && jf < f.N;
      const __half* fbase = reinterpret_cast<const __half*>(f.data) + (f_ok ? jf * f.sN : 0);
      const int Hh = f_ok ? f.H : 0;  // an invalid frame index reads as all-out-of-bounds
      const float x = cx * L.scale[lvl], y = cy * L.scale[lvl];
      const int fx = floor_i(x), fy = floor_i(y);
      const float dx = x - floorf(x), dy = y - floorf(y);
      int mnx = (lane < 9) ? fx : 0x7fffffff, mny = (lane < 9) ? fy : 0x7fffffff;
      int mxx = (lane < 9) ? fx : (int)0x80000000, mxy = (lane < 9) ? fy : (int)0x80000000;
#pragma unroll
      for (int o = 8; o > 0; o >>= 1) {   // lanes 0..15 cover the 9 pixels
        mnx = min(mnx, __shfl_xor_sync(0xffffffffu, mnx, o));
        mny = min(mny, __shfl_xor_sync(0xffffffffu, mny, o));
        mxx = max(mxx, __shfl_xor_sync(0xffffffffu, mxx, o));
        mxy = max(mxy, __shfl_xor_sync(0xffffffffu, mxy, o));
      }
      mnx = __shfl_sync(0xffffffffu, mnx, 0); mny = __shfl_sync(0xffffffffu, mny, 0);
      mxx = __shfl_sync(0xffffffffu, mxx, 0); mxy = __shfl_sync(0xffffffffu, mxy, 0);
      int pbase;  // this pixel's window origin inside the staged window, as a Ds offset
      if (mxx - mnx <= kUni - kWin && mxy - mny <= kUni - kWin) {
        window_mma<C, kUni, NTB>(afrag, fbase, Hh, f.W, (int)f.sH, (int)f.sW, mny - kCorrR,
                            mnx - kCorrR, 0, 8, Ds, lane);
        pbase = lane * kPosPad + (fy - mny) * kUni + (fx - mnx);
      } else {
        // spread-out patch (strong zoom / rotation): one private 8x8 window per pixel
        for (int p = 0; p < 9; p++) {
          const int pfx = __shfl_sync(0xffffffffu, fx, p), pfy = __shfl_sync(0xffffffffu, fy, p);
          window_mma<C, kWin, NTB>(afrag, fbase, Hh, f.W, (int)f.sH, (int)f.sW, pfy - kCorrR,
                              pfx - kCorrR, p, p, Ds, lane);
        }
        pbase = lane * kPosPad;
      }
      __syncwarp();
      // bilinear blend (correlation_kernel.cu:221-230) as two lerps, fp32
#pragma unroll
      for (int it = 0; it < NIT; it++) {
        const uint32_t tv = (tab[it >> 1] >> ((it & 1) * 16)) & 0xffffu;
        const int pix = tv & 15, off = tv >> 4;
        const int base = __shfl_sync(0xffffffffu, pbase, pix);
        const float wx = __shfl_sync(0xffffffffu, dx, pix), wy = __shfl_sync(0xffffffffu, dy, pix);
        const float* dp = Ds + base + off;
        const float c00 = dp[0], c01 = dp[1], c10 = dp[kUni], c11 = dp[kUni + 1];
        const float top = c00 + wx * (c01 - c00);
        const float bot = c10 + wx * (c11 - c10);
        const float v = top + wy * (bot - top);
        if (NL == 2 && lvl == 0) {
          if (it & 1) acc0[it >> 1].y = __float2half_rn(v); else acc0[it >> 1].x = __float2half_rn(v);
        } else {
          const int o = it * 32 + lane;
          if (o < NOUT) {
            if (NL == 2)
              reinterpret_cast<__half2*>(out + (int64_t)e * out_ld)[o] =
                  __halves2half2((it & 1) ? acc0[it >> 1].y : acc0[it >> 1].x, __float2half_rn(v));
            else
              out[(int64_t)e * out_ld + o] = __float2half_rn(v);
          }
        }
      }
      __syncwarp();
    }
  }
}

static bool fast_path_ok(const FmapView& g, const CorrLevels& L, int dtype, int R, int E) {
  if (dtype != RVO_F16 || R != kCorrR || g.H != 3 || g.W != 3) return false;
  if (g.C != 128 || g.sC != 1) return false;
  if (L.n < 1 || L.n > 2) return false;
  auto al16 = [](const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; };
  if (!al16(g.data) || (g.sN % 8) || (g.sH % 8) || (g.sW % 8)) return false;
  const int64_t i32max = 0x7fffffff;
  for (int l = 0; l < L.n; l++) {
    const FmapView& f = L.f[l];
    if (f.C != g.C || f.sC != 1 || !al16(f.data) || (f.sN % 8) || (f.sH % 8) || (f.sW % 8))
      return false;
    if (f.sH < 0 || f.sW < 0 || f.H * f.sH + f.W * f.sW > i32max) return false;  // 32-bit in-frame offsets
  }
  (void)E;
  return true;
}

static int corr_launch(const rvo_fmap_t* fmap1, const rvo_fmap_t* pyr, const float* scale,
                       int nlevels, const float* coords, const int64_t* kk, const int64_t* jj,
                       int64_t pmod, int64_t fmod, int E, int radius, void* out, int64_t out_ld,
                       cudaStream_t st, const char* who) {
  RVO_CHECK_ARG(E >= 0, "%s: E=%d", who, E);
  RVO_CHECK_ARG(fmap1 && pyr && nlevels >= 1 && nlevels <= kMaxLevels, "%s: bad levels", who);
  RVO_CHECK_ARG(radius >= 0 && radius <= 7, "%s: radius %d unsupported", who, radius);
  if (E == 0) return RVO_OK;
  RVO_CHECK_ARG(coords && kk && jj && out && fmap1->data, "%s: null pointer", who);
  {
    const int64_t row = (int64_t)(2 * radius + 1) * (2 * radius + 1) * fmap1->H * fmap1->W * nlevels;
    if (out_ld <= 0) out_ld = row;
    RVO_CHECK_ARG(out_ld >= row, "%s: output row stride %lld < %lld", who, (long long)out_ld,
                  (long long)row);
  }
  RVO_CHECK_ARG(fmap1->H == fmap1->W && fmap1->H >= 1 && fmap1->H <= 9, "%s: patch size", who);
  FmapView g = view_of(fmap1);
  CorrLevels L;
  L.n = nlevels;
  for (int l = 0; l < nlevels; l++) {
    RVO_CHECK_ARG(pyr[l].data && pyr[l].dtype == fmap1->dtype && pyr[l].C == fmap1->C,
                  "%s: level %d dtype/channels differ from fmap1", who, l);
    L.f[l] = view_of(&pyr[l]);
    L.scale[l] = scale ? scale[l] : 1.0f;
  }
  for (int l = nlevels; l < kMaxLevels; l++) { L.f[l] = L.f[0]; L.scale[l] = 1.0f; }

  if (fast_path_ok(g, L, fmap1->dtype, radius, E) &&
      (nlevels == 1 || (out_ld % 2 == 0 && (reinterpret_cast<uintptr_t>(out) & 3u) == 0))) {
    // development switch (tools/microbench.py): RVO_CORR_VARIANT = 10*NTB + OCC
    static const int variant = getenv("RVO_CORR_VARIANT") ? atoi(getenv("RVO_CORR_VARIANT")) : 22;
    const int occ = variant % 10;
    int grid = cdiv(E, kWarpsPerCta);
    if (grid > sm_budget() * occ) grid = sm_budget() * occ;   // persistent: `occ` resident CTAs per SM
#define RVO_CORR_LAUNCH(NL_, NTB_, OCC_)                                                       \
    corr_mma_kernel<128, NL_, NTB_, OCC_><<<grid, kWarpsPerCta * 32, 0, st>>>(                  \
        (const __half*)g.data, g.N, g.sN, g.sH, g.sW, L, coords, kk, jj, pmod, fmod, E,         \
        (__half*)out, out_ld)
#define RVO_CORR_PICK(NL_)                                                                     \
    switch (variant) {                                                                         \
      case 12: RVO_CORR_LAUNCH(NL_, 1, 2); break;                                              \
      case 13: RVO_CORR_LAUNCH(NL_, 1, 3); break;                                              \
      case 14: RVO_CORR_LAUNCH(NL_, 1, 4); break;                                              \
      case 22: RVO_CORR_LAUNCH(NL_, 2, 2); break;                                              \
      case 24: RVO_CORR_LAUNCH(NL_, 2, 4); break;                                              \
      case 23: RVO_CORR_LAUNCH(NL_, 2, 3); break;                                              \
      default: RVO_CORR_LAUNCH(NL_, 2, 2); break;                                              \
    }
    if (nlevels == 2) { RVO_CORR_PICK(2) } else { RVO_CORR_PICK(1) }
#undef RVO_CORR_PICK
#undef RVO_CORR_LAUNCH
    RVO_LAUNCH_CHECK("corr_mma_kernel");
    return RVO_OK;
  }
  const int D = 2 * radius + 2;
  const size_t smem = (size_t)8 * D * D * sizeof(float);
  const int64_t nwork = (int64_t)E * g.H * g.W * nlevels;
  int64_t grid = (nwork + 7) / 8;
  if (grid > (int64_t)sm_budget() * 64) grid = (int64_t)sm_budget() * 64;
  if (fmap1->dtype == RVO_F16)
    corr_generic_kernel<__half><<<(int)grid, 256, smem, st>>>(g, L, coords, kk, jj, pmod, fmod, E,
                                                              radius, (__half*)out, out_ld);
  else if (fmap1->dtype == RVO_F32)
    corr_generic_kernel<float><<<(int)grid, 256, smem, st>>>(g, L, coords, kk, jj, pmod, fmod, E,
                                                             radius, (float*)out, out_ld);
  else
    RVO_CHECK_ARG(false, "%s: dtype %d unsupported", who, fmap1->dtype);
  RVO_LAUNCH_CHECK("corr_generic_kernel");
  return RVO_OK;
}

}  // namespace rvo

using namespace rvo;

extern "C" int rvo_patchify_forward(const rvo_fmap_t* net, const float* coords, int M, int radius,
                                    void* out, void* stream) {
  RVO_CHECK_ARG(net && net->data && coords && out, "rvo_patchify_forward: null pointer");
  RVO_CHECK_ARG(M >= 0 && radius >= 0 && radius <= 15, "rvo_patchify_forward: M=%d R=%d", M, radius);
  if (M == 0 || net->N == 0) return RVO_OK;
  FmapView v = view_of(net);
  const int D = 2 * radius + 2;
  const int64_t total = (int64_t)v.N * M * v.C * D * D;
  int64_t grid = (total + 255) / 256;
  if (grid > (int64_t)sm_budget() * 32) grid = (int64_t)sm_budget() * 32;
  const bool c_fast = (v.sC == 1);
  cudaStream_t st = (cudaStream_t)stream;
  if (net->dtype == RVO_F16)
    patchify_raw_kernel<__half><<<(int)grid, 256, 0, st>>>(v, coords, M, radius, c_fast, (__half*)out);
  else if (net->dtype == RVO_F32)
    patchify_raw_kernel<float><<<(int)grid, 256, 0, st>>>(v, coords, M, radius, c_fast, (float*)out);
  else
    RVO_CHECK_ARG(false, "rvo_patchify_forward: dtype %d unsupported", net->dtype);
  RVO_LAUNCH_CHECK("patchify_raw_kernel");
  return RVO_OK;
}

extern "C" int rvo_patchify_bilinear(const rvo_fmap_t* net, const float* coords, int M, int radius,
                                     void* out, int out_dtype, int64_t oB, int64_t oM, int64_t oC,
                                     int64_t oH, int64_t oW, void* stream) {
  RVO_CHECK_ARG(net && net->data && coords && out, "rvo_patchify_bilinear: null pointer");
  RVO_CHECK_ARG(M >= 0 && radius >= 0 && radius <= 15, "rvo_patchify_bilinear: M=%d R=%d", M, radius);
  if (M == 0 || net->N == 0) return RVO_OK;
  FmapView v = view_of(net);
  const int d = 2 * radius + 1;
  const int64_t total = (int64_t)v.N * M * v.C * d * d;
  int64_t grid = (total + 255) / 256;
  if (grid > (int64_t)sm_budget() * 32) grid = (int64_t)sm_budget() * 32;
  const bool c_fast = (v.sC == 1);
  cudaStream_t st = (cudaStream_t)stream;
#define RVO_PB(TI, TO)                                                                          \
  patchify_bilinear_kernel<TI, TO><<<(int)grid, 256, 0, st>>>(v, coords, M, radius, c_fast,     \
                                                              (TO*)out, oB, oM, oC, oH, oW)
  if (net->dtype == RVO_F16 && out_dtype == RVO_F16) RVO_PB(__half, __half);
  else if (net->dtype == RVO_F16 && out_dtype == RVO_F32) RVO_PB(__half, float);
  else if (net->dtype == RVO_F32 && out_dtype == RVO_F32) RVO_PB(float, float);
  else if (net->dtype == RVO_F32 && out_dtype == RVO_F16) RVO_PB(float, __half);
  else RVO_CHECK_ARG(false, "rvo_patchify_bilinear: dtype %d -> %d unsupported", net->dtype, out_dtype);
#undef RVO_PB
  RVO_LAUNCH_CHECK("patchify_bilinear_kernel");
  return RVO_OK;
}

extern "C" int rvo_corr_forward(const rvo_fmap_t* fmap1, const rvo_fmap_t* fmap2,
                                const float* coords, const int64_t* ii, const int64_t* jj, int E,
                                int radius, void* out, void* stream) {
  const float one = 1.0f;
  return corr_launch(fmap1, fmap2, &one, 1, coords, ii, jj, 0, 0, E, radius, out, 0,
                     (cudaStream_t)stream, "rvo_corr_forward");
}

extern "C" int rvo_corr_pyramid(const rvo_fmap_t* fmap1, const rvo_fmap_t* pyr, const float* scale,
                                int nlevels, const float* coords, const int64_t* kk,
                                const int64_t* jj, int64_t pmod, int64_t fmod, int E, int radius,
                                void* out, int64_t out_ld, void* stream) {
  return corr_launch(fmap1, pyr, scale, nlevels, coords, kk, jj, pmod, fmod, E, radius, out, out_ld,
                     (cudaStream_t)stream, "rvo_corr_pyramid");
}

// Host-buffer variant.  The views must describe DENSE host tensors (any stride order): the whole
// allocation [0, max offset] is copied.
static int64_t view_span_elems(const rvo_fmap_t* f) {
  return (f->N - 1) * f->sN + (f->C - 1) * f->sC + (f->H - 1) * f->sH + (f->W - 1) * f->sW + 1;
}

extern "C" int rvo_corr_pyramid_host(const rvo_fmap_t* fmap1, const rvo_fmap_t* pyr,
                                     const float* scale, int nlevels, const float* coords,
                                     const int64_t* kk, const int64_t* jj, int64_t pmod,
                                     int64_t fmod, int E, int radius, void* out, void* stream) {
  RVO_CHECK_ARG(fmap1 && pyr && nlevels >= 1 && nlevels <= kMaxLevels, "rvo_corr_pyramid_host: args");
  cudaStream_t st = (cudaStream_t)stream;
  const size_t es = fmap1->dtype == RVO_F16 ? 2 : 4;
  const int PP = fmap1->H * fmap1->W;
  const int d = 2 * radius + 1;
  void* dptr[1 + kMaxLevels] = {nullptr, nullptr, nullptr};
  float* d_coords = nullptr; int64_t *d_kk = nullptr, *d_jj = nullptr; void* d_out = nullptr;
  rvo_fmap_t v1 = *fmap1, vp[kMaxLevels];
  int rc = RVO_OK;
  auto fail = [&](cudaError_t e, const char* w) { rc = cuda_fail(e, w); };
  cudaError_t ce;
#define H2D(dst, src, bytes)                                                                    \
  if (rc == RVO_OK && (ce = cudaMalloc((void**)&dst, bytes ? bytes : 1)) != cudaSuccess) fail(ce, "cudaMalloc"); \
  if (rc == RVO_OK && (ce = cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, st)) != cudaSuccess) fail(ce, "H2D");
  H2D(dptr[0], fmap1->data, (size_t)view_span_elems(fmap1) * es);
  v1.data = dptr[0];
  for (int l = 0; l < nlevels; l++) {
    vp[l] = pyr[l];
    H2D(dptr[1 + l], pyr[l].data, (size_t)view_span_elems(&pyr[l]) * es);
    vp[l].data = dptr[1 + l];
  }
  H2D(d_coords, coords, (size_t)E * 2 * PP * sizeof(float));
  H2D(d_kk, kk, (size_t)E * sizeof(int64_t));
  H2D(d_jj, jj, (size_t)E * sizeof(int64_t));
#undef H2D
  const size_t out_bytes = (size_t)E * d * d * PP * nlevels * es;
  if (rc == RVO_OK && (ce = cudaMalloc(&d_out, out_bytes ? out_bytes : 1)) != cudaSuccess) fail(ce, "cudaMalloc");
  if (rc == RVO_OK)
    rc = corr_launch(&v1, vp, scale, nlevels, d_coords, d_kk, d_jj, pmod, fmod, E, radius, d_out, 0, st,
                     "rvo_corr_pyramid_host");
  if (rc == RVO_OK && (ce = cudaMemcpyAsync(out, d_out, out_bytes, cudaMemcpyDeviceToHost, st)) != cudaSuccess) fail(ce, "D2H");
  cudaError_t se = cudaStreamSynchronize(st);
  if (rc == RVO_OK && se != cudaSuccess) fail(se, "cudaStreamSynchronize");
  for (auto p : dptr) if (p) cudaFree(p);
  if (d_coords) cudaFree(d_coords);
  if (d_kk) cudaFree(d_kk);
  if (d_jj) cudaFree(d_jj);
  if (d_out) cudaFree(d_out);
  return rc;
}
