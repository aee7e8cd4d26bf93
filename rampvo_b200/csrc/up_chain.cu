// up_chain.cu — one row-local stretch of the update operator (ramp/net.py:69-90, ramp/blocks.py:15-50) as ONE
// persistent tcgen05 kernel, sm_100a.  See include/rampvo_b200.h (rvo_up_chain) for the contract.
//
// Why: run layer by layer (up_gemm.cu + update_ops.cu) every nn.Linear reads and writes its [E, 384] activations
// through HBM / L2 (69.6 MB per layer at E = 45 312) and every LayerNorm / residual / gate is one more pass over
// the fp32 hidden state: ~1.5 GB per update for ~230 MB of algorithmic traffic.  Here a CTA owns a tile of 128
// edge rows for a whole stretch of layers:
//   * the activation tile lives in shared memory (6 K blocks of [128 x 64] fp16 in the UMMA K-major SWIZZLE_128B
//     layout, 96 KB) and is rewritten IN PLACE by the epilogue of each layer — the next layer's A operand never
//     leaves the SM;
//   * the weights of the current layer stream through a 5-stage TMA ring of [192 x 64] boxes (24 KB each; all CTAs
//     read the same 288 KB per layer, an L2 hit after the first touch);
//   * one elected thread issues tcgen05.mma M128 N192 K16 pairs into a 384-column fp32 accumulator in TMEM;
//   * 16 worker warps (TMEM lane quadrant x 96-column part, thread = row) run the prologue (cp.async row gather, or
//     the SoftAgg expand + LayerNorm) and the epilogues: bias, fp16 rounding of the Linear output, ReLU, LayerNorm
//     (row statistics exchanged through shared memory; two passes over TMEM, the pre-norm value parked in TMEM by
//     tcgen05.st), residual adds against the fp32 hidden state, the GatedResidual tail and the two 384 -> 2 heads;
//   * the first layer of the correlation MLP (K = 1008) streams its A operand through the same six slots.
// MMA and epilogue of one tile do not overlap (the accumulator fills 384 of the 512 TMEM columns and the A tile is
// rewritten in place), but the weight ring keeps prefetching during the epilogue and the 148 CTAs are not in
// lockstep; the win is the traffic: per stretch the hidden state is read once and written once.
#include <type_traits>

#include "common.cuh"
#include "tcgen05.cuh"

namespace rvo {

constexpr int kCcC = 384;                          // width of the update operator
constexpr int kCcM = 128;                          // rows per tile = MMA M
constexpr int kCcNH = 192;                         // MMA N (two per K step cover the 384 outputs)
constexpr int kCcSlots = 6;                        // A-tile K blocks resident in shared memory
constexpr int kCcSlotBytes = kCcM * 128;           // 16 384
constexpr int kCcWStage = kCcNH * 128;             // 24 576
constexpr int kCcWStages = 5;
constexpr int kCcWorkers = 384;                    // 12 warps (warps 4..15); 512 threads in all = 128 registers each
constexpr int kCcThreads = 128 + kCcWorkers;       // warp 0: weight TMA, warp 1: MMA issuer, warps 2-3 idle
constexpr int kCcSmemBytes = kCcSlots * kCcSlotBytes + kCcWStages * kCcWStage + 1024;
constexpr int kCcParts = 3;                        // column parts per row (128 columns each)
constexpr int kCcPartCols = kCcC / kCcParts;       // 128
constexpr int kCcChunks = kCcPartCols / 16;        // 16-column chunks per thread

struct ChainMaps {
  TcTmap m[RVO_CHAIN_MAX_LAYERS];
  TcTmap out16;                                    // row-major fp16 [M, 384] output that equals the activation tile
};

__device__ __forceinline__ float round_h(float v) { return __half2float(__float2half_rn(v)); }

__device__ __forceinline__ void tmem_st16(uint32_t taddr, const float* v) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};\n" ::"r"(taddr),
      "r"(__float_as_uint(v[0])), "r"(__float_as_uint(v[1])), "r"(__float_as_uint(v[2])), "r"(__float_as_uint(v[3])),
      "r"(__float_as_uint(v[4])), "r"(__float_as_uint(v[5])), "r"(__float_as_uint(v[6])), "r"(__float_as_uint(v[7])),
      "r"(__float_as_uint(v[8])), "r"(__float_as_uint(v[9])), "r"(__float_as_uint(v[10])), "r"(__float_as_uint(v[11])),
      "r"(__float_as_uint(v[12])), "r"(__float_as_uint(v[13])), "r"(__float_as_uint(v[14])), "r"(__float_as_uint(v[15]))
      : "memory");
  asm volatile("tcgen05.wait::st.sync.aligned;\n" ::: "memory");
}

// 8 halves (one 16-byte chunk) <-> 8 floats, by value (no pointer punning: the chunks live in registers)
__device__ __forceinline__ void unpack2(uint32_t u, float& lo, float& hi) {
  lo = __half2float(__ushort_as_half((unsigned short)(u & 0xffffu)));
  hi = __half2float(__ushort_as_half((unsigned short)(u >> 16)));
}
__device__ __forceinline__ void unpack8(uint4 u, float* f) {
  unpack2(u.x, f[0], f[1]);
  unpack2(u.y, f[2], f[3]);
  unpack2(u.z, f[4], f[5]);
  unpack2(u.w, f[6], f[7]);
}
__device__ __forceinline__ uint32_t pack2(float lo, float hi) {
  return (uint32_t)__half_as_ushort(__float2half_rn(lo)) | ((uint32_t)__half_as_ushort(__float2half_rn(hi)) << 16);
}
__device__ __forceinline__ uint4 pack8(const float* f) {
  return make_uint4(pack2(f[0], f[1]), pack2(f[2], f[3]), pack2(f[4], f[5]), pack2(f[6], f[7]));
}

// v[0..16) += bias16[c0 .. c0+16)   (uniform address across the warp: one broadcast transaction per 16 bytes)
__device__ __forceinline__ void add_bias16(const __half* __restrict__ bias, int c0, float* v) {
  const uint4* p = reinterpret_cast<const uint4*>(bias + c0);
#pragma unroll
  for (int g = 0; g < 2; g++) {
    float b[8];
    unpack8(__ldg(p + g), b);
#pragma unroll
    for (int i = 0; i < 8; i++) v[8 * g + i] += b[i];
  }
}

// v = (v - mean) * rstd * gamma + beta over 16 columns (gamma / beta already offset; uniform addresses)
__device__ __forceinline__ void layer_norm16(float* v, float mean, float rstd, const float* __restrict__ gamma,
                                             const float* __restrict__ beta) {
#pragma unroll
  for (int g = 0; g < 4; g++) {
    const float4 ga = __ldg(reinterpret_cast<const float4*>(gamma) + g);
    const float4 be = __ldg(reinterpret_cast<const float4*>(beta) + g);
    v[4 * g] = (v[4 * g] - mean) * rstd * ga.x + be.x;
    v[4 * g + 1] = (v[4 * g + 1] - mean) * rstd * ga.y + be.y;
    v[4 * g + 2] = (v[4 * g + 2] - mean) * rstd * ga.z + be.z;
    v[4 * g + 3] = (v[4 * g + 3] - mean) * rstd * ga.w + be.w;
  }
}

// ---- tile-blocked private layouts.  In the epilogues a thread owns a ROW, so a row-major [E, 384] array makes every
// warp access touch 32 different lines.  Arrays that only the chains themselves read (the hidden state between
// stretches, the per-CTA residual and gate scratch) are therefore stored as [block][16-byte chunk][row 0..127]:
// the 32 rows of a warp are 32 consecutive 16-byte pieces — one fully coalesced 512-byte access.
__device__ __forceinline__ size_t blk32_off(int blk, int row, int c) {   // fp32, c % 4 == 0, in floats
  return (((size_t)blk * (kCcC / 4) + (size_t)(c >> 2)) * kCcM + (size_t)row) << 2;
}
__device__ __forceinline__ size_t blk16_off(int blk, int row, int c) {   // fp16, c % 8 == 0, in halves
  return (((size_t)blk * (kCcC / 8) + (size_t)(c >> 3)) * kCcM + (size_t)row) << 3;
}
__device__ __forceinline__ void load16_blk32(const float* __restrict__ base, int blk, int row, int c0, float* v) {
#pragma unroll
  for (int g = 0; g < 4; g++) {
    const float4 t = *reinterpret_cast<const float4*>(base + blk32_off(blk, row, c0 + 4 * g));
    v[4 * g] = t.x; v[4 * g + 1] = t.y; v[4 * g + 2] = t.z; v[4 * g + 3] = t.w;
  }
}
__device__ __forceinline__ void store16_blk32(float* __restrict__ base, int blk, int row, int c0, const float* v) {
#pragma unroll
  for (int g = 0; g < 4; g++)
    *reinterpret_cast<float4*>(base + blk32_off(blk, row, c0 + 4 * g)) =
        make_float4(v[4 * g], v[4 * g + 1], v[4 * g + 2], v[4 * g + 3]);
}
__device__ __forceinline__ void load16_blk16(const __half* __restrict__ base, int blk, int row, int c0, uint4* g) {
  g[0] = *reinterpret_cast<const uint4*>(base + blk16_off(blk, row, c0));
  g[1] = *reinterpret_cast<const uint4*>(base + blk16_off(blk, row, c0 + 8));
}
__device__ __forceinline__ void store16_blk16(__half* __restrict__ base, int blk, int row, int c0, const float* v) {
  *reinterpret_cast<uint4*>(base + blk16_off(blk, row, c0)) = pack8(v);
  *reinterpret_cast<uint4*>(base + blk16_off(blk, row, c0 + 8)) = pack8(v + 8);
}

// ---- row-major arrays of the interface (the incoming / outgoing hidden state, the SoftAgg f | g rows): 32-byte
// (full sector) accesses per thread
__device__ __forceinline__ void load16_row32(const float* __restrict__ p, float* v) {
#pragma unroll
  for (int g = 0; g < 2; g++)
    asm volatile("ld.global.v8.f32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];\n"
                 : "=f"(v[8 * g]), "=f"(v[8 * g + 1]), "=f"(v[8 * g + 2]), "=f"(v[8 * g + 3]), "=f"(v[8 * g + 4]),
                   "=f"(v[8 * g + 5]), "=f"(v[8 * g + 6]), "=f"(v[8 * g + 7])
                 : "l"(p + 8 * g));
}
__device__ __forceinline__ void store16_row32(float* __restrict__ p, const float* v) {
#pragma unroll
  for (int g = 0; g < 2; g++)
    asm volatile("st.global.v8.f32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};\n" ::"l"(p + 8 * g), "f"(v[8 * g]),
                 "f"(v[8 * g + 1]), "f"(v[8 * g + 2]), "f"(v[8 * g + 3]), "f"(v[8 * g + 4]), "f"(v[8 * g + 5]),
                 "f"(v[8 * g + 6]), "f"(v[8 * g + 7])
                 : "memory");
}
__device__ __forceinline__ void store16_row16(__half* __restrict__ p, const float* v) {
  const uint4 lo = pack8(v), hi = pack8(v + 8);
  asm volatile("st.global.v8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};\n" ::"l"(p), "r"(lo.x), "r"(lo.y),
               "r"(lo.z), "r"(lo.w), "r"(hi.x), "r"(hi.y), "r"(hi.z), "r"(hi.w)
               : "memory");
}

// the 16 columns [c0, c0+16) of row r of the activation tile (fp16, K-major SWIZZLE_128B K blocks)
__device__ __forceinline__ void store16_tile(uint32_t A_u, int r, int c0, const float* v) {
#pragma unroll
  for (int g = 0; g < 2; g++) {
    const int col = c0 + 8 * g;
    const uint32_t addr = A_u + (uint32_t)(col >> 6) * kCcSlotBytes + (uint32_t)r * 128u +
                          (uint32_t)((((col & 63) >> 3) ^ (r & 7)) << 4);
    const uint4 u = pack8(v + 8 * g);
    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};\n" ::"r"(addr), "r"(u.x), "r"(u.y), "r"(u.z), "r"(u.w)
                 : "memory");
  }
}

__device__ __forceinline__ void tma_store_2d(const void* tmap, uint32_t src, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];\n" ::"l"(
                   reinterpret_cast<uint64_t>(tmap)),
               "r"(src), "r"(c0), "r"(c1)
               : "memory");
}

// the CTAs of a cluster run the same tiles-per-CTA count and the same layer program, so every weight stage is needed
// by all of them at (nearly) the same time: each CTA fetches 1/csize of the stage and multicasts it to every CTA of
// the cluster — the L2 is read once per cluster instead of once per CTA (measured: a Linear layer's time follows the
// number of CTAs that stream the same 288 KB from L2, 1.9 / 3.9 / 6.0 us per tile at 37 / 74 / 148 readers)
__device__ __forceinline__ void tma_load_2d_mc(uint32_t dst, const void* tmap, int c0, int c1, uint64_t* bar,
                                               uint16_t mask) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1, {%3, "
      "%4}], [%2], %5;\n" ::"r"(dst),
      "l"(reinterpret_cast<uint64_t>(tmap)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "h"(mask)
      : "memory");
}
__device__ __forceinline__ void umma_commit_mc(uint64_t* b, uint16_t mask) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;\n" ::"r"(
                   smem_u32(b)),
               "h"(mask)
               : "memory");
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;\n" ::: "memory");
}

__device__ __forceinline__ void workers_sync() { asm volatile("bar.sync 1, %0;\n" ::"n"(kCcWorkers) : "memory"); }

// one 16-column chunk of global operands, prefetched one chunk ahead of its use: x = 16 fp32, g / h = 16 fp16 each
struct Pre {
  float4 x[4];
  uint4 g[2], h[2];
};
__device__ __forceinline__ void pre_x(const Pre& p, float* v) {
#pragma unroll
  for (int i = 0; i < 4; i++) { v[4 * i] = p.x[i].x; v[4 * i + 1] = p.x[i].y; v[4 * i + 2] = p.x[i].z; v[4 * i + 3] = p.x[i].w; }
}
__device__ __forceinline__ void pre_blk32(Pre& p, const float* __restrict__ base, int blk, int row, int c0) {
#pragma unroll
  for (int g = 0; g < 4; g++) p.x[g] = *reinterpret_cast<const float4*>(base + blk32_off(blk, row, c0 + 4 * g));
}
__device__ __forceinline__ void pre_row32(Pre& p, const float* __restrict__ q) {   // row-major: two 32-byte loads
#pragma unroll
  for (int g = 0; g < 2; g++) {
    float r0, r1, r2, r3, r4, r5, r6, r7;
    asm volatile("ld.global.v8.f32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];\n"
                 : "=f"(r0), "=f"(r1), "=f"(r2), "=f"(r3), "=f"(r4), "=f"(r5), "=f"(r6), "=f"(r7)
                 : "l"(q + 8 * g));
    p.x[2 * g] = make_float4(r0, r1, r2, r3);
    p.x[2 * g + 1] = make_float4(r4, r5, r6, r7);
  }
}

// chunks 0..kCcChunks-1 of this thread's part, chunk i+1 fetched while chunk i is processed; `first` = chunk 0,
// fetched by the caller before it waited for the accumulator
template <class Fetch, class Chunk>
__device__ __forceinline__ void pipelined_chunks(int cbase, Pre first, Fetch fetch, Chunk chunk) {
  Pre p0 = first, p1;
#pragma unroll 1
  for (int ci = 0; ci < kCcChunks; ci += 2) {
    const int c0 = cbase + 16 * ci;
    p1 = fetch(c0 + 16);
    chunk(c0, p0);
    if (ci + 2 < kCcChunks) p0 = fetch(c0 + 32);
    chunk(c0 + 16, p1);
  }
}

// TMEM column of output column c: the two N = 192 halves sit at columns 0 and 256
__device__ __forceinline__ uint32_t acc_col(int c) { return (uint32_t)(c + (c >= kCcNH ? 256 - kCcNH : 0)); }

// -DRVO_DEBUG builds: clock64 stamps of CTA 0 (MMA issuer: operands ready / last MMA issued; worker warp 4: accumulator
// ready / epilogue done), read back with rvo_up_chain_trace (tools/chain_bench.py --trace)
#ifdef RVO_DEBUG
__device__ long long g_cc_trace[4 * 64];
#define CC_TRACE(slot, i) do { if (blockIdx.x == 0 && (i) < 64) g_cc_trace[(slot) * 64 + (i)] = clock64(); } while (0)
#else
#define CC_TRACE(slot, i) do { } while (0)
#endif

__global__ void __launch_bounds__(kCcThreads, 1)
up_chain_kernel(const rvo_chain_t a, const __grid_constant__ ChainMaps maps) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ uint64_t wfull[kCcWStages], wempty[kCcWStages], afull[kCcSlots], aempty[kCcSlots], hfull, tfull;
  __shared__ uint32_t tmem_base_s;
  __shared__ float4 xch[kCcM][4];             // row statistics / head partial sums of the four column parts
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int M = a.M, n_layers = a.n_layers;
  const int n_tiles = (M + kCcM - 1) / kCcM;
  // every CTA of a cluster must consume the same weight stages: the same number of tiles for all (a tile index past
  // the last one is a phantom tile: zero rows in, nothing stored)
  const int n_iter = (n_tiles + (int)gridDim.x - 1) / (int)gridDim.x;
  uint32_t csize, crank;
  asm volatile("mov.u32 %0, %%cluster_nctaid.x;\n" : "=r"(csize));
  asm volatile("mov.u32 %0, %%cluster_ctarank;\n" : "=r"(crank));
  const uint16_t cmask = (uint16_t)((1u << csize) - 1u);

  if (tid == 0) {
    for (int s = 0; s < kCcWStages; s++) {
      mbar_init(&wfull[s], 1);
      mbar_init(&wempty[s], csize);                      // the MMA issuer of every CTA of the cluster releases a stage
    }
    for (int s = 0; s < kCcSlots; s++) {
      mbar_init(&afull[s], kCcWorkers);
      mbar_init(&aempty[s], 1);
    }
    mbar_init(&hfull, kCcWorkers);
    mbar_init(&tfull, 1);
    asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;\n" ::"r"(smem_u32(&tmem_base_s))
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
  __syncthreads();
  cluster_sync_all();                                     // remote barriers are initialised before anyone signals them
  asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
  const uint32_t tmem_base = tmem_base_s;
  const uint32_t A_u = smem_u32(smem), W_u = A_u + kCcSlots * kCcSlotBytes;
  const bool stream0 = a.prologue == RVO_CHAIN_PRO_ROWS;     // layer 0 takes its A operand through the slot ring

  if (warp == 0) {
    // ===== weight producer: (tile, layer, K block, N half) in the order the MMA issuer consumes them =====
    if (lane == 0) {
      int s = 0, ph = 0;
      const int slice_rows = kCcNH / (int)csize;           // this CTA's share of a [192 x 64] stage
      const uint32_t slice_off = crank * (uint32_t)(slice_rows * 128);
      for (int it = 0; it < n_iter; it++)
        for (int l = 0; l < n_layers; l++) {
          const int nkb = (a.layer[l].K + 63) >> 6;
          for (int kb = 0; kb < nkb; kb++)
#pragma unroll
            for (int nh = 0; nh < 2; nh++) {
              mbar_wait(&wempty[s], ph ^ 1);               // every CTA of the cluster has consumed this stage
              mbar_expect_tx(&wfull[s], (uint32_t)kCcWStage);
              if (csize == 1)
                tma_load_2d(W_u + s * kCcWStage, &maps.m[l], kb * 64, nh * kCcNH, &wfull[s]);
              else
                tma_load_2d_mc(W_u + s * kCcWStage + slice_off, &maps.m[l], kb * 64,
                               nh * kCcNH + (int)crank * slice_rows, &wfull[s], cmask);
              if (++s == kCcWStages) { s = 0; ph ^= 1; }
            }
        }
    }
  } else if (warp == 1) {
    // ===== MMA issuer =====
    int s = 0, ph = 0;
    uint32_t a_par = 0, h_ph = 0;
    constexpr uint32_t idesc = umma_idesc_f16(kCcM, kCcNH);
    int tr = 0;
    for (int it = 0; it < n_iter; it++)
      for (int l = 0; l < n_layers; l++, tr++) {
        const int nkb = (a.layer[l].K + 63) >> 6;
        const bool stream = stream0 && l == 0;
        if (!stream) {                                   // whole A tile written by the workers (prologue / epilogue)
          mbar_wait_spin(&hfull, h_ph);
          h_ph ^= 1;
          asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
          asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
        }
        if (lane == 0) CC_TRACE(0, tr);
        for (int kb = 0; kb < nkb; kb++) {
          const int slot = stream ? kb % kCcSlots : kb;
          if (stream) {
            mbar_wait_spin(&afull[slot], (a_par >> slot) & 1u);
            a_par ^= 1u << slot;
            asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
            asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
          }
#pragma unroll
          for (int nh = 0; nh < 2; nh++) {
            mbar_wait_spin(&wfull[s], ph);
            asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
            if (lane == 0) {
              const uint64_t da0 = umma_desc(A_u + slot * kCcSlotBytes), db0 = umma_desc(W_u + s * kCcWStage);
#pragma unroll
              for (int k = 0; k < 4; k++)
                umma_f16(tmem_base + nh * 256, da0 + (uint64_t)(2 * k), db0 + (uint64_t)(2 * k), idesc,
                         (kb | k) ? 1u : 0u);
              if (csize == 1) umma_commit(&wempty[s]);
              else umma_commit_mc(&wempty[s], cmask);
            }
            __syncwarp();
            if (++s == kCcWStages) { s = 0; ph ^= 1; }
          }
          if (stream && lane == 0) umma_commit(&aempty[slot]);
          __syncwarp();
        }
        if (lane == 0) {
          umma_commit(&tfull);
          CC_TRACE(1, tr);
        }
        __syncwarp();
      }
  } else if (warp >= 4) {
    // ===== workers: thread = (row, 128-column part), eight chunks of 16 columns each =====
    const int wtid = tid - 128, wwarp = wtid >> 5;
    const int q = warp & 3, part = wwarp >> 2;             // warps 4..15: every quadrant gets parts 0..2
    const int row = q * 32 + lane;
    const uint32_t trow = tmem_base + ((uint32_t)(q * 32) << 16);
    const int cbase = part * kCcPartCols;
    uint32_t e_par = (1u << kCcSlots) - 1u, t_ph = 0;
    const __half* a16 = reinterpret_cast<const __half*>(a.a16);
    const __half* hya = reinterpret_cast<const __half*>(a.hy_a);
    const __half* hyb = reinterpret_cast<const __half*>(a.hy_b);
    __half* scr16 = reinterpret_cast<__half*>(a.scratch16);
    const int cta = blockIdx.x;
    bool tile_store_pending = false;                       // a TMA store still reads the activation tile

    for (int it = 0; it < n_iter; it++) {
      const int t = blockIdx.x + it * gridDim.x;           // >= n_tiles: phantom tile
      const int e = t * kCcM + row;
      const bool valid = e < M;

      // ---------------- prologue ----------------
      if (stream0) {
        // thread owns 16-byte chunk `ch` of rows r0, r0 + 48 (and r0 + 96 when r0 < 32) of every K block
        const int ch = wtid & 7, r0 = wtid >> 3;
        const int nj = r0 < 32 ? 3 : 2;
        int64_t src[3];
#pragma unroll
        for (int j = 0; j < 3; j++) {
          const int ee = t * kCcM + r0 + 48 * j;
          src[j] = (j < nj && ee < M) ? (a.gather ? a.gather[ee] : (int64_t)ee) : -1;
        }
        const int K0 = a.layer[0].K, nkb = (K0 + 63) >> 6;
        for (int kb = 0; kb < nkb; kb++) {
          const int slot = kb % kCcSlots;
          mbar_wait(&aempty[slot], (e_par >> slot) & 1u);
          e_par ^= 1u << slot;
          const int col = kb * 64 + ch * 8;
          const uint32_t dst0 = A_u + slot * kCcSlotBytes + r0 * 128 + (uint32_t)((ch ^ (r0 & 7)) << 4);
#pragma unroll
          for (int j = 0; j < 3; j++) {                     // (r0 + 48 j) & 7 == r0 & 7
            if (j < nj) {
              const bool ok = src[j] >= 0 && col < K0;
              cp_async16(dst0 + j * (48 * 128), a16 + (ok ? src[j] * a.lda + col : 0), ok ? 16u : 0u);
            }
          }
          asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];\n" ::"r"(smem_u32(&afull[slot]))
                       : "memory");
        }
      } else {
        // SoftAgg expand (blocks.py:47-48, net.py:84-85) [+ the first LayerNorm of Update.gru, net.py:47]
        const bool with_ln = a.prologue == RVO_CHAIN_PRO_EXPAND_LN;
        const __half* ra = valid ? hya + (size_t)a.grp_a[e] * kCcC : nullptr;
        const __half* rb = (valid && hyb) ? hyb + (size_t)a.grp_b[e] * kCcC : nullptr;
        float s1 = 0.f, s2 = 0.f;
        auto fetch = [&](int c0) {
          Pre p;
          if (valid) {
            pre_blk32(p, a.x32, t, row, c0);
            p.g[0] = __ldg(reinterpret_cast<const uint4*>(ra + c0));
            p.g[1] = __ldg(reinterpret_cast<const uint4*>(ra + c0) + 1);
            if (rb) {
              p.h[0] = __ldg(reinterpret_cast<const uint4*>(rb + c0));
              p.h[1] = __ldg(reinterpret_cast<const uint4*>(rb + c0) + 1);
            }
          }
          return p;
        };
        auto chunk = [&](int c0, const Pre& p) {
          float v[16];
          if (valid) {
            pre_x(p, v);
#pragma unroll
            for (int hf = 0; hf < 2; hf++) {
              float h8[8];
              unpack8(p.g[hf], h8);
#pragma unroll
              for (int i = 0; i < 8; i++) v[8 * hf + i] += h8[i];
              if (rb) {
                unpack8(p.h[hf], h8);
#pragma unroll
                for (int i = 0; i < 8; i++) v[8 * hf + i] += h8[i];
              }
            }
          } else {
#pragma unroll
            for (int i = 0; i < 16; i++) v[i] = 0.f;
          }
          if (with_ln) {
#pragma unroll
            for (int i = 0; i < 16; i++) { s1 += v[i]; s2 += v[i] * v[i]; }
            tmem_st16(trow + acc_col(c0), v);
          } else {
            store16_tile(A_u, row, c0, v);
          }
        };
        pipelined_chunks(cbase, fetch(cbase), fetch, chunk);
        if (with_ln) {
          xch[row][part] = make_float4(s1, s2, 0.f, 0.f);
          workers_sync();
          float t1 = 0.f, t2 = 0.f;
#pragma unroll
          for (int p = 0; p < kCcParts; p++) {
            const float4 u = xch[row][p];
            t1 += u.x; t2 += u.y;
          }
          const float mean = t1 * (1.0f / kCcC);
          const float rstd = rsqrtf(fmaxf(t2 * (1.0f / kCcC) - mean * mean, 0.f) + 1e-3f);
#pragma unroll 2
          for (int ci = 0; ci < kCcChunks; ci++) {
            const int c0 = cbase + 16 * ci;
            float v[16];
            tmem_ld16(trow + acc_col(c0), v);
            layer_norm16(v, mean, rstd, a.pro_gamma + c0, a.pro_beta + c0);
            store16_blk32(a.scratch32, cta, row, c0, v);
            store16_tile(A_u, row, c0, v);
          }
          asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
        }
        asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
        mbar_arrive(&hfull);
      }

      // ---------------- layers ----------------
      for (int l = 0; l < n_layers; l++) {
        const rvo_chain_layer_t& L = a.layer[l];
        const __half* bias = reinterpret_cast<const __half*>(L.bias16);
        const bool has_next = l + 1 < n_layers;
        const int epi = L.epilogue;
        const bool writes_tile = epi == RVO_CHAIN_EPI_RELU || epi == RVO_CHAIN_EPI_LN_RELU ||
                                 epi == RVO_CHAIN_EPI_GATED_LN ||
                                 ((epi == RVO_CHAIN_EPI_RES || epi == RVO_CHAIN_EPI_ADD3_LN) && (has_next || a.out16));
        const bool tile_out16 = (epi == RVO_CHAIN_EPI_RES || epi == RVO_CHAIN_EPI_ADD3_LN) && a.out16;

        if (epi == RVO_CHAIN_EPI_RES) {
          // v = res32[e] + t ; out32[e] = v ; tile = half(v)  — the residual chunk is fetched one chunk ahead
          auto fetch = [&](int c0) {
            Pre p;
            if (valid) pre_blk32(p, a.res32, t, row, c0);
            return p;
          };
          const Pre first = fetch(cbase);
          mbar_wait(&tfull, t_ph);
          if (wtid == 0) CC_TRACE(2, it * n_layers + l);
          asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
          if (tile_store_pending && writes_tile) {
            if (wtid == 0) asm volatile("cp.async.bulk.wait_group.read 0;\n" ::: "memory");
            workers_sync();
            tile_store_pending = false;
          }
          auto chunk = [&](int c0, const Pre& p) {
            float v[16], x[16];
            tmem_ld16(trow + acc_col(c0), v);
            add_bias16(bias, c0, v);
            pre_x(p, x);
#pragma unroll
            for (int i = 0; i < 16; i++) v[i] = (valid ? x[i] : 0.f) + round_h(v[i]);
            if (valid) store16_blk32(a.out32, t, row, c0, v);
            if (writes_tile) store16_tile(A_u, row, c0, v);
          };
          pipelined_chunks(cbase, first, fetch, chunk);
        } else if (epi == RVO_CHAIN_EPI_GATED_LN || epi == RVO_CHAIN_EPI_GATED_HEADS) {
          // y = residual + half(gate * t); then LayerNorm (GATED_LN) or the heads (GATED_HEADS)
          const bool heads = epi == RVO_CHAIN_EPI_GATED_HEADS;
          auto fetch = [&](int c0) {
            Pre p;
            pre_blk32(p, a.scratch32, cta, row, c0);
            p.g[0] = *reinterpret_cast<const uint4*>(scr16 + blk16_off(cta, row, c0));
            p.g[1] = *reinterpret_cast<const uint4*>(scr16 + blk16_off(cta, row, c0 + 8));
            return p;
          };
          const Pre first = fetch(cbase);
          mbar_wait(&tfull, t_ph);
          if (wtid == 0) CC_TRACE(2, it * n_layers + l);
          asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
          if (tile_store_pending && writes_tile) {
            if (wtid == 0) asm volatile("cp.async.bulk.wait_group.read 0;\n" ::: "memory");
            workers_sync();
            tile_store_pending = false;
          }
          float s1 = 0.f, s2 = 0.f;
          float acc4[4] = {0.f, 0.f, 0.f, 0.f};
          float* o32 = heads ? a.out32 + (size_t)e * kCcC : nullptr;
          // one chunk: y = residual + half(gate * t), then the head partial sums (kHeads) or the row statistics
          auto chunk_t = [&](auto heads_c, int c0, const Pre& p) {
            constexpr bool kHeads = decltype(heads_c)::value;
            float v[16], x[16];
            tmem_ld16(trow + acc_col(c0), v);
            add_bias16(bias, c0, v);
            pre_x(p, x);
#pragma unroll
            for (int hf = 0; hf < 2; hf++) {
              float gt[8];
              unpack8(p.g[hf], gt);
#pragma unroll
              for (int i = 0; i < 8; i++) v[8 * hf + i] = x[8 * hf + i] + round_h(gt[i] * round_h(v[8 * hf + i]));
            }
            if constexpr (kHeads) {
              if (valid) store16_row32(o32 + c0, v);
              // heads run in fp16 under autocast: relu(net) is rounded to fp16 before the Linear
#pragma unroll
              for (int gq = 0; gq < 4; gq++) {
                const float4 wd0 = __ldg(reinterpret_cast<const float4*>(a.Wd + c0) + gq);
                const float4 wd1 = __ldg(reinterpret_cast<const float4*>(a.Wd + kCcC + c0) + gq);
                const float h0 = round_h(fmaxf(v[4 * gq], 0.f)), h1 = round_h(fmaxf(v[4 * gq + 1], 0.f));
                const float h2 = round_h(fmaxf(v[4 * gq + 2], 0.f)), h3 = round_h(fmaxf(v[4 * gq + 3], 0.f));
                acc4[0] += h0 * wd0.x + h1 * wd0.y + h2 * wd0.z + h3 * wd0.w;
                acc4[1] += h0 * wd1.x + h1 * wd1.y + h2 * wd1.z + h3 * wd1.w;
                const float4 ww0 = __ldg(reinterpret_cast<const float4*>(a.Ww + c0) + gq);
                const float4 ww1 = __ldg(reinterpret_cast<const float4*>(a.Ww + kCcC + c0) + gq);
                acc4[2] += h0 * ww0.x + h1 * ww0.y + h2 * ww0.z + h3 * ww0.w;
                acc4[3] += h0 * ww1.x + h1 * ww1.y + h2 * ww1.z + h3 * ww1.w;
              }
            } else {
#pragma unroll
              for (int i = 0; i < 16; i++) { s1 += v[i]; s2 += v[i] * v[i]; }
              tmem_st16(trow + acc_col(c0), v);
            }
          };
          if (heads) pipelined_chunks(cbase, first, fetch, [&](int c0, const Pre& p) { chunk_t(std::true_type{}, c0, p); });
          else pipelined_chunks(cbase, first, fetch, [&](int c0, const Pre& p) { chunk_t(std::false_type{}, c0, p); });
          if (heads) {
            xch[row][part] = make_float4(acc4[0], acc4[1], acc4[2], acc4[3]);
            workers_sync();
            if (part == 0 && valid) {
              float s4[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
              for (int p = 0; p < kCcParts; p++) {
                const float4 u = xch[row][p];
                s4[0] += u.x; s4[1] += u.y; s4[2] += u.z; s4[3] += u.w;
              }
              const float d0 = round_h(s4[0] + a.bd[0]), d1 = round_h(s4[1] + a.bd[1]);
              const float w0 = round_h(s4[2] + a.bw[0]), w1 = round_h(s4[3] + a.bw[1]);
              reinterpret_cast<float2*>(a.delta)[e] = make_float2(d0, d1);
              reinterpret_cast<float2*>(a.weight)[e] =
                  make_float2(round_h(1.0f / (1.0f + __expf(-w0))), round_h(1.0f / (1.0f + __expf(-w1))));
            }
          } else {
            xch[row][part] = make_float4(s1, s2, 0.f, 0.f);
            workers_sync();
            float t1 = 0.f, t2 = 0.f;
#pragma unroll
            for (int p = 0; p < kCcParts; p++) {
              const float4 u = xch[row][p];
              t1 += u.x; t2 += u.y;
            }
            const float mean = t1 * (1.0f / kCcC);
            const float rstd = rsqrtf(fmaxf(t2 * (1.0f / kCcC) - mean * mean, 0.f) + 1e-3f);
#pragma unroll 2
            for (int ci = 0; ci < kCcChunks; ci++) {
              const int c0 = cbase + 16 * ci;
              float v[16];
              tmem_ld16(trow + acc_col(c0), v);
              layer_norm16(v, mean, rstd, L.gamma + c0, L.beta + c0);
              store16_blk32(a.scratch32, cta, row, c0, v);
              store16_tile(A_u, row, c0, v);
            }
          }
        } else if (epi == RVO_CHAIN_EPI_ADD3_LN) {
          // out32[e] = LN((net_in[e] + imap[idx % mod]) + t)
          const float* nin = a.net_in + (size_t)e * kCcC;
          const __half* im = nullptr;
          if (valid) {
            int64_t k = a.imap_idx[e];
            if (a.imap_mod > 0) k %= a.imap_mod;
            im = reinterpret_cast<const __half*>(a.imap16) + (size_t)k * kCcC;
          }
          auto fetch = [&](int c0) {
            Pre p;
            if (valid) {
              pre_row32(p, nin + c0);
              p.g[0] = __ldg(reinterpret_cast<const uint4*>(im + c0));
              p.g[1] = __ldg(reinterpret_cast<const uint4*>(im + c0) + 1);
            }
            return p;
          };
          const Pre first = fetch(cbase);
          mbar_wait(&tfull, t_ph);
          if (wtid == 0) CC_TRACE(2, it * n_layers + l);
          asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
          if (tile_store_pending && writes_tile) {
            if (wtid == 0) asm volatile("cp.async.bulk.wait_group.read 0;\n" ::: "memory");
            workers_sync();
            tile_store_pending = false;
          }
          float s1 = 0.f, s2 = 0.f;
          auto chunk = [&](int c0, const Pre& p) {
            float v[16];
            tmem_ld16(trow + acc_col(c0), v);
            add_bias16(bias, c0, v);
            if (valid) {
              float x[16];
              pre_x(p, x);
#pragma unroll
              for (int hf = 0; hf < 2; hf++) {
                float m8[8];
                unpack8(p.g[hf], m8);
#pragma unroll
                for (int i = 0; i < 8; i++) v[8 * hf + i] = (x[8 * hf + i] + m8[i]) + round_h(v[8 * hf + i]);
              }
            }
#pragma unroll
            for (int i = 0; i < 16; i++) { s1 += v[i]; s2 += v[i] * v[i]; }
            tmem_st16(trow + acc_col(c0), v);
          };
          pipelined_chunks(cbase, first, fetch, chunk);
          xch[row][part] = make_float4(s1, s2, 0.f, 0.f);
          workers_sync();
          float t1 = 0.f, t2 = 0.f;
#pragma unroll
          for (int p = 0; p < kCcParts; p++) {
            const float4 u = xch[row][p];
            t1 += u.x; t2 += u.y;
          }
          const float mean = t1 * (1.0f / kCcC);
          const float rstd = rsqrtf(fmaxf(t2 * (1.0f / kCcC) - mean * mean, 0.f) + 1e-3f);
#pragma unroll 2
          for (int ci = 0; ci < kCcChunks; ci++) {
            const int c0 = cbase + 16 * ci;
            float v[16];
            tmem_ld16(trow + acc_col(c0), v);
            layer_norm16(v, mean, rstd, L.gamma + c0, L.beta + c0);
            if (valid) store16_blk32(a.out32, t, row, c0, v);
            if (writes_tile) store16_tile(A_u, row, c0, v);
          }
        } else {
          // ---- epilogues without global loads: RELU, LN_RELU, STORE16, GATE ----
          mbar_wait(&tfull, t_ph);
          if (wtid == 0) CC_TRACE(2, it * n_layers + l);
          asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
          if (tile_store_pending && writes_tile) {
            if (wtid == 0) asm volatile("cp.async.bulk.wait_group.read 0;\n" ::: "memory");
            workers_sync();
            tile_store_pending = false;
          }
          if (epi == RVO_CHAIN_EPI_RELU) {
#pragma unroll 2
            for (int ci = 0; ci < kCcChunks; ci++) {
              const int c0 = cbase + 16 * ci;
              float v[16];
              tmem_ld16(trow + acc_col(c0), v);
              add_bias16(bias, c0, v);
#pragma unroll
              for (int i = 0; i < 16; i++) v[i] = fmaxf(v[i], 0.f);
              store16_tile(A_u, row, c0, v);
            }
          } else if (epi == RVO_CHAIN_EPI_STORE16) {
            __half* y = reinterpret_cast<__half*>(L.y16) + (size_t)e * L.ldy;
#pragma unroll 2
            for (int ci = 0; ci < kCcChunks; ci++) {
              const int c0 = cbase + 16 * ci;
              float v[16];
              tmem_ld16(trow + acc_col(c0), v);
              add_bias16(bias, c0, v);
              if (valid) store16_row16(y + c0, v);
            }
          } else if (epi == RVO_CHAIN_EPI_GATE) {
#pragma unroll 2
            for (int ci = 0; ci < kCcChunks; ci++) {
              const int c0 = cbase + 16 * ci;
              float v[16];
              tmem_ld16(trow + acc_col(c0), v);
              add_bias16(bias, c0, v);
#pragma unroll
              for (int i = 0; i < 16; i++) v[i] = 1.0f / (1.0f + __expf(-round_h(v[i])));
              store16_blk16(scr16, cta, row, c0, v);
            }
          } else {                                           // LN_RELU: next A = half(relu(LN(t)))
            float s1 = 0.f, s2 = 0.f;
#pragma unroll 2
            for (int ci = 0; ci < kCcChunks; ci++) {
              const int c0 = cbase + 16 * ci;
              float v[16];
              tmem_ld16(trow + acc_col(c0), v);
              add_bias16(bias, c0, v);
#pragma unroll
              for (int i = 0; i < 16; i++) {
                v[i] = round_h(v[i]);
                s1 += v[i];
                s2 += v[i] * v[i];
              }
            }
            xch[row][part] = make_float4(s1, s2, 0.f, 0.f);
            workers_sync();
            float t1 = 0.f, t2 = 0.f;
#pragma unroll
            for (int p = 0; p < kCcParts; p++) {
              const float4 u = xch[row][p];
              t1 += u.x; t2 += u.y;
            }
            const float mean = t1 * (1.0f / kCcC);
            const float rstd = rsqrtf(fmaxf(t2 * (1.0f / kCcC) - mean * mean, 0.f) + 1e-3f);
#pragma unroll 2
            for (int ci = 0; ci < kCcChunks; ci++) {
              const int c0 = cbase + 16 * ci;
              float v[16];
              tmem_ld16(trow + acc_col(c0), v);
              add_bias16(bias, c0, v);
#pragma unroll
              for (int i = 0; i < 16; i++) v[i] = round_h(v[i]);
              layer_norm16(v, mean, rstd, L.gamma + c0, L.beta + c0);
#pragma unroll
              for (int i = 0; i < 16; i++) v[i] = fmaxf(v[i], 0.f);
              store16_tile(A_u, row, c0, v);
            }
          }
        }
        t_ph ^= 1;
        asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
        if (writes_tile) asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
        if (wtid == 0) CC_TRACE(3, it * n_layers + l);
        if (tile_out16) {
          // out16 = half(v) is exactly the activation tile: six TMA box stores straight out of the K-block slots
          workers_sync();
          if (wtid == 0 && t < n_tiles) {
#pragma unroll
            for (int kb = 0; kb < kCcSlots; kb++)
              tma_store_2d(&maps.out16, A_u + kb * kCcSlotBytes, kb * 64, t * kCcM);
            asm volatile("cp.async.bulk.commit_group;\n" ::: "memory");
          }
          tile_store_pending = true;
        }
        if (has_next) mbar_arrive(&hfull);
      }
      // the exchange buffer, the scratch rows and the activation tile are reused by the next tile
      if (tile_store_pending) {
        if (wtid == 0) asm volatile("cp.async.bulk.wait_group.read 0;\n" ::: "memory");
        tile_store_pending = false;
      }
      workers_sync();
    }
    if (stream0) asm volatile("cp.async.wait_all;\n" ::: "memory");
    if (wtid == 0) asm volatile("cp.async.bulk.wait_group 0;\n" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
  __syncthreads();
  cluster_sync_all();                                     // no CTA leaves while a peer may still signal its barriers
  if (warp == 1) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;\n" ::"r"(tmem_base) : "memory");
  }
}

}  // namespace rvo

using namespace rvo;

// Cluster size of the chain kernel: the largest of 4 / 2 / 1 whose co-resident clusters cover (nearly) the whole
// device — the kernel is persistent, a cluster that does not fit in the first wave would serialise behind it.
// g_chain_ctas = CTAs of one wave at that size.  rvo_up_chain_set_cluster forces a size (benchmarks).
static int g_chain_csize = 0, g_chain_ctas = 0, g_chain_forced = 0;

static int chain_cluster_size() {
  if (g_chain_csize) return g_chain_csize;
  if (cudaFuncSetAttribute(up_chain_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kCcSmemBytes) != cudaSuccess)
    return -1;
  const int cand[3] = {4, 2, 1};
  for (int i = 0; i < 3; i++) {
    const int cs = cand[i];
    if (g_chain_forced && cs != g_chain_forced) continue;
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.gridDim = dim3(kNumSMs / cs * cs);
    cfg.blockDim = dim3(kCcThreads);
    cfg.dynamicSmemBytes = kCcSmemBytes;
    cudaLaunchAttribute attr;
    attr.id = cudaLaunchAttributeClusterDimension;
    attr.val.clusterDim.x = cs;
    attr.val.clusterDim.y = 1;
    attr.val.clusterDim.z = 1;
    cfg.attrs = &attr;
    cfg.numAttrs = 1;
    int n = 0;
    if (cudaOccupancyMaxActiveClusters(&n, up_chain_kernel, &cfg) != cudaSuccess) {
      cudaGetLastError();
      continue;
    }
    int ctas = n * cs;
    if (ctas > kNumSMs) ctas = kNumSMs / cs * cs;
    if (ctas >= (g_chain_forced ? cs : (cs == 1 ? 1 : kNumSMs * 7 / 8))) {   // accept 4 only if >= 129 CTAs fit, ...
      g_chain_csize = cs;
      g_chain_ctas = ctas;
      return cs;
    }
  }
  return -1;
}

extern "C" int rvo_up_chain_set_cluster(int cluster_size) {
  RVO_CHECK_ARG(cluster_size == 0 || cluster_size == 1 || cluster_size == 2 || cluster_size == 4,
                "rvo_up_chain_set_cluster: %d (0 = automatic, 1, 2 or 4)", cluster_size);
  g_chain_forced = cluster_size;
  g_chain_csize = 0;
  return RVO_OK;
}

extern "C" int rvo_up_chain_info(int* cluster_size, int* ctas) {
  RVO_CHECK_ARG(chain_cluster_size() > 0, "rvo_up_chain_info: no cluster size of up_chain_kernel fits this device");
  if (cluster_size) *cluster_size = g_chain_csize;
  if (ctas) *ctas = g_chain_ctas;
  return RVO_OK;
}

#ifdef RVO_DEBUG
extern "C" int rvo_up_chain_trace(long long* host_out) {
  RVO_CUDA(cudaDeviceSynchronize());
  RVO_CUDA(cudaMemcpyFromSymbol(host_out, g_cc_trace, sizeof(long long) * 4 * 64));
  return RVO_OK;
}
#endif

extern "C" int64_t rvo_up_chain_scratch_rows(void) { return (int64_t)kNumSMs * kCcM; }

extern "C" int rvo_up_chain(const rvo_chain_t* c, void* stream) {
  RVO_CHECK_ARG(c != nullptr, "rvo_up_chain: null descriptor");
  RVO_CHECK_ARG(c->M >= 0 && c->n_layers >= 1 && c->n_layers <= RVO_CHAIN_MAX_LAYERS,
                "rvo_up_chain: M=%d n_layers=%d (1..%d)", c->M, c->n_layers, RVO_CHAIN_MAX_LAYERS);
  if (c->M == 0) return RVO_OK;
  const bool rows = c->prologue == RVO_CHAIN_PRO_ROWS;
  RVO_CHECK_ARG(rows || c->prologue == RVO_CHAIN_PRO_EXPAND || c->prologue == RVO_CHAIN_PRO_EXPAND_LN,
                "rvo_up_chain: prologue %d", c->prologue);
  if (rows) {
    RVO_CHECK_ARG(c->a16 && (reinterpret_cast<uintptr_t>(c->a16) & 15u) == 0 && c->lda % 8 == 0 &&
                      c->lda >= c->layer[0].K,
                  "rvo_up_chain: a16 must be 16-byte aligned with a row pitch (multiple of 8) >= K");
  } else {
    RVO_CHECK_ARG(c->x32 && c->hy_a && c->grp_a, "rvo_up_chain: expand prologue needs x32 / hy_a / grp_a");
    RVO_CHECK_ARG((c->hy_b == nullptr) == (c->grp_b == nullptr), "rvo_up_chain: hy_b and grp_b go together");
    RVO_CHECK_ARG(c->layer[0].K == kCcC, "rvo_up_chain: expand prologue needs K = 384 in layer 0");
    if (c->prologue == RVO_CHAIN_PRO_EXPAND_LN)
      RVO_CHECK_ARG(c->pro_gamma && c->pro_beta && c->scratch32, "rvo_up_chain: EXPAND_LN needs gamma / beta / scratch32");
  }
  ChainMaps maps;
  memset(&maps, 0, sizeof(maps));
  const int csize = chain_cluster_size();
  RVO_CHECK_ARG(csize > 0, "rvo_up_chain: no cluster size of up_chain_kernel fits this device");
  bool gate_pending = false;
  for (int l = 0; l < c->n_layers; l++) {
    const rvo_chain_layer_t& L = c->layer[l];
    const bool last = l + 1 == c->n_layers;
    RVO_CHECK_ARG(L.w16 && L.bias16, "rvo_up_chain: layer %d has no weights / bias", l);
    RVO_CHECK_ARG((reinterpret_cast<uintptr_t>(L.bias16) & 15u) == 0, "rvo_up_chain: layer %d bias alignment", l);
    RVO_CHECK_ARG(L.K > 0 && L.K % 8 == 0 && (L.K == kCcC || (l == 0 && rows)),
                  "rvo_up_chain: layer %d K=%d (384, or a multiple of 8 in a streamed layer 0)", l, L.K);
    switch (L.epilogue) {
      case RVO_CHAIN_EPI_RELU:
        RVO_CHECK_ARG(!last, "rvo_up_chain: layer %d: RELU must feed another layer", l);
        break;
      case RVO_CHAIN_EPI_LN_RELU:
        RVO_CHECK_ARG(!last && L.gamma && L.beta, "rvo_up_chain: layer %d: LN_RELU needs gamma / beta and a next layer", l);
        break;
      case RVO_CHAIN_EPI_ADD3_LN:
        RVO_CHECK_ARG(L.gamma && L.beta && c->net_in && c->imap16 && c->imap_idx && c->out32,
                      "rvo_up_chain: layer %d: ADD3_LN needs gamma / beta / net_in / imap16 / imap_idx / out32", l);
        break;
      case RVO_CHAIN_EPI_RES:
        RVO_CHECK_ARG(c->res32 && c->out32, "rvo_up_chain: layer %d: RES needs res32 / out32", l);
        break;
      case RVO_CHAIN_EPI_STORE16:
        RVO_CHECK_ARG(L.y16 && L.ldy >= kCcC && L.ldy % 8 == 0 && (reinterpret_cast<uintptr_t>(L.y16) & 15u) == 0,
                      "rvo_up_chain: layer %d: STORE16 destination / pitch", l);
        break;
      case RVO_CHAIN_EPI_GATE:
        RVO_CHECK_ARG(!last && c->scratch16, "rvo_up_chain: layer %d: GATE needs scratch16 and a next layer", l);
        gate_pending = true;
        break;
      case RVO_CHAIN_EPI_GATED_LN:
        RVO_CHECK_ARG(!last && gate_pending && L.gamma && L.beta && c->scratch32 && c->scratch16 &&
                          c->prologue == RVO_CHAIN_PRO_EXPAND_LN,
                      "rvo_up_chain: layer %d: GATED_LN needs a GATE before it, gamma / beta, scratch and EXPAND_LN", l);
        gate_pending = false;
        break;
      case RVO_CHAIN_EPI_GATED_HEADS:
        RVO_CHECK_ARG(last && gate_pending && c->scratch32 && c->scratch16 && c->out32 && c->Wd && c->bd && c->Ww &&
                          c->bw && c->delta && c->weight && c->prologue == RVO_CHAIN_PRO_EXPAND_LN,
                      "rvo_up_chain: layer %d: GATED_HEADS must be last, after a GATE, with heads / out32 / scratch", l);
        gate_pending = false;
        break;
      default:
        RVO_CHECK_ARG(false, "rvo_up_chain: layer %d: epilogue %d", l, L.epilogue);
    }
    int rc = make_tmap_2d_f16(L.w16, kCcC, L.K, L.K, kCcNH / csize, &maps.m[l], "rvo_up_chain(w)");
    if (rc != RVO_OK) return rc;
  }
  if (c->out16) {
    int rc = make_tmap_2d_f16(c->out16, c->M, kCcC, kCcC, kCcM, &maps.out16, "rvo_up_chain(out16)");
    if (rc != RVO_OK) return rc;
  }
  const int n_tiles = (c->M + kCcM - 1) / kCcM;
  int grid = (n_tiles + csize - 1) / csize * csize;          // whole clusters
  if (grid > g_chain_ctas) grid = g_chain_ctas;
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = dim3(grid);
  cfg.blockDim = dim3(kCcThreads);
  cfg.dynamicSmemBytes = kCcSmemBytes;
  cfg.stream = (cudaStream_t)stream;
  cudaLaunchAttribute attr;
  attr.id = cudaLaunchAttributeClusterDimension;
  attr.val.clusterDim.x = csize;
  attr.val.clusterDim.y = 1;
  attr.val.clusterDim.z = 1;
  cfg.attrs = &attr;
  cfg.numAttrs = 1;
  RVO_CUDA(cudaLaunchKernelEx(&cfg, up_chain_kernel, *c, maps));
  RVO_LAUNCH_CHECK("up_chain_kernel");
  return RVO_OK;
}
