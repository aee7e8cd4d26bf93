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
constexpr int kCcWorkers = 512;                    // 16 warps
constexpr int kCcThreads = 64 + kCcWorkers;        // warp 0: weight TMA, warp 1: MMA issuer
constexpr int kCcSmemBytes = kCcSlots * kCcSlotBytes + kCcWStages * kCcWStage + 1024;
constexpr int kCcParts = 4;                        // column parts per row (96 columns each)
constexpr int kCcPartCols = kCcC / kCcParts;       // 96

struct ChainMaps {
  TcTmap m[RVO_CHAIN_MAX_LAYERS];
};

__device__ __forceinline__ float round_h(float v) { return __half2float(__float2half_rn(v)); }

__device__ __forceinline__ void tmem_st32(uint32_t taddr, const float* v) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
      "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};\n" ::"r"(taddr),
      "r"(__float_as_uint(v[0])), "r"(__float_as_uint(v[1])), "r"(__float_as_uint(v[2])), "r"(__float_as_uint(v[3])),
      "r"(__float_as_uint(v[4])), "r"(__float_as_uint(v[5])), "r"(__float_as_uint(v[6])), "r"(__float_as_uint(v[7])),
      "r"(__float_as_uint(v[8])), "r"(__float_as_uint(v[9])), "r"(__float_as_uint(v[10])), "r"(__float_as_uint(v[11])),
      "r"(__float_as_uint(v[12])), "r"(__float_as_uint(v[13])), "r"(__float_as_uint(v[14])), "r"(__float_as_uint(v[15])),
      "r"(__float_as_uint(v[16])), "r"(__float_as_uint(v[17])), "r"(__float_as_uint(v[18])), "r"(__float_as_uint(v[19])),
      "r"(__float_as_uint(v[20])), "r"(__float_as_uint(v[21])), "r"(__float_as_uint(v[22])), "r"(__float_as_uint(v[23])),
      "r"(__float_as_uint(v[24])), "r"(__float_as_uint(v[25])), "r"(__float_as_uint(v[26])), "r"(__float_as_uint(v[27])),
      "r"(__float_as_uint(v[28])), "r"(__float_as_uint(v[29])), "r"(__float_as_uint(v[30])), "r"(__float_as_uint(v[31]))
      : "memory");
  asm volatile("tcgen05.wait::st.sync.aligned;\n" ::: "memory");
}

// 8 halves (one 16-byte chunk) -> 8 floats
__device__ __forceinline__ void unpack8(const uint4& u, float* f) {
  const __half2* h = reinterpret_cast<const __half2*>(&u);
#pragma unroll
  for (int i = 0; i < 4; i++) {
    const float2 t = __half22float2(h[i]);
    f[2 * i] = t.x;
    f[2 * i + 1] = t.y;
  }
}
__device__ __forceinline__ uint4 pack8(const float* f) {
  uint4 u;
  __half2* h = reinterpret_cast<__half2*>(&u);
#pragma unroll
  for (int i = 0; i < 4; i++) h[i] = __floats2half2_rn(f[2 * i], f[2 * i + 1]);
  return u;
}

// v[0..32) += bias16[c0 .. c0+32)   (uniform address across the warp: one broadcast transaction per 16 bytes)
__device__ __forceinline__ void add_bias32(const __half* __restrict__ bias, int c0, float* v) {
  const uint4* p = reinterpret_cast<const uint4*>(bias + c0);
#pragma unroll
  for (int g = 0; g < 4; g++) {
    float b[8];
    unpack8(__ldg(p + g), b);
#pragma unroll
    for (int i = 0; i < 8; i++) v[8 * g + i] += b[i];
  }
}

// 32 fp32 values of row `p` (already offset to the first column) — 128 contiguous bytes
__device__ __forceinline__ void load32_f32(const float* __restrict__ p, float* v) {
  const float4* q = reinterpret_cast<const float4*>(p);
#pragma unroll
  for (int g = 0; g < 8; g++) {
    const float4 t = q[g];
    v[4 * g] = t.x; v[4 * g + 1] = t.y; v[4 * g + 2] = t.z; v[4 * g + 3] = t.w;
  }
}
__device__ __forceinline__ void store32_f32(float* __restrict__ p, const float* v) {
  float4* q = reinterpret_cast<float4*>(p);
#pragma unroll
  for (int g = 0; g < 8; g++) q[g] = make_float4(v[4 * g], v[4 * g + 1], v[4 * g + 2], v[4 * g + 3]);
}
__device__ __forceinline__ void store32_f16(__half* __restrict__ p, const float* v) {
  uint4* q = reinterpret_cast<uint4*>(p);
#pragma unroll
  for (int g = 0; g < 4; g++) q[g] = pack8(v + 8 * g);
}

// the 32 columns [c0, c0+32) of row r of the activation tile (fp16, K-major SWIZZLE_128B K blocks)
__device__ __forceinline__ void store32_tile(uint32_t A_u, int r, int c0, const float* v) {
#pragma unroll
  for (int g = 0; g < 4; g++) {
    const int col = c0 + 8 * g;
    const uint32_t addr = A_u + (uint32_t)(col >> 6) * kCcSlotBytes + (uint32_t)r * 128u +
                          (uint32_t)((((col & 63) >> 3) ^ (r & 7)) << 4);
    const uint4 u = pack8(v + 8 * g);
    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};\n" ::"r"(addr), "r"(u.x), "r"(u.y), "r"(u.z), "r"(u.w)
                 : "memory");
  }
}

__device__ __forceinline__ void workers_sync() { asm volatile("bar.sync 1, %0;\n" ::"n"(kCcWorkers) : "memory"); }

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// TMEM column of output column c: the two N = 192 halves sit at columns 0 and 256
__device__ __forceinline__ uint32_t acc_col(int c) { return (uint32_t)(c + (c >= kCcNH ? 256 - kCcNH : 0)); }

__global__ void __launch_bounds__(kCcThreads, 1)
up_chain_kernel(const rvo_chain_t a, const __grid_constant__ ChainMaps maps) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ uint64_t wfull[kCcWStages], wempty[kCcWStages], afull[kCcSlots], aempty[kCcSlots], hfull, tfull;
  __shared__ uint32_t tmem_base_s;
  __shared__ float4 xch[kCcM][kCcParts];             // row statistics / head partial sums of the four column parts
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int M = a.M, n_layers = a.n_layers;
  const int n_tiles = (M + kCcM - 1) / kCcM;

  if (tid == 0) {
    for (int s = 0; s < kCcWStages; s++) {
      mbar_init(&wfull[s], 1);
      mbar_init(&wempty[s], 1);
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
  asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
  const uint32_t tmem_base = tmem_base_s;
  const uint32_t A_u = smem_u32(smem), W_u = A_u + kCcSlots * kCcSlotBytes;
  const bool stream0 = a.prologue == RVO_CHAIN_PRO_ROWS;     // layer 0 takes its A operand through the slot ring

  if (warp == 0) {
    // ===== weight producer: (tile, layer, K block, N half) in the order the MMA issuer consumes them =====
    if (lane == 0) {
      int s = 0, ph = 0;
      for (int t = blockIdx.x; t < n_tiles; t += gridDim.x)
        for (int l = 0; l < n_layers; l++) {
          const int nkb = (a.layer[l].K + 63) >> 6;
          for (int kb = 0; kb < nkb; kb++)
#pragma unroll
            for (int nh = 0; nh < 2; nh++) {
              mbar_wait(&wempty[s], ph ^ 1);
              mbar_expect_tx(&wfull[s], (uint32_t)kCcWStage);
              tma_load_2d(W_u + s * kCcWStage, &maps.m[l], kb * 64, nh * kCcNH, &wfull[s]);
              if (++s == kCcWStages) { s = 0; ph ^= 1; }
            }
        }
    }
  } else if (warp == 1) {
    // ===== MMA issuer =====
    int s = 0, ph = 0;
    uint32_t a_par = 0, h_ph = 0;
    constexpr uint32_t idesc = umma_idesc_f16(kCcM, kCcNH);
    for (int t = blockIdx.x; t < n_tiles; t += gridDim.x)
      for (int l = 0; l < n_layers; l++) {
        const int nkb = (a.layer[l].K + 63) >> 6;
        const bool stream = stream0 && l == 0;
        if (!stream) {                                   // whole A tile written by the workers (prologue / epilogue)
          mbar_wait_spin(&hfull, h_ph);
          h_ph ^= 1;
          asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
          asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
        }
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
              umma_commit(&wempty[s]);
            }
            __syncwarp();
            if (++s == kCcWStages) { s = 0; ph ^= 1; }
          }
          if (stream && lane == 0) umma_commit(&aempty[slot]);
          __syncwarp();
        }
        if (lane == 0) umma_commit(&tfull);
        __syncwarp();
      }
  } else {
    // ===== workers: thread = (row, 96-column part) in the epilogues =====
    const int wtid = tid - 64, wwarp = wtid >> 5;
    const int q = warp & 3, part = wwarp >> 2;             // warps 2..17: every quadrant gets parts 0..3
    const int row = q * 32 + lane;
    const uint32_t trow = tmem_base + ((uint32_t)(q * 32) << 16);
    const int cbase = part * kCcPartCols;
    uint32_t e_par = (1u << kCcSlots) - 1u, t_ph = 0;
    const __half* a16 = reinterpret_cast<const __half*>(a.a16);
    float* scr32 = a.scratch32 ? a.scratch32 + ((size_t)blockIdx.x * kCcM + row) * kCcC : nullptr;
    __half* scr16 = a.scratch16 ? reinterpret_cast<__half*>(a.scratch16) + ((size_t)blockIdx.x * kCcM + row) * kCcC
                                : nullptr;

    for (int t = blockIdx.x; t < n_tiles; t += gridDim.x) {
      const int e = t * kCcM + row;
      const bool valid = e < M;

      // ---------------- prologue ----------------
      if (stream0) {
        // thread owns 16-byte chunk `ch` of rows r0 and r0 + 64 of every K block
        const int ch = wtid & 7, r0 = wtid >> 3;
        int64_t src[2];
#pragma unroll
        for (int j = 0; j < 2; j++) {
          const int ee = t * kCcM + r0 + 64 * j;
          int64_t g = ee < M ? (a.gather ? a.gather[ee] : (int64_t)ee) : -1;
          src[j] = g;
        }
        const int K0 = a.layer[0].K, nkb = (K0 + 63) >> 6;
        for (int kb = 0; kb < nkb; kb++) {
          const int slot = kb % kCcSlots;
          mbar_wait(&aempty[slot], (e_par >> slot) & 1u);
          e_par ^= 1u << slot;
          const int col = kb * 64 + ch * 8;
          const uint32_t dst0 = A_u + slot * kCcSlotBytes + r0 * 128 + (uint32_t)((ch ^ (r0 & 7)) << 4);
#pragma unroll
          for (int j = 0; j < 2; j++) {                     // (r0 + 64) & 7 == r0 & 7
            const bool ok = src[j] >= 0 && col < K0;
            cp_async16(dst0 + j * (64 * 128), a16 + (ok ? src[j] * a.lda + col : 0), ok ? 16u : 0u);
          }
          asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];\n" ::"r"(smem_u32(&afull[slot]))
                       : "memory");
        }
      } else {
        // SoftAgg expand (+ the first LayerNorm of Update.gru): warp per row, lane owns 4 columns of each 128
        const __half* hya = reinterpret_cast<const __half*>(a.hy_a);
        const __half* hyb = reinterpret_cast<const __half*>(a.hy_b);
        const bool with_ln = a.prologue == RVO_CHAIN_PRO_EXPAND_LN;
        for (int r = wwarp; r < kCcM; r += kCcWorkers / 32) {
          const int ee = t * kCcM + r;
          float v[3][4];
          if (ee < M) {
            const int ga = a.grp_a[ee];
            const int gb = hyb ? a.grp_b[ee] : 0;
#pragma unroll
            for (int j = 0; j < 3; j++) {
              const float4 x = reinterpret_cast<const float4*>(a.x32 + (size_t)ee * kCcC + 128 * j)[lane];
              const uint2 ha = reinterpret_cast<const uint2*>(hya + (size_t)ga * kCcC + 128 * j)[lane];
              const float2 a0 = __half22float2(*reinterpret_cast<const __half2*>(&ha.x));
              const float2 a1 = __half22float2(*reinterpret_cast<const __half2*>(&ha.y));
              v[j][0] = x.x + a0.x; v[j][1] = x.y + a0.y; v[j][2] = x.z + a1.x; v[j][3] = x.w + a1.y;
              if (hyb) {
                const uint2 hb = reinterpret_cast<const uint2*>(hyb + (size_t)gb * kCcC + 128 * j)[lane];
                const float2 b0 = __half22float2(*reinterpret_cast<const __half2*>(&hb.x));
                const float2 b1 = __half22float2(*reinterpret_cast<const __half2*>(&hb.y));
                v[j][0] += b0.x; v[j][1] += b0.y; v[j][2] += b1.x; v[j][3] += b1.y;
              }
            }
          } else {
#pragma unroll
            for (int j = 0; j < 3; j++)
#pragma unroll
              for (int i = 0; i < 4; i++) v[j][i] = 0.f;
          }
          if (with_ln) {                                   // two-pass LayerNorm in registers, eps 1e-3
            float s = 0.f;
#pragma unroll
            for (int j = 0; j < 3; j++)
#pragma unroll
              for (int i = 0; i < 4; i++) s += v[j][i];
            const float mean = warp_sum(s) * (1.0f / kCcC);
            float ss = 0.f;
#pragma unroll
            for (int j = 0; j < 3; j++)
#pragma unroll
              for (int i = 0; i < 4; i++) { const float d = v[j][i] - mean; ss += d * d; }
            const float rstd = rsqrtf(warp_sum(ss) * (1.0f / kCcC) + 1e-3f);
#pragma unroll
            for (int j = 0; j < 3; j++) {
              const float4 g = reinterpret_cast<const float4*>(a.pro_gamma + 128 * j)[lane];
              const float4 b = reinterpret_cast<const float4*>(a.pro_beta + 128 * j)[lane];
              v[j][0] = (v[j][0] - mean) * rstd * g.x + b.x;
              v[j][1] = (v[j][1] - mean) * rstd * g.y + b.y;
              v[j][2] = (v[j][2] - mean) * rstd * g.z + b.z;
              v[j][3] = (v[j][3] - mean) * rstd * g.w + b.w;
            }
            float* sr = a.scratch32 + ((size_t)blockIdx.x * kCcM + r) * kCcC;
#pragma unroll
            for (int j = 0; j < 3; j++)
              reinterpret_cast<float4*>(sr + 128 * j)[lane] = make_float4(v[j][0], v[j][1], v[j][2], v[j][3]);
          }
#pragma unroll
          for (int j = 0; j < 3; j++) {
            const int col = 128 * j + 4 * lane;
            const uint32_t addr = A_u + (uint32_t)(col >> 6) * kCcSlotBytes + (uint32_t)r * 128u +
                                  (uint32_t)((((col & 63) >> 3) ^ (r & 7)) << 4) + (uint32_t)((col & 7) * 2);
            const __half2 p0 = __floats2half2_rn(v[j][0], v[j][1]), p1 = __floats2half2_rn(v[j][2], v[j][3]);
            asm volatile("st.shared.v2.b32 [%0], {%1, %2};\n" ::"r"(addr), "r"(*reinterpret_cast<const uint32_t*>(&p0)),
                         "r"(*reinterpret_cast<const uint32_t*>(&p1))
                         : "memory");
          }
        }
        // the scratch rows written above are read back by OTHER threads in the GATED epilogues
        if (with_ln) __threadfence_block();
        asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
        mbar_arrive(&hfull);
      }

      // ---------------- layers ----------------
      for (int l = 0; l < n_layers; l++) {
        const rvo_chain_layer_t& L = a.layer[l];
        const __half* bias = reinterpret_cast<const __half*>(L.bias16);
        const bool has_next = l + 1 < n_layers;
        mbar_wait(&tfull, t_ph);
        t_ph ^= 1;
        asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
        const int epi = L.epilogue;

        if (epi == RVO_CHAIN_EPI_RELU) {
#pragma unroll 1
          for (int ci = 0; ci < 3; ci++) {
            const int c0 = cbase + 32 * ci;
            float v[32];
            tmem_ld32(trow + acc_col(c0), v);
            add_bias32(bias, c0, v);
#pragma unroll
            for (int i = 0; i < 32; i++) v[i] = fmaxf(v[i], 0.f);
            store32_tile(A_u, row, c0, v);
          }
        } else if (epi == RVO_CHAIN_EPI_STORE16) {
          __half* y = reinterpret_cast<__half*>(L.y16) + (size_t)e * L.ldy;
#pragma unroll 1
          for (int ci = 0; ci < 3; ci++) {
            const int c0 = cbase + 32 * ci;
            float v[32];
            tmem_ld32(trow + acc_col(c0), v);
            add_bias32(bias, c0, v);
            if (valid) store32_f16(y + c0, v);
          }
        } else if (epi == RVO_CHAIN_EPI_GATE) {
#pragma unroll 1
          for (int ci = 0; ci < 3; ci++) {
            const int c0 = cbase + 32 * ci;
            float v[32];
            tmem_ld32(trow + acc_col(c0), v);
            add_bias32(bias, c0, v);
#pragma unroll
            for (int i = 0; i < 32; i++) v[i] = 1.0f / (1.0f + __expf(-round_h(v[i])));
            store32_f16(scr16 + c0, v);
          }
        } else if (epi == RVO_CHAIN_EPI_RES) {
          const float* res = a.res32 + (size_t)e * kCcC;
          float* o32 = a.out32 + (size_t)e * kCcC;
          __half* o16 = a.out16 ? reinterpret_cast<__half*>(a.out16) + (size_t)e * kCcC : nullptr;
#pragma unroll 1
          for (int ci = 0; ci < 3; ci++) {
            const int c0 = cbase + 32 * ci;
            float v[32], x[32];
            tmem_ld32(trow + acc_col(c0), v);
            add_bias32(bias, c0, v);
            if (valid) {
              load32_f32(res + c0, x);
#pragma unroll
              for (int i = 0; i < 32; i++) v[i] = x[i] + round_h(v[i]);
              store32_f32(o32 + c0, v);
              if (o16) store32_f16(o16 + c0, v);
            }
            if (has_next) store32_tile(A_u, row, c0, v);
          }
        } else if (epi == RVO_CHAIN_EPI_GATED_HEADS) {
          float* o32 = a.out32 + (size_t)e * kCcC;
          float acc4[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll 1
          for (int ci = 0; ci < 3; ci++) {
            const int c0 = cbase + 32 * ci;
            float v[32], x[32];
            tmem_ld32(trow + acc_col(c0), v);
            add_bias32(bias, c0, v);
            load32_f32(scr32 + c0, x);
#pragma unroll
            for (int g = 0; g < 4; g++) {
              float gt[8];
              unpack8(reinterpret_cast<const uint4*>(scr16 + c0)[g], gt);
#pragma unroll
              for (int i = 0; i < 8; i++) v[8 * g + i] = x[8 * g + i] + round_h(gt[i] * round_h(v[8 * g + i]));
            }
            if (valid) store32_f32(o32 + c0, v);
            // heads run in fp16 under autocast: relu(net) is rounded to fp16 before the Linear
#pragma unroll
            for (int g = 0; g < 8; g++) {
              const float4 wd0 = __ldg(reinterpret_cast<const float4*>(a.Wd + c0) + g);
              const float4 wd1 = __ldg(reinterpret_cast<const float4*>(a.Wd + kCcC + c0) + g);
              const float4 ww0 = __ldg(reinterpret_cast<const float4*>(a.Ww + c0) + g);
              const float4 ww1 = __ldg(reinterpret_cast<const float4*>(a.Ww + kCcC + c0) + g);
              const float h0 = round_h(fmaxf(v[4 * g], 0.f)), h1 = round_h(fmaxf(v[4 * g + 1], 0.f));
              const float h2 = round_h(fmaxf(v[4 * g + 2], 0.f)), h3 = round_h(fmaxf(v[4 * g + 3], 0.f));
              acc4[0] += h0 * wd0.x + h1 * wd0.y + h2 * wd0.z + h3 * wd0.w;
              acc4[1] += h0 * wd1.x + h1 * wd1.y + h2 * wd1.z + h3 * wd1.w;
              acc4[2] += h0 * ww0.x + h1 * ww0.y + h2 * ww0.z + h3 * ww0.w;
              acc4[3] += h0 * ww1.x + h1 * ww1.y + h2 * ww1.z + h3 * ww1.w;
            }
          }
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
          // ---- LayerNorm epilogues: pass 1 forms the pre-norm value and its row statistics, pass 2 normalises ----
          const float* nin = a.net_in ? a.net_in + (size_t)e * kCcC : nullptr;
          const __half* im = nullptr;
          if (epi == RVO_CHAIN_EPI_ADD3_LN && valid) {
            int64_t k = a.imap_idx[e];
            if (a.imap_mod > 0) k %= a.imap_mod;
            im = reinterpret_cast<const __half*>(a.imap16) + (size_t)k * kCcC;
          }
          float s1 = 0.f, s2 = 0.f;
#pragma unroll 1
          for (int ci = 0; ci < 3; ci++) {
            const int c0 = cbase + 32 * ci;
            float v[32];
            tmem_ld32(trow + acc_col(c0), v);
            add_bias32(bias, c0, v);
            if (epi == RVO_CHAIN_EPI_LN_RELU) {
#pragma unroll
              for (int i = 0; i < 32; i++) v[i] = round_h(v[i]);
            } else if (epi == RVO_CHAIN_EPI_ADD3_LN) {
              if (valid) {
                float x[32];
                load32_f32(nin + c0, x);
#pragma unroll
                for (int g = 0; g < 4; g++) {
                  float m8[8];
                  unpack8(__ldg(reinterpret_cast<const uint4*>(im + c0) + g), m8);
#pragma unroll
                  for (int i = 0; i < 8; i++) v[8 * g + i] = (x[8 * g + i] + m8[i]) + round_h(v[8 * g + i]);
                }
              }
              tmem_st32(trow + acc_col(c0), v);
            } else {                                          // GATED_LN
              float x[32];
              load32_f32(scr32 + c0, x);
#pragma unroll
              for (int g = 0; g < 4; g++) {
                float gt[8];
                unpack8(reinterpret_cast<const uint4*>(scr16 + c0)[g], gt);
#pragma unroll
                for (int i = 0; i < 8; i++) v[8 * g + i] = x[8 * g + i] + round_h(gt[i] * round_h(v[8 * g + i]));
              }
              tmem_st32(trow + acc_col(c0), v);
            }
#pragma unroll
            for (int i = 0; i < 32; i++) { s1 += v[i]; s2 += v[i] * v[i]; }
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
          float* o32 = a.out32 ? a.out32 + (size_t)e * kCcC : nullptr;
          __half* o16 = a.out16 ? reinterpret_cast<__half*>(a.out16) + (size_t)e * kCcC : nullptr;
#pragma unroll 1
          for (int ci = 0; ci < 3; ci++) {
            const int c0 = cbase + 32 * ci;
            float v[32];
            tmem_ld32(trow + acc_col(c0), v);
            if (epi == RVO_CHAIN_EPI_LN_RELU) {
              add_bias32(bias, c0, v);
#pragma unroll
              for (int i = 0; i < 32; i++) v[i] = round_h(v[i]);
            }
#pragma unroll
            for (int g = 0; g < 8; g++) {
              const float4 ga = __ldg(reinterpret_cast<const float4*>(L.gamma + c0) + g);
              const float4 be = __ldg(reinterpret_cast<const float4*>(L.beta + c0) + g);
              v[4 * g] = (v[4 * g] - mean) * rstd * ga.x + be.x;
              v[4 * g + 1] = (v[4 * g + 1] - mean) * rstd * ga.y + be.y;
              v[4 * g + 2] = (v[4 * g + 2] - mean) * rstd * ga.z + be.z;
              v[4 * g + 3] = (v[4 * g + 3] - mean) * rstd * ga.w + be.w;
            }
            if (epi == RVO_CHAIN_EPI_LN_RELU) {
#pragma unroll
              for (int i = 0; i < 32; i++) v[i] = fmaxf(v[i], 0.f);
            } else if (epi == RVO_CHAIN_EPI_ADD3_LN) {
              if (valid) {
                store32_f32(o32 + c0, v);
                if (o16) store32_f16(o16 + c0, v);
              }
            } else {
              store32_f32(scr32 + c0, v);
            }
            if (has_next) store32_tile(A_u, row, c0, v);
          }
        }
        asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
        if (has_next) {
          asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
          mbar_arrive(&hfull);
        }
      }
      // the exchange buffer and (GATED stretches) the scratch rows are reused by the next tile
      workers_sync();
    }
    if (stream0) asm volatile("cp.async.wait_all;\n" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
  __syncthreads();
  if (warp == 1) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;\n" ::"r"(tmem_base) : "memory");
  }
}

}  // namespace rvo

using namespace rvo;

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
    int rc = make_tmap_2d_f16(L.w16, kCcC, L.K, L.K, kCcNH, &maps.m[l], "rvo_up_chain(w)");
    if (rc != RVO_OK) return rc;
  }
  const int n_tiles = (c->M + kCcM - 1) / kCcM;
  const int grid = n_tiles < kNumSMs ? n_tiles : kNumSMs;
  RVO_CUDA(cudaFuncSetAttribute(up_chain_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kCcSmemBytes));
  up_chain_kernel<<<grid, kCcThreads, kCcSmemBytes, (cudaStream_t)stream>>>(*c, maps);
  RVO_LAUNCH_CHECK("up_chain_kernel");
  return RVO_OK;
}
