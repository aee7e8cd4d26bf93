// corr_tc.cu — altcorr lookup as tile GEMMs on the 5th-gen tensor cores (tcgen05 + TMEM), sm_100a.
//
// The per-edge kernel (altcorr.cu) re-reads every 8x8x128 window from L2 once per patch pixel: 2.3 GB of
// L2->SM traffic per update for 256 MB of algorithmic bytes, and its mma.sync path tops out near
// 620 TFLOP/s.  Here the loop is turned inside out:
//
//   bin      every (edge, patch pixel, level) row is assigned to the 16x16-position tile of its target
//            frame that contains its 8x8 window (tiles step by 9 so every window fits in exactly one
//            tile): histogram + scan + scatter on the device, no host sync;
//   GEMM     one CTA per (tile, <=128 rows): A = the rows' 128-channel patch vectors gathered into
//            shared memory, B = the tile's 256 feature vectors (zero-filled outside the map), both in
//            the canonical K-major SWIZZLE_128B layout; ONE elected thread issues 8 tcgen05.mma
//            (M=128, N=256, K=16) into a 128x256 fp32 accumulator in TMEM;
//   epilogue thread r owns TMEM lane r: tcgen05.ld brings one tile row (16 columns) at a time, the
//            thread picks its own 8-wide window, does the separable bilinear blend and stages its
//            7x7 outputs in shared memory; rows are then written with coalesced 100-byte stores.
//
// Output layout ("tile layout", consumed by the update operator with a permuted first-layer weight):
//   out[e, ((lvl*9 + pix)*7 + a)*8 + b], a = y offset, b = x offset (0..6); b = 7 is a zero pad so that
//   every (row, a) is one aligned 16-byte store.  rvo_corr_pyramid keeps the reference's [E, 882] layout.
#include <stdlib.h>

#include <cuda.h>
#include <cub/cub.cuh>

#include "common.cuh"

namespace rvo {

constexpr int kTcR = 3;            // lookup radius
constexpr int kTcWin = 8;          // raw window 8x8
constexpr int kTcTile = 16;        // tile edge (positions)
constexpr int kTcStep = 9;         // tile step = tile - window + 1
constexpr int kTcRows = 128;       // rows (edge, pixel, level) per CTA = MMA M
constexpr int kTcC = 128;          // channels = MMA K total
constexpr int kTcGroup = 56;       // output halves per (level, pixel) group: 7 rows of 8 (7 used)
constexpr int kTcThreads = 128;
constexpr int kTcMaxLevels = 2;

struct TcLevel {
  const __half* data;
  int N, H, W;
  int64_t sN, sH, sW;
  float scale;
  int TX, TY, binbase;
};

struct TcGeom {
  TcLevel lv[kTcMaxLevels];
  int nlevels, nbins;
};

__device__ __forceinline__ int tc_floor(float v) {
  float f = floorf(v);
  if (!(f > -1.0e6f)) f = -1.0e6f;
  if (f > 1.0e6f) f = 1.0e6f;
  return (int)f;
}

// ------------------------------------------------------------------ binning ----

// one thread per (edge, level): the 9 patch pixels of an edge almost always share one tile, so their
// histogram updates are aggregated into one atomicAdd per distinct bin (ranks stay unique per bin)
__global__ void __launch_bounds__(256)
tc_bin_count_kernel(TcGeom G, const float* __restrict__ coords, const int64_t* __restrict__ kk,
                    const int64_t* __restrict__ jj, int64_t pmod, int64_t fmod, int64_t gN, int E,
                    int32_t* __restrict__ cnt, int32_t* __restrict__ rowbin,
                    int32_t* __restrict__ rowrank, __half* __restrict__ out, int64_t out_ld) {
  const int NL = G.nlevels;
  const int64_t T = (int64_t)E * NL;
  for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < T;
       t += (int64_t)gridDim.x * blockDim.x) {
    const int lvl = (int)(t % NL);
    const int e = (int)(t / NL);
    const bool l1 = lvl != 0;
    const float scale = l1 ? G.lv[1].scale : G.lv[0].scale;
    const int LW = l1 ? G.lv[1].W : G.lv[0].W, LH = l1 ? G.lv[1].H : G.lv[0].H;
    const int LN = l1 ? G.lv[1].N : G.lv[0].N, TX = l1 ? G.lv[1].TX : G.lv[0].TX;
    const int TY = l1 ? G.lv[1].TY : G.lv[0].TY, base = l1 ? G.lv[1].binbase : G.lv[0].binbase;
    int64_t ip = kk[e], jf = jj[e];
    if (pmod > 0) ip %= pmod;
    if (fmod > 0) jf %= fmod;
    const bool idx_ok = ip >= 0 && ip < gN && jf >= 0 && jf < LN;
    int bins[9];
#pragma unroll
    for (int pix = 0; pix < 9; pix++) {
      const float x = coords[(int64_t)e * 18 + pix] * scale;
      const float y = coords[(int64_t)e * 18 + 9 + pix] * scale;
      const int x0 = tc_floor(x) - kTcR, y0 = tc_floor(y) - kTcR;
      const bool ok = idx_ok && x0 > -kTcWin && x0 < LW && y0 > -kTcWin && y0 < LH;
      bins[pix] = ok ? base + ((int)jf * TY + (y0 + kTcWin) / kTcStep) * TX + (x0 + kTcWin) / kTcStep : -1;
    }
    int ranks[9];
#pragma unroll
    for (int pix = 0; pix < 9; pix++) ranks[pix] = -1;
#pragma unroll
    for (int pix = 0; pix < 9; pix++) {
      if (bins[pix] < 0 || ranks[pix] >= 0) continue;
      int n = 0;
#pragma unroll
      for (int q = 0; q < 9; q++) n += (q >= pix && bins[q] == bins[pix]) ? 1 : 0;
      int r0 = atomicAdd(&cnt[bins[pix]], n);
#pragma unroll
      for (int q = 0; q < 9; q++)
        if (q >= pix && bins[q] == bins[pix]) ranks[q] = r0++;
    }
#pragma unroll
    for (int pix = 0; pix < 9; pix++) {
      const int64_t r = ((int64_t)e * 9 + pix) * NL + lvl;
      rowbin[r] = bins[pix];
      if (bins[pix] >= 0) {
        rowrank[r] = ranks[pix];
      } else {
        // window entirely outside the map (or invalid index): the 7x7 outputs are zero
        uint4* o = reinterpret_cast<uint4*>(out + (int64_t)e * out_ld + (lvl * 9 + pix) * kTcGroup);
#pragma unroll
        for (int q = 0; q < kTcGroup / 8; q++) o[q] = make_uint4(0u, 0u, 0u, 0u);
      }
    }
  }
}

struct __align__(16) TcHdr {
  int nrows, rbase, lvl, f, X0, Y0, flags, pad1;   // flags: bit 0 = same tile as the previous block, bit 1 = as the next
};

// one CTA: exclusive scans of the per-bin row counts and block counts (counts staged in shared
// memory with coalesced loads)
constexpr int kScanMaxBins = 49152;        // 192 KB of dynamic shared memory
__global__ void __launch_bounds__(1024)
tc_bin_scan_kernel(const int32_t* __restrict__ cnt, int nbins, int32_t* __restrict__ rowstart,
                   int32_t* __restrict__ blkstart, int32_t* __restrict__ total_blocks) {
  typedef cub::BlockScan<int, 1024> Scan;
  __shared__ typename Scan::TempStorage tmp;
  extern __shared__ int32_t c_s[];            // [nbins]
  for (int b = threadIdx.x; b < nbins; b += 1024) c_s[b] = cnt[b];
  __syncthreads();
  const int per = (nbins + 1023) / 1024;
  const int b0 = threadIdx.x * per, b1 = min(nbins, b0 + per);
  int rows = 0, blks = 0;
  for (int b = b0; b < b1; b++) {
    rows += c_s[b];
    blks += (c_s[b] + kTcRows - 1) / kTcRows;
  }
  int rpre, bpre, btot;
  Scan(tmp).ExclusiveSum(rows, rpre);
  __syncthreads();
  Scan(tmp).ExclusiveSum(blks, bpre, btot);
  for (int b = b0; b < b1; b++) {
    const int c = c_s[b];
    rowstart[b] = rpre;
    blkstart[b] = bpre;
    rpre += c;
    bpre += (c + kTcRows - 1) / kTcRows;
  }
  if (threadIdx.x == 0) total_blocks[0] = btot;
}

// one thread per tile: the headers of its blocks
__global__ void __launch_bounds__(256)
tc_block_hdr_kernel(TcGeom G, const int32_t* __restrict__ cnt, const int32_t* __restrict__ rowstart,
                    const int32_t* __restrict__ blkstart, TcHdr* __restrict__ hdr) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= G.nbins) return;
  const int c = cnt[b];
  if (c == 0) return;
  const int lvl = (G.nlevels > 1 && b >= G.lv[1].binbase) ? 1 : 0;
  const bool l1 = lvl != 0;
  const int TX = l1 ? G.lv[1].TX : G.lv[0].TX, TY = l1 ? G.lv[1].TY : G.lv[0].TY;
  int t = b - (l1 ? G.lv[1].binbase : G.lv[0].binbase);
  const int tx = t % TX; t /= TX;
  const int ty = t % TY;
  TcHdr h;
  h.lvl = lvl; h.f = t / TY;
  h.X0 = tx * kTcStep - kTcWin; h.Y0 = ty * kTcStep - kTcWin;
  h.pad1 = 0;
  const int rs = rowstart[b], bs = blkstart[b];
  for (int i = 0; i * kTcRows < c; i++) {
    h.nrows = min(kTcRows, c - i * kTcRows);
    h.rbase = rs + i * kTcRows;
    h.flags = (i > 0 ? 1 : 0) | ((i + 1) * kTcRows < c ? 2 : 0);
    hdr[bs + i] = h;
  }
}

// everything the tile kernel needs to know about a row, written in bin order by the scatter pass so
// that the main kernel does one coalesced 32-byte load per row instead of a chain of dependent loads
struct __align__(16) TcRow {
  long long src;     // element offset of the patch-pixel vector inside gmap
  long long out;     // element offset of the row's output group
  float dx, dy;      // bilinear weights
  int ox, oy;        // window origin inside the 16x16 tile, 0..8
};

__global__ void __launch_bounds__(256)
tc_bin_scatter_kernel(TcGeom G, const float* __restrict__ coords, const int64_t* __restrict__ kk,
                      int64_t pmod, int64_t g_sN, int64_t g_sH, int64_t g_sW, int64_t out_ld, int64_t R,
                      const int32_t* __restrict__ rowbin, const int32_t* __restrict__ rowrank,
                      const int32_t* __restrict__ rowstart, TcRow* __restrict__ rows) {
  const int NL = G.nlevels;
  for (int64_t r = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; r < R;
       r += (int64_t)gridDim.x * blockDim.x) {
    const int bin = rowbin[r];
    if (bin < 0) continue;
    const int lvl = (int)(r % NL);
    const int64_t ep = r / NL;
    const int pix = (int)(ep % 9);
    const int e = (int)(ep / 9);
    const TcLevel& L = G.lv[lvl];
    int64_t ip = kk[e];
    if (pmod > 0) ip %= pmod;
    const float x = coords[(int64_t)e * 18 + pix] * L.scale;
    const float y = coords[(int64_t)e * 18 + 9 + pix] * L.scale;
    const float fxf = floorf(x), fyf = floorf(y);
    TcRow rec;
    rec.src = ip * g_sN + (pix / 3) * g_sH + (pix % 3) * g_sW;
    rec.out = (long long)e * out_ld + (lvl * 9 + pix) * kTcGroup;
    rec.dx = x - fxf;
    rec.dy = y - fyf;
    rec.ox = ((int)fxf - kTcR + kTcWin) % kTcStep;
    rec.oy = ((int)fyf - kTcR + kTcWin) % kTcStep;
    rows[rowstart[bin] + rowrank[r]] = rec;
  }
}

// ------------------------------------------------------------------ tcgen05 helpers ----

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return (uint32_t)__cvta_generic_to_shared(p);
}

// K-major, SWIZZLE_128B operand tile: rows of 128 bytes, 8-row groups 1024 bytes apart
__device__ __forceinline__ uint64_t umma_desc(uint32_t smem_addr) {
  uint64_t d = (uint64_t)((smem_addr >> 4) & 0x3FFFu);   // start address, 16-byte units
  d |= (uint64_t)(1024u >> 4) << 32;                      // stride byte offset (8 rows)
  d |= (uint64_t)1 << 46;                                 // descriptor version (sm_100)
  d |= (uint64_t)2 << 61;                                 // SWIZZLE_128B
  return d;
}

// kind::f16, A/B = fp16 K-major, D = fp32, M = 128, N = 256
constexpr uint32_t kIdesc = (1u << 4) | ((256u >> 3) << 17) | ((128u >> 4) << 24);

__device__ __forceinline__ void umma_f16(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
      "}\n" ::"r"(tmem_d), "l"(da), "l"(db), "r"(kIdesc), "r"(accumulate)
      : "memory");
}

__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src, uint32_t src_bytes) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(dst), "l"(src), "r"(src_bytes)
               : "memory");
}

__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float* v) {
  uint32_t r[16];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];\n"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]),
        "=r"(r[15])
      : "r"(taddr)
      : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;\n" ::: "memory");
#pragma unroll
  for (int i = 0; i < 16; i++) v[i] = __uint_as_float(r[i]);
}

constexpr int kTcSmemA = 2 * kTcRows * 128;            // two K blocks of [128 rows x 128 B] = 32 KB
constexpr int kTcSmemB = 2 * 256 * 128;                // two K blocks of [256 rows x 128 B] = 64 KB
constexpr int kStage1 = 17;                            // floats per lane of the window-row stage

// ------------------------------------------------------------------ pipelined persistent variant ----
//
// Same math, canonical Blackwell structure: one persistent CTA per SM, warp-specialised, three
// pipelines through mbarriers —
//   producers (2 x 4 warps)  group s fills shared-memory stage s (cp.async A rows + B tile) for the
//                            blocks it == s (mod 2); the two groups run concurrently so one group's
//                            L2 latency hides behind the other's issue;
//   MMA issuer (1 warp)      one thread, 8 tcgen05.mma per block into one of two 256-column TMEM
//                            accumulators; tcgen05.commit releases the smem stage and publishes the
//                            accumulator;
//   epilogue (2 x 4 warps)   group s drains TMEM stage s: TMEM lane quadrant per warp, window
//                            extraction + blend + 16-byte stores while the next block is multiplied.
constexpr int kPipeThreads = 640;
constexpr int kPipeProducers = 128;                                      // per stage: warps 0-3 even blocks, 4-7 odd
constexpr int kPipeMmaWarp = 8;
constexpr int kPipeEpiWarp0 = 12;                                        // warps 12-15: even blocks, 16-19: odd
constexpr int kPipeStageBytes = kTcSmemA + kTcSmemB;                     // 96 KB
constexpr int kPipeEpiBytes = 2 * 128 * kStage1 * 4;                     // 17 408
constexpr int kPipeSmemBytes = 2 * kPipeStageBytes + kPipeEpiBytes + 1024;

__device__ __forceinline__ void mbar_init(uint64_t* b, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(smem_u32(b)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* b) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];\n" ::"r"(smem_u32(b)) : "memory");
}
// Waiting warps share issue slots with the producers: poll, then back off with nanosleep so that
// a spinning epilogue / MMA warp does not starve the warps that are doing the work.
__device__ __forceinline__ void mbar_wait(uint64_t* b, uint32_t parity) {
  const uint32_t addr = smem_u32(b);
  uint32_t done;
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t"
      "}\n"
      : "=r"(done)
      : "r"(addr), "r"(parity)
      : "memory");
  while (!done) {
    __nanosleep(64);
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t"
        "}\n"
        : "=r"(done)
        : "r"(addr), "r"(parity)
        : "memory");
  }
}
__device__ __forceinline__ void umma_commit(uint64_t* b) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n" ::"r"(
                   smem_u32(b))
               : "memory");
}

__device__ __forceinline__ TcHdr ld_hdr(const TcHdr* __restrict__ hdr, int b) {
  const uint4* p = reinterpret_cast<const uint4*>(hdr + b);
  const uint4 a = p[0], c = p[1];
  TcHdr h;
  h.nrows = (int)a.x; h.rbase = (int)a.y; h.lvl = (int)a.z; h.f = (int)a.w;
  h.X0 = (int)c.x; h.Y0 = (int)c.y; h.flags = (int)c.z; h.pad1 = 0;
  return h;
}

__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float* v) {
  uint32_t r[32];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];\n"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
        "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
        "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;\n" ::: "memory");
#pragma unroll
  for (int i = 0; i < 32; i++) v[i] = __uint_as_float(r[i]);
}

__global__ void __launch_bounds__(kPipeThreads, 1)
corr_tile_pipe_kernel(TcGeom G, const __half* __restrict__ gmap, const TcHdr* __restrict__ hdr,
                      const TcRow* __restrict__ rows, const int32_t* __restrict__ total_blocks,
                      __half* __restrict__ out) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ uint64_t full_bar[2], empty_bar[2], tfull_bar[2], tempty_bar[2];
  __shared__ uint32_t tmem_base_s;
  __shared__ long long rowsrc[2][2][kTcRows];   // [producer group][iteration parity][row]
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int nblk = total_blocks[0];

  if (tid == 0) {
    for (int s = 0; s < 2; s++) {
      mbar_init(&full_bar[s], kPipeProducers);
      mbar_init(&empty_bar[s], 1);
      mbar_init(&tfull_bar[s], 1);
      mbar_init(&tempty_bar[s], 128);
    }
    asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
  }
  if (warp == kPipeMmaWarp) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;\n" ::"r"(
                     smem_u32(&tmem_base_s))
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
  const uint32_t tmem_base = tmem_base_s;

  if (warp < 2 * kPipeProducers / 32) {
    // ===== producers: group pg fills stage pg; thread owns 16-byte chunk `ch` of rows / positions q0 + 8 j =====
    const int pg = warp >> 2, ptid = tid & (kPipeProducers - 1);
    const int ch = ptid & 15, q0 = ptid >> 4;                    // q0 in 0..7
    // swizzled chunk offset inside a 128-byte row: row & 7 == q0 for every row this thread touches
    const uint32_t swz = (uint32_t)(((ch & 7) ^ q0) << 4);
    const uint32_t kboff_a = (ch >> 3) * (kTcRows * 128), kboff_b = (ch >> 3) * (256 * 128);
    const int s = pg;
    TcHdr Hn;
    if ((int)blockIdx.x + pg * (int)gridDim.x < nblk) Hn = ld_hdr(hdr, blockIdx.x + pg * gridDim.x);
    for (int it = pg; blockIdx.x + it * (int)gridDim.x < nblk; it += 2) {
      const int b = blockIdx.x + it * gridDim.x;
      const TcHdr B = Hn;
      if (b + 2 * (int)gridDim.x < nblk) Hn = ld_hdr(hdr, b + 2 * gridDim.x);   // next header in flight
      const long long my_src = (ptid < B.nrows) ? rows[B.rbase + ptid].src : -1;
      mbar_wait(&empty_bar[s], ((it >> 1) & 1) ^ 1);        // first use of a stage passes immediately
      // level parameters into registers (a dynamically indexed struct would be re-read from the
      // constant bank on every use)
      const bool l1 = B.lvl != 0;
      const __half* ldata = l1 ? G.lv[1].data : G.lv[0].data;
      const int LH = l1 ? G.lv[1].H : G.lv[0].H, LW = l1 ? G.lv[1].W : G.lv[0].W;
      const int sH = (int)(l1 ? G.lv[1].sH : G.lv[0].sH), sW = (int)(l1 ? G.lv[1].sW : G.lv[0].sW);
      const int64_t sN = l1 ? G.lv[1].sN : G.lv[0].sN;
      const uint32_t As_u = smem_u32(smem + s * kPipeStageBytes);
      const uint32_t Bs_u = As_u + kTcSmemA;
      long long* rsrc = rowsrc[pg][(it >> 1) & 1];
      rsrc[ptid] = my_src;
      // B: positions p = q0 + 8 j (j < 32): py = j >> 1, px = q0 + 8 (j & 1).  Two fixed columns per
      // thread, sixteen rows; in-frame offsets are 32-bit (checked on the host).
      {
        const __half* fbase = ldata + (int64_t)B.f * sN + ch * 8;
        const int xa = B.X0 + q0, xb = xa + 8;
        const bool oka = (unsigned)xa < (unsigned)LW, okb = (unsigned)xb < (unsigned)LW;
        const int offa = oka ? xa * sW : 0, offb = okb ? xb * sW : 0;
        uint32_t dst = Bs_u + kboff_b + swz + q0 * 128;
        int y = B.Y0;
#pragma unroll 4
        for (int jy = 0; jy < 16; jy++, y++, dst += 2 * (8 * 128)) {
          const bool oky = (unsigned)y < (unsigned)LH;
          const int rowoff = oky ? y * sH : 0;
          cp_async16(dst, fbase + rowoff + offa, (oky && oka) ? 16u : 0u);
          cp_async16(dst + 8 * 128, fbase + rowoff + offb, (oky && okb) ? 16u : 0u);
        }
      }
      if (pg == 0) asm volatile("bar.sync 1, 128;\n" ::: "memory");
      else asm volatile("bar.sync 2, 128;\n" ::: "memory");
      // A: rows r = q0 + 8 j (j < 16)
      {
        uint32_t dst = As_u + kboff_a + swz + q0 * 128;
        const __half* gsrc = gmap + ch * 8;
        const long long* rs = rsrc + q0;
#pragma unroll 4
        for (int j = 0; j < 16; j++, dst += 8 * 128) {
          const long long off = rs[8 * j];
          const bool ok = off >= 0;
          cp_async16(dst, gsrc + (ok ? off : 0), ok ? 16u : 0u);
        }
      }
      asm volatile("cp.async.commit_group;\n" ::: "memory");
      asm volatile("cp.async.wait_group 0;\n" ::: "memory");
      asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");   // generic-proxy writes -> async proxy
      mbar_arrive(&full_bar[s]);
    }
  } else if (warp == kPipeMmaWarp) {
    // ===== MMA issuer =====
    int it = 0;
    for (int b = blockIdx.x; b < nblk; b += gridDim.x, it++) {
      const int s = it & 1;
      mbar_wait(&full_bar[s], (it >> 1) & 1);
      mbar_wait(&tempty_bar[s], ((it >> 1) & 1) ^ 1);
      asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
      if (lane == 0) {
        const uint32_t As_u = smem_u32(smem + s * kPipeStageBytes);
        const uint32_t Bs_u = As_u + kTcSmemA;
#pragma unroll
        for (int kb = 0; kb < 2; kb++)
#pragma unroll
          for (int k = 0; k < 4; k++)
            umma_f16(tmem_base + s * 256, umma_desc(As_u + kb * (kTcRows * 128) + k * 32),
                     umma_desc(Bs_u + kb * (256 * 128) + k * 32), (kb | k) ? 1u : 0u);
        umma_commit(&empty_bar[s]);     // smem stage may be refilled once these MMAs have read it
        umma_commit(&tfull_bar[s]);     // accumulator ready for the epilogue
      }
      __syncwarp();
    }
  } else if (warp >= kPipeEpiWarp0) {
    // ===== epilogue: group g (4 warps) drains TMEM stage g, i.e. blocks it == g (mod 2) =====
    const int g = (warp - kPipeEpiWarp0) >> 2;
    const int q = warp & 3;                               // TMEM lane quadrant of this warp
    const int row = q * 32 + lane;                        // accumulator row == TMEM lane
    float* S1 = reinterpret_cast<float*>(smem + 2 * kPipeStageBytes) + (g * 128 + row) * kStage1;
    const uint32_t tlane = tmem_base + g * 256 + ((uint32_t)(q * 32) << 16);
    TcHdr Hn;
    if ((int)blockIdx.x + g * (int)gridDim.x < nblk) Hn = ld_hdr(hdr, blockIdx.x + g * gridDim.x);
    for (int it = g; blockIdx.x + it * (int)gridDim.x < nblk; it += 2) {
      const int b = blockIdx.x + it * gridDim.x;
      const TcHdr B = Hn;
      if (b + 2 * (int)gridDim.x < nblk) Hn = ld_hdr(hdr, b + 2 * gridDim.x);
      int ox = 0, oy = 1 << 20;                           // inactive rows never match a window row
      float dx = 0.f, dy = 0.f;
      __half* orow = out;
      if (row < B.nrows) {
        const uint4* rp = reinterpret_cast<const uint4*>(rows + B.rbase + row);
        const uint4 r0 = rp[0], r1 = rp[1];
        orow = out + (((long long)r0.w << 32) | (long long)r0.z);
        dx = __uint_as_float(r1.x);
        dy = __uint_as_float(r1.y);
        ox = (int)r1.z;
        oy = (int)r1.w;
      }
      mbar_wait(&tfull_bar[g], (it >> 1) & 1);
      asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
      float hprev[7];
#pragma unroll
      for (int i = 0; i < 7; i++) hprev[i] = 0.f;
#pragma unroll 1
      for (int wy2 = 0; wy2 < kTcTile; wy2 += 2) {
        float v[32];
        tmem_ld32(tlane + wy2 * kTcTile, v);              // two tile rows per TMEM load
#pragma unroll
        for (int hh = 0; hh < 2; hh++) {
          const int a = wy2 + hh - oy;
          if (a >= 0 && a <= 7) {
#pragma unroll
            for (int i = 0; i < 16; i++) S1[i] = v[hh * 16 + i];
            float c[8], h[7];
#pragma unroll
            for (int i = 0; i < 8; i++) c[i] = S1[ox + i];
#pragma unroll
            for (int i = 0; i < 7; i++) h[i] = c[i] + dx * (c[i + 1] - c[i]);
            if (a >= 1) {
              float o[7];
#pragma unroll
              for (int i = 0; i < 7; i++) o[i] = hprev[i] + dy * (h[i] - hprev[i]);
              const __half2 p0 = __floats2half2_rn(o[0], o[1]), p1 = __floats2half2_rn(o[2], o[3]);
              const __half2 p2 = __floats2half2_rn(o[4], o[5]), p3 = __floats2half2_rn(o[6], 0.f);
              uint4 u;
              u.x = *reinterpret_cast<const uint32_t*>(&p0);
              u.y = *reinterpret_cast<const uint32_t*>(&p1);
              u.z = *reinterpret_cast<const uint32_t*>(&p2);
              u.w = *reinterpret_cast<const uint32_t*>(&p3);
              *reinterpret_cast<uint4*>(orow + (a - 1) * 8) = u;  // one aligned 16-byte store per (row, a)
            }
#pragma unroll
            for (int i = 0; i < 7; i++) hprev[i] = h[i];
          }
        }
      }
      asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
      mbar_arrive(&tempty_bar[g]);
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
  __syncthreads();
  if (warp == kPipeMmaWarp) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;\n" ::"r"(tmem_base) : "memory");
  }
}

// ------------------------------------------------------------------ TMA variant ----
//
// corr_tile_tma_kernel: the same block list, restructured around the memory system —
//   B tiles   one thread issues two cp.async.bulk.tensor (TMA) box loads per tile: the 4-D tensor map
//             over the channels-last frame ring {C, W, H, N} with box {64, 16, 16, 1} and
//             SWIZZLE_128B lands exactly in the UMMA canonical layout, and out-of-map positions
//             (negative coordinates included) are zero-filled by the hardware.  A tile is loaded
//             ONCE for all consecutive row blocks that share it (level 2 has ~5 blocks per tile);
//   chunks    a CTA takes chunks of kChunk consecutive blocks (so tile sharing survives the
//             persistent schedule) strided by the grid;
//   A rows    4 producer warps gather the 128 patch-pixel vectors with cp.async, two blocks in
//             flight (completion of block b is published after block b+1 has been issued);
//   MMA       as before, 8 x tcgen05.mma M128 N256 K16 per block into one of two TMEM accumulators;
//   epilogue  2 x 4 warps; warps whose 32 rows are all past nrows skip the drain, the window-row
//             staging uses 16-byte shared-memory stores.
constexpr int kChunkLog2 = 3;
constexpr int kChunk = 1 << kChunkLog2;
constexpr int kTmaThreads = 640;              // warps 0-7 A producers, 8 TMA, 9 MMA, 12-19 epilogue
constexpr int kTmaTmaWarp = 8;
constexpr int kTmaMmaWarp = 9;
constexpr int kTmaEpiWarp0 = 12;
constexpr int kTmaOffB = 0;                   // 2 x 64 KB
constexpr int kTmaOffA = 2 * kTcSmemB;        // 2 x 32 KB
constexpr int kTmaOffS = kTmaOffA + 2 * kTcSmemA;
constexpr int kStageV = 20;                   // floats per thread of the window-row stage (16 + pad, 16-B aligned)
constexpr int kTmaSmemBytes = kTmaOffS + 256 * kStageV * 4 + 1024;

__device__ __forceinline__ void mbar_expect_tx(uint64_t* b, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(smem_u32(b)), "r"(bytes)
               : "memory");
}

__device__ __forceinline__ void tma_load_4d(uint32_t dst, const void* tmap, int c0, int c1, int c2, int c3,
                                            uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, "
      "%5, %6}], [%2];\n" ::"r"(dst),
      "l"(reinterpret_cast<uint64_t>(tmap)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}

struct __align__(64) TcTmap {
  unsigned char bytes[128];                   // CUtensorMap
};

// block index of the i-th block of this CTA's schedule
__device__ __forceinline__ int tma_block(int i) {
  return ((int)blockIdx.x + (i >> kChunkLog2) * (int)gridDim.x) * kChunk + (i & (kChunk - 1));
}

__global__ void __launch_bounds__(kTmaThreads, 1)
corr_tile_tma_kernel(const __grid_constant__ TcTmap tm0, const __grid_constant__ TcTmap tm1,
                     const __half* __restrict__ gmap, const TcHdr* __restrict__ hdr,
                     const TcRow* __restrict__ rows, const int32_t* __restrict__ total_blocks,
                     __half* __restrict__ out) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ uint64_t bfull[2], bempty[2], afull[2], aempty[2], tfull[2], tempty[2];
  __shared__ uint32_t tmem_base_s;
  __shared__ long long rowsrc[2][kTcRows];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int nblk = total_blocks[0];

  if (tid == 0) {
    for (int s = 0; s < 2; s++) {
      mbar_init(&bfull[s], 1);
      mbar_init(&bempty[s], 1);
      mbar_init(&afull[s], 128);
      mbar_init(&aempty[s], 1);
      mbar_init(&tfull[s], 1);
      mbar_init(&tempty[s], 128);
    }
    asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
  }
  if (warp == kTmaMmaWarp) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;\n" ::"r"(
                     smem_u32(&tmem_base_s))
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
  const uint32_t tmem_base = tmem_base_s;

  if (warp < 8) {
    // ===== A producers: group pg (4 warps) fills stage pg for blocks i == pg (mod 2); the two groups
    // run concurrently so one block's gather is in flight while the other is published.  Thread owns
    // 16-byte chunk `ch` of rows q0 + 8 j =====
    const int pg = warp >> 2, ptid = tid & 127;
    const int ch = ptid & 15, q0 = ptid >> 4;
    const uint32_t swz = (uint32_t)(((ch & 7) ^ q0) << 4);
    const uint32_t kboff_a = (ch >> 3) * (kTcRows * 128);
    const __half* gsrc = gmap + ch * 8;
    const uint32_t dst0 = smem_u32(smem + kTmaOffA + pg * kTcSmemA) + kboff_a + swz + q0 * 128;
    TcHdr Hn;
    if (tma_block(pg) < nblk) Hn = ld_hdr(hdr, tma_block(pg));
    for (int i = pg;; i += 2) {
      const int b = tma_block(i);
      if (b >= nblk) break;
      const TcHdr B = Hn;
      if (tma_block(i + 2) < nblk) Hn = ld_hdr(hdr, tma_block(i + 2));
      const long long my_src = (ptid < B.nrows) ? rows[B.rbase + ptid].src : -1;
      mbar_wait(&aempty[pg], ((i >> 1) & 1) ^ 1);
      rowsrc[pg][ptid] = my_src;
      if (pg == 0) asm volatile("bar.sync 1, 128;\n" ::: "memory");
      else asm volatile("bar.sync 2, 128;\n" ::: "memory");
      uint32_t dst = dst0;
      const long long* rs = rowsrc[pg] + q0;
#pragma unroll 4
      for (int j = 0; j < 16; j++, dst += 8 * 128) {
        const long long off = rs[8 * j];
        const bool ok = off >= 0;
        cp_async16(dst, gsrc + (ok ? off : 0), ok ? 16u : 0u);
      }
      asm volatile("cp.async.commit_group;\n" ::: "memory");
      asm volatile("cp.async.wait_group 0;\n" ::: "memory");
      asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");   // generic-proxy writes -> async proxy
      mbar_arrive(&afull[pg]);
    }
  } else if (warp == kTmaTmaWarp) {
    // ===== TMA producer: one thread, one tile load per run of blocks that share the tile =====
    if (lane == 0) {
      int ib = 0;
      TcHdr Hn;
      if (tma_block(0) < nblk) Hn = ld_hdr(hdr, tma_block(0));
      for (int i = 0;; i++) {
        const int b = tma_block(i);
        if (b >= nblk) break;
        const TcHdr B = Hn;
        if (tma_block(i + 1) < nblk) Hn = ld_hdr(hdr, tma_block(i + 1));
        const bool newB = (i & (kChunk - 1)) == 0 || !(B.flags & 1);
        if (!newB) continue;
        const int s = ib & 1;
        mbar_wait(&bempty[s], ((ib >> 1) & 1) ^ 1);
        const void* tm = B.lvl ? (const void*)&tm1 : (const void*)&tm0;
        const uint32_t Bs_u = smem_u32(smem + kTmaOffB + s * kTcSmemB);
        mbar_expect_tx(&bfull[s], (uint32_t)kTcSmemB);
        tma_load_4d(Bs_u, tm, 0, B.X0, B.Y0, B.f, &bfull[s]);
        tma_load_4d(Bs_u + 256 * 128, tm, 64, B.X0, B.Y0, B.f, &bfull[s]);
        ib++;
      }
    }
  } else if (warp == kTmaMmaWarp) {
    // ===== MMA issuer =====
    int ib = -1;
    TcHdr Hn;
    if (tma_block(0) < nblk) Hn = ld_hdr(hdr, tma_block(0));
    for (int i = 0;; i++) {
      const int b = tma_block(i);
      if (b >= nblk) break;
      const TcHdr B = Hn;
      if (tma_block(i + 1) < nblk) Hn = ld_hdr(hdr, tma_block(i + 1));
      const bool newB = (i & (kChunk - 1)) == 0 || !(B.flags & 1);
      const bool lastB = (i & (kChunk - 1)) == kChunk - 1 || !(B.flags & 2);
      if (newB) {
        ib++;
        mbar_wait(&bfull[ib & 1], (ib >> 1) & 1);
      }
      const int s = i & 1, sb = ib & 1;
      mbar_wait(&afull[s], (i >> 1) & 1);
      mbar_wait(&tempty[s], ((i >> 1) & 1) ^ 1);
      asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
      if (lane == 0) {
        const uint32_t As_u = smem_u32(smem + kTmaOffA + s * kTcSmemA);
        const uint32_t Bs_u = smem_u32(smem + kTmaOffB + sb * kTcSmemB);
#pragma unroll
        for (int kb = 0; kb < 2; kb++)
#pragma unroll
          for (int k = 0; k < 4; k++)
            umma_f16(tmem_base + s * 256, umma_desc(As_u + kb * (kTcRows * 128) + k * 32),
                     umma_desc(Bs_u + kb * (256 * 128) + k * 32), (kb | k) ? 1u : 0u);
        umma_commit(&aempty[s]);
        umma_commit(&tfull[s]);
        if (lastB) umma_commit(&bempty[sb]);
      }
      __syncwarp();
    }
  } else if (warp >= kTmaEpiWarp0) {
    // ===== epilogue: group g (4 warps) drains TMEM accumulator g, i.e. blocks i == g (mod 2) =====
    const int g = (warp - kTmaEpiWarp0) >> 2;
    const int q = warp & 3;
    const int row = q * 32 + lane;
    float* S1 = reinterpret_cast<float*>(smem + kTmaOffS) + (g * 128 + row) * kStageV;
    const uint32_t tlane = tmem_base + g * 256 + ((uint32_t)(q * 32) << 16);
    TcHdr Hn;
    if (tma_block(g) < nblk) Hn = ld_hdr(hdr, tma_block(g));
    for (int i = g;; i += 2) {
      const int b = tma_block(i);
      if (b >= nblk) break;
      const TcHdr B = Hn;
      if (tma_block(i + 2) < nblk) Hn = ld_hdr(hdr, tma_block(i + 2));
      int ox = 0, oy = 1 << 20;
      float dx = 0.f, dy = 0.f;
      __half* orow = out;
      if (row < B.nrows) {
        const uint4* rp = reinterpret_cast<const uint4*>(rows + B.rbase + row);
        const uint4 r0 = rp[0], r1 = rp[1];
        orow = out + (((long long)r0.w << 32) | (long long)r0.z);
        dx = __uint_as_float(r1.x);
        dy = __uint_as_float(r1.y);
        ox = (int)r1.z;
        oy = (int)r1.w;
      }
      mbar_wait(&tfull[g], (i >> 1) & 1);
      asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
      if (q * 32 < B.nrows) {                               // warp-uniform: this warp owns live rows
        float hprev[7];
#pragma unroll
        for (int k = 0; k < 7; k++) hprev[k] = 0.f;
#pragma unroll 1
        for (int wy2 = 0; wy2 < kTcTile; wy2 += 2) {
          float v[32];
          tmem_ld32(tlane + wy2 * kTcTile, v);
#pragma unroll
          for (int hh = 0; hh < 2; hh++) {
            const int a = wy2 + hh - oy;
            if (a >= 0 && a <= 7) {
#pragma unroll
              for (int k = 0; k < 4; k++)
                *reinterpret_cast<float4*>(S1 + 4 * k) =
                    make_float4(v[hh * 16 + 4 * k], v[hh * 16 + 4 * k + 1], v[hh * 16 + 4 * k + 2],
                                v[hh * 16 + 4 * k + 3]);
              float c[8], h[7];
#pragma unroll
              for (int k = 0; k < 8; k++) c[k] = S1[ox + k];
#pragma unroll
              for (int k = 0; k < 7; k++) h[k] = c[k] + dx * (c[k + 1] - c[k]);
              if (a >= 1) {
                float o[7];
#pragma unroll
                for (int k = 0; k < 7; k++) o[k] = hprev[k] + dy * (h[k] - hprev[k]);
                const __half2 p0 = __floats2half2_rn(o[0], o[1]), p1 = __floats2half2_rn(o[2], o[3]);
                const __half2 p2 = __floats2half2_rn(o[4], o[5]), p3 = __floats2half2_rn(o[6], 0.f);
                uint4 u;
                u.x = *reinterpret_cast<const uint32_t*>(&p0);
                u.y = *reinterpret_cast<const uint32_t*>(&p1);
                u.z = *reinterpret_cast<const uint32_t*>(&p2);
                u.w = *reinterpret_cast<const uint32_t*>(&p3);
                *reinterpret_cast<uint4*>(orow + (a - 1) * 8) = u;
              }
#pragma unroll
              for (int k = 0; k < 7; k++) hprev[k] = h[k];
            }
          }
        }
      }
      asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
      mbar_arrive(&tempty[g]);
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
  __syncthreads();
  if (warp == kTmaMmaWarp) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;\n" ::"r"(tmem_base) : "memory");
  }
}

// ------------------------------------------------------------------ host ----

struct TcWs {
  int32_t *cnt, *rowstart, *blkstart, *total, *rowbin, *rowrank;
  TcRow* rows;
  TcHdr* hdr;
  size_t total_bytes;
  int64_t maxblocks;
};

static inline size_t al256(size_t v) { return (v + 255) / 256 * 256; }

static TcWs tc_layout(void* base, int64_t R, int nbins) {
  TcWs w;
  char* c = reinterpret_cast<char*>(base);
  size_t off = 0;
  auto take = [&](size_t bytes) { char* r = c + off; off += al256(bytes ? bytes : 4); return r; };
  w.maxblocks = R / kTcRows + nbins + 1;
  w.cnt = (int32_t*)take((size_t)nbins * 4);
  w.rowstart = (int32_t*)take((size_t)(nbins + 1) * 4);
  w.blkstart = (int32_t*)take((size_t)(nbins + 1) * 4);
  w.total = (int32_t*)take(16);
  w.rowbin = (int32_t*)take((size_t)R * 4);
  w.rowrank = (int32_t*)take((size_t)R * 4);
  w.rows = (TcRow*)take((size_t)R * sizeof(TcRow));
  w.hdr = (TcHdr*)take((size_t)w.maxblocks * sizeof(TcHdr));
  w.total_bytes = off;
  return w;
}

static int tc_geom(const rvo_fmap_t* pyr, const float* scale, int nlevels, TcGeom* G, const char* who) {
  RVO_CHECK_ARG(pyr && nlevels >= 1 && nlevels <= kTcMaxLevels, "%s: bad levels", who);
  G->nlevels = nlevels;
  int base = 0;
  for (int l = 0; l < nlevels; l++) {
    const rvo_fmap_t& p = pyr[l];
    RVO_CHECK_ARG(p.dtype == RVO_F16 && p.C == kTcC && p.sC == 1, "%s: level %d must be channels-last fp16 with 128 channels", who, l);
    RVO_CHECK_ARG((reinterpret_cast<uintptr_t>(p.data) & 15u) == 0 && p.sN % 8 == 0 && p.sH % 8 == 0 && p.sW % 8 == 0,
                  "%s: level %d is not 16-byte aligned", who, l);
    RVO_CHECK_ARG(p.sH >= 0 && p.sW >= 0 && (int64_t)p.H * p.sH + (int64_t)p.W * p.sW < 0x7fffffff,
                  "%s: level %d frame too large for 32-bit offsets", who, l);
    TcLevel& L = G->lv[l];
    L.data = (const __half*)p.data;
    L.N = p.N; L.H = p.H; L.W = p.W;
    L.sN = p.sN; L.sH = p.sH; L.sW = p.sW;
    L.scale = scale ? scale[l] : 1.0f;
    L.TX = (p.W + kTcWin - 1) / kTcStep + 1;
    L.TY = (p.H + kTcWin - 1) / kTcStep + 1;
    L.binbase = base;
    base += p.N * L.TX * L.TY;
  }
  for (int l = nlevels; l < kTcMaxLevels; l++) G->lv[l] = G->lv[0];
  G->nbins = base;
  return RVO_OK;
}

}  // namespace rvo


namespace rvo {

// cuTensorMapEncodeTiled through the runtime's driver entry point (no link-time libcuda dependency)
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*,
                                  CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion,
                                  CUtensorMapFloatOOBfill);

static EncodeTiledFn encode_tiled_fn() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  }
  return fn;
}

// 4-D map over a channels-last level {C = 128, W, H, N}; box = one K block of a tile: {64, 16, 16, 1}
static int tc_make_tmap(const TcLevel& L, TcTmap* out) {
  EncodeTiledFn enc = encode_tiled_fn();
  RVO_CHECK_ARG(enc != nullptr, "rvo_corr_tiles: cuTensorMapEncodeTiled is not available in this driver");
  static_assert(sizeof(CUtensorMap) == sizeof(TcTmap), "CUtensorMap size");
  const cuuint64_t dims[4] = {(cuuint64_t)kTcC, (cuuint64_t)L.W, (cuuint64_t)L.H, (cuuint64_t)L.N};
  const cuuint64_t strides[3] = {(cuuint64_t)L.sW * 2, (cuuint64_t)L.sH * 2, (cuuint64_t)L.sN * 2};
  const cuuint32_t box[4] = {64, (cuuint32_t)kTcTile, (cuuint32_t)kTcTile, 1};
  const cuuint32_t estr[4] = {1, 1, 1, 1};
  CUresult r = enc(reinterpret_cast<CUtensorMap*>(out), CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 4,
                   const_cast<__half*>(L.data), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  RVO_CHECK_ARG(r == CUDA_SUCCESS, "rvo_corr_tiles: cuTensorMapEncodeTiled failed (%d) for a level of %d x %d x %d",
                (int)r, L.N, L.H, L.W);
  return RVO_OK;
}

}  // namespace rvo

using namespace rvo;

extern "C" int64_t rvo_corr_tiles_ws_bytes(const rvo_fmap_t* pyr, int nlevels, int E) {
  TcGeom G;
  if (E < 0 || tc_geom(pyr, nullptr, nlevels, &G, "rvo_corr_tiles_ws_bytes") != RVO_OK) return -1;
  return (int64_t)tc_layout(nullptr, (int64_t)E * 9 * nlevels, G.nbins).total_bytes;
}

extern "C" int rvo_corr_tiles(const rvo_fmap_t* fmap1, const rvo_fmap_t* pyr, const float* scale,
                              int nlevels, const float* coords, const int64_t* kk, const int64_t* jj,
                              int64_t pmod, int64_t fmod, int E, void* out, int64_t out_ld, void* ws,
                              int64_t ws_bytes, void* stream) {
  RVO_CHECK_ARG(E >= 0, "rvo_corr_tiles: E=%d", E);
  if (E == 0) return RVO_OK;
  RVO_CHECK_ARG(fmap1 && fmap1->data && coords && kk && jj && out && ws, "rvo_corr_tiles: null pointer");
  RVO_CHECK_ARG(fmap1->dtype == RVO_F16 && fmap1->C == kTcC && fmap1->sC == 1 && fmap1->H == 3 && fmap1->W == 3,
                "rvo_corr_tiles: fmap1 must be channels-last fp16 [Np,128,3,3]");
  RVO_CHECK_ARG((reinterpret_cast<uintptr_t>(fmap1->data) & 15u) == 0 && fmap1->sN % 8 == 0 &&
                    fmap1->sH % 8 == 0 && fmap1->sW % 8 == 0, "rvo_corr_tiles: fmap1 alignment");
  TcGeom G;
  int rc = tc_geom(pyr, scale, nlevels, &G, "rvo_corr_tiles");
  if (rc != RVO_OK) return rc;
  RVO_CHECK_ARG(out_ld >= 9 * nlevels * kTcGroup && out_ld % 8 == 0 && (reinterpret_cast<uintptr_t>(out) & 15u) == 0,
                "rvo_corr_tiles: output rows need >= %d halves, stride %% 8 == 0, 16-byte alignment", 9 * nlevels * kTcGroup);
  const int64_t R = (int64_t)E * 9 * nlevels;
  RVO_CHECK_ARG(R < 0x7fffffff, "rvo_corr_tiles: too many rows");
  TcWs w = tc_layout(ws, R, G.nbins);
  RVO_CHECK_ARG((int64_t)w.total_bytes <= ws_bytes, "rvo_corr_tiles: workspace %lld < %lld bytes",
                (long long)ws_bytes, (long long)w.total_bytes);
  cudaStream_t st = (cudaStream_t)stream;
  RVO_CUDA(cudaMemsetAsync(w.cnt, 0, (size_t)G.nbins * 4, st));
  RVO_CHECK_ARG(G.nbins <= kScanMaxBins, "rvo_corr_tiles: %d tiles exceed the scan capacity", G.nbins);
  int grid = (int)((R + 255) / 256);
  if (grid > kNumSMs * 16) grid = kNumSMs * 16;
  tc_bin_count_kernel<<<(int)(((int64_t)E * nlevels + 255) / 256), 256, 0, st>>>(G, coords, kk, jj, pmod, fmod, fmap1->N, E, w.cnt, w.rowbin,
                                            w.rowrank, (__half*)out, out_ld);
  RVO_LAUNCH_CHECK("tc_bin_count_kernel");
  RVO_CUDA(cudaFuncSetAttribute(tc_bin_scan_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                kScanMaxBins * 4));
  tc_bin_scan_kernel<<<1, 1024, (size_t)G.nbins * 4, st>>>(w.cnt, G.nbins, w.rowstart, w.blkstart, w.total);
  RVO_LAUNCH_CHECK("tc_bin_scan_kernel");
  tc_block_hdr_kernel<<<(G.nbins + 255) / 256, 256, 0, st>>>(G, w.cnt, w.rowstart, w.blkstart, w.hdr);
  RVO_LAUNCH_CHECK("tc_block_hdr_kernel");
  tc_bin_scatter_kernel<<<grid, 256, 0, st>>>(G, coords, kk, pmod, fmap1->sN, fmap1->sH, fmap1->sW, out_ld, R,
                                              w.rowbin, w.rowrank, w.rowstart, w.rows);
  RVO_LAUNCH_CHECK("tc_bin_scatter_kernel");
  static const bool legacy = getenv("RVO_CORR_LEGACY") != nullptr;   // debugging aid: the cp.async-only kernel
  if (legacy) {
    RVO_CUDA(cudaFuncSetAttribute(corr_tile_pipe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kPipeSmemBytes));
    corr_tile_pipe_kernel<<<kNumSMs, kPipeThreads, kPipeSmemBytes, st>>>(G, (const __half*)fmap1->data, w.hdr,
                                                                         w.rows, w.total, (__half*)out);
    RVO_LAUNCH_CHECK("corr_tile_pipe_kernel");
    return RVO_OK;
  }
  TcTmap tm[kTcMaxLevels];
  for (int l = 0; l < kTcMaxLevels; l++) {
    rc = tc_make_tmap(G.lv[l], &tm[l]);
    if (rc != RVO_OK) return rc;
  }
  RVO_CUDA(cudaFuncSetAttribute(corr_tile_tma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kTmaSmemBytes));
  corr_tile_tma_kernel<<<kNumSMs, kTmaThreads, kTmaSmemBytes, st>>>(tm[0], tm[1], (const __half*)fmap1->data,
                                                                    w.hdr, w.rows, w.total, (__half*)out);
  RVO_LAUNCH_CHECK("corr_tile_tma_kernel");
  return RVO_OK;
}
