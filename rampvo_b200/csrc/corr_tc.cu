// corr_tc.cu — altcorr lookup as tile GEMMs on the 5th-gen tensor cores (tcgen05 + TMEM), sm_100a.
//
// The per-edge kernel (altcorr.cu) re-reads every 8x8x128 window from L2 once per patch pixel: 2.3 GB of
// L2->SM traffic per update for 256 MB of algorithmic bytes, and its mma.sync path tops out near
// 620 TFLOP/s.  Here the loop is turned inside out:
//
//   bin      every (edge, patch pixel, level) row is assigned to the 16x16-position tile of its target
//            frame that contains its 8x8 window (tiles step by 9 so every window fits in exactly one
//            tile) and, inside the tile, ordered by the tile row its window starts in: a histogram pass
//            (register-aggregated per edge) and a scatter pass whose CTAs each scan the tile table into
//            shared memory for themselves (no scan launch), on the device, no host sync;
//   GEMM     persistent CTAs walk the (tile, <=128 rows) blocks: A = the rows' 128-channel patch vectors
//            gathered with cp.async, B = the tile's 256 feature vectors brought by TMA (zero-filled
//            outside the map), both in the canonical K-major SWIZZLE_128B layout; ONE elected thread
//            issues 8 tcgen05.mma (M=128, N = 16 x covered tile rows <= 256, K=16) into a 128x256 fp32
//            accumulator in TMEM;
//   epilogue thread r owns TMEM lane r: tcgen05.ld brings two tile rows at a time, the thread picks its
//            own 8-wide window in registers, does the separable bilinear blend and writes each of its
//            7 output rows as one aligned 16-byte store.
//   (corr_tile_tma_kernel below documents the pipeline and what was measured to bound it.)
//
// Output layout ("tile layout", consumed by the update operator with a permuted first-layer weight):
//   out[e, ((lvl*9 + pix)*7 + a)*8 + b], a = y offset, b = x offset (0..6); b = 7 is a zero pad so that
//   every (row, a) is one aligned 16-byte store.  rvo_corr_pyramid keeps the reference's [E, 882] layout.
#include <stdlib.h>

#include <cuda.h>
#include <cub/cub.cuh>

#include "common.cuh"
#include "tcgen05.cuh"

namespace rvo {

constexpr int kTcR = 3;            // lookup radius
constexpr int kTcWin = 8;          // raw window 8x8
constexpr int kTcTile = 16;        // tile edge (positions)
constexpr int kTcStep = 9;         // tile step = tile - window + 1
constexpr int kTcRows = 128;       // rows (edge, pixel, level) per CTA = MMA M
constexpr int kTcC = 128;          // channels = MMA K total
constexpr int kTcGroup = 56;       // output halves per (level, pixel) group: 7 rows of 8 (7 used)
constexpr int kTcMaxLevels = 2;
constexpr int kTcSub = 12;         // counters per tile: one per window-origin row oy = 0..8 (12 keeps int4 alignment)

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

// index % mod for ring-buffer indices: 32-bit arithmetic whenever the operands allow it (a 64-bit
// modulo is ~100 instructions)
__device__ __forceinline__ int64_t tc_mod(int64_t v, int64_t mod) {
  if (mod <= 0) return v;
  if ((uint64_t)v < 0x80000000ull && mod < 0x80000000ll) return (int64_t)((uint32_t)v % (uint32_t)mod);
  return v % mod;
}

// Both binning passes run one thread per (edge, level) and see the edge's 9 patch pixels together: they almost
// always share a tile and fall into ~3 window-origin rows, so everything is aggregated in registers first.
struct TcEdgeBins {
  int sub[9];        // (tile, oy) sub-bin of every pixel's row, -1: window entirely outside the map / invalid index
  int tile[9];
  unsigned same[9];  // bit q: pixel q falls into the same sub-bin
  unsigned tfirst;   // bit p: pixel p is the first of its tile
};

__device__ __forceinline__ void tc_edge_bins(const TcGeom& G, int lvl, const float* __restrict__ cp, int64_t ip,
                                             int64_t jf, int64_t gN, TcEdgeBins& B, float (&xs)[9], float (&ys)[9]) {
  const bool l1 = lvl != 0;
  const float scale = l1 ? G.lv[1].scale : G.lv[0].scale;
  const int LW = l1 ? G.lv[1].W : G.lv[0].W, LH = l1 ? G.lv[1].H : G.lv[0].H;
  const int LN = l1 ? G.lv[1].N : G.lv[0].N, TX = l1 ? G.lv[1].TX : G.lv[0].TX;
  const int TY = l1 ? G.lv[1].TY : G.lv[0].TY, base = l1 ? G.lv[1].binbase : G.lv[0].binbase;
  const bool idx_ok = ip >= 0 && ip < gN && jf >= 0 && jf < LN;
#pragma unroll
  for (int pix = 0; pix < 9; pix++) {
    xs[pix] = cp[pix] * scale;
    ys[pix] = cp[9 + pix] * scale;
    const int x0 = tc_floor(xs[pix]) - kTcR, y0 = tc_floor(ys[pix]) - kTcR;
    const bool ok = idx_ok && x0 > -kTcWin && x0 < LW && y0 > -kTcWin && y0 < LH;
    // sub-bin = (tile, oy): rows of a tile end up sorted by the tile row their window starts in, so a
    // warp of the tile kernel's epilogue reads few more than 8 accumulator rows
    const int ty = (y0 + kTcWin) / kTcStep;
    B.tile[pix] = ok ? base + ((int)jf * TY + ty) * TX + (x0 + kTcWin) / kTcStep : -1;
    B.sub[pix] = ok ? B.tile[pix] * kTcSub + (y0 + kTcWin) - ty * kTcStep : -1;
  }
  // 36 comparisons each for the sub-bins and the tiles
  B.tfirst = 0x1ffu;
#pragma unroll
  for (int pix = 0; pix < 9; pix++) B.same[pix] = 1u << pix;
#pragma unroll
  for (int pix = 1; pix < 9; pix++)
#pragma unroll
    for (int q = 0; q < pix; q++) {
      if (B.sub[q] == B.sub[pix]) {
        B.same[pix] |= 1u << q;
        B.same[q] |= 1u << pix;
      }
      if (B.tile[q] == B.tile[pix]) B.tfirst &= ~(1u << pix);
    }
}

// histogram pass: one fire-and-forget add per distinct sub-bin and per distinct tile of the edge
__global__ void __launch_bounds__(256)
tc_bin_count_kernel(TcGeom G, const float* __restrict__ coords, const int64_t* __restrict__ kk,
                    const int64_t* __restrict__ jj, int64_t pmod, int64_t fmod, int64_t gN, int E,
                    int32_t* __restrict__ cnt, int32_t* __restrict__ tot, int32_t* __restrict__ rowsub,
                    int2* __restrict__ edge_pf) {
  const int NL = G.nlevels;
  const int T = E * NL;
  for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < T; t += gridDim.x * blockDim.x) {
    const int e = t / NL;
    const int lvl = t - e * NL;
    TcEdgeBins B;
    float xs[9], ys[9];
    const int64_t ip = tc_mod(kk[e], pmod), jf = tc_mod(jj[e], fmod);
    tc_edge_bins(G, lvl, coords + (int64_t)e * 18, ip, jf, gN, B, xs, ys);
    // what the row-parallel scatter pass would otherwise recompute per row: the rows' sub-bins and the edge's ring
    // indices (only read for rows with a sub-bin, whose indices are in range)
    if (lvl == 0) edge_pf[e] = make_int2((int)ip, (int)jf);
#pragma unroll
    for (int pix = 0; pix < 9; pix++) rowsub[(e * 9 + pix) * NL + lvl] = B.sub[pix];
#pragma unroll
    for (int pix = 0; pix < 9; pix++) {
      if (B.sub[pix] < 0) continue;
      if ((B.same[pix] & ((1u << pix) - 1u)) == 0) atomicAdd(&cnt[B.sub[pix]], __popc(B.same[pix]));
      if ((B.tfirst >> pix) & 1u) {
        int n = 0;
#pragma unroll
        for (int q = 0; q < 9; q++) n += (B.tile[q] == B.tile[pix]) ? 1 : 0;
        atomicAdd(&tot[B.tile[pix]], n);
      }
    }
  }
}

struct __align__(16) TcHdr {
  int nrows, rbase, lvl, f, X0, Y0, flags, oy_max;   // flags: bit 0 = same tile as the previous block, bit 1 = as the
};                                                   // next, bits 8.. = smallest window-origin row of the block's rows;
                                                     // oy_max = the largest (written by the block's LAST row)

// everything the tile kernel needs to know about a row, written in bin order by the scatter pass so
// that the main kernel does one coalesced 32-byte load per row instead of a chain of dependent loads
struct __align__(16) TcRow {
  long long src;     // element offset of the patch-pixel vector inside gmap
  long long out;     // element offset of the row's output group
  float dx, dy;      // bilinear weights
  int ox, oy;        // window origin inside the 16x16 tile, 0..8
};

// Exclusive scans of the per-tile row totals and 128-row block counts -> first row and first block of every tile.
// The table is small (~10 k tiles, 40 KB), so there is no separate scan launch: EVERY CTA of the scatter pass scans
// it for itself into shared memory (coalesced loads that hit the L2, two block scans) and then serves its rows'
// lookups from there — one launch and one dependent global round trip less than a scan kernel on a single SM.
constexpr int kScanMaxBins = 24576;          // 6 B x bins + 4 KB of dynamic shared memory
constexpr int kScatterThreads = 1024;

// scatter pass, one thread per row: a second set of counters (cursor, zeroed with cnt / tot) hands the row its slot
// inside its sub-bin — first row of the tile + rows of the lower sub-bins + cursor — and the thread writes the row's
// record there.  The row that opens a 128-row block also writes the block header.
__global__ void __launch_bounds__(kScatterThreads)
tc_bin_scatter_kernel(TcGeom G, const float* __restrict__ coords, const int32_t* __restrict__ rowsub,
                      const int2* __restrict__ edge_pf, int g_sN, int g_sH, int g_sW, int64_t out_ld, int64_t R,
                      const int32_t* __restrict__ cnt, int32_t* __restrict__ cursor, const int32_t* __restrict__ tot,
                      TcRow* __restrict__ rows, TcHdr* __restrict__ hdr, int32_t* __restrict__ total_blocks,
                      __half* __restrict__ out) {
  typedef cub::BlockScan<int, kScatterThreads, cub::BLOCK_SCAN_WARP_SCANS> Scan;
  __shared__ typename Scan::TempStorage tmp_r, tmp_b;
  extern __shared__ int32_t scan_sm[];
  const int nbins = G.nbins;
  int32_t* bfirst_s = scan_sm;                               // [threads] first block of the thread's first tile
  int32_t* r_s = scan_sm + kScatterThreads;                  // [nbins + 1] row totals -> first row of the tile
  uint16_t* b_s = reinterpret_cast<uint16_t*>(r_s + nbins + 1);   // [nbins] first block, relative to bfirst_s
  for (int b = threadIdx.x; b < nbins; b += kScatterThreads) r_s[b] = tot[b];
  __syncthreads();
  // thread t owns tiles t*per .. t*per + per - 1
  const int per = (nbins + kScatterThreads - 1) / kScatterThreads;
  {
    const int b0 = threadIdx.x * per;
    int nrow = 0, blks = 0;
    for (int i = 0; i < per; i++) {
      const int c = (b0 + i < nbins) ? r_s[b0 + i] : 0;
      nrow += c;
      blks += (c + kTcRows - 1) / kTcRows;
    }
    int rpre, bpre, rtot, btot;
    Scan(tmp_r).ExclusiveSum(nrow, rpre, rtot);
    Scan(tmp_b).ExclusiveSum(blks, bpre, btot);
    if (threadIdx.x == 0) r_s[nbins] = rtot;
    if (blockIdx.x == 0 && threadIdx.x == 0) total_blocks[0] = btot;
    // in place: counts -> exclusive row offsets; block offsets relative to the thread's first block fit 16 bits
    bfirst_s[threadIdx.x] = bpre;
    int brel = 0;
    for (int i = 0; i < per; i++) {
      if (b0 + i < nbins) {
        const int c = r_s[b0 + i];
        r_s[b0 + i] = rpre;
        b_s[b0 + i] = (uint16_t)brel;
        rpre += c;
        brel += (c + kTcRows - 1) / kTcRows;
      }
    }
  }
  __syncthreads();
  const int NL = G.nlevels;
  const int Ri = (int)R;                                     // < 2^31 (checked on the host)
  for (int r = blockIdx.x * blockDim.x + threadIdx.x; r < Ri; r += gridDim.x * blockDim.x) {
    const int sub = rowsub[r];
    const bool ok = sub >= 0;
    const int ep = NL == 2 ? r >> 1 : r / NL;
    const int lvl = r - ep * NL;
    const int e = ep / 9;
    const int pix = ep - e * 9;
    // neighbouring lanes are the pixels of one edge and mostly share a sub-bin: one atomic per distinct sub-bin of
    // the warp, its lanes take consecutive slots
    const unsigned act = __activemask();
    const unsigned lane = threadIdx.x & 31;
    const unsigned peers = __match_any_sync(act, ok ? sub : -1 - (int)lane);
    const int leader = __ffs(peers) - 1;
    int rank = 0;
    if (ok && (int)lane == leader) rank = atomicAdd(&cursor[sub], __popc(peers));
    rank = __shfl_sync(act, rank, leader) + __popc(peers & ((1u << lane) - 1u));
    const long long out_off = (long long)e * out_ld + (lvl * 9 + pix) * kTcGroup;
    if (!ok) {
      // window entirely outside the map (or invalid index): the 7x7 outputs are zero
      uint4* o = reinterpret_cast<uint4*>(out + out_off);
#pragma unroll
      for (int q = 0; q < kTcGroup / 8; q++) o[q] = make_uint4(0u, 0u, 0u, 0u);
      continue;
    }
    const int tile = sub / kTcSub, oy = sub - tile * kTcSub;
    const float scale = lvl ? G.lv[1].scale : G.lv[0].scale;
    const float x = coords[e * 18 + pix] * scale;
    const float y = coords[e * 18 + 9 + pix] * scale;
    const float fxf = floorf(x), fyf = floorf(y);
    const int wx = (int)fxf - kTcR + kTcWin, wy = (int)fyf - kTcR + kTcWin;   // > 0: the row has a sub-bin
    const int ox = wx % kTcStep;
    const int2 pf = edge_pf[e];
    TcRow rec;
    rec.src = pf.x * g_sN + (pix / 3) * g_sH + (pix % 3) * g_sW;               // < 2^31 (checked on the host)
    rec.out = out_off;
    rec.dx = x - fxf;
    rec.dy = y - fyf;
    rec.ox = ox;
    rec.oy = oy;
    const int4* c4 = reinterpret_cast<const int4*>(cnt) + tile * 3;
    const int4 v0 = c4[0], v1 = c4[1];
    // rows of a tile are ordered by oy: first row of the tile + rows of the lower sub-bins + rank
    const int rs = r_s[tile];
    int pos = rs + rank;
    pos += (oy > 0 ? v0.x : 0) + (oy > 1 ? v0.y : 0) + (oy > 2 ? v0.z : 0) + (oy > 3 ? v0.w : 0);
    pos += (oy > 4 ? v1.x : 0) + (oy > 5 ? v1.y : 0) + (oy > 6 ? v1.z : 0) + (oy > 7 ? v1.w : 0);
    rows[pos] = rec;
    // the row that opens a 128-row block of its tile also writes the block's header
    const int k = pos - rs;
    const int total = r_s[tile + 1] - rs;
    if ((k & (kTcRows - 1)) == 0) {
      // (field by field: oy_max of the same header belongs to another thread)
      int* h = reinterpret_cast<int*>(hdr + (bfirst_s[tile / per] + b_s[tile] + k / kTcRows));
      *reinterpret_cast<int4*>(h) = make_int4(min(kTcRows, total - k), pos, lvl, pf.y);
      *reinterpret_cast<int2*>(h + 4) = make_int2(wx - ox - kTcWin, wy - oy - kTcWin);
      h[6] = (k > 0 ? 1 : 0) | (k + kTcRows < total ? 2 : 0) | (oy << 8);
    }
    // rows of a block are ordered by oy: its last row knows the largest one, which bounds the accumulator rows
    // (= MMA N) the block needs
    if ((k & (kTcRows - 1)) == kTcRows - 1 || k == total - 1)
      reinterpret_cast<int*>(hdr + (bfirst_s[tile / per] + b_s[tile] + k / kTcRows))[7] = oy;
  }
}

// ------------------------------------------------------------------ tcgen05 helpers ----

constexpr int kTcSmemA = 2 * kTcRows * 128;            // two K blocks of [128 rows x 128 B] = 32 KB
constexpr int kTcSmemB = 2 * 256 * 128;                // two K blocks of [256 rows x 128 B] = 64 KB

// ------------------------------------------------------------------ mbarrier / TMEM helpers ----

__device__ __forceinline__ TcHdr ld_hdr(const TcHdr* __restrict__ hdr, int b) {
  const uint4* p = reinterpret_cast<const uint4*>(hdr + b);
  const uint4 a = p[0], c = p[1];
  TcHdr h;
  h.nrows = (int)a.x; h.rbase = (int)a.y; h.lvl = (int)a.z; h.f = (int)a.w;
  h.X0 = (int)c.x; h.Y0 = (int)c.y; h.flags = (int)c.z; h.oy_max = (int)c.w;
  return h;
}

// ------------------------------------------------------------------ TMA variant ----
//
// corr_tile_tma_kernel: the same block list, restructured around what bounds it.  Measured on B200:
// a tcgen05.mma M128 N256 K16 with both operands in shared memory takes ~200 cycles (operand fetch,
// ~60 B/clk), so a block costs ~1600 tensor cycles and its 64 KB tile ~40 B/clk/SM of L2 bandwidth;
// everything else has to hide behind those two.
//   B tiles   one thread issues two cp.async.bulk.tensor (TMA) box loads per tile: the 4-D tensor map
//             over the channels-last frame ring {C, W, H, N} with box {64, 16, 16, 1} and
//             SWIZZLE_128B lands exactly in the UMMA canonical layout, and out-of-map positions
//             (negative coordinates included) are zero-filled by the hardware.  A tile is loaded
//             ONCE for all consecutive row blocks that share it (level 2 has ~5 blocks per tile);
//   chunks    a CTA takes chunks of consecutive blocks (so tile sharing survives the persistent
//             schedule) strided by the grid, 8 blocks or 1 block long (tma_block);
//   A rows    two producer groups (4 warps each, alternate blocks) gather the 128 patch-pixel
//             vectors with cp.async into a 3-stage ring: the gather of block i starts when the MMAs
//             of block i - 3 have retired;
//   rows      row r of a block sits in TMEM lane r; odd blocks are rotated by 64 lanes so that the
//             half-empty level-1 blocks keep all four schedulers busy;
//   MMA       8 x tcgen05.mma M128 N256 K16 per block into one of two TMEM accumulators;
//   epilogue  2 accumulators x 4 lane quadrants = 8 warps.  Rows of a tile are sorted by window-origin
//             row, so a warp reads only the accumulator rows its windows cover, two rows per tcgen05.ld.
//             The body is one branch-free basic block (only the 16-byte stores are predicated): the
//             per-thread x offset is applied by 19 selects for its high bits and a 5-tap blend for the
//             low bits — the half-rate ALU pipe (FSEL) is what bounds the epilogue, so work is moved to
//             the FMA pipe — followed by the vertical blend and the fp16 pack.
constexpr int kTmaThreads = 640;              // warps 0-7 A producers, 8 TMA, 9 MMA, 12-19 epilogue
constexpr int kTmaTmaWarp = 8;
constexpr int kTmaMmaWarp = 9;
constexpr int kTmaEpiWarp0 = 12;
constexpr int kTmaAStages = 3;
constexpr int kTmaOffB = 0;                   // 2 x 64 KB
constexpr int kTmaOffA = 2 * kTcSmemB;        // 3 x 32 KB
constexpr int kTmaSmemBytes = kTmaOffA + kTmaAStages * kTcSmemA + 1024;

// Persistent schedule: a CTA takes chunks of 2^CL consecutive blocks, strided by the grid.  A chunk keeps the blocks
// of one tile together (its features are loaded once per run of blocks), which pays when the tiles hold several
// blocks each (precise.yaml: chunks of 8, 836 vs 849 us); with about one block per level-1 tile (default.yaml) short
// chunks mix cheap and expensive blocks evenly over the CTAs instead (µs, chunks of 8 / 4 / 2 / 1: 101 / 99 / 95-97 /
// 96 on a uniform synthetic graph, 105.7 / - / 99-101 / 98.6 on the running VO's graph).  The host picks 8 or 1.
// block index of the i-th block of this CTA's schedule
template <int CL>
__device__ __forceinline__ int tma_block(int i) {
  return (((int)blockIdx.x + (i >> CL) * (int)gridDim.x) << CL) + (i & ((1 << CL) - 1));
}

// row of a block that lives in TMEM lane (quadrant q, lane l).
//   more than 64 rows: row r sits in lane r (fewest partially filled warps — the epilogue is instruction-bound); odd
//     blocks of a CTA's schedule (rot = 1) are rotated by 64 lanes, so that partly filled blocks load the four
//     schedulers evenly (lane quadrant == scheduler: even blocks fill quadrants 0,1 first, odd blocks 2,3);
//   at most 64 rows (the typical level-1 block, ~52 rows): every row sits in TWO lanes, r and r + 64 — the M = 128
//     MMA computes the copy for free — and the epilogue warps of quadrants q and q + 2 each take half of the
//     accumulator rows the windows cover.  The epilogue of a block is a latency chain of one warp per quadrant
//     (measured 2 500-3 000 cycles for a 26-row warp against 1 350 for the block's MMAs, and the two-deep
//     accumulator ring makes the MMAs of block i + 2 wait for it); halving the chain on otherwise idle
//     schedulers is what shortens the period.
constexpr int kTcDupRows = 64;
__device__ __forceinline__ int tma_row_of(int q, int l, int nrows, int rot) {
  if (nrows <= kTcDupRows) {
    const int r = (q * 32 + l) & 63;
    return r < nrows ? r : -1;
  }
  const int r = (q * 32 + l + 64 * rot) & 127;
  return r < nrows ? r : -1;
}

// h[k] = lerp_x(v[ox + k], v[ox + k + 1]), k = 0..6, for a per-thread window offset ox = 0..8.
// The epilogue is bound by the half-rate ALU pipe (FSEL), so only bits 8 and 4 of ox go through
// selects (19 FSEL); the low two bits are folded into the horizontal blend as a 5-tap filter with
// per-thread weights w[j] = (j == s)(1 - dx) + (j == s + 1) dx, s = ox & 3 — 35 FMUL/FFMA on the FMA pipe
// instead of 17 FSEL + 14 FADD/FFMA.  Zero taps contribute exact zeros.
__device__ __forceinline__ void epi_hrow(const float* v, bool p8, bool p4, const float (&w)[5], float (&h)[7]) {
  float t0[15], t1[11];
#pragma unroll
  for (int k = 0; k < 8; k++) t0[k] = p8 ? v[k + 8] : v[k];
#pragma unroll
  for (int k = 8; k < 15; k++) t0[k] = v[k];
#pragma unroll
  for (int k = 0; k < 11; k++) t1[k] = p4 ? t0[k + 4] : t0[k];
#pragma unroll
  for (int k = 0; k < 7; k++) {
    float acc = w[0] * t1[k];
    acc = fmaf(w[1], t1[k + 1], acc);
    acc = fmaf(w[2], t1[k + 2], acc);
    acc = fmaf(w[3], t1[k + 3], acc);
    h[k] = fmaf(w[4], t1[k + 4], acc);
  }
}

// o = lerp_y(h0, h1) rounded to fp16: 7 values + zero pad = one aligned 16-byte vector
__device__ __forceinline__ uint4 epi_pack(const float (&h0)[7], const float (&h1)[7], float dy) {
  float o[7];
#pragma unroll
  for (int k = 0; k < 7; k++) o[k] = h0[k] + dy * (h1[k] - h0[k]);
  const __half2 q0 = __floats2half2_rn(o[0], o[1]), q1 = __floats2half2_rn(o[2], o[3]);
  const __half2 q2 = __floats2half2_rn(o[4], o[5]), q3 = __floats2half2_rn(o[6], 0.f);
  uint4 u;
  u.x = *reinterpret_cast<const uint32_t*>(&q0);
  u.y = *reinterpret_cast<const uint32_t*>(&q1);
  u.z = *reinterpret_cast<const uint32_t*>(&q2);
  u.w = *reinterpret_cast<const uint32_t*>(&q3);
  return u;
}

// debugging aids, env RVO_CORR_DBG (bit mask, 0 in production; results are WRONG with bits 1-8 set — they exist to
// time the pipeline's parts in isolation, tools/corr_bench.py): 1 = epilogue reads TMEM but skips the arithmetic and
// stores, 2 = epilogue only hands the accumulator back, 4 = no A gathers, 8 = no TMA tile loads, 16 = per-block
// clock64 stamps of CTA 0 (read back with rvo_corr_trace, printed by `tools/corr_bench.py trace`)
__device__ long long g_tc_trace[8 * 256];
#define TC_TRACE(slot, i) do { if ((dbg & 16) && blockIdx.x == 0 && (i) < 256) g_tc_trace[(slot) * 256 + (i)] = clock64(); } while (0)

template <int CL>
__global__ void __launch_bounds__(kTmaThreads, 1)
corr_tile_tma_kernel(const __grid_constant__ TcTmap tm0, const __grid_constant__ TcTmap tm1,
                     const __half* __restrict__ gmap, const TcHdr* __restrict__ hdr,
                     const TcRow* __restrict__ rows, const int32_t* __restrict__ total_blocks,
                     __half* __restrict__ out, int dbg_arg) {
#ifdef RVO_DEBUG
  const int dbg = dbg_arg;          // -DRVO_DEBUG builds only (python -m rampvo_b200.build --debug)
#else
  constexpr int dbg = 0;            // release library: every debug branch below is compiled out
#endif
  extern __shared__ uint8_t smem_raw[];
  __shared__ uint64_t bfull[2], bempty[2], afull[kTmaAStages], aempty[kTmaAStages], tfull[2], tempty[2];
  __shared__ uint32_t tmem_base_s;
  __shared__ int rowsrc[2][kTcRows];            // per producer group: gmap element offsets (< 2^31, host-checked)
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int nblk = total_blocks[0];

  if (tid == 0) {
    for (int s = 0; s < 2; s++) {
      mbar_init(&bfull[s], 1);
      mbar_init(&bempty[s], 1);
      mbar_init(&tfull[s], 1);
      mbar_init(&tempty[s], 128);
    }
    for (int s = 0; s < kTmaAStages; s++) {
      mbar_init(&afull[s], 128);
      mbar_init(&aempty[s], 1);
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
    // ===== A producers: group pg takes blocks i == pg (mod 2), stage i % 3.  Thread owns 16-byte chunk
    // `ch` of the M slots q0 + 8 j; headers run two own-blocks ahead and the row sources one ahead of
    // the gather, so neither L2 round trip sits on the per-block critical path =====
    const int pg = warp >> 2, ptid = tid & 127;
    const int ch = ptid & 15, q0 = ptid >> 4;
    const uint32_t swz = (uint32_t)(((ch & 7) ^ q0) << 4);
    const uint32_t kboff_a = (ch >> 3) * (kTcRows * 128);
    const __half* gsrc = gmap + ch * 8;
    const uint32_t dst0 = smem_u32(smem + kTmaOffA) + kboff_a + swz + q0 * 128;
    const int my_q = ptid >> 5, my_l = ptid & 31;             // the M slot whose source this thread fetches
    TcHdr H1, H2;
    int src_n = -1;
    if (tma_block<CL>(pg) < nblk) {
      H1 = ld_hdr(hdr, tma_block<CL>(pg));
      const int r = tma_row_of(my_q, my_l, H1.nrows, pg);
      if (r >= 0) src_n = (int)rows[H1.rbase + r].src;
    }
    if (tma_block<CL>(pg + 2) < nblk) H1 = ld_hdr(hdr, tma_block<CL>(pg + 2));
    for (int i = pg;; i += 2) {
      const int b = tma_block<CL>(i);
      if (b >= nblk) break;
      const int my_src = src_n;
      src_n = -1;
      if (tma_block<CL>(i + 4) < nblk) H2 = ld_hdr(hdr, tma_block<CL>(i + 4));
      if (tma_block<CL>(i + 2) < nblk) {
        const int r = tma_row_of(my_q, my_l, H1.nrows, pg);
        if (r >= 0) src_n = (int)rows[H1.rbase + r].src;
      }
      H1 = H2;
      const int s = i % kTmaAStages, ph = (i / kTmaAStages) & 1;
      mbar_wait(&aempty[s], ph ^ 1);
      if (ptid == 0) TC_TRACE(0, i);
      rowsrc[pg][ptid] = my_src;
      if (pg == 0) asm volatile("bar.sync 1, 128;\n" ::: "memory");
      else asm volatile("bar.sync 2, 128;\n" ::: "memory");
      uint32_t dst = dst0 + s * kTcSmemA;
      const int* rs = rowsrc[pg] + q0;
#pragma unroll 4
      for (int j = 0; j < 16; j++, dst += 8 * 128) {
        const int off = rs[8 * j];
        const bool ok = off >= 0;
        if (!(dbg & 4)) cp_async16(dst, gsrc + (ok ? off : 0), ok ? 16u : 0u);
      }
      // hardware-triggered arrival when this thread's copies have landed: the producer goes on to the next block's
      // gather instead of sitting out a memory round trip per block (the consumer issues the proxy fence)
      asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];\n" ::"r"(smem_u32(&afull[s])) : "memory");
      if (ptid == 0) TC_TRACE(1, i);
      // the group's next write of rowsrc[pg] comes after its next bar.sync partner has read this one:
      // every thread passes the loop above before any thread can pass the next iteration's barrier
      if (pg == 0) asm volatile("bar.sync 3, 128;\n" ::: "memory");
      else asm volatile("bar.sync 4, 128;\n" ::: "memory");
    }
    asm volatile("cp.async.wait_all;\n" ::: "memory");
  } else if (warp == kTmaTmaWarp) {
    // ===== TMA producer: one thread, one tile load per run of blocks that share the tile =====
    if (lane == 0) {
      int ib = 0;
      TcHdr Hn;
      if (tma_block<CL>(0) < nblk) Hn = ld_hdr(hdr, tma_block<CL>(0));
      for (int i = 0;; i++) {
        const int b = tma_block<CL>(i);
        if (b >= nblk) break;
        const TcHdr B = Hn;
        if (tma_block<CL>(i + 1) < nblk) Hn = ld_hdr(hdr, tma_block<CL>(i + 1));
        const bool newB = (i & ((1 << CL) - 1)) == 0 || !(B.flags & 1);
        if (!newB) continue;
        const int s = ib & 1;
        mbar_wait(&bempty[s], ((ib >> 1) & 1) ^ 1);
        const void* tm = B.lvl ? (const void*)&tm1 : (const void*)&tm0;
        const uint32_t Bs_u = smem_u32(smem + kTmaOffB + s * kTcSmemB);
        if (dbg & 8) {
          mbar_arrive(&bfull[s]);
        } else {
          mbar_expect_tx(&bfull[s], (uint32_t)kTcSmemB);
          tma_load_4d(Bs_u, tm, 0, B.X0, B.Y0, B.f, &bfull[s]);
          tma_load_4d(Bs_u + 256 * 128, tm, 64, B.X0, B.Y0, B.f, &bfull[s]);
        }
        ib++;
      }
    }
  } else if (warp == kTmaMmaWarp) {
    // ===== MMA issuer: ONE thread runs the whole loop (the other lanes of the warp go straight to the final barrier).
    // The tensor pipe executes an M128 N256 K16 step at its 128-cycle floor, but its queue is shallow — the thread is
    // blocked in the issue of a block until the previous block's steps have executed — so whatever the thread spends
    // between two blocks (barrier waits, fences, warp re-convergence) leaves the pipe idle. =====
    if (lane == 0) {
      int ib = -1;
      TcHdr Hn;
      if (tma_block<CL>(0) < nblk) Hn = ld_hdr(hdr, tma_block<CL>(0));
      int sa = 0, pa = 0;
      for (int i = 0;; i++) {
        const int b = tma_block<CL>(i);
        if (b >= nblk) break;
        const TcHdr B = Hn;
        if (tma_block<CL>(i + 1) < nblk) Hn = ld_hdr(hdr, tma_block<CL>(i + 1));
        const bool newB = (i & ((1 << CL) - 1)) == 0 || !(B.flags & 1);
        const bool lastB = (i & ((1 << CL) - 1)) == (1 << CL) - 1 || !(B.flags & 2);
        TC_TRACE(2, i);
        // descriptors before the waits: nothing but the MMAs themselves follows the last barrier.  One descriptor per
        // stage, advanced by constants (the address field counts 16-byte units and cannot carry out of its 14 bits
        // below 256 KB).  Only the tile rows the block's windows cover are computed (rows are ordered by
        // window-origin row): accumulator column 0 is tile row wbase, N = 16 positions per covered row.
        if (newB) ib++;
        const int st = i & 1, sb = ib & 1;
        const uint32_t As_u = smem_u32(smem + kTmaOffA + sa * kTcSmemA);
        const uint32_t Bs_u = smem_u32(smem + kTmaOffB + sb * kTcSmemB);
        const int wbase = (B.flags >> 8) & ~1;
        const int nt = B.oy_max + kTcWin - wbase;
        const uint64_t da0 = umma_desc(As_u), db0 = umma_desc(Bs_u + wbase * (kTcTile * 128));
        const uint32_t idesc = umma_idesc_f16(128, nt * kTcTile);
        const uint32_t dcol = tmem_base + st * 256;
        mbar_wait_spin(&tempty[st], ((i >> 1) & 1) ^ 1);
        if (newB) mbar_wait_spin(&bfull[sb], (ib >> 1) & 1);
        mbar_wait_spin(&afull[sa], pa);
        asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");   // A rows: cp.async writes -> tensor-core reads
        asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
        TC_TRACE(3, i);
#pragma unroll
        for (int kb = 0; kb < 2; kb++)
#pragma unroll
          for (int k = 0; k < 4; k++)
            umma_f16(dcol, da0 + (uint64_t)((kb * (kTcRows * 128) + k * 32) >> 4),
                     db0 + (uint64_t)((kb * (256 * 128) + k * 32) >> 4), idesc, (kb | k) ? 1u : 0u);
        umma_commit(&aempty[sa]);
        umma_commit(&tfull[st]);
        if (lastB) umma_commit(&bempty[sb]);
        TC_TRACE(4, i);
        if (++sa == kTmaAStages) { sa = 0; pa ^= 1; }
      }
    }
  } else if (warp >= kTmaEpiWarp0) {
    // ===== epilogue: accumulator g, TMEM lane quadrant q; this group's blocks are i = g, g + 2, ... =====
    const int ew = warp - kTmaEpiWarp0;
    const int g = ew >> 2, q = warp & 3;
    // window rows a_first .. a_last of every live row are this warp's: it stores output rows a_first .. a_last - 1
    // (measured: splitting them over two warps per quadrant does not pay, both land on the same scheduler)
    constexpr int a_first = 0, a_last = 7;
    const uint32_t tlane = tmem_base + g * 256 + ((uint32_t)(q * 32) << 16);
    TcHdr B, H1, H2;                                          // headers of blocks i, i + 2, i + 4
    uint4 r0n = make_uint4(0, 0, 0, 0), r1n = make_uint4(0, 0, 0, 0);
    if (tma_block<CL>(g) < nblk) {
      B = ld_hdr(hdr, tma_block<CL>(g));
      const int r = tma_row_of(q, lane, B.nrows, g);
      if (r >= 0) {
        const uint4* rp = reinterpret_cast<const uint4*>(rows + B.rbase + r);
        r0n = rp[0]; r1n = rp[1];
      }
    }
    if (tma_block<CL>(g + 2) < nblk) H1 = ld_hdr(hdr, tma_block<CL>(g + 2));
    for (int i = g;; i += 2) {
      const int b = tma_block<CL>(i);
      if (b >= nblk) break;
      const uint4 r0 = r0n, r1 = r1n;
      if (tma_block<CL>(i + 4) < nblk) H2 = ld_hdr(hdr, tma_block<CL>(i + 4));
      if (tma_block<CL>(i + 2) < nblk) {                         // next own block's row record in flight
        const int r = tma_row_of(q, lane, H1.nrows, g);
        if (r >= 0) {
          const uint4* rp = reinterpret_cast<const uint4*>(rows + H1.rbase + r);
          r0n = rp[0]; r1n = rp[1];
        }
      }
      const bool live = tma_row_of(q, lane, B.nrows, g) >= 0;
      __half* orow = out + (((long long)r0.w << 32) | (long long)r0.z);
      const float dx = __uint_as_float(r1.x), dy = __uint_as_float(r1.y);
      const int ox = (int)r1.z;
      const int oy = live ? (int)r1.w : (1 << 20);           // dead rows never match an accumulator row
      const bool p8 = (ox & 8) != 0, p4 = (ox & 4) != 0;
      float w[5];
      {
        const int sx = ox & 3;
#pragma unroll
        for (int j = 0; j < 5; j++) w[j] = (j == sx ? 1.0f - dx : 0.0f) + (j == sx + 1 ? dx : 0.0f);
      }
      // warp-uniform range of accumulator rows (tcgen05.ld is warp-collective)
      const int lo = __reduce_min_sync(0xffffffffu, live ? oy : 64);
      const int hi = __reduce_max_sync(0xffffffffu, live ? oy : -1);
      mbar_wait(&tfull[g], (i >> 1) & 1);
      asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
      if (q == 0 && lane == 0) TC_TRACE(5, i);
      if (hi >= 0 && !(dbg & 2)) {
        float hprev[7];
#pragma unroll
        for (int k = 0; k < 7; k++) hprev[k] = 0.f;
        // accumulator rows w_lo .. w_hi in trips of two; a block of duplicated rows (tma_row_of) is split between the
        // warps of quadrants q and q + 2: the second one starts from the blended row just above its first trip
        const int wbase = (B.flags >> 8) & ~1;              // accumulator column 0 = tile row wbase (see the MMA issuer)
        int w_lo = (lo + a_first) & ~1, w_hi = hi + a_last;
        if (B.nrows <= kTcDupRows) {
          const int trips = ((w_hi - w_lo) >> 1) + 1, first = (trips + 1) >> 1;
          if (q < 2) {
            w_hi = w_lo + 2 * first - 1;
          } else {
            w_lo += 2 * first;
            if (w_lo <= w_hi && !(dbg & 1)) {
              float v[16];
              tmem_ld16(tlane + (w_lo - 1 - wbase) * kTcTile, v);
              epi_hrow(v, p8, p4, w, hprev);
            }
          }
        }
#pragma unroll 1
        for (int wy = w_lo; wy <= w_hi; wy += 2) {                 // accumulator rows wy, wy + 1
          float v[32];
          tmem_ld32(tlane + (wy - wbase) * kTcTile, v);
          if (dbg & 1) continue;
          // one basic block for both rows (values are computed unconditionally, only the stores are
          // predicated), so the selects (ALU), blends (FMA) and conversions of independent rows interleave
          float hA[7], hB[7];
          epi_hrow(v, p8, p4, w, hA);
          epi_hrow(v + 16, p8, p4, w, hB);
          const uint4 uA = epi_pack(hprev, hA, dy), uB = epi_pack(hA, hB, dy);
          const int a = wy - oy;                             // window row of accumulator row wy
          if (a > a_first && a <= a_last) *reinterpret_cast<uint4*>(orow + (a - 1) * 8) = uA;
          if (a >= a_first && a < a_last) *reinterpret_cast<uint4*>(orow + a * 8) = uB;
#pragma unroll
          for (int k = 0; k < 7; k++) hprev[k] = hB[k];
        }
      }
      asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
      if (q == 0 && lane == 0) TC_TRACE(6, i);
      mbar_arrive(&tempty[g]);
      B = H1;
      H1 = H2;
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
  int32_t *cnt, *cursor, *tot, *total, *rowsub;
  int2* edge_pf;
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
  w.cnt = (int32_t*)take((size_t)nbins * kTcSub * 4);   // cnt, cursor and tot are adjacent: one memset
  w.cursor = (int32_t*)take((size_t)nbins * kTcSub * 4);
  w.tot = (int32_t*)take((size_t)nbins * 4);
  w.total = (int32_t*)take(16);
  w.rowsub = (int32_t*)take((size_t)R * 4);
  w.edge_pf = (int2*)take((size_t)(R / 9 + 1) * sizeof(int2));
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

extern "C" int rvo_corr_trace(long long* host_out) {
  RVO_CUDA(cudaDeviceSynchronize());
  RVO_CUDA(cudaMemcpyFromSymbol(host_out, g_tc_trace, sizeof(long long) * 8 * 256));
  return RVO_OK;
}

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
  RVO_CHECK_ARG((int64_t)fmap1->N * fmap1->sN < 0x7fffffff, "rvo_corr_tiles: fmap1 too large for 32-bit offsets");
  TcWs w = tc_layout(ws, R, G.nbins);
  RVO_CHECK_ARG((int64_t)w.total_bytes <= ws_bytes, "rvo_corr_tiles: workspace %lld < %lld bytes",
                (long long)ws_bytes, (long long)w.total_bytes);
  cudaStream_t st = (cudaStream_t)stream;
  RVO_CUDA(cudaMemsetAsync(w.cnt, 0, (size_t)((char*)w.total - (char*)w.cnt), st));   // cnt + cursor + tot: one memset
  RVO_CHECK_ARG(G.nbins <= kScanMaxBins, "rvo_corr_tiles: %d tiles exceed the scan capacity", G.nbins);
  const int64_t T = (int64_t)E * nlevels;
  tc_bin_count_kernel<<<(int)((T + 255) / 256), 256, 0, st>>>(G, coords, kk, jj, pmod, fmod, fmap1->N, E, w.cnt, w.tot,
                                                              w.rowsub, w.edge_pf);
  RVO_LAUNCH_CHECK("tc_bin_count_kernel");
  // scatter pass, every CTA with its own copy of the scanned tile table in shared memory
  const size_t scan_smem = (size_t)(G.nbins + 1) * 6 + (size_t)kScatterThreads * 4 + 64;
  RVO_CUDA(cudaFuncSetAttribute(tc_bin_scatter_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                kScanMaxBins * 6 + kScatterThreads * 4 + 64));
  int grid = sm_budget() * (scan_smem <= 100 * 1024 ? 2 : 1);         // two CTAs per SM while the table allows it
  if ((int64_t)grid * kScatterThreads > R) grid = (int)((R + kScatterThreads - 1) / kScatterThreads);
  tc_bin_scatter_kernel<<<grid, kScatterThreads, scan_smem, st>>>(G, coords, w.rowsub, w.edge_pf, (int)fmap1->sN,
                                                                (int)fmap1->sH, (int)fmap1->sW, out_ld, R, w.cnt, w.cursor,
                                                                w.tot, w.rows, w.hdr, w.total, (__half*)out);
  RVO_LAUNCH_CHECK("tc_bin_scatter_kernel");
#ifdef RVO_DEBUG
  const char* dbg_s = getenv("RVO_CORR_DBG");      // debugging aid: see the kernel's dbg bits
  const int dbg = dbg_s ? atoi(dbg_s) : 0;
#else
  const int dbg = 0;                               // the release library never reads the environment
#endif
  TcTmap tm[kTcMaxLevels];
  for (int l = 0; l < kTcMaxLevels; l++) {
    rc = tc_make_tmap(G.lv[l], &tm[l]);
    if (rc != RVO_OK) return rc;
  }
  // schedule granularity (tma_block): single blocks while the level-1 tiles hold about one 128-row block each
  const int64_t bins1 = (int64_t)G.lv[0].N * G.lv[0].TX * G.lv[0].TY;
  const bool short_chunks = (int64_t)E * 9 < 96 * bins1;
  auto kern = short_chunks ? corr_tile_tma_kernel<0> : corr_tile_tma_kernel<3>;
  RVO_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, kTmaSmemBytes));
  kern<<<sm_budget(), kTmaThreads, kTmaSmemBytes, st>>>(tm[0], tm[1], (const __half*)fmap1->data, w.hdr, w.rows, w.total,
                                                       (__half*)out, dbg);
  RVO_LAUNCH_CHECK("corr_tile_tma_kernel");
  return RVO_OK;
}
