// frame_ops.cu — the per-frame glue of the hot path as a handful of kernels (sm_100a):
//
//   rvo_select_patches   event-biased patch selection (ramp/utils.py:186-226 get_coords_from_topk_events with
//                        nms_image :157-183): |events| -> 4x4 average pool -> channel mean -> transpose ->
//                        NMS (k x k max filter) -> top-M -> (x = idx / H' as a TRUE division, y = idx % H').
//                        Reference: abs, avg_pool2d, transpose, mean, max_pool2d, eq, mul, flatten, topk (a
//                        108 us single-CTA gatherTopK + a key/value sort), div, remainder, stack = 12+ launches.
//                        Here: one grid-wide scoring pass + one CTA doing NMS and an exact top-M.
//   rvo_pyramid_level2   level-2 feature map = 4x4 average pool (ramp/Ramp_vo.py:381), channels-last.
//   rvo_copy_segments    up to 8 independent device-to-device copies in one launch (ring-buffer slot writes of a
//                        new frame: fmap1, fmap2, gmap, imap, patches — ramp/Ramp_vo.py:376-381).
//   rvo_event_stack      raw events (x, y, p in arrival order) -> the 5-bin int8 event stack
//                        (utils/transformers.py:128-161 EventToStack_Numpy), written as the fp32 tensor the encoder
//                        consumes (and optionally as int8).
#include "common.cuh"

namespace rvo {

// ------------------------------------------------------------------ patch selection ----

// score[x' * H4 + y'] = mean_c( avgpool4x4(|ev[c]|) )[y', x']   (transposed map, utils.py:196-198)
// avg_pool2d accumulates row-major over the 4x4 window in fp32 and divides by 16; mean = sum * (1/C) with the
// factor formed like ATen's MeanOps (float(n_out) / float(n_in)).  For the integer-valued event stacks of the
// reference every partial sum is exact, so the result is bit-identical to the torch ops.
__global__ void __launch_bounds__(256)
sel_score_kernel(const float* __restrict__ ev, int C, int H, int W, float factor, float* __restrict__ score) {
  const int H4 = H / 4, W4 = W / 4;
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= H4 * W4) return;
  const int y4 = t / W4, x4 = t % W4;
  float acc = 0.0f;
  for (int c = 0; c < C; c++) {
    const float* p = ev + ((size_t)c * H + 4 * y4) * W + 4 * x4;
    float s = 0.0f;
#pragma unroll
    for (int dy = 0; dy < 4; dy++) {
      const float4 v = *reinterpret_cast<const float4*>(p + (size_t)dy * W);
      s = __fadd_rn(s, fabsf(v.x));
      s = __fadd_rn(s, fabsf(v.y));
      s = __fadd_rn(s, fabsf(v.z));
      s = __fadd_rn(s, fabsf(v.w));
    }
    acc = __fadd_rn(acc, __fdiv_rn(s, 16.0f));
  }
  score[(size_t)x4 * H4 + y4] = __fmul_rn(acc, factor);
}

constexpr int kSelThreads = 1024;
constexpr int kSelMaxM = 1024;

// One CTA: border suppression, k x k NMS (separable max), exact top-M with torch.topk's CUDA ordering
// (values descending; equal values by ascending flat index — sbtopk gathers in index order and its key/value
// sort is stable), coordinates.  Shared memory: two n-float planes + the candidate keys.
__global__ void __launch_bounds__(kSelThreads, 1)
sel_topk_kernel(const float* __restrict__ score, int n_rows /*W4*/, int n_cols /*H4*/, int border, int nms,
                int M, int gather_order, float* __restrict__ coords /*[M,2]*/, long long* __restrict__ idx_out,
                float* __restrict__ val_out) {
  extern __shared__ float sm[];
  const int n = n_rows * n_cols;
  float* a = sm;                    // scores, then NMS-filtered scores
  float* b = sm + n;                // row-direction max
  __shared__ unsigned long long top[kSelMaxM];
  __shared__ unsigned int hist[256];
  __shared__ unsigned int s_cnt, s_digit, s_need;
  __shared__ unsigned long long s_prefix;
  const int tid = threadIdx.x;
  const int r = nms / 2;

  for (int t = tid; t < n; t += kSelThreads) {
    float v = score[t];
    if (border > 0) {
      const int row = t / n_cols, col = t % n_cols;
      if (row < border || row >= n_rows - border || col < border || col >= n_cols - border) v = 0.0f;
    }
    a[t] = v;
  }
  __syncthreads();
  if (nms > 0) {
    for (int t = tid; t < n; t += kSelThreads) {
      const int row = t / n_cols, col = t % n_cols;
      float m = a[t];
      for (int d = -r; d <= r; d++) {
        const int c = col + d;
        if (c >= 0 && c < n_cols) m = fmaxf(m, a[row * n_cols + c]);
      }
      b[t] = m;
    }
    __syncthreads();
    for (int t = tid; t < n; t += kSelThreads) {
      const int row = t / n_cols, col = t % n_cols;
      float m = b[t];
      for (int d = -r; d <= r; d++) {
        const int rr = row + d;
        if (rr >= 0 && rr < n_rows) m = fmaxf(m, b[rr * n_cols + col]);
      }
      const float v = a[t];
      a[t] = (m == v) ? v : 0.0f;        // x * (maxpool(x) == x), utils.py:181-183 (scores are >= 0)
    }
    __syncthreads();
  }
  // key = value bits (non-negative floats order like unsigned ints) : inverted index -> larger key = better
  auto key_of = [&](int t) -> unsigned long long {
    return ((unsigned long long)__float_as_uint(a[t]) << 32) | (unsigned long long)(0xffffffffu - (unsigned)t);
  };
  // radix select (MSB first, 8 bits per pass) of the M-th largest key; all keys are distinct
  if (tid == 0) { s_prefix = 0ull; s_need = (unsigned)M; }
  __syncthreads();
  for (int pass = 0; pass < 8; pass++) {
    const int shift = 56 - 8 * pass;
    for (int t = tid; t < 256; t += kSelThreads) hist[t] = 0;
    __syncthreads();
    const unsigned long long prefix = s_prefix;
    const unsigned long long mask = pass == 0 ? 0ull : (~0ull << (shift + 8));
    for (int t0 = 0; t0 < n; t0 += kSelThreads) {
      const int t = t0 + tid;
      bool ok = false;
      unsigned digit = 0;
      if (t < n) {
        const unsigned long long k = key_of(t);
        ok = (k & mask) == prefix;
        digit = (unsigned)(k >> shift) & 255u;
      }
      // warp-aggregated histogram update: one atomic per distinct digit per warp
      const unsigned active = __ballot_sync(0xffffffffu, ok);
      if (ok) {
        const unsigned peers = __match_any_sync(active, digit);
        if ((int)(__ffs(peers) - 1) == (tid & 31)) atomicAdd(&hist[digit], __popc(peers));
      }
    }
    __syncthreads();
    if (tid == 0) {
      unsigned need = s_need, acc = 0;
      int d = 255;
      for (; d > 0; d--) {
        if (acc + hist[d] >= need) break;
        acc += hist[d];
      }
      s_digit = (unsigned)d;
      s_need = need - acc;               // rank of the pivot inside bin d
      s_prefix = prefix | ((unsigned long long)d << shift);
    }
    __syncthreads();
  }
  const unsigned long long pivot = s_prefix;      // exactly M keys are >= pivot
  if (tid == 0) s_cnt = 0;
  for (int t = tid; t < kSelMaxM; t += kSelThreads) top[t] = 0ull;
  __syncthreads();
  for (int t = tid; t < n; t += kSelThreads) {
    const unsigned long long k = key_of(t);
    if (k >= pivot) {
      const unsigned slot = atomicAdd(&s_cnt, 1u);
      if (slot < (unsigned)kSelMaxM) top[slot] = k;
    }
  }
  __syncthreads();
  // bitonic sort (descending) of the kSelMaxM slots; empty slots are 0 and sink to the end
  for (int k2 = 2; k2 <= kSelMaxM; k2 <<= 1) {
    for (int j = k2 >> 1; j > 0; j >>= 1) {
      const int i = tid, l = i ^ j;
      if (l > i) {
        const unsigned long long x = top[i], y = top[l];
        const bool desc = (i & k2) == 0;
        if (desc ? (x < y) : (x > y)) { top[i] = y; top[l] = x; }
      }
      __syncthreads();
    }
  }
  if (gather_order) {
    // torch.topk's PRE-sort order (sbtopk::gatherTopK): values above the k-th value by ascending index, then the
    // ties with the k-th value by ascending index.  Used for M <= 32, where torch finishes with an unstable
    // bitonic sort that the caller reproduces by running the same torch sort on these values.
    const float pv = __uint_as_float((unsigned)(pivot >> 32));
    unsigned long long k = ~0ull;
    if (tid < M) {
      k = top[tid];
      const unsigned idx = 0xffffffffu - (unsigned)(k & 0xffffffffull);
      const float v = __uint_as_float((unsigned)(k >> 32));
      k = ((unsigned long long)(v == pv ? 1u : 0u) << 32) | (unsigned long long)idx;      // (class, index) order
    }
    __syncthreads();
    top[tid] = k;
    __syncthreads();
    for (int k2 = 2; k2 <= kSelMaxM; k2 <<= 1) {
      for (int j = k2 >> 1; j > 0; j >>= 1) {
        const int i = tid, l = i ^ j;
        if (l > i) {
          const unsigned long long x = top[i], y = top[l];
          const bool asc = (i & k2) == 0;
          if (asc ? (x > y) : (x < y)) { top[i] = y; top[l] = x; }
        }
        __syncthreads();
      }
    }
    for (int m = tid; m < M; m += kSelThreads) {
      const unsigned idx = (unsigned)(top[m] & 0xffffffffull);
      idx_out[m] = (long long)idx;
      val_out[m] = a[idx];
    }
    return;
  }
  // rows = indices / H' : torch's CUDA true division by a host scalar multiplies by the reciprocal
  // (div_true_kernel_cuda: a * (1 / b)), which is what the reference computes on the GPU (utils.py:212)
  const float inv = __fdiv_rn(1.0f, (float)n_cols);
  for (int m = tid; m < M; m += kSelThreads) {
    const unsigned idx = 0xffffffffu - (unsigned)(top[m] & 0xffffffffull);
    coords[2 * m] = __fmul_rn((float)idx, inv);
    coords[2 * m + 1] = (float)(idx % (unsigned)n_cols);
    if (idx_out) idx_out[m] = (long long)idx;
    if (val_out) val_out[m] = a[idx];
  }
}

// ------------------------------------------------------------------ pyramid level 2 ----

// out[y, x, c] = mean of the 4x4 block, fp32 accumulation in row-major window order then one rounding
// (F.avg_pool2d on an fp16 tensor accumulates in float, ramp/Ramp_vo.py:381)
template <typename T>
__global__ void __launch_bounds__(256)
pyramid2_kernel(const T* __restrict__ f, int H, int W, int C, T* __restrict__ out) {
  const int H4 = H / 4, W4 = W / 4;
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= (int64_t)H4 * W4 * C) return;
  const int c = (int)(t % C);
  const int x4 = (int)((t / C) % W4), y4 = (int)(t / ((int64_t)C * W4));
  float s = 0.0f;
#pragma unroll
  for (int dy = 0; dy < 4; dy++)
#pragma unroll
    for (int dx = 0; dx < 4; dx++)
      s = __fadd_rn(s, (float)f[((size_t)(4 * y4 + dy) * W + 4 * x4 + dx) * C + c]);
  out[t] = (T)__fdiv_rn(s, 16.0f);
}

// ------------------------------------------------------------------ multi-segment copy ----

struct CopySegs {
  const void* src[8];
  void* dst[8];
  int64_t bytes[8];
  int64_t start[9];      // prefix sums in 16-byte (or 1-byte) units
  int n;
};

template <typename V>
__global__ void __launch_bounds__(256)
copy_segments_kernel(const CopySegs s) {
  const int64_t total = s.start[s.n];
  for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (int64_t)gridDim.x * blockDim.x) {
    int k = 0;
    while (k + 1 < s.n && t >= s.start[k + 1]) k++;
    const int64_t o = t - s.start[k];
    reinterpret_cast<V*>(s.dst[k])[o] = reinterpret_cast<const V*>(s.src[k])[o];
  }
}

// ------------------------------------------------------------------ event stack ----

// utils/transformers.py:149-161: bin b = int32( (num_bins * float32(i)) / N ) by ARRIVAL ORDER i, then
// np.add.at(voxel, (b, y, x), p) for in-bounds integer pixels.  Sums of +-1 in fp32 are exact in any order.
__global__ void __launch_bounds__(256)
event_scatter_kernel(const uint16_t* __restrict__ x, const uint16_t* __restrict__ y, const float* __restrict__ p,
                     int64_t n_events, int bins, int H, int W, float* __restrict__ acc) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_events; i += (int64_t)gridDim.x * blockDim.x) {
    const int b = (int)__fdiv_rn(__fmul_rn((float)bins, (float)i), (float)n_events);
    const int xi = x[i], yi = y[i];
    if (xi < W && yi < H && b >= 0 && b < bins) atomicAdd(&acc[((size_t)b * H + yi) * W + xi], p[i]);
  }
}

// voxel_grid.astype("int8") (transformers.py:159): C conversion float -> int32 (truncation) -> int8 (wrap)
__global__ void __launch_bounds__(256)
event_finalize_kernel(float* __restrict__ acc, int64_t n, int8_t* __restrict__ out8) {
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n) return;
  const int8_t v = (int8_t)(int)acc[t];
  acc[t] = (float)v;
  if (out8) out8[t] = v;
}

// ------------------------------------------------------------------ frame commit ----

struct FrameCommit {
  const float* patches_new;   // [M,3,P,P] staged patches of the new frame (x, y, disparity = 1)
  const float* clr;           // [M,3] colours sampled by patchify (RGB in [-0.5,1.5])
  const float* depth_rand;    // [M] uniform draws (before initialisation) or null
  float* patches;             // [N,M,3,P,P]
  int64_t* tstamps;           // [N]
  float* intrinsics;          // [N,4]
  int64_t* index;             // [N,M]
  int64_t* index_map;         // [N]
  uint8_t* colors;            // [N,M,3]
  float intr[4];              // already divided by RES
  int n, M, P, N;
  int64_t counter, m_next;
  int median_frames;          // > 0: depth = median of the depths of the last `median_frames` frames
};

constexpr int kCommitThreads = 1024;
constexpr int kCommitMaxMedian = 2048;      // patch depths entering the median (3 frames x 300 patches = 900)

// ramp/Ramp_vo.py:345-372 in one launch: timestamps / intrinsics / index rows, colour conversion, depth
// initialisation (uniform draw, or the lower median of the last frames' depths like torch.median) and the
// write of the new patches into slot n.
__global__ void __launch_bounds__(kCommitThreads, 1)
frame_commit_kernel(const FrameCommit a) {
  __shared__ float srt[kCommitMaxMedian];
  __shared__ float s_med;
  const int tid = threadIdx.x;
  const int PP = a.P * a.P;
  if (tid == 0) {
    a.tstamps[a.n] = a.counter;
    if (a.n + 1 < a.N) a.index_map[a.n + 1] = a.m_next;
  }
  if (tid < 4) a.intrinsics[(size_t)a.n * 4 + tid] = a.intr[tid];
  if (a.n + 1 < a.N)
    for (int t = tid; t < a.M; t += kCommitThreads) a.index[(size_t)(a.n + 1) * a.M + t] = a.n + 1;
  // colours: (clr[:, [2,1,0]] + 0.5) * (255/2) -> uint8 (Ramp_vo.py:352-353; float -> uint8 truncates)
  for (int t = tid; t < a.M * 3; t += kCommitThreads) {
    const int m = t / 3, c = t % 3;
    const float v = __fmul_rn(__fadd_rn(a.clr[m * 3 + (2 - c)], 0.5f), 255.0f / 2);
    a.colors[((size_t)a.n * a.M + m) * 3 + c] = (uint8_t)(int)v;
  }
  float med = 0.0f;
  if (a.median_frames > 0) {
    // torch.median(patches_[n-3:n, :, 2]) (Ramp_vo.py:370-371) = the lower median of F*M*P*P values.  Every pixel
    // of a patch carries the same inverse depth (BA and the initialisation write all P*P of them, ba_cuda.cu:
    // 225-227), so the sorted list is the sorted list of the F*M patch depths with each entry repeated P*P times
    // and its element (cnt-1)/2 is element ((cnt-1)/2) / (P*P) of the short list: sort F*M values, not F*M*P*P.
    const int cntp = a.median_frames * a.M;
    int cap = 1;
    while (cap < cntp) cap <<= 1;
    for (int t = tid; t < cap; t += kCommitThreads) {
      float v = __int_as_float(0x7f800000);
      if (t < cntp)
        v = a.patches[((size_t)(a.n - a.median_frames) * a.M + t) * 3 * PP + 2 * PP + (a.P >= 2 ? a.P + 1 : 0)];
      srt[t] = v;
    }
    __syncthreads();
    for (int k2 = 2; k2 <= cap; k2 <<= 1)
      for (int j = k2 >> 1; j > 0; j >>= 1) {
        for (int i = tid; i < cap; i += kCommitThreads) {
          const int l = i ^ j;
          if (l > i) {
            const float x = srt[i], y = srt[l];
            const bool asc = (i & k2) == 0;
            if (asc ? (x > y) : (x < y)) { srt[i] = y; srt[l] = x; }
          }
        }
        __syncthreads();
      }
    if (tid == 0) s_med = srt[((cntp * PP - 1) / 2) / PP];
    __syncthreads();
    med = s_med;
  }
  for (int t = tid; t < a.M * 3 * PP; t += kCommitThreads) {
    const int m = t / (3 * PP), ch = (t / PP) % 3;
    float v = a.patches_new[t];
    if (ch == 2) v = a.median_frames > 0 ? med : (a.depth_rand ? a.depth_rand[m] : v);
    a.patches[(size_t)a.n * a.M * 3 * PP + t] = v;
  }
}

}  // namespace rvo

using namespace rvo;

extern "C" int rvo_frame_commit(const float* patches_new, const float* clr, const float* depth_rand,
                                float* patches, int64_t* tstamps, float* intrinsics, int64_t* index,
                                int64_t* index_map, uint8_t* colors, const float* intr4, int n, int M, int P,
                                int N, int64_t counter, int64_t m_next, int median_frames, void* stream) {
  RVO_CHECK_ARG(patches_new && clr && patches && tstamps && intrinsics && index && index_map && colors && intr4,
                "rvo_frame_commit: null pointer");
  RVO_CHECK_ARG(n >= 0 && n < N && M >= 1 && P >= 1, "rvo_frame_commit: n=%d N=%d M=%d P=%d", n, N, M, P);
  RVO_CHECK_ARG(median_frames >= 0 && median_frames <= n, "rvo_frame_commit: median over %d frames at n=%d",
                median_frames, n);
  RVO_CHECK_ARG((int64_t)median_frames * M <= kCommitMaxMedian,
                "rvo_frame_commit: median over %d patches (max %d)", median_frames * M, kCommitMaxMedian);
  FrameCommit a;
  a.patches_new = patches_new; a.clr = clr; a.depth_rand = depth_rand; a.patches = patches;
  a.tstamps = tstamps; a.intrinsics = intrinsics; a.index = index; a.index_map = index_map; a.colors = colors;
  for (int k = 0; k < 4; k++) a.intr[k] = intr4[k];
  a.n = n; a.M = M; a.P = P; a.N = N; a.counter = counter; a.m_next = m_next; a.median_frames = median_frames;
  frame_commit_kernel<<<1, kCommitThreads, 0, (cudaStream_t)stream>>>(a);
  RVO_LAUNCH_CHECK("frame_commit_kernel");
  return RVO_OK;
}

extern "C" int64_t rvo_select_ws_bytes(int H, int W) {
  if (H < 4 || W < 4) return -1;
  return (int64_t)(H / 4) * (W / 4) * sizeof(float);
}

extern "C" int rvo_select_patches(const float* events, int C, int H, int W, int M, int border, int nms,
                                  int gather_order, float* coords, int64_t* idx_out, float* val_out, void* ws,
                                  int64_t ws_bytes, void* stream) {
  RVO_CHECK_ARG(events && ws, "rvo_select_patches: null pointer");
  RVO_CHECK_ARG(gather_order ? (idx_out && val_out) : (coords != nullptr),
                "rvo_select_patches: gather order needs idx_out + val_out, sorted order needs coords");
  RVO_CHECK_ARG(C >= 1 && H >= 4 && W >= 4 && H % 4 == 0 && W % 4 == 0, "rvo_select_patches: events [%d,%d,%d]", C, H, W);
  RVO_CHECK_ARG((reinterpret_cast<uintptr_t>(events) & 15u) == 0, "rvo_select_patches: events must be 16-byte aligned");
  const int H4 = H / 4, W4 = W / 4, n = H4 * W4;
  RVO_CHECK_ARG(M >= 1 && M <= kSelMaxM && M <= n, "rvo_select_patches: M=%d (1..%d)", M, kSelMaxM);
  RVO_CHECK_ARG(nms >= 0 && (nms == 0 || nms % 2 == 1), "rvo_select_patches: NMS window %d must be odd", nms);
  RVO_CHECK_ARG(border >= 0, "rvo_select_patches: border %d", border);
  RVO_CHECK_ARG(ws_bytes >= (int64_t)n * (int64_t)sizeof(float), "rvo_select_patches: workspace too small");
  const size_t smem = 2 * (size_t)n * sizeof(float);
  RVO_CHECK_ARG(smem <= 200 * 1024, "rvo_select_patches: %dx%d score map does not fit shared memory", H4, W4);
  cudaStream_t st = (cudaStream_t)stream;
  float* score = (float*)ws;
  const float factor = (float)n / (float)((int64_t)n * C);
  sel_score_kernel<<<cdiv(n, 256), 256, 0, st>>>(events, C, H, W, factor, score);
  RVO_LAUNCH_CHECK("sel_score_kernel");
  RVO_CUDA(cudaFuncSetAttribute(sel_topk_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  sel_topk_kernel<<<1, kSelThreads, smem, st>>>(score, W4, H4, border, nms, M, gather_order, coords,
                                                (long long*)idx_out, val_out);
  RVO_LAUNCH_CHECK("sel_topk_kernel");
  return RVO_OK;
}

extern "C" int rvo_pyramid_level2(const void* fmap, int dtype, int H, int W, int C, void* out, void* stream) {
  RVO_CHECK_ARG(fmap && out, "rvo_pyramid_level2: null pointer");
  RVO_CHECK_ARG(H >= 4 && W >= 4 && C >= 1, "rvo_pyramid_level2: fmap [%d,%d,%d]", H, W, C);
  const int64_t n = (int64_t)(H / 4) * (W / 4) * C;
  cudaStream_t st = (cudaStream_t)stream;
  if (dtype == RVO_F16)
    pyramid2_kernel<__half><<<cdiv(n, 256), 256, 0, st>>>((const __half*)fmap, H, W, C, (__half*)out);
  else if (dtype == RVO_F32)
    pyramid2_kernel<float><<<cdiv(n, 256), 256, 0, st>>>((const float*)fmap, H, W, C, (float*)out);
  else
    RVO_CHECK_ARG(false, "rvo_pyramid_level2: dtype %d", dtype);
  RVO_LAUNCH_CHECK("pyramid2_kernel");
  return RVO_OK;
}

extern "C" int rvo_copy_segments(const void* const* src, void* const* dst, const int64_t* bytes, int n,
                                 void* stream) {
  RVO_CHECK_ARG(n >= 0 && n <= 8, "rvo_copy_segments: %d segments (max 8)", n);
  if (n == 0) return RVO_OK;
  RVO_CHECK_ARG(src && dst && bytes, "rvo_copy_segments: null pointer");
  CopySegs s;
  bool vec = true;
  for (int k = 0; k < n; k++) {
    RVO_CHECK_ARG(bytes[k] >= 0 && (bytes[k] == 0 || (src[k] && dst[k])), "rvo_copy_segments: segment %d", k);
    vec = vec && bytes[k] % 16 == 0 && (reinterpret_cast<uintptr_t>(src[k]) & 15u) == 0 &&
          (reinterpret_cast<uintptr_t>(dst[k]) & 15u) == 0;
  }
  s.n = n;
  s.start[0] = 0;
  for (int k = 0; k < n; k++) {
    s.src[k] = src[k];
    s.dst[k] = dst[k];
    s.bytes[k] = bytes[k];
    s.start[k + 1] = s.start[k] + (vec ? bytes[k] / 16 : bytes[k]);
  }
  const int64_t total = s.start[n];
  if (total == 0) return RVO_OK;
  int grid = cdiv(total, 256);
  if (grid > sm_budget() * 16) grid = sm_budget() * 16;
  if (vec)
    copy_segments_kernel<uint4><<<grid, 256, 0, (cudaStream_t)stream>>>(s);
  else
    copy_segments_kernel<uint8_t><<<grid, 256, 0, (cudaStream_t)stream>>>(s);
  RVO_LAUNCH_CHECK("copy_segments_kernel");
  return RVO_OK;
}

extern "C" int rvo_event_stack(const uint16_t* x, const uint16_t* y, const float* p, int64_t n_events, int bins,
                               int H, int W, float* stack_f32, int8_t* stack_i8, void* stream) {
  RVO_CHECK_ARG(stack_f32, "rvo_event_stack: null output");
  RVO_CHECK_ARG(bins >= 1 && H >= 1 && W >= 1 && n_events >= 0, "rvo_event_stack: bad sizes");
  RVO_CHECK_ARG(n_events < (1 << 24), "rvo_event_stack: more than 2^24 events per stack (fp32 index arithmetic "
                                      "of the reference stops being exact)");
  cudaStream_t st = (cudaStream_t)stream;
  const int64_t n = (int64_t)bins * H * W;
  RVO_CUDA(cudaMemsetAsync(stack_f32, 0, (size_t)n * sizeof(float), st));
  if (n_events < 2) {                        // transformers.py:146-147: fewer than two events -> empty stack
    if (stack_i8) RVO_CUDA(cudaMemsetAsync(stack_i8, 0, (size_t)n, st));
    return RVO_OK;
  }
  RVO_CHECK_ARG(x && y && p, "rvo_event_stack: null event arrays");
  int grid = cdiv(n_events, 256);
  if (grid > sm_budget() * 8) grid = sm_budget() * 8;
  event_scatter_kernel<<<grid, 256, 0, st>>>(x, y, p, n_events, bins, H, W, stack_f32);
  RVO_LAUNCH_CHECK("event_scatter_kernel");
  event_finalize_kernel<<<cdiv(n, 256), 256, 0, st>>>(stack_f32, n, stack_i8);
  RVO_LAUNCH_CHECK("event_finalize_kernel");
  return RVO_OK;
}

// ------------------------------------------------------------------ patch-graph step of a new frame ----
//
// What Ramp_vo does to the edge list between two recurrent updates when no keyframe is dropped
// (ramp/Ramp_vo.py:203-208 remove_factors of the edges whose source frame left the removal window, :194-201 +
// :312-325 append_factors of the new frame's forward and backward edges) costs the reference — and cost this
// repo's host code — a dozen small tensor ops (boolean index, three cats, a row gather of the hidden state) on the
// critical path between the keyframe decision and the next update.  Here it is two launches:
//   edges_step_kernel   ONE CTA: order-preserving compaction of (ii, jj, kk) by `ii >= lim` (chunked block scan),
//                       then the appended edges generated from (n, M, r); src_row[e'] = the old row of every
//                       surviving edge, -1 for a new edge;
//   net_rows_kernel     the hidden state rows follow: net_out[e'] = net_in[src_row[e']] or 0.
namespace rvo {

constexpr int kEsThreads = 1024;

// One CTA per tile of kEsThreads edges.  The ranks of the surviving edges need the number of survivors in all earlier
// tiles: every CTA publishes its count as (epoch << 32 | count) in tile_state[] right after counting and sums the
// words of its predecessors (spinning until they carry this launch's epoch — predecessors have lower block indices,
// so they are scheduled no later than their waiters).  The last CTA, which knows the total, appends the new edges.
// (A single CTA doing all of it is bound by what one SM can move: 42 us for the 2.9 MB of a 47 712-edge list.)
__global__ void __launch_bounds__(kEsThreads)
edges_step_kernel(const int64_t* __restrict__ ii, const int64_t* __restrict__ jj, const int64_t* __restrict__ kk,
                  int E0, int drop_k, int lim, int n, int M, int r, int64_t* __restrict__ ii_o,
                  int64_t* __restrict__ jj_o, int64_t* __restrict__ kk_o, int32_t* __restrict__ src_row, int E_expected,
                  float* __restrict__ status, unsigned long long* __restrict__ tile_state, unsigned int epoch) {
  __shared__ int wsum[kEsThreads / 32];
  __shared__ int base_s;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int T = (int)gridDim.x, j = (int)blockIdx.x;
  const int idx = j * kEsThreads + tid;
  int64_t vi = 0, vj = 0, vk = 0;
  bool keep = false;
  if (idx < E0) {
    vi = ii[idx]; vj = jj[idx]; vk = kk[idx];
    keep = true;
    if (drop_k >= 0) {
      // keyframe drop_k goes first (Ramp_vo.py:249-262): its edges are removed, later frames / patches move down
      keep = vi != drop_k && vj != drop_k;
      if (vi > drop_k) { vi -= 1; vk -= M; }
      if (vj > drop_k) vj -= 1;
    }
    keep = keep && vi >= lim;
  }
  const unsigned bal = __ballot_sync(0xffffffffu, keep);
  if (lane == 0) wsum[warp] = __popc(bal);
  __syncthreads();
  if (warp == 0) {                                     // exclusive scan of the 32 warp counts; publish the tile count
    const int c = wsum[lane];
    int incl = c;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int v = __shfl_up_sync(0xffffffffu, incl, o);
      if (lane >= o) incl += v;
    }
    wsum[lane] = incl - c;
    const int total = __shfl_sync(0xffffffffu, incl, 31);
    if (lane == 0) {
      __threadfence();
      atomicExch(&tile_state[j], ((unsigned long long)epoch << 32) | (unsigned int)total);
    }
    // survivors in the earlier tiles
    int before = 0;
    for (int i0 = 0; i0 < j; i0 += 32) {
      const int i = i0 + lane;
      int c2 = 0;
      if (i < j) {
        unsigned long long w;
        do {
          w = *reinterpret_cast<volatile unsigned long long*>(&tile_state[i]);
        } while ((unsigned int)(w >> 32) != epoch);
        c2 = (int)(unsigned int)(w & 0xffffffffull);
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) c2 += __shfl_xor_sync(0xffffffffu, c2, o);
      before += c2;
    }
    if (lane == 0) base_s = before;
  }
  __syncthreads();
  if (keep) {
    const int pos = base_s + wsum[warp] + __popc(bal & ((1u << lane) - 1u));
    ii_o[pos] = vi; jj_o[pos] = vj; kk_o[pos] = vk;
    src_row[pos] = idx;
  }
  if (j != T - 1) return;
  // ---- last CTA: the edges of the new frame follow the survivors
  __shared__ int kept_s;
  if (tid == 0) {
    const unsigned long long w = *reinterpret_cast<volatile unsigned long long*>(&tile_state[j]);
    kept_s = base_s + (int)(unsigned int)(w & 0xffffffffull);
  }
  __syncthreads();
  const int kept = E0 > 0 ? kept_s : 0;
  // forward edges: every patch of frames [f0, f1) -> frame n-1 (Ramp_vo.py:312-318)
  const int f0 = max(n - r, 0), f1 = max(n - 1, 0);
  const int n_f = M * (f1 - f0);
  for (int t = tid; t < n_f; t += kEsThreads) {
    const int k = M * f0 + t;
    ii_o[kept + t] = k / M; jj_o[kept + t] = n - 1; kk_o[kept + t] = k;
    src_row[kept + t] = -1;
  }
  // backward edges: every patch of frame n-1 -> frames [j0, n), patch-major (Ramp_vo.py:320-325, 'ij' meshgrid)
  const int j0 = max(n - r, 0), nj = n - j0;
  const int n_b = n >= 1 ? M * nj : 0;
  for (int t = tid; t < n_b; t += kEsThreads) {
    const int p = t / nj, jt = j0 + (t - p * nj);
    const int o = kept + n_f + t;
    ii_o[o] = n - 1; jj_o[o] = jt; kk_o[o] = (int64_t)M * (n - 1) + p;
    src_row[o] = -1;
  }
  if (tid == 0) status[0] = kept + n_f + n_b == E_expected ? 0.0f : (float)(kept + n_f + n_b + 1);
}

// warp per row of C = 384 floats
__global__ void __launch_bounds__(256)
net_rows_kernel(const float* __restrict__ net_in, const int32_t* __restrict__ src_row, int E, int C4,
                float* __restrict__ net_out) {
  const int lane = threadIdx.x & 31;
  for (int e = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; e < E; e += (gridDim.x * blockDim.x) >> 5) {
    const int s = src_row[e];
    const float4* src = reinterpret_cast<const float4*>(net_in) + (size_t)(s < 0 ? 0 : s) * C4;
    float4* dst = reinterpret_cast<float4*>(net_out) + (size_t)e * C4;
    for (int c = lane; c < C4; c += 32) dst[c] = s < 0 ? make_float4(0.f, 0.f, 0.f, 0.f) : src[c];
  }
}

}  // namespace rvo

extern "C" int rvo_net_rows(const float* net_in, const int32_t* src_row, int E, int C, float* net_out, void* stream) {
  RVO_CHECK_ARG(E >= 0 && C % 4 == 0, "rvo_net_rows: bad sizes");
  if (E == 0) return RVO_OK;
  RVO_CHECK_ARG(net_in && src_row && net_out && net_out != net_in, "rvo_net_rows: null / aliased pointer");
  int grid = rvo::cdiv((int64_t)E * 32, 256);
  if (grid > rvo::sm_budget() * 8) grid = rvo::sm_budget() * 8;
  rvo::net_rows_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(net_in, src_row, E, C / 4, net_out);
  RVO_LAUNCH_CHECK("net_rows_kernel");
  return RVO_OK;
}

extern "C" int64_t rvo_edges_step_tiles(int E0) { return E0 <= 0 ? 1 : (E0 + rvo::kEsThreads - 1) / rvo::kEsThreads; }

extern "C" int rvo_edges_step(const int64_t* ii, const int64_t* jj, const int64_t* kk, int E0, int drop_k, int lim,
                              int n, int M, int r, int64_t* ii_out, int64_t* jj_out, int64_t* kk_out, int E_new,
                              int32_t* src_row, float* status, uint64_t* tile_state, uint32_t epoch,
                              const float* net_in, int C, float* net_out, void* stream) {
  RVO_CHECK_ARG(E0 >= 0 && E_new >= 0 && n >= 1 && M >= 1 && r >= 1, "rvo_edges_step: bad sizes");
  RVO_CHECK_ARG(ii_out && jj_out && kk_out && src_row && status && tile_state && (E0 == 0 || (ii && jj && kk)),
                "rvo_edges_step: null pointer");
  RVO_CHECK_ARG(!net_out || ((net_in || E0 == 0) && C % 4 == 0 && net_out != net_in), "rvo_edges_step: hidden-state buffers");
  cudaStream_t st = (cudaStream_t)stream;
  const int tiles = (int)rvo_edges_step_tiles(E0);
  rvo::edges_step_kernel<<<tiles, rvo::kEsThreads, 0, st>>>(
      ii, jj, kk, E0, drop_k, lim, n, M, r, ii_out, jj_out, kk_out, src_row, E_new, status,
      reinterpret_cast<unsigned long long*>(tile_state), epoch);
  RVO_LAUNCH_CHECK("edges_step_kernel");
  if (net_out && E_new > 0) {
    int grid = rvo::cdiv((int64_t)E_new * 32, 256);
    if (grid > rvo::sm_budget() * 8) grid = rvo::sm_budget() * 8;
    rvo::net_rows_kernel<<<grid, 256, 0, st>>>(net_in, src_row, E_new, C / 4, net_out);
    RVO_LAUNCH_CHECK("net_rows_kernel");
  }
  return RVO_OK;
}
