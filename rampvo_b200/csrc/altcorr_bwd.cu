// altcorr_bwd.cu — the training-time gradients of ramp.altcorr:
//
//   rvo_corr_backward       cuda_corr.backward (ramp/altcorr/correlation.cpp:50-55; corr_backward_kernel
//                           correlation_kernel.cu:140-190 + the blend adjoint built with torch ops at :236-262).
//   rvo_patchify_backward   cuda_corr.patchify_backward (correlation_kernel.cu:50-80,306-331).
//
// Reference: one thread per (edge, pixel, window cell) doing 2*C scalar atomicAdds, after FOUR zero-initialised
// [E,D,D,P,P] temporaries and ~12 elementwise launches that un-blend the gradient.  Here one WARP owns an
// (edge, patch pixel) row: the 64 un-blended window gradients are formed in registers (two per lane, the 4-corner
// adjoint evaluated on the fly), lanes then split the channels — the gradient of the patch feature is reduced in
// registers over the whole window (1 atomic per channel instead of 64) and only the scatter into the frame map,
// whose targets genuinely collide across edges, stays atomic.  fp32 accumulation for every input dtype.
#include "common.cuh"

namespace rvo {

template <typename T>
__device__ __forceinline__ float ldf(const T* p) { return (float)*p; }

// grid: one warp per (edge e, patch pixel i0*P + j0)
template <typename T>
__global__ void __launch_bounds__(256)
corr_backward_kernel(const rvo_fmap_t f1, const rvo_fmap_t f2, const float* __restrict__ coords,
                     const int64_t* __restrict__ ii, const int64_t* __restrict__ jj,
                     const float* __restrict__ grad, int E, int P, int R, float* __restrict__ g1,
                     float* __restrict__ g2) {
  const int lane = threadIdx.x & 31;
  const int64_t row = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int PP = P * P;
  if (row >= (int64_t)E * PP) return;
  const int e = (int)(row / PP), pix = (int)(row % PP);
  const int D = 2 * R + 2, d = D - 1, C = f1.C, H2 = f2.H, W2 = f2.W;
  const int64_t ix = ii[e], jx = jj[e];
  const float x = coords[((size_t)e * 2 + 0) * PP + pix], y = coords[((size_t)e * 2 + 1) * PP + pix];
  const float fx = floorf(x), fy = floorf(y);
  const float dx = x - fx, dy = y - fy;
  const int x0 = (int)fx - R, y0 = (int)fy - R;
  // un-blended gradient of window cell (a, b) = sum over the <= 4 blended outputs that used it; the incoming
  // gradient is laid out [e, b' (x offset), a' (y offset), i0, j0] (correlation_kernel.cu:232,245)
  float G[2];
#pragma unroll
  for (int k = 0; k < 2; k++) {
    const int cell = lane + 32 * k;
    float g = 0.f;
    if (cell < D * D) {
      const int a = cell / D, b = cell % D;
      const float* ge = grad + (size_t)e * d * d * PP + pix;
      auto at = [&](int aa, int bb) -> float {
        return (aa >= 0 && aa < d && bb >= 0 && bb < d) ? ge[((size_t)bb * d + aa) * PP] : 0.f;
      };
      g = (1.f - dx) * (1.f - dy) * at(a, b) + dx * (1.f - dy) * at(a, b - 1) + (1.f - dx) * dy * at(a - 1, b) +
          dx * dy * at(a - 1, b - 1);
    }
    G[k] = g;
  }
  const T* p1 = (const T*)f1.data + ix * f1.sN + (pix / P) * f1.sH + (pix % P) * f1.sW;
  const T* p2 = (const T*)f2.data + jx * f2.sN;
  float* q1 = g1 + ((size_t)ix * C) * PP + pix;                         // [Np, C, P, P] dense
  float* q2 = g2 + ((size_t)jx * C) * H2 * W2;                          // [Nf, C, H2, W2] dense
  for (int c0 = 0; c0 < C; c0 += 32) {                                  // all lanes iterate: the shuffles are warp-wide
    const int c = c0 + lane;
    const bool live = c < C;
    const float v1 = live ? ldf(p1 + (int64_t)c * f1.sC) : 0.f;
    float acc = 0.f;
    for (int cell = 0; cell < D * D; cell++) {
      const float g = __shfl_sync(0xffffffffu, G[cell >> 5], cell & 31);
      const int i1 = y0 + cell / D, j1 = x0 + cell % D;
      if (live && g != 0.f && i1 >= 0 && i1 < H2 && j1 >= 0 && j1 < W2) {
        acc = fmaf(g, ldf(p2 + (int64_t)c * f2.sC + (int64_t)i1 * f2.sH + (int64_t)j1 * f2.sW), acc);
        atomicAdd(q2 + ((size_t)c * H2 + i1) * W2 + j1, g * v1);
      }
    }
    if (live) atomicAdd(q1 + (size_t)c * PP, acc);
  }
}

// thread per (m, c, a, b): net_grad[c, floor(y)+a-R, floor(x)+b-R] += patch_grad[m, c, a, b]
template <typename T>
__global__ void __launch_bounds__(256)
patchify_backward_kernel(const T* __restrict__ pg, const float* __restrict__ coords, int M, int C, int H, int W,
                         int R, float* __restrict__ out) {
  const int D = 2 * R + 2;
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= (int64_t)M * C * D * D) return;
  const int b = (int)(t % D), a = (int)((t / D) % D);
  const int c = (int)((t / (D * D)) % C), m = (int)(t / ((int64_t)D * D * C));
  const int i = (int)floorf(coords[2 * m + 1]) + a - R, j = (int)floorf(coords[2 * m]) + b - R;
  if (i >= 0 && i < H && j >= 0 && j < W) atomicAdd(out + ((size_t)c * H + i) * W + j, (float)pg[t]);
}

}  // namespace rvo

using namespace rvo;

extern "C" int rvo_corr_backward(const rvo_fmap_t* fmap1, const rvo_fmap_t* fmap2, const float* coords,
                                 const int64_t* ii, const int64_t* jj, const float* corr_grad, int E, int radius,
                                 float* fmap1_grad, float* fmap2_grad, void* stream) {
  RVO_CHECK_ARG(fmap1 && fmap2 && fmap1_grad && fmap2_grad, "rvo_corr_backward: null pointer");
  RVO_CHECK_ARG(fmap1->dtype == fmap2->dtype && (fmap1->dtype == RVO_F16 || fmap1->dtype == RVO_F32),
                "rvo_corr_backward: dtype");
  RVO_CHECK_ARG(fmap1->C == fmap2->C && fmap1->H == fmap1->W && fmap1->H >= 1, "rvo_corr_backward: shapes");
  RVO_CHECK_ARG(radius >= 0 && (2 * radius + 2) * (2 * radius + 2) <= 64, "rvo_corr_backward: radius %d (max 3)", radius);
  RVO_CHECK_ARG(E >= 0, "rvo_corr_backward: E=%d", E);
  cudaStream_t st = (cudaStream_t)stream;
  const int P = fmap1->H;
  RVO_CUDA(cudaMemsetAsync(fmap1_grad, 0, (size_t)fmap1->N * fmap1->C * P * P * sizeof(float), st));
  RVO_CUDA(cudaMemsetAsync(fmap2_grad, 0, (size_t)fmap2->N * fmap2->C * fmap2->H * fmap2->W * sizeof(float), st));
  if (E == 0) return RVO_OK;
  RVO_CHECK_ARG(coords && ii && jj && corr_grad, "rvo_corr_backward: null pointer");
  const int64_t warps = (int64_t)E * P * P;
  const int grid = cdiv(warps * 32, 256);
  if (fmap1->dtype == RVO_F16)
    corr_backward_kernel<__half><<<grid, 256, 0, st>>>(*fmap1, *fmap2, coords, ii, jj, corr_grad, E, P, radius,
                                                       fmap1_grad, fmap2_grad);
  else
    corr_backward_kernel<float><<<grid, 256, 0, st>>>(*fmap1, *fmap2, coords, ii, jj, corr_grad, E, P, radius,
                                                      fmap1_grad, fmap2_grad);
  RVO_LAUNCH_CHECK("corr_backward_kernel");
  return RVO_OK;
}

extern "C" int rvo_patchify_backward(const void* patch_grad, int dtype, const float* coords, int M, int C, int H,
                                     int W, int radius, float* net_grad, void* stream) {
  RVO_CHECK_ARG(net_grad && M >= 0 && C >= 1 && H >= 1 && W >= 1 && radius >= 0, "rvo_patchify_backward: arguments");
  cudaStream_t st = (cudaStream_t)stream;
  RVO_CUDA(cudaMemsetAsync(net_grad, 0, (size_t)C * H * W * sizeof(float), st));
  if (M == 0) return RVO_OK;
  RVO_CHECK_ARG(patch_grad && coords, "rvo_patchify_backward: null pointer");
  const int D = 2 * radius + 2;
  const int64_t n = (int64_t)M * C * D * D;
  if (dtype == RVO_F16)
    patchify_backward_kernel<__half><<<cdiv(n, 256), 256, 0, st>>>((const __half*)patch_grad, coords, M, C, H, W,
                                                                   radius, net_grad);
  else if (dtype == RVO_F32)
    patchify_backward_kernel<float><<<cdiv(n, 256), 256, 0, st>>>((const float*)patch_grad, coords, M, C, H, W,
                                                                  radius, net_grad);
  else
    RVO_CHECK_ARG(false, "rvo_patchify_backward: dtype %d", dtype);
  RVO_LAUNCH_CHECK("patchify_backward_kernel");
  return RVO_OK;
}
