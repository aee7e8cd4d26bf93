// scene_lstm.cu — the recurrent front end of the SingleScale RAMP encoder (ramp/extractor.py:187-269,
// MergerLSTMsceneEncoder) for one event stack + one image per call, as two launches:
//
//   scene_presence_kernel   events_are_present / image_is_present = any(x != 0) (extractor.py:253-254) into two
//                           DEVICE flags — the reference reads them on the host (two torch.any syncs per frame)
//   scene_lstm_kernel       per pixel: one nn.LSTM step for the events and one for the image with the (h, c)
//                           state CARRIED across calls (extractor.py:242-243), then the shared 1x1 super-state
//                           convolution applied for the events and for the image when present (:255-258).
//
// The reference runs two cuDNN RNNs with batch = H*W = 307 200 sequences of length 1 plus permute/contiguous
// copies both ways (:238-247).  Here a thread owns a pixel: inputs and states are planar [C, H*W] (coalesced),
// the 2.9 k weights live in shared memory, everything is fp32; the super state leaves as the fp16 channels-last
// [H, W, 16] tensor (15 channels + one zero pad so that a pixel is 32 bytes) the CNNs consume, and as the planar
// fp32 state of the next call.
#include "common.cuh"

namespace rvo {

constexpr int kSlHid = 15;          // output_lstm_dim (net.py:105)
constexpr int kSlMaxIn = 8;

__global__ void __launch_bounds__(256)
scene_presence_kernel(const float* __restrict__ ev, int64_t n_ev, const float* __restrict__ im, int64_t n_im,
                      int32_t* __restrict__ flags) {
  bool e = false, i = false;
  for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < n_ev; t += (int64_t)gridDim.x * blockDim.x)
    e |= ev[t] != 0.0f;
  for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < n_im; t += (int64_t)gridDim.x * blockDim.x)
    i |= im[t] != 0.0f;
  if (__any_sync(0xffffffffu, e) && (threadIdx.x & 31) == 0) atomicOr(&flags[0], 1);
  if (__any_sync(0xffffffffu, i) && (threadIdx.x & 31) == 0) atomicOr(&flags[1], 1);
}

struct SceneLstmParams {
  // packed fp32 parameter block (host layout, see rvo_scene_lstm_params_floats):
  //   ev: W_ih [4h, Ce] | W_hh [4h, h] | b_ih + b_hh [4h];  im: the same with Ci;  superstate W [h, 2h] | b [h]
  const float* w;
  int Ce, Ci;
};

__device__ __forceinline__ float sigmoidf_(float x) { return 1.0f / (1.0f + __expf(-x)); }

// one LSTM step for one pixel (torch gate order i, f, g, o; nn.LSTM: c' = f*c + i*g, h' = o*tanh(c'))
template <int CIN>
__device__ __forceinline__ void lstm_step(const float* __restrict__ Wih, const float* __restrict__ Whh,
                                          const float* __restrict__ b, const float* x, const float* hp,
                                          const float* cp, float* hn, float* cn) {
  constexpr int H = kSlHid;
#pragma unroll
  for (int u = 0; u < H; u++) {
    float g4[4];
#pragma unroll
    for (int g = 0; g < 4; g++) {
      const int r = g * H + u;
      float s = b[r];
#pragma unroll
      for (int c = 0; c < CIN; c++) s = fmaf(Wih[r * CIN + c], x[c], s);
#pragma unroll
      for (int c = 0; c < H; c++) s = fmaf(Whh[r * H + c], hp[c], s);
      g4[g] = s;
    }
    const float c1 = sigmoidf_(g4[1]) * cp[u] + sigmoidf_(g4[0]) * tanhf(g4[2]);
    cn[u] = c1;
    hn[u] = sigmoidf_(g4[3]) * tanhf(c1);
  }
}

template <int CE, int CI>
__global__ void __launch_bounds__(128)
scene_lstm_kernel(const float* __restrict__ wpack, const float* __restrict__ ev, const float* __restrict__ im,
                  int64_t HW, float* __restrict__ st_ev /*[2,h,HW] h then c*/, float* __restrict__ st_im,
                  float* __restrict__ super /*[h,HW]*/, const int32_t* __restrict__ flags, int first,
                  __half* __restrict__ out /*[HW,16]*/) {
  constexpr int H = kSlHid, G = 4 * kSlHid;
  constexpr int n_ev = G * CE + G * H + G, n_im = G * CI + G * H + G, n_ss = H * 2 * H + H;
  __shared__ float w[n_ev + n_im + n_ss];
  for (int t = threadIdx.x; t < n_ev + n_im + n_ss; t += blockDim.x) w[t] = wpack[t];
  __syncthreads();
  const float* We_ih = w;
  const float* We_hh = We_ih + G * CE;
  const float* be = We_hh + G * H;
  const float* Wi_ih = w + n_ev;
  const float* Wi_hh = Wi_ih + G * CI;
  const float* bi = Wi_hh + G * H;
  const float* Ws = w + n_ev + n_im;
  const float* bs = Ws + H * 2 * H;
  const bool ev_present = flags[0] != 0, im_present = flags[1] != 0;

  for (int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; p < HW; p += (int64_t)gridDim.x * blockDim.x) {
    float x[kSlMaxIn], hp[H], cp[H], he[H], hi[H], cn[H];
    // events
#pragma unroll
    for (int c = 0; c < CE; c++) x[c] = ev[(size_t)c * HW + p];
#pragma unroll
    for (int c = 0; c < H; c++) {
      hp[c] = first ? 0.0f : st_ev[(size_t)c * HW + p];
      cp[c] = first ? 0.0f : st_ev[(size_t)(H + c) * HW + p];
    }
    lstm_step<CE>(We_ih, We_hh, be, x, hp, cp, he, cn);
#pragma unroll
    for (int c = 0; c < H; c++) {
      st_ev[(size_t)c * HW + p] = he[c];
      st_ev[(size_t)(H + c) * HW + p] = cn[c];
    }
    // image
#pragma unroll
    for (int c = 0; c < CI; c++) x[c] = im[(size_t)c * HW + p];
#pragma unroll
    for (int c = 0; c < H; c++) {
      hp[c] = first ? 0.0f : st_im[(size_t)c * HW + p];
      cp[c] = first ? 0.0f : st_im[(size_t)(H + c) * HW + p];
    }
    lstm_step<CI>(Wi_ih, Wi_hh, bi, x, hp, cp, hi, cn);
#pragma unroll
    for (int c = 0; c < H; c++) {
      st_im[(size_t)c * HW + p] = hi[c];
      st_im[(size_t)(H + c) * HW + p] = cn[c];
    }
    // super state: ss = W [ss_prev ; data] + b, first with the event embedding, then with the image embedding
    float ss[H], tmp[H];
#pragma unroll
    for (int c = 0; c < H; c++) ss[c] = first ? 0.0f : super[(size_t)c * HW + p];
    if (ev_present) {
#pragma unroll
      for (int o = 0; o < H; o++) {
        float s = bs[o];
#pragma unroll
        for (int c = 0; c < H; c++) s = fmaf(Ws[o * 2 * H + c], ss[c], s);
#pragma unroll
        for (int c = 0; c < H; c++) s = fmaf(Ws[o * 2 * H + H + c], he[c], s);
        tmp[o] = s;
      }
#pragma unroll
      for (int c = 0; c < H; c++) ss[c] = tmp[c];
    }
    if (im_present) {
#pragma unroll
      for (int o = 0; o < H; o++) {
        float s = bs[o];
#pragma unroll
        for (int c = 0; c < H; c++) s = fmaf(Ws[o * 2 * H + c], ss[c], s);
#pragma unroll
        for (int c = 0; c < H; c++) s = fmaf(Ws[o * 2 * H + H + c], hi[c], s);
        tmp[o] = s;
      }
#pragma unroll
      for (int c = 0; c < H; c++) ss[c] = tmp[c];
    }
#pragma unroll
    for (int c = 0; c < H; c++) super[(size_t)c * HW + p] = ss[c];
    // fp16 channels-last [HW, 16], channel 15 = 0: two 16-byte stores per pixel
    __align__(16) __half o16[16];
#pragma unroll
    for (int c = 0; c < H; c++) o16[c] = __float2half(ss[c]);
    o16[15] = __float2half(0.0f);
    uint4* dst = reinterpret_cast<uint4*>(out + (size_t)p * 16);
    dst[0] = reinterpret_cast<const uint4*>(o16)[0];
    dst[1] = reinterpret_cast<const uint4*>(o16)[1];
  }
}

}  // namespace rvo

using namespace rvo;

extern "C" int rvo_scene_lstm_params_floats(int Ce, int Ci) {
  const int H = kSlHid, G = 4 * kSlHid;
  if (Ce < 1 || Ci < 1 || Ce > kSlMaxIn || Ci > kSlMaxIn) return -1;
  return (G * Ce + G * H + G) + (G * Ci + G * H + G) + (H * 2 * H + H);
}

extern "C" int rvo_scene_lstm_forward(const float* params, int Ce, int Ci, const float* events, const float* image,
                                      int H, int W, float* state_ev, float* state_im, float* super_state,
                                      int32_t* flags, int first, void* out16, void* stream) {
  RVO_CHECK_ARG(params && events && image && state_ev && state_im && super_state && flags && out16,
                "rvo_scene_lstm_forward: null pointer");
  RVO_CHECK_ARG(Ce == 5 && Ci == 3, "rvo_scene_lstm_forward: %d event bins / %d image channels (built for the "
                                    "5-bin stack + RGB of every shipped config)", Ce, Ci);
  RVO_CHECK_ARG(H >= 1 && W >= 1, "rvo_scene_lstm_forward: %dx%d", H, W);
  RVO_CHECK_ARG((reinterpret_cast<uintptr_t>(out16) & 15u) == 0, "rvo_scene_lstm_forward: output alignment");
  cudaStream_t st = (cudaStream_t)stream;
  const int64_t HW = (int64_t)H * W;
  RVO_CUDA(cudaMemsetAsync(flags, 0, 2 * sizeof(int32_t), st));
  scene_presence_kernel<<<sm_budget() * 4, 256, 0, st>>>(events, HW * Ce, image, HW * Ci, flags);
  RVO_LAUNCH_CHECK("scene_presence_kernel");
  int grid = cdiv(HW, 128);
  if (grid > sm_budget() * 16) grid = sm_budget() * 16;
  scene_lstm_kernel<5, 3><<<grid, 128, 0, st>>>(params, events, image, HW, state_ev, state_im, super_state, flags,
                                                first, (__half*)out16);
  RVO_LAUNCH_CHECK("scene_lstm_kernel");
  return RVO_OK;
}
