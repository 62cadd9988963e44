// Convolutional front-end for the Atari-shaped configs.
//
// The reference's live networks are MLPs over flat observations; its only convolutional spec is the dead draft
// .ipynb_checkpoints/models-checkpoint.py:48-66 (Nature-CNN trunk: Conv 8/4 -> 4/2 -> 3/1 -> Flatten -> Linear(3136, 512))
// and :93-121 (the RND conv stacks).  There is no reference arithmetic to pin against, so parity is DEFINED against
// torch.nn.Conv2d evaluated in fp64 (tests/test_gpu_conv.py).
//
// A convolution is lowered to the dense layers this library already has: im2col gathers every receptive field into a
// row, `cols [N*OH*OW, C*KH*KW]`, the layer itself is then ppx_tc_linear / ppx_linear_* on (cols, W [C*KH*KW, Cout]) --
// forward, weight gradient (cols^T dY) and data gradient (dY W^T) all reuse the parity-checked GEMM kernels and their
// fused bias / activation epilogues -- and col2im folds the data gradient of the rows back onto the input.  The output
// of a layer, [N*OH*OW, Cout], IS the next layer's input in NHWC, so only the first layer reads the frames in the
// caller's NCHW order.  No padding, square stride (what the draft uses).
//
// Both kernels are plain HBM streams: im2col reads the input once through L1/L2 (a pixel is re-read by up to
// (K/stride)^2 patches, from cache) and writes KH*KW*C floats per output position; col2im is written as a GATHER (one
// thread per input element sums the <= ceil(K/stride)^2 patch entries that cover it, in a fixed order), so it needs no
// atomics and is deterministic.
#include "common.cuh"

namespace ppx {
namespace {

struct ConvShape {
  int N, C, H, W, KH, KW, stride, OH, OW;
  int nchw;            // 1: x is [N,C,H,W] and a patch is ordered (c, kh, kw) -- torch's weight.view(Cout, -1) order
                       // 0: x is [N,H,W,C] and a patch is ordered (kh, kw, c)
};

__device__ __forceinline__ int64_t x_index(const ConvShape& s, int n, int c, int h, int w) {
  return s.nchw ? (((int64_t)n * s.C + c) * s.H + h) * s.W + w : (((int64_t)n * s.H + h) * s.W + w) * s.C + c;
}

// one thread per element of cols; consecutive threads = consecutive k of one row (coalesced stores; for NHWC the loads of
// a (kh, kw) group are contiguous in c as well)
__global__ void __launch_bounds__(256) im2col_kernel(ConvShape s, const float* __restrict__ x, float* __restrict__ cols) {
  const int K = s.C * s.KH * s.KW;
  const int64_t total = (int64_t)s.N * s.OH * s.OW * K;
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
    const int k = (int)(e % K);
    const int64_t row = e / K;
    const int ow = (int)(row % s.OW), oh = (int)((row / s.OW) % s.OH), n = (int)(row / ((int64_t)s.OW * s.OH));
    int c, kh, kw;
    if (s.nchw) { kw = k % s.KW; kh = (k / s.KW) % s.KH; c = k / (s.KW * s.KH); }
    else { c = k % s.C; kw = (k / s.C) % s.KW; kh = k / (s.C * s.KW); }
    cols[e] = __ldg(x + x_index(s, n, c, oh * s.stride + kh, ow * s.stride + kw));
  }
}

// four consecutive k per thread (16-byte loads and stores): valid when a patch's innermost run is a multiple of 4 floats and
// 16-byte aligned in the source -- NHWC with C % 4 == 0, or NCHW with KW % 4 == 0, stride % 4 == 0, W % 4 == 0
__global__ void __launch_bounds__(256) im2col_vec4_kernel(ConvShape s, const float* __restrict__ x, float* __restrict__ cols) {
  const int K = s.C * s.KH * s.KW, K4 = K / 4;
  const int64_t total = (int64_t)s.N * s.OH * s.OW * K4;
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
    const int k = (int)(e % K4) * 4;
    const int64_t row = e / K4;
    const int ow = (int)(row % s.OW), oh = (int)((row / s.OW) % s.OH), n = (int)(row / ((int64_t)s.OW * s.OH));
    int c, kh, kw;
    if (s.nchw) { kw = k % s.KW; kh = (k / s.KW) % s.KH; c = k / (s.KW * s.KH); }
    else { c = k % s.C; kw = (k / s.C) % s.KW; kh = k / (s.C * s.KW); }
    const float4 v = __ldg(reinterpret_cast<const float4*>(x + x_index(s, n, c, oh * s.stride + kh, ow * s.stride + kw)));
    *reinterpret_cast<float4*>(cols + row * K + k) = v;
  }
}

// one thread per element of dx (in x's own layout): dx[n,c,h,w] = sum over the patches (oh, ow) and taps (kh, kw) with
// oh*stride + kh == h, ow*stride + kw == w of dcols[(n,oh,ow), k(c,kh,kw)], kh then kw ascending
__global__ void __launch_bounds__(256) col2im_kernel(ConvShape s, const float* __restrict__ dcols, float* __restrict__ dx) {
  const int K = s.C * s.KH * s.KW;
  const int64_t total = (int64_t)s.N * s.C * s.H * s.W;
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
    int n, c, h, w;
    if (s.nchw) { w = (int)(e % s.W); h = (int)((e / s.W) % s.H); c = (int)((e / ((int64_t)s.W * s.H)) % s.C); n = (int)(e / ((int64_t)s.W * s.H * s.C)); }
    else { c = (int)(e % s.C); w = (int)((e / s.C) % s.W); h = (int)((e / ((int64_t)s.C * s.W)) % s.H); n = (int)(e / ((int64_t)s.C * s.W * s.H)); }
    float acc = 0.f;
    for (int kh = h % s.stride; kh < s.KH; kh += s.stride) {
      const int oh = (h - kh) / s.stride;
      if (h - kh < 0 || oh >= s.OH) continue;
      for (int kw = w % s.stride; kw < s.KW; kw += s.stride) {
        const int ow = (w - kw) / s.stride;
        if (w - kw < 0 || ow >= s.OW) continue;
        const int k = s.nchw ? (c * s.KH + kh) * s.KW + kw : (kh * s.KW + kw) * s.C + c;
        acc += __ldg(dcols + (((int64_t)n * s.OH + oh) * s.OW + ow) * K + k);
      }
    }
    dx[e] = acc;
  }
}

__device__ __forceinline__ float act_deriv(float h, int act) {       // derivative through the post-activation value h
  switch (act) {
    case PPX_ACT_TANH: return 1.f - h * h;
    case PPX_ACT_LEAKY_RELU: return h > 0.f ? 1.f : 0.01f;
    case PPX_ACT_ELU: return h > 0.f ? 1.f : h + 1.f;
    case PPX_ACT_RELU: return h > 0.f ? 1.f : 0.f;
    default: return 1.f;
  }
}
__global__ void __launch_bounds__(256) act_bwd_mul_kernel(const float* __restrict__ d, const float* __restrict__ h, int64_t n, int act,
                                                          float* __restrict__ out) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    out[i] = d[i] * act_deriv(h[i], act);
}

int make_shape(ConvShape* s, int nchw, int N, int C, int H, int W, int KH, int KW, int stride, const char* who) {
  PPX_REQUIRE(N >= 1 && C >= 1 && H >= 1 && W >= 1 && KH >= 1 && KW >= 1 && stride >= 1 && KH <= H && KW <= W,
              "%s: bad shape N=%d C=%d H=%d W=%d K=%dx%d stride=%d", who, N, C, H, W, KH, KW, stride);
  *s = ConvShape{N, C, H, W, KH, KW, stride, (H - KH) / stride + 1, (W - KW) / stride + 1, nchw ? 1 : 0};
  return PPX_OK;
}

inline unsigned grid_for(int64_t total) { return (unsigned)std::min<int64_t>(ceil_div(total, 256), (int64_t)sm_count() * 16); }

}  // namespace
}  // namespace ppx

using namespace ppx;

extern "C" int ppx_im2col(const float* x, int nchw, int N, int C, int H, int W, int KH, int KW, int stride, float* cols,
                          void* stream) {
  PPX_REQUIRE(x && cols, "im2col: null pointer");
  ConvShape s;
  int rc = make_shape(&s, nchw, N, C, H, W, KH, KW, stride, "im2col");
  if (rc) return rc;
  const int64_t total = (int64_t)N * s.OH * s.OW * C * KH * KW;
  const bool aligned = (((uintptr_t)x | (uintptr_t)cols) & 15) == 0;
  const bool vec = aligned && (nchw ? (KW % 4 == 0 && stride % 4 == 0 && W % 4 == 0) : (C % 4 == 0));
  if (vec) im2col_vec4_kernel<<<grid_for(total / 4), 256, 0, (cudaStream_t)stream>>>(s, x, cols);
  else im2col_kernel<<<grid_for(total), 256, 0, (cudaStream_t)stream>>>(s, x, cols);
  return after_launch("im2col");
}

extern "C" int ppx_col2im(const float* dcols, int nchw, int N, int C, int H, int W, int KH, int KW, int stride, float* dx,
                          void* stream) {
  PPX_REQUIRE(dcols && dx, "col2im: null pointer");
  ConvShape s;
  int rc = make_shape(&s, nchw, N, C, H, W, KH, KW, stride, "col2im");
  if (rc) return rc;
  col2im_kernel<<<grid_for((int64_t)N * C * H * W), 256, 0, (cudaStream_t)stream>>>(s, dcols, dx);
  return after_launch("col2im");
}

extern "C" int ppx_act_bwd_mul(const float* d, const float* h, int64_t n, int act, float* out, void* stream) {
  PPX_REQUIRE(d && h && out && n >= 0, "act_bwd_mul: bad arguments");
  if (n == 0) return PPX_OK;
  act_bwd_mul_kernel<<<grid_for(n), 256, 0, (cudaStream_t)stream>>>(d, h, n, act, out);
  return after_launch("act_bwd_mul");
}
