// Max pooling over channels_last (NHWC) activations for the inference plan of the encoder (frozen_encoder.py): the
// stem's nn.MaxPool2d(3, stride 2, padding 1) (torchvision resnet, child 3 of the truncated encoder the reference
// builds at Models_RESNET50_TRUNCATE_GRAM_with_Attention.py:17). ATen's max_pool_forward_nhwc takes 0.9 ms for the
// 256 x 112 x 112 x 64 fp32 stem output (it also produces the argmax indices backward needs); inference needs only
// the values: one 16-byte channel vector per thread, the k x k window read through L1/L2, 1.03 GB of compulsory HBM
// traffic. Max is exact, so the result is bit-identical to the reference's.
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace gh {

struct MaxPoolParams {
  const void* in;
  void* out;
  int B, H, W, C, OH, OW;
  int k, stride, pad;
  long long total_vec;      // B * OH * OW * (C / VEC)
};

template <typename T> struct PoolVec;
template <> struct PoolVec<float> {
  static constexpr int kElems = 4;
  __device__ static void max_into(float4& a, const float4& b) {
    a.x = (b.x > a.x || b.x != b.x) ? b.x : a.x;      // NaN-propagating, like ATen's (val > max) || isnan(val)
    a.y = (b.y > a.y || b.y != b.y) ? b.y : a.y;
    a.z = (b.z > a.z || b.z != b.z) ? b.z : a.z;
    a.w = (b.w > a.w || b.w != b.w) ? b.w : a.w;
  }
  __device__ static float4 lowest() { const float m = -__int_as_float(0x7f800000); return make_float4(m, m, m, m); }
};
template <> struct PoolVec<__nv_bfloat16> {
  static constexpr int kElems = 8;
  __device__ static void max_into(float4& a, const float4& b) {
    __nv_bfloat162* pa = reinterpret_cast<__nv_bfloat162*>(&a);
    const __nv_bfloat162* pb = reinterpret_cast<const __nv_bfloat162*>(&b);
#pragma unroll
    for (int i = 0; i < 4; ++i) pa[i] = __hmax2_nan(pa[i], pb[i]);
  }
  __device__ static float4 lowest() {
    const uint32_t m = 0xff80ff80u;                      // two bf16 -inf
    const float f = __uint_as_float(m);
    return make_float4(f, f, f, f);
  }
};

// grid (ceil(OW / ppc), ceil(OH / kPoolRows), B), ppc = 256 / (C / VEC) output columns per CTA: no index divisions beyond
// one 32-bit one per thread. A thread walks kPoolRows output rows of its (column, channel vector): the horizontal maximum
// of every input row is computed once and rows shared by two vertically adjacent windows (stride < k) stay in registers.
constexpr int kPoolRows = 4;

template <typename T>
__global__ void __launch_bounds__(256) maxpool2d_nhwc_kernel(const MaxPoolParams p) {
  constexpr int VEC = PoolVec<T>::kElems;
  const int cv = p.C / VEC;
  const int ppc = 256 / cv;
  const int t = threadIdx.x;
  const int oxl = t / cv, c = t - oxl * cv;
  const int ox = blockIdx.x * ppc + oxl;
  if (oxl >= ppc || ox >= p.OW) return;
  const int b = blockIdx.z;
  const int oy0 = blockIdx.y * kPoolRows;
  const float4* __restrict__ in = reinterpret_cast<const float4*>(p.in) + (long long)b * p.H * p.W * cv + c;
  float4* __restrict__ out = reinterpret_cast<float4*>(p.out) + ((long long)b * p.OH * p.OW) * cv + c;
  const int x0 = ox * p.stride - p.pad;

  auto row_max = [&](int y) {                           // max over the window's columns of input row y (valid rows only)
    float4 m = PoolVec<T>::lowest();
    const float4* row = in + (long long)y * p.W * cv;
    for (int dx = 0; dx < p.k; ++dx) {
      const int x = x0 + dx;
      if (x >= 0 && x < p.W) PoolVec<T>::max_into(m, __ldg(row + (long long)x * cv));
    }
    return m;
  };

  if (p.k == 3 && p.stride == 2) {                      // the ResNet stem: rows 2oy-1, 2oy, 2oy+1; the last is shared
    int y = oy0 * 2 - p.pad;
    float4 carry = (y >= 0 && y < p.H) ? row_max(y) : PoolVec<T>::lowest();
#pragma unroll
    for (int r = 0; r < kPoolRows; ++r) {
      const int oy = oy0 + r;
      if (oy >= p.OH) break;
      const int y1 = oy * 2 - p.pad + 1, y2 = y1 + 1;
      const float4 a = (y1 >= 0 && y1 < p.H) ? row_max(y1) : PoolVec<T>::lowest();
      const float4 d = (y2 >= 0 && y2 < p.H) ? row_max(y2) : PoolVec<T>::lowest();
      float4 m = carry;
      PoolVec<T>::max_into(m, a);
      PoolVec<T>::max_into(m, d);
      out[((long long)oy * p.OW + ox) * cv] = m;
      carry = d;
    }
    return;
  }
  for (int r = 0; r < kPoolRows; ++r) {                 // any other window
    const int oy = oy0 + r;
    if (oy >= p.OH) break;
    float4 m = PoolVec<T>::lowest();
    for (int dy = 0; dy < p.k; ++dy) {
      const int y = oy * p.stride - p.pad + dy;
      if (y >= 0 && y < p.H) PoolVec<T>::max_into(m, row_max(y));
    }
    out[((long long)oy * p.OW + ox) * cv] = m;
  }
}

// ----------------------------------------------------------------------------------------------------------------------
// Space-to-depth staging of the image batch for the stem of the inference plan. cuDNN has no good NHWC kernel for the
// stem's 7x7 stride-2 convolution over 3 channels (1.33 ms at batch 256, ~9x off its roofline); the same convolution
// written as a 4x4 stride-1 convolution over the 2x2 space-to-depth image (12 channels, padded to 16) runs on its
// regular tensor-core kernels in 0.58 ms. With tap i' = i + 1 = 2a + r the input row 2y + i - 3 is row r of cell
// y + a - 2, so  z[b, Y + 2, X + 2, c*4 + r*2 + s] = x[b, c, 2Y + r, 2X + s]  inside a zero border of 2 (top/left) and
// 1 (bottom/right) makes it an unpadded convolution; channels 12..15 are zero. One thread per cell: 6 8-byte loads
// (x-contiguous input) or 12 scalar ones, 64 bytes (fp32) / 32 bytes (bf16) stored.
struct StemS2DParams {
  const float* x;
  void* z;
  long long s_img, s_c, s_y, s_x;     // element strides of x
  int B, H, W;                        // H, W even
  int HP, WP;                         // H/2 + 3, W/2 + 3
  long long total;                    // B * HP * WP
};

template <typename TO>
__global__ void __launch_bounds__(256) stem_space_to_depth_kernel(const StemS2DParams p) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < p.total; i += (long long)gridDim.x * blockDim.x) {
    const int xp = (int)(i % p.WP);
    long long r = i / p.WP;
    const int yp = (int)(r % p.HP);
    const int b = (int)(r / p.HP);
    float v[16];
#pragma unroll
    for (int k = 0; k < 16; ++k) v[k] = 0.f;
    const int Y = yp - 2, X = xp - 2;
    if (Y >= 0 && Y < p.H / 2 && X >= 0 && X < p.W / 2) {
      const float* base = p.x + (long long)b * p.s_img + (long long)(2 * Y) * p.s_y + (long long)(2 * X) * p.s_x;
#pragma unroll
      for (int c = 0; c < 3; ++c)
#pragma unroll
        for (int rr = 0; rr < 2; ++rr) {
          const float* q = base + (long long)c * p.s_c + (long long)rr * p.s_y;
          if (p.s_x == 1 && (reinterpret_cast<uintptr_t>(q) & 7) == 0) {
            const float2 t = __ldg(reinterpret_cast<const float2*>(q));
            v[c * 4 + rr * 2] = t.x;
            v[c * 4 + rr * 2 + 1] = t.y;
          } else {
            v[c * 4 + rr * 2] = __ldg(q);
            v[c * 4 + rr * 2 + 1] = __ldg(q + p.s_x);
          }
        }
    }
    if constexpr (sizeof(TO) == 4) {
      float4* out = reinterpret_cast<float4*>(p.z) + i * 4;
#pragma unroll
      for (int k = 0; k < 4; ++k) out[k] = make_float4(v[4 * k], v[4 * k + 1], v[4 * k + 2], v[4 * k + 3]);
    } else {
      uint4* out = reinterpret_cast<uint4*>(p.z) + i * 2;
#pragma unroll
      for (int k = 0; k < 2; ++k) {
        uint4 w;
        __nv_bfloat162 h0 = __floats2bfloat162_rn(v[8 * k], v[8 * k + 1]), h1 = __floats2bfloat162_rn(v[8 * k + 2], v[8 * k + 3]);
        __nv_bfloat162 h2 = __floats2bfloat162_rn(v[8 * k + 4], v[8 * k + 5]), h3 = __floats2bfloat162_rn(v[8 * k + 6], v[8 * k + 7]);
        w.x = *reinterpret_cast<uint32_t*>(&h0); w.y = *reinterpret_cast<uint32_t*>(&h1);
        w.z = *reinterpret_cast<uint32_t*>(&h2); w.w = *reinterpret_cast<uint32_t*>(&h3);
        out[k] = w;
      }
    }
  }
}

}  // namespace gh
