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

template <typename T>
__global__ void __launch_bounds__(256) maxpool2d_nhwc_kernel(const MaxPoolParams p) {
  constexpr int VEC = PoolVec<T>::kElems;
  const int cv = p.C / VEC;
  const float4* __restrict__ in = reinterpret_cast<const float4*>(p.in);
  float4* __restrict__ out = reinterpret_cast<float4*>(p.out);
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < p.total_vec;
       i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(i % cv);
    long long r = i / cv;
    const int ox = (int)(r % p.OW); r /= p.OW;
    const int oy = (int)(r % p.OH);
    const int b = (int)(r / p.OH);
    const int y0 = oy * p.stride - p.pad, x0 = ox * p.stride - p.pad;
    float4 m = PoolVec<T>::lowest();
    for (int dy = 0; dy < p.k; ++dy) {
      const int y = y0 + dy;
      if (y < 0 || y >= p.H) continue;
      for (int dx = 0; dx < p.k; ++dx) {
        const int x = x0 + dx;
        if (x < 0 || x >= p.W) continue;
        PoolVec<T>::max_into(m, __ldg(in + (((long long)b * p.H + y) * p.W + x) * cv + c));
      }
    }
    out[i] = m;
  }
}

}  // namespace gh
