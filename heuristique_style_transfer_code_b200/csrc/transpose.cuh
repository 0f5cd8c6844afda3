// Batched 2-D transpose with optional fp32 -> bf16 cast:  out[b][s][r] = cast(in[b][r][s]), output rows `out_pitch`
// elements apart (>= R: lets the caller pad C x HW rows to the 16 B multiple a TMA tensor map needs, e.g. HW = 196 bf16).
//
// Hand-off between a channels_last (NHWC) backbone and the Gram kernels: the reference's
// gram_matrix() starts with activations.view(b, ch, h*w) (Models/Models_RESNET50_TRUNCATE_GRAM_with_Attention.py:27-28),
// i.e. it needs every image as a C x HW matrix with HW contiguous. A channels_last activation is the HW x C transpose
// of that; this kernel produces the C x HW view in one HBM-bound pass (and, for the backward, turns the fp32 NCHW
// gradient back into the NHWC tensor in the activation's dtype), instead of torch's generic strided copy.
// 64 x 64 tiles through shared memory: both the global reads and the global writes are full 128 B (bf16) / 256 B (fp32)
// row segments; the tile row pitch of 65 words keeps the transposed shared-memory reads conflict-free.
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>

namespace gh {

template <typename TI>
__device__ __forceinline__ float tr_load(const TI* p);
template <>
__device__ __forceinline__ float tr_load<float>(const float* p) { return __ldg(p); }
template <>
__device__ __forceinline__ float tr_load<__nv_bfloat16>(const __nv_bfloat16* p) {
  return __bfloat162float(__ldg(p));
}
template <typename TO>
__device__ __forceinline__ void tr_store(TO* p, float v);
template <>
__device__ __forceinline__ void tr_store<float>(float* p, float v) { *p = v; }
template <>
__device__ __forceinline__ void tr_store<__nv_bfloat16>(__nv_bfloat16* p, float v) { *p = __float2bfloat16_rn(v); }

// grid (ceil(S/64), ceil(R/64), B), block (64, 4)
template <typename TI, typename TO>
__global__ void __launch_bounds__(256) transpose_cast_kernel(const TI* __restrict__ in, TO* __restrict__ out, int R, int S,
                                                             long long out_pitch) {
  __shared__ float tile[64][65];
  const long long img = (long long)blockIdx.z * R * S;
  const long long img_out = (long long)blockIdx.z * S * out_pitch;
  const int s0 = blockIdx.x * 64, r0 = blockIdx.y * 64;
  const int tx = threadIdx.x, ty = threadIdx.y;
#pragma unroll
  for (int i = 0; i < 16; ++i) {
    const int r = r0 + ty + 4 * i, s = s0 + tx;
    if (r < R && s < S) tile[ty + 4 * i][tx] = tr_load<TI>(in + img + (long long)r * S + s);
  }
  __syncthreads();
#pragma unroll
  for (int i = 0; i < 16; ++i) {
    const int s = s0 + ty + 4 * i, r = r0 + tx;
    if (r < R && s < S) tr_store<TO>(out + img_out + (long long)s * out_pitch + r, tile[tx][ty + 4 * i]);
  }
}

}  // namespace gh
