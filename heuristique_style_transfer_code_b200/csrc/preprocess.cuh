// Camera-mode preprocessing on the GPU: one kernel turns a raw BGR uint8 frame into the normalised fp32 CHW tensor the
// encoder consumes, bit-identically to what the reference does on the host per frame
// (functions/functions_RESNET50_Truncate_Gram_Attention.py:499-501: cv2.cvtColor(BGR2RGB) -> Image.fromarray ->
// transform; test_RESNET50_Truncate_gram_attention.py:61-66: Resize [+ CenterCrop] + ToTensor + Normalize).
//
// torchvision's Resize on a PIL image is Pillow's ImagingResample with the bilinear (triangle) filter: separable,
// support widened by the down-scale factor (antialiasing), 8-bit fixed point -- coefficients with 22 fractional bits,
// a horizontal pass rounded to uint8, then a vertical pass rounded to uint8. The coefficient tables depend only on the
// sizes; the host computes them once (streaming.py) and a CenterCrop is just the sub-range of output coordinates the
// tables are built for. One thread = one output pixel (three channels): for every contributing input row it forms
// the horizontally filtered uint8, then filters those vertically, divides by 255 and normalises in IEEE fp32.
// The frame is read straight from the uploaded uint8 buffer (6 MB at 1080p, L2 resident); nothing else touches HBM.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace gh {

constexpr int kPreprocPrecisionBits = 32 - 8 - 2;

struct PreprocParams {
  const uint8_t* frame;     // (H, W, 3) uint8, `pitch` bytes between rows
  long long pitch;
  int H, W;
  int OH, OW;               // output (after crop) extent
  const int* hx_min;        // [OW]   first contributing input column
  const int* hx_size;       // [OW]   number of contributing columns
  const int* hk;            // [OW][hkmax] integer coefficients (sum = 2^22)
  int hkmax;
  const int* vy_min;        // [OH]
  const int* vy_size;       // [OH]
  const int* vk;            // [OH][vkmax]
  int vkmax;
  float* out;               // (3, OH, OW) fp32
  int bgr;                  // 1: the frame is BGR and output channel c reads input channel 2 - c (cv2.cvtColor BGR2RGB)
  float mean[3], std[3];
};

__device__ __forceinline__ int preproc_clip8(int acc) {
  const int v = acc >> kPreprocPrecisionBits;
  return v < 0 ? 0 : (v > 255 ? 255 : v);
}

__global__ void __launch_bounds__(256) preprocess_frame_kernel(const PreprocParams p) {
  const int ox = blockIdx.x * blockDim.x + threadIdx.x;
  const int oy = blockIdx.y;
  if (ox >= p.OW) return;
  const int x0 = p.hx_min[ox], nx = p.hx_size[ox];
  const int y0 = p.vy_min[oy], ny = p.vy_size[oy];
  const int* __restrict__ hk = p.hk + (long long)ox * p.hkmax;
  const int* __restrict__ vk = p.vk + (long long)oy * p.vkmax;
  int acc[3] = {1 << (kPreprocPrecisionBits - 1), 1 << (kPreprocPrecisionBits - 1), 1 << (kPreprocPrecisionBits - 1)};
  for (int j = 0; j < ny; ++j) {
    const uint8_t* __restrict__ row = p.frame + (long long)(y0 + j) * p.pitch + (long long)x0 * 3;
    int h0 = 1 << (kPreprocPrecisionBits - 1), h1 = h0, h2 = h0;
    for (int i = 0; i < nx; ++i) {
      const int k = __ldg(hk + i);
      h0 += (int)__ldg(row + 3 * i) * k;
      h1 += (int)__ldg(row + 3 * i + 1) * k;
      h2 += (int)__ldg(row + 3 * i + 2) * k;
    }
    const int kv = __ldg(vk + j);
    acc[0] += preproc_clip8(h0) * kv;
    acc[1] += preproc_clip8(h1) * kv;
    acc[2] += preproc_clip8(h2) * kv;
  }
  const long long plane = (long long)p.OH * p.OW;
  float* o = p.out + (long long)oy * p.OW + ox;
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    const int src = p.bgr ? 2 - c : c;
    const float v = (float)preproc_clip8(acc[src]) / 255.0f;          // ToTensor
    o[c * plane] = (v - p.mean[c]) / p.std[c];                        // Normalize (IEEE division: no fast-math)
  }
}

// ---- loader-side counterpart: a uint8 image batch normalised on the device -------------------------------------------------
// The reference's loaders end in ToTensor + Normalize on the host (test_RESNET50_Truncate_gram_attention.py:64-65,
// train_best_RESNET50_Truncate_gram_attention.py:42-43) and every batch crosses PCIe as fp32: 154 MB for 256 images at
// 224 x 224, 2.8 ms at the 55 GB/s the host link gives one GPU -- hidden behind the forward then, but the bound of the step
// when eight ranks share the host (25 GB/s each). Uploading the uint8 pixels (38.5 MB) and applying the two transforms
// here moves a quarter of the bytes; the arithmetic is the host's: float(b) / 255 (ToTensor), (v - mean[c]) / std[c]
// (Normalize), IEEE fp32 with correctly rounded divisions, so the fp32 batch is bit-identical to the loader's.
// A byte has 256 values and a plane one (mean, std): each CTA works inside ONE plane (grid x = plane, grid y = chunk of it,
// so no 64-bit division per element), builds the plane's 256-entry table of results once -- thread t computes
// ((float)t / 255 - mean) / std with the two correctly rounded divisions -- and the pixels become shared-memory lookups:
// the arithmetic per value is the host's, done once per byte value instead of once per pixel (the first version, which
// divided per pixel behind a 64-bit channel computation, reached 0.35 of the HBM roofline:
// profiles/r4b_normalize_u8_timing_first_version.log). One thread = kNormUnroll groups of VEC pixels, 256 threads apart:
// all loads of a thread are issued before the first use; with VEC = 4 a warp loads 128 B and stores 512 B contiguous per
// instruction.
struct NormalizeU8Params {
  const uint8_t* src;       // (N, C, H, W) uint8, dense
  float* dst;               // (N, C, H, W) fp32, dense
  long long hw;             // H * W
  int groups;               // per plane: hw / VEC
  int C;
  float mean[4], std[4];
};

constexpr int kNormUnroll = 8;

template <int VEC>
__global__ void __launch_bounds__(256) normalize_u8_kernel(const NormalizeU8Params p) {
  __shared__ float table[256];
  const long long plane = blockIdx.x;
  const int c = (int)(blockIdx.x % (unsigned)p.C);
  const float m = c == 0 ? p.mean[0] : c == 1 ? p.mean[1] : c == 2 ? p.mean[2] : p.mean[3];
  const float s = c == 0 ? p.std[0] : c == 1 ? p.std[1] : c == 2 ? p.std[2] : p.std[3];
  table[threadIdx.x] = __fdiv_rn(__fdiv_rn((float)threadIdx.x, 255.0f) - m, s);     // ToTensor, Normalize
  __syncthreads();
  const uint8_t* __restrict__ src = p.src + plane * p.hw;
  float* __restrict__ dst = p.dst + plane * p.hw;
  const int first = blockIdx.y * (256 * kNormUnroll) + threadIdx.x;
  if (VEC == 4) {
    unsigned raw[kNormUnroll];
#pragma unroll
    for (int j = 0; j < kNormUnroll; ++j) {
      const int g = first + j * 256;
      raw[j] = g < p.groups ? __ldg(reinterpret_cast<const unsigned*>(src) + g) : 0u;
    }
#pragma unroll
    for (int j = 0; j < kNormUnroll; ++j) {
      const int g = first + j * 256;
      if (g < p.groups) {
        float4 o;
        o.x = table[raw[j] & 0xffu];
        o.y = table[(raw[j] >> 8) & 0xffu];
        o.z = table[(raw[j] >> 16) & 0xffu];
        o.w = table[raw[j] >> 24];
        __stcs(reinterpret_cast<float4*>(dst) + g, o);
      }
    }
  } else {
#pragma unroll
    for (int j = 0; j < kNormUnroll; ++j) {
      const int g = first + j * 256;
      if (g < p.groups) dst[g] = table[__ldg(src + g)];
    }
  }
}

}  // namespace gh
