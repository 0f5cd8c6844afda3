// Gram head of the Multi-PatchGAN discriminator (reference Models/Models_Multi_PatchGAN.py:198-256).
//
//   patch_pool_kernel      one warp per (image, collected layer, channel) plane: a single pass over the projected
//                          feature maps produces the 4x4 adaptive-average bins of every channel and the per-channel
//                          sum / sum of squares the folded-in :198 layer norm needs. HBM-bound: every map is read once
//                          (the reference reads it for layer_norm, writes it, reads it again for the pooling).
//   patch_gram_kernel      one CTA per (image, layer): whole-map layer norm applied to the 16 bin means, layer norm
//                          over (D, 4, 4), the D x D Gram over the 16 positions, its Frobenius norm.
//   patch_attention_kernel the two 8-head nn.MultiheadAttention layers over the L stacked layer tokens, the mean over
//                          layers and the classifier, one image at a time per CTA; every intermediate stays in shared
//                          memory (L <= 8 tokens of ndf <= 128 features), weights come through L1/L2.
// The Linear(D*D -> ndf) between the two is the tcgen05 split-bf16 GEMM (gh_gemm_f32) over all L*B Gram rows at once.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace gh {

constexpr int kPatchMaxLayers = 8;
constexpr int kPatchPool = 4;
constexpr int kPatchBins = kPatchPool * kPatchPool;   // 16 positions per channel after pooling
constexpr int kPatchMaxD = 128;
constexpr int kPatchThreads = 256;

struct PatchLayer {
  const float* x;
  int H, W;
  long long s_img, s_c, s_y, s_x;   // element strides (NCHW: C*H*W, H*W, W, 1)
};

struct PatchGramParams {
  PatchLayer layer[kPatchMaxLayers];
  int L, B, D;
  int ln_input;        // 1: apply F.layer_norm(x, x.shape[1:]) (:198) to the map before pooling
  float* gram;         // (L, B, D*D)
  float* gram_norm;    // (L, B)
};

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// sum over the CTA of one double per thread; result broadcast to every thread. red: >= 32 doubles of shared memory.
__device__ __forceinline__ double block_sum(double v, double* red) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  v = warp_sum(v);
  __syncthreads();
  if (lane == 0) red[warp] = v;
  __syncthreads();
  double t = (lane < (int)(blockDim.x >> 5)) ? red[lane] : 0.0;
  t = warp_sum(t);
  return t;
}

// ---- pass over the maps: one warp per (image, layer, channel) plane ---------------------------------------------------
// Fast path (rows contiguous, W <= 32*KMAX): lane <-> fixed columns x = lane + 32k, so a lane's column-bin membership is
// a constant mask; rows are walked in segments over which the row-bin membership is constant, and inside a segment the
// inner loop is one load, one add and one fma per element. Column sums are folded into the 16 bins once per segment.
template <int KMAX>
__device__ __forceinline__ void pool_plane_rows(const float* __restrict__ plane, int H, int W, long long s_y,
                                                const int* ys, const int* ye, const int* xs, const int* xe,
                                                float* acc, float& sum, float& sq) {
  const int lane = threadIdx.x & 31;
  unsigned colmask[KMAX];
  float colsum[KMAX];
  float rowacc[kPatchPool][KMAX];
#pragma unroll
  for (int i = 0; i < kPatchPool; ++i)
#pragma unroll
    for (int k = 0; k < KMAX; ++k) rowacc[i][k] = 0.f;
#pragma unroll
  for (int k = 0; k < KMAX; ++k) {
    const int x = lane + 32 * k;
    unsigned m = 0;
#pragma unroll
    for (int j = 0; j < kPatchPool; ++j) m |= (x >= xs[j] && x < xe[j] && x < W) ? (1u << j) : 0u;
    colmask[k] = m;
    colsum[k] = 0.f;
  }
  int y = 0;
  while (y < H) {
    unsigned rm = 0;
#pragma unroll
    for (int i = 0; i < kPatchPool; ++i) rm |= (y >= ys[i] && y < ye[i]) ? (1u << i) : 0u;
    int yend = H;                                     // the segment ends at the next bin boundary above y
#pragma unroll
    for (int i = 0; i < kPatchPool; ++i) {
      if (ys[i] > y && ys[i] < yend) yend = ys[i];
      if (ye[i] > y && ye[i] < yend) yend = ye[i];
    }
    for (int yy = y; yy < yend; yy += 4) {
      float v[4][KMAX];
#pragma unroll
      for (int r = 0; r < 4; ++r) {
        const float* row = plane + (long long)(yy + r) * s_y;
#pragma unroll
        for (int k = 0; k < KMAX; ++k) {
          const int x = lane + 32 * k;
          v[r][k] = (yy + r < yend && x < W) ? __ldg(row + x) : 0.f;
        }
      }
#pragma unroll
      for (int r = 0; r < 4; ++r)
#pragma unroll
        for (int k = 0; k < KMAX; ++k) {
          colsum[k] += v[r][k];
          sq = fmaf(v[r][k], v[r][k], sq);
        }
    }
#pragma unroll
    for (int i = 0; i < kPatchPool; ++i)
      if ((rm >> i) & 1u) {                            // warp-uniform
#pragma unroll
        for (int k = 0; k < KMAX; ++k) rowacc[i][k] += colsum[k];
      }
#pragma unroll
    for (int k = 0; k < KMAX; ++k) {
      sum += colsum[k];
      colsum[k] = 0.f;
    }
    y = yend;
  }
#pragma unroll
  for (int i = 0; i < kPatchPool; ++i)
#pragma unroll
    for (int j = 0; j < kPatchPool; ++j)
#pragma unroll
      for (int k = 0; k < KMAX; ++k) acc[i * kPatchPool + j] += ((colmask[k] >> j) & 1u) ? rowacc[i][k] : 0.f;
}

// Vector path (rows contiguous and 16 B aligned, W % 4 == 0, W <= 4*LPR): LPR lanes cover one row with a float4 each, so one
// load instruction fetches 32/LPR rows; U of them are in flight per lane before the first add.
template <int LPR, int U>
__device__ __forceinline__ void pool_plane_vec4(const float* __restrict__ plane, int H, int W, long long s_y,
                                                const int* ys, const int* ye, const int* xs, const int* xe,
                                                float* acc, float& sum, float& sq) {
  constexpr int RPI = 32 / LPR;                       // rows per load instruction
  const int lane = threadIdx.x & 31;
  const int cg = (lane % LPR) * 4, rsub = lane / LPR;
  const bool colvalid = cg < W;
  unsigned colmask[4];
  float colsum[4] = {0.f, 0.f, 0.f, 0.f};
  float rowacc[kPatchPool][4];
#pragma unroll
  for (int i = 0; i < kPatchPool; ++i)
#pragma unroll
    for (int e = 0; e < 4; ++e) rowacc[i][e] = 0.f;
#pragma unroll
  for (int e = 0; e < 4; ++e) {
    const int x = cg + e;
    unsigned m = 0;
#pragma unroll
    for (int j = 0; j < kPatchPool; ++j) m |= (x >= xs[j] && x < xe[j] && x < W) ? (1u << j) : 0u;
    colmask[e] = m;
  }
  int y = 0;
  while (y < H) {
    unsigned rm = 0;
#pragma unroll
    for (int i = 0; i < kPatchPool; ++i) rm |= (y >= ys[i] && y < ye[i]) ? (1u << i) : 0u;
    int yend = H;
#pragma unroll
    for (int i = 0; i < kPatchPool; ++i) {
      if (ys[i] > y && ys[i] < yend) yend = ys[i];
      if (ye[i] > y && ye[i] < yend) yend = ye[i];
    }
    for (int yy = y; yy < yend; yy += U * RPI) {
      float4 v[U];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const int row = yy + u * RPI + rsub;
        v[u] = (row < yend && colvalid) ? __ldg(reinterpret_cast<const float4*>(plane + (long long)row * s_y + cg))
                                        : make_float4(0.f, 0.f, 0.f, 0.f);
      }
#pragma unroll
      for (int u = 0; u < U; ++u) {
        colsum[0] += v[u].x; colsum[1] += v[u].y; colsum[2] += v[u].z; colsum[3] += v[u].w;
        sq = fmaf(v[u].x, v[u].x, sq); sq = fmaf(v[u].y, v[u].y, sq);
        sq = fmaf(v[u].z, v[u].z, sq); sq = fmaf(v[u].w, v[u].w, sq);
      }
    }
#pragma unroll
    for (int i = 0; i < kPatchPool; ++i)
      if ((rm >> i) & 1u) {                            // warp-uniform
#pragma unroll
        for (int e = 0; e < 4; ++e) rowacc[i][e] += colsum[e];
      }
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      sum += colsum[e];
      colsum[e] = 0.f;
    }
    y = yend;
  }
  // once per plane: this lane's four columns into the column bins they belong to
#pragma unroll
  for (int i = 0; i < kPatchPool; ++i)
#pragma unroll
    for (int j = 0; j < kPatchPool; ++j)
#pragma unroll
      for (int e = 0; e < 4; ++e) acc[i * kPatchPool + j] += ((colmask[e] >> j) & 1u) ? rowacc[i][e] : 0.f;
}

// Small planes (H*W <= kPatchSmallPlane): the walk above costs ~1000 instructions per plane whatever its size. Here
// the warp copies its plane into shared memory (accumulating sum and sum of squares on the way), then lane i < 16 adds
// up the (at most (H/4+1) x (W/4+1)) elements of bin i: ~10x fewer instructions for the 14x14 .. 28x28 maps.
constexpr int kPatchSmallPlane = 1024;

__device__ __forceinline__ float pool_plane_small(const float* __restrict__ plane, int H, int W, long long s_y,
                                                  long long s_x, float* buf, float& sum, float& sq) {
  const int lane = threadIdx.x & 31;
  const int n = H * W;
  if (s_x == 1 && s_y == W) {
    for (int i = lane; i < n; i += 32) {
      const float v = __ldg(plane + i);
      buf[i] = v;
      sum += v;
      sq = fmaf(v, v, sq);
    }
  } else {
    for (int i = lane; i < n; i += 32) {
      const int y = i / W, x = i - y * W;
      const float v = __ldg(plane + (long long)y * s_y + (long long)x * s_x);
      buf[i] = v;
      sum += v;
      sq = fmaf(v, v, sq);
    }
  }
  __syncwarp();
  float mean = 0.f;
  if (lane < kPatchBins) {
    const int by = lane / kPatchPool, bx = lane % kPatchPool;
    const int y0 = (by * H) / kPatchPool, y1 = ((by + 1) * H + kPatchPool - 1) / kPatchPool;
    const int x0 = (bx * W) / kPatchPool, x1 = ((bx + 1) * W + kPatchPool - 1) / kPatchPool;
    float s = 0.f;
    for (int y = y0; y < y1; ++y)
      for (int x = x0; x < x1; ++x) s += buf[y * W + x];
    mean = s / (float)((y1 - y0) * (x1 - x0));
  }
  __syncwarp();
  return mean;
}

// Any strides / any width: lanes along x, every element tested against the four column bins.
__device__ __forceinline__ void pool_plane_generic(const float* __restrict__ plane, int H, int W, long long s_y,
                                                   long long s_x, const int* ys, const int* ye, const int* xs,
                                                   const int* xe, float* acc, float& sum, float& sq) {
  const int lane = threadIdx.x & 31;
  for (int y = 0; y < H; ++y) {
    const float* row = plane + (long long)y * s_y;
    float col[kPatchPool] = {0.f, 0.f, 0.f, 0.f};
    for (int x = lane; x < W; x += 32) {
      const float v = __ldg(row + (long long)x * s_x);
      sum += v;
      sq = fmaf(v, v, sq);
#pragma unroll
      for (int j = 0; j < kPatchPool; ++j) col[j] += (x >= xs[j] && x < xe[j]) ? v : 0.f;
    }
#pragma unroll
    for (int i = 0; i < kPatchPool; ++i)
      if (y >= ys[i] && y < ye[i]) {                 // warp-uniform
#pragma unroll
        for (int j = 0; j < kPatchPool; ++j) acc[i * kPatchPool + j] += col[j];
      }
  }
}

// grid (B * ceil(D / 8), L), 8 warps: warp w of CTA (b, chunk) owns channel chunk*8 + w of image b in layer blockIdx.y.
// pooled: (L, B, D, 16) bin means; stats: (L, B, D, 2) per-channel sum and sum of squares (doubles).
__global__ void __launch_bounds__(kPatchThreads, 3) patch_pool_kernel(const PatchGramParams p, float* __restrict__ pooled,
                                                                   double* __restrict__ stats) {
  const int l = blockIdx.y;
  const PatchLayer ly = p.layer[l];
  const int H = ly.H, W = ly.W, D = p.D;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int chunks = (D + 7) / 8;
  const int b = blockIdx.x / chunks, c = (blockIdx.x % chunks) * 8 + warp;
  if (c >= D) return;
  int ys[kPatchPool], ye[kPatchPool], xs[kPatchPool], xe[kPatchPool];
#pragma unroll
  for (int i = 0; i < kPatchPool; ++i) {            // ATen adaptive bins: floor(i*n/4) .. ceil((i+1)*n/4)
    ys[i] = (i * H) / kPatchPool;  ye[i] = ((i + 1) * H + kPatchPool - 1) / kPatchPool;
    xs[i] = (i * W) / kPatchPool;  xe[i] = ((i + 1) * W + kPatchPool - 1) / kPatchPool;
  }
  const float* plane = ly.x + (long long)b * ly.s_img + (long long)c * ly.s_c;
  const long long plane_id = ((long long)l * p.B + b) * D + c;
  if (H * W <= kPatchSmallPlane) {                    // uniform over the CTA (one layer per blockIdx.y)
    extern __shared__ float small_planes[];
    float sum = 0.f, sq = 0.f;
    const float mean = pool_plane_small(plane, H, W, ly.s_y, ly.s_x, small_planes + warp * kPatchSmallPlane, sum, sq);
    if (lane < kPatchBins) pooled[plane_id * kPatchBins + lane] = mean;
    const float fsum = warp_sum(sum), fsq = warp_sum(sq);
    if (lane == 0) {
      stats[plane_id * 2] = (double)fsum;
      stats[plane_id * 2 + 1] = (double)fsq;
    }
    return;
  }
  float acc[kPatchBins];
#pragma unroll
  for (int i = 0; i < kPatchBins; ++i) acc[i] = 0.f;
  float sum = 0.f, sq = 0.f;
  const bool vec = ly.s_x == 1 && (W & 3) == 0 && (ly.s_y & 3) == 0 && (ly.s_c & 3) == 0 && (ly.s_img & 3) == 0 &&
                   (reinterpret_cast<uintptr_t>(ly.x) & 15) == 0;
  if (vec && W <= 32) pool_plane_vec4<8, 4>(plane, H, W, ly.s_y, ys, ye, xs, xe, acc, sum, sq);
  else if (vec && W <= 64) pool_plane_vec4<16, 7>(plane, H, W, ly.s_y, ys, ye, xs, xe, acc, sum, sq);
  else if (vec && W <= 128) pool_plane_vec4<32, 7>(plane, H, W, ly.s_y, ys, ye, xs, xe, acc, sum, sq);
  else if (ly.s_x == 1 && W <= 32) pool_plane_rows<1>(plane, H, W, ly.s_y, ys, ye, xs, xe, acc, sum, sq);
  else if (ly.s_x == 1 && W <= 64) pool_plane_rows<2>(plane, H, W, ly.s_y, ys, ye, xs, xe, acc, sum, sq);
  else if (ly.s_x == 1 && W <= 128) pool_plane_rows<4>(plane, H, W, ly.s_y, ys, ye, xs, xe, acc, sum, sq);
  else if (ly.s_x == 1 && W <= 256) pool_plane_rows<8>(plane, H, W, ly.s_y, ys, ye, xs, xe, acc, sum, sq);
  else pool_plane_generic(plane, H, W, ly.s_y, ly.s_x, ys, ye, xs, xe, acc, sum, sq);

  // halving butterfly: 16 values per lane -> lane (i << 1) holds bin i summed over the warp (16 shuffles instead of 80)
#pragma unroll
  for (int half = 8, o = 16; half >= 1; half >>= 1, o >>= 1) {
    const bool up = (lane & o) != 0;
#pragma unroll
    for (int i = 0; i < half; ++i) {
      const float send = up ? acc[i] : acc[i + half];
      const float keep = up ? acc[i + half] : acc[i];
      acc[i] = keep + __shfl_xor_sync(0xffffffffu, send, o);
    }
  }
  const float total = acc[0] + __shfl_xor_sync(0xffffffffu, acc[0], 1);
  const int bin = ((lane >> 4) & 1) * 8 + ((lane >> 3) & 1) * 4 + ((lane >> 2) & 1) * 2 + ((lane >> 1) & 1);
  if ((lane & 1) == 0) {
    const int by = bin / kPatchPool, bx = bin % kPatchPool;   // arithmetic, not ys[by]: keeps the bin tables in registers
    const int cnt = (((by + 1) * H + kPatchPool - 1) / kPatchPool - (by * H) / kPatchPool) *
                    (((bx + 1) * W + kPatchPool - 1) / kPatchPool - (bx * W) / kPatchPool);
    pooled[plane_id * kPatchBins + bin] = total / (float)cnt;
  }
  const float fsum = warp_sum(sum), fsq = warp_sum(sq);
  if (lane == 0) {
    stats[plane_id * 2] = (double)fsum;
    stats[plane_id * 2 + 1] = (double)fsq;
  }
}

// grid (B, L): whole-map layer norm folded in (:198), layer norm over (D, 4, 4) (:213), Gram (:217-220), norm (:223).
__global__ void __launch_bounds__(kPatchThreads) patch_gram_kernel(const PatchGramParams p, const float* __restrict__ pooled,
                                                                   const double* __restrict__ stats) {
  extern __shared__ float smem[];
  float* Z = smem;                                   // [D][16] pooled (then normalised) map
  double* red = reinterpret_cast<double*>(Z + p.D * kPatchBins);   // 32 doubles
  const int b = blockIdx.x, l = blockIdx.y;
  const int H = p.layer[l].H, W = p.layer[l].W, D = p.D;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
  const int n = D * kPatchBins;
  const long long base = ((long long)l * p.B + b) * D;
  for (int i = threadIdx.x; i < n; i += blockDim.x) Z[i] = pooled[base * kPatchBins + i];
  if (p.ln_input) {                                  // pool(LN(x)) = (pool(x) - mean) * rstd: pooling is linear
    double tsum = 0.0, tsq = 0.0;
    for (int c = threadIdx.x; c < D; c += blockDim.x) {
      tsum += stats[(base + c) * 2];
      tsq += stats[(base + c) * 2 + 1];
    }
    const double cnt = (double)D * H * W;
    const double s1 = block_sum(tsum, red);
    const double s2 = block_sum(tsq, red);
    const double mean = s1 / cnt;
    double var = s2 / cnt - mean * mean;
    var = var > 0.0 ? var : 0.0;
    const float fm = (float)mean, fr = (float)(1.0 / sqrt(var + 1e-5));
    for (int i = threadIdx.x; i < n; i += blockDim.x) Z[i] = (Z[i] - fm) * fr;   // each thread rewrites what it loaded
  }
  __syncthreads();

  // ---- layer norm over the (D, 4, 4) pooled map (:213) ----
  double ps = 0.0;
  for (int i = threadIdx.x; i < n; i += blockDim.x) ps += (double)Z[i];
  const float mean2 = (float)(block_sum(ps, red) / n);
  double pv = 0.0;
  for (int i = threadIdx.x; i < n; i += blockDim.x) { const float d = Z[i] - mean2; pv += (double)(d * d); }
  const float rstd2 = (float)(1.0 / sqrt(block_sum(pv, red) / n + 1e-5));
  for (int i = threadIdx.x; i < n; i += blockDim.x) Z[i] = (Z[i] - mean2) * rstd2;
  __syncthreads();

  // ---- Gram over the 16 positions (:217-220) and its Frobenius norm (:223) ----
  const float inv = (float)(1.0 / ((double)kPatchBins + 1e-6));
  float* out = p.gram + ((long long)l * p.B + b) * D * D;
  double fro = 0.0;
  for (int d0 = 0; d0 < D; d0 += 32) {               // lane <-> column d (coalesced stores), warp <-> rows c
    const int d = d0 + lane;
    float zd[kPatchBins];
#pragma unroll
    for (int k = 0; k < kPatchBins; ++k) zd[k] = (d < D) ? Z[d * kPatchBins + k] : 0.f;
    for (int c = warp; c < D; c += nwarps) {
      float s = 0.f;
#pragma unroll
      for (int k = 0; k < kPatchBins; k += 4) {
        const float4 t = *reinterpret_cast<const float4*>(&Z[c * kPatchBins + k]);   // broadcast
        s = fmaf(zd[k], t.x, s); s = fmaf(zd[k + 1], t.y, s); s = fmaf(zd[k + 2], t.z, s); s = fmaf(zd[k + 3], t.w, s);
      }
      s *= inv;
      if (d < D) {
        out[c * D + d] = s;
        fro += (double)(s * s);
      }
    }
  }
  const double f2 = block_sum(fro, red);
  if (threadIdx.x == 0) p.gram_norm[(long long)l * p.B + b] = (float)sqrt(f2);
}

// ----------------------------------------------------------------------------------------------------------------------
struct PatchAttnParams {
  const float* feat;                       // (L, B, E) projected Gram features
  const float* w_in[2];  const float* b_in[2];     // (3E, E), (3E)   attention_per_layer, attention_per_patch
  const float* w_out[2]; const float* b_out[2];    // (E, E), (E)
  const float* w_c;      const float* b_c;         // (nc, E), (nc)
  int L, B, E, heads, nc;
  float* emb;                              // (B, E)
  float* logits;                           // (B, nc)
};

// Y[l][j] = bias[j] + sum_k X[l][k] * Wt[j][k]   for l < L, j < N;  X, Y in shared memory (row pitch ldx / ldy), W global.
template <int MAXL>
__device__ __forceinline__ void rows_linear(const float* X, int ldx, const float* __restrict__ Wt, const float* __restrict__ bias,
                                            float* Y, int ldy, int L, int N, int K, float scale_first, int n_scaled) {
  for (int j = threadIdx.x; j < N; j += blockDim.x) {
    float acc[MAXL];
#pragma unroll
    for (int l = 0; l < MAXL; ++l) acc[l] = 0.f;
    const float4* w4 = reinterpret_cast<const float4*>(Wt + (long long)j * K);
    for (int k4 = 0; k4 < K / 4; ++k4) {
      const float4 w = __ldg(w4 + k4);
#pragma unroll
      for (int l = 0; l < MAXL; ++l)
        if (l < L) {
          const float4 x = *reinterpret_cast<const float4*>(X + l * ldx + k4 * 4);   // broadcast across the warp
          acc[l] = fmaf(x.x, w.x, acc[l]); acc[l] = fmaf(x.y, w.y, acc[l]);
          acc[l] = fmaf(x.z, w.z, acc[l]); acc[l] = fmaf(x.w, w.w, acc[l]);
        }
    }
    const float bj = bias ? __ldg(bias + j) : 0.f;
    const float sc = (j < n_scaled) ? scale_first : 1.f;
#pragma unroll
    for (int l = 0; l < MAXL; ++l)
      if (l < L) Y[l * ldy + j] = (acc[l] + bj) * sc;
  }
}

__global__ void __launch_bounds__(kPatchThreads) patch_attention_kernel(const PatchAttnParams p) {
  extern __shared__ float smem[];
  const int L = p.L, E = p.E, H = p.heads, hd = E / H;
  float* X = smem;                       // [L][E]   tokens in / attention output / next tokens
  float* QKV = X + L * E;                // [L][3E]
  float* P = QKV + L * 3 * E;            // [H][L][L]
  float* O = P + H * L * L;              // [L][E]
  const float qscale = sqrtf(1.0f / (float)hd);      // torch: q * math.sqrt(1.0 / head_dim)

  for (int b = blockIdx.x; b < p.B; b += gridDim.x) {
    __syncthreads();
    for (int i = threadIdx.x; i < L * E; i += blockDim.x) {
      const int l = i / E, e = i - l * E;
      X[i] = p.feat[((long long)l * p.B + b) * E + e];
    }
    __syncthreads();
    for (int a = 0; a < 2; ++a) {
      rows_linear<kPatchMaxLayers>(X, E, p.w_in[a], p.b_in[a], QKV, 3 * E, L, 3 * E, E, qscale, E);
      __syncthreads();
      for (int i = threadIdx.x; i < H * L * L; i += blockDim.x) {       // scores
        const int h = i / (L * L), r = i - h * L * L, l = r / L, m = r - l * L;
        const float* q = QKV + l * 3 * E + h * hd;
        const float* k = QKV + m * 3 * E + E + h * hd;
        float s = 0.f;
        for (int d = 0; d < hd; ++d) s = fmaf(q[d], k[d], s);
        P[i] = s;
      }
      __syncthreads();
      for (int i = threadIdx.x; i < H * L; i += blockDim.x) {           // softmax over keys
        float* row = P + i * L;
        float mx = row[0];
        for (int m = 1; m < L; ++m) mx = fmaxf(mx, row[m]);
        float sum = 0.f;
        for (int m = 0; m < L; ++m) { const float e = expf(row[m] - mx); row[m] = e; sum += e; }
        const float r = 1.f / sum;
        for (int m = 0; m < L; ++m) row[m] *= r;
      }
      __syncthreads();
      for (int i = threadIdx.x; i < L * E; i += blockDim.x) {           // attention-weighted values
        const int l = i / E, e = i - l * E, h = e / hd;
        const float* pr = P + (h * L + l) * L;
        float s = 0.f;
        for (int m = 0; m < L; ++m) s = fmaf(pr[m], QKV[m * 3 * E + 2 * E + e], s);
        O[i] = s;
      }
      __syncthreads();
      rows_linear<kPatchMaxLayers>(O, E, p.w_out[a], p.b_out[a], X, E, L, E, E, 1.f, 0);
      __syncthreads();
    }
    for (int e = threadIdx.x; e < E; e += blockDim.x) {                 // mean over the layer tokens (:247)
      float s = 0.f;
      for (int l = 0; l < L; ++l) s += X[l * E + e];
      s /= (float)L;
      O[e] = s;
      p.emb[(long long)b * E + e] = s;
    }
    __syncthreads();
    for (int j = threadIdx.x >> 5; j < p.nc; j += blockDim.x >> 5) {    // classifier (:256): one warp per class
      const float* w = p.w_c + (long long)j * E;
      float s = 0.f;
      for (int e = threadIdx.x & 31; e < E; e += 32) s = fmaf(O[e], __ldg(w + e), s);
      s = warp_sum(s);
      if ((threadIdx.x & 31) == 0) p.logits[(long long)b * p.nc + j] = s + __ldg(p.b_c + j);
    }
  }
}

}  // namespace gh
