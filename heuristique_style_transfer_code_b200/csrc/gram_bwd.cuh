// Gram backward on tcgen05:  dF_b = (dG_b + dG_b^T) F_b * scale,   F_b, dF_b : C x HW.
//
// This is what autograd derives for the reference's bmm(features, features^T).div(h*w)
// (Models/Models_RESNET50_TRUNCATE_GRAM_with_Attention.py:26-30) followed, on the classification path, by
// adaptive_avg_pool2d (:51-52). In POOL mode dG is never materialised: it is the block-constant upsampling of the
// g x g descriptor gradient, dG[c][d] = dP[c/k][d/k] / k^2, and the symmetrised bf16 operand tile is generated in
// shared memory from the 4 KB dP of the image. In DENSE mode (gram_matrix() used directly, e.g. style transfer,
// functions/functions_RESNET50_Truncate_Gram_Attention.py:291-296) dG is read from HBM.
//
// The product is computed transposed so the result leaves TMEM already coalesced along HW:
//   D^T[hw][c] = sum_d  F^T[hw][d] * M[c][d],    M = dG + dG^T (symmetric)
//   A tile = F^T  [128 hw][64 d]  (producers gather 8 channel rows per 16 B chunk; loads coalesced along hw)
//   B tile = M    [256 c ][64 d]  (x NB blocks of 256 output channels)
//   unit   = (image b, 128-wide hw tile, block of 256*NB output channels); K loop over all C input channels
//   TMEM   = 512 columns: two 256-column accumulator buffers when NB == 1 (epilogue of unit i overlaps MMAs of
//            unit i+1), one 512-column buffer when NB == 2.
#pragma once
#include "common.cuh"

namespace gh {

constexpr int kGbProducerWarps = 8;
constexpr int kGbProducerThreads = 256;
constexpr int kGbEpiWarp0 = 8;
constexpr int kGbMmaWarp = 12;
constexpr int kGbThreads = 13 * 32;
constexpr uint32_t kGbATileBytes = 128 * kRowBytes;   // 16 KB
constexpr uint32_t kGbBBlkBytes = 256 * kRowBytes;    // 32 KB
constexpr uint32_t kGbRingBytes = 192 * 1024;
constexpr int kGbMaxStages = 4;
constexpr int kGbMaxG = 64;
constexpr uint32_t kGbSmemBytes = kGbRingBytes + kGbMaxG * kGbMaxG * 4 + 1024 + 256;

struct GramBwdParams {
  const void* F;
  long long img_stride, row_stride;
  int B, C, HW;
  int mode;                 // GRAM_POOL / GRAM_DENSE
  const float* dP;          // POOL: (B, L, g*g) slice base of this stage
  long long dp_img_stride;
  int g, kshift;            // pool factor k = 1 << kshift
  const float* dG;          // DENSE: (B, C, C)
  float* dF;
  long long df_img_stride, df_row_stride;
  float scale;
  int nHT, nCB, NB, nkb;
  int stages;
  uint32_t stage_bytes;
  int nacc;
  int total_units;
};

struct GramBwdUnit {
  int b, ht, cb;
};
__device__ __forceinline__ GramBwdUnit gram_bwd_decode(const GramBwdParams& p, int u) {
  GramBwdUnit w;
  w.cb = u % p.nCB;
  const int t = u / p.nCB;
  w.ht = t % p.nHT;
  w.b = t / p.nHT;
  return w;
}

__device__ __forceinline__ void named_bar_sync(uint32_t id, uint32_t nthreads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

// SRC: 0 = fp32 features, 2 = bf16 features.
template <int SRC>
__global__ void __launch_bounds__(kGbThreads, 1) gram_bwd_kernel(const GramBwdParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t sym_smem = smem_base + kGbRingBytes;
  float* sym = reinterpret_cast<float*>(smem_raw + (sym_smem - smem_u32(smem_raw)));
  const uint32_t bars = sym_smem + kGbMaxG * kGbMaxG * 4;
  const uint32_t bar_full = bars, bar_empty = bars + 8 * kGbMaxStages;
  const uint32_t bar_tfull = bars + 16 * kGbMaxStages, bar_tempty = bar_tfull + 16;
  const uint32_t tmem_slot = bar_tempty + 16;
  volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(smem_raw + (tmem_slot - smem_u32(smem_raw)));

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    for (int s = 0; s < kGbMaxStages; ++s) {
      mbar_init(bar_full + 8 * s, kGbProducerWarps);
      mbar_init(bar_empty + 8 * s, 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(bar_tfull + 8 * a, 1);
      mbar_init(bar_tempty + 8 * a, 4);
    }
    mbar_fence_init();
  }
  if (warp == kGbMmaWarp) tmem_alloc(tmem_slot, 512);
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem_base = *tmem_slot_ptr;
  const int nstages = p.stages;

  if (warp < kGbProducerWarps) {
    // =========================== producers ===========================
    const int tid = threadIdx.x;
    uint32_t stage = 0, phase = 0;
    int cur_b = -1;
    for (int u = blockIdx.x; u < p.total_units; u += gridDim.x) {
      const GramBwdUnit w = gram_bwd_decode(p, u);
      if (p.mode == GRAM_POOL && w.b != cur_b) {
        // sym[i][j] = dP[i][j] + dP[j][i] for this image
        named_bar_sync(1, kGbProducerThreads);   // every producer is done reading the previous image's table
        const float* dp = p.dP + (long long)w.b * p.dp_img_stride;
        const int gg = p.g * p.g;
        for (int i = tid; i < gg; i += kGbProducerThreads) {
          const int r = i / p.g, c = i - r * p.g;
          sym[i] = __ldg(dp + i) + __ldg(dp + c * p.g + r);
        }
        named_bar_sync(1, kGbProducerThreads);
        cur_b = w.b;
      }
      const int hw0 = w.ht * 128;
      for (int kb = 0; kb < p.nkb; ++kb) {
        mbar_wait(bar_empty + 8 * stage, phase ^ 1u, 100u + stage);
        const uint32_t a_smem = smem_base + stage * p.stage_bytes;
        const uint32_t b_smem = a_smem + kGbATileBytes;
        // ---- A tile: F^T [128 hw][64 d]; thread -> hw row (tid & 127), 4 of the 8 d-chunks
        {
          const int row = tid & 127;
          const int hw = hw0 + row;
          const bool hv = hw < p.HW;
          const int jbase = (tid >> 7) * 4;
          if (SRC == 0) {
            const float* fb = reinterpret_cast<const float*>(p.F) + (long long)w.b * p.img_stride + hw;
            float x[32];
#pragma unroll
            for (int jj = 0; jj < 4; ++jj)
#pragma unroll
              for (int e = 0; e < 8; ++e) {
                const int d = kb * 64 + (jbase + jj) * 8 + e;
                x[jj * 8 + e] = (hv && d < p.C) ? __ldg(fb + (long long)d * p.row_stride) : 0.f;
              }
#pragma unroll
            for (int jj = 0; jj < 4; ++jj) {
              const uint32_t off = sw128_off((uint32_t)row, (uint32_t)((jbase + jj) * 8));
              sts_u4(a_smem + off, pack_bf16x2(x[jj * 8 + 0], x[jj * 8 + 1]), pack_bf16x2(x[jj * 8 + 2], x[jj * 8 + 3]),
                     pack_bf16x2(x[jj * 8 + 4], x[jj * 8 + 5]), pack_bf16x2(x[jj * 8 + 6], x[jj * 8 + 7]));
            }
          } else {
            const unsigned short* fb = reinterpret_cast<const unsigned short*>(p.F) + (long long)w.b * p.img_stride + hw;
            unsigned int x[32];
#pragma unroll
            for (int jj = 0; jj < 4; ++jj)
#pragma unroll
              for (int e = 0; e < 8; ++e) {
                const int d = kb * 64 + (jbase + jj) * 8 + e;
                x[jj * 8 + e] = (hv && d < p.C) ? (unsigned int)__ldg(fb + (long long)d * p.row_stride) : 0u;
              }
#pragma unroll
            for (int jj = 0; jj < 4; ++jj) {
              const uint32_t off = sw128_off((uint32_t)row, (uint32_t)((jbase + jj) * 8));
              sts_u4(a_smem + off, x[jj * 8 + 0] | (x[jj * 8 + 1] << 16), x[jj * 8 + 2] | (x[jj * 8 + 3] << 16),
                     x[jj * 8 + 4] | (x[jj * 8 + 5] << 16), x[jj * 8 + 6] | (x[jj * 8 + 7] << 16));
            }
          }
        }
        // ---- B tile(s): M [256*NB c][64 d]; thread -> row n = tid + 256*i, all 8 chunks
        for (int i = 0; i < p.NB; ++i) {
          const int n = tid + 256 * i;
          const int c = w.cb * 256 * p.NB + n;
          const uint32_t blk = b_smem + (uint32_t)i * kGbBBlkBytes;
          if (p.mode == GRAM_POOL) {
            const float* srow = sym + (c >> p.kshift) * p.g;
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              const int d0 = kb * 64 + j * 8;
              uint32_t pk[4];
              if (c >= p.C) {
                pk[0] = pk[1] = pk[2] = pk[3] = 0u;
              } else if (p.kshift >= 3) {
                const float v = (d0 < p.C) ? srow[d0 >> p.kshift] : 0.f;
                pk[0] = pk[1] = pk[2] = pk[3] = pack_bf16x2(v, v);
              } else {
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                  const int da = d0 + 2 * e, db = da + 1;
                  pk[e] = pack_bf16x2(da < p.C ? srow[da >> p.kshift] : 0.f, db < p.C ? srow[db >> p.kshift] : 0.f);
                }
              }
              sts_u4(blk + sw128_off((uint32_t)(tid), (uint32_t)(j * 8)), pk[0], pk[1], pk[2], pk[3]);
            }
          } else {
            const float* gb = p.dG + (long long)w.b * p.C * p.C;
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              float x[8];
#pragma unroll
              for (int e = 0; e < 8; ++e) {
                const int d = kb * 64 + j * 8 + e;
                x[e] = (c < p.C && d < p.C) ? (__ldg(gb + (long long)c * p.C + d) + __ldg(gb + (long long)d * p.C + c)) : 0.f;
              }
              sts_u4(blk + sw128_off((uint32_t)(tid), (uint32_t)(j * 8)), pack_bf16x2(x[0], x[1]), pack_bf16x2(x[2], x[3]),
                     pack_bf16x2(x[4], x[5]), pack_bf16x2(x[6], x[7]));
            }
          }
        }
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) mbar_arrive(bar_full + 8 * stage);
        if ((int)++stage == nstages) { stage = 0; phase ^= 1u; }
      }
    }
  } else if (warp == kGbMmaWarp) {
    // =========================== MMA issuer ===========================
    uint32_t stage = 0, phase = 0, it = 0;
    const uint32_t idesc = make_idesc_bf16(128, 256);
    for (int u = blockIdx.x; u < p.total_units; u += gridDim.x, ++it) {
      const uint32_t ab = (p.nacc == 2) ? (it & 1u) : 0u;
      const uint32_t use = (p.nacc == 2) ? (it >> 1) : it;        // how many times this buffer was used before
      mbar_wait(bar_tempty + 8 * ab, (use & 1u) ^ 1u, 200u + ab);
      tc_fence_after_sync();
      const uint32_t acc_col = ab * 256u;
      for (int kb = 0; kb < p.nkb; ++kb) {
        mbar_wait(bar_full + 8 * stage, phase, 300u + stage);
        tc_fence_after_sync();
        if (elect_one()) {   // elect.sync, not a lane test: see gram_fwd_pair.cuh
          const uint32_t a_smem = smem_base + stage * p.stage_bytes;
          const uint32_t b_smem = a_smem + kGbATileBytes;
#pragma unroll
          for (uint32_t ks = 0; ks < kTileK / kUmmaK; ++ks) {
            if ((int)(kb * 64 + ks * 16) < p.C) {
              const uint32_t koff = ks * 32u;
              for (int i = 0; i < p.NB; ++i)
                umma_bf16(tmem_base + acc_col + 256u * i, make_smem_desc_sw128(a_smem + koff),
                          make_smem_desc_sw128(b_smem + (uint32_t)i * kGbBBlkBytes + koff), idesc, (uint32_t)kb | ks);
            }
          }
          umma_commit(bar_empty + 8 * stage);
          if (kb + 1 == p.nkb) umma_commit(bar_tfull + 8 * ab);
        }
        __syncwarp();
        if ((int)++stage == nstages) { stage = 0; phase ^= 1u; }
      }
    }
  } else {
    // =========================== epilogue ===========================
    const int q = warp - kGbEpiWarp0;
    uint32_t it = 0;
    for (int u = blockIdx.x; u < p.total_units; u += gridDim.x, ++it) {
      const GramBwdUnit w = gram_bwd_decode(p, u);
      const uint32_t ab = (p.nacc == 2) ? (it & 1u) : 0u;
      const uint32_t use = (p.nacc == 2) ? (it >> 1) : it;
      mbar_wait(bar_tfull + 8 * ab, use & 1u, 400u + ab);
      tc_fence_after_sync();
      const int hw = w.ht * 128 + q * 32 + lane;
      const bool hv = hw < p.HW;
      float* ob = p.dF + (long long)w.b * p.df_img_stride + hw;
      const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + ab * 256u;
      const int ncols = 256 * p.NB;
#pragma unroll 1
      for (int n0 = 0; n0 < ncols; n0 += 32) {
        const int c0 = w.cb * ncols + n0;
        if (c0 >= p.C) break;   // warp-uniform
        float v[32];
        tmem_ld32(taddr + (uint32_t)n0, v);
        if (hv) {
#pragma unroll
          for (int j = 0; j < 32; ++j)
            if (c0 + j < p.C) ob[(long long)(c0 + j) * p.df_row_stride] = v[j] * p.scale;   // 128 B per warp per j
        }
      }
      tc_fence_before_sync();
      __syncwarp();
      if (lane == 0) mbar_arrive(bar_tempty + 8 * ab);
    }
  }

  tc_fence_before_sync();
  __syncthreads();
  if (warp == kGbMmaWarp) {
    tc_fence_after_sync();
    tmem_dealloc(tmem_base, 512);
  }
}

}  // namespace gh
