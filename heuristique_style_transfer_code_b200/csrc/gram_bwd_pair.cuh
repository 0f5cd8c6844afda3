// Gram backward on CTA pairs:  dF_b[c][x] = scale * sum_d M_b[c][d] F_b[d][x],   M = dG + dG^T  (C x C, symmetric).
//
// Autograd of reference Models/Models_RESNET50_TRUNCATE_GRAM_with_Attention.py:26-30 (bmm + div) and :51-52
// (adaptive_avg_pool2d), driven by loss.backward() at functions/functions_RESNET50_Truncate_Gram_Attention.py:135.
//
// One cluster of two CTAs (one TPC) produces a 256-channel x NT-position tile of dF per unit with
// tcgen05.mma.cta_group::2 (M = 256, N = NT <= 256):
//   A = M   [256 c][K d]   K-major SWIZZLE_128B; each CTA GENERATES its 128 rows in shared memory from the image's
//                          g x g descriptor gradient (POOL: M[c][d] = dP[c/k][d/k] + dP[d/k][c/k], block-constant) or
//                          from the dense dG (DENSE), so dG never exists in HBM on the classification path
//   B = F   [K d][NT x]    MN-major SWIZZLE_128B: rows of F are contiguous along x, which is the canonical MN-major
//                          layout, so the feature map is TMA-staged exactly as cuDNN left it (fp32 -> kind::tf32,
//                          bf16 -> kind::f16): no cast pass, no register staging; each CTA loads its NT/2 columns
//   D = dF  [128 c][NT x]  per CTA in TMEM, two accumulator buffers: the epilogue of unit i overlaps the MMAs of i+1
//   epilogue: tcgen05.ld (lane = channel, 32 positions) -> *scale -> swizzled staging tile -> TMA store, so the
//             4 B/element gradient leaves the SM as full 128 B lines without occupying the LSU
// Warps: 0 = TMA producer, 1 = TMEM owner + MMA issuer (leader CTA only), 2-5 = epilogue, 6-21 = A-tile generators
// in four groups of four warps; group i owns ring stage i, so four K chunks are generated concurrently and the
// per-chunk handshake (mbarrier wait, proxy fence, arrive) of one group overlaps the stores of the others.
#pragma once
#include "pair.cuh"
#include "gram_fwd.cuh"   // GramMode
#include "gram_bwd.cuh"   // named_bar_sync

namespace gh {

constexpr int kBpStages = 4;
constexpr uint32_t kBpTileBytes = 16384;                    // A: [128 c][128 B]; B: <= 128 x-columns x (K chunk) x elem
constexpr uint32_t kBpStageBytes = 2 * kBpTileBytes;
constexpr int kBpStoreBufs = 4;                             // staging tiles per epilogue warp = TMA stores it keeps in flight
constexpr uint32_t kBpStoreBytes = 4 * kBpStoreBufs * 4096; // per epilogue warp: kBpStoreBufs [32 c][32 x] fp32 staging tiles
constexpr int kBpMaxG = 32;                                 // pooled size handled here (larger g: the ldg kernels)
constexpr int kBpGroups = kBpStages;                        // generator groups; group i fills ring stage i
constexpr int kBpTableFloats = kBpMaxG * kBpMaxG + kBpMaxG; // g x g table + one row of zeros (rows beyond C)
constexpr uint32_t kBpSymBytes = kBpGroups * 2 * kBpTableFloats * 4;   // per group: current and next image's table
constexpr uint32_t kBpSmemBytes = kBpStages * kBpStageBytes + kBpStoreBytes + kBpSymBytes + 1024 + 256;
constexpr int kBpGroupThreads = 128;                        // one thread per A row
constexpr int kBpThreads = 6 * 32 + kBpGroups * kBpGroupThreads;

struct GramBwdPairParams {
  int B, C, HW;
  int mode;                 // GRAM_POOL / GRAM_DENSE
  const float* dP;          // POOL: (B, L, g*g) slice base of this stage
  long long dp_img_stride;
  int g, kshift;
  const float* dG;          // DENSE: (B, C, C)
  float scale;
  int NT;                   // x-tile width of a pair (multiple of 32, <= 256); each CTA supplies NT/2 columns of B
  int nHT, nCB;             // x tiles, 256-channel output blocks
  int nkc;                  // K chunks (32 input channels for tf32, 64 for bf16)
  int total_units;
};

struct GramBwdPairUnit {
  int b, ht, cb;
};
__device__ __forceinline__ GramBwdPairUnit gbp_decode(const GramBwdPairParams& p, int u) {
  GramBwdPairUnit w;
  w.cb = u % p.nCB;
  const int t = u / p.nCB;
  w.ht = t % p.nHT;
  w.b = t / p.nHT;
  return w;
}

__device__ __forceinline__ uint32_t f32_to_tf32_rna(float x) {
  uint32_t r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
  return r;
}

// MODE: GRAM_POOL (A generated from the g x g descriptor gradient) or GRAM_DENSE (A = dG + dG^T read from HBM).
template <int KIND, int MODE>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kBpThreads, 1)
    gram_bwd_pair_kernel(const GramBwdPairParams p, const __grid_constant__ CUtensorMap tmF,
                         const __grid_constant__ CUtensorMap tmD) {
  using T = KindTraits<KIND>;
  constexpr uint32_t KC = T::kElemsPerRow;                  // input channels per K chunk
  constexpr uint32_t kAtomBytesB = KC * kRowBytes;          // one 128 B-wide x block of the B tile: 4 KB | 8 KB
  constexpr uint32_t kStepBytesB = (T::kUmmaK / 8) * 1024;  // B advance per MMA: K/8 groups of 8 K-rows

  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t store_smem = smem_base + kBpStages * kBpStageBytes;
  const uint32_t sym_smem = store_smem + kBpStoreBytes;
  float* sym = reinterpret_cast<float*>(smem_raw + (sym_smem - smem_u32(smem_raw)));
  const uint32_t bars = sym_smem + kBpSymBytes;
  const uint32_t bar_full = bars, bar_empty = bars + 8 * kBpStages;
  const uint32_t bar_tfull = bars + 16 * kBpStages, bar_tempty = bar_tfull + 16;
  const uint32_t tmem_slot = bar_tempty + 16;
  volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(smem_raw + (tmem_slot - smem_u32(smem_raw)));

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const int pair = (int)cluster_id_x(), npairs = (int)cluster_nclusters_x();

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tmF);
    tma_prefetch_desc(&tmD);
    for (int s = 0; s < kBpStages; ++s) {
      mbar_init(bar_full + 8 * s, 1 + 2 * (kBpGroupThreads / 32)); // leader's expect_tx + the stage's generator warps of both CTAs
      mbar_init(bar_empty + 8 * s, 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(bar_tfull + 8 * a, 1);
      mbar_init(bar_tempty + 8 * a, 8);                            // 4 epilogue warps x 2 CTAs
    }
    mbar_fence_init();
  }
  if (warp == 1) tmem_alloc2(tmem_slot, 512);
  tc_fence_before_sync();
  cluster_sync_all();
  tc_fence_after_sync();
  const uint32_t tmem_base = *tmem_slot_ptr;

  const uint32_t full_leader = mapa_u32(bar_full, 0);      // shared::cluster addresses of the leader's barriers
  const uint32_t tempty_leader = mapa_u32(bar_tempty, 0);
  const int na = (p.NT / 2 + (int)KC - 1) / (int)KC;       // 128 B-wide x blocks each CTA loads per K chunk

  if (warp == 0) {
    // =========================== TMA producer: this CTA's NT/2 columns of F ===========================
    uint32_t stage = 0, phase = 0;
    for (int u = pair; u < p.total_units; u += npairs) {
      const GramBwdPairUnit w = gbp_decode(p, u);
      const int x0 = w.ht * p.NT + (int)rank * (p.NT / 2);
      for (int kc = 0; kc < p.nkc; ++kc) {
        mbar_wait(bar_empty + 8 * stage, phase ^ 1u, 100u + stage);
        if (lane == 0) {
          if (rank == 0) mbar_arrive_expect_tx(bar_full + 8 * stage, 2u * (uint32_t)na * kAtomBytesB);
          const uint32_t b_tile = smem_base + stage * kBpStageBytes + kBpTileBytes;
          for (int j = 0; j < na; ++j)
            tma_load_3d_pair(b_tile + (uint32_t)j * kAtomBytesB, &tmF, full_leader + 8 * stage, x0 + j * (int)KC,
                             kc * (int)KC, w.b);
        }
        __syncwarp();
        if (++stage == kBpStages) { stage = 0; phase ^= 1u; }
      }
    }
  } else if (warp == 1) {
    // =========================== MMA issuer (leader CTA) ===========================
    if (rank == 0) {
      const uint32_t idesc = make_idesc(T::kFormat, 256, (uint32_t)p.NT, 0, 1);
      uint32_t stage = 0, phase = 0, it = 0;
      for (int u = pair; u < p.total_units; u += npairs, ++it) {
        const uint32_t ab = it & 1u, use = it >> 1;
        mbar_wait_cl(bar_tempty + 8 * ab, (use & 1u) ^ 1u, 200u + ab);
        tc_fence_after_sync();
        for (int kc = 0; kc < p.nkc; ++kc) {
          mbar_wait_cl(bar_full + 8 * stage, phase, 300u + stage);
          tc_fence_after_sync();
          if (lane == 0) {
            const uint32_t a_tile = smem_base + stage * kBpStageBytes;
            const uint32_t b_tile = a_tile + kBpTileBytes;
#pragma unroll
            for (uint32_t ks = 0; ks < KC / T::kUmmaK; ++ks) {
              if ((int)(kc * KC + ks * T::kUmmaK) >= p.C) break;
              const uint64_t db = (KIND == KIND_TF32) ? make_smem_desc_sw128b32_mnmajor(b_tile + ks * kStepBytesB, kAtomBytesB)
                                                      : make_smem_desc_sw128_mnmajor(b_tile + ks * kStepBytesB, kAtomBytesB);
              umma2<KIND>(tmem_base + ab * 256u, make_smem_desc_sw128(a_tile + ks * 32u), db, idesc, (uint32_t)kc | ks);
            }
            umma_commit2(bar_empty + 8 * stage);
            if (kc + 1 == p.nkc) umma_commit2(bar_tfull + 8 * ab);
          }
          __syncwarp();
          if (++stage == kBpStages) { stage = 0; phase ^= 1u; }
        }
      }
    }
  } else if (warp < 6) {
    // =========================== epilogue: TMEM -> scale -> staging -> TMA store ===========================
    const int q = warp & 3;                                 // TMEM lane quarter this warp may read
    const uint32_t my_store = store_smem + (uint32_t)(warp - 2) * (kBpStoreBufs * 4096u);
    uint32_t it = 0, buf = 0;
    for (int u = pair; u < p.total_units; u += npairs, ++it) {
      const GramBwdPairUnit w = gbp_decode(p, u);
      const uint32_t ab = it & 1u, use = it >> 1;
      mbar_wait(bar_tfull + 8 * ab, use & 1u, 400u + ab);
      tc_fence_after_sync();
      const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + ab * 256u;
      const int crow0 = w.cb * 256 + (int)rank * 128 + q * 32;
#pragma unroll 1
      for (int n0 = 0; n0 < p.NT; n0 += 32) {
        const int x = w.ht * p.NT + n0;
        if (x >= p.HW) break;                               // warp-uniform
        float v[32];
        tmem_ld32(taddr + (uint32_t)n0, v);
        if (lane == 0) tma_store_wait_read<kBpStoreBufs - 1>();   // the staging tile about to be reused has been read
        __syncwarp();
        const uint32_t tile = my_store + buf * 4096u + (uint32_t)lane * kRowBytes;
#pragma unroll
        for (int j = 0; j < 8; ++j)
          sts_u4(tile + ((((uint32_t)j) ^ ((uint32_t)lane & 7u)) << 4), __float_as_uint(v[4 * j] * p.scale),
                 __float_as_uint(v[4 * j + 1] * p.scale), __float_as_uint(v[4 * j + 2] * p.scale),
                 __float_as_uint(v[4 * j + 3] * p.scale));
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0 && crow0 < p.C) {
          tma_store_3d(&tmD, my_store + buf * 4096u, x, crow0, w.b);
          tma_store_commit();
        }
        buf = (buf + 1u) % kBpStoreBufs;
      }
      tc_fence_before_sync();
      __syncwarp();
      if (lane == 0) mbar_arrive_cluster(tempty_leader + 8 * ab);
    }
    if (lane == 0) tma_store_wait_all<0>();
    __syncwarp();
  } else {
    // =========================== A-tile generators: this CTA's 128 rows of M ===========================
    // Group grp fills stage grp, i.e. the K chunks n = unit_index * nkc + kc with n % 4 == grp. Thread = A row.
    // POOL: every 16 B chunk of the row is one value of the per-image table sym = dP + dP^T repeated (the pooling
    // factor k is >= the elements of a chunk): 8 shared loads, 8 conversions and 8 16 B stores per K chunk. Each
    // group keeps its own copy of the table; the NEXT unit's is fetched into registers while the current unit is
    // generated and published through the group's second buffer.
    constexpr int EPC = 16 / (int)T::kElemBytes;            // elements per 16 B chunk: 4 (tf32) / 8 (bf16)
    constexpr int NE = kBpMaxG * kBpMaxG / kBpGroupThreads; // table entries per thread
    const int grp = (warp - 6) >> 2;
    const int gt = threadIdx.x - 6 * 32 - grp * kBpGroupThreads;   // 0..127 = A row
    const uint32_t row = (uint32_t)gt, sw = row & 7u;
    const uint32_t row_off = (row >> 3) * kAtomBytes + sw * kRowBytes;
    const int gg = p.g * p.g;
    float* table = sym + grp * 2 * kBpTableFloats;
    float pre[NE];
    auto fetch_table = [&](int b) {
      const float* dp = p.dP + (long long)b * p.dp_img_stride;
#pragma unroll
      for (int e = 0; e < NE; ++e) {
        const int i = gt + e * kBpGroupThreads;
        pre[e] = 0.f;
        if (i < gg) {
          const int r = i / p.g, cc = i - r * p.g;
          pre[e] = __ldg(dp + i) + __ldg(dp + cc * p.g + r);
        }
      }
    };
    auto publish_table = [&](int which) {
#pragma unroll
      for (int e = 0; e < NE; ++e) {
        const int i = gt + e * kBpGroupThreads;
        if (i < gg) table[which * kBpTableFloats + i] = pre[e];
      }
    };
    uint32_t phase = 0;
    int cur = 0;
    const uint32_t bar_id = 1u + (uint32_t)grp;
    if (MODE == GRAM_POOL) {
      if (gt < kBpMaxG) {                                   // the zero rows
        table[gg + gt] = 0.f;
        table[kBpTableFloats + gg + gt] = 0.f;
      }
      if (pair < p.total_units) {
        fetch_table(gbp_decode(p, pair).b);
        publish_table(0);
      }
      named_bar_sync(bar_id, kBpGroupThreads);
    }
    int n0 = 0;                                             // sequence number of the unit's first K chunk
    for (int u = pair; u < p.total_units; u += npairs, n0 += p.nkc) {
      const GramBwdPairUnit w = gbp_decode(p, u);
      const bool has_next = u + npairs < p.total_units;
      if (MODE == GRAM_POOL && has_next) fetch_table(gbp_decode(p, u + npairs).b);
      const int c = w.cb * 256 + (int)rank * 128 + (int)row;
      const bool row_ok = c < p.C;
      const float* srow = table + cur * kBpTableFloats + (row_ok ? (c >> p.kshift) * p.g : gg);
      for (int kc = (grp - (n0 & 3) + 4) & 3; kc < p.nkc; kc += kBpGroups) {
        mbar_wait(bar_empty + 8 * grp, phase ^ 1u, 500u + grp);
        const uint32_t a_row = smem_base + (uint32_t)grp * kBpStageBytes + row_off;
        const int dbase = kc * (int)KC;
        if (MODE == GRAM_POOL) {
          const bool full_chunk = dbase + (int)KC <= p.C;
          float v[8];
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const int d0 = dbase + j * EPC;
            v[j] = (full_chunk || d0 < p.C) ? srow[d0 >> p.kshift] : 0.f;
          }
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const uint32_t t = (KIND == KIND_TF32) ? f32_to_tf32_rna(v[j]) : pack_bf16x2(v[j], v[j]);
            sts_u4(a_row + ((((uint32_t)j) ^ sw) << 4), t, t, t, t);
          }
        } else {
          const float* gb = p.dG + (long long)w.b * p.C * p.C;
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const int d0 = dbase + j * EPC;
            float x[EPC];
#pragma unroll
            for (int e = 0; e < EPC; ++e) {
              const int d = d0 + e;
              x[e] = (row_ok && d < p.C) ? (__ldg(gb + (long long)c * p.C + d) + __ldg(gb + (long long)d * p.C + c)) : 0.f;
            }
            const uint32_t addr = a_row + ((((uint32_t)j) ^ sw) << 4);
            if constexpr (KIND == KIND_TF32) {
              sts_u4(addr, f32_to_tf32_rna(x[0]), f32_to_tf32_rna(x[1]), f32_to_tf32_rna(x[2]), f32_to_tf32_rna(x[3]));
            } else {
              sts_u4(addr, pack_bf16x2(x[0], x[1]), pack_bf16x2(x[2], x[3]), pack_bf16x2(x[4], x[5]), pack_bf16x2(x[6], x[7]));
            }
          }
        }
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) mbar_arrive_cluster(full_leader + 8 * grp);
        phase ^= 1u;
      }
      if (MODE == GRAM_POOL) {
        if (has_next) publish_table(cur ^ 1);               // last read one unit ago, before the barrier that ended it
        named_bar_sync(bar_id, kBpGroupThreads);            // the group sees the next table and is done with this one
        cur ^= 1;
      }
    }
  }

  __syncwarp();
  tc_fence_before_sync();
  cluster_sync_all();
  if (warp == 1) {
    tc_fence_after_sync();
    tmem_dealloc2(tmem_base, 512);
  }
}

// x-tile plan: nHT tiles of NT (multiple of 32, 64 <= NT <= 256) covering HW with the least padded MMA work.
inline void gbp_plan_tiles(int HW, int* NT, int* nHT) {
  int best_nt = 256, best_n = (HW + 255) / 256;
  long long best_cost = (long long)best_nt * best_n;
  const int n0 = (HW + 255) / 256;
  for (int n = n0; n <= n0 + 3; ++n) {
    int nt = ((HW + n - 1) / n + 31) / 32 * 32;
    if (nt < 64) nt = 64;
    if (nt > 256) continue;
    const long long cost = (long long)nt * n;
    if (cost < best_cost) { best_cost = cost; best_nt = nt; best_n = n; }
  }
  *NT = best_nt;
  *nHT = best_n;
}

}  // namespace gh
