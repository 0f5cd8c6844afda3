// Batched symmetric Gram forward on CTA pairs:  G_b = F_b F_b^T * scale  with TMA-staged operands.
//
// Same contract as gram_fwd.cuh (reference Models/Models_RESNET50_TRUNCATE_GRAM_with_Attention.py:26-30 and, in POOL
// mode, :51-55), different machine mapping:
//   unit      = (image b, 256x256 super-tile I <= J of the Gram, K-range kp of ksplit), walked by a CLUSTER of two CTAs
//   MMA       = tcgen05.mma.cta_group::2, M = 256 (rows of block I, 128 per CTA), N = 256 (rows of block J, 128 per CTA):
//               each CTA holds a 128 x 256 fp32 accumulator (256 TMEM columns), so TWO accumulator buffers fit and the
//               pooled epilogue of unit i overlaps the MMAs of unit i+1
//   operands  = TMA (cp.async.bulk.tensor) straight from the NCHW feature map into K-major SWIZZLE_128B tiles, one
//               128-row box per operand half per CTA; bf16 features feed kind::f16, fp32 features feed kind::tf32 as
//               they are (no cast pass, no register staging); a diagonal super-tile loads ONE box per CTA and uses it as
//               both its A half and its B half
//   ring      = 6 stages x (A 16 KB + B 16 KB) per CTA, up to 192 KB of loads in flight per SM
//   epilogue  = the k x k pooling / mirrored dense store of gram_fwd.cuh on this CTA's 128 rows; the redundant
//               lower-left 128 x 128 block of a diagonal super-tile is computed (M is indivisible) but never read
// Warps: 0 = TMA producer, 1 = TMEM owner + MMA issuer (leader CTA only), 2-5 = epilogue.
#pragma once
#include "pair.cuh"
#include "gram_fwd.cuh"

namespace gh {

constexpr int kFpStages = 6;
constexpr uint32_t kFpTileBytes = 16384;                       // [128 rows][128 B]
constexpr uint32_t kFpStageBytes = 2 * kFpTileBytes;
constexpr uint32_t kFpSmemBytes = kFpStages * kFpStageBytes + 1024 + 256;
constexpr int kFpThreads = 6 * 32;

// Epilogue of one unit on this CTA's accumulator (rows I*256 + rank*128 .. +127, columns J*256 .. +255).
template <int KP>
__device__ __forceinline__ void gfp_epilogue_unit(const GramFwdParams& p, const GramUnit& w, uint32_t tmem_acc, int rank,
                                                  int q, bool atomics, int lane) {
  const int row0 = w.I * 256 + rank * 128;
  if (row0 >= p.C) return;                                     // this CTA's rows are all padding
  const int c_row = row0 + q * 32 + lane;
  float* outp = p.out + (long long)w.b * p.out_img_stride;
  const uint32_t lane_addr = tmem_acc + ((uint32_t)(q * 32) << 16);
#pragma unroll 1
  for (int g0 = 0; g0 < 256; g0 += 128) {
    const int c_grp = w.J * 256 + g0;
    if (c_grp >= p.C) break;                                   // padding columns
    if ((c_grp >> 7) < (row0 >> 7)) continue;                  // below the diagonal: the mirror of a block done elsewhere
    const bool mirror = (c_grp >> 7) > (row0 >> 7);
    if constexpr (KP >= 8) {
      gram_epi_pool_group128<KP>(lane_addr + (uint32_t)g0, c_row, c_grp, p.C, p.g, p.scale, outp, mirror, atomics, lane);
    } else {
#pragma unroll 1
      for (int n0 = 0; n0 < 128; n0 += 32) {
        const int c_col0 = c_grp + n0;
        if (c_col0 >= p.C) break;
        float v[32];
        tmem_ld32(lane_addr + (uint32_t)(g0 + n0), v);
        if (KP > 0)
          gram_epi_pool_chunk<(KP > 0 ? KP : 1)>(v, c_row, c_col0, p.C, p.g, p.scale, outp, mirror, atomics, lane);
        else gram_epi_dense_chunk(v, c_row, c_col0, p.C, p.scale, outp, mirror, atomics);
      }
    }
  }
}

// p.nkb counts K blocks of KindTraits<KIND>::kElemsPerRow positions (64 for bf16, 32 for fp32/tf32).
// NHWC = false: features with x contiguous (NCHW): K-major operand tiles, 3-D map (x, c, b), one 128-row box per tile.
// NHWC = true : channels_last features (c contiguous): the same 128 channels x KB positions arrive as MN-major tiles
//               ([atom][KB positions][128 B of channels]) through ONE 4-D box per tile (make_tensor_map_nhwc_mn); only
//               the producer's coordinates and the descriptors differ, the accumulators and the epilogue are the same.
template <int KIND, int KP, bool NHWC>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kFpThreads, 1)
    gram_fwd_pair_kernel(const GramFwdParams p, const __grid_constant__ CUtensorMap tmap) {
  using T = KindTraits<KIND>;
  constexpr int KB = (int)T::kElemsPerRow;
  constexpr int E = (int)T::kElemsPerRow;                          // channels per 128 B atom (NHWC)

  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t bars = smem_base + kFpStages * kFpStageBytes;
  const uint32_t bar_full = bars, bar_empty = bars + 8 * kFpStages;
  const uint32_t bar_tfull = bars + 16 * kFpStages, bar_tempty = bar_tfull + 16;
  const uint32_t tmem_slot = bar_tempty + 16;
  volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(smem_raw + (tmem_slot - smem_u32(smem_raw)));

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const int pair = (int)cluster_id_x(), npairs = (int)cluster_nclusters_x();

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tmap);
    for (int s = 0; s < kFpStages; ++s) {
      mbar_init(bar_full + 8 * s, 1);      // the leader's arrive.expect_tx; both CTAs' TMA bytes complete on it
      mbar_init(bar_empty + 8 * s, 1);     // multicast tcgen05.commit
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(bar_tfull + 8 * a, 1);
      mbar_init(bar_tempty + 8 * a, 8);    // 4 epilogue warps x 2 CTAs
    }
    mbar_fence_init();
  }
  if (warp == 1) tmem_alloc2(tmem_slot, 512);
  tc_fence_before_sync();
  cluster_sync_all();
  tc_fence_after_sync();
  const uint32_t tmem_base = *tmem_slot_ptr;

  const uint32_t full_leader = mapa_u32(bar_full, 0);
  const uint32_t tempty_leader = mapa_u32(bar_tempty, 0);
  // One thread per warp issues the TMA / tcgen05 instructions. It is chosen with elect.sync, not `lane == 0`: ptxas
  // then knows the guarded region is single-threaded and feeds the instructions' uniform-register operands with plain
  // R2UR / uniform-datapath arithmetic; with a lane test it wraps every UTCHMMA / UTMALDG in an elect-broadcast-retry
  // loop (~150 cycles per instruction, which alone capped the tensor pipe at ~60 % in the first version).
  const bool elected = elect_one();

  if (warp == 0) {
    // =========================== TMA producer (one thread): this CTA's 128 rows of block I (and of block J) ==========
    if (elected) {
      uint32_t stage = 0, phase = 0;
      for (int u = pair; u < p.total_units; u += npairs) {
        const GramUnit w = gram_decode_unit(p, u);
        const bool diag = (w.I == w.J);
        const int rowA = w.I * 256 + (int)rank * 128, rowB = w.J * 256 + (int)rank * 128;
        const uint32_t tx_bytes = (diag ? 2u : 4u) * kFpTileBytes;
        for (int kb = w.kb0; kb < w.kb1; ++kb) {
          mbar_wait(bar_empty + 8 * stage, phase ^ 1u, 100u + stage);
          if (rank == 0) mbar_arrive_expect_tx(bar_full + 8 * stage, tx_bytes);
          const uint32_t a_tile = smem_base + stage * kFpStageBytes;
          if constexpr (NHWC) {
            tma_load_4d_pair(a_tile, &tmap, full_leader + 8 * stage, 0, kb * KB, rowA / E, w.b);
            if (!diag) tma_load_4d_pair(a_tile + kFpTileBytes, &tmap, full_leader + 8 * stage, 0, kb * KB, rowB / E, w.b);
          } else {
            tma_load_3d_pair(a_tile, &tmap, full_leader + 8 * stage, kb * KB, rowA, w.b);
            if (!diag) tma_load_3d_pair(a_tile + kFpTileBytes, &tmap, full_leader + 8 * stage, kb * KB, rowB, w.b);
          }
          if (++stage == kFpStages) { stage = 0; phase ^= 1u; }
        }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    // =========================== MMA issuer (one thread of the leader CTA) ===========================
    // Descriptors = a constant for stage 0 plus stage * 2048 and k-step * 2 (16 B units); a K block that lies fully
    // inside HW issues its four MMAs without bound checks. Keeping this loop to a few dozen instructions matters: one
    // thread executes ~10 cycles per dependent instruction and a K block's MMAs take only 512 cycles.
    if (rank == 0 && elected) {
      const uint32_t idesc = make_idesc(T::kFormat, 256, 256, NHWC ? 1 : 0, NHWC ? 1 : 0);
      constexpr uint32_t kChanBlockBytes = (uint32_t)KB * kRowBytes;   // NHWC: one 128 B-wide channel block of a tile
      const uint64_t d0 = !NHWC ? make_smem_desc_sw128(smem_base)
                                : (KIND == KIND_TF32 ? make_smem_desc_sw128b32_mnmajor(smem_base, kChanBlockBytes)
                                                     : make_smem_desc_sw128_mnmajor(smem_base, kChanBlockBytes));
      // per MMA: K-major tiles advance 32 B inside the 128 B row, MN-major ones by UMMA_K position rows of 128 B
      constexpr uint64_t kStageInc = kFpStageBytes >> 4, kTileInc = kFpTileBytes >> 4,
                         kKInc = NHWC ? (T::kUmmaK * kRowBytes) >> 4 : 32u >> 4;
      const int full_kb = p.HW / KB;
      uint32_t stage = 0, phase = 0, it = 0;
      for (int u = pair; u < p.total_units; u += npairs, ++it) {
        const GramUnit w = gram_decode_unit(p, u);
        const uint64_t b_off = (w.I == w.J) ? 0 : kTileInc;
        const uint32_t ab = it & 1u, use = it >> 1;
        mbar_wait_cl(bar_tempty + 8 * ab, (use & 1u) ^ 1u, 200u + ab);
        tc_fence_after_sync();
        const uint32_t acc = tmem_base + ab * 256u;
        for (int kb = w.kb0; kb < w.kb1; ++kb) {
          mbar_wait_cl(bar_full + 8 * stage, phase, 300u + stage);
          tc_fence_after_sync();
          const uint64_t da = d0 + stage * kStageInc, db = da + b_off;
          if (kb < full_kb) {
            umma2<KIND>(acc, da, db, idesc, kb != w.kb0 ? 1u : 0u);
            umma2<KIND>(acc, da + kKInc, db + kKInc, idesc, 1u);
            umma2<KIND>(acc, da + 2 * kKInc, db + 2 * kKInc, idesc, 1u);
            umma2<KIND>(acc, da + 3 * kKInc, db + 3 * kKInc, idesc, 1u);
          } else {
            for (uint32_t ks = 0; ks < T::kElemsPerRow / T::kUmmaK; ++ks) {
              if ((int)(kb * KB + ks * T::kUmmaK) >= p.HW) break;     // K tail: whole k-steps past HW are skipped
              umma2<KIND>(acc, da + ks * kKInc, db + ks * kKInc, idesc, (uint32_t)(kb - w.kb0) | ks);
            }
          }
          umma_commit2(bar_empty + 8 * stage);
          if (kb + 1 == w.kb1) umma_commit2(bar_tfull + 8 * ab);
          if (++stage == kFpStages) { stage = 0; phase ^= 1u; }
        }
      }
    }
    __syncwarp();
  } else {
    // =========================== epilogue ===========================
    const int q = warp & 3;
    const bool atomics = p.use_atomics != 0;
    uint32_t it = 0;
    for (int u = pair; u < p.total_units; u += npairs, ++it) {
      const GramUnit w = gram_decode_unit(p, u);
      const uint32_t ab = it & 1u, use = it >> 1;
      mbar_wait(bar_tfull + 8 * ab, use & 1u, 400u + ab);
      tc_fence_after_sync();
      gfp_epilogue_unit<KP>(p, w, tmem_base + ab * 256u, (int)rank, q, atomics, lane);
      tc_fence_before_sync();
      __syncwarp();
      if (lane == 0) mbar_arrive_cluster(tempty_leader + 8 * ab);
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

}  // namespace gh
