// Gram forward for bf16 feature maps: operands TMA-staged straight into the UMMA SWIZZLE_128B layout.
//
// Same units, smem ring, TMEM plan, MMA issue and epilogue as gram_fwd.cuh (they are shared code); only the producer
// differs: when the activations are already bf16 (backbone under autocast) no conversion is needed, so one thread
// issues cp.async.bulk.tensor (3-D tensor map over (HW, C, B), box 64 x 256 x 1, 128 B swizzle) and the 32 KB stage
// lands asynchronously with mbarrier complete_tx -- up to 6 stages (192 KB) in flight per SM with no register staging.
// Out-of-bounds rows / columns (C or HW tails) are zero-filled by the TMA unit.
// Requirements (else the launcher uses the ld.global producers of gram_fwd.cuh): base pointer 16 B aligned, row and
// image pitches multiples of 16 B (HW % 8 == 0 for dense NCHW: 3136, 784, 12544 ... but not 196 or 49).
// Warps: 0 = TMA producer, 1 = TMEM owner + MMA issuer, 2-5 = epilogue (warp % 4 = TMEM lane quarter).
#pragma once
#include <cuda.h>
#include "gram_fwd.cuh"

namespace gh {

constexpr int kGtThreads = 6 * 32;

__device__ __forceinline__ void tma_load_3d(uint32_t dst_smem, const CUtensorMap* tmap, uint32_t bar, int c0, int c1,
                                            int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
      ::"r"(dst_smem), "l"(reinterpret_cast<uint64_t>(tmap)), "r"(bar), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* tmap) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(tmap)) : "memory");
}

template <int KP>
__global__ void __launch_bounds__(kGtThreads, 1) gram_fwd_tma_kernel(const GramFwdParams p,
                                                                     const __grid_constant__ CUtensorMap tmap) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t bars = smem_base + kGfStages * kGfStageBytes;
  const uint32_t bar_full = bars, bar_empty = bars + 8 * kGfStages;
  const uint32_t bar_tfull = bars + 16 * kGfStages, bar_tempty = bar_tfull + 8;
  const uint32_t tmem_slot = bar_tempty + 8;
  volatile uint32_t* tmem_slot_ptr =
      reinterpret_cast<volatile uint32_t*>(smem_raw + (tmem_slot - smem_u32(smem_raw)));
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tmap);
    for (int s = 0; s < kGfStages; ++s) {
      mbar_init(bar_full + 8 * s, 1);     // the producer's arrive.expect_tx; the TMA completes the byte count
      mbar_init(bar_empty + 8 * s, 1);
    }
    mbar_init(bar_tfull, 1);
    mbar_init(bar_tempty, 4);
    mbar_fence_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, kGfTmemCols);
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem_base = *tmem_slot_ptr;

  if (warp == 0) {
    // =========================== TMA producer (one thread) ===========================
    if (lane == 0) {
      uint32_t stage = 0, phase = 0;
      GfItem it;
      bool have = gf_item_first(it, p);
      while (have) {
        mbar_wait(bar_empty + 8 * stage, phase ^ 1u, 100u + stage);
        mbar_arrive_expect_tx(bar_full + 8 * stage, kGfStageBytes);
        const int blk = it.h == 0 ? it.w.I : it.w.J;
        tma_load_3d(smem_base + stage * kGfStageBytes, &tmap, bar_full + 8 * stage, it.kb * 64, blk * 256, it.w.b);
        if (++stage == kGfStages) { stage = 0; phase ^= 1u; }
        have = gf_item_next(it, p);
      }
    }
  } else if (warp == 1) {
    // =========================== MMA issuer ===========================
    uint32_t stage = 0, phase = 0, acc_phase = 0;
    for (int u = blockIdx.x; u < p.total_units; u += gridDim.x) {
      const GramUnit w = gram_decode_unit(p, u);
      gf_issue_unit(p, w, smem_base, bar_full, bar_empty, bar_tfull, bar_tempty, tmem_base, stage, phase, acc_phase, lane);
      acc_phase ^= 1u;
    }
  } else {
    // =========================== epilogue ===========================
    const int q = warp & 3;
    uint32_t acc_phase = 0;
    const bool atomics = p.use_atomics != 0;
    for (int u = blockIdx.x; u < p.total_units; u += gridDim.x) {
      const GramUnit w = gram_decode_unit(p, u);
      mbar_wait(bar_tfull, acc_phase, 400u);
      tc_fence_after_sync();
      gf_epilogue_unit<KP>(p, w, tmem_base, q, 0, 1, atomics, lane);
      tc_fence_before_sync();
      __syncwarp();
      if (lane == 0) mbar_arrive(bar_tempty);
      acc_phase ^= 1u;
    }
  }

  tc_fence_before_sync();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after_sync();
    tmem_dealloc(tmem_base, kGfTmemCols);
  }
}

// Host side: cuTensorMapEncodeTiled through the runtime's driver entry point (no link-time dependency on libcuda).
typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

inline PFN_encodeTiled get_encode_tiled() {
  static PFN_encodeTiled fn = nullptr;
  static bool tried = false;
  if (!tried) {
    tried = true;
    void* ptr = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<PFN_encodeTiled>(ptr);
  }
  return fn;
}

// Returns false when the layout cannot be described (caller falls back to the ld.global producers).
inline bool make_feature_tensor_map(CUtensorMap* map, const void* F, long long img_stride, long long row_stride, int B,
                                    int C, int HW) {
  PFN_encodeTiled enc = get_encode_tiled();
  if (!enc) return false;
  if ((reinterpret_cast<uintptr_t>(F) & 15u) || (row_stride * 2) % 16 || (img_stride * 2) % 16) return false;
  if (row_stride < HW) return false;
  cuuint64_t dims[3] = {(cuuint64_t)HW, (cuuint64_t)C, (cuuint64_t)B};
  cuuint64_t strides[2] = {(cuuint64_t)row_stride * 2, (cuuint64_t)img_stride * 2};
  cuuint32_t box[3] = {64, 256, 1};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(F), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS;
}

}  // namespace gh
