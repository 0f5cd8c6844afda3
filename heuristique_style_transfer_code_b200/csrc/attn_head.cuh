// Attention + classifier tail of the head, forward and backward, in fp32.
//
// Restates what the reference obtains from nn.MultiheadAttention(embed_dim=g*g, num_heads=1)(X, X, X),
// .mean(dim=0) and nn.Linear (Models/Models_RESNET50_TRUNCATE_GRAM_with_Attention.py:58-61 / :110-114;
// math in SURVEY.md Appendix A):
//   QKV = X W_in^T + b_in ; S = Q K^T / sqrt(E) per image (L x L, L <= 8) ; A = softmax_keys(S) ; O = A V
//   Obar = mean_l O[l]    ; emb = Obar W_o^T + b_o  (mean commutes with out_proj) ; logits = emb W_c^T + b_c
// The linear layers run on tcgen05 with split-bf16 operands (umma_gemm.cuh: fp32-level accuracy, needed because the
// scores feeding the softmax reach 1e4-1e5 on a random-init encoder); this file holds the per-image attention core,
// the classifier, the bias-gradient reductions and the fp32 FMA GEMM used for shapes the tensor-core kernel does not
// take (unaligned E, the nc-wide classifier products).
#pragma once
#include "common.cuh"

namespace gh {

// ---------------------------------------------------------------------------------------------
// Strided fp32 GEMM: D[m][n] (+)= sum_k A(m,k) B(k,n) + bias[n]
//   A(m,k) = A[m*a_sm + k*a_sk], B(k,n) = B[k*b_sk + n*b_sn], D row-major with leading dimension ldd.
//   AK = true when a_sk == 1 (A rows contiguous in k), BK = true when b_sk == 1.
// 64x64 tile, 16-deep slices, 256 threads, 4x4 outputs per thread.
// ---------------------------------------------------------------------------------------------
template <bool AK, bool BK>
__global__ void __launch_bounds__(256) sgemm_strided_kernel(const float* __restrict__ A, long long a_sm, long long a_sk,
                                                            const float* __restrict__ Bm, long long b_sk, long long b_sn,
                                                            const float* __restrict__ bias, float* __restrict__ D,
                                                            long long ldd, int M, int N, int K, int accumulate) {
  __shared__ float As[16][64 + 4];
  __shared__ float Bs[16][64 + 4];
  const int tid = threadIdx.x;
  const int m0 = blockIdx.y * 64, n0 = blockIdx.x * 64;
  const int tx = tid & 15, ty = tid >> 4;
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

  for (int k0 = 0; k0 < K; k0 += 16) {
    // 64 x 16 elements per operand, 4 per thread; the fast-varying thread index follows the contiguous axis
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const int idx = tid + e * 256;
      int mm, kk;
      if (AK) { kk = idx & 15; mm = idx >> 4; } else { mm = idx & 63; kk = idx >> 6; }
      const int m = m0 + mm, k = k0 + kk;
      As[kk][mm] = (m < M && k < K) ? __ldg(A + m * a_sm + k * a_sk) : 0.f;
      int nn, kb;
      if (BK) { kb = idx & 15; nn = idx >> 4; } else { nn = idx & 63; kb = idx >> 6; }
      const int n = n0 + nn, k2 = k0 + kb;
      Bs[kb][nn] = (n < N && k2 < K) ? __ldg(Bm + k2 * b_sk + n * b_sn) : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < 16; ++kk) {
      float a[4], b[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) a[i] = As[kk][ty * 4 + i];
#pragma unroll
      for (int j = 0; j < 4; ++j) b[j] = Bs[kk][tx * 4 + j];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int m = m0 + ty * 4 + i;
    if (m >= M) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int n = n0 + tx * 4 + j;
      if (n >= N) continue;
      float v = acc[i][j] + (bias ? bias[n] : 0.f);
      float* d = D + m * ldd + n;
      *d = accumulate ? (*d + v) : v;
    }
  }
}

inline cudaError_t launch_sgemm(const float* A, long long a_sm, long long a_sk, const float* Bm, long long b_sk,
                                long long b_sn, const float* bias, float* D, long long ldd, int M, int N, int K,
                                int accumulate, cudaStream_t st) {
  if (M <= 0 || N <= 0) return cudaSuccess;
  dim3 grid((N + 63) / 64, (M + 63) / 64), block(256);
  const bool ak = (a_sk == 1), bk = (b_sk == 1);
  if (ak && bk) sgemm_strided_kernel<true, true><<<grid, block, 0, st>>>(A, a_sm, a_sk, Bm, b_sk, b_sn, bias, D, ldd, M, N, K, accumulate);
  else if (ak) sgemm_strided_kernel<true, false><<<grid, block, 0, st>>>(A, a_sm, a_sk, Bm, b_sk, b_sn, bias, D, ldd, M, N, K, accumulate);
  else if (bk) sgemm_strided_kernel<false, true><<<grid, block, 0, st>>>(A, a_sm, a_sk, Bm, b_sk, b_sn, bias, D, ldd, M, N, K, accumulate);
  else sgemm_strided_kernel<false, false><<<grid, block, 0, st>>>(A, a_sm, a_sk, Bm, b_sk, b_sn, bias, D, ldd, M, N, K, accumulate);
  return cudaGetLastError();
}

// out[n] = sum_m X[m][n]   (bias gradients). One thread per column, 32 columns per block row-strip loop.
__global__ void colsum_kernel(const float* __restrict__ X, long long ldx, float* __restrict__ out, int M, int N) {
  __shared__ float red[8][33];
  const int n = blockIdx.x * 32 + threadIdx.x;
  float s = 0.f;
  if (n < N)
    for (int m = threadIdx.y; m < M; m += 8) s += X[m * ldx + n];
  red[threadIdx.y][threadIdx.x] = s;
  __syncthreads();
  if (threadIdx.y == 0 && n < N) {
    float t = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) t += red[i][threadIdx.x];
    out[n] = t;
  }
}

constexpr int kMaxL = 8;

__device__ __forceinline__ float block_sum_128(float v, float* scratch) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
  __syncthreads();
  if (l == 0) scratch[w] = v;
  __syncthreads();
  return scratch[0] + scratch[1] + scratch[2] + scratch[3];
}

// One CTA (128 threads) per image. QKV: (B*L, 3E) rows ordered b*L + l. Writes probs (B, L, L) and Obar (B, E).
// All L*L score dot products are accumulated in one pass over E (each thread keeps L*L partial sums), followed by a
// single block reduction, instead of L*L separate reductions.
template <int LT>
__global__ void __launch_bounds__(128) attn_core_fwd_kernel(const float* __restrict__ QKV, float* __restrict__ probs,
                                                            float* __restrict__ obar, int L, int E) {
  __shared__ float red[4][LT * LT];
  __shared__ float sc[LT * LT];
  const int b = blockIdx.x;
  const float* base = QKV + (long long)b * L * 3 * E;
  const float inv_sqrt_e = rsqrtf((float)E);
  float acc[LT * LT];
#pragma unroll
  for (int i = 0; i < LT * LT; ++i) acc[i] = 0.f;
  for (int e = threadIdx.x; e < E; e += 128) {
    float qv[LT], kv[LT];
#pragma unroll
    for (int l = 0; l < LT; ++l) {
      qv[l] = l < L ? base[(long long)l * 3 * E + e] * inv_sqrt_e : 0.f;   // torch scales q before the product
      kv[l] = l < L ? base[(long long)l * 3 * E + E + e] : 0.f;
    }
#pragma unroll
    for (int l = 0; l < LT; ++l)
#pragma unroll
      for (int m = 0; m < LT; ++m) acc[l * LT + m] = fmaf(qv[l], kv[m], acc[l * LT + m]);
  }
  const int w = threadIdx.x >> 5, ln = threadIdx.x & 31;
#pragma unroll
  for (int i = 0; i < LT * LT; ++i) {
    float v = acc[i];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if (ln == 0) red[w][i] = v;
  }
  __syncthreads();
  if (threadIdx.x < LT * LT) sc[threadIdx.x] = red[0][threadIdx.x] + red[1][threadIdx.x] + red[2][threadIdx.x] + red[3][threadIdx.x];
  __syncthreads();
  if (threadIdx.x < L) {
    const int l = threadIdx.x;
    float mx = -INFINITY;
    for (int m = 0; m < L; ++m) mx = fmaxf(mx, sc[l * LT + m]);
    float den = 0.f;
    for (int m = 0; m < L; ++m) den += expf(sc[l * LT + m] - mx);
    for (int m = 0; m < L; ++m) {
      const float pr = expf(sc[l * LT + m] - mx) / den;
      sc[l * LT + m] = pr;
      probs[((long long)b * L + l) * L + m] = pr;
    }
  }
  __syncthreads();
  // Obar[e] = (1/L) sum_l sum_m A[l][m] V[m][e] = sum_m (mean_l A[l][m]) V[m][e]
  float wm[LT];
#pragma unroll
  for (int m = 0; m < LT; ++m) {
    float a = 0.f;
    for (int l = 0; l < L; ++l) a += sc[l * LT + m];
    wm[m] = m < L ? a / (float)L : 0.f;
  }
  for (int e = threadIdx.x; e < E; e += 128) {
    float o = 0.f;
#pragma unroll
    for (int m = 0; m < LT; ++m)
      if (m < L) o = fmaf(wm[m], base[(long long)m * 3 * E + 2 * E + e], o);
    obar[(long long)b * E + e] = o;
  }
}

// logits[b][c] = <emb[b], W_c[c]> + b_c[c]: one warp per (image, class) pair, 4 pairs per 128-thread block.
__global__ void __launch_bounds__(128) classifier_fwd_kernel(const float* __restrict__ emb, const float* __restrict__ Wc,
                                                             const float* __restrict__ bc, float* __restrict__ logits,
                                                             int B, int E, int nc) {
  const int pair = blockIdx.x * 4 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (pair >= B * nc) return;
  const int b = pair / nc, c = pair % nc;
  const float* x = emb + (long long)b * E;
  const float* w = Wc + (long long)c * E;
  float s = 0.f;
  for (int e = lane; e < E; e += 32) s = fmaf(x[e], __ldg(w + e), s);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if (lane == 0) logits[pair] = s + (bc ? bc[c] : 0.f);
}

// Backward of the per-image core. dObar: (B, E). Writes dQKV (B*L, 3E).
__global__ void __launch_bounds__(128) attn_core_bwd_kernel(const float* __restrict__ QKV, const float* __restrict__ probs,
                                                            const float* __restrict__ dobar, float* __restrict__ dQKV,
                                                            int L, int E) {
  __shared__ float sA[kMaxL * kMaxL];
  __shared__ float sdS[kMaxL * kMaxL];
  __shared__ float sdAv[kMaxL];
  __shared__ float scratch[4];
  const int b = blockIdx.x;
  const float* base = QKV + (long long)b * L * 3 * E;
  float* dbase = dQKV + (long long)b * L * 3 * E;
  const float* dO = dobar + (long long)b * E;
  const float inv_l = 1.f / (float)L;
  const float inv_sqrt_e = rsqrtf((float)E);
  if (threadIdx.x < L * L) {
    const int l = threadIdx.x / L, m = threadIdx.x % L;
    sA[l * kMaxL + m] = probs[((long long)b * L + l) * L + m];
  }
  // dA[l][m] = <dO[l], V[m]> with dO[l] = dObar / L for every l  -> depends on m only
  for (int m = 0; m < L; ++m) {
    const float* vp = base + (long long)m * 3 * E + 2 * E;
    float s = 0.f;
    for (int e = threadIdx.x; e < E; e += 128) s = fmaf(dO[e], vp[e], s);
    s = block_sum_128(s, scratch);
    if (threadIdx.x == 0) sdAv[m] = s * inv_l;
  }
  __syncthreads();
  if (threadIdx.x < L) {
    const int l = threadIdx.x;
    float dot = 0.f;
    for (int m = 0; m < L; ++m) dot += sdAv[m] * sA[l * kMaxL + m];
    for (int m = 0; m < L; ++m) sdS[l * kMaxL + m] = sA[l * kMaxL + m] * (sdAv[m] - dot) * inv_sqrt_e;
  }
  __syncthreads();
  float colA[kMaxL];
  for (int m = 0; m < L; ++m) {
    float a = 0.f;
    for (int l = 0; l < L; ++l) a += sA[l * kMaxL + m];
    colA[m] = a * inv_l;
  }
  for (int e = threadIdx.x; e < E; e += 128) {
    float qv[kMaxL], kv[kMaxL];
    for (int l = 0; l < L; ++l) {
      qv[l] = base[(long long)l * 3 * E + e];
      kv[l] = base[(long long)l * 3 * E + E + e];
    }
    const float doe = dO[e];
    for (int l = 0; l < L; ++l) {
      float dq = 0.f, dk = 0.f;
      for (int m = 0; m < L; ++m) {
        dq = fmaf(sdS[l * kMaxL + m], kv[m], dq);   // dQ[l] = sum_m dS[l][m] K[m]
        dk = fmaf(sdS[m * kMaxL + l], qv[m], dk);   // dK[l] = sum_m dS[m][l] Q[m]
      }
      dbase[(long long)l * 3 * E + e] = dq;
      dbase[(long long)l * 3 * E + E + e] = dk;
      dbase[(long long)l * 3 * E + 2 * E + e] = colA[l] * doe;   // dV[l] = sum_i A[i][l] dO[i]
    }
  }
}

}  // namespace gh
