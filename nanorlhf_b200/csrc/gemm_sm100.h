// Host-visible parameter block + epilogue selectors of the tcgen05 GEMM family (gemm_sm100.cu).
#pragma once
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>

namespace nrl {

enum GemmEpilogue : int {
  EPI_STORE = 0,     // D = A B^T (+ bias)                      -> bf16 [M,N]
  EPI_LOGPROB = 1,   // fused lm-head log-prob partials          -> float4 [n_splits, M]
  EPI_DLOGITS = 2,   // dZ = (onehot - softmax) * g / T          -> bf16 [M,N]
  EPI_SWIGLU = 4,    // D[M, N/2] = silu(gate) * up with [32 gate | 32 up] interleaved weight rows   -> bf16
  EPI_MERGE = 3,     // D = addend + scale * (A B^T)  (K-BC LoRA merge: W + (alpha/r) B A) -> bf16 [M,N]
};

struct GemmParams {
  int M, N, K;
  int n_splits;                 // EPI_LOGPROB: vocab splits (work items = m_tiles * n_splits)
  float scale;                  // 1 / temperature for the log-prob epilogues
  const __nv_bfloat16* bias;    // EPI_STORE: optional [N]
  const int* targets;           // EPI_LOGPROB / EPI_DLOGITS: [M] int32
  const float* lse;             // EPI_DLOGITS: [M]
  const float* grad_logp;       // EPI_DLOGITS: [M]
  float* partials;              // EPI_LOGPROB: [n_splits, M, 4]
  const float* row_scale;       // fp8 GEMM: [M] per-token dequant scale of A
  const float* col_scale;       // fp8 GEMM: [N] per-output-channel dequant scale of B
  const __nv_bfloat16* addend;  // EPI_MERGE: [M, N] row-major, row stride addend_stride elements
  long addend_stride;
  int act;                      // EPI_STORE: 0 = none, 1 = exact (erf) GELU applied after the bias
  // EPI_MERGE, K-BC across ranks: when non-null the merged tile is written with `multimem.st` to this NVLS multicast
  // address (element [0,0] of D in every rank's sampler arena, row stride mc_stride elements) instead of a local TMA store
  __nv_bfloat16* mc_out;
  long mc_stride;
};

}  // namespace nrl

extern "C" cudaError_t nrl_gemm_bf16_tn(const CUtensorMap* tmA, const CUtensorMap* tmB, const CUtensorMap* tmD,
                                        const nrl::GemmParams* p, int block_n, int epi, int num_sms,
                                        cudaStream_t stream);
// experimental 2-CTA (cta_group::2) bf16 GEMM, EPI_STORE only (gemm_sm100_2cta.cu); B map box = 128 rows
extern "C" cudaError_t nrl_gemm_bf16_tn_2cta(const CUtensorMap* tmA, const CUtensorMap* tmB, const CUtensorMap* tmD,
                                             const nrl::GemmParams* p, int swiglu, int num_sms, cudaStream_t stream);
extern "C" cudaError_t nrl_gemm_fp8_tn(const CUtensorMap* tmA, const CUtensorMap* tmB, const CUtensorMap* tmD,
                                       const nrl::GemmParams* p, int block_n, int epi, int num_sms, cudaStream_t stream);
extern "C" cudaError_t nrl_lmhead_combine(const float* partials, int M, int n_splits, float* logp, float* entropy,
                                          float* lse, cudaStream_t stream);
