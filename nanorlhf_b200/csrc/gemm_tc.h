// Host-visible parameter block of the general tcgen05 GEMM (gemm_tc.cu).
#pragma once
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>

namespace nrl {
namespace tc {

enum TcEpilogue : int {
  EPI_BF16 = 0,      // D (bf16, TMA store) = alpha * acc (+ bias) (GELU)
  EPI_F32 = 1,       // out_f32 (row-major, plain stores) (+)= alpha * acc
  EPI_SWIGLU = 2,    // D[M, N/2] (bf16, TMA store) = silu(gate) * up; B rows interleaved per 32 features [32 gate | 32 up]
};

struct TcParams {
  int M, N, K;                  // D is [M, N]; K = contraction length of the first operand pair
  int K2;                       // contraction length of the second operand pair (0 = none)
  const float* row_scale;       // FP8 kernels: [M] per-token dequantisation scale of A
  const float* col_scale;       // FP8 kernels: [N] per-output-channel dequantisation scale of B
  int batch;                    // BATCH kernels: number of independent [M,K] x [N,K]^T problems behind 3D tensor maps (else 1)
  float alpha;                  // multiplies the accumulator
  const __nv_bfloat16* bias;    // EPI_BF16: optional [N]
  int act;                      // EPI_BF16: 0 = none, 1 = exact (erf) GELU after the bias
  float* out_f32;               // EPI_F32: [M, N] fp32, row stride out_f32_stride elements
  long out_f32_stride;
  int accumulate;               // EPI_F32: 1 = add to what is there
  // split-K (SPLIT kernels; small outputs with a long contraction -- LoRA wgrads, rank-r projections): every tile is
  // computed by `splits` CTAs over disjoint k-ranges; partial tiles go to `ws` (fp32), the last CTA to arrive on the
  // tile's counter sums them in split order (deterministic) and stores the bf16 result.  Counters reset themselves.
  int splits;
  float* ws;                    // [splits, tiles, 128, BN] fp32
  int* counters;                // [tiles] int32, zero between launches
  __nv_bfloat16* out_bf16;      // SPLIT: D written with plain stores (row stride out_stride elements)
  long out_stride;
};

}  // namespace tc
}  // namespace nrl

extern "C" cudaError_t nrl_gemm_tc_splitk(const CUtensorMap* maps, const nrl::tc::TcParams* p, int bn, int a_mn, int b_mn, int num_sms,
                                          cudaStream_t stream);
extern "C" cudaError_t nrl_gemm_tc_fp8(const CUtensorMap* maps, const nrl::tc::TcParams* p, int cg, int bn, int swiglu, int num_sms,
                                       cudaStream_t stream);
// batched K-major x K-major GEMM, bf16 out; maps = {A, B, D} 3D (tma_host.h make_tma_3d), B box rows = bn / cg
extern "C" cudaError_t nrl_gemm_tc_batched(const CUtensorMap* maps, const nrl::tc::TcParams* p, int cg, int bn, int num_sms,
                                           cudaStream_t stream);
extern "C" cudaError_t nrl_gemm_tc(const CUtensorMap* maps, const nrl::tc::TcParams* p, int cg, int bn, int a_mn, int b_mn,
                                   int epi, int num_sms, cudaStream_t stream);
