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
};

struct TcParams {
  int M, N, K;                  // D is [M, N]; K = contraction length of the first operand pair
  int K2;                       // contraction length of the second operand pair (0 = none)
  float alpha;                  // multiplies the accumulator
  const __nv_bfloat16* bias;    // EPI_BF16: optional [N]
  int act;                      // EPI_BF16: 0 = none, 1 = exact (erf) GELU after the bias
  float* out_f32;               // EPI_F32: [M, N] fp32, row stride out_f32_stride elements
  long out_f32_stride;
  int accumulate;               // EPI_F32: 1 = add to what is there
};

}  // namespace tc
}  // namespace nrl

extern "C" cudaError_t nrl_gemm_tc(const CUtensorMap* maps, const nrl::tc::TcParams* p, int cg, int bn, int a_mn, int b_mn,
                                   int epi, int num_sms, cudaStream_t stream);
