// C launch API of the non-GEMM kernels (elementwise.cu, rl_kernels.cu, sampling.cu, attention.cu, comm.cu).
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace nrl {
struct AdamHyper {
  float lr, beta1, beta2, eps, wd;
  float step_size;    // lr / (1 - beta1^t)
  float inv_bc2;      // 1 / (1 - beta2^t)
  float grad_scale;   // multiplies the gradient before use (1/world, 1/accum ...)
};
}  // namespace nrl

extern "C" {
cudaError_t nrl_add_layernorm(const void* x, const void* residual, const void* w, const void* b, void* y, int rows, int d,
                              float eps, cudaStream_t s);
cudaError_t nrl_rmsnorm(const void* x, const void* residual, const void* w, void* y, void* residual_out, float* rstd,
                        int rows, int d, float eps, cudaStream_t s);
cudaError_t nrl_rmsnorm_bwd(const void* x, const void* w, const void* gy, const float* rstd, void* gx, int rows, int d,
                            cudaStream_t s);
cudaError_t nrl_rope(const void* x, void* y, const float* cos_t, const float* sin_t, int T, int H, int D,
                     long x_stride_t, long y_stride_t, float sin_sign, cudaStream_t s);
cudaError_t nrl_swiglu(const void* gate, const void* up, long in_stride, void* out, long T, int F, cudaStream_t s);
cudaError_t nrl_swiglu_bwd(const void* gate, const void* up, long in_stride, const void* gout, void* dgate, void* dup,
                           long out_stride, long T, int F, cudaStream_t s);

cudaError_t nrl_gae_scan(const float* rewards, const float* values, float* adv, float* returns, int B, int T,
                         float gamma, float lam, cudaStream_t s);
cudaError_t nrl_policy_loss(const float* new_lp, const float* old_lp, const float* adv, const uint8_t* mask,
                            const float* ref_lp, float cliprange, float kl_coef, long n, float* grad_unnorm,
                            float* acc, cudaStream_t s);
cudaError_t nrl_value_loss(const float* vpred, const float* vold, const float* ret, const uint8_t* mask, float clip,
                           long n, float* grad_unnorm, float* acc, cudaStream_t s);
cudaError_t nrl_adamw_flat(void* param, const void* grad, void* m, void* v, float* master, long n, int moments_bf16,
                           nrl::AdamHyper h, cudaStream_t s);

cudaError_t nrl_sample(const void* logits, int is_bf16, long row_stride, int rows, int V, float temperature,
                       float top_p, unsigned long long seed, unsigned long long step, const int* row_ids,
                       const int* row_steps, int* out_tokens, int impl, cudaStream_t s);
}

extern "C" {
cudaError_t nrl_kv_cache_write(const void* k, const void* v, long k_stride_t, long v_stride_t, void* k_cache,
                               void* v_cache, const int* slot_mapping, const int* src_index, int T, int Hkv, int head_dim, int page,
                               cudaStream_t s);
cudaError_t nrl_paged_decode(const void* q, long q_stride_s, const void* k_cache, const void* v_cache,
                             const int* block_tables, const int* context_lens, void* out, float* part_o,
                             float* part_ml, int S, int Hq, int Hkv, int head_dim, int page, int max_blocks, int splits,
                             float scale, cudaStream_t s);
}

extern "C" {
cudaError_t nrl_allreduce_adam(const void* const* grad_ptrs, void* const* param_ptrs, const void* grad_mc, void* param_mc,
                               void* m, void* v, float* master, long lo, long n, int world, int rank, int moments_bf16,
                               int use_multicast, nrl::AdamHyper h, int max_blocks, cudaStream_t s);
cudaError_t nrl_allreduce_sum(void* const* buf_ptrs, long lo, long n, int world, int rank, float scale, int max_blocks,
                              cudaStream_t s);
}

extern "C" {
cudaError_t nrl_attn_varlen_fwd(const void* q, const void* k, const void* v, void* out, float* lse, long qs, long ks,
                                long vs, long os, const int* cu, int num_seqs, int total, int Hq, int Hkv, int D,
                                float scale, int causal, const void* rel_a, const void* rel_b, const short* lut,
                                int lut_center, int NB, cudaStream_t s);
cudaError_t nrl_attn_varlen_bwd(const void* dout, const void* q, const void* k, const void* v, const void* o,
                                const float* lse, float* delta, void* dq, void* dk, void* dv, long qs, long ks, long vs,
                                long os, const int* cu, int num_seqs, int total, int Hq, int Hkv, int D, float scale,
                                cudaStream_t s);
// tcgen05 forward (attention_fwd_tc.cu): causal, head_dim 128, q/k/v addressed through TMA maps over [T, H*D]
cudaError_t nrl_attn_fwd_tc(const CUtensorMap* tmQ, const CUtensorMap* tmK, const CUtensorMap* tmV, void* out, float* lse,
                            long o_stride_t, const int* cu, int num_seqs, int total, int Hq, int Hkv, float scale,
                            cudaStream_t s, long long* prof = nullptr);
// TMA-fed DeBERTa disentangled attention (attention_varlen.cu); maps = {Q, K, V, relA, relB}
cudaError_t nrl_deberta_attn_fwd(const CUtensorMap* maps, void* out, float* lse, long os, const int* cu, int num_seqs,
                                 int total, int Hq, float scale, const short* lut, int lut_center, int NB, int bn, cudaStream_t s);
// tcgen05 backward (attention_bwd_tc.cu): delta + dK/dV + dQ; `maps` = 8 tensor maps (see the .cu)
cudaError_t nrl_attn_bwd_tc(const CUtensorMap* maps, const void* o, const void* dout, const float* lse, float* delta,
                            void* dq, void* dk, void* dv, long o_stride_t, long dq_stride_t, long dkv_stride_t,
                            const int* cu, int num_seqs, int total, int Hq, int Hkv, float scale, cudaStream_t s);
}

extern "C" cudaError_t nrl_quant_rows_e4m3(const void* x, long x_stride, void* q, long q_stride, float* scale, int M, int K,
                                           cudaStream_t s);

extern "C" {
cudaError_t nrl_kv_cache_write_fp8(const void* k, const void* v, long k_stride_t, long v_stride_t, void* kq, void* vq, float* ks,
                                   float* vs, const int* slot_mapping, const int* src_index, int pairs, int Hkv, int head_dim,
                                   int page, cudaStream_t s);
cudaError_t nrl_paged_decode_fp8(const void* q, long q_stride_s, const void* kq, const void* vq, const float* ks, const float* vs,
                                 const int* block_tables, const int* context_lens, void* out, float* part_o, float* part_ml, int S,
                                 int Hq, int Hkv, int head_dim, int page, int max_blocks, int splits, float scale, cudaStream_t s);
}

extern "C" cudaError_t nrl_rope_kv_write(void* qkv, long stride_s, const float* cos_t, const float* sin_t, void* k_cache, void* v_cache,
                                         float* k_scale, float* v_scale, const int* slot_mapping, int S, int Hq, int Hkv, int head_dim,
                                         int page, int kv8, cudaStream_t s);
