// Packed-varlen flash attention, forward + backward (K1 prefill, K9/K14 training + log-prob passes) and the
// fused DeBERTa-v3 disentangled-attention forward (K8).
//
//   * tokens are packed [T_total, H, D] with cu_seqlens -- no padding; GQA by head index (kv = q / G);
//   * online softmax, scores never leave registers; mma.sync m16n8k16 bf16 tensor-core math with fp32
//     accumulation, ldmatrix from XOR-swizzled shared-memory tiles, cp.async double buffering;
//   * backward = three kernels: delta = rowsum(dO * O); dK/dV (one CTA per KV block, loops query blocks
//     and the G query heads of the kv head, S^T formulation); dQ (one CTA per query block) -- no atomics,
//     deterministic;
//   * REL_BIAS: score += A[i, c(i-j)] + B[j, c(i-j)] with c = log-bucketed relative position: the c2p and
//     p2c terms of DeBERTa's disentangled attention, gathered from per-row / per-key tables that stay in
//     shared memory, so the three score terms are fused in one flash-style pass
//     (reference path: HF eager DeBERTa, /root/reference/GRPO/grpo.py:189-192).
// These replace flash-attn-2 (attn_implementation="flash_attention_2", /root/reference/GRPO/grpo.py:219).
#include "common.cuh"
#include "kernels.h"

namespace nrl {

NRL_DEVICE void ldsm_x4(uint32_t (&r)[4], uint32_t addr) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}
NRL_DEVICE void ldsm_x4_t(uint32_t (&r)[4], uint32_t addr) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}
NRL_DEVICE void mma16816(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

// byte offset of 16-byte chunk c of row r in a [rows][D x bf16] tile, XOR-swizzled (conflict-free ldmatrix)
template <int D>
NRL_DEVICE uint32_t tile_off(int r, int c) {
  return static_cast<uint32_t>(r * (D * 2) + ((c ^ (r & 7)) << 4));
}

// one bf16 from shared memory (32-bit shared address) as fp32
NRL_DEVICE float lds_bf16(uint32_t addr) {
  unsigned short h;
  asm volatile("ld.shared.u16 %0, [%1];" : "=h"(h) : "r"(addr));
  return __uint_as_float(static_cast<uint32_t>(h) << 16);
}

struct AttnParams {
  const __nv_bfloat16 *q, *k, *v, *o, *dout;
  __nv_bfloat16 *out, *dq, *dk, *dv;
  float *lse, *delta;                    // [Hq, T_total]
  const int* cu_seqlens;                 // [S+1]
  long q_stride_t, k_stride_t, v_stride_t, o_stride_t;   // token strides in elements (head stride = D)
  int num_seqs, total_tokens, Hq, Hkv, G;
  float scale, scale_log2;
  // DeBERTa relative-position bias
  const __nv_bfloat16 *rel_a, *rel_b;    // [Hq, T_total, NB]
  const short* bucket_lut;               // [2 * lut_center + 1] : delta -> clamped bucket index
  int lut_center, NB;
};

// Map a block index to (sequence, row block) for packed sequences; returns false when out of range.
template <int BM>
NRL_DEVICE bool locate_block(const int* cu, int num_seqs, int blk, int& seq, int& m_blk, int& seq_start, int& seq_len) {
  __shared__ int s_info[4];
  if (threadIdx.x == 0) {
    int acc = 0, found = 0;
    for (int s = 0; s < num_seqs; ++s) {
      int a = cu[s], b = cu[s + 1];
      int nb = (b - a + BM - 1) / BM;
      if (blk < acc + nb) {
        s_info[0] = s; s_info[1] = blk - acc; s_info[2] = a; s_info[3] = b - a;
        found = 1;
        break;
      }
      acc += nb;
    }
    if (!found) s_info[0] = -1;
  }
  __syncthreads();
  seq = s_info[0];
  if (seq < 0) return false;
  m_blk = s_info[1]; seq_start = s_info[2]; seq_len = s_info[3];
  return true;
}

// cooperative async load of a [ROWS][D] tile from rows [row0, row0+ROWS) of a packed [T, H, D] tensor
template <int D, int ROWS, int THREADS>
NRL_DEVICE void load_tile_async(uint8_t* smem_tile, const __nv_bfloat16* base, long stride_t, int head, int row0,
                                int row_limit) {
  constexpr int CH = D / 8;
  for (int i = threadIdx.x; i < ROWS * CH; i += THREADS) {
    const int r = i / CH, c = i % CH;
    const bool ok = (row0 + r) < row_limit;
    const __nv_bfloat16* src = base + static_cast<long>(ok ? (row0 + r) : row0) * stride_t + head * D + c * 8;
    cp_async_16_zfill(smem_tile + tile_off<D>(r, c), src, ok);
  }
}

constexpr int kRelBW = 112;   // staged window of the bias tables: >= 64 + 32 - 1 + 7 columns, multiple of 8
constexpr int kRelStride = 120;   // smem row stride of a window (elements): 240 B = 60 words -> rows land 28 banks apart

// first (8-aligned) table column needed by the score tile (rows q0..q0+63, keys k0..k0+BN-1)
NRL_DEVICE int rel_window_lo(const AttnParams& p, int q0, int k0, int BN) {
  const int dmin = q0 - (k0 + BN - 1);
  return p.bucket_lut[min(max(dmin, -p.lut_center), p.lut_center) + p.lut_center] & ~7;
}

// =================================================================================================
// forward
// =================================================================================================
template <int D, bool CAUSAL, bool REL_BIAS, int BN>
__global__ void __launch_bounds__(128) flash_fwd_kernel(AttnParams p) {
  constexpr int BM = 64, KS = D / 16, NT = BN / 8;
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* sQ = smem;                              // [64][D]
  uint8_t* sK = sQ + BM * D * 2;                   // [2][BN][D]
  uint8_t* sV = sK + 2 * BN * D * 2;               // [2][BN][D]
  uint8_t* sA = sV + 2 * BN * D * 2;               // REL_BIAS: [2][BM][kRelStride] bf16 (column window, see below)
  uint8_t* sB = sA + (REL_BIAS ? 2 * BM * kRelStride * 2 : 0);   // REL_BIAS: [2][BN][kRelStride] bf16

  int seq, m_blk, seq_start, seq_len;
  if (!locate_block<BM>(p.cu_seqlens, p.num_seqs, blockIdx.x, seq, m_blk, seq_start, seq_len)) return;
  const int head = blockIdx.y, kvh = head / p.G;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, t4 = lane & 3;
  const int q0 = m_blk * BM;                                   // first query row (within the sequence)
  const int n_blocks = CAUSAL ? min((seq_len + BN - 1) / BN, (q0 + BM + BN - 1) / BN) : (seq_len + BN - 1) / BN;

  const __nv_bfloat16* qbase = p.q + static_cast<long>(seq_start) * p.q_stride_t;
  const __nv_bfloat16* kbase = p.k + static_cast<long>(seq_start) * p.k_stride_t;
  const __nv_bfloat16* vbase = p.v + static_cast<long>(seq_start) * p.v_stride_t;

  load_tile_async<D, BM, 128>(sQ, qbase, p.q_stride_t, head, q0, seq_len);
  auto load_kv = [&](int nb, int buf) {
    load_tile_async<D, BN, 128>(sK + buf * BN * D * 2, kbase, p.k_stride_t, kvh, nb * BN, seq_len);
    load_tile_async<D, BN, 128>(sV + buf * BN * D * 2, vbase, p.v_stride_t, kvh, nb * BN, seq_len);
    if (REL_BIAS) {
      // The bucket index c(i-j) is monotone in (i-j) with slope <= 1, so a (64 x BN) score tile touches at
      // most 64+BN-1 consecutive table columns: stage only that window of A (query rows) and B (key rows).
      // (Windowing A as well as B keeps the CTA at ~70 KB of smem: 3 CTAs / SM instead of 2.)
      const int CH = kRelBW / 8;
      const int c_lo = rel_window_lo(p, q0, nb * BN, BN);
      const long tbase = (static_cast<long>(head) * p.total_tokens + seq_start) * p.NB;
      const __nv_bfloat16* abase = p.rel_a + tbase;
      const __nv_bfloat16* bbase = p.rel_b + tbase;
      uint8_t* dst_a = sA + static_cast<long>(buf) * BM * kRelStride * 2;
      uint8_t* dst_b = sB + static_cast<long>(buf) * BN * kRelStride * 2;
      for (int i = threadIdx.x; i < (BM + BN) * CH; i += 128) {
        const int r = i / CH, c = i % CH;
        const bool col_ok = c_lo + c * 8 < p.NB;
        if (r < BM) {
          const bool ok = ((q0 + r) < seq_len) && col_ok;
          cp_async_16_zfill(dst_a + (static_cast<long>(r) * kRelStride + c * 8) * 2,
                            abase + static_cast<long>(ok ? q0 + r : q0) * p.NB + (ok ? c_lo + c * 8 : 0), ok);
        } else {
          const int rb = r - BM;
          const bool ok = ((nb * BN + rb) < seq_len) && col_ok;
          cp_async_16_zfill(dst_b + (static_cast<long>(rb) * kRelStride + c * 8) * 2,
                            bbase + static_cast<long>(ok ? nb * BN + rb : nb * BN) * p.NB + (ok ? c_lo + c * 8 : 0), ok);
        }
      }
    }
  };
  load_kv(0, 0);
  cp_async_commit();

  float o[D / 8][4];
#pragma unroll
  for (int i = 0; i < D / 8; ++i) o[i][0] = o[i][1] = o[i][2] = o[i][3] = 0.f;
  float m_run[2] = {-INFINITY, -INFINITY}, l_run[2] = {0.f, 0.f};
  uint32_t qf[KS][4];
  const int row_a = q0 + warp * 16 + g, row_b = row_a + 8;     // the two query rows this thread owns

  for (int nb = 0; nb < n_blocks; ++nb) {
    const int buf = nb & 1;
    if (nb + 1 < n_blocks) load_kv(nb + 1, buf ^ 1);
    cp_async_commit();
    cp_async_wait<1>();
    __syncthreads();
    if (nb == 0) {
      const uint32_t qb = smem_u32(sQ);
      const int mrow = warp * 16 + (lane & 7) + ((lane >> 3) & 1) * 8, mcol = lane >> 4;
#pragma unroll
      for (int ks = 0; ks < KS; ++ks) ldsm_x4(qf[ks], qb + tile_off<D>(mrow, ks * 2 + mcol));
    }
    const uint32_t kb = smem_u32(sK + buf * BN * D * 2), vb = smem_u32(sV + buf * BN * D * 2);

    float s[NT][4];
#pragma unroll
    for (int i = 0; i < NT; ++i) s[i][0] = s[i][1] = s[i][2] = s[i][3] = 0.f;
#pragma unroll
    for (int ks = 0; ks < KS; ++ks) {
#pragma unroll
      for (int np = 0; np < NT / 2; ++np) {
        uint32_t kf[4];
        const int mrow = np * 16 + (lane & 7) + (lane >> 4) * 8, mcol = (lane >> 3) & 1;
        ldsm_x4(kf, kb + tile_off<D>(mrow, ks * 2 + mcol));
        mma16816(s[np * 2], qf[ks], kf[0], kf[1]);
        mma16816(s[np * 2 + 1], qf[ks], kf[2], kf[3]);
      }
    }
    // ---- scale, bias, mask ----
    const int key0 = nb * BN;
    if (REL_BIAS) {
      const int c_lo_cur = rel_window_lo(p, q0, key0, BN);
      const short* lut = p.bucket_lut + p.lut_center;
      const int dmin = max(q0 - (key0 + BN - 1), -p.lut_center), dmax = min(q0 + BM - 1 - key0, p.lut_center);
      const int c_min = lut[dmin], c_max = lut[dmax];
      const __nv_bfloat16* aw = reinterpret_cast<const __nv_bfloat16*>(sA + static_cast<long>(buf) * BM * kRelStride * 2);
      const __nv_bfloat16* bw = reinterpret_cast<const __nv_bfloat16*>(sB + static_cast<long>(buf) * BN * kRelStride * 2);
      const __nv_bfloat16* ar_a = aw + static_cast<long>(row_a - q0) * kRelStride;
      const __nv_bfloat16* ar_b = aw + static_cast<long>(row_b - q0) * kRelStride;
      if (c_min == c_max) {
        // far-from-diagonal tile: one bucket for every (i, j) -> bias = A[i, c0] + B[j, c0] (rank-1, no lookups)
        const int cw = c_min - c_lo_cur;
        const float a0 = __bfloat162float(ar_a[cw]), a1 = __bfloat162float(ar_b[cw]);
#pragma unroll
        for (int nt = 0; nt < NT; ++nt) {
          const int kl = nt * 8 + t4 * 2;
          const float b0 = __bfloat162float(bw[static_cast<long>(kl) * kRelStride + cw]);
          const float b1 = __bfloat162float(bw[static_cast<long>(kl + 1) * kRelStride + cw]);
          s[nt][0] += a0 + b0; s[nt][1] += a0 + b1; s[nt][2] += a1 + b0; s[nt][3] += a1 + b1;
        }
      } else {
        // General tile: one LUT read + two table reads per score.  Everything is 32-bit shared-memory addressing
        // with compile-time offsets; the thread's second row (row_a + 8) sees the deltas of the first row shifted
        // by one 8-key group, so its bucket indices are reused (10 LUT reads per thread instead of 16).
        // The LUT carries a 64-entry margin on both sides (build_bucket_lut), so padded-tail rows/keys stay in range;
        // their window offsets are clamped by the unsigned min below and their scores are masked.
        const short* lut_row = lut + (row_a - key0 - t4 * 2);
        uint32_t cw2[NT + 1][2];                      // 2 * (bucket - window start) of row_a at key group nt-1, parity e1
#pragma unroll
        for (int g8 = 0; g8 <= NT; ++g8)
#pragma unroll
          for (int e1 = 0; e1 < 2; ++e1) {
            const int c = lut_row[-((g8 - 1) * 8 + e1)];
            cw2[g8][e1] = min(static_cast<uint32_t>(c - c_lo_cur), static_cast<uint32_t>(kRelBW - 1)) * 2u;
          }
        const uint32_t sa_a = smem_u32(aw) + static_cast<uint32_t>(row_a - q0) * (kRelStride * 2);
        const uint32_t sa_b = sa_a + 8 * kRelStride * 2;
        const uint32_t sb0 = smem_u32(bw) + static_cast<uint32_t>(t4 * 2) * (kRelStride * 2);
#pragma unroll
        for (int nt = 0; nt < NT; ++nt)
#pragma unroll
          for (int e1 = 0; e1 < 2; ++e1) {
            const uint32_t brow = sb0 + (nt * 8 + e1) * (kRelStride * 2);
            const uint32_t wa = cw2[nt + 1][e1], wb = cw2[nt][e1];
            s[nt][e1] += lds_bf16(sa_a + wa) + lds_bf16(brow + wa);
            s[nt][2 + e1] += lds_bf16(sa_b + wb) + lds_bf16(brow + wb);
          }
      }
    }
    // ---- mask: only blocks that cross the end of the sequence (or the causal diagonal) have dead entries ----
    if ((key0 + BN > seq_len) || (CAUSAL && key0 + BN - 1 > q0)) {
#pragma unroll
      for (int nt = 0; nt < NT; ++nt)
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const int key = key0 + nt * 8 + t4 * 2 + (e & 1);
          const int row = (e < 2) ? row_a : row_b;
          if ((key >= seq_len) || (CAUSAL && key > row)) s[nt][e] = -INFINITY;
        }
    }
    // ---- online softmax (rows a = c0,c1 ; b = c2,c3); scores stay unscaled, the scale rides in the exp2 FMA ----
    float mx[2] = {-INFINITY, -INFINITY};
#pragma unroll
    for (int nt = 0; nt < NT; ++nt) {
      mx[0] = fmaxf(mx[0], fmaxf(s[nt][0], s[nt][1]));
      mx[1] = fmaxf(mx[1], fmaxf(s[nt][2], s[nt][3]));
    }
#pragma unroll
    for (int r = 0; r < 2; ++r) {
      mx[r] = fmaxf(mx[r], __shfl_xor_sync(0xffffffffu, mx[r], 1));
      mx[r] = fmaxf(mx[r], __shfl_xor_sync(0xffffffffu, mx[r], 2));
    }
    float corr[2], ps[2] = {0.f, 0.f};
#pragma unroll
    for (int r = 0; r < 2; ++r) {
      // key 0 is live for every row (also for the causal case), so mn is finite from the first block on
      const float mn = fmaxf(m_run[r], mx[r] * p.scale_log2);
      corr[r] = exp2f(m_run[r] - mn);
      m_run[r] = mn;
    }
#pragma unroll
    for (int nt = 0; nt < NT; ++nt)
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const int r = e >> 1;
        const float pv = exp2f(fmaf(s[nt][e], p.scale_log2, -m_run[r]));
        s[nt][e] = pv;
        ps[r] += pv;
      }
#pragma unroll
    for (int r = 0; r < 2; ++r) {
      ps[r] += __shfl_xor_sync(0xffffffffu, ps[r], 1);
      ps[r] += __shfl_xor_sync(0xffffffffu, ps[r], 2);
      l_run[r] = l_run[r] * corr[r] + ps[r];
    }
#pragma unroll
    for (int i = 0; i < D / 8; ++i) {
      o[i][0] *= corr[0]; o[i][1] *= corr[0]; o[i][2] *= corr[1]; o[i][3] *= corr[1];
    }
    // ---- O += P V ----
#pragma unroll
    for (int ks = 0; ks < BN / 16; ++ks) {
      uint32_t pa[4];
      pa[0] = pack_bf16x2(s[2 * ks][0], s[2 * ks][1]);
      pa[1] = pack_bf16x2(s[2 * ks][2], s[2 * ks][3]);
      pa[2] = pack_bf16x2(s[2 * ks + 1][0], s[2 * ks + 1][1]);
      pa[3] = pack_bf16x2(s[2 * ks + 1][2], s[2 * ks + 1][3]);
#pragma unroll
      for (int nd = 0; nd < D / 16; ++nd) {
        uint32_t vf[4];
        const int mrow = ks * 16 + (lane & 7) + ((lane >> 3) & 1) * 8, mcol = lane >> 4;
        ldsm_x4_t(vf, vb + tile_off<D>(mrow, nd * 2 + mcol));
        mma16816(o[nd * 2], pa, vf[0], vf[1]);
        mma16816(o[nd * 2 + 1], pa, vf[2], vf[3]);
      }
    }
    __syncthreads();
  }
  cp_async_wait<0>();

  // ---- epilogue ----
  const float inv_a = l_run[0] > 0.f ? 1.f / l_run[0] : 0.f, inv_b = l_run[1] > 0.f ? 1.f / l_run[1] : 0.f;
  __nv_bfloat16* obase = p.out + static_cast<long>(seq_start) * p.o_stride_t + head * D;
#pragma unroll
  for (int i = 0; i < D / 8; ++i) {
    const int col = i * 8 + t4 * 2;
    if (row_a < seq_len)
      *reinterpret_cast<uint32_t*>(obase + static_cast<long>(row_a) * p.o_stride_t + col) = pack_bf16x2(o[i][0] * inv_a, o[i][1] * inv_a);
    if (row_b < seq_len)
      *reinterpret_cast<uint32_t*>(obase + static_cast<long>(row_b) * p.o_stride_t + col) = pack_bf16x2(o[i][2] * inv_b, o[i][3] * inv_b);
  }
  if (p.lse != nullptr && t4 == 0) {
    const float ln2 = 0.6931471805599453f;
    float* lse = p.lse + static_cast<long>(head) * p.total_tokens + seq_start;
    if (row_a < seq_len) lse[row_a] = (m_run[0] + log2f(l_run[0])) * ln2;
    if (row_b < seq_len) lse[row_b] = (m_run[1] + log2f(l_run[1])) * ln2;
  }
}

// =================================================================================================
// DeBERTa disentangled attention, TMA-fed (K8).  Same math as flash_fwd_kernel<64, false, true, 32>, but every
// tile -- Q, K_j, V_j and the two sliding bias-table windows -- arrives by TMA behind one mbarrier per buffer, issued
// by a single thread.  In the cp.async version the per-thread address arithmetic of staging a (64+32) x 112 window
// every 32 keys was ~40 % of all executed instructions.
//   smem: Q 8 KB | K[2] 8 KB | V[2] 8 KB | A-window[2] 64 x 224 B | B-window[2] 32 x 224 B | barriers  (~67 KB, 3 CTAs/SM)
// =================================================================================================
template <int BN>
__global__ void __launch_bounds__(128)
deberta_attn_fwd_tma_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                            const __grid_constant__ CUtensorMap tmV, const __grid_constant__ CUtensorMap tmA,
                            const __grid_constant__ CUtensorMap tmB, AttnParams p) {
  constexpr int D = 64, BM = 64, KS = D / 16, NT = BN / 8;
  constexpr int kW = (BN == 32) ? 112 : 136;      // window columns: >= 64 + BN - 1 + 7, multiple of 8
  constexpr int kRelRowB = kW * 2;               // dense TMA box row (bytes)
  constexpr int kQ = 0, kK = BM * D * 2, kV = kK + 2 * BN * D * 2, kA = kV + 2 * BN * D * 2,
                kB = kA + 2 * BM * kRelRowB, kBar = kB + 2 * BN * kRelRowB;
  constexpr uint32_t kBlockTx = 2 * BN * D * 2 + (BM + BN) * kRelRowB;
  extern __shared__ __align__(1024) uint8_t smem[];
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem + kBar);      // [0] Q, [1..2] per-buffer block data

  int seq, m_blk, seq_start, seq_len;
  if (!locate_block<BM>(p.cu_seqlens, p.num_seqs, blockIdx.x, seq, m_blk, seq_start, seq_len)) return;
  const int head = blockIdx.y;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, t4 = lane & 3;
  const int q0 = m_blk * BM;
  const int n_blocks = (seq_len + BN - 1) / BN;
  const int tab_row0 = head * p.total_tokens + seq_start;           // row of this sequence in the [H*T, NB] tables

  if (threadIdx.x == 0) {
    mbar_init(&bar[0], 1); mbar_init(&bar[1], 1); mbar_init(&bar[2], 1);
    fence_mbar_init();
  }
  __syncthreads();
  auto issue_block = [&](int nb, int buf) {                         // thread 0 only
    const int key0 = nb * BN;
    const int c_lo = rel_window_lo(p, q0, key0, BN);
    uint64_t* b = &bar[1 + buf];
    mbar_arrive_expect_tx(b, kBlockTx);
    tma_load_2d(smem + kK + buf * BN * D * 2, &tmK, b, head * D, seq_start + key0);
    tma_load_2d(smem + kV + buf * BN * D * 2, &tmV, b, head * D, seq_start + key0);
    tma_load_2d(smem + kA + buf * BM * kRelRowB, &tmA, b, c_lo, tab_row0 + q0);
    tma_load_2d(smem + kB + buf * BN * kRelRowB, &tmB, b, c_lo, tab_row0 + key0);
  };
  if (threadIdx.x == 0) {
    mbar_arrive_expect_tx(&bar[0], BM * D * 2);
    tma_load_2d(smem + kQ, &tmQ, &bar[0], head * D, seq_start + q0);
    issue_block(0, 0);
  }

  float o[D / 8][4];
#pragma unroll
  for (int i = 0; i < D / 8; ++i) o[i][0] = o[i][1] = o[i][2] = o[i][3] = 0.f;
  float m_run[2] = {-INFINITY, -INFINITY}, l_run[2] = {0.f, 0.f};
  uint32_t qf[KS][4];
  const int row_a = q0 + warp * 16 + g, row_b = row_a + 8;
  const short* lut = p.bucket_lut + p.lut_center;

  mbar_wait(&bar[0], 0);
  {
    const uint32_t qb = smem_u32(smem + kQ);
    const int mrow = warp * 16 + (lane & 7) + ((lane >> 3) & 1) * 8, mcol = lane >> 4;
#pragma unroll
    for (int ks = 0; ks < KS; ++ks) ldsm_x4(qf[ks], qb + tile_off<D>(mrow, ks * 2 + mcol));
  }

  for (int nb = 0; nb < n_blocks; ++nb) {
    const int buf = nb & 1;
    // buffer buf^1 was last read in iteration nb-1, which ended with __syncthreads()
    if (threadIdx.x == 0 && nb + 1 < n_blocks) issue_block(nb + 1, buf ^ 1);
    mbar_wait(&bar[1 + buf], (nb >> 1) & 1);
    const uint32_t kb = smem_u32(smem + kK + buf * BN * D * 2), vb = smem_u32(smem + kV + buf * BN * D * 2);

    float s[NT][4];
#pragma unroll
    for (int i = 0; i < NT; ++i) s[i][0] = s[i][1] = s[i][2] = s[i][3] = 0.f;
#pragma unroll
    for (int ks = 0; ks < KS; ++ks) {
#pragma unroll
      for (int np = 0; np < NT / 2; ++np) {
        uint32_t kf[4];
        const int mrow = np * 16 + (lane & 7) + (lane >> 4) * 8, mcol = (lane >> 3) & 1;
        ldsm_x4(kf, kb + tile_off<D>(mrow, ks * 2 + mcol));
        mma16816(s[np * 2], qf[ks], kf[0], kf[1]);
        mma16816(s[np * 2 + 1], qf[ks], kf[2], kf[3]);
      }
    }
    // ---- bias: A[i, c(i-j)] + B[j, c(i-j)] from the two windows ----
    const int key0 = nb * BN;
    {
      const int c_lo_cur = rel_window_lo(p, q0, key0, BN);
      const int dmin = max(q0 - (key0 + BN - 1), -p.lut_center), dmax = min(q0 + BM - 1 - key0, p.lut_center);
      const int c_min = lut[dmin], c_max = lut[dmax];
      const uint32_t sa_a = smem_u32(smem + kA + buf * BM * kRelRowB) + static_cast<uint32_t>(row_a - q0) * kRelRowB;
      const uint32_t sa_b = sa_a + 8 * kRelRowB;
      const uint32_t sb0 = smem_u32(smem + kB + buf * BN * kRelRowB) + static_cast<uint32_t>(t4 * 2) * kRelRowB;
      if (c_min == c_max) {
        // far-from-diagonal tile: one bucket for every (i, j) -> bias = A[i, c0] + B[j, c0] (rank-1, no lookups)
        const uint32_t cw = static_cast<uint32_t>(c_min - c_lo_cur) * 2u;
        const float a0 = lds_bf16(sa_a + cw), a1 = lds_bf16(sa_b + cw);
#pragma unroll
        for (int nt = 0; nt < NT; ++nt) {
          const float b0 = lds_bf16(sb0 + (nt * 8) * kRelRowB + cw), b1 = lds_bf16(sb0 + (nt * 8 + 1) * kRelRowB + cw);
          s[nt][0] += a0 + b0; s[nt][1] += a0 + b1; s[nt][2] += a1 + b0; s[nt][3] += a1 + b1;
        }
      } else {
        // general tile (see flash_fwd_kernel): row_a + 8 reuses row_a's bucket indices one key group later
        const short* lut_row = lut + (row_a - key0 - t4 * 2);
        uint32_t cw2[NT + 1][2];
#pragma unroll
        for (int g8 = 0; g8 <= NT; ++g8)
#pragma unroll
          for (int e1 = 0; e1 < 2; ++e1) {
            const int c = lut_row[-((g8 - 1) * 8 + e1)];
            cw2[g8][e1] = min(static_cast<uint32_t>(c - c_lo_cur), static_cast<uint32_t>(kW - 1)) * 2u;
          }
#pragma unroll
        for (int nt = 0; nt < NT; ++nt)
#pragma unroll
          for (int e1 = 0; e1 < 2; ++e1) {
            const uint32_t brow = sb0 + (nt * 8 + e1) * kRelRowB;
            const uint32_t wa = cw2[nt + 1][e1], wb = cw2[nt][e1];
            s[nt][e1] += lds_bf16(sa_a + wa) + lds_bf16(brow + wa);
            s[nt][2 + e1] += lds_bf16(sa_b + wb) + lds_bf16(brow + wb);
          }
      }
    }
    if (key0 + BN > seq_len) {
#pragma unroll
      for (int nt = 0; nt < NT; ++nt)
#pragma unroll
        for (int e = 0; e < 4; ++e)
          if (key0 + nt * 8 + t4 * 2 + (e & 1) >= seq_len) s[nt][e] = -INFINITY;
    }
    // ---- online softmax (unscaled scores, the scale rides in the exp2 FMA) ----
    float mx[2] = {-INFINITY, -INFINITY};
#pragma unroll
    for (int nt = 0; nt < NT; ++nt) {
      mx[0] = fmaxf(mx[0], fmaxf(s[nt][0], s[nt][1]));
      mx[1] = fmaxf(mx[1], fmaxf(s[nt][2], s[nt][3]));
    }
#pragma unroll
    for (int r = 0; r < 2; ++r) {
      mx[r] = fmaxf(mx[r], __shfl_xor_sync(0xffffffffu, mx[r], 1));
      mx[r] = fmaxf(mx[r], __shfl_xor_sync(0xffffffffu, mx[r], 2));
    }
    float corr[2], ps[2] = {0.f, 0.f};
#pragma unroll
    for (int r = 0; r < 2; ++r) {
      const float mn = fmaxf(m_run[r], mx[r] * p.scale_log2);      // key 0 is always live: finite from block 0 on
      corr[r] = exp2f(m_run[r] - mn);
      m_run[r] = mn;
    }
#pragma unroll
    for (int nt = 0; nt < NT; ++nt)
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const float pv = exp2f(fmaf(s[nt][e], p.scale_log2, -m_run[e >> 1]));
        s[nt][e] = pv;
        ps[e >> 1] += pv;
      }
#pragma unroll
    for (int r = 0; r < 2; ++r) {
      ps[r] += __shfl_xor_sync(0xffffffffu, ps[r], 1);
      ps[r] += __shfl_xor_sync(0xffffffffu, ps[r], 2);
      l_run[r] = l_run[r] * corr[r] + ps[r];
    }
#pragma unroll
    for (int i = 0; i < D / 8; ++i) {
      o[i][0] *= corr[0]; o[i][1] *= corr[0]; o[i][2] *= corr[1]; o[i][3] *= corr[1];
    }
    // ---- O += P V ----
#pragma unroll
    for (int ks = 0; ks < BN / 16; ++ks) {
      uint32_t pa[4];
      pa[0] = pack_bf16x2(s[2 * ks][0], s[2 * ks][1]);
      pa[1] = pack_bf16x2(s[2 * ks][2], s[2 * ks][3]);
      pa[2] = pack_bf16x2(s[2 * ks + 1][0], s[2 * ks + 1][1]);
      pa[3] = pack_bf16x2(s[2 * ks + 1][2], s[2 * ks + 1][3]);
#pragma unroll
      for (int nd = 0; nd < D / 16; ++nd) {
        uint32_t vf[4];
        const int mrow = ks * 16 + (lane & 7) + ((lane >> 3) & 1) * 8, mcol = lane >> 4;
        ldsm_x4_t(vf, vb + tile_off<D>(mrow, nd * 2 + mcol));
        mma16816(o[nd * 2], pa, vf[0], vf[1]);
        mma16816(o[nd * 2 + 1], pa, vf[2], vf[3]);
      }
    }
    __syncthreads();
  }

  const float inv_a = l_run[0] > 0.f ? 1.f / l_run[0] : 0.f, inv_b = l_run[1] > 0.f ? 1.f / l_run[1] : 0.f;
  __nv_bfloat16* obase = p.out + static_cast<long>(seq_start) * p.o_stride_t + head * D;
#pragma unroll
  for (int i = 0; i < D / 8; ++i) {
    const int col = i * 8 + t4 * 2;
    if (row_a < seq_len)
      *reinterpret_cast<uint32_t*>(obase + static_cast<long>(row_a) * p.o_stride_t + col) = pack_bf16x2(o[i][0] * inv_a, o[i][1] * inv_a);
    if (row_b < seq_len)
      *reinterpret_cast<uint32_t*>(obase + static_cast<long>(row_b) * p.o_stride_t + col) = pack_bf16x2(o[i][2] * inv_b, o[i][3] * inv_b);
  }
  if (p.lse != nullptr && t4 == 0) {
    const float ln2 = 0.6931471805599453f;
    float* lse = p.lse + static_cast<long>(head) * p.total_tokens + seq_start;
    if (row_a < seq_len) lse[row_a] = (m_run[0] + log2f(l_run[0])) * ln2;
    if (row_b < seq_len) lse[row_b] = (m_run[1] + log2f(l_run[1])) * ln2;
  }
}

// Causal forward with TWO 16-row MMA tiles per warp (BM = 128): every K / V fragment fetched with ldmatrix feeds two
// MMAs instead of one, which moves the kernel from shared-memory-bandwidth bound to tensor-pipe bound.
template <int D, int BN>
__global__ void __launch_bounds__(128) flash_fwd_causal_mt2_kernel(AttnParams p) {
  constexpr int BM = 128, KS = D / 16, NT = BN / 8, MT = 2;
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* sQ = smem;                              // [128][D]
  uint8_t* sK = sQ + BM * D * 2;                   // [2][BN][D]
  uint8_t* sV = sK + 2 * BN * D * 2;               // [2][BN][D]

  int seq, m_blk, seq_start, seq_len;
  if (!locate_block<BM>(p.cu_seqlens, p.num_seqs, blockIdx.x, seq, m_blk, seq_start, seq_len)) return;
  const int head = blockIdx.y, kvh = head / p.G;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, t4 = lane & 3;
  const int q0 = m_blk * BM;
  const int n_blocks = min((seq_len + BN - 1) / BN, (q0 + BM + BN - 1) / BN);
  const __nv_bfloat16* qbase = p.q + static_cast<long>(seq_start) * p.q_stride_t;
  const __nv_bfloat16* kbase = p.k + static_cast<long>(seq_start) * p.k_stride_t;
  const __nv_bfloat16* vbase = p.v + static_cast<long>(seq_start) * p.v_stride_t;

  load_tile_async<D, BM, 128>(sQ, qbase, p.q_stride_t, head, q0, seq_len);
  auto load_kv = [&](int nb, int buf) {
    load_tile_async<D, BN, 128>(sK + buf * BN * D * 2, kbase, p.k_stride_t, kvh, nb * BN, seq_len);
    load_tile_async<D, BN, 128>(sV + buf * BN * D * 2, vbase, p.v_stride_t, kvh, nb * BN, seq_len);
  };
  load_kv(0, 0);
  cp_async_commit();

  float o[MT][D / 8][4];
  float m_run[MT][2], l_run[MT][2];
#pragma unroll
  for (int mt = 0; mt < MT; ++mt) {
#pragma unroll
    for (int i = 0; i < D / 8; ++i) o[mt][i][0] = o[mt][i][1] = o[mt][i][2] = o[mt][i][3] = 0.f;
    m_run[mt][0] = m_run[mt][1] = -INFINITY;
    l_run[mt][0] = l_run[mt][1] = 0.f;
  }
  const int wrow0 = q0 + warp * 32;                 // first query row of this warp
  const uint32_t qb = smem_u32(sQ);

  for (int nb = 0; nb < n_blocks; ++nb) {
    const int buf = nb & 1;
    if (nb + 1 < n_blocks) load_kv(nb + 1, buf ^ 1);
    cp_async_commit();
    cp_async_wait<1>();
    __syncthreads();
    const int key0 = nb * BN;
    if (key0 <= wrow0 + 31) {                        // causal: later key blocks are fully masked for this warp
      const uint32_t kb = smem_u32(sK + buf * BN * D * 2), vb = smem_u32(sV + buf * BN * D * 2);
      float s[MT][NT][4];
#pragma unroll
      for (int mt = 0; mt < MT; ++mt)
#pragma unroll
        for (int i = 0; i < NT; ++i) s[mt][i][0] = s[mt][i][1] = s[mt][i][2] = s[mt][i][3] = 0.f;
#pragma unroll
      for (int ks = 0; ks < KS; ++ks) {
        uint32_t qf[MT][4];
#pragma unroll
        for (int mt = 0; mt < MT; ++mt) {
          const int mrow = warp * 32 + mt * 16 + (lane & 7) + ((lane >> 3) & 1) * 8, mcol = lane >> 4;
          ldsm_x4(qf[mt], qb + tile_off<D>(mrow, ks * 2 + mcol));
        }
#pragma unroll
        for (int np = 0; np < NT / 2; ++np) {
          uint32_t kf[4];
          const int mrow = np * 16 + (lane & 7) + (lane >> 4) * 8, mcol = (lane >> 3) & 1;
          ldsm_x4(kf, kb + tile_off<D>(mrow, ks * 2 + mcol));
#pragma unroll
          for (int mt = 0; mt < MT; ++mt) {
            mma16816(s[mt][np * 2], qf[mt], kf[0], kf[1]);
            mma16816(s[mt][np * 2 + 1], qf[mt], kf[2], kf[3]);
          }
        }
      }
#pragma unroll
      for (int mt = 0; mt < MT; ++mt) {
        const int row_a = wrow0 + mt * 16 + g, row_b = row_a + 8;
        float mx[2] = {-INFINITY, -INFINITY};
#pragma unroll
        for (int nt = 0; nt < NT; ++nt)
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const int key = key0 + nt * 8 + t4 * 2 + (e & 1);
            const int row = (e < 2) ? row_a : row_b;
            const bool dead = (key >= seq_len) || (key > row);
            const float x = dead ? -INFINITY : s[mt][nt][e] * p.scale_log2;
            s[mt][nt][e] = x;
            mx[e >> 1] = fmaxf(mx[e >> 1], x);
          }
        float corr[2], ps[2] = {0.f, 0.f};
#pragma unroll
        for (int r = 0; r < 2; ++r) {
          mx[r] = fmaxf(mx[r], __shfl_xor_sync(0xffffffffu, mx[r], 1));
          mx[r] = fmaxf(mx[r], __shfl_xor_sync(0xffffffffu, mx[r], 2));
          const float mn = fmaxf(m_run[mt][r], mx[r]);
          corr[r] = (mn == -INFINITY) ? 1.f : exp2f(m_run[mt][r] - mn);
          m_run[mt][r] = mn;
        }
#pragma unroll
        for (int nt = 0; nt < NT; ++nt)
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const int r = e >> 1;
            const float pv = (m_run[mt][r] == -INFINITY) ? 0.f : exp2f(s[mt][nt][e] - m_run[mt][r]);
            s[mt][nt][e] = pv;
            ps[r] += pv;
          }
#pragma unroll
        for (int r = 0; r < 2; ++r) {
          ps[r] += __shfl_xor_sync(0xffffffffu, ps[r], 1);
          ps[r] += __shfl_xor_sync(0xffffffffu, ps[r], 2);
          l_run[mt][r] = l_run[mt][r] * corr[r] + ps[r];
        }
#pragma unroll
        for (int i = 0; i < D / 8; ++i) {
          o[mt][i][0] *= corr[0]; o[mt][i][1] *= corr[0]; o[mt][i][2] *= corr[1]; o[mt][i][3] *= corr[1];
        }
      }
#pragma unroll
      for (int ks = 0; ks < BN / 16; ++ks) {
        uint32_t pa[MT][4];
#pragma unroll
        for (int mt = 0; mt < MT; ++mt) {
          pa[mt][0] = pack_bf16x2(s[mt][2 * ks][0], s[mt][2 * ks][1]);
          pa[mt][1] = pack_bf16x2(s[mt][2 * ks][2], s[mt][2 * ks][3]);
          pa[mt][2] = pack_bf16x2(s[mt][2 * ks + 1][0], s[mt][2 * ks + 1][1]);
          pa[mt][3] = pack_bf16x2(s[mt][2 * ks + 1][2], s[mt][2 * ks + 1][3]);
        }
#pragma unroll
        for (int nd = 0; nd < D / 16; ++nd) {
          uint32_t vf[4];
          const int mrow = ks * 16 + (lane & 7) + ((lane >> 3) & 1) * 8, mcol = lane >> 4;
          ldsm_x4_t(vf, vb + tile_off<D>(mrow, nd * 2 + mcol));
#pragma unroll
          for (int mt = 0; mt < MT; ++mt) {
            mma16816(o[mt][nd * 2], pa[mt], vf[0], vf[1]);
            mma16816(o[mt][nd * 2 + 1], pa[mt], vf[2], vf[3]);
          }
        }
      }
    }
    __syncthreads();
  }
  cp_async_wait<0>();

  __nv_bfloat16* obase = p.out + static_cast<long>(seq_start) * p.o_stride_t + head * D;
  const float ln2 = 0.6931471805599453f;
#pragma unroll
  for (int mt = 0; mt < MT; ++mt) {
    const int row_a = wrow0 + mt * 16 + g, row_b = row_a + 8;
    const float inv_a = l_run[mt][0] > 0.f ? 1.f / l_run[mt][0] : 0.f, inv_b = l_run[mt][1] > 0.f ? 1.f / l_run[mt][1] : 0.f;
#pragma unroll
    for (int i = 0; i < D / 8; ++i) {
      const int col = i * 8 + t4 * 2;
      if (row_a < seq_len)
        *reinterpret_cast<uint32_t*>(obase + static_cast<long>(row_a) * p.o_stride_t + col) = pack_bf16x2(o[mt][i][0] * inv_a, o[mt][i][1] * inv_a);
      if (row_b < seq_len)
        *reinterpret_cast<uint32_t*>(obase + static_cast<long>(row_b) * p.o_stride_t + col) = pack_bf16x2(o[mt][i][2] * inv_b, o[mt][i][3] * inv_b);
    }
    if (p.lse != nullptr && t4 == 0) {
      float* lse = p.lse + static_cast<long>(head) * p.total_tokens + seq_start;
      if (row_a < seq_len) lse[row_a] = (m_run[mt][0] + log2f(l_run[mt][0])) * ln2;
      if (row_b < seq_len) lse[row_b] = (m_run[mt][1] + log2f(l_run[mt][1])) * ln2;
    }
  }
}

// =================================================================================================
// backward
// =================================================================================================
// delta[h, t] = sum_d dO[t,h,d] * O[t,h,d]
__global__ void attn_bwd_delta_kernel(AttnParams p, int D) {
  const long idx = blockIdx.x * static_cast<long>(blockDim.x >> 5) + (threadIdx.x >> 5);   // (t, h) pair per warp
  const long total = static_cast<long>(p.total_tokens) * p.Hq;
  if (idx >= total) return;
  const int lane = threadIdx.x & 31;
  const long t = idx / p.Hq;
  const int h = idx % p.Hq;
  const __nv_bfloat16* o = p.o + t * p.o_stride_t + h * D;
  const __nv_bfloat16* d = p.dout + t * p.o_stride_t + h * D;
  float acc = 0.f;
  for (int c = lane * 2; c < D; c += 64) {
    float2 a = unpack_bf16x2(*reinterpret_cast<const uint32_t*>(o + c));
    float2 b = unpack_bf16x2(*reinterpret_cast<const uint32_t*>(d + c));
    acc += a.x * b.x + a.y * b.y;
  }
  acc = warp_sum(acc);
  if (lane == 0) p.delta[static_cast<long>(h) * p.total_tokens + t] = acc;
}

// dK / dV: one CTA per (kv block of 64 keys, kv head); 4 warps x 16 keys; loops the G q-heads and the query blocks.
// Works on S^T = K Q^T so that the key dimension is the MMA M dimension and dK/dV accumulate in registers.
template <int D>
__global__ void __launch_bounds__(128) flash_bwd_dkdv_kernel(AttnParams p) {
  constexpr int BNK = 64, BQ = 32, KS = D / 16;
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* sK = smem;                        // [64][D]
  uint8_t* sV = sK + BNK * D * 2;            // [64][D]
  uint8_t* sQ = sV + BNK * D * 2;            // [2][BQ][D]
  uint8_t* sdO = sQ + 2 * BQ * D * 2;        // [2][BQ][D]
  float* sLse = reinterpret_cast<float*>(sdO + 2 * BQ * D * 2);   // [2][BQ]
  float* sDelta = sLse + 2 * BQ;                                   // [2][BQ]

  int seq, n_blk, seq_start, seq_len;
  if (!locate_block<BNK>(p.cu_seqlens, p.num_seqs, blockIdx.x, seq, n_blk, seq_start, seq_len)) return;
  const int kvh = blockIdx.y;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, t4 = lane & 3;
  const int k0 = n_blk * BNK;
  const __nv_bfloat16* qbase = p.q + static_cast<long>(seq_start) * p.q_stride_t;
  const __nv_bfloat16* kbase = p.k + static_cast<long>(seq_start) * p.k_stride_t;
  const __nv_bfloat16* vbase = p.v + static_cast<long>(seq_start) * p.v_stride_t;
  const __nv_bfloat16* dobase = p.dout + static_cast<long>(seq_start) * p.o_stride_t;

  load_tile_async<D, BNK, 128>(sK, kbase, p.k_stride_t, kvh, k0, seq_len);
  load_tile_async<D, BNK, 128>(sV, vbase, p.v_stride_t, kvh, k0, seq_len);
  cp_async_commit();

  float dk[D / 8][4], dv[D / 8][4];
#pragma unroll
  for (int i = 0; i < D / 8; ++i) {
    dk[i][0] = dk[i][1] = dk[i][2] = dk[i][3] = 0.f;
    dv[i][0] = dv[i][1] = dv[i][2] = dv[i][3] = 0.f;
  }
  const int q_first_blk = k0 / BQ;                               // causal: queries >= first key of the block
  const int q_blocks = (seq_len + BQ - 1) / BQ;
  const int iters_per_head = q_blocks - q_first_blk;
  const int total_iters = iters_per_head * p.G;

  auto load_q = [&](int it, int buf) {
    const int head = kvh * p.G + it / iters_per_head;
    const int qb = q_first_blk + it % iters_per_head;
    load_tile_async<D, BQ, 128>(sQ + buf * BQ * D * 2, qbase, p.q_stride_t, head, qb * BQ, seq_len);
    load_tile_async<D, BQ, 128>(sdO + buf * BQ * D * 2, dobase, p.o_stride_t, head, qb * BQ, seq_len);
    if (threadIdx.x < BQ) {
      const int r = qb * BQ + threadIdx.x;
      const long off = static_cast<long>(head) * p.total_tokens + seq_start + r;
      sLse[buf * BQ + threadIdx.x] = (r < seq_len) ? p.lse[off] : INFINITY;
      sDelta[buf * BQ + threadIdx.x] = (r < seq_len) ? p.delta[off] : 0.f;
    }
  };
  if (total_iters > 0) load_q(0, 0);
  cp_async_commit();
  const float log2e = 1.4426950408889634f;
  const int key_a = k0 + warp * 16 + g, key_b = key_a + 8;

  for (int it = 0; it < total_iters; ++it) {
    const int buf = it & 1;
    if (it + 1 < total_iters) load_q(it + 1, buf ^ 1);
    cp_async_commit();
    cp_async_wait<1>();
    __syncthreads();
    const int qb = q_first_blk + it % iters_per_head;
    const int qrow0 = qb * BQ;
    const uint32_t kaddr = smem_u32(sK), vaddr = smem_u32(sV);
    const uint32_t qaddr = smem_u32(sQ + buf * BQ * D * 2), doaddr = smem_u32(sdO + buf * BQ * D * 2);

    // S^T[16 keys x 32 queries] = K Q^T ; dP^T = V dO^T
    float st[BQ / 8][4], dpt[BQ / 8][4];
#pragma unroll
    for (int i = 0; i < BQ / 8; ++i) {
      st[i][0] = st[i][1] = st[i][2] = st[i][3] = 0.f;
      dpt[i][0] = dpt[i][1] = dpt[i][2] = dpt[i][3] = 0.f;
    }
#pragma unroll
    for (int ks = 0; ks < KS; ++ks) {
      uint32_t ka[4], va[4];
      const int arow = warp * 16 + (lane & 7) + ((lane >> 3) & 1) * 8, acol = lane >> 4;
      ldsm_x4(ka, kaddr + tile_off<D>(arow, ks * 2 + acol));
      ldsm_x4(va, vaddr + tile_off<D>(arow, ks * 2 + acol));
#pragma unroll
      for (int np = 0; np < BQ / 16; ++np) {
        uint32_t qf[4], df[4];
        const int brow = np * 16 + (lane & 7) + (lane >> 4) * 8, bcol = (lane >> 3) & 1;
        ldsm_x4(qf, qaddr + tile_off<D>(brow, ks * 2 + bcol));
        ldsm_x4(df, doaddr + tile_off<D>(brow, ks * 2 + bcol));
        mma16816(st[np * 2], ka, qf[0], qf[1]);
        mma16816(st[np * 2 + 1], ka, qf[2], qf[3]);
        mma16816(dpt[np * 2], va, df[0], df[1]);
        mma16816(dpt[np * 2 + 1], va, df[2], df[3]);
      }
    }
    // P^T = exp(S^T*scale - lse[q]) ; dS^T = P^T * (dP^T - delta[q]) * scale
#pragma unroll
    for (int nt = 0; nt < BQ / 8; ++nt)
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const int ql = nt * 8 + t4 * 2 + (e & 1);
        const int qrow = qrow0 + ql;
        const int key = (e < 2) ? key_a : key_b;
        const bool dead = (key >= seq_len) || (qrow >= seq_len) || (key > qrow);
        const float pv = dead ? 0.f : exp2f(st[nt][e] * p.scale_log2 - sLse[buf * BQ + ql] * log2e);
        const float ds = pv * (dpt[nt][e] - sDelta[buf * BQ + ql]) * p.scale;
        st[nt][e] = pv;
        dpt[nt][e] = ds;
      }
    // dV += P^T dO ; dK += dS^T Q   (A operands from registers, B = dO / Q with queries as the K dimension)
#pragma unroll
    for (int ks = 0; ks < BQ / 16; ++ks) {
      uint32_t pa[4], da[4];
      pa[0] = pack_bf16x2(st[2 * ks][0], st[2 * ks][1]);
      pa[1] = pack_bf16x2(st[2 * ks][2], st[2 * ks][3]);
      pa[2] = pack_bf16x2(st[2 * ks + 1][0], st[2 * ks + 1][1]);
      pa[3] = pack_bf16x2(st[2 * ks + 1][2], st[2 * ks + 1][3]);
      da[0] = pack_bf16x2(dpt[2 * ks][0], dpt[2 * ks][1]);
      da[1] = pack_bf16x2(dpt[2 * ks][2], dpt[2 * ks][3]);
      da[2] = pack_bf16x2(dpt[2 * ks + 1][0], dpt[2 * ks + 1][1]);
      da[3] = pack_bf16x2(dpt[2 * ks + 1][2], dpt[2 * ks + 1][3]);
#pragma unroll
      for (int nd = 0; nd < D / 16; ++nd) {
        uint32_t dof[4], qf[4];
        const int mrow = ks * 16 + (lane & 7) + ((lane >> 3) & 1) * 8, mcol = lane >> 4;
        ldsm_x4_t(dof, doaddr + tile_off<D>(mrow, nd * 2 + mcol));
        ldsm_x4_t(qf, qaddr + tile_off<D>(mrow, nd * 2 + mcol));
        mma16816(dv[nd * 2], pa, dof[0], dof[1]);
        mma16816(dv[nd * 2 + 1], pa, dof[2], dof[3]);
        mma16816(dk[nd * 2], da, qf[0], qf[1]);
        mma16816(dk[nd * 2 + 1], da, qf[2], qf[3]);
      }
    }
    __syncthreads();
  }
  cp_async_wait<0>();
  __nv_bfloat16* dkb = p.dk + static_cast<long>(seq_start) * p.k_stride_t + kvh * D;
  __nv_bfloat16* dvb = p.dv + static_cast<long>(seq_start) * p.v_stride_t + kvh * D;
#pragma unroll
  for (int i = 0; i < D / 8; ++i) {
    const int col = i * 8 + t4 * 2;
    if (key_a < seq_len) {
      *reinterpret_cast<uint32_t*>(dkb + static_cast<long>(key_a) * p.k_stride_t + col) = pack_bf16x2(dk[i][0], dk[i][1]);
      *reinterpret_cast<uint32_t*>(dvb + static_cast<long>(key_a) * p.v_stride_t + col) = pack_bf16x2(dv[i][0], dv[i][1]);
    }
    if (key_b < seq_len) {
      *reinterpret_cast<uint32_t*>(dkb + static_cast<long>(key_b) * p.k_stride_t + col) = pack_bf16x2(dk[i][2], dk[i][3]);
      *reinterpret_cast<uint32_t*>(dvb + static_cast<long>(key_b) * p.v_stride_t + col) = pack_bf16x2(dv[i][2], dv[i][3]);
    }
  }
}

// dQ: one CTA per (query block of 64, q head); loops kv blocks 0..diag.
template <int D>
__global__ void __launch_bounds__(128) flash_bwd_dq_kernel(AttnParams p) {
  constexpr int BM = 64, BN = 32, KS = D / 16;
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* sQ = smem;                       // [64][D]
  uint8_t* sdO = sQ + BM * D * 2;           // [64][D]
  uint8_t* sK = sdO + BM * D * 2;           // [2][BN][D]
  uint8_t* sV = sK + 2 * BN * D * 2;        // [2][BN][D]

  int seq, m_blk, seq_start, seq_len;
  if (!locate_block<BM>(p.cu_seqlens, p.num_seqs, blockIdx.x, seq, m_blk, seq_start, seq_len)) return;
  const int head = blockIdx.y, kvh = head / p.G;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, t4 = lane & 3;
  const int q0 = m_blk * BM;
  const int n_blocks = min((seq_len + BN - 1) / BN, (q0 + BM + BN - 1) / BN);
  const __nv_bfloat16* qbase = p.q + static_cast<long>(seq_start) * p.q_stride_t;
  const __nv_bfloat16* kbase = p.k + static_cast<long>(seq_start) * p.k_stride_t;
  const __nv_bfloat16* vbase = p.v + static_cast<long>(seq_start) * p.v_stride_t;
  const __nv_bfloat16* dobase = p.dout + static_cast<long>(seq_start) * p.o_stride_t;

  load_tile_async<D, BM, 128>(sQ, qbase, p.q_stride_t, head, q0, seq_len);
  load_tile_async<D, BM, 128>(sdO, dobase, p.o_stride_t, head, q0, seq_len);
  auto load_kv = [&](int nb, int buf) {
    load_tile_async<D, BN, 128>(sK + buf * BN * D * 2, kbase, p.k_stride_t, kvh, nb * BN, seq_len);
    load_tile_async<D, BN, 128>(sV + buf * BN * D * 2, vbase, p.v_stride_t, kvh, nb * BN, seq_len);
  };
  load_kv(0, 0);
  cp_async_commit();

  const int row_a = q0 + warp * 16 + g, row_b = row_a + 8;
  const long loff = static_cast<long>(head) * p.total_tokens + seq_start;
  const float log2e = 1.4426950408889634f;
  const float lse_a = (row_a < seq_len) ? p.lse[loff + row_a] * log2e : INFINITY;
  const float lse_b = (row_b < seq_len) ? p.lse[loff + row_b] * log2e : INFINITY;
  const float del_a = (row_a < seq_len) ? p.delta[loff + row_a] : 0.f;
  const float del_b = (row_b < seq_len) ? p.delta[loff + row_b] : 0.f;

  float dq[D / 8][4];
#pragma unroll
  for (int i = 0; i < D / 8; ++i) dq[i][0] = dq[i][1] = dq[i][2] = dq[i][3] = 0.f;

  for (int nb = 0; nb < n_blocks; ++nb) {
    const int buf = nb & 1;
    if (nb + 1 < n_blocks) load_kv(nb + 1, buf ^ 1);
    cp_async_commit();
    cp_async_wait<1>();
    __syncthreads();
    const uint32_t qaddr = smem_u32(sQ), doaddr = smem_u32(sdO);
    const uint32_t kaddr = smem_u32(sK + buf * BN * D * 2), vaddr = smem_u32(sV + buf * BN * D * 2);
    float s[BN / 8][4], dp[BN / 8][4];
#pragma unroll
    for (int i = 0; i < BN / 8; ++i) {
      s[i][0] = s[i][1] = s[i][2] = s[i][3] = 0.f;
      dp[i][0] = dp[i][1] = dp[i][2] = dp[i][3] = 0.f;
    }
#pragma unroll
    for (int ks = 0; ks < KS; ++ks) {
      uint32_t qa[4], da[4];
      const int arow = warp * 16 + (lane & 7) + ((lane >> 3) & 1) * 8, acol = lane >> 4;
      ldsm_x4(qa, qaddr + tile_off<D>(arow, ks * 2 + acol));
      ldsm_x4(da, doaddr + tile_off<D>(arow, ks * 2 + acol));
#pragma unroll
      for (int np = 0; np < BN / 16; ++np) {
        uint32_t kf[4], vf[4];
        const int brow = np * 16 + (lane & 7) + (lane >> 4) * 8, bcol = (lane >> 3) & 1;
        ldsm_x4(kf, kaddr + tile_off<D>(brow, ks * 2 + bcol));
        ldsm_x4(vf, vaddr + tile_off<D>(brow, ks * 2 + bcol));
        mma16816(s[np * 2], qa, kf[0], kf[1]);
        mma16816(s[np * 2 + 1], qa, kf[2], kf[3]);
        mma16816(dp[np * 2], da, vf[0], vf[1]);
        mma16816(dp[np * 2 + 1], da, vf[2], vf[3]);
      }
    }
    const int key0 = nb * BN;
#pragma unroll
    for (int nt = 0; nt < BN / 8; ++nt)
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const int key = key0 + nt * 8 + t4 * 2 + (e & 1);
        const int row = (e < 2) ? row_a : row_b;
        const bool dead = (key >= seq_len) || (row >= seq_len) || (key > row);
        const float pv = dead ? 0.f : exp2f(s[nt][e] * p.scale_log2 - ((e < 2) ? lse_a : lse_b));
        s[nt][e] = pv * (dp[nt][e] - ((e < 2) ? del_a : del_b)) * p.scale;     // dS
      }
    // dQ += dS K  (B = K with keys as the contraction dim -> transposed ldmatrix)
#pragma unroll
    for (int ks = 0; ks < BN / 16; ++ks) {
      uint32_t sa[4];
      sa[0] = pack_bf16x2(s[2 * ks][0], s[2 * ks][1]);
      sa[1] = pack_bf16x2(s[2 * ks][2], s[2 * ks][3]);
      sa[2] = pack_bf16x2(s[2 * ks + 1][0], s[2 * ks + 1][1]);
      sa[3] = pack_bf16x2(s[2 * ks + 1][2], s[2 * ks + 1][3]);
#pragma unroll
      for (int nd = 0; nd < D / 16; ++nd) {
        uint32_t kf[4];
        const int mrow = ks * 16 + (lane & 7) + ((lane >> 3) & 1) * 8, mcol = lane >> 4;
        ldsm_x4_t(kf, kaddr + tile_off<D>(mrow, nd * 2 + mcol));
        mma16816(dq[nd * 2], sa, kf[0], kf[1]);
        mma16816(dq[nd * 2 + 1], sa, kf[2], kf[3]);
      }
    }
    __syncthreads();
  }
  cp_async_wait<0>();
  __nv_bfloat16* dqb = p.dq + static_cast<long>(seq_start) * p.q_stride_t + head * D;
#pragma unroll
  for (int i = 0; i < D / 8; ++i) {
    const int col = i * 8 + t4 * 2;
    if (row_a < seq_len) *reinterpret_cast<uint32_t*>(dqb + static_cast<long>(row_a) * p.q_stride_t + col) = pack_bf16x2(dq[i][0], dq[i][1]);
    if (row_b < seq_len) *reinterpret_cast<uint32_t*>(dqb + static_cast<long>(row_b) * p.q_stride_t + col) = pack_bf16x2(dq[i][2], dq[i][3]);
  }
}

template <typename K>
static cudaError_t set_smem(K kern, int bytes) {
  return cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
}

}  // namespace nrl

using namespace nrl;

static AttnParams make_params(const void* q, const void* k, const void* v, long qs, long ks, long vs, long os,
                              const int* cu, int num_seqs, int total, int Hq, int Hkv, float scale) {
  AttnParams p{};
  p.q = static_cast<const __nv_bfloat16*>(q);
  p.k = static_cast<const __nv_bfloat16*>(k);
  p.v = static_cast<const __nv_bfloat16*>(v);
  p.q_stride_t = qs; p.k_stride_t = ks; p.v_stride_t = vs; p.o_stride_t = os;
  p.cu_seqlens = cu; p.num_seqs = num_seqs; p.total_tokens = total; p.Hq = Hq; p.Hkv = Hkv; p.G = Hq / Hkv;
  p.scale = scale; p.scale_log2 = scale * 1.4426950408889634f;
  return p;
}

// grid upper bound: every sequence contributes ceil(len/BM) blocks <= total/BM + num_seqs
extern "C" cudaError_t nrl_attn_varlen_fwd(const void* q, const void* k, const void* v, void* out, float* lse, long qs,
                                           long ks, long vs, long os, const int* cu, int num_seqs, int total, int Hq,
                                           int Hkv, int D, float scale, int causal, const void* rel_a, const void* rel_b,
                                           const short* lut, int lut_center, int NB, cudaStream_t s) {
  if (total == 0) return cudaSuccess;
  AttnParams p = make_params(q, k, v, qs, ks, vs, os, cu, num_seqs, total, Hq, Hkv, scale);
  p.out = static_cast<__nv_bfloat16*>(out);
  p.lse = lse;
  p.rel_a = static_cast<const __nv_bfloat16*>(rel_a);
  p.rel_b = static_cast<const __nv_bfloat16*>(rel_b);
  p.bucket_lut = lut; p.lut_center = lut_center; p.NB = NB;
  dim3 grid(total / 64 + num_seqs, Hq);
  cudaError_t e;
  if (rel_a != nullptr) {
    if (D != 64 || causal) return cudaErrorInvalidValue;
    constexpr int BN = 32;
    const int smem = 64 * 64 * 2 + 4 * BN * 64 * 2 + 2 * (64 + BN) * kRelStride * 2;
    auto kern = flash_fwd_kernel<64, false, true, BN>;
    if ((e = set_smem(kern, smem)) != cudaSuccess) return e;
    kern<<<grid, 128, smem, s>>>(p);
  } else if (D == 128 && causal) {
    constexpr int BN = 32;
    const int smem = 128 * 128 * 2 + 4 * BN * 128 * 2;
    auto kern = flash_fwd_causal_mt2_kernel<128, BN>;
    if ((e = set_smem(kern, smem)) != cudaSuccess) return e;
    dim3 grid2(total / 128 + num_seqs, Hq);
    kern<<<grid2, 128, smem, s>>>(p);
  } else if (D == 64 && causal) {
    const int smem = 64 * 64 * 2 + 4 * 64 * 64 * 2;
    auto kern = flash_fwd_kernel<64, true, false, 64>;
    if ((e = set_smem(kern, smem)) != cudaSuccess) return e;
    kern<<<grid, 128, smem, s>>>(p);
  } else if (D == 64 && !causal) {
    const int smem = 64 * 64 * 2 + 4 * 64 * 64 * 2;
    auto kern = flash_fwd_kernel<64, false, false, 64>;
    if ((e = set_smem(kern, smem)) != cudaSuccess) return e;
    kern<<<grid, 128, smem, s>>>(p);
  } else {
    return cudaErrorInvalidValue;
  }
  return cudaGetLastError();
}

extern "C" cudaError_t nrl_attn_varlen_bwd(const void* dout, const void* q, const void* k, const void* v, const void* o,
                                           const float* lse, float* delta, void* dq, void* dk, void* dv, long qs, long ks,
                                           long vs, long os, const int* cu, int num_seqs, int total, int Hq, int Hkv,
                                           int D, float scale, cudaStream_t s) {
  if (total == 0) return cudaSuccess;
  if (D != 128 && D != 64) return cudaErrorInvalidValue;
  AttnParams p = make_params(q, k, v, qs, ks, vs, os, cu, num_seqs, total, Hq, Hkv, scale);
  p.o = static_cast<const __nv_bfloat16*>(o);
  p.dout = static_cast<const __nv_bfloat16*>(dout);
  p.lse = const_cast<float*>(lse);
  p.delta = delta;
  p.dq = static_cast<__nv_bfloat16*>(dq);
  p.dk = static_cast<__nv_bfloat16*>(dk);
  p.dv = static_cast<__nv_bfloat16*>(dv);
  const long pairs = static_cast<long>(total) * Hq;
  attn_bwd_delta_kernel<<<static_cast<int>((pairs + 7) / 8), 256, 0, s>>>(p, D);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return e;
  dim3 grid_kv(total / 64 + num_seqs, Hkv), grid_q(total / 64 + num_seqs, Hq);
  if (D == 128) {
    const int smem_kv = 2 * 64 * 128 * 2 + 4 * 32 * 128 * 2 + 4 * 32 * 4;
    const int smem_q = 2 * 64 * 128 * 2 + 4 * 32 * 128 * 2;
    if ((e = set_smem(flash_bwd_dkdv_kernel<128>, smem_kv)) != cudaSuccess) return e;
    if ((e = set_smem(flash_bwd_dq_kernel<128>, smem_q)) != cudaSuccess) return e;
    flash_bwd_dkdv_kernel<128><<<grid_kv, 128, smem_kv, s>>>(p);
    flash_bwd_dq_kernel<128><<<grid_q, 128, smem_q, s>>>(p);
  } else {
    const int smem_kv = 2 * 64 * 64 * 2 + 4 * 32 * 64 * 2 + 4 * 32 * 4;
    const int smem_q = 2 * 64 * 64 * 2 + 4 * 32 * 64 * 2;
    if ((e = set_smem(flash_bwd_dkdv_kernel<64>, smem_kv)) != cudaSuccess) return e;
    if ((e = set_smem(flash_bwd_dq_kernel<64>, smem_q)) != cudaSuccess) return e;
    flash_bwd_dkdv_kernel<64><<<grid_kv, 128, smem_kv, s>>>(p);
    flash_bwd_dq_kernel<64><<<grid_q, 128, smem_q, s>>>(p);
  }
  return cudaGetLastError();
}

// TMA-fed DeBERTa attention: maps = {Q [T,H*64] box 64x64 sw128, K box 32 rows, V box 32 rows, relA [H*T,NB] box 64 x 112
// (no swizzle), relB box 32 x 112}
extern "C" cudaError_t nrl_deberta_attn_fwd(const CUtensorMap* maps, void* out, float* lse, long os, const int* cu, int num_seqs,
                                            int total, int Hq, float scale, const short* lut, int lut_center, int NB, int bn,
                                            cudaStream_t s) {
  using namespace nrl;
  if (total == 0) return cudaSuccess;
  AttnParams p = make_params(nullptr, nullptr, nullptr, 0, 0, 0, os, cu, num_seqs, total, Hq, Hq, scale);
  p.out = static_cast<__nv_bfloat16*>(out);
  p.lse = lse;
  p.bucket_lut = lut; p.lut_center = lut_center; p.NB = NB;
  cudaError_t e;
  dim3 grid(total / 64 + num_seqs, Hq);
  if (bn == 64) {
    constexpr int smem = 64 * 64 * 2 + 4 * 64 * 64 * 2 + 2 * (64 + 64) * 136 * 2 + 64;
    if ((e = set_smem(deberta_attn_fwd_tma_kernel<64>, smem)) != cudaSuccess) return e;
    deberta_attn_fwd_tma_kernel<64><<<grid, 128, smem, s>>>(maps[0], maps[1], maps[2], maps[3], maps[4], p);
  } else {
    constexpr int smem = 64 * 64 * 2 + 4 * 32 * 64 * 2 + 2 * (64 + 32) * 112 * 2 + 64;
    if ((e = set_smem(deberta_attn_fwd_tma_kernel<32>, smem)) != cudaSuccess) return e;
    deberta_attn_fwd_tma_kernel<32><<<grid, 128, smem, s>>>(maps[0], maps[1], maps[2], maps[3], maps[4], p);
  }
  return cudaGetLastError();
}
