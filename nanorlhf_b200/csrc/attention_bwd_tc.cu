// tcgen05 / TMEM / TMA flash-attention BACKWARD (causal, packed varlen, GQA, head_dim 128) for sm_100a.
//
// Three kernels, no atomics (same decomposition as the mma.sync version in attention_varlen.cu, which stays the oracle):
//   delta   : delta[h, t] = sum_d dO * O                                   (one warp per (token, head))
//   dK / dV : one CTA per (128-key block, kv head).  K_j, V_j stay in smem; the CTA streams (q head of the group,
//             64-row query block) pairs:   S^T = K_j Q_i^T,  dP^T = V_j dO_i^T          (tcgen05, TMEM lanes = keys)
//             P^T = exp2(S^T*c - lse_i),  dS^T = P^T o (dP^T - delta_i) * scale         (softmax warps, one key row each)
//             dV_j += P^T dO_i,  dK_j += dS^T Q_i                                         (tcgen05, accumulators in TMEM)
//   dQ      : one CTA per (128-row query block, q head).  Q_i, dO_i stay in smem; the CTA streams 64-key blocks:
//             S = Q_i K_j^T,  dP = dO_i V_j^T  ->  dS = P o (dP - delta) * scale  ->  dQ_i += dS K_j
// Shared structure: warp 0 = TMA producer (3-stage ring), warp 1 = single-thread MMA issuer, warps 2..5 and 6..9 = two
// elementwise warpgroups that take alternate blocks (each owns one S/dP TMEM buffer and one P/dS smem buffer), so
// the exp2/convert work of block n overlaps the MMAs of block n+1.  The operands that must be read "transposed"
// (dO_i, Q_i for dV/dK; K_j for dQ) are consumed in place as MN-major UMMA operands - nothing is transposed in smem.
// Replaces the flash-attn-2 backward the reference reaches through HF (attn_implementation="flash_attention_2",
// /root/reference/GRPO/grpo.py:219; training step /root/reference/GRPO/grpo_trainer.py:652).
#include "common.cuh"
#include "kernels.h"

namespace nrl {

constexpr int kBwdThreads = 320;
constexpr int kP128 = 128 * 128;     // [128 rows][64 x bf16] swizzled panel (16 KB)
constexpr int kT128 = 2 * kP128;     // 128 x 128 bf16 tile (32 KB)
constexpr int kP64 = 64 * 128;       // [64 rows][64 x bf16] panel (8 KB)
constexpr int kT64 = 2 * kP64;       // 64 x 128 bf16 tile (16 KB)
constexpr int kPS = 128 * 128;       // [128 rows][64 x bf16] P / dS tile (16 KB)
constexpr int kRing = 3;

struct AttnBwdParams {
  const int* cu_seqlens;
  const float* lse;           // [Hq, T]
  const float* delta;         // [Hq, T]
  __nv_bfloat16 *dq, *dk, *dv;
  long dq_stride_t, dkv_stride_t;
  int num_seqs, total_tokens, Hq, G;
  float scale, scale_log2;
};

// (sequence, block) lookup for packed sequences; result broadcast through `info` (dynamic smem, 4 ints)
template <int BM>
NRL_DEVICE bool bwd_locate(const int* cu, int num_seqs, int blk, int* info, int& m_blk, int& seq_start, int& seq_len) {
  if (threadIdx.x == 0) {
    int acc = 0, found = 0;
    for (int s = 0; s < num_seqs; ++s) {
      const int a = cu[s], b = cu[s + 1];
      const int nb = (b - a + BM - 1) / BM;
      if (blk < acc + nb) {
        info[0] = 1; info[1] = blk - acc; info[2] = a; info[3] = b - a;
        found = 1;
        break;
      }
      acc += nb;
    }
    if (!found) info[0] = 0;
  }
  __syncthreads();
  if (!info[0]) return false;
  m_blk = info[1]; seq_start = info[2]; seq_len = info[3];
  return true;
}

// 32 fp32 values (as bit patterns) -> 16 packed bf16x2, written in place into the low half
NRL_DEVICE void store_row_chunk(uint8_t* row_base, int r, int c32, const uint32_t (&pk)[16]) {
  // columns c32*32 .. +31 of a 64-column (128-byte) row: 16-byte chunks c32*4 .. +3, 128B-swizzled by row
#pragma unroll
  for (int q4 = 0; q4 < 4; ++q4) {
    const int ch = c32 * 4 + q4;
    *reinterpret_cast<uint4*>(row_base + ((ch ^ (r & 7)) << 4)) = make_uint4(pk[q4 * 4], pk[q4 * 4 + 1], pk[q4 * 4 + 2], pk[q4 * 4 + 3]);
  }
}

__global__ void attn_bwd_delta_tc_kernel(const __nv_bfloat16* __restrict__ o, const __nv_bfloat16* __restrict__ dout,
                                         float* __restrict__ delta, long o_stride_t, int total, int Hq) {
  const long idx = blockIdx.x * static_cast<long>(blockDim.x >> 5) + (threadIdx.x >> 5);
  if (idx >= static_cast<long>(total) * Hq) return;
  const int lane = threadIdx.x & 31;
  const long t = idx / Hq;
  const int h = idx % Hq;
  const uint2 a = *reinterpret_cast<const uint2*>(o + t * o_stride_t + h * 128 + lane * 4);
  const uint2 b = *reinterpret_cast<const uint2*>(dout + t * o_stride_t + h * 128 + lane * 4);
  const float2 a0 = unpack_bf16x2(a.x), a1 = unpack_bf16x2(a.y), b0 = unpack_bf16x2(b.x), b1 = unpack_bf16x2(b.y);
  float acc = a0.x * b0.x + a0.y * b0.y + a1.x * b1.x + a1.y * b1.y;
  acc = warp_sum(acc);
  if (lane == 0) delta[static_cast<long>(h) * total + t] = acc;
}

// =====================================================================================================================
// dK / dV
// =====================================================================================================================
struct DkdvSmem {
  static constexpr int kK = 0;
  static constexpr int kV = kT128;
  static constexpr int kQ = 2 * kT128;                     // ring
  static constexpr int kDO = kQ + kRing * kT64;            // ring
  static constexpr int kP = kDO + kRing * kT64;            // [wg]
  static constexpr int kDS = kP + 2 * kPS;                 // [wg]
  static constexpr int kLD = kDS + 2 * kPS;                // [wg][parity][lse2 x 64 | delta x 64] floats
  static constexpr int kBar = kLD + 2 * 2 * 128 * 4;
  static constexpr int kTotal = kBar + 256;
};

__global__ void __launch_bounds__(kBwdThreads, 1)
attn_bwd_dkdv_tc_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmDO,
                        const __grid_constant__ CUtensorMap tmK, const __grid_constant__ CUtensorMap tmV, AttnBwdParams p) {
  using L = DkdvSmem;
  extern __shared__ __align__(1024) uint8_t smem[];
  uint64_t* kv_full = reinterpret_cast<uint64_t*>(smem + L::kBar);
  uint64_t* qdo_full = kv_full + 1;            // [kRing]
  uint64_t* qdo_empty = qdo_full + kRing;      // [kRing]
  uint64_t* s_full = qdo_empty + kRing;        // [wg]  S^T and dP^T of the block are in TMEM
  uint64_t* pds_ready = s_full + 2;            // [wg]  P^T / dS^T of the block are in smem
  uint64_t* pds_free = pds_ready + 2;          // [wg]  the dV / dK MMAs that read them retired
  uint64_t* done_bar = pds_free + 2;
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(done_bar + 1);
  int* info = reinterpret_cast<int*>(tmem_ptr + 4);      // its own 16-byte slot: tcgen05.alloc writes next to it

  int kb, seq_start, seq_len;
  if (!bwd_locate<128>(p.cu_seqlens, p.num_seqs, blockIdx.x, info, kb, seq_start, seq_len)) return;
  const int kvh = blockIdx.y;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int key0 = kb * 128;
  const int i_begin = key0 / 64;                                   // first 64-row query block that sees these keys
  const int n_qb = (seq_len + 63) / 64 - i_begin;                  // >= 1
  const int n_iter = p.G * n_qb;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmQ); tma_prefetch_desc(&tmDO); tma_prefetch_desc(&tmK); tma_prefetch_desc(&tmV);
  }
  if (warp == 1 && lane == 0) {
    mbar_init(kv_full, 1);
    for (int i = 0; i < kRing; ++i) { mbar_init(&qdo_full[i], 1); mbar_init(&qdo_empty[i], 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(&s_full[i], 1); mbar_init(&pds_ready[i], 4); mbar_init(&pds_free[i], 1); }
    mbar_init(done_bar, 1);
    fence_mbar_init();
  }
  if (warp == 1) { __syncwarp(); tmem_alloc(tmem_ptr, 512); }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;
  // TMEM columns: S^T[wg] 0/64, dP^T[wg] 128/192, dV 256..383, dK 384..511

  if (warp == 0) {
    // ============================== TMA producer ==============================
    if (elect_one()) {
      mbar_arrive_expect_tx(kv_full, 2 * kT128);
      tma_load_2d(smem + L::kK, &tmK, kv_full, kvh * 128, seq_start + key0);
      tma_load_2d(smem + L::kK + kP128, &tmK, kv_full, kvh * 128 + 64, seq_start + key0);
      tma_load_2d(smem + L::kV, &tmV, kv_full, kvh * 128, seq_start + key0);
      tma_load_2d(smem + L::kV + kP128, &tmV, kv_full, kvh * 128 + 64, seq_start + key0);
      int st = 0;
      uint32_t phase = 0;
      for (int it = 0; it < n_iter; ++it) {
        const int head = kvh * p.G + it / n_qb;
        const int row = seq_start + (i_begin + it % n_qb) * 64;
        mbar_wait(&qdo_empty[st], phase ^ 1);
        mbar_arrive_expect_tx(&qdo_full[st], 2 * kT64);
        uint8_t* sq = smem + L::kQ + st * kT64;
        uint8_t* sd = smem + L::kDO + st * kT64;
        tma_load_2d(sq, &tmQ, &qdo_full[st], head * 128, row);
        tma_load_2d(sq + kP64, &tmQ, &qdo_full[st], head * 128 + 64, row);
        tma_load_2d(sd, &tmDO, &qdo_full[st], head * 128, row);
        tma_load_2d(sd + kP64, &tmDO, &qdo_full[st], head * 128 + 64, row);
        if (++st == kRing) { st = 0; phase ^= 1; }
      }
    }
  } else if (warp == 1) {
    // ============================== MMA issuer ==============================
    if (elect_one()) {
      constexpr uint32_t idesc_s = make_idesc(128, 64, 1, 1);          // [128 keys] x [64 q], both K-major (over d)
      constexpr uint32_t idesc_g = make_idesc_bmn(128, 128);          // [128 keys] x [128 d], B MN-major (over q)
      const uint32_t k_addr = smem_u32(smem + L::kK), v_addr = smem_u32(smem + L::kV);
      mbar_wait(kv_full, 0);
      tc_fence_after();
      int st_s = 0;
      uint32_t ph_s = 0;
      auto issue_sdp = [&](int it) {
        const int b = it & 1;
        mbar_wait(&qdo_full[st_s], ph_s);
        tc_fence_after();
        const uint32_t q_addr = smem_u32(smem + L::kQ + st_s * kT64);
        const uint32_t do_addr = smem_u32(smem + L::kDO + st_s * kT64);
#pragma unroll
        for (int k = 0; k < 8; ++k)
          umma_f16(tmem_base + b * 64, make_smem_desc_sw128(k_addr + (k >> 2) * kP128 + (k & 3) * 32),
                   make_smem_desc_sw128(q_addr + (k >> 2) * kP64 + (k & 3) * 32), idesc_s, k != 0 ? 1u : 0u);
#pragma unroll
        for (int k = 0; k < 8; ++k)
          umma_f16(tmem_base + 128 + b * 64, make_smem_desc_sw128(v_addr + (k >> 2) * kP128 + (k & 3) * 32),
                   make_smem_desc_sw128(do_addr + (k >> 2) * kP64 + (k & 3) * 32), idesc_s, k != 0 ? 1u : 0u);
        umma_commit(&s_full[b]);
        if (++st_s == kRing) { st_s = 0; ph_s ^= 1; }
      };
      issue_sdp(0);
      int st = 0;
      for (int it = 0; it < n_iter; ++it) {
        const int b = it & 1;
        if (it + 1 < n_iter) issue_sdp(it + 1);             // overlaps the elementwise work of block it
        mbar_wait(&pds_ready[b], (it >> 1) & 1);
        tc_fence_after();
        const uint32_t p_addr = smem_u32(smem + L::kP + b * kPS);
        const uint32_t ds_addr = smem_u32(smem + L::kDS + b * kPS);
        const uint32_t q_addr = smem_u32(smem + L::kQ + st * kT64);
        const uint32_t do_addr = smem_u32(smem + L::kDO + st * kT64);
#pragma unroll
        for (int k = 0; k < 4; ++k)            // 4 x 16 query rows
          umma_f16(tmem_base + 256, make_smem_desc_sw128(p_addr + k * 32),
                   make_smem_desc_sw128_mn(do_addr + k * 16 * 128, kP64), idesc_g, (it | k) != 0 ? 1u : 0u);
#pragma unroll
        for (int k = 0; k < 4; ++k)
          umma_f16(tmem_base + 384, make_smem_desc_sw128(ds_addr + k * 32),
                   make_smem_desc_sw128_mn(q_addr + k * 16 * 128, kP64), idesc_g, (it | k) != 0 ? 1u : 0u);
        umma_commit(&qdo_empty[st]);
        umma_commit(&pds_free[b]);
        if (++st == kRing) st = 0;
      }
      umma_commit(done_bar);
    }
    __syncwarp();
  } else {
    // ============================== elementwise warpgroups ==============================
    const int wg = (warp - 2) >> 2;
    const int quad = warp & 3;
    const int r = quad * 32 + lane;                          // key row in tile == TMEM lane
    const int t128 = (warp - 2 - wg * 4) * 32 + lane;        // 0..127 within the warpgroup
    const uint32_t lane_off = static_cast<uint32_t>(quad * 32) << 16;
    const int kk = key0 + r;
    float* ld_base = reinterpret_cast<float*>(smem + L::kLD) + wg * 256;
    auto load_ld = [&](int it) -> float {
      // threads 0..63 fetch lse (in log2 units), 64..127 fetch delta, of query row (t128 & 63) of block `it`
      if (it >= n_iter) return 0.f;
      const int head = kvh * p.G + it / n_qb;
      const int qq = (i_begin + it % n_qb) * 64 + (t128 & 63);
      if (qq >= seq_len) return t128 < 64 ? INFINITY : 0.f;          // lse = +inf  =>  P = 0 for rows past the end
      const long off = static_cast<long>(head) * p.total_tokens + seq_start + qq;
      return t128 < 64 ? p.lse[off] * 1.4426950408889634f : p.delta[off];
    };
    float ld_next = load_ld(wg);
    for (int it = wg; it < n_iter; it += 2) {
      const int n = it >> 1;
      float* ldbuf = ld_base + (n & 1) * 128;
      ldbuf[t128] = ld_next;
      named_barrier_sync(1 + wg, 128);
      ld_next = load_ld(it + 2);
      const int q0 = (i_begin + it % n_qb) * 64;
      const bool diag = q0 < key0 + 127;                     // some (key, query) pairs of the block are masked
      mbar_wait(&s_full[wg], n & 1);
      tc_fence_after();
      uint32_t pk[2][16], dk_[2][16];
#pragma unroll
      for (int c32 = 0; c32 < 2; ++c32) {
        uint32_t sv[32], dv_[32];
        tmem_ld_32x32b_x32(tmem_base + wg * 64 + c32 * 32 + lane_off, sv);
        tmem_ld_32x32b_x32(tmem_base + 128 + wg * 64 + c32 * 32 + lane_off, dv_);
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 32; i += 2) {
          const float2 l2 = *reinterpret_cast<const float2*>(ldbuf + c32 * 32 + i);
          const float2 dl = *reinterpret_cast<const float2*>(ldbuf + 64 + c32 * 32 + i);
          float p0 = exp2f(fmaf(__uint_as_float(sv[i]), p.scale_log2, -l2.x));
          float p1 = exp2f(fmaf(__uint_as_float(sv[i + 1]), p.scale_log2, -l2.y));
          if (diag) {
            if (q0 + c32 * 32 + i < kk) p0 = 0.f;
            if (q0 + c32 * 32 + i + 1 < kk) p1 = 0.f;
          }
          const float d0 = p0 * (__uint_as_float(dv_[i]) - dl.x) * p.scale;
          const float d1 = p1 * (__uint_as_float(dv_[i + 1]) - dl.y) * p.scale;
          pk[c32][i / 2] = pack_bf16x2(p0, p1);
          dk_[c32][i / 2] = pack_bf16x2(d0, d1);
        }
      }
      // the dV / dK MMAs of this warpgroup's previous block (it-2) must be done with P^T / dS^T
      if (it >= 2) {
        mbar_wait(&pds_free[wg], (n - 1) & 1);
        tc_fence_after();
      }
      uint8_t* prow = smem + L::kP + wg * kPS + r * 128;
      uint8_t* drow = smem + L::kDS + wg * kPS + r * 128;
      store_row_chunk(prow, r, 0, pk[0]);
      store_row_chunk(prow, r, 1, pk[1]);
      store_row_chunk(drow, r, 0, dk_[0]);
      store_row_chunk(drow, r, 1, dk_[1]);
      fence_proxy_async_smem();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&pds_ready[wg]);
    }
    // ---- epilogue: warpgroup 0 writes dV, warpgroup 1 writes dK ----
    mbar_wait(done_bar, 0);
    tc_fence_after();
    const bool row_ok = kk < seq_len;
    __nv_bfloat16* dst = (wg == 0 ? p.dv : p.dk) + static_cast<long>(seq_start + kk) * p.dkv_stride_t + kvh * 128;
#pragma unroll 1
    for (int c = 0; c < 4; ++c) {
      uint32_t ov[32];
      tmem_ld_32x32b_x32(tmem_base + 256 + wg * 128 + c * 32 + lane_off, ov);     // warp-converged (.sync.aligned)
      tmem_ld_wait();
      if (row_ok) {
#pragma unroll
        for (int q4 = 0; q4 < 4; ++q4) {
          uint4 o4;
          o4.x = pack_bf16x2(__uint_as_float(ov[q4 * 8]), __uint_as_float(ov[q4 * 8 + 1]));
          o4.y = pack_bf16x2(__uint_as_float(ov[q4 * 8 + 2]), __uint_as_float(ov[q4 * 8 + 3]));
          o4.z = pack_bf16x2(__uint_as_float(ov[q4 * 8 + 4]), __uint_as_float(ov[q4 * 8 + 5]));
          o4.w = pack_bf16x2(__uint_as_float(ov[q4 * 8 + 6]), __uint_as_float(ov[q4 * 8 + 7]));
          *reinterpret_cast<uint4*>(dst + c * 32 + q4 * 8) = o4;
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) { tc_fence_after(); tmem_dealloc(tmem_base, 512); }
}

// =====================================================================================================================
// dQ
// =====================================================================================================================
struct DqSmem {
  static constexpr int kQ = 0;
  static constexpr int kDO = kT128;
  static constexpr int kK = 2 * kT128;                     // ring
  static constexpr int kV = kK + kRing * kT64;             // ring
  static constexpr int kDS = kV + kRing * kT64;            // [wg]
  static constexpr int kBar = kDS + 2 * kPS;
  static constexpr int kTotal = kBar + 256;
};

__global__ void __launch_bounds__(kBwdThreads, 1)
attn_bwd_dq_tc_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmDO,
                      const __grid_constant__ CUtensorMap tmK, const __grid_constant__ CUtensorMap tmV, AttnBwdParams p) {
  using L = DqSmem;
  extern __shared__ __align__(1024) uint8_t smem[];
  uint64_t* q_full = reinterpret_cast<uint64_t*>(smem + L::kBar);
  uint64_t* kv_full = q_full + 1;              // [kRing]
  uint64_t* kv_empty = kv_full + kRing;        // [kRing]
  uint64_t* s_full = kv_empty + kRing;         // [wg]
  uint64_t* ds_ready = s_full + 2;             // [wg]
  uint64_t* ds_free = ds_ready + 2;            // [wg]
  uint64_t* done_bar = ds_free + 2;
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(done_bar + 1);
  int* info = reinterpret_cast<int*>(tmem_ptr + 4);      // its own 16-byte slot: tcgen05.alloc writes next to it

  int mb, seq_start, seq_len;
  // latest (longest) query blocks first
  if (!bwd_locate<128>(p.cu_seqlens, p.num_seqs, gridDim.x - 1 - blockIdx.x, info, mb, seq_start, seq_len)) return;
  const int head = blockIdx.y, kvh = head / p.G;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int q0 = mb * 128;
  const int n_kb = min((seq_len + 63) / 64, (q0 + 128 + 63) / 64);

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmQ); tma_prefetch_desc(&tmDO); tma_prefetch_desc(&tmK); tma_prefetch_desc(&tmV);
  }
  if (warp == 1 && lane == 0) {
    mbar_init(q_full, 1);
    for (int i = 0; i < kRing; ++i) { mbar_init(&kv_full[i], 1); mbar_init(&kv_empty[i], 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(&s_full[i], 1); mbar_init(&ds_ready[i], 4); mbar_init(&ds_free[i], 1); }
    mbar_init(done_bar, 1);
    fence_mbar_init();
  }
  if (warp == 1) { __syncwarp(); tmem_alloc(tmem_ptr, 512); }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;
  // TMEM columns: S[wg] 0/64, dP[wg] 128/192, dQ 256..383

  if (warp == 0) {
    if (elect_one()) {
      mbar_arrive_expect_tx(q_full, 2 * kT128);
      tma_load_2d(smem + L::kQ, &tmQ, q_full, head * 128, seq_start + q0);
      tma_load_2d(smem + L::kQ + kP128, &tmQ, q_full, head * 128 + 64, seq_start + q0);
      tma_load_2d(smem + L::kDO, &tmDO, q_full, head * 128, seq_start + q0);
      tma_load_2d(smem + L::kDO + kP128, &tmDO, q_full, head * 128 + 64, seq_start + q0);
      int st = 0;
      uint32_t phase = 0;
      for (int j = 0; j < n_kb; ++j) {
        mbar_wait(&kv_empty[st], phase ^ 1);
        mbar_arrive_expect_tx(&kv_full[st], 2 * kT64);
        uint8_t* sk = smem + L::kK + st * kT64;
        uint8_t* sv = smem + L::kV + st * kT64;
        const int row = seq_start + j * 64;
        tma_load_2d(sk, &tmK, &kv_full[st], kvh * 128, row);
        tma_load_2d(sk + kP64, &tmK, &kv_full[st], kvh * 128 + 64, row);
        tma_load_2d(sv, &tmV, &kv_full[st], kvh * 128, row);
        tma_load_2d(sv + kP64, &tmV, &kv_full[st], kvh * 128 + 64, row);
        if (++st == kRing) { st = 0; phase ^= 1; }
      }
    }
  } else if (warp == 1) {
    if (elect_one()) {
      constexpr uint32_t idesc_s = make_idesc(128, 64, 1, 1);          // [128 q] x [64 keys], both K-major (over d)
      constexpr uint32_t idesc_g = make_idesc_bmn(128, 128);          // [128 q] x [128 d], B = K_j MN-major (over keys)
      const uint32_t q_addr = smem_u32(smem + L::kQ), do_addr = smem_u32(smem + L::kDO);
      mbar_wait(q_full, 0);
      tc_fence_after();
      int st_s = 0;
      uint32_t ph_s = 0;
      auto issue_sdp = [&](int j) {
        const int b = j & 1;
        mbar_wait(&kv_full[st_s], ph_s);
        tc_fence_after();
        const uint32_t k_addr = smem_u32(smem + L::kK + st_s * kT64);
        const uint32_t v_addr = smem_u32(smem + L::kV + st_s * kT64);
#pragma unroll
        for (int k = 0; k < 8; ++k)
          umma_f16(tmem_base + b * 64, make_smem_desc_sw128(q_addr + (k >> 2) * kP128 + (k & 3) * 32),
                   make_smem_desc_sw128(k_addr + (k >> 2) * kP64 + (k & 3) * 32), idesc_s, k != 0 ? 1u : 0u);
#pragma unroll
        for (int k = 0; k < 8; ++k)
          umma_f16(tmem_base + 128 + b * 64, make_smem_desc_sw128(do_addr + (k >> 2) * kP128 + (k & 3) * 32),
                   make_smem_desc_sw128(v_addr + (k >> 2) * kP64 + (k & 3) * 32), idesc_s, k != 0 ? 1u : 0u);
        umma_commit(&s_full[b]);
        if (++st_s == kRing) { st_s = 0; ph_s ^= 1; }
      };
      issue_sdp(0);
      int st = 0;
      for (int j = 0; j < n_kb; ++j) {
        const int b = j & 1;
        if (j + 1 < n_kb) issue_sdp(j + 1);
        mbar_wait(&ds_ready[b], (j >> 1) & 1);
        tc_fence_after();
        const uint32_t ds_addr = smem_u32(smem + L::kDS + b * kPS);
        const uint32_t k_addr = smem_u32(smem + L::kK + st * kT64);
#pragma unroll
        for (int k = 0; k < 4; ++k)            // 4 x 16 keys
          umma_f16(tmem_base + 256, make_smem_desc_sw128(ds_addr + k * 32),
                   make_smem_desc_sw128_mn(k_addr + k * 16 * 128, kP64), idesc_g, (j | k) != 0 ? 1u : 0u);
        umma_commit(&kv_empty[st]);
        umma_commit(&ds_free[b]);
        if (++st == kRing) st = 0;
      }
      umma_commit(done_bar);
    }
    __syncwarp();
  } else {
    const int wg = (warp - 2) >> 2;
    const int quad = warp & 3;
    const int r = quad * 32 + lane;                          // query row in tile == TMEM lane
    const uint32_t lane_off = static_cast<uint32_t>(quad * 32) << 16;
    const int row = q0 + r;
    const bool row_ok = row < seq_len;
    const long loff = static_cast<long>(head) * p.total_tokens + seq_start + row;
    const float lse2 = row_ok ? p.lse[loff] * 1.4426950408889634f : INFINITY;     // +inf => P = 0
    const float dl = row_ok ? p.delta[loff] : 0.f;
    for (int j = wg; j < n_kb; j += 2) {
      const int n = j >> 1;
      const int key0 = j * 64;
      const bool edge = (key0 + 63 > q0) || (key0 + 64 > seq_len);
      mbar_wait(&s_full[wg], n & 1);
      tc_fence_after();
      uint32_t dk_[2][16];
#pragma unroll
      for (int c32 = 0; c32 < 2; ++c32) {
        uint32_t sv[32], dv_[32];
        tmem_ld_32x32b_x32(tmem_base + wg * 64 + c32 * 32 + lane_off, sv);
        tmem_ld_32x32b_x32(tmem_base + 128 + wg * 64 + c32 * 32 + lane_off, dv_);
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 32; i += 2) {
          float p0 = exp2f(fmaf(__uint_as_float(sv[i]), p.scale_log2, -lse2));
          float p1 = exp2f(fmaf(__uint_as_float(sv[i + 1]), p.scale_log2, -lse2));
          if (edge) {
            const int key = key0 + c32 * 32 + i;
            if (key > row || key >= seq_len) p0 = 0.f;
            if (key + 1 > row || key + 1 >= seq_len) p1 = 0.f;
          }
          const float d0 = p0 * (__uint_as_float(dv_[i]) - dl) * p.scale;
          const float d1 = p1 * (__uint_as_float(dv_[i + 1]) - dl) * p.scale;
          dk_[c32][i / 2] = pack_bf16x2(d0, d1);
        }
      }
      if (j >= 2) {
        mbar_wait(&ds_free[wg], (n - 1) & 1);
        tc_fence_after();
      }
      uint8_t* drow = smem + L::kDS + wg * kPS + r * 128;
      store_row_chunk(drow, r, 0, dk_[0]);
      store_row_chunk(drow, r, 1, dk_[1]);
      fence_proxy_async_smem();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&ds_ready[wg]);
    }
    // ---- epilogue: each warpgroup writes 64 of the 128 dQ columns ----
    mbar_wait(done_bar, 0);
    tc_fence_after();
    __nv_bfloat16* dst = p.dq + static_cast<long>(seq_start + row) * p.dq_stride_t + head * 128 + wg * 64;
#pragma unroll 1
    for (int c = 0; c < 2; ++c) {
      uint32_t ov[32];
      tmem_ld_32x32b_x32(tmem_base + 256 + wg * 64 + c * 32 + lane_off, ov);
      tmem_ld_wait();
      if (row_ok) {
#pragma unroll
        for (int q4 = 0; q4 < 4; ++q4) {
          uint4 o4;
          o4.x = pack_bf16x2(__uint_as_float(ov[q4 * 8]), __uint_as_float(ov[q4 * 8 + 1]));
          o4.y = pack_bf16x2(__uint_as_float(ov[q4 * 8 + 2]), __uint_as_float(ov[q4 * 8 + 3]));
          o4.z = pack_bf16x2(__uint_as_float(ov[q4 * 8 + 4]), __uint_as_float(ov[q4 * 8 + 5]));
          o4.w = pack_bf16x2(__uint_as_float(ov[q4 * 8 + 6]), __uint_as_float(ov[q4 * 8 + 7]));
          *reinterpret_cast<uint4*>(dst + c * 32 + q4 * 8) = o4;
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) { tc_fence_after(); tmem_dealloc(tmem_base, 512); }
}

}  // namespace nrl

// maps: [0] Q box 64 rows, [1] dO box 64 rows, [2] K box 128 rows, [3] V box 128 rows  (dK/dV kernel)
//       [4] Q box 128 rows, [5] dO box 128 rows, [6] K box 64 rows, [7] V box 64 rows  (dQ kernel)
extern "C" cudaError_t nrl_attn_bwd_tc(const CUtensorMap* maps, const void* o, const void* dout, const float* lse,
                                       float* delta, void* dq, void* dk, void* dv, long o_stride_t, long dq_stride_t,
                                       long dkv_stride_t, const int* cu, int num_seqs, int total, int Hq, int Hkv,
                                       float scale, cudaStream_t s) {
  using namespace nrl;
  if (total == 0) return cudaSuccess;
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(attn_bwd_dkdv_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, DkdvSmem::kTotal);
    if (e != cudaSuccess) return e;
    e = cudaFuncSetAttribute(attn_bwd_dq_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, DqSmem::kTotal);
    if (e != cudaSuccess) return e;
    configured = true;
  }
  const long pairs = static_cast<long>(total) * Hq;
  attn_bwd_delta_tc_kernel<<<static_cast<int>((pairs + 7) / 8), 256, 0, s>>>(
      static_cast<const __nv_bfloat16*>(o), static_cast<const __nv_bfloat16*>(dout), delta, o_stride_t, total, Hq);
  AttnBwdParams p;
  p.cu_seqlens = cu; p.lse = lse; p.delta = delta;
  p.dq = static_cast<__nv_bfloat16*>(dq); p.dk = static_cast<__nv_bfloat16*>(dk); p.dv = static_cast<__nv_bfloat16*>(dv);
  p.dq_stride_t = dq_stride_t; p.dkv_stride_t = dkv_stride_t;
  p.num_seqs = num_seqs; p.total_tokens = total; p.Hq = Hq; p.G = Hq / Hkv;
  p.scale = scale; p.scale_log2 = scale * 1.4426950408889634f;
  dim3 grid_kv(total / 128 + num_seqs, Hkv), grid_q(total / 128 + num_seqs, Hq);
  attn_bwd_dkdv_tc_kernel<<<grid_kv, kBwdThreads, DkdvSmem::kTotal, s>>>(maps[0], maps[1], maps[2], maps[3], p);
  attn_bwd_dq_tc_kernel<<<grid_q, kBwdThreads, DqSmem::kTotal, s>>>(maps[4], maps[5], maps[6], maps[7], p);
  return cudaGetLastError();
}
