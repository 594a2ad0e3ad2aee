// tcgen05 / TMEM / TMA flash-attention FORWARD (causal, packed varlen, GQA, head_dim 128) for sm_100a.
//
// One CTA = 256 query rows of one head as TWO 128-row tiles (A, B) that share one K/V ring, so every K/V byte
// fetched from L2 feeds two tiles and the two softmax warpgroups hide each other's latencies (2 warps per SMSP).
// KV is streamed in blocks of 64 keys.
//   warp 0       TMA producer : Q tiles once, then K_j / V_j (3-stage ring), 128B-swizzled panels of 64 dims
//   warp 1       MMA issuer   : S_X,j = Q_X K_j^T  (8 x tcgen05.mma 128x64x16, both operands K-major) -> TMEM S_X[j%2]
//                               O_X  += P_X,j V_j  (4 x tcgen05.mma 128x128x16, A = P in smem K-major, B = V MN-major)
//   warps 2..5   softmax of tile A, warps 6..9 softmax of tile B: thread i owns query row i (TMEM lane i):
//                               tcgen05.ld S -> mask -> online max with lazy (2^8) rescale -> exp2 -> bf16 P into
//                               swizzled smem; rescales O in TMEM when needed; final O / l and LSE go straight to HBM
// TMEM (512 cols): S_A[2] 0..127, S_B[2] 128..255, O_A 256..383, O_B 384..511.
// SMEM: Q 64 KB + 3 x (K 16 KB + V 16 KB) + P 2 tiles x 2 buffers x 16 KB = 224 KB.
// The softmax never waits for the tensor pipe on the common path: P is double buffered (pv_done[X][b] guards reuse) and O is
// only rescaled - which needs PV of the previous block retired - when a row maximum grew by more than 2^8.
// Backward: attention_bwd_tc.cu.  Oracle: the mma.sync kernels in attention_varlen.cu.  Together they replace flash-attn-2
// (attn_implementation="flash_attention_2", /root/reference/GRPO/grpo.py:219) on the training, log-prob and prefill paths.
#include <cstdlib>

#include "common.cuh"
#include "kernels.h"

namespace nrl {

constexpr int kTcD = 128;                 // head dim
constexpr int kTcBM = 256;                // query rows per CTA (two tiles of 128)
constexpr int kTcBN = 64;                 // keys per block
constexpr int kQPanel = 128 * 128;        // [128 rows][64 x bf16] swizzled panel (16 KB)
constexpr int kQTile = 2 * kQPanel;       // one 128 x 128 Q tile (32 KB)
constexpr int kKVPanel = 64 * 128;        // [64 keys][64 x bf16] (8 KB)
constexpr int kKVTile = 2 * kKVPanel;     // [64 keys][128 dims] (16 KB)
constexpr int kPTile = 128 * 128;         // [128 rows][64 keys] bf16 (16 KB)
constexpr int kStages = 3;
constexpr int kTcThreads = 320;

// Cycle accounting / knock-out experiments (bench/attn_prof.py) are compiled in only with -DNRL_ATTN_PROFILE
// (NRL_ATTN_PROFILE=1 python -m nanorlhf_b200.csrc.build); production builds carry neither clock reads nor tests.
#ifdef NRL_ATTN_PROFILE
#define NRL_TWAIT(slot, ...)                                              \
  do {                                                                   \
    const long long t_ = clock64();                                     \
    mbar_wait(__VA_ARGS__);                                             \
    tw[slot] += clock64() - t_;                                         \
  } while (0)
#define NRL_DBG(bit) (p.dbg & (bit))
#else
#define NRL_TWAIT(slot, ...) mbar_wait(__VA_ARGS__)
#define NRL_DBG(bit) false
#endif

struct AttnTcParams {
  const int* cu_seqlens;
  __nv_bfloat16* out;
  float* lse;                 // [Hq, T_total]
  long o_stride_t;
  int num_seqs, total_tokens, Hq, G;
  float scale_log2;
  long long* prof;            // optional [grid][40] cycle counters (NRL_ATTN_PROFILE builds, bench/attn_prof.py)
  int dbg;                    // timing experiments only (NRL_ATTN_DBG): 1 skip S MMAs, 2 skip PV MMAs, 4 skip exps, 8 skip K/V loads
};

struct AttnTcSmem {
  static constexpr int kQ = 0;                                   // 2 tiles
  static constexpr int kK = 2 * kQTile;                          // kStages
  static constexpr int kV = kK + kStages * kKVTile;              // kStages
  static constexpr int kP = kV + kStages * kKVTile;              // [tile][buf]
  static constexpr int kBar = kP + 4 * kPTile;
  static constexpr int kTotal = kBar + 256;
};

__device__ bool tc_locate_block(const int* cu, int num_seqs, int blk, int& m_blk, int& seq_start, int& seq_len) {
  __shared__ int s_info[4];
  if (threadIdx.x == 0) {
    int acc = 0, found = 0;
    for (int s = 0; s < num_seqs; ++s) {
      const int a = cu[s], b = cu[s + 1];
      const int nb = (b - a + kTcBM - 1) / kTcBM;
      if (blk < acc + nb) {
        s_info[0] = 1; s_info[1] = blk - acc; s_info[2] = a; s_info[3] = b - a;
        found = 1;
        break;
      }
      acc += nb;
    }
    if (!found) s_info[0] = 0;
  }
  __syncthreads();
  if (!s_info[0]) return false;
  m_blk = s_info[1]; seq_start = s_info[2]; seq_len = s_info[3];
  return true;
}

__global__ void __launch_bounds__(kTcThreads, 1)
attn_fwd_tc_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                   const __grid_constant__ CUtensorMap tmV, AttnTcParams p) {
  using L = AttnTcSmem;
  extern __shared__ __align__(1024) uint8_t smem[];
  uint64_t* q_full = reinterpret_cast<uint64_t*>(smem + L::kBar);
  uint64_t* kv_full = q_full + 1;              // [kStages]
  uint64_t* kv_empty = kv_full + kStages;      // [kStages]
  uint64_t* s_full = kv_empty + kStages;       // [tile][2]
  uint64_t* p_ready = s_full + 4;              // [tile][2]
  uint64_t* pv_done = p_ready + 4;             // [tile][2]  PV_X of a block with parity b retired: P_X[b] reusable, O_X current
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(pv_done + 4);

  int m_blk, seq_start, seq_len;
  // heaviest (latest) query blocks are launched first: longest-processing-time order trims the causal tail
  if (!tc_locate_block(p.cu_seqlens, p.num_seqs, gridDim.x - 1 - blockIdx.x, m_blk, seq_start, seq_len)) return;
  const int head = blockIdx.y, kvh = head / p.G;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int q0 = m_blk * kTcBM;
  const int kv_blocks = (seq_len + kTcBN - 1) / kTcBN;
  // number of KV blocks each tile visits (causal); tile B may lie entirely behind the end of the sequence
  const int nA = min(kv_blocks, (q0 + 128 + kTcBN - 1) / kTcBN);
  const int nB = (q0 + 128 < seq_len) ? min(kv_blocks, (q0 + 256 + kTcBN - 1) / kTcBN) : 0;
  const int n_blocks = max(nA, nB);

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmQ);
    tma_prefetch_desc(&tmK);
    tma_prefetch_desc(&tmV);
  }
  if (warp == 1 && lane == 0) {
    mbar_init(q_full, 1);
    for (int i = 0; i < kStages; ++i) {
      mbar_init(&kv_full[i], 1);
      mbar_init(&kv_empty[i], 1);
    }
    for (int i = 0; i < 4; ++i) {
      mbar_init(&s_full[i], 1);
      mbar_init(&p_ready[i], 4);        // one arrive per softmax warp of the tile
      mbar_init(&pv_done[i], 1);
    }
    fence_mbar_init();
  }
  if (warp == 1) {
    __syncwarp();
    tmem_alloc(tmem_ptr, 512);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;
#ifdef NRL_ATTN_PROFILE
  long long tw[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  const long long t_begin = clock64();
#endif

  if (warp == 0) {
    // ============================== TMA producer ==============================
    if (elect_one()) {
      mbar_arrive_expect_tx(q_full, 2 * kQTile);
#pragma unroll
      for (int t = 0; t < 2; ++t) {
        tma_load_2d(smem + L::kQ + t * kQTile, &tmQ, q_full, head * kTcD, seq_start + q0 + t * 128);
        tma_load_2d(smem + L::kQ + t * kQTile + kQPanel, &tmQ, q_full, head * kTcD + 64, seq_start + q0 + t * 128);
      }
      int st = 0;
      uint32_t phase = 0;
      for (int j = 0; j < n_blocks; ++j) {
        NRL_TWAIT(0, &kv_empty[st], phase ^ 1);
        uint8_t* sk = smem + L::kK + st * kKVTile;
        uint8_t* sv = smem + L::kV + st * kKVTile;
        if (NRL_DBG(8)) {
          mbar_arrive(&kv_full[st]);
          if (++st == kStages) { st = 0; phase ^= 1; }
          continue;
        }
        mbar_arrive_expect_tx(&kv_full[st], 2 * kKVTile);
        const int row = seq_start + j * kTcBN;
        tma_load_2d(sk, &tmK, &kv_full[st], kvh * kTcD, row);
        tma_load_2d(sk + kKVPanel, &tmK, &kv_full[st], kvh * kTcD + 64, row);
        tma_load_2d(sv, &tmV, &kv_full[st], kvh * kTcD, row);
        tma_load_2d(sv + kKVPanel, &tmV, &kv_full[st], kvh * kTcD + 64, row);
        if (++st == kStages) { st = 0; phase ^= 1; }
      }
    }
  } else if (warp == 1) {
    // ============================== MMA issuer ==============================
    // elect.sync (not `lane == 0`) lets the compiler prove single-thread execution; otherwise every tcgen05.mma is
    // wrapped in an ELECT/branch loop whose latency exceeds these short (32-64 cycle) MMAs
    if (elect_one()) {
      constexpr uint32_t idesc_s = make_idesc(128, kTcBN, 1, 1);        // Q K^T : both K-major
      constexpr uint32_t idesc_o = make_idesc_bmn(128, kTcD);          // P V   : B = V is MN-major
      NRL_TWAIT(1, q_full, 0);
      tc_fence_after();
      int st_s = 0;                       // ring stage of the block whose S is issued next
      uint32_t ph_s = 0;
      auto issue_s = [&](int j) {
        NRL_TWAIT(2, &kv_full[st_s], ph_s);
        tc_fence_after();
        const uint32_t k_addr = smem_u32(smem + L::kK + st_s * kKVTile);
#pragma unroll
        for (int t = 0; t < 2; ++t) {
          if (j < (t == 0 ? nA : nB)) {
            const uint32_t q_addr = smem_u32(smem + L::kQ + t * kQTile);
            const uint32_t t_s = tmem_base + t * 128 + (j & 1) * kTcBN;
#pragma unroll
            for (int k = 0; k < 8; ++k)        // 8 x 16 dims; dims 0-63 in panel 0, 64-127 in panel 1
              if (!NRL_DBG(1)) umma_f16(t_s, make_smem_desc_sw128(q_addr + (k >> 2) * kQPanel + (k & 3) * 32),
                       make_smem_desc_sw128(k_addr + (k >> 2) * kKVPanel + (k & 3) * 32), idesc_s, k != 0 ? 1u : 0u);
            umma_commit(&s_full[t * 2 + (j & 1)]);
          }
        }
        if (++st_s == kStages) { st_s = 0; ph_s ^= 1; }
      };
      issue_s(0);
      int st = 0;
      for (int j = 0; j < n_blocks; ++j) {
        if (j + 1 < n_blocks) issue_s(j + 1);               // overlaps the softmax of block j
        const uint32_t v_addr = smem_u32(smem + L::kV + st * kKVTile);
#pragma unroll
        for (int t = 0; t < 2; ++t) {
          if (j < (t == 0 ? nA : nB)) {
            NRL_TWAIT(3, &p_ready[t * 2 + (j & 1)], (j >> 1) & 1);         // P_X,j in smem, O_X rescaled
            tc_fence_after();
            const uint32_t p_addr = smem_u32(smem + L::kP + (t * 2 + (j & 1)) * kPTile);
#pragma unroll
            for (int k = 0; k < 4; ++k)        // 4 x 16 keys
              if (!NRL_DBG(2)) umma_f16(tmem_base + 256 + t * 128, make_smem_desc_sw128(p_addr + k * 32),
                       make_smem_desc_sw128_mn(v_addr + k * 16 * 128, kKVPanel), idesc_o, (j | k) != 0 ? 1u : 0u);
            umma_commit(&pv_done[t * 2 + (j & 1)]);
          }
        }
        umma_commit(&kv_empty[st]);            // K_j / V_j smem reusable
        if (++st == kStages) st = 0;
      }
    }
    __syncwarp();
  } else {
    // ============================== softmax / correction / epilogue ==============================
    const int tile = (warp - 2) >> 2;                         // 0 = A, 1 = B
    const int n_mine = tile == 0 ? nA : nB;
    const int quad = warp & 3;
    const int r = quad * 32 + lane;                          // row in tile == TMEM lane
    const int row0 = q0 + tile * 128;                        // first row of the tile (within the sequence)
    const int row = row0 + r;
    const uint32_t lane_off = static_cast<uint32_t>(quad * 32) << 16;
    const uint32_t t_o = tmem_base + 256 + tile * 128 + lane_off;
    float m_run = -INFINITY, l_run = 0.f;
    for (int j = 0; j < n_mine; ++j) {
      const int b = j & 1;
      NRL_TWAIT(4, &s_full[tile * 2 + b], (j >> 1) & 1);
      tc_fence_after();
      const uint32_t t_s = tmem_base + tile * 128 + b * kTcBN + lane_off;
      uint32_t sv[2][32];
      tmem_ld_32x32b_x32(t_s, sv[0]);
      tmem_ld_32x32b_x32(t_s + 32, sv[1]);
      tmem_ld_wait();
      const int key0 = j * kTcBN;
      // only blocks on the diagonal of this tile and a ragged last block need masking (warpgroup-uniform test)
      if (key0 + kTcBN - 1 > row0 || key0 + kTcBN > seq_len) {
#pragma unroll
        for (int c = 0; c < 2; ++c)
#pragma unroll
          for (int i = 0; i < 32; ++i) {
            const int key = key0 + c * 32 + i;
            if (key >= seq_len || key > row) sv[c][i] = 0xff800000u;      // -inf
          }
      }
      float mx4[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};         // 4 independent chains
#pragma unroll
      for (int c = 0; c < 2; ++c)
#pragma unroll
        for (int i = 0; i < 32; ++i) mx4[i & 3] = fmaxf(mx4[i & 3], __uint_as_float(sv[c][i]));
      const float mx = fmaxf(fmaxf(mx4[0], mx4[1]), fmaxf(mx4[2], mx4[3])) * p.scale_log2;
      // lazy rescale: keep the stale maximum while the new one is < 2^8 above it (p <= 256 is harmless in bf16 and
      // l / O stay consistent with m_run); key 0 is live for every row so m_new is finite from block 0 on
      const float m_new = (mx - m_run > 8.f) ? mx : m_run;
      const float corr = exp2f(m_run - m_new);
      float ps4[4] = {0.f, 0.f, 0.f, 0.f};
      if (!NRL_DBG(4))
      // exponentials (in place: sv[c][i/2] <- bf16x2(p_i, p_i+1)); the wait below overlaps them
#pragma unroll
      for (int c = 0; c < 2; ++c)
#pragma unroll
        for (int i = 0; i < 32; i += 2) {
          const float p0 = exp2f(fmaf(__uint_as_float(sv[c][i]), p.scale_log2, -m_new));
          const float p1 = exp2f(fmaf(__uint_as_float(sv[c][i + 1]), p.scale_log2, -m_new));
          ps4[(i >> 1) & 3] += p0 + p1;
          sv[c][i / 2] = pack_bf16x2(p0, p1);
        }
      const float ps = (ps4[0] + ps4[1]) + (ps4[2] + ps4[3]);
      // P_X[b] was last read by PV_X(j-2).  pv_done[X][b] completes once per two blocks and cannot run ahead of this
      // thread (its next completion needs our p_ready), so the parity wait is unambiguous.
      if (j >= 2) {
        NRL_TWAIT(5, &pv_done[tile * 2 + b], ((j >> 1) - 1) & 1);
        tc_fence_after();
      }
      uint8_t* rowp = smem + L::kP + (tile * 2 + b) * kPTile + r * 128;
#pragma unroll
      for (int c = 0; c < 2; ++c)
#pragma unroll
        for (int q4 = 0; q4 < 4; ++q4) {
          const int ch = c * 4 + q4;           // 16-byte chunk of the 128-byte (64-key) row
          *reinterpret_cast<uint4*>(rowp + ((ch ^ (r & 7)) << 4)) =
              make_uint4(sv[c][q4 * 4], sv[c][q4 * 4 + 1], sv[c][q4 * 4 + 2], sv[c][q4 * 4 + 3]);
        }
      l_run = l_run * corr + ps;
      // ---- rescale O in TMEM when this row's maximum moved (rare with the lazy threshold; skipped warp-wide) ----
      // only then does this block depend on PV_X(j-1) having retired: the softmax never waits for the tensor pipe
      // on the common path
      const bool need = (j > 0) && (corr != 1.f);
      if (__any_sync(0xffffffffu, need)) {
        NRL_TWAIT(6, &pv_done[tile * 2 + (b ^ 1)], ((j - 1) >> 1) & 1);
        tc_fence_after();
#pragma unroll 1
        for (int c = 0; c < 4; ++c) {
          uint32_t ov[32];
          tmem_ld_32x32b_x32(t_o + c * 32, ov);
          tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < 32; ++i) ov[i] = __float_as_uint(__uint_as_float(ov[i]) * corr);
          tmem_st_32x32b_x32(t_o + c * 32, ov);
        }
        tmem_st_wait();
      }
      m_run = m_new;
      fence_proxy_async_smem();              // P writes visible to the tensor-core (async) proxy
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&p_ready[tile * 2 + b]);
    }
    // ---- epilogue: O / l -> bf16, LSE ----
    if (n_mine > 0) {
      NRL_TWAIT(7, &pv_done[tile * 2 + ((n_mine - 1) & 1)], ((n_mine - 1) >> 1) & 1);
      tc_fence_after();
      const float inv = l_run > 0.f ? 1.f / l_run : 0.f;
      const bool row_ok = row < seq_len;
      __nv_bfloat16* orow = p.out + static_cast<long>(seq_start + row) * p.o_stride_t + head * kTcD;
#pragma unroll 1
      for (int c = 0; c < 4; ++c) {
        uint32_t ov[32];
        tmem_ld_32x32b_x32(t_o + c * 32, ov);
        tmem_ld_wait();
        if (row_ok) {
#pragma unroll
          for (int q4 = 0; q4 < 4; ++q4) {
            uint4 o4;
            o4.x = pack_bf16x2(__uint_as_float(ov[q4 * 8]) * inv, __uint_as_float(ov[q4 * 8 + 1]) * inv);
            o4.y = pack_bf16x2(__uint_as_float(ov[q4 * 8 + 2]) * inv, __uint_as_float(ov[q4 * 8 + 3]) * inv);
            o4.z = pack_bf16x2(__uint_as_float(ov[q4 * 8 + 4]) * inv, __uint_as_float(ov[q4 * 8 + 5]) * inv);
            o4.w = pack_bf16x2(__uint_as_float(ov[q4 * 8 + 6]) * inv, __uint_as_float(ov[q4 * 8 + 7]) * inv);
            *reinterpret_cast<uint4*>(orow + c * 32 + q4 * 8) = o4;
          }
        }
      }
      if (row_ok && p.lse != nullptr)
        p.lse[static_cast<long>(head) * p.total_tokens + seq_start + row] = (m_run + log2f(l_run)) * 0.6931471805599453f;
    }
  }

#ifdef NRL_ATTN_PROFILE
  if (p.prof != nullptr && lane == 0 && (warp == 0 || warp == 1 || warp == 2 || warp == 6)) {
    long long* dst = p.prof + (static_cast<long>(blockIdx.y) * gridDim.x + blockIdx.x) * 40 + (warp == 0 ? 0 : warp == 1 ? 10 : warp == 2 ? 20 : 30);
    for (int i = 0; i < 8; ++i) dst[i] = tw[i];
    dst[8] = clock64() - t_begin;
    dst[9] = n_blocks;
  }
#endif
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

}  // namespace nrl

extern "C" cudaError_t nrl_attn_fwd_tc(const CUtensorMap* tmQ, const CUtensorMap* tmK, const CUtensorMap* tmV, void* out,
                                       float* lse, long o_stride_t, const int* cu, int num_seqs, int total, int Hq, int Hkv,
                                       float scale, cudaStream_t s, long long* prof) {
  using namespace nrl;
  if (total == 0) return cudaSuccess;
  AttnTcParams p;
  p.cu_seqlens = cu; p.out = static_cast<__nv_bfloat16*>(out); p.lse = lse; p.o_stride_t = o_stride_t;
  p.num_seqs = num_seqs; p.total_tokens = total; p.Hq = Hq; p.G = Hq / Hkv;
  p.scale_log2 = scale * 1.4426950408889634f;
  static const int dbg = getenv("NRL_ATTN_DBG") ? atoi(getenv("NRL_ATTN_DBG")) : 0;
  p.dbg = dbg;
  p.prof = prof;
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(attn_fwd_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, AttnTcSmem::kTotal);
    if (e != cudaSuccess) return e;
    configured = true;
  }
  dim3 grid(total / kTcBM + num_seqs, Hq);
  attn_fwd_tc_kernel<<<grid, kTcThreads, AttnTcSmem::kTotal, s>>>(*tmQ, *tmK, *tmV, p);
  return cudaGetLastError();
}
