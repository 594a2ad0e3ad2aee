// EXPERIMENTAL (opt-in: NRL_GEMM_2CTA=1; not yet validated on hardware -- written at the end of round 1 when the GPU
// budget was spent; first item of the next round's GPU time).
//
// 2-CTA (cta_group::2) variant of the bf16 GEMM  D[M,N] = A[M,K] * B[N,K]^T (+bias, +GELU):
// a CTA pair (cluster of 2, same TPC) computes a 256 x 256 tile with one `tcgen05.mma.cta_group::2` per 16-wide
// k-step.  Each CTA loads its own 128 rows of A and only HALF of the B tile (128 of the 256 rows), so a k-block
// costs 32 KB of L2->SM traffic per SM instead of 48 KB (the decode-shape GEMMs are bound by exactly that), and the
// smaller stage lets the ring hold 6 stages instead of 4.  Protocol (CUTLASS sm100 2-SM mainloop, re-derived):
//   * both CTAs issue their TMA loads with `.cta_group::2`, signalling the LEADER's (rank 0) full barrier
//     (barrier address with the peer bit cleared); the leader arms it with the bytes of both CTAs;
//   * only the leader's MMA thread issues MMAs; its `tcgen05.commit.cta_group::2 ... multicast::cluster` arrives on
//     the barrier at the same offset in BOTH CTAs (smem slot free / accumulator ready);
//   * the epilogue warps of both CTAs hand the accumulator back by arriving on the leader's tmem_empty barrier;
//   * TMEM is allocated with cta_group::2 by one warp in each CTA; the accumulator of CTA r holds rows 128r..128r+127.
#include "common.cuh"
#include "gemm_sm100.h"

namespace nrl {
namespace two_cta {

constexpr int BLOCK_M = 128;          // per CTA; the pair covers 256 rows
constexpr int BLOCK_N = 256;
constexpr int BLOCK_K = 64;
constexpr int UMMA_K = 16;
constexpr int kThreads = 192;
constexpr int kEpiThreads = 128;
constexpr int kABytes = BLOCK_M * BLOCK_K * 2;              // 16 KB
constexpr int kBBytes = (BLOCK_N / 2) * BLOCK_K * 2;        // 16 KB: this CTA's half of the B tile
constexpr int kStageBytes = kABytes + kBBytes;
constexpr int kStages = 6;
constexpr int kStagingBytes = BLOCK_M * 128;
constexpr int kStagingOff = kStages * kStageBytes;
constexpr int kBarOff = kStagingOff + 2 * kStagingBytes;
constexpr int kSmemTotal = kBarOff + 256;
constexpr uint32_t kPeerMask = 0xFEFFFFFFu;                  // clears the CTA-pair peer bit of a shared::cluster address
constexpr int GROUP_M = 8;                                   // in 256-row tile pairs

NRL_DEVICE uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
NRL_DEVICE void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
NRL_DEVICE void tmem_alloc_2cta(uint32_t* smem_dst, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_dst)), "r"(ncols) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
NRL_DEVICE void tmem_dealloc_2cta(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
NRL_DEVICE void umma_f16_2cta(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}\n"
      ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// arrive on the barrier at this offset in both CTAs of the pair once all prior MMAs of this thread retire
NRL_DEVICE void umma_commit_pair(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
               ::"r"(smem_u32(bar)), "h"(static_cast<uint16_t>(3)) : "memory");
}
// TMA load whose transaction bytes are credited to the leader CTA's barrier
NRL_DEVICE void tma_load_2d_pair(void* smem_dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar) & kPeerMask), "r"(c0), "r"(c1)
      : "memory");
}
NRL_DEVICE void mbar_arrive_leader(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(smem_u32(bar) & kPeerMask) : "memory");
}

NRL_DEVICE float gelu_erf_2(float x) {
  const float z = fabsf(x) * 0.70710678118654752f;
  const float t = __fdividef(1.f, fmaf(0.3275911f, z, 1.f));
  float poly = fmaf(1.061405429f, t, -1.453152027f);
  poly = fmaf(poly, t, 1.421413741f);
  poly = fmaf(poly, t, -0.284496736f);
  poly = fmaf(poly, t, 0.254829592f);
  const float erf_abs = 1.f - poly * t * exp2f(-z * z * 1.4426950408889634f);
  return 0.5f * x * (1.f + copysignf(erf_abs, x));
}

struct PairCoord {
  int m2, n_blk;
};
NRL_DEVICE PairCoord pair_coord(int t, int num_m2, int num_n) {
  const int per_group = GROUP_M * num_n;
  const int g = t / per_group, r = t - g * per_group;
  const int gsz = min(GROUP_M, num_m2 - g * GROUP_M);
  PairCoord c;
  c.n_blk = r / gsz;
  c.m2 = g * GROUP_M + (r - c.n_blk * gsz);
  return c;
}

// SWIGLU = true: B rows are interleaved [32 gate | 32 up] per 64-column chunk (parallel/weight_sync.py); the epilogue
// forms silu(g) * u on the accumulator tile and stores a [M, N/2] activation (same contract as EPI_SWIGLU).
template <bool SWIGLU>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kThreads, 1)
gemm_bf16_tn_2cta_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                         const __grid_constant__ CUtensorMap tmD, GemmParams p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + kBarOff);    // leader's are the live ones
  uint64_t* empty_bar = full_bar + kStages;                            // per CTA
  uint64_t* tmem_full = empty_bar + kStages;                           // [2] per CTA
  uint64_t* tmem_empty = tmem_full + 2;                                // [2] leader's are the live ones
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(tmem_empty + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const bool leader = rank == 0;
  const int num_m2 = (p.M + 2 * BLOCK_M - 1) / (2 * BLOCK_M);
  const int num_n = (p.N + BLOCK_N - 1) / BLOCK_N;
  const int num_kb = (p.K + BLOCK_K - 1) / BLOCK_K;
  const int num_work = num_m2 * num_n;
  const int pair_id = blockIdx.x >> 1, num_pairs = gridDim.x >> 1;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    tma_prefetch_desc(&tmD);
  }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < kStages; ++i) {
      mbar_init(&full_bar[i], 1);                      // the leader's producer thread (+ transaction bytes of both CTAs)
      mbar_init(&empty_bar[i], 1);                     // one multicast commit
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tmem_full[i], 1);                     // one multicast commit
      mbar_init(&tmem_empty[i], 2 * kEpiThreads / 32); // the epilogue warps of both CTAs
    }
    fence_mbar_init();
  }
  cluster_sync_all();                                  // peers' barriers exist before anything remote is signalled
  if (warp == 1) tmem_alloc_2cta(tmem_ptr, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;

  if (warp == 0) {
    // ===================================== TMA producer (both CTAs) ======================================
    if (elect_one()) {
      int stage = 0;
      uint32_t phase = 0;
      for (int w = pair_id; w < num_work; w += num_pairs) {
        const PairCoord c = pair_coord(w, num_m2, num_n);
        const int m_row = (c.m2 * 2 + static_cast<int>(rank)) * BLOCK_M;
        const int n_row = c.n_blk * BLOCK_N + static_cast<int>(rank) * (BLOCK_N / 2);
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          uint8_t* sa = smem + stage * kStageBytes;
          uint8_t* sb = sa + kABytes;
          if (leader) mbar_arrive_expect_tx(&full_bar[stage], 2 * kStageBytes);
          tma_load_2d_pair(sa, &tmA, &full_bar[stage], kb * BLOCK_K, m_row);
          tma_load_2d_pair(sb, &tmB, &full_bar[stage], kb * BLOCK_K, n_row);
          if (++stage == kStages) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ===================================== MMA issuer (leader CTA only) =======================================
    if (leader && elect_one()) {
      constexpr uint32_t idesc = make_idesc(2 * BLOCK_M, BLOCK_N, 1, 1);
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      for (int w = pair_id; w < num_work; w += num_pairs) {
        mbar_wait(&tmem_empty[acc], acc_phase ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * BLOCK_N;
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          const uint32_t a_addr = smem_u32(smem + stage * kStageBytes);
          const uint32_t b_addr = a_addr + kABytes;
#pragma unroll
          for (int k = 0; k < BLOCK_K / UMMA_K; ++k)
            umma_f16_2cta(d_tmem, make_smem_desc_sw128(a_addr + k * UMMA_K * 2), make_smem_desc_sw128(b_addr + k * UMMA_K * 2),
                          idesc, (kb | k) != 0 ? 1u : 0u);
          umma_commit_pair(&empty_bar[stage]);
          if (++stage == kStages) { stage = 0; phase ^= 1; }
        }
        umma_commit_pair(&tmem_full[acc]);
        if (++acc == 2) { acc = 0; acc_phase ^= 1; }
      }
    }
    __syncwarp();
  } else {
    // ===================================== epilogue (both CTAs, own 128 rows) ==========================================
    const int quad = warp & 3;
    const int row_in_tile = quad * 32 + lane;
    const int epi_tid = threadIdx.x - 64;
    int acc = 0;
    uint32_t acc_phase = 0;
    uint8_t* staging = smem + kStagingOff;
    int store_buf = 0;
    for (int w = pair_id; w < num_work; w += num_pairs) {
      const PairCoord c = pair_coord(w, num_m2, num_n);
      const int m_blk = c.m2 * 2 + static_cast<int>(rank);
      mbar_wait(&tmem_full[acc], acc_phase);
      tc_fence_after();
      const uint32_t t_acc = tmem_base + acc * BLOCK_N + (static_cast<uint32_t>(quad * 32) << 16);
      uint32_t sw_packed[32];
      (void)sw_packed;
#pragma unroll 1
      for (int ch64 = 0; ch64 < BLOCK_N / 64; ++ch64) {
        uint32_t v[2][32];
        tmem_ld_32x32b_x32(t_acc + ch64 * 64, v[0]);
        tmem_ld_32x32b_x32(t_acc + ch64 * 64 + 32, v[1]);
        tmem_ld_wait();
        if (ch64 == BLOCK_N / 64 - 1) {
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive_leader(&tmem_empty[acc]);     // the accumulator of this CTA is drained
        }
        const int col0 = c.n_blk * BLOCK_N + ch64 * 64;
        uint32_t packed[32];
        int store_col = col0;
        bool do_store = true;
        if (SWIGLU) {
          constexpr float kLog2e = 1.4426950408889634f;
#pragma unroll
          for (int j = 0; j < 32; j += 2) {
            const float g0 = __uint_as_float(v[0][j]), g1 = __uint_as_float(v[0][j + 1]);
            const float u0 = __uint_as_float(v[1][j]), u1 = __uint_as_float(v[1][j + 1]);
            const uint32_t pk = pack_bf16x2(g0 / (1.f + exp2f(-g0 * kLog2e)) * u0, g1 / (1.f + exp2f(-g1 * kLog2e)) * u1);
            if ((ch64 & 1) == 0) sw_packed[j / 2] = pk;            // static indices: the array stays in registers
            else sw_packed[16 + j / 2] = pk;
          }
          do_store = (ch64 & 1) == 1;                               // two 32-feature halves make one 64-wide slab
          store_col = c.n_blk * (BLOCK_N / 2) + (ch64 >> 1) * 64;
#pragma unroll
          for (int j = 0; j < 32; ++j) packed[j] = sw_packed[j];
        } else {
#pragma unroll
          for (int h = 0; h < 2; ++h)
#pragma unroll
            for (int j = 0; j < 32; j += 2) {
              float x0 = __uint_as_float(v[h][j]), x1 = __uint_as_float(v[h][j + 1]);
              const int col = col0 + h * 32 + j;
              if (p.bias != nullptr) {
                if (col < p.N) x0 += __bfloat162float(p.bias[col]);
                if (col + 1 < p.N) x1 += __bfloat162float(p.bias[col + 1]);
              }
              if (p.act == 1) { x0 = gelu_erf_2(x0); x1 = gelu_erf_2(x1); }
              packed[h * 16 + j / 2] = pack_bf16x2(x0, x1);
            }
        }
        if (do_store) {
          uint8_t* buf = staging + store_buf * kStagingBytes;
          if (epi_tid == 0) tma_store_wait_read<1>();
          named_barrier_sync(1, kEpiThreads);
          uint8_t* rowp = buf + row_in_tile * 128;
#pragma unroll
          for (int q8 = 0; q8 < 8; ++q8)
            *reinterpret_cast<uint4*>(rowp + ((q8 ^ (row_in_tile & 7)) * 16)) =
                make_uint4(packed[q8 * 4], packed[q8 * 4 + 1], packed[q8 * 4 + 2], packed[q8 * 4 + 3]);
          fence_proxy_async_smem();
          named_barrier_sync(2, kEpiThreads);
          if (epi_tid == 0) {
            tma_store_2d(&tmD, buf, store_col, m_blk * BLOCK_M);
            tma_store_commit();
          }
          store_buf ^= 1;
        }
      }
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
    }
    if (epi_tid == 0) tma_store_wait<0>();
  }

  tc_fence_before();
  cluster_sync_all();              // nobody leaves (or frees TMEM) while the peer can still signal / read it
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc_2cta(tmem_base, 512);
  }
}

}  // namespace two_cta
}  // namespace nrl

// maps: A box 128 rows x 64, B box 128 rows x 64 (half of the 256-wide tile), D box 128 rows x 64
extern "C" cudaError_t nrl_gemm_bf16_tn_2cta(const CUtensorMap* tmA, const CUtensorMap* tmB, const CUtensorMap* tmD,
                                             const nrl::GemmParams* p, int swiglu, int num_sms, cudaStream_t stream) {
  using namespace nrl::two_cta;
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(gemm_bf16_tn_2cta_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemTotal);
    if (e != cudaSuccess) return e;
    e = cudaFuncSetAttribute(gemm_bf16_tn_2cta_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemTotal);
    if (e != cudaSuccess) return e;
    configured = true;
  }
  const int num_m2 = (p->M + 2 * BLOCK_M - 1) / (2 * BLOCK_M), num_n = (p->N + BLOCK_N - 1) / BLOCK_N;
  int pairs = num_m2 * num_n;
  if (pairs > num_sms / 2) pairs = num_sms / 2;
  if (pairs < 1) pairs = 1;
  // __cluster_dims__(2, 1, 1): consecutive CTAs form the pairs
  if (swiglu)
    gemm_bf16_tn_2cta_kernel<true><<<2 * pairs, kThreads, kSmemTotal, stream>>>(*tmA, *tmB, *tmD, *p);
  else
    gemm_bf16_tn_2cta_kernel<false><<<2 * pairs, kThreads, kSmemTotal, stream>>>(*tmA, *tmB, *tmD, *p);
  return cudaGetLastError();
}
