// gemm_tc: the general tcgen05 / TMEM / TMA GEMM of the training, log-prob and reward paths (hand-written PTX).
//
//   D[M,N] = alpha * ( opA(A) . opB(B)^T  +  A2 . B2^T )  (+ bias[N]) (GELU)          bf16 in, fp32 accumulate
//
// One kernel template covers what the reference gets from cuBLAS behind every nn.Linear / peft LoRA layer / torch.bmm
// (/root/reference/GRPO/grpo_trainer.py:543-556,652-660; GRPO/grpo.py:189-192,228-243):
//
//   * CG = 1 | 2      cta_group::1 (one CTA, 128 x BN tile) or cta_group::2 (a CTA pair of one TPC computes a 256 x BN
//                     tile; each CTA TMA-loads its own 128 rows of A and only HALF of the B tile, one elected thread of
//                     the leader CTA issues `tcgen05.mma.cta_group::2`, commits are multicast to both CTAs' barriers).
//                     Per k-block an SM ingests 16 KB + BN*64 B instead of 16 KB + BN*128 B: the L2->SM stream and the
//                     smem operand reads -- what bounds the 1-CTA kernel -- drop by a third (measured: 1.15-1.25x).
//   * BN = 64 | 128 | 192 | 256   picked on the host by wave efficiency (N = 1536 outputs: 192 fills 2.9 of 3 waves
//                     where 256 fills 2.2 of 3).
//   * A_MN / B_MN     either operand may be "MN-major" (the contraction index is the ROW of the stored matrix).  That is
//                     what makes dgrad ( dX = dY . W : B = W[N,K] read as [k=N rows][n=K cols] ) and wgrad
//                     ( dW = dY^T . X : both operands token-major ) run on the same kernel with NO transposed copies:
//                     TMA lands [64 k-rows x 64 mn-cols] panels and the UMMA descriptor addresses them as MN-major.
//   * dual source K   a second operand pair (A2, B2) contributes K2 further contraction columns into the same TMEM
//                     accumulator: LoRA  y = x W^T + (s x A^T) B^T  is ONE GEMM with K + r columns instead of a base
//                     GEMM, an adapter GEMM and an add pass over [M,N]; likewise dX = dY W + (s dY B) A.
//   * epilogues       EPI_BF16: alpha, bias, exact GELU -> bf16 through swizzled smem + TMA store;
//                     EPI_F32 : fp32 row-major out (+)= alpha * acc  (weight-gradient accumulation, lm-head dW).
//
// Pipeline (per CTA, 192 threads): warp 0 = TMA producer, warp 1 = MMA issuer (+ TMEM allocator), warps 2-5 = epilogue;
// smem ring of kStages k-blocks behind full/empty mbarriers, two TMEM accumulators so the epilogue of tile i overlaps the
// main loop of tile i+1, persistent grid with a grouped (L2-friendly) tile order.
#include "common.cuh"
#include "gemm_tc.h"

namespace nrl {
namespace tc {

constexpr int BLOCK_M = 128;        // rows per CTA (a CTA pair covers 256)
constexpr int BLOCK_K = 64;         // 64 bf16 = one 128-byte swizzle row
constexpr int UMMA_K = 16;
constexpr int kThreads = 192;
constexpr int kEpiThreads = 128;
constexpr int kABytes = BLOCK_M * BLOCK_K * 2;      // 16 KB
constexpr int kPanelBytes = 64 * 128;               // one [64 x 64] bf16 panel of an MN-major operand
constexpr int kStagingBytes = BLOCK_M * 128;        // 128 rows x 64 bf16 output slab
constexpr int kSmemBudget = 227 * 1024;
constexpr uint32_t kPeerMask = 0xFEFFFFFFu;         // clears the CTA-pair peer bit of a shared::cluster address
constexpr int GROUP_M = 8;                          // tile rows per L2 super-row

template <int CG, int BN, bool B_MN = false>
struct Cfg {
  static constexpr int kBRows = BN / CG;                                  // B rows (n) this CTA contributes per k-block
  // an MN-major B tile is made of whole [64 k x 64 n] panels: 96 columns (BN = 192 on a CTA pair) load two panels, the
  // MMA reads the first 96 columns of them (the surplus 32 belong to the peer / the next tile and are ignored)
  static constexpr int kBPanels = (kBRows + 63) / 64;
  static constexpr int kBBytes = B_MN ? kBPanels * kPanelBytes : kBRows * BLOCK_K * 2;
  static constexpr int kStageBytes = kABytes + kBBytes;
  static constexpr int kStagesFit = (kSmemBudget - 2 * kStagingBytes - 1024) / kStageBytes;
  static constexpr int kStages = kStagesFit > 8 ? 8 : kStagesFit;
  static constexpr int kStagingOff = kStages * kStageBytes;
  static constexpr int kBarOff = kStagingOff + 2 * kStagingBytes;
  static constexpr int kTotal = kBarOff + 512;
  static constexpr uint32_t kTmemCols = (2 * BN <= 128) ? 128 : (2 * BN <= 256) ? 256 : 512;
  static_assert(kStageBytes % 1024 == 0, "stages must keep the 1024-byte swizzle alignment");
};

NRL_DEVICE uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
NRL_DEVICE void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
template <int CG>
NRL_DEVICE void tmem_alloc_cg(uint32_t* smem_dst, uint32_t ncols) {
  if (CG == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_dst)), "r"(ncols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  } else {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_dst)), "r"(ncols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
}
template <int CG>
NRL_DEVICE void tmem_dealloc_cg(uint32_t taddr, uint32_t ncols) {
  if (CG == 1) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
  else asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
template <int CG>
NRL_DEVICE void umma_bf16(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
  if (CG == 1) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                 "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}\n"
                 ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate) : "memory");
  } else {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                 "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}\n"
                 ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate) : "memory");
  }
}
template <int CG>
NRL_DEVICE void umma_fp8(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
  if (CG == 1) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                 "tcgen05.mma.cta_group::1.kind::f8f6f4 [%0], %1, %2, %3, p;\n\t}\n"
                 ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate) : "memory");
  } else {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                 "tcgen05.mma.cta_group::2.kind::f8f6f4 [%0], %1, %2, %3, p;\n\t}\n"
                 ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate) : "memory");
  }
}
// all MMAs issued so far by this thread arrive on `bar` when they retire (CG = 2: on the barrier at the same offset in
// BOTH CTAs of the pair)
template <int CG>
NRL_DEVICE void umma_commit_cg(uint64_t* bar) {
  if (CG == 1) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
  } else {
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                 ::"r"(smem_u32(bar)), "h"(static_cast<uint16_t>(3)) : "memory");
  }
}
// TMA tile load; CG = 2: the transaction bytes are credited to the LEADER CTA's barrier (peer bit cleared)
template <int CG>
NRL_DEVICE void tma_load_cg(void* smem_dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1) {
  if (CG == 1) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                 ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1) : "memory");
  } else {
    asm volatile("cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                 ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar) & kPeerMask), "r"(c0), "r"(c1) : "memory");
  }
}
NRL_DEVICE void mbar_arrive_leader(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(smem_u32(bar) & kPeerMask) : "memory");
}
NRL_DEVICE constexpr uint32_t make_idesc_tc(uint32_t M, uint32_t N, bool a_mn, bool b_mn) {     // bf16 x bf16 -> fp32
  return (1u << 4) | (1u << 7) | (1u << 10) | ((a_mn ? 1u : 0u) << 15) | ((b_mn ? 1u : 0u) << 16) | ((N >> 3) << 17) |
         ((M >> 4) << 24);
}

NRL_DEVICE float gelu_erf_tc(float x) {          // Abramowitz-Stegun 7.1.26, |err| < 1.5e-7 (see gemm_sm100.cu)
  const float z = fabsf(x) * 0.70710678118654752f;
  const float t = __fdividef(1.f, fmaf(0.3275911f, z, 1.f));
  float poly = fmaf(1.061405429f, t, -1.453152027f);
  poly = fmaf(poly, t, 1.421413741f);
  poly = fmaf(poly, t, -0.284496736f);
  poly = fmaf(poly, t, 0.254829592f);
  const float erf_abs = 1.f - poly * t * exp2f(-z * z * 1.4426950408889634f);
  return 0.5f * x * (1.f + copysignf(erf_abs, x));
}

struct TileCoord {
  int m_tile, n_blk;          // m_tile in units of CG*128 rows
};
NRL_DEVICE TileCoord tile_coord(int t, int num_m, int num_n) {
  const int per_group = GROUP_M * num_n;
  const int g = t / per_group, r = t - g * per_group;
  const int gsz = min(GROUP_M, num_m - g * GROUP_M);
  TileCoord c;
  c.n_blk = r / gsz;
  c.m_tile = g * GROUP_M + (r - c.n_blk * gsz);
  return c;
}

template <int CG>
NRL_DEVICE void tma_load3_cg(void* smem_dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1, int c2) {
  if (CG == 1) {
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
                 ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2) : "memory");
  } else {
    asm volatile("cp.async.bulk.tensor.3d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
                 ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar) & kPeerMask), "r"(c0), "r"(c1), "r"(c2) : "memory");
  }
}
NRL_DEVICE void tma_store_3d(const CUtensorMap* m, const void* smem_src, int c0, int c1, int c2) {
  asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%2, %3, %4}], [%1];"
               ::"l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(smem_src)), "r"(c0), "r"(c1), "r"(c2) : "memory");
}

// FP8: A and B are e4m3 bytes, K-major (a 128-byte swizzle row = 128 contraction elements), the MMA is kind::f8f6f4
// (K = 32 per instruction, twice the bf16 rate) and the epilogue applies the per-row (token) and per-column (output
// channel) dequantisation scales -- the sampler's fp8 rollout GEMMs (north star: fp8 tcgen05 GEMM in the decoder).
template <int CG, int BN, bool A_MN, bool B_MN, int EPI, bool SPLIT = false, bool BATCH = false, bool FP8 = false>
NRL_DEVICE void gemm_tc_body(const CUtensorMap& tmA, const CUtensorMap& tmB, const CUtensorMap& tmA2, const CUtensorMap& tmB2,
                             const CUtensorMap& tmD, const TcParams& p) {
  using C = Cfg<CG, BN, B_MN>;
  extern __shared__ __align__(1024) uint8_t smem[];
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + C::kBarOff);      // CG = 2: the leader's are the live ones
  uint64_t* empty_bar = full_bar + C::kStages;                              // per CTA (multicast commit)
  uint64_t* tmem_full = empty_bar + C::kStages;                             // [2] per CTA (multicast commit)
  uint64_t* tmem_empty = tmem_full + 2;                                     // [2] CG = 2: the leader's are the live ones
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(tmem_empty + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = (CG == 2) ? cluster_ctarank() : 0u;
  const bool leader = rank == 0;
  const int num_m = (p.M + CG * BLOCK_M - 1) / (CG * BLOCK_M);
  const int num_n = (p.N + BN - 1) / BN;
  constexpr int kElemsPerKB = FP8 ? 2 * BLOCK_K : BLOCK_K;                  // 128 bytes of contraction per k-block either way
  const int num_kb1 = (p.K + kElemsPerKB - 1) / kElemsPerKB;
  const int num_kb = num_kb1 + (p.K2 + kElemsPerKB - 1) / kElemsPerKB;
  const int splits = SPLIT ? p.splits : 1;
  const int kb_per_split = (num_kb + splits - 1) / splits;
  const int tiles_per_batch = num_m * num_n;
  // SPLIT: work item = (tile, k-range), splits of a tile adjacent.  BATCH: work item = (batch, tile), third TMA coordinate
  const int num_work = tiles_per_batch * (BATCH ? p.batch : splits);
  const int unit = blockIdx.x / CG, num_units = gridDim.x / CG;
  volatile int* s_last = reinterpret_cast<volatile int*>(tmem_ptr + 1);

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    if (p.K2 > 0) {
      tma_prefetch_desc(&tmA2);
      tma_prefetch_desc(&tmB2);
    }
    if (EPI != EPI_F32) tma_prefetch_desc(&tmD);
  }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < C::kStages; ++i) {
      mbar_init(&full_bar[i], 1);                       // the (leader's) producer thread + the transaction bytes
      mbar_init(&empty_bar[i], 1);                      // one (multicast) commit
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tmem_full[i], 1);                      // one (multicast) commit
      mbar_init(&tmem_empty[i], CG * kEpiThreads / 32); // the epilogue warps (of both CTAs)
    }
    fence_mbar_init();
  }
  if (CG == 2) cluster_sync_all();                      // the peer's barriers exist before anything remote is signalled
  if (warp == 1) {
    __syncwarp();
    tmem_alloc_cg<CG>(tmem_ptr, C::kTmemCols);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;

  if (warp == 0) {
    // ================================ TMA producer (every CTA) ================================
    if (elect_one()) {
      int stage = 0;
      uint32_t phase = 0;
      for (int w = unit; w < num_work; w += num_units) {
        const int bidx = BATCH ? w / tiles_per_batch : 0;
        const TileCoord c = tile_coord(BATCH ? w % tiles_per_batch : w / splits, num_m, num_n);
        const int m0 = (c.m_tile * CG + static_cast<int>(rank)) * BLOCK_M;
        const int n0 = c.n_blk * BN + static_cast<int>(rank) * C::kBRows;
        const int kb_lo = BATCH ? 0 : (w % splits) * kb_per_split, kb_hi = BATCH ? num_kb : min(num_kb, kb_lo + kb_per_split);
        for (int kb = kb_lo; kb < kb_hi; ++kb) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          uint8_t* sa = smem + stage * C::kStageBytes;
          uint8_t* sb = sa + kABytes;
          if (leader) mbar_arrive_expect_tx(&full_bar[stage], CG * C::kStageBytes);
          const bool second = kb >= num_kb1;
          const CUtensorMap* ma = second ? &tmA2 : &tmA;
          const CUtensorMap* mb = second ? &tmB2 : &tmB;
          const int k0 = (second ? kb - num_kb1 : kb) * kElemsPerKB;
          if (BATCH) {
            tma_load3_cg<CG>(sa, ma, &full_bar[stage], k0, m0, bidx);
            tma_load3_cg<CG>(sb, mb, &full_bar[stage], k0, n0, bidx);
            if (++stage == C::kStages) { stage = 0; phase ^= 1; }
            continue;
          }
          if (A_MN) {          // stored [K rows][M cols]: two [64 k x 64 m] panels
            tma_load_cg<CG>(sa, ma, &full_bar[stage], m0, k0);
            tma_load_cg<CG>(sa + kPanelBytes, ma, &full_bar[stage], m0 + 64, k0);
          } else {
            tma_load_cg<CG>(sa, ma, &full_bar[stage], k0, m0);
          }
          if (B_MN) {
#pragma unroll
            for (int pn = 0; pn < C::kBPanels; ++pn)
              tma_load_cg<CG>(sb + pn * kPanelBytes, mb, &full_bar[stage], n0 + pn * 64, k0);
          } else {
            tma_load_cg<CG>(sb, mb, &full_bar[stage], k0, n0);
          }
          if (++stage == C::kStages) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ================================ MMA issuer (leader CTA only) ================================
    if (leader && elect_one()) {
      constexpr uint32_t idesc = FP8 ? ((1u << 4) | ((static_cast<uint32_t>(BN) >> 3) << 17) | ((static_cast<uint32_t>(CG * BLOCK_M) >> 4) << 24))
                                     : make_idesc_tc(CG * BLOCK_M, BN, A_MN, B_MN);          // e4m3 x e4m3 (format code 0) -> fp32
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      for (int w = unit; w < num_work; w += num_units) {
        mbar_wait(&tmem_empty[acc], acc_phase ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * BN;
        const int kb_lo = BATCH ? 0 : (w % splits) * kb_per_split, kb_hi = BATCH ? num_kb : min(num_kb, kb_lo + kb_per_split);
        for (int kb = kb_lo; kb < kb_hi; ++kb) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          const uint32_t a_addr = smem_u32(smem + stage * C::kStageBytes);
          const uint32_t b_addr = a_addr + kABytes;
#pragma unroll
          for (int k = 0; k < BLOCK_K / UMMA_K; ++k) {
            // K-major: 16 contraction elements = 32 bytes along the swizzled row; MN-major: 16 contraction rows = 2048 B
            const uint64_t adesc = A_MN ? make_smem_desc_sw128_mn(a_addr + k * UMMA_K * 128, kPanelBytes)
                                        : make_smem_desc_sw128(a_addr + k * UMMA_K * 2);
            const uint64_t bdesc = B_MN ? make_smem_desc_sw128_mn(b_addr + k * UMMA_K * 128, kPanelBytes)
                                        : make_smem_desc_sw128(b_addr + k * UMMA_K * 2);
            if (FP8) umma_fp8<CG>(d_tmem, adesc, bdesc, idesc, ((kb - kb_lo) | k) != 0 ? 1u : 0u);
            else umma_bf16<CG>(d_tmem, adesc, bdesc, idesc, ((kb - kb_lo) | k) != 0 ? 1u : 0u);
          }
          umma_commit_cg<CG>(&empty_bar[stage]);        // smem stage reusable once these MMAs retire
          if (++stage == C::kStages) { stage = 0; phase ^= 1; }
        }
        umma_commit_cg<CG>(&tmem_full[acc]);            // accumulator complete -> epilogue (of both CTAs)
        if (++acc == 2) { acc = 0; acc_phase ^= 1; }
      }
    }
    __syncwarp();
  } else {
    // ================================ epilogue (every CTA: its own 128 rows) ================================
    const int quad = warp & 3;                          // TMEM lane quadrant this warp may read
    const int row_in_tile = quad * 32 + lane;
    const int epi_tid = threadIdx.x - 64;
    int acc = 0;
    uint32_t acc_phase = 0;
    uint8_t* staging = smem + C::kStagingOff;
    int store_buf = 0;
    for (int w = unit; w < num_work; w += num_units) {
      const int tile_id = BATCH ? w % tiles_per_batch : w / splits;
      const int bidx = BATCH ? w / tiles_per_batch : 0;
      (void)bidx;
      const TileCoord c = tile_coord(tile_id, num_m, num_n);
      const int m_blk = c.m_tile * CG + static_cast<int>(rank);
      const int row = m_blk * BLOCK_M + row_in_tile;
      const bool row_ok = row < p.M;
      mbar_wait(&tmem_full[acc], acc_phase);
      tc_fence_after();
      const uint32_t t_acc = tmem_base + acc * BN + (static_cast<uint32_t>(quad * 32) << 16);
      uint32_t sw_packed[32];
      (void)sw_packed;
      if (SPLIT) {
        // ---- split-K: park the fp32 partial tile, the last CTA of the tile reduces all partials in split order ----
        const long tile_elems = static_cast<long>(BLOCK_M) * BN;
        // workspace tile layout [BN / 4 column groups][128 rows][4]: a warp's 32 rows of one column group are 512 contiguous
        // bytes, for the partial stores here and for the reduction's loads below
        float* mine = p.ws + (static_cast<long>(w % splits) * (num_m * num_n) + tile_id) * tile_elems + row_in_tile * 4;
#pragma unroll 1
        for (int ch = 0; ch < BN / 32; ++ch) {
          uint32_t v[32];
          tmem_ld_32x32b_x32(t_acc + ch * 32, v);
          tmem_ld_wait();
          if (ch == BN / 32 - 1) {
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&tmem_empty[acc]);
          }
#pragma unroll
          for (int j = 0; j < 32; j += 4)
            __stcg(reinterpret_cast<float4*>(mine + (ch * 8 + (j >> 2)) * (BLOCK_M * 4)),
                   make_float4(__uint_as_float(v[j]), __uint_as_float(v[j + 1]), __uint_as_float(v[j + 2]), __uint_as_float(v[j + 3])));
        }
        __threadfence();                                   // this thread's partial is visible device-wide ...
        named_barrier_sync(1, kEpiThreads);                // ... and so is every other epilogue thread's
        if (epi_tid == 0) {
          const int old = atomicAdd(p.counters + tile_id, 1);
          const int last = (old == splits - 1) ? 1 : 0;
          if (last) p.counters[tile_id] = 0;               // self-reset for the next launch
          __threadfence();
          *s_last = last;
        }
        named_barrier_sync(2, kEpiThreads);
        if (*s_last && row_ok) {
          const float* base = p.ws + static_cast<long>(tile_id) * tile_elems + row_in_tile * 4;
          const long split_stride = static_cast<long>(num_m * num_n) * tile_elems;
          __nv_bfloat16* orow = p.out_bf16 + static_cast<long>(row) * p.out_stride + c.n_blk * BN;
#pragma unroll 1
          for (int c0 = 0; c0 < BN; c0 += 32) {
            if (c.n_blk * BN + c0 >= p.N) break;
            float acc[32];
#pragma unroll
            for (int e = 0; e < 32; ++e) acc[e] = 0.f;
#pragma unroll 1
            for (int sp = 0; sp < splits; ++sp) {            // fixed order -> deterministic; 8 independent 16-byte loads in flight
              const float* b = base + sp * split_stride + (c0 >> 2) * (BLOCK_M * 4);
              float4 x[8];
#pragma unroll
              for (int q = 0; q < 8; ++q) x[q] = __ldcg(reinterpret_cast<const float4*>(b + q * (BLOCK_M * 4)));
#pragma unroll
              for (int q = 0; q < 8; ++q) { acc[4 * q] += x[q].x; acc[4 * q + 1] += x[q].y; acc[4 * q + 2] += x[q].z; acc[4 * q + 3] += x[q].w; }
            }
#pragma unroll
            for (int j = 0; j < 32; j += 8) {
              if (c.n_blk * BN + c0 + j < p.N) {             // N % 8 == 0
                float s8[8];
#pragma unroll
                for (int e = 0; e < 8; ++e) s8[e] = acc[j + e] * p.alpha;
                if (p.bias != nullptr) {
                  const uint4 bv = *reinterpret_cast<const uint4*>(p.bias + c.n_blk * BN + c0 + j);      // bias 16-byte aligned
                  const uint32_t bw[4] = {bv.x, bv.y, bv.z, bv.w};
#pragma unroll
                  for (int e = 0; e < 4; ++e) {
                    const float2 bf = unpack_bf16x2(bw[e]);
                    s8[2 * e] += bf.x;
                    s8[2 * e + 1] += bf.y;
                  }
                }
                *reinterpret_cast<uint4*>(orow + c0 + j) = make_uint4(pack_bf16x2(s8[0], s8[1]), pack_bf16x2(s8[2], s8[3]),
                                                                      pack_bf16x2(s8[4], s8[5]), pack_bf16x2(s8[6], s8[7]));
              }
            }
          }
        }
        named_barrier_sync(1, kEpiThreads);                // s_last is rewritten by the next work item
        if (++acc == 2) { acc = 0; acc_phase ^= 1; }
        continue;
      }
#pragma unroll 1
      for (int ch = 0; ch < BN / 64; ++ch) {
        uint32_t v[2][32];
        tmem_ld_32x32b_x32(t_acc + ch * 64, v[0]);
        tmem_ld_32x32b_x32(t_acc + ch * 64 + 32, v[1]);
        tmem_ld_wait();
        if (ch == BN / 64 - 1) {                        // every TMEM read of this accumulator has retired
          tc_fence_before();
          __syncwarp();
          if (lane == 0) {
            if (CG == 2) mbar_arrive_leader(&tmem_empty[acc]);
            else mbar_arrive(&tmem_empty[acc]);
          }
        }
        const int col0 = c.n_blk * BN + ch * 64;
        if (EPI == EPI_F32) {
          // fp32 row-major out (+)= alpha * acc : thread = row, 2 x 128 contiguous bytes per chunk
          if (row_ok) {
            float* orow = p.out_f32 + static_cast<long>(row) * p.out_f32_stride + col0;
#pragma unroll
            for (int h = 0; h < 2; ++h)
#pragma unroll
              for (int j = 0; j < 32; j += 4) {
                const int col = col0 + h * 32 + j;
                if (col < p.N) {                        // N % 4 == 0 (checked on the host)
                  float4 o = make_float4(__uint_as_float(v[h][j]) * p.alpha, __uint_as_float(v[h][j + 1]) * p.alpha,
                                         __uint_as_float(v[h][j + 2]) * p.alpha, __uint_as_float(v[h][j + 3]) * p.alpha);
                  float4* dst = reinterpret_cast<float4*>(orow + h * 32 + j);
                  if (p.accumulate) {
                    const float4 old = *dst;
                    o.x += old.x; o.y += old.y; o.z += old.z; o.w += old.w;
                  }
                  *dst = o;
                }
              }
          }
        } else if (EPI == EPI_SWIGLU) {
          // every 64-column chunk of the accumulator is [32 gate | 32 up] of the same 32 features (weight rows interleaved at
          // weight-refresh time, parallel/weight_sync.py): silu(g) * u is formed in registers, two chunks make one 64-wide
          // output slab -- the [tokens, 2F] gate_up tensor never exists
          constexpr float kLog2e = 1.4426950408889634f;
          const float rs8 = (FP8 && row_ok) ? p.row_scale[row] : 1.f;
#pragma unroll
          for (int j = 0; j < 32; j += 2) {
            float g0 = __uint_as_float(v[0][j]), g1 = __uint_as_float(v[0][j + 1]);
            float u0 = __uint_as_float(v[1][j]), u1 = __uint_as_float(v[1][j + 1]);
            if (FP8) {
              const int cg_ = min(col0 + j, p.N - 2), cu_ = min(col0 + 32 + j, p.N - 2);
              g0 *= rs8 * p.col_scale[cg_]; g1 *= rs8 * p.col_scale[cg_ + 1];
              u0 *= rs8 * p.col_scale[cu_]; u1 *= rs8 * p.col_scale[cu_ + 1];
            }
            const uint32_t pk = pack_bf16x2(g0 / (1.f + exp2f(-g0 * kLog2e)) * u0, g1 / (1.f + exp2f(-g1 * kLog2e)) * u1);
            if ((ch & 1) == 0) sw_packed[j / 2] = pk;     // static indices: the array stays in registers
            else sw_packed[16 + j / 2] = pk;
          }
          if (ch & 1) {
            uint8_t* buf = staging + store_buf * kStagingBytes;
            if (epi_tid == 0) tma_store_wait_read<1>();
            named_barrier_sync(1, kEpiThreads);
            uint8_t* rowp = buf + row_in_tile * 128;
#pragma unroll
            for (int q8 = 0; q8 < 8; ++q8)
              *reinterpret_cast<uint4*>(rowp + ((q8 ^ (row_in_tile & 7)) * 16)) =
                  make_uint4(sw_packed[q8 * 4], sw_packed[q8 * 4 + 1], sw_packed[q8 * 4 + 2], sw_packed[q8 * 4 + 3]);
            fence_proxy_async_smem();
            named_barrier_sync(2, kEpiThreads);
            if (epi_tid == 0) {
              tma_store_2d(&tmD, buf, c.n_blk * (BN / 2) + (ch >> 1) * 64, m_blk * BLOCK_M);
              tma_store_commit();
            }
            store_buf ^= 1;
          }
        } else {
          uint8_t* buf = staging + store_buf * kStagingBytes;
          if (epi_tid == 0) tma_store_wait_read<1>();   // the store that last used `buf` has drained
          named_barrier_sync(1, kEpiThreads);
          uint32_t packed[32];
#pragma unroll
          for (int h = 0; h < 2; ++h)
#pragma unroll
            for (int j = 0; j < 32; j += 2) {
              float x0 = __uint_as_float(v[h][j]) * p.alpha, x1 = __uint_as_float(v[h][j + 1]) * p.alpha;
              const int col = col0 + h * 32 + j;
              if (FP8) {
                const float rs = row_ok ? p.row_scale[row] : 0.f;
                x0 *= rs * ((col < p.N) ? p.col_scale[col] : 0.f);
                x1 *= rs * ((col + 1 < p.N) ? p.col_scale[col + 1] : 0.f);
              }
              if (p.bias != nullptr) {
                if (col < p.N) x0 += __bfloat162float(p.bias[col]);
                if (col + 1 < p.N) x1 += __bfloat162float(p.bias[col + 1]);
              }
              if (p.act == 1) {
                x0 = gelu_erf_tc(x0);
                x1 = gelu_erf_tc(x1);
              }
              packed[h * 16 + j / 2] = pack_bf16x2(x0, x1);
            }
          uint8_t* rowp = buf + row_in_tile * 128;
#pragma unroll
          for (int q8 = 0; q8 < 8; ++q8)
            *reinterpret_cast<uint4*>(rowp + ((q8 ^ (row_in_tile & 7)) * 16)) =
                make_uint4(packed[q8 * 4], packed[q8 * 4 + 1], packed[q8 * 4 + 2], packed[q8 * 4 + 3]);
          fence_proxy_async_smem();
          named_barrier_sync(2, kEpiThreads);
          if (epi_tid == 0) {
            if (BATCH) tma_store_3d(&tmD, buf, col0, m_blk * BLOCK_M, bidx);
            else tma_store_2d(&tmD, buf, col0, m_blk * BLOCK_M);
            tma_store_commit();
          }
          store_buf ^= 1;
        }
      }
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
    }
    if (EPI != EPI_F32 && !SPLIT && epi_tid == 0) tma_store_wait<0>();
  }

  tc_fence_before();
  if (CG == 2) cluster_sync_all();     // nobody leaves (or frees TMEM) while the peer can still signal / read it
  else __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc_cg<CG>(tmem_base, C::kTmemCols);
  }
}

template <int BN, bool A_MN, bool B_MN, int EPI>
__global__ void __launch_bounds__(kThreads, 1)
gemm_tc_cg1_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                   const __grid_constant__ CUtensorMap tmA2, const __grid_constant__ CUtensorMap tmB2,
                   const __grid_constant__ CUtensorMap tmD, const TcParams p) {
  gemm_tc_body<1, BN, A_MN, B_MN, EPI>(tmA, tmB, tmA2, tmB2, tmD, p);
}
template <int BN, bool A_MN, bool B_MN, int EPI>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kThreads, 1)
gemm_tc_cg2_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                   const __grid_constant__ CUtensorMap tmA2, const __grid_constant__ CUtensorMap tmB2,
                   const __grid_constant__ CUtensorMap tmD, const TcParams p) {
  gemm_tc_body<2, BN, A_MN, B_MN, EPI>(tmA, tmB, tmA2, tmB2, tmD, p);
}

template <int BN>
__global__ void __launch_bounds__(kThreads, 1)
gemm_tc_batched_cg1_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                           const __grid_constant__ CUtensorMap tmD, const TcParams p) {
  gemm_tc_body<1, BN, false, false, EPI_BF16, false, true>(tmA, tmB, tmA, tmB, tmD, p);
}
template <int BN>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kThreads, 1)
gemm_tc_batched_cg2_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                           const __grid_constant__ CUtensorMap tmD, const TcParams p) {
  gemm_tc_body<2, BN, false, false, EPI_BF16, false, true>(tmA, tmB, tmA, tmB, tmD, p);
}
template <int BN, int EPI>
__global__ void __launch_bounds__(kThreads, 1)
gemm_tc_fp8_cg1_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                       const __grid_constant__ CUtensorMap tmD, const TcParams p) {
  gemm_tc_body<1, BN, false, false, EPI, false, false, true>(tmA, tmB, tmA, tmB, tmD, p);
}
template <int BN, int EPI>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kThreads, 1)
gemm_tc_fp8_cg2_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                       const __grid_constant__ CUtensorMap tmD, const TcParams p) {
  gemm_tc_body<2, BN, false, false, EPI, false, false, true>(tmA, tmB, tmA, tmB, tmD, p);
}
template <int CG, int BN, int EPI>
static cudaError_t launch_fp8(const CUtensorMap* maps, const TcParams& p, int num_sms, cudaStream_t stream) {
  using C = Cfg<CG, BN, false>;
  const int num_m = (p.M + CG * BLOCK_M - 1) / (CG * BLOCK_M), num_n = (p.N + BN - 1) / BN;
  int units = num_m * num_n;
  if (units > num_sms / CG) units = num_sms / CG;
  if (units < 1) units = 1;
  static bool configured = false;
  if (CG == 1) {
    auto kern = gemm_tc_fp8_cg1_kernel<BN, EPI>;
    if (!configured) {
      cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, C::kTotal);
      if (e != cudaSuccess) return e;
      configured = true;
    }
    kern<<<units, kThreads, C::kTotal, stream>>>(maps[0], maps[1], maps[2], p);
  } else {
    auto kern = gemm_tc_fp8_cg2_kernel<BN, EPI>;
    if (!configured) {
      cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, C::kTotal);
      if (e != cudaSuccess) return e;
      configured = true;
    }
    kern<<<2 * units, kThreads, C::kTotal, stream>>>(maps[0], maps[1], maps[2], p);
  }
  return cudaGetLastError();
}

template <int CG, int BN>
static cudaError_t launch_batched(const CUtensorMap* maps, const TcParams& p, int num_sms, cudaStream_t stream) {
  using C = Cfg<CG, BN, false>;
  const int num_m = (p.M + CG * BLOCK_M - 1) / (CG * BLOCK_M), num_n = (p.N + BN - 1) / BN;
  long units = static_cast<long>(num_m) * num_n * p.batch;
  if (units > num_sms / CG) units = num_sms / CG;
  if (units < 1) units = 1;
  static bool configured = false;
  if (CG == 1) {
    auto kern = gemm_tc_batched_cg1_kernel<BN>;
    if (!configured) {
      cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, C::kTotal);
      if (e != cudaSuccess) return e;
      configured = true;
    }
    kern<<<static_cast<int>(units), kThreads, C::kTotal, stream>>>(maps[0], maps[1], maps[2], p);
  } else {
    auto kern = gemm_tc_batched_cg2_kernel<BN>;
    if (!configured) {
      cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, C::kTotal);
      if (e != cudaSuccess) return e;
      configured = true;
    }
    kern<<<static_cast<int>(2 * units), kThreads, C::kTotal, stream>>>(maps[0], maps[1], maps[2], p);
  }
  return cudaGetLastError();
}

template <int BN, bool A_MN, bool B_MN>
__global__ void __launch_bounds__(kThreads, 1)
gemm_tc_splitk_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const TcParams p) {
  gemm_tc_body<1, BN, A_MN, B_MN, EPI_BF16, true>(tmA, tmB, tmA, tmB, tmA, p);
}

template <int BN, bool A_MN, bool B_MN>
static cudaError_t launch_splitk(const CUtensorMap* maps, const TcParams& p, int num_sms, cudaStream_t stream) {
  using C = Cfg<1, BN, B_MN>;
  auto kern = gemm_tc_splitk_kernel<BN, A_MN, B_MN>;
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, C::kTotal);
    if (e != cudaSuccess) return e;
    configured = true;
  }
  const int tiles = ((p.M + BLOCK_M - 1) / BLOCK_M) * ((p.N + BN - 1) / BN);
  int grid = tiles * p.splits;
  if (grid > num_sms) grid = num_sms;
  kern<<<grid, kThreads, C::kTotal, stream>>>(maps[0], maps[1], p);
  return cudaGetLastError();
}

template <int CG, int BN, bool A_MN, bool B_MN, int EPI>
static cudaError_t launch(const CUtensorMap* maps, const TcParams& p, int num_sms, cudaStream_t stream) {
  using C = Cfg<CG, BN, B_MN>;
  const int num_m = (p.M + CG * BLOCK_M - 1) / (CG * BLOCK_M), num_n = (p.N + BN - 1) / BN;
  int units = num_m * num_n;
  if (units > num_sms / CG) units = num_sms / CG;
  if (units < 1) units = 1;
  static bool configured = false;
  if (CG == 1) {
    auto kern = gemm_tc_cg1_kernel<BN, A_MN, B_MN, EPI>;
    if (!configured) {
      cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, C::kTotal);
      if (e != cudaSuccess) return e;
      configured = true;
    }
    kern<<<units, kThreads, C::kTotal, stream>>>(maps[0], maps[1], maps[2], maps[3], maps[4], p);
  } else {
    auto kern = gemm_tc_cg2_kernel<BN, A_MN, B_MN, EPI>;
    if (!configured) {
      cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, C::kTotal);
      if (e != cudaSuccess) return e;
      configured = true;
    }
    kern<<<2 * units, kThreads, C::kTotal, stream>>>(maps[0], maps[1], maps[2], maps[3], maps[4], p);   // consecutive CTAs pair up
  }
  return cudaGetLastError();
}

}  // namespace tc
}  // namespace nrl

// e4m3 x e4m3 GEMM with per-row x per-column scales; maps = {A (uint8 [M,K], box {128, 128}), B (uint8 [N,K], box {bn/cg, 128}), D}
extern "C" cudaError_t nrl_gemm_tc_fp8(const CUtensorMap* maps, const nrl::tc::TcParams* p, int cg, int bn, int swiglu, int num_sms,
                                       cudaStream_t stream) {
  using namespace nrl::tc;
  if (p->row_scale == nullptr || p->col_scale == nullptr || p->K2 != 0) return cudaErrorInvalidValue;
#define NRL_TC_FP8(CG_, BN_)                                                                                     \
  if (cg == CG_ && bn == BN_)                                                                                    \
    return swiglu ? launch_fp8<CG_, BN_, EPI_SWIGLU>(maps, *p, num_sms, stream) : launch_fp8<CG_, BN_, EPI_BF16>(maps, *p, num_sms, stream);
  NRL_TC_FP8(1, 128)
  NRL_TC_FP8(1, 256)
  NRL_TC_FP8(2, 128)
  NRL_TC_FP8(2, 256)
#undef NRL_TC_FP8
  if (!swiglu && cg == 2 && bn == 192) return launch_fp8<2, 192, EPI_BF16>(maps, *p, num_sms, stream);
  if (!swiglu && cg == 1 && bn == 192) return launch_fp8<1, 192, EPI_BF16>(maps, *p, num_sms, stream);
  return cudaErrorInvalidValue;
}

extern "C" cudaError_t nrl_gemm_tc_batched(const CUtensorMap* maps, const nrl::tc::TcParams* p, int cg, int bn, int num_sms,
                                           cudaStream_t stream) {
  using namespace nrl::tc;
  if (p->batch < 1 || p->K2 != 0) return cudaErrorInvalidValue;
  if (cg == 1 && bn == 128) return launch_batched<1, 128>(maps, *p, num_sms, stream);
  if (cg == 1 && bn == 256) return launch_batched<1, 256>(maps, *p, num_sms, stream);
  if (cg == 2 && bn == 128) return launch_batched<2, 128>(maps, *p, num_sms, stream);
  if (cg == 2 && bn == 256) return launch_batched<2, 256>(maps, *p, num_sms, stream);
  return cudaErrorInvalidValue;
}

// split-K variant (cta_group::1, bf16 output with plain stores, no second operand pair); maps = {A, B}
extern "C" cudaError_t nrl_gemm_tc_splitk(const CUtensorMap* maps, const nrl::tc::TcParams* p, int bn, int a_mn, int b_mn, int num_sms,
                                          cudaStream_t stream) {
  using namespace nrl::tc;
  if (p->splits < 1 || p->ws == nullptr || p->counters == nullptr || p->out_bf16 == nullptr || p->K2 != 0) return cudaErrorInvalidValue;
#define NRL_TC_SPLIT(BN_, AMN_, BMN_) \
  if (bn == BN_ && a_mn == (AMN_ ? 1 : 0) && b_mn == (BMN_ ? 1 : 0)) return launch_splitk<BN_, AMN_, BMN_>(maps, *p, num_sms, stream);
  NRL_TC_SPLIT(64, false, false)
  NRL_TC_SPLIT(128, false, false)
  NRL_TC_SPLIT(64, false, true)
  NRL_TC_SPLIT(128, false, true)
  NRL_TC_SPLIT(64, true, true)
  NRL_TC_SPLIT(128, true, true)
#undef NRL_TC_SPLIT
  return cudaErrorInvalidValue;
}

// maps = {A, B, A2, B2, D}; operand boxes (tma_host.h): K-major operand [rows, K]: box {rows = 128 (A) | BN/CG (B), 64};
// MN-major operand [K, cols]: box {64, 64}; D: box {128, 64}.  Unused maps may repeat a valid one.
extern "C" cudaError_t nrl_gemm_tc(const CUtensorMap* maps, const nrl::tc::TcParams* p, int cg, int bn, int a_mn, int b_mn,
                                   int epi, int num_sms, cudaStream_t stream) {
  using namespace nrl::tc;
#define NRL_TC_CASE(CG_, BN_, AMN_, BMN_, EPI_)                                                             \
  if (cg == CG_ && bn == BN_ && a_mn == (AMN_ ? 1 : 0) && b_mn == (BMN_ ? 1 : 0) && epi == EPI_)             \
    return launch<CG_, BN_, AMN_, BMN_, EPI_>(maps, *p, num_sms, stream);
  // K-major x K-major ("TN": forward linears, LoRA down-projection, DeBERTa)
  NRL_TC_CASE(1, 64, false, false, EPI_BF16)
  NRL_TC_CASE(1, 128, false, false, EPI_BF16)
  NRL_TC_CASE(1, 192, false, false, EPI_BF16)
  NRL_TC_CASE(1, 256, false, false, EPI_BF16)
  NRL_TC_CASE(2, 128, false, false, EPI_BF16)
  NRL_TC_CASE(2, 192, false, false, EPI_BF16)
  NRL_TC_CASE(2, 256, false, false, EPI_BF16)
  NRL_TC_CASE(1, 128, false, false, EPI_SWIGLU)
  NRL_TC_CASE(1, 256, false, false, EPI_SWIGLU)
  NRL_TC_CASE(2, 128, false, false, EPI_SWIGLU)
  NRL_TC_CASE(2, 256, false, false, EPI_SWIGLU)
  // K-major x MN-major ("NN": dgrad  dX = dY W, dH = dZ W_lm)
  NRL_TC_CASE(1, 64, false, true, EPI_BF16)
  NRL_TC_CASE(1, 128, false, true, EPI_BF16)
  NRL_TC_CASE(1, 192, false, true, EPI_BF16)
  NRL_TC_CASE(1, 256, false, true, EPI_BF16)
  NRL_TC_CASE(2, 128, false, true, EPI_BF16)
  NRL_TC_CASE(2, 192, false, true, EPI_BF16)
  NRL_TC_CASE(2, 256, false, true, EPI_BF16)
  // MN-major x MN-major ("NT": wgrad  dW = dY^T X)
  NRL_TC_CASE(1, 64, true, true, EPI_BF16)
  NRL_TC_CASE(1, 128, true, true, EPI_BF16)
  NRL_TC_CASE(1, 256, true, true, EPI_BF16)
  NRL_TC_CASE(2, 256, true, true, EPI_BF16)
  NRL_TC_CASE(1, 64, true, true, EPI_F32)
  NRL_TC_CASE(1, 128, true, true, EPI_F32)
  NRL_TC_CASE(1, 256, true, true, EPI_F32)
  NRL_TC_CASE(2, 256, true, true, EPI_F32)
#undef NRL_TC_CASE
  return cudaErrorInvalidValue;
}
