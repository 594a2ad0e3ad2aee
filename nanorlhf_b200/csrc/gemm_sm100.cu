// tcgen05 / TMEM / TMA GEMM family for sm_100a (hand-written PTX, no CUTLASS).
//
//   D[M,N] = A[M,K] * B[N,K]^T          A, B bf16 row-major with K contiguous ("TN", nn.Linear layout)
//
// One persistent CTA per SM, 6 warps:
//   warp 0      TMA producer   : cp.async.bulk.tensor 128B-swizzled A/B tiles -> smem ring (mbarrier tx)
//   warp 1      MMA issuer     : one thread issues tcgen05.mma (128 x BLOCK_N x 16), accumulators in TMEM,
//                                tcgen05.commit releases smem stages / publishes accumulators
//   warps 2..5  epilogue       : tcgen05.ld TMEM -> registers -> fused epilogue; accumulators are
//                                double-buffered in TMEM so the epilogue of tile i overlaps the mainloop of i+1
//
// Epilogues (template EPI):
//   EPI_STORE   bf16 store (+ optional bias[N]) through swizzled smem + TMA store           (K3: linear layers)
//   EPI_LOGPROB fused lm-head log-prob (K-LP fwd): per-row online (max, sum-exp, sum exp*z) over the vocab
//               range of the work item + target logit; nothing of size [M,V] is ever written
//   EPI_SWIGLU  gate_up projection with silu(gate)*up applied in the epilogue (interleaved weight rows)
//   EPI_DLOGITS K-LP backward: dZ = (onehot(target) - softmax(z)) * g / T in bf16 via TMA store
//
// Reference call sites this replaces: cuBLAS GEMMs behind every nn.Linear and the
// logits -> /T -> log_softmax -> gather chain (/root/reference/GRPO/grpo_trainer.py:543-549,652-660).
#include "common.cuh"
#include "gemm_sm100.h"

namespace nrl {

constexpr int BLOCK_M = 128;
constexpr int BLOCK_K = 64;        // 64 bf16 = 128 bytes = one swizzle-128B row
constexpr int UMMA_K = 16;
constexpr int kNumThreads = 192;
constexpr int kNumEpiThreads = 128;
constexpr int kStagingBytes = BLOCK_M * 128;   // 128 rows x 64 bf16
constexpr int GROUP_M = 16;

// GELU(x) = x/2 * (1 + erf(x / sqrt 2)) with erf from Abramowitz-Stegun 7.1.26 (|error| < 1.5e-7): two MUFU ops and
// seven FMAs, cheap enough to hide under the next tile's MMAs (libdevice erff made the epilogue the bottleneck).
NRL_DEVICE float gelu_erf(float x) {
  const float z = fabsf(x) * 0.70710678118654752f;
  const float t = __fdividef(1.f, fmaf(0.3275911f, z, 1.f));
  float poly = fmaf(1.061405429f, t, -1.453152027f);
  poly = fmaf(poly, t, 1.421413741f);
  poly = fmaf(poly, t, -0.284496736f);
  poly = fmaf(poly, t, 0.254829592f);
  const float erf_abs = 1.f - poly * t * exp2f(-z * z * 1.4426950408889634f);
  return 0.5f * x * (1.f + copysignf(erf_abs, x));
}

template <int BLOCK_N>
struct SmemLayout {
  static constexpr int kABytes = BLOCK_M * BLOCK_K * 2;
  static constexpr int kBBytes = BLOCK_N * BLOCK_K * 2;
  static constexpr int kStageBytes = kABytes + kBBytes;
  static constexpr int kStages = (BLOCK_N >= 192) ? 4 : 6;
  static constexpr int kStagingOff = kStages * kStageBytes;
  static constexpr int kBarOff = kStagingOff + 2 * kStagingBytes;
  static constexpr int kTotal = kBarOff + 256;
};

struct TileCoord {
  int m_blk, n_blk;
};

// grouped ordering: GROUP_M m-tiles x all n-tiles form a super-row so the A panel stays in L2
NRL_DEVICE TileCoord tile_coord(int t, int num_m, int num_n) {
  int per_group = GROUP_M * num_n;
  int g = t / per_group;
  int r = t - g * per_group;
  int gsz = min(GROUP_M, num_m - g * GROUP_M);
  TileCoord c;
  c.n_blk = r / gsz;
  c.m_blk = g * GROUP_M + (r - c.n_blk * gsz);
  return c;
}

// FP8 = true: A and B are e4m3 bytes (one k-block = 128 elements = the same 128-byte swizzle row), the MMA is
// tcgen05 kind::f8f6f4 (K = 32 per instruction, 2x the bf16 rate) and the epilogue applies the per-row (token)
// and per-column (output channel) dequantisation scales.
template <int BLOCK_N, int EPI, bool FP8 = false>
__global__ void __launch_bounds__(kNumThreads, 1)
gemm_bf16_tn_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                    const __grid_constant__ CUtensorMap tmD, GemmParams p) {
  using L = SmemLayout<BLOCK_N>;
  extern __shared__ __align__(1024) uint8_t smem[];
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + L::kBarOff);
  uint64_t* empty_bar = full_bar + L::kStages;
  uint64_t* tmem_full = empty_bar + L::kStages;
  uint64_t* tmem_empty = tmem_full + 2;
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(tmem_empty + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int num_m = (p.M + BLOCK_M - 1) / BLOCK_M;
  const int num_n_total = (p.N + BLOCK_N - 1) / BLOCK_N;
  constexpr int kElemsPerKBlock = FP8 ? 128 : BLOCK_K;          // 128 bytes of K per stage either way
  const int num_kb = (p.K + kElemsPerKBlock - 1) / kElemsPerKBlock;
  // EPI_LOGPROB: a work item is (m tile, vocab split) and walks n tiles [n_begin, n_end) itself
  const int n_splits = (EPI == EPI_LOGPROB) ? p.n_splits : 1;
  const int n_per_split = (EPI == EPI_LOGPROB) ? (num_n_total + n_splits - 1) / n_splits : 1;
  const int num_work = (EPI == EPI_LOGPROB) ? num_m * n_splits : num_m * num_n_total;
  constexpr uint32_t kTmemCols = (2 * BLOCK_N <= 256) ? 256 : 512;      // power of two >= two accumulators

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    if (EPI != EPI_LOGPROB) tma_prefetch_desc(&tmD);
  }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < L::kStages; ++i) {
      mbar_init(&full_bar[i], 1);
      mbar_init(&empty_bar[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tmem_full[i], 1);
      mbar_init(&tmem_empty[i], kNumEpiThreads / 32);
    }
    fence_mbar_init();
  }
  if (warp == 1) {
    __syncwarp();
    tmem_alloc(tmem_ptr, kTmemCols);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;

  if (warp == 0) {
    // ===================================== TMA producer ======================================
    if (elect_one()) {
      int stage = 0;
      uint32_t phase = 0;
      for (int w = blockIdx.x; w < num_work; w += gridDim.x) {
        int m_blk, n_begin, n_end;
        if (EPI == EPI_LOGPROB) {
          m_blk = w % num_m;
          int sp = w / num_m;
          n_begin = sp * n_per_split;
          n_end = min(num_n_total, n_begin + n_per_split);
        } else {
          TileCoord c = tile_coord(w, num_m, num_n_total);
          m_blk = c.m_blk;
          n_begin = c.n_blk;
          n_end = n_begin + 1;
        }
        for (int n_blk = n_begin; n_blk < n_end; ++n_blk) {
          for (int kb = 0; kb < num_kb; ++kb) {
            mbar_wait(&empty_bar[stage], phase ^ 1);
            uint8_t* sa = smem + stage * L::kStageBytes;
            uint8_t* sb = sa + L::kABytes;
            mbar_arrive_expect_tx(&full_bar[stage], L::kStageBytes);
            tma_load_2d(sa, &tmA, &full_bar[stage], kb * kElemsPerKBlock, m_blk * BLOCK_M);
            tma_load_2d(sb, &tmB, &full_bar[stage], kb * kElemsPerKBlock, n_blk * BLOCK_N);
            if (++stage == L::kStages) { stage = 0; phase ^= 1; }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===================================== MMA issuer =======================================
    // elect.sync instead of `lane == 0`: the compiler can then prove single-thread execution and emits the
    // UTCHMMAs back to back instead of wrapping each one in an ELECT/branch loop (~13 instructions per MMA)
    if (elect_one()) {
      constexpr uint32_t idesc = FP8 ? make_idesc(BLOCK_M, BLOCK_N, 0, 0)    // e4m3 x e4m3 -> fp32
                                     : make_idesc(BLOCK_M, BLOCK_N, 1, 1);   // bf16 x bf16 -> fp32
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      for (int w = blockIdx.x; w < num_work; w += gridDim.x) {
        int n_begin = 0, n_end = 1;
        if (EPI == EPI_LOGPROB) {
          int sp = w / num_m;
          n_begin = sp * n_per_split;
          n_end = min(num_n_total, n_begin + n_per_split);
        }
        for (int n_blk = n_begin; n_blk < n_end; ++n_blk) {
          mbar_wait(&tmem_empty[acc], acc_phase ^ 1);
          tc_fence_after();
          const uint32_t d_tmem = tmem_base + acc * BLOCK_N;
          for (int kb = 0; kb < num_kb; ++kb) {
            mbar_wait(&full_bar[stage], phase);
            tc_fence_after();
            const uint32_t a_addr = smem_u32(smem + stage * L::kStageBytes);
            const uint32_t b_addr = a_addr + L::kABytes;
#pragma unroll
            for (int k = 0; k < BLOCK_K / UMMA_K; ++k) {
              uint64_t adesc = make_smem_desc_sw128(a_addr + k * UMMA_K * 2);
              uint64_t bdesc = make_smem_desc_sw128(b_addr + k * UMMA_K * 2);
              if (FP8) umma_f8(d_tmem, adesc, bdesc, idesc, (kb | k) != 0 ? 1u : 0u);
              else umma_f16(d_tmem, adesc, bdesc, idesc, (kb | k) != 0 ? 1u : 0u);
            }
            umma_commit(&empty_bar[stage]);          // smem stage reusable once these MMAs retire
            if (++stage == L::kStages) { stage = 0; phase ^= 1; }
          }
          umma_commit(&tmem_full[acc]);              // accumulator complete -> epilogue
          if (++acc == 2) { acc = 0; acc_phase ^= 1; }
        }
      }
    }
    __syncwarp();
  } else {
    // ===================================== epilogue ==========================================
    const int quad = warp & 3;                        // TMEM lane quadrant this warp may access
    const int row_in_tile = quad * 32 + lane;
    const int epi_tid = threadIdx.x - 64;
    int acc = 0;
    uint32_t acc_phase = 0;
    uint8_t* staging = smem + L::kStagingOff;
    int store_buf = 0;
    constexpr float kLog2e = 1.4426950408889634f;

    for (int w = blockIdx.x; w < num_work; w += gridDim.x) {
      int m_blk, n_begin, n_end, split = 0;
      if (EPI == EPI_LOGPROB) {
        m_blk = w % num_m;
        split = w / num_m;
        n_begin = split * n_per_split;
        n_end = min(num_n_total, n_begin + n_per_split);
      } else {
        TileCoord c = tile_coord(w, num_m, num_n_total);
        m_blk = c.m_blk;
        n_begin = c.n_blk;
        n_end = n_begin + 1;
      }
      const int row = m_blk * BLOCK_M + row_in_tile;
      const bool row_ok = row < p.M;
      // per-row state of the fused log-prob epilogues
      float run_max = -INFINITY, run_sum = 0.f, run_ez = 0.f, tgt_logit = -INFINITY;
      int target = -1;
      float row_lse = 0.f, row_g = 0.f;
      if (EPI == EPI_LOGPROB || EPI == EPI_DLOGITS) {
        target = row_ok ? p.targets[row] : -1;
        if (EPI == EPI_DLOGITS && row_ok) {
          row_lse = p.lse[row];
          row_g = p.grad_logp[row] * p.scale;          // d logp/dz carries the 1/T factor
        }
      }

      for (int n_blk = n_begin; n_blk < n_end; ++n_blk) {
        mbar_wait(&tmem_full[acc], acc_phase);
        tc_fence_after();
        const uint32_t t_acc = tmem_base + acc * BLOCK_N + (static_cast<uint32_t>(quad * 32) << 16);
        uint32_t swiglu_packed[32];
        (void)swiglu_packed;
#pragma unroll 1
        for (int c = 0; c < BLOCK_N / 64; ++c) {
          uint32_t v[2][32];
          tmem_ld_32x32b_x32(t_acc + c * 64, v[0]);
          tmem_ld_32x32b_x32(t_acc + c * 64 + 32, v[1]);
          tmem_ld_wait();
          if (c == BLOCK_N / 64 - 1) {
            // every TMEM read of this accumulator has retired: hand it back to the MMA warp
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&tmem_empty[acc]);
          }
          const int col0 = n_blk * BLOCK_N + c * 64;

          if (EPI == EPI_LOGPROB) {
            // online (max, sum exp, sum exp*z) in base-2 domain; z = acc * (1/T)
            float cmax = -INFINITY;
#pragma unroll
            for (int h = 0; h < 2; ++h)
#pragma unroll
              for (int j = 0; j < 32; ++j) {
                float z = __uint_as_float(v[h][j]) * p.scale;
                int col = col0 + h * 32 + j;
                z = (col < p.N) ? z : -INFINITY;
                v[h][j] = __float_as_uint(z);
                cmax = fmaxf(cmax, z);
                if (col == target) tgt_logit = z;
              }
            float new_max = fmaxf(run_max, cmax);
            if (new_max > -INFINITY) {
              float corr = exp2f((run_max - new_max) * kLog2e);   // exp2(-inf) = 0 on the first chunk
              run_sum *= corr;
              run_ez *= corr;
              const float mb = new_max * kLog2e;
#pragma unroll
              for (int h = 0; h < 2; ++h)
#pragma unroll
                for (int j = 0; j < 32; ++j) {
                  float z = __uint_as_float(v[h][j]);
                  float e = exp2f(fmaf(z, kLog2e, -mb));
                  run_sum += e;
                  run_ez = fmaf(e, (z == -INFINITY) ? 0.f : z, run_ez);
                }
              run_max = new_max;
            }
          } else if (EPI == EPI_SWIGLU) {
            // ---- fused SwiGLU: the weight rows are interleaved so that every 64-column chunk of the tile is
            //      [32 gate | 32 up] of the same 32 features; two chunks make one 64-wide output slab ----
            uint32_t cur[16];
            const float rs = (FP8 && row_ok) ? p.row_scale[row] : 1.f;
#pragma unroll
            for (int j = 0; j < 32; j += 2) {
              float g0 = __uint_as_float(v[0][j]), g1 = __uint_as_float(v[0][j + 1]);
              float u0 = __uint_as_float(v[1][j]), u1 = __uint_as_float(v[1][j + 1]);
              if (FP8) {
                const int cg = min(col0 + j, p.N - 2), cu = min(col0 + 32 + j, p.N - 2);
                g0 *= rs * p.col_scale[cg]; g1 *= rs * p.col_scale[cg + 1];
                u0 *= rs * p.col_scale[cu]; u1 *= rs * p.col_scale[cu + 1];
              }
              cur[j / 2] = pack_bf16x2(g0 / (1.f + exp2f(-g0 * kLog2e)) * u0, g1 / (1.f + exp2f(-g1 * kLog2e)) * u1);
            }
            if ((c & 1) == 0) {
#pragma unroll
              for (int j = 0; j < 16; ++j) swiglu_packed[j] = cur[j];
            } else {
#pragma unroll
              for (int j = 0; j < 16; ++j) swiglu_packed[16 + j] = cur[j];
              uint8_t* buf = staging + store_buf * kStagingBytes;
              if (epi_tid == 0) tma_store_wait_read<1>();
              named_barrier_sync(1, kNumEpiThreads);
              uint8_t* rowp = buf + row_in_tile * 128;
#pragma unroll
              for (int ch = 0; ch < 8; ++ch) {
                uint4 q = make_uint4(swiglu_packed[ch * 4], swiglu_packed[ch * 4 + 1], swiglu_packed[ch * 4 + 2], swiglu_packed[ch * 4 + 3]);
                *reinterpret_cast<uint4*>(rowp + ((ch ^ (row_in_tile & 7)) * 16)) = q;
              }
              fence_proxy_async_smem();
              named_barrier_sync(2, kNumEpiThreads);
              if (epi_tid == 0) {
                tma_store_2d(&tmD, buf, n_blk * (BLOCK_N / 2) + (c >> 1) * 64, m_blk * BLOCK_M);
                tma_store_commit();
              }
              store_buf ^= 1;
            }
          } else if (EPI == EPI_MERGE && p.mc_out != nullptr) {
            // ---- K-BC: W' = W + s B A straight into EVERY rank's sampler arena with multimem.st (the NVSwitch replicates
            //      each store; no per-peer recompute, no NCCL broadcast).  The tile is staged through shared memory so that
            //      a warp writes four full 128-byte row segments per instruction instead of 32 scattered 16-byte pieces ----
            uint8_t* buf = staging + store_buf * kStagingBytes;
            named_barrier_sync(1, kNumEpiThreads);            // the previous chunk's readers are done with `buf`
            {
              uint32_t packed[32];
#pragma unroll
              for (int h = 0; h < 2; ++h)
#pragma unroll
                for (int j = 0; j < 32; j += 2) {
                  float x0 = __uint_as_float(v[h][j]) * p.scale, x1 = __uint_as_float(v[h][j + 1]) * p.scale;
                  const int col = col0 + h * 32 + j;
                  if (row_ok && col + 1 < p.N) {
                    const float2 w2 = unpack_bf16x2(*reinterpret_cast<const uint32_t*>(p.addend + static_cast<long>(row) * p.addend_stride + col));
                    x0 += w2.x;
                    x1 += w2.y;
                  }
                  packed[h * 16 + j / 2] = pack_bf16x2(x0, x1);
                }
              uint8_t* rowp = buf + row_in_tile * 128;
#pragma unroll
              for (int ch = 0; ch < 8; ++ch)
                *reinterpret_cast<uint4*>(rowp + ((ch ^ (row_in_tile & 7)) * 16)) =
                    make_uint4(packed[ch * 4], packed[ch * 4 + 1], packed[ch * 4 + 2], packed[ch * 4 + 3]);
            }
            named_barrier_sync(2, kNumEpiThreads);
#pragma unroll
            for (int it = 0; it < 8; ++it) {
              const int idx = it * kNumEpiThreads + epi_tid;  // 1024 16-byte pieces: piece = (row r, chunk ch)
              const int r = idx >> 3, ch = idx & 7;
              const int grow = m_blk * BLOCK_M + r, gcol = col0 + ch * 8;
              if (grow < p.M && gcol < p.N) {                 // N % 8 == 0 (checked on the host)
                const uint4 q = *reinterpret_cast<const uint4*>(buf + r * 128 + ((ch ^ (r & 7)) * 16));
                asm volatile("multimem.st.relaxed.sys.global.v4.bf16x2 [%0], {%1,%2,%3,%4};"
                             ::"l"(p.mc_out + static_cast<long>(grow) * p.mc_stride + gcol), "r"(q.x), "r"(q.y), "r"(q.z), "r"(q.w) : "memory");
              }
            }
            store_buf ^= 1;
          } else {
            // ---- bf16 tile store through swizzled smem + TMA ----
            uint8_t* buf = staging + store_buf * kStagingBytes;
            if (epi_tid == 0) tma_store_wait_read<1>();   // the store that last used `buf` has drained
            named_barrier_sync(1, kNumEpiThreads);
            uint32_t packed[32];
#pragma unroll
            for (int h = 0; h < 2; ++h)
#pragma unroll
              for (int j = 0; j < 32; j += 2) {
                float x0 = __uint_as_float(v[h][j]);
                float x1 = __uint_as_float(v[h][j + 1]);
                const int col = col0 + h * 32 + j;
                if (EPI == EPI_MERGE) {
                  // K-BC: merged weight = W + (alpha/r) * (B A); W streamed straight from HBM in the epilogue
                  x0 *= p.scale;
                  x1 *= p.scale;
                  if (row_ok && col + 1 < p.N) {
                    float2 w2 = unpack_bf16x2(*reinterpret_cast<const uint32_t*>(p.addend + static_cast<long>(row) * p.addend_stride + col));
                    x0 += w2.x;
                    x1 += w2.y;
                  } else if (row_ok && col < p.N) {
                    x0 += __bfloat162float(p.addend[static_cast<long>(row) * p.addend_stride + col]);
                  }
                } else if (EPI == EPI_STORE) {
                  if (FP8) {
                    const float rs = row_ok ? p.row_scale[row] : 0.f;
                    x0 *= rs * ((col < p.N) ? p.col_scale[col] : 0.f);
                    x1 *= rs * ((col + 1 < p.N) ? p.col_scale[col + 1] : 0.f);
                  }
                  if (p.bias != nullptr) {
                    if (col < p.N) x0 += __bfloat162float(p.bias[col]);
                    if (col + 1 < p.N) x1 += __bfloat162float(p.bias[col + 1]);
                  }
                  if (p.act == 1) {           // exact-form GELU on the fp32 accumulator
                    x0 = gelu_erf(x0);
                    x1 = gelu_erf(x1);
                  }
                } else {  // EPI_DLOGITS
                  float p0 = exp2f((x0 * p.scale - row_lse) * kLog2e);
                  float p1 = exp2f((x1 * p.scale - row_lse) * kLog2e);
                  x0 = ((col == target) ? 1.f : 0.f) - p0;
                  x1 = ((col + 1 == target) ? 1.f : 0.f) - p1;
                  x0 = row_ok ? x0 * row_g : 0.f;
                  x1 = row_ok ? x1 * row_g : 0.f;
                }
                packed[h * 16 + j / 2] = pack_bf16x2(x0, x1);
              }
            uint8_t* rowp = buf + row_in_tile * 128;
#pragma unroll
            for (int ch = 0; ch < 8; ++ch) {
              uint4 q = make_uint4(packed[ch * 4], packed[ch * 4 + 1], packed[ch * 4 + 2], packed[ch * 4 + 3]);
              *reinterpret_cast<uint4*>(rowp + ((ch ^ (row_in_tile & 7)) * 16)) = q;
            }
            fence_proxy_async_smem();
            named_barrier_sync(2, kNumEpiThreads);
            if (epi_tid == 0) {
              tma_store_2d(&tmD, buf, col0, m_blk * BLOCK_M);
              tma_store_commit();
            }
            store_buf ^= 1;
          }
        }
        if (++acc == 2) { acc = 0; acc_phase ^= 1; }
      }

      if (EPI == EPI_LOGPROB && row_ok) {
        // partial statistics of this (row, vocab split); combined by lmhead_combine_kernel
        float4 part = make_float4(run_max, run_sum, run_ez, tgt_logit);
        reinterpret_cast<float4*>(p.partials)[static_cast<size_t>(split) * p.M + row] = part;
      }
    }
    if (EPI != EPI_LOGPROB && epi_tid == 0) tma_store_wait<0>();
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, kTmemCols);
  }
}

// Combine the per-split partials of the fused log-prob GEMM into logp / entropy / lse.
__global__ void lmhead_combine_kernel(const float4* __restrict__ partials, int M, int n_splits,
                                      float* __restrict__ logp, float* __restrict__ entropy,
                                      float* __restrict__ lse_out) {
  int row = blockIdx.x * blockDim.x + threadIdx.x;
  if (row >= M) return;
  float mx = -INFINITY;
  for (int s = 0; s < n_splits; ++s) mx = fmaxf(mx, partials[static_cast<size_t>(s) * M + row].x);
  float sum = 0.f, ez = 0.f, tgt = -INFINITY;
  for (int s = 0; s < n_splits; ++s) {
    float4 q = partials[static_cast<size_t>(s) * M + row];
    float c = (q.x == -INFINITY) ? 0.f : __expf(q.x - mx);
    sum += q.y * c;
    ez += q.z * c;
    tgt = fmaxf(tgt, q.w);
  }
  float lse = mx + __logf(sum);
  logp[row] = tgt - lse;
  if (entropy != nullptr) entropy[row] = lse - ez / sum;
  if (lse_out != nullptr) lse_out[row] = lse;
}

// ------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------
template <int BLOCK_N, int EPI, bool FP8 = false>
static cudaError_t launch_impl(const CUtensorMap& tmA, const CUtensorMap& tmB, const CUtensorMap& tmD,
                               const GemmParams& p, int num_sms, cudaStream_t stream) {
  using L = SmemLayout<BLOCK_N>;
  auto kern = gemm_bf16_tn_kernel<BLOCK_N, EPI, FP8>;
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, L::kTotal);
    if (e != cudaSuccess) return e;
    configured = true;
  }
  const int num_m = (p.M + BLOCK_M - 1) / BLOCK_M;
  const int num_n = (p.N + BLOCK_N - 1) / BLOCK_N;
  int work = (EPI == EPI_LOGPROB) ? num_m * p.n_splits : num_m * num_n;
  int grid = work < num_sms ? work : num_sms;
  if (grid < 1) grid = 1;
  kern<<<grid, kNumThreads, L::kTotal, stream>>>(tmA, tmB, tmD, p);
  return cudaGetLastError();
}

}  // namespace nrl

extern "C" cudaError_t nrl_gemm_bf16_tn(const CUtensorMap* tmA, const CUtensorMap* tmB, const CUtensorMap* tmD,
                                        const nrl::GemmParams* p, int block_n, int epi, int num_sms,
                                        cudaStream_t stream) {
  using namespace nrl;
  if (block_n == 256) {
    if (epi == EPI_STORE) return launch_impl<256, EPI_STORE>(*tmA, *tmB, *tmD, *p, num_sms, stream);
    if (epi == EPI_LOGPROB) return launch_impl<256, EPI_LOGPROB>(*tmA, *tmB, *tmD, *p, num_sms, stream);
    if (epi == EPI_DLOGITS) return launch_impl<256, EPI_DLOGITS>(*tmA, *tmB, *tmD, *p, num_sms, stream);
    if (epi == EPI_MERGE) return launch_impl<256, EPI_MERGE>(*tmA, *tmB, *tmD, *p, num_sms, stream);
    if (epi == EPI_SWIGLU) return launch_impl<256, EPI_SWIGLU>(*tmA, *tmB, *tmD, *p, num_sms, stream);
  } else if (block_n == 192) {
    // 192-wide tiles: N = 1536-class outputs (o_proj / down_proj) fill 128 of 148 SMs at M = 2048 instead of 96
    if (epi == EPI_STORE) return launch_impl<192, EPI_STORE>(*tmA, *tmB, *tmD, *p, num_sms, stream);
  } else if (block_n == 128) {
    if (epi == EPI_SWIGLU) return launch_impl<128, EPI_SWIGLU>(*tmA, *tmB, *tmD, *p, num_sms, stream);
    if (epi == EPI_MERGE) return launch_impl<128, EPI_MERGE>(*tmA, *tmB, *tmD, *p, num_sms, stream);
    if (epi == EPI_STORE) return launch_impl<128, EPI_STORE>(*tmA, *tmB, *tmD, *p, num_sms, stream);
    if (epi == EPI_LOGPROB) return launch_impl<128, EPI_LOGPROB>(*tmA, *tmB, *tmD, *p, num_sms, stream);
    if (epi == EPI_DLOGITS) return launch_impl<128, EPI_DLOGITS>(*tmA, *tmB, *tmD, *p, num_sms, stream);
  }
  return cudaErrorInvalidValue;
}

extern "C" cudaError_t nrl_gemm_fp8_tn(const CUtensorMap* tmA, const CUtensorMap* tmB, const CUtensorMap* tmD,
                                       const nrl::GemmParams* p, int block_n, int epi, int num_sms, cudaStream_t stream) {
  using namespace nrl;
  if (p->row_scale == nullptr || p->col_scale == nullptr) return cudaErrorInvalidValue;
  if (block_n == 256) {
    if (epi == EPI_STORE) return launch_impl<256, EPI_STORE, true>(*tmA, *tmB, *tmD, *p, num_sms, stream);
    if (epi == EPI_SWIGLU) return launch_impl<256, EPI_SWIGLU, true>(*tmA, *tmB, *tmD, *p, num_sms, stream);
  } else if (block_n == 128) {
    if (epi == EPI_STORE) return launch_impl<128, EPI_STORE, true>(*tmA, *tmB, *tmD, *p, num_sms, stream);
    if (epi == EPI_SWIGLU) return launch_impl<128, EPI_SWIGLU, true>(*tmA, *tmB, *tmD, *p, num_sms, stream);
  }
  return cudaErrorInvalidValue;
}

extern "C" cudaError_t nrl_lmhead_combine(const float* partials, int M, int n_splits, float* logp, float* entropy,
                                          float* lse, cudaStream_t stream) {
  int threads = 256;
  int blocks = (M + threads - 1) / threads;
  nrl::lmhead_combine_kernel<<<blocks, threads, 0, stream>>>(reinterpret_cast<const float4*>(partials), M, n_splits,
                                                             logp, entropy, lse);
  return cudaGetLastError();
}
