// K2: paged-KV decode attention (one new query token per sequence) + KV-cache page writer.
//
// Decode attention on the rollout path is HBM-bound (28 KB of KV per context token for Qwen2.5-1.5B;
// ~100 GB per step at 2048 sequences x 1.7k context), so the design goal is "read every KV byte
// exactly once at full HBM rate":
//   * one CTA per (sequence, kv head, split); all G = Hq/Hkv query heads of the kv head are processed
//     against each KV page read (GQA: 6x / 7x fewer bytes than per-q-head kernels);
//   * a page (16 tokens x 128 dims) is one contiguous 4 KB chunk per K and V -> 16-byte cp.async
//     into XOR-swizzled shared memory, 3-stage per-warp pipeline (96 KB / CTA, 2 CTAs / SM);
//   * math on tensor cores via mma.sync m16n8k16 (the G <= 8 query heads padded to M = 16): the
//     kernel is bandwidth-bound, legacy HMMA issue rate is ~2.5x above what HBM can feed;
//   * each warp streams its own pages with private online-softmax state; warps (and optional
//     KV splits) are merged once at the end.
// Reference path replaced: vLLM PagedAttention / FlashInfer decode inside llm.generate
// (/root/reference/GRPO/grpo_trainer.py:142).
#include "common.cuh"
#include "kernels.h"

namespace nrl {

constexpr int kHeadDim = 128;
constexpr int kPage = 16;                         // tokens per KV page
constexpr int kDecWarps = 4;
constexpr int kDecStages = 3;
constexpr int kPageBytes = kPage * kHeadDim * 2;  // 4096 (K or V of one page, one kv head)

NRL_DEVICE void ldmatrix_x4(uint32_t (&r)[4], uint32_t addr) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}
NRL_DEVICE void ldmatrix_x4_trans(uint32_t (&r)[4], uint32_t addr) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}
NRL_DEVICE void mma_bf16_16816(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

// byte offset of 16-byte chunk `c` (0..15) of row `r` inside a [rows][128 x bf16] swizzled tile
NRL_DEVICE uint32_t swz(int r, int c) { return static_cast<uint32_t>(r * 256 + ((c ^ (r & 7)) << 4)); }

// ---- KV cache write ----------------------------------------------------------------------------
// k, v: [T, Hkv, D] (row stride given) ; caches: [num_blocks, Hkv, kPage, D]; slot = block * kPage + offset
__global__ void kv_cache_write_kernel(const __nv_bfloat16* __restrict__ k, const __nv_bfloat16* __restrict__ v,
                                      long k_stride_t, long v_stride_t, __nv_bfloat16* __restrict__ k_cache,
                                      __nv_bfloat16* __restrict__ v_cache, const int* __restrict__ slot_mapping,
                                      const int* __restrict__ src_index, int T, int Hkv) {
  const int vec_per_tok = Hkv * (kHeadDim / 8);
  const long total = static_cast<long>(T) * vec_per_tok;
  for (long i = blockIdx.x * static_cast<long>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<long>(gridDim.x) * blockDim.x) {
    const int pair = i / vec_per_tok;
    const int r = i % vec_per_tok;
    const long t = src_index ? src_index[pair] : pair;
    const int h = r / (kHeadDim / 8);
    const int c = r % (kHeadDim / 8);
    const int slot = slot_mapping[pair];
    if (slot < 0) continue;
    const long dst = ((static_cast<long>(slot / kPage) * Hkv + h) * kPage + (slot % kPage)) * kHeadDim + c * 8;
    *reinterpret_cast<uint4*>(k_cache + dst) = *reinterpret_cast<const uint4*>(k + t * k_stride_t + h * kHeadDim + c * 8);
    *reinterpret_cast<uint4*>(v_cache + dst) = *reinterpret_cast<const uint4*>(v + t * v_stride_t + h * kHeadDim + c * 8);
  }
}

// ---- decode attention ----------------------------------------------------------------------------
struct DecodeParams {
  const __nv_bfloat16* q;        // [S, Hq, D], row stride q_stride_s
  const __nv_bfloat16* k_cache;  // [num_blocks, Hkv, kPage, D]
  const __nv_bfloat16* v_cache;
  const int* block_tables;       // [S, max_blocks]
  const int* context_lens;       // [S]  number of valid KV tokens (including the current one)
  __nv_bfloat16* out;            // [S, Hq, D]
  float* part_o;                 // [S, Hkv, splits, 8, D]   (splits > 1)
  float* part_ml;                // [S, Hkv, splits, 8, 2]
  long q_stride_s;
  int max_blocks, Hq, Hkv, G, splits;
  float scale_log2;              // softmax scale * log2(e)
};

__global__ void __launch_bounds__(kDecWarps * 32, 2) paged_decode_kernel(DecodeParams p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  // layout: [warp][stage][K page | V page] , then Q tile [16][128] bf16 (4 KB)
  uint8_t* q_tile = smem + kDecWarps * kDecStages * 2 * kPageBytes;
  const int seq = blockIdx.x, kvh = blockIdx.y, split = blockIdx.z;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int g = lane >> 2, t4 = lane & 3;
  const int ctx = p.context_lens[seq];
  const int n_pages = (ctx + kPage - 1) / kPage;
  const int pages_per_split = (n_pages + p.splits - 1) / p.splits;
  const int page_lo = split * pages_per_split;
  const int page_hi = min(n_pages, page_lo + pages_per_split);
  const int* bt = p.block_tables + static_cast<long>(seq) * p.max_blocks;

  // ---- Q tile: G real rows, zero padded to 16 (scores are scaled in fp32 after the MMA) ----
  for (int i = threadIdx.x; i < 16 * 16; i += blockDim.x) {
    const int r = i >> 4, c = i & 15;
    uint4 val = make_uint4(0, 0, 0, 0);
    if (r < p.G) {
      val = *reinterpret_cast<const uint4*>(p.q + seq * p.q_stride_s + (kvh * p.G + r) * kHeadDim + c * 8);
    }
    *reinterpret_cast<uint4*>(q_tile + swz(r, c)) = val;
  }
  __syncthreads();
  uint32_t qf[8][4];
  {
    const uint32_t qbase = smem_u32(q_tile);
    // matrices: (rows 0-7, k0..7) (rows 8-15, k0..7) (rows 0-7, k8..15) (rows 8-15, k8..15)
    const int mrow = (lane & 7) + ((lane >> 3) & 1) * 8;
    const int mcol = (lane >> 4);
#pragma unroll
    for (int ks = 0; ks < 8; ++ks) ldmatrix_x4(qf[ks], qbase + swz(mrow, ks * 2 + mcol));
  }

  uint8_t* my_smem = smem + warp * kDecStages * 2 * kPageBytes;
  const int my_first = page_lo + warp;
  const int my_count = (my_first < page_hi) ? (page_hi - my_first + kDecWarps - 1) / kDecWarps : 0;

  auto issue = [&](int it) {
    if (it < my_count) {
      const int page = my_first + it * kDecWarps;
      const long blk = bt[page];
      const uint8_t* ksrc = reinterpret_cast<const uint8_t*>(p.k_cache) + (blk * p.Hkv + kvh) * kPageBytes;
      const uint8_t* vsrc = reinterpret_cast<const uint8_t*>(p.v_cache) + (blk * p.Hkv + kvh) * kPageBytes;
      uint8_t* kd = my_smem + (it % kDecStages) * 2 * kPageBytes;
      uint8_t* vd = kd + kPageBytes;
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const int chunk = lane + j * 32;          // 256 chunks of 16 B per page
        const int r = chunk >> 4, c = chunk & 15;
        cp_async_16(kd + swz(r, c), ksrc + chunk * 16);
        cp_async_16(vd + swz(r, c), vsrc + chunk * 16);
      }
    }
    cp_async_commit();
  };

  float o[16][4];
#pragma unroll
  for (int i = 0; i < 16; ++i) o[i][0] = o[i][1] = o[i][2] = o[i][3] = 0.f;
  float m_run = -INFINITY, l_run = 0.f;     // for row g (rows >= 8 are padding)

  issue(0);
  issue(1);
  for (int it = 0; it < my_count; ++it) {
    issue(it + 2);
    cp_async_wait<2>();
    __syncwarp();
    const uint32_t kb = smem_u32(my_smem + (it % kDecStages) * 2 * kPageBytes);
    const uint32_t vb = kb + kPageBytes;
    const int tok0 = (my_first + it * kDecWarps) * kPage;

    // ---- S = Q K^T : 16 x 16 ----
    float s[2][4] = {{0, 0, 0, 0}, {0, 0, 0, 0}};
#pragma unroll
    for (int ks = 0; ks < 8; ++ks) {
      // matrices: (tok 0-7, dims k0..7) (tok 0-7, dims k8..15) (tok 8-15, dims k0..7) (tok 8-15, dims k8..15)
      uint32_t kf[4];
      const int mrow = (lane & 7) + (lane >> 4) * 8;
      const int mcol = (lane >> 3) & 1;
      ldmatrix_x4(kf, kb + swz(mrow, ks * 2 + mcol));
      mma_bf16_16816(s[0], qf[ks], kf[0], kf[1]);
      mma_bf16_16816(s[1], qf[ks], kf[2], kf[3]);
    }
    // ---- mask + online softmax on row g (c0,c1 of each n-tile) ----
    float tmax = -INFINITY;
#pragma unroll
    for (int nt = 0; nt < 2; ++nt)
#pragma unroll
      for (int j = 0; j < 2; ++j) {
        const int tok = tok0 + nt * 8 + t4 * 2 + j;
        s[nt][j] = (tok >= ctx) ? -INFINITY : s[nt][j] * p.scale_log2;
        tmax = fmaxf(tmax, s[nt][j]);
      }
    tmax = fmaxf(tmax, __shfl_xor_sync(0xffffffffu, tmax, 1));
    tmax = fmaxf(tmax, __shfl_xor_sync(0xffffffffu, tmax, 2));
    const float m_new = fmaxf(m_run, tmax);
    const float corr = (m_new == -INFINITY) ? 1.f : exp2f(m_run - m_new);
    float psum = 0.f;
    float pr[2][2];
#pragma unroll
    for (int nt = 0; nt < 2; ++nt)
#pragma unroll
      for (int j = 0; j < 2; ++j) {
        pr[nt][j] = (m_new == -INFINITY) ? 0.f : exp2f(s[nt][j] - m_new);
        psum += pr[nt][j];
      }
    psum += __shfl_xor_sync(0xffffffffu, psum, 1);
    psum += __shfl_xor_sync(0xffffffffu, psum, 2);
    l_run = l_run * corr + psum;
    m_run = m_new;
#pragma unroll
    for (int i = 0; i < 16; ++i) { o[i][0] *= corr; o[i][1] *= corr; }
    // ---- O += P V ----
    uint32_t pa[4];
    pa[0] = pack_bf16x2(pr[0][0], pr[0][1]);   // row g,   tokens 2t,2t+1
    pa[1] = 0u;                                // row g+8 (padding)
    pa[2] = pack_bf16x2(pr[1][0], pr[1][1]);   // row g,   tokens 8+2t..
    pa[3] = 0u;
#pragma unroll
    for (int nd = 0; nd < 8; ++nd) {
      // matrices: (tok 0-7, dims n0..7) (tok 8-15, dims n0..7) (tok 0-7, dims n0+8..15) (tok 8-15, dims n0+8..15)
      uint32_t vf[4];
      const int mrow = (lane & 7) + ((lane >> 3) & 1) * 8;
      const int mcol = (lane >> 4);
      ldmatrix_x4_trans(vf, vb + swz(mrow, nd * 2 + mcol));
      mma_bf16_16816(o[nd * 2], pa, vf[0], vf[1]);
      mma_bf16_16816(o[nd * 2 + 1], pa, vf[2], vf[3]);
    }
    __syncwarp();
  }
  cp_async_wait<0>();
  __syncthreads();

  // ---- merge the 4 warps through shared memory (re-using the pipeline buffers) ----
  float* mo = reinterpret_cast<float*>(smem);                 // [warp][8][128]
  float* mml = mo + kDecWarps * 8 * kHeadDim;                 // [warp][8][2]
#pragma unroll
  for (int nd = 0; nd < 16; ++nd) {
    mo[(warp * 8 + g) * kHeadDim + nd * 8 + t4 * 2] = o[nd][0];
    mo[(warp * 8 + g) * kHeadDim + nd * 8 + t4 * 2 + 1] = o[nd][1];
  }
  if (t4 == 0) {
    mml[(warp * 8 + g) * 2] = m_run;
    mml[(warp * 8 + g) * 2 + 1] = l_run;
  }
  __syncthreads();
  for (int i = threadIdx.x; i < p.G * kHeadDim; i += blockDim.x) {
    const int r = i / kHeadDim, d = i % kHeadDim;
    float M = -INFINITY;
#pragma unroll
    for (int w = 0; w < kDecWarps; ++w) M = fmaxf(M, mml[(w * 8 + r) * 2]);
    float L = 0.f, acc = 0.f;
#pragma unroll
    for (int w = 0; w < kDecWarps; ++w) {
      const float mw = mml[(w * 8 + r) * 2];
      const float c = (mw == -INFINITY) ? 0.f : exp2f(mw - M);
      L += mml[(w * 8 + r) * 2 + 1] * c;
      acc += mo[(w * 8 + r) * kHeadDim + d] * c;
    }
    if (p.splits == 1) {
      p.out[(static_cast<long>(seq) * p.Hq + kvh * p.G + r) * kHeadDim + d] = __float2bfloat16(L > 0.f ? acc / L : 0.f);
    } else {
      const long base = ((static_cast<long>(seq) * p.Hkv + kvh) * p.splits + split) * 8 + r;
      p.part_o[base * kHeadDim + d] = acc;
      if (d == 0) {
        p.part_ml[base * 2] = M;
        p.part_ml[base * 2 + 1] = L;
      }
    }
  }
}

__global__ void decode_merge_splits_kernel(DecodeParams p) {
  const int seq = blockIdx.x, kvh = blockIdx.y;
  for (int i = threadIdx.x; i < p.G * kHeadDim; i += blockDim.x) {
    const int r = i / kHeadDim, d = i % kHeadDim;
    const long base = ((static_cast<long>(seq) * p.Hkv + kvh) * p.splits) * 8 + r;
    float M = -INFINITY;
    for (int s = 0; s < p.splits; ++s) M = fmaxf(M, p.part_ml[(base + s * 8) * 2]);
    float L = 0.f, acc = 0.f;
    for (int s = 0; s < p.splits; ++s) {
      const float ms = p.part_ml[(base + s * 8) * 2];
      const float c = (ms == -INFINITY) ? 0.f : exp2f(ms - M);
      L += p.part_ml[(base + s * 8) * 2 + 1] * c;
      acc += p.part_o[(base + s * 8) * kHeadDim + d] * c;
    }
    p.out[(static_cast<long>(seq) * p.Hq + kvh * p.G + r) * kHeadDim + d] = __float2bfloat16(L > 0.f ? acc / L : 0.f);
  }
}

}  // namespace nrl

using namespace nrl;

extern "C" cudaError_t nrl_kv_cache_write(const void* k, const void* v, long k_stride_t, long v_stride_t, void* k_cache,
                                          void* v_cache, const int* slot_mapping, const int* src_index, int T, int Hkv,
                                          int head_dim, int page, cudaStream_t s) {
  if (head_dim != kHeadDim || page != kPage) return cudaErrorInvalidValue;
  if (T == 0) return cudaSuccess;
  long total = static_cast<long>(T) * Hkv * (kHeadDim / 8);
  long blocks = (total + 255) / 256;
  if (blocks > 148L * 8) blocks = 148L * 8;
  kv_cache_write_kernel<<<static_cast<int>(blocks), 256, 0, s>>>(
      static_cast<const __nv_bfloat16*>(k), static_cast<const __nv_bfloat16*>(v), k_stride_t, v_stride_t,
      static_cast<__nv_bfloat16*>(k_cache), static_cast<__nv_bfloat16*>(v_cache), slot_mapping, src_index, T, Hkv);
  return cudaGetLastError();
}

extern "C" cudaError_t nrl_paged_decode(const void* q, long q_stride_s, const void* k_cache, const void* v_cache,
                                        const int* block_tables, const int* context_lens, void* out, float* part_o,
                                        float* part_ml, int S, int Hq, int Hkv, int head_dim, int page,
                                        int max_blocks, int splits, float scale, cudaStream_t s) {
  if (head_dim != kHeadDim || page != kPage || Hq % Hkv != 0 || Hq / Hkv > 8) return cudaErrorInvalidValue;
  if (S == 0) return cudaSuccess;
  DecodeParams p;
  p.q = static_cast<const __nv_bfloat16*>(q);
  p.k_cache = static_cast<const __nv_bfloat16*>(k_cache);
  p.v_cache = static_cast<const __nv_bfloat16*>(v_cache);
  p.block_tables = block_tables;
  p.context_lens = context_lens;
  p.out = static_cast<__nv_bfloat16*>(out);
  p.part_o = part_o;
  p.part_ml = part_ml;
  p.q_stride_s = q_stride_s;
  p.max_blocks = max_blocks;
  p.Hq = Hq;
  p.Hkv = Hkv;
  p.G = Hq / Hkv;
  p.splits = splits;
  p.scale_log2 = scale * 1.4426950408889634f;
  const int smem = kDecWarps * kDecStages * 2 * kPageBytes + 16 * kHeadDim * 2;
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(paged_decode_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e != cudaSuccess) return e;
    configured = true;
  }
  dim3 grid(S, Hkv, splits);
  paged_decode_kernel<<<grid, kDecWarps * 32, smem, s>>>(p);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return e;
  if (splits > 1) {
    decode_merge_splits_kernel<<<dim3(S, Hkv), 128, 0, s>>>(p);
    e = cudaGetLastError();
  }
  return e;
}
