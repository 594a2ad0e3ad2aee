// K5 sampler: temperature + nucleus (top-p) + seeded categorical draw, or arg-max at temperature 0.
//
// One CTA per sequence, four streaming passes over the row of logits (never sorted, never copied):
//   1. row maximum
//   2. probability-mass histogram, one bin per octave below the maximum   -> octave holding the top-p boundary
//   3. 64-slice histogram inside that octave                              -> cut-off resolved to 2^(1/64) = 1.1 %
//   4. kept mass per thread; draw u ~ Philox(seed, row, step) in (0, kept mass] and walk the kept tokens
// Histograms are lane-private shared-memory columns ([bin][lane]: bank == lane), so there are no atomics.
//
// Two kernels run that algorithm:
//   sample_top_p_smem_kernel   bf16 rows up to 4 x 110 k tokens, top_p >= 0.5: a thread-block CLUSTER of C CTAs owns a row, each CTA
//                              pulls its 1/C slice into shared memory ONCE with bulk async copies (TMA, mbarrier completion); the
//                              partial results (max, mass, candidate, mass above the candidate) are exchanged through distributed
//                              shared memory (st.shared::cluster + barrier.cluster).  HBM sees every logit once.  No histograms:
//                              the nucleus condition is verified for the drawn token only (exact; see the kernel).
//   sample_top_p_kernel        the streaming histogram kernel described above (fp32 logits, rows that are not 16-byte sliceable,
//                              larger vocabularies, top_p < 0.5): one CTA per row, every pass re-reads the row through L2.
// Reference: vLLM SamplingParams(temperature, top_p=0.95, seed=...) in vllm_generate
// (/root/reference/GRPO/grpo_trainer.py:127) and the T=0 greedy pass of ReMax (remax_trainer.py:167).
#include <curand_kernel.h>

#include <algorithm>
#include <cstdlib>
#include <string>

#include "common.cuh"
#include "kernels.h"

namespace nrl {

constexpr int kSampThreads = 256;       // 8 warps x 8 KB private histograms = 64 KB: 3 CTAs (rows) per SM
constexpr int kPerLane = kSampThreads / 32;

template <typename T>
struct RowVec;
template <>
struct RowVec<float> {
  static constexpr int N = 4;
  static NRL_DEVICE void load(const float* z, int vec, float (&o)[4]) {
    float4 a = reinterpret_cast<const float4*>(z)[vec];
    o[0] = a.x; o[1] = a.y; o[2] = a.z; o[3] = a.w;
  }
};
template <>
struct RowVec<__nv_bfloat16> {
  static constexpr int N = 8;
  static NRL_DEVICE void load(const __nv_bfloat16* z, int vec, float (&o)[8]) {
    uint4 a = reinterpret_cast<const uint4*>(z)[vec];
    uint32_t w[4] = {a.x, a.y, a.z, a.w};
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      float2 f = unpack_bf16x2(w[j]);
      o[2 * j] = f.x; o[2 * j + 1] = f.y;
    }
  }
};

// Visit every element of a row: thread t owns vectors t, t+T, t+2T, ...; four independent 16-byte loads are
// issued before any of them is consumed (memory-level parallelism: the passes are latency- not bandwidth-bound).
template <typename T, typename F>
NRL_DEVICE void for_each_elem(const T* z, int nvec, int V, F&& f) {
  constexpr int VN = RowVec<T>::N;
  constexpr int U = 4;
  for (int v0 = threadIdx.x; v0 < nvec; v0 += U * kSampThreads) {
    float x[U][VN];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int v = v0 + u * kSampThreads;
      if (v < nvec) RowVec<T>::load(z, v, x[u]);
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int v = v0 + u * kSampThreads;
      if (v < nvec) {
#pragma unroll
        for (int j = 0; j < VN; ++j)
          if (v * VN + j < V) f(v * VN + j, x[u][j]);
      }
    }
  }
}

NRL_DEVICE float block_reduce_sum(float v, float* red) {
  v = warp_sum(v);
  __syncthreads();
  if (lane_id() == 0) red[threadIdx.x >> 5] = v;
  __syncthreads();
  float t = (threadIdx.x < (blockDim.x >> 5)) ? red[threadIdx.x] : 0.f;
  if (threadIdx.x < 32) t = warp_sum(t);
  if (threadIdx.x == 0) red[0] = t;
  __syncthreads();
  return red[0];
}
NRL_DEVICE float block_reduce_max(float v, float* red) {
  v = warp_max(v);
  __syncthreads();
  if (lane_id() == 0) red[threadIdx.x >> 5] = v;
  __syncthreads();
  float t = (threadIdx.x < (blockDim.x >> 5)) ? red[threadIdx.x] : -INFINITY;
  if (threadIdx.x < 32) t = warp_max(t);
  if (threadIdx.x == 0) red[0] = t;
  __syncthreads();
  return red[0];
}

// Top-p by two-level (64 x 64) mass histograms with NO atomics: every lane owns a private column of a
// per-warp shared-memory histogram laid out [bin][lane] (bank == lane: conflict-free read-modify-write).
// Level 1 bins one octave of probability each (relative to the row maximum); level 2 splits the octave that
// contains the nucleus boundary into 64 slices, so the cut-off probability is resolved to 2^(1/64) = 1.1 %.
NRL_DEVICE void hist_clear(float* h) {
  for (int i = threadIdx.x; i < (kSampThreads / 32) * 64 * 32; i += kSampThreads) h[i] = 0.f;
  __syncthreads();
}
// sum the private columns: result in s_acc[0..63]
NRL_DEVICE void hist_reduce(const float* h, float* s_acc) {
  __syncthreads();
  // 64 bins x (warps*32) columns; thread t handles bin t/4 (kSampThreads = 256 -> 4 threads per bin)
  constexpr int TPB = kSampThreads / 64;
  const int bin = threadIdx.x / TPB, part = threadIdx.x % TPB;
  float t = 0.f;
  for (int w = 0; w < kSampThreads / 32; ++w)
    for (int l = part; l < 32; l += TPB) t += h[(w * 64 + bin) * 32 + l];
#pragma unroll
  for (int o = TPB / 2; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
  if (part == 0) s_acc[bin] = t;
  __syncthreads();
}

template <typename T>
__global__ void __launch_bounds__(kSampThreads) sample_top_p_kernel(const T* __restrict__ logits, long row_stride,
                                                                    int V, float inv_temp, float top_p,
                                                                    unsigned long long seed, unsigned long long step,
                                                                    const int* __restrict__ row_ids,
                                                                    const int* __restrict__ row_steps,
                                                                    int* __restrict__ out_tokens) {
  extern __shared__ float s_hist[];                      // [warps][64][32]
  __shared__ float red[32];
  __shared__ float s_acc[64];
  __shared__ float chunk_sum[kSampThreads];
  __shared__ float s_resid;
  __shared__ int s_owner;
  const int row = blockIdx.x;
  const T* z = logits + static_cast<long>(row) * row_stride;
  const int tid = threadIdx.x;
  constexpr int VN = RowVec<T>::N;
  const int nvec = (V + VN - 1) / VN;
  const float sc = inv_temp * 1.4426950408889634f;      // logits -> log2 domain
  float* my_hist = s_hist + (tid >> 5) * 64 * 32 + (tid & 31);

  // ---- pass 1: row maximum ----
  float mx = -INFINITY;
  for_each_elem(z, nvec, V, [&](int, float x) { mx = fmaxf(mx, x * sc); });
  mx = block_reduce_max(mx, red);

  float thresh = INFINITY;                                // keep tokens with (mx - z) < thresh  [octaves below the max]
  if (top_p < 1.f) {
    // ---- pass 2: one-octave mass histogram ----
    hist_clear(s_hist);
    for_each_elem(z, nvec, V, [&](int, float x) {
      const float d = mx - x * sc;
      const int b = min(63, static_cast<int>(d));
      my_hist[b * 32] += exp2f(-d);
    });
    hist_reduce(s_hist, s_acc);
    float total = 0.f;
    for (int k = 0; k < 64; ++k) total += s_acc[k];
    const float target = top_p * total;
    int B = 63;
    float before = 0.f, cum = 0.f;
    for (int k = 0; k < 64; ++k) {
      if (cum + s_acc[k] >= target) { B = k; before = cum; break; }
      cum += s_acc[k];
    }
    __syncthreads();
    // ---- pass 3: 1/64-octave histogram inside octave B ----
    hist_clear(s_hist);
    for_each_elem(z, nvec, V, [&](int, float x) {
      const float d = mx - x * sc;
      const int b = min(63, static_cast<int>(d));
      if (b == B) {
        const int f = min(63, static_cast<int>((d - static_cast<float>(B)) * 64.f));
        my_hist[f * 32] += exp2f(-d);
      }
    });
    hist_reduce(s_hist, s_acc);
    int Fc = 63;
    cum = before;
    for (int k = 0; k < 64; ++k) {
      cum += s_acc[k];
      if (cum >= target) { Fc = k; break; }
    }
    thresh = (B >= 63 && Fc >= 63) ? INFINITY : static_cast<float>(B) + static_cast<float>(Fc + 1) * (1.f / 64.f);
    __syncthreads();
  }

  // ---- pass 4: kept mass per thread (thread-major order), then draw and walk ----
  float mine = 0.f;
  for_each_elem(z, nvec, V, [&](int, float x) {
    const float d = mx - x * sc;
    if (d < thresh) mine += exp2f(-d);
  });
  chunk_sum[tid] = mine;
  __syncthreads();
  if (tid < 32) {
    float local[kPerLane];
    float tsum = 0.f;
#pragma unroll
    for (int i = 0; i < kPerLane; ++i) { local[i] = chunk_sum[tid * kPerLane + i]; tsum += local[i]; }
    float incl = tsum;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      float y = __shfl_up_sync(0xffffffffu, incl, o);
      if (tid >= o) incl += y;
    }
    const float total = __shfl_sync(0xffffffffu, incl, 31);
    float u = 0.f;
    if (tid == 0) {
      curandStatePhilox4_32_10_t st;
      curand_init(seed, static_cast<unsigned long long>(row_ids ? row_ids[row] : row),
                  step + (row_steps ? static_cast<unsigned long long>(row_steps[row]) : 0ull), &st);
      u = curand_uniform(&st) * total;              // (0, total]
    }
    u = __shfl_sync(0xffffffffu, u, 0);
    const float excl = incl - tsum;
    unsigned ball = __ballot_sync(0xffffffffu, (incl >= u) && (tsum > 0.f));
    int wl = ball ? (__ffs(ball) - 1) : 31;
    if (tid == wl) {
      float acc2 = excl;
      int owner = tid * kPerLane + kPerLane - 1;
      float resid = 0.f;
      bool hit = false;
      for (int i = 0; i < kPerLane; ++i) {
        if (!hit && local[i] > 0.f && acc2 + local[i] >= u) { owner = tid * kPerLane + i; resid = u - acc2; hit = true; }
        if (!hit) acc2 += local[i];
      }
      if (!hit) {   // numerical slack: fall back to the last thread with mass
        for (int i = kPerLane - 1; i >= 0; --i) if (local[i] > 0.f) { owner = tid * kPerLane + i; resid = local[i]; break; }
      }
      s_owner = owner;
      s_resid = resid;
    }
  }
  __syncthreads();
  if (tid == s_owner) {
    const float resid = s_resid;
    float acc2 = 0.f;
    int tok = -1, last_kept = -1;
    for (int v = tid; v < nvec && tok < 0; v += kSampThreads) {
      float x[VN];
      RowVec<T>::load(z, v, x);
#pragma unroll
      for (int j = 0; j < VN; ++j)
        if (tok < 0 && v * VN + j < V) {
          const float d = mx - x[j] * sc;
          if (d < thresh) {
            last_kept = v * VN + j;
            acc2 += exp2f(-d);
            if (acc2 >= resid) tok = v * VN + j;
          }
        }
    }
    if (tok < 0) tok = last_kept >= 0 ? last_kept : 0;
    out_tokens[row] = tok;
  }
}


// ---- shared-memory-resident variant: cluster of C CTAs per row, exact nucleus by rejection ----------------------------------
// With the row resident in shared memory a pass costs ~1 us, but the lane-private histogram updates of the streaming kernel are
// a serial read-modify-write chain per thread (ncu: IPC 1.0 at 8 warps per SM when ported as is).  This kernel needs no
// histogram: it draws a token from the FULL softmax (per-thread partial sums -> owner thread walks its elements), then
// verifies the nucleus condition for that token alone with one more streaming pass -- the token is kept iff the mass of the
// strictly more probable tokens is < top_p (exactly torch's `(cumsum - p) < top_p` rule, with no 1.1 % cut-off resolution); a
// rejected draw (probability ~ 1 - top_p) is redrawn from the next Philox offset, reusing the partial sums.
NRL_DEVICE uint32_t cluster_ctarank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
NRL_DEVICE uint32_t cluster_nctarank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_nctarank;" : "=r"(r)); return r; }
NRL_DEVICE void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// the address of `p` (a shared-memory object of this CTA) in the shared memory of CTA `rank` of the cluster
NRL_DEVICE uint32_t map_to_rank(const void* p, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(smem_u32(p)), "r"(rank));
  return r;
}
NRL_DEVICE void st_cluster_f32(uint32_t addr, float v) { asm volatile("st.shared::cluster.f32 [%0], %1;" ::"r"(addr), "f"(v) : "memory"); }
NRL_DEVICE void bulk_load_1d(void* smem_dst, const void* gsrc, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(smem_dst)),
               "l"(gsrc), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}

constexpr int kMaxCluster = 4;
constexpr int kSmemThreads = 512;                                 // 16 warps; two CTAs per SM when the slice is <= ~100 KB
constexpr int kSmemPerLane = kSmemThreads / 32;
constexpr int kSliceBytesMax = 216 * 1024;
constexpr int kMaxAttempts = 32;                                  // P(32 rejections) <= (1 - top_p)^32; dispatch keeps top_p >= 0.5
constexpr int kXSlots = 6;                                        // max, mass, and (candidate, mass above) x attempt parity

// every CTA of the cluster publishes `n` floats into slot [my rank] of `dst` in every CTA; after the cluster barrier all CTAs hold
// all contributions and combine them in rank order (identical result everywhere).  A CTA can run at most one cluster barrier
// ahead of its peers, and a slot is reused four barriers later at the earliest, so it is never rewritten while still being read.
NRL_DEVICE void cluster_publish(float (*dst)[2], const float* src, int n, uint32_t my_rank, uint32_t nrank) {
  if (threadIdx.x < static_cast<unsigned>(n) * nrank) {
    const uint32_t peer = threadIdx.x / n;
    const int k = threadIdx.x % n;
    st_cluster_f32(map_to_rank(&dst[my_rank][k], peer), src[k]);
  }
  cluster_sync_all();
}

NRL_DEVICE uint4 lds_v4(uint32_t saddr) {
  uint4 r;
  asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "r"(saddr));
  return r;
}
// visit the slice: thread t owns 16-byte vectors t, t + T, t + 2T, ... (ascending), four loads in flight
template <typename F>
NRL_DEVICE void slice_for_each(const __nv_bfloat16* zs, int nvec, F&& f) {
  constexpr int U = 4;
  const uint32_t zb = smem_u32(zs);
  for (int v0 = threadIdx.x; v0 < nvec; v0 += U * kSmemThreads) {
    uint4 raw[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int v = v0 + u * kSmemThreads;
      raw[u] = lds_v4(zb + static_cast<uint32_t>(min(v, nvec - 1)) * 16u);
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int v = v0 + u * kSmemThreads;
      if (v < nvec) {
        const uint32_t w[4] = {raw[u].x, raw[u].y, raw[u].z, raw[u].w};
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const float2 x = unpack_bf16x2(w[j]);
          f(v * 8 + 2 * j, x.x);
          f(v * 8 + 2 * j + 1, x.y);
        }
      }
    }
  }
}

__global__ void __launch_bounds__(kSmemThreads, 2) sample_top_p_smem_kernel(const __nv_bfloat16* __restrict__ logits, long row_stride,
                                                                            int V, int slice, float inv_temp, float top_p,
                                                                            unsigned long long seed, unsigned long long step,
                                                                            const int* __restrict__ row_ids,
                                                                            const int* __restrict__ row_steps,
                                                                            int* __restrict__ out_tokens) {
  extern __shared__ __align__(128) unsigned char s_dyn[];
  const __nv_bfloat16* zs = reinterpret_cast<const __nv_bfloat16*>(s_dyn);      // this CTA's slice of the row
  __shared__ __align__(8) uint64_t s_bar;
  __shared__ float red[32];
  __shared__ float s_pub[2];
  __shared__ float s_x[kXSlots][kMaxCluster][2];           // exchange slots, written by the peers
  __shared__ float chunk_sum[kSmemThreads];
  __shared__ float s_resid, s_u;
  __shared__ int s_owner;
  const uint32_t crank = cluster_ctarank(), csize = cluster_nctarank();
  const int row = blockIdx.x / static_cast<int>(csize);
  const int tid = threadIdx.x;
  const int e0 = static_cast<int>(crank) * slice;                                // first token of this CTA's slice
  const int Vl = max(0, min(slice, V - e0));                                     // tokens in the slice (multiple of 8)
  const int nvec = Vl / 8;
  const float sc = inv_temp * 1.4426950408889634f;                               // logits -> log2 domain

  // A CTA's shared memory may only be written by its peers once it is known to be running: arrive now, wait right before the
  // first exchange (the load and the first pass hide the barrier).
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  // ---- the only global read of the row: 1/C of it, as a few bulk async copies ----
  if (tid == 0) {
    mbar_init(&s_bar, 1);
    fence_mbar_init();
    const uint32_t bytes = static_cast<uint32_t>(Vl) * 2u;
    if (bytes > 0) {
      mbar_arrive_expect_tx(&s_bar, bytes);
      const unsigned char* src = reinterpret_cast<const unsigned char*>(logits + static_cast<long>(row) * row_stride + e0);
      for (uint32_t off = 0; off < bytes; off += 32768u) bulk_load_1d(s_dyn + off, src + off, min(32768u, bytes - off), &s_bar);
    } else {
      mbar_arrive(&s_bar);
    }
  }
  __syncthreads();                                         // the barrier is initialised before anyone polls it
  mbar_wait(&s_bar, 0);

  // ---- pass 1: row maximum ----
  float mx = -INFINITY;
  slice_for_each(zs, nvec, [&](int, float x) { mx = fmaxf(mx, x * sc); });
  mx = block_reduce_max(mx, red);
  if (tid == 0) s_pub[0] = mx;
  __syncthreads();
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");      // every CTA of the cluster has started
  cluster_publish(s_x[0], s_pub, 1, crank, csize);
  mx = s_x[0][0][0];
  for (uint32_t r = 1; r < csize; ++r) mx = fmaxf(mx, s_x[0][r][0]);

  // ---- pass 2: un-normalised probability mass per thread, per CTA, per row ----
  float mine = 0.f;
  slice_for_each(zs, nvec, [&](int, float x) { mine += exp2f(x * sc - mx); });
  chunk_sum[tid] = mine;
  const float cta_mass = block_reduce_sum(mine, red);      // (its barriers also publish chunk_sum)
  if (tid == 0) s_pub[0] = cta_mass;
  __syncthreads();
  cluster_publish(s_x[1], s_pub, 1, crank, csize);
  float Z = 0.f, excl_cta = 0.f;
  int last_with_mass = 0;
  for (uint32_t r = 0; r < csize; ++r) {
    if (r == crank) excl_cta = Z;
    Z += s_x[1][r][0];
    if (s_x[1][r][0] > 0.f) last_with_mass = static_cast<int>(r);
  }
  const unsigned long long rid = static_cast<unsigned long long>(row_ids ? row_ids[row] : row);
  const unsigned long long off0 = (step + (row_steps ? static_cast<unsigned long long>(row_steps[row]) : 0ull)) * kMaxAttempts;

#pragma unroll 1
  for (int attempt = 0; attempt < kMaxAttempts; ++attempt) {
    if (tid == 0) {
      curandStatePhilox4_32_10_t st;                       // a function of (seed, row id, tokens generated, attempt) only:
      curand_init(seed, rid, off0 + attempt, &st);         // every CTA of the cluster derives the same draw
      s_u = curand_uniform(&st) * Z;                       // (0, Z]
    }
    __syncthreads();
    const float u = s_u;
    int owner_cta = last_with_mass;
    {
      float acc = 0.f;
      for (uint32_t r = 0; r < csize; ++r) {
        if (s_x[1][r][0] > 0.f && acc + s_x[1][r][0] >= u) { owner_cta = static_cast<int>(r); break; }
        acc += s_x[1][r][0];
      }
    }
    if (owner_cta == static_cast<int>(crank)) {            // CTA-uniform branch
      const float ul = fminf(fmaxf(u - excl_cta, 0.f), cta_mass);     // the draw, local to this CTA's mass
      if (tid < 32) {
        float local[kSmemPerLane];
        float tsum = 0.f;
#pragma unroll
        for (int i = 0; i < kSmemPerLane; ++i) { local[i] = chunk_sum[tid * kSmemPerLane + i]; tsum += local[i]; }
        float incl = tsum;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
          float y = __shfl_up_sync(0xffffffffu, incl, o);
          if (tid >= o) incl += y;
        }
        const float excl = incl - tsum;
        const unsigned has = __ballot_sync(0xffffffffu, tsum > 0.f);
        const unsigned ball = __ballot_sync(0xffffffffu, (incl >= ul) && (tsum > 0.f));
        const int wl = ball ? (__ffs(ball) - 1) : (has ? 31 - __clz(has) : 31);      // numerical slack: the last lane with mass
        if (tid == wl) {
          float acc2 = excl;
          int owner = tid * kSmemPerLane + kSmemPerLane - 1;
          float resid = 0.f;
          bool hit = false;
#pragma unroll
          for (int i = 0; i < kSmemPerLane; ++i) {
            if (!hit && local[i] > 0.f && acc2 + local[i] >= ul) { owner = tid * kSmemPerLane + i; resid = ul - acc2; hit = true; }
            if (!hit) acc2 += local[i];
          }
          if (!hit) {   // numerical slack: the last thread with mass takes it
#pragma unroll
            for (int i = kSmemPerLane - 1; i >= 0; --i)
              if (!hit && local[i] > 0.f) { owner = tid * kSmemPerLane + i; resid = local[i]; hit = true; }
          }
          s_owner = owner;
          s_resid = resid;
        }
      }
      __syncthreads();
      if (tid == s_owner) {                                // walk this thread's elements in the order pass 2 summed them
        const float resid = s_resid;
        const uint32_t zb = smem_u32(zs);
        float acc2 = 0.f, d_tok = 0.f, d_last = 0.f;
        int tok = -1, last = 0;
        for (int v = tid; v < nvec && tok < 0; v += kSmemThreads) {
          const uint4 raw = lds_v4(zb + static_cast<uint32_t>(v) * 16u);
          const uint32_t w[4] = {raw.x, raw.y, raw.z, raw.w};
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const float2 xx = unpack_bf16x2(w[j >> 1]);
            const float e = ((j & 1) ? xx.y : xx.x) * sc - mx;
            if (tok < 0) {
              acc2 += exp2f(e);
              last = v * 8 + j;
              d_last = -e;
              if (acc2 >= resid) { tok = last; d_tok = -e; }
            }
          }
        }
        if (tok < 0) { tok = last; d_tok = d_last; }
        s_pub[0] = __int_as_float(e0 + tok);
        s_pub[1] = d_tok;                                  // octaves below the row maximum
      }
    }
    __syncthreads();
    float (*slot)[2] = s_x[2 + 2 * (attempt & 1)];
    cluster_publish(slot, s_pub, 2, crank, csize);         // only the owner's entry is read
    const int tok = __float_as_int(slot[owner_cta][0]);
    const float d_tok = slot[owner_cta][1];
    bool accept = top_p >= 1.f || attempt == kMaxAttempts - 1;
    if (!accept) {
      // ---- verification pass: mass of the strictly more probable tokens ----
      float above = 0.f;
      slice_for_each(zs, nvec, [&](int, float x) {
        const float d = mx - x * sc;
        above += (d < d_tok) ? exp2f(-d) : 0.f;
      });
      above = block_reduce_sum(above, red);
      if (tid == 0) s_pub[0] = above;
      __syncthreads();
      float (*slot2)[2] = s_x[3 + 2 * (attempt & 1)];
      cluster_publish(slot2, s_pub, 1, crank, csize);
      float tot = 0.f;
      for (uint32_t r = 0; r < csize; ++r) tot += slot2[r][0];
      accept = tot < top_p * Z;                            // identical on every CTA
    }
    if (accept) {
      if (crank == 0 && tid == 0) out_tokens[row] = tok;
      return;                                              // no remote access follows the last cluster barrier
    }
  }
}

template <typename T>
__global__ void __launch_bounds__(kSampThreads) argmax_kernel(const T* __restrict__ logits, long row_stride, int V,
                                                              int* __restrict__ out_tokens) {
  __shared__ float s_val[32];
  __shared__ int s_idx[32];
  const int row = blockIdx.x;
  const T* z = logits + static_cast<long>(row) * row_stride;
  float best = -INFINITY;
  int bi = 0x7fffffff;
  constexpr int VN = RowVec<T>::N;
  const int nvec = (V + VN - 1) / VN;
  for (int v = threadIdx.x; v < nvec; v += kSampThreads) {
    float x[VN];
    RowVec<T>::load(z, v, x);
#pragma unroll
    for (int j = 0; j < VN; ++j) {
      int i = v * VN + j;
      if (i < V && (x[j] > best || (x[j] == best && i < bi))) { best = x[j]; bi = i; }
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    float ov = __shfl_xor_sync(0xffffffffu, best, o);
    int oi = __shfl_xor_sync(0xffffffffu, bi, o);
    if (ov > best || (ov == best && oi < bi)) { best = ov; bi = oi; }
  }
  if (lane_id() == 0) { s_val[threadIdx.x >> 5] = best; s_idx[threadIdx.x >> 5] = bi; }
  __syncthreads();
  if (threadIdx.x < 32) {
    const bool has = threadIdx.x < (kSampThreads >> 5);
    best = has ? s_val[threadIdx.x] : -INFINITY;
    bi = has ? s_idx[threadIdx.x] : 0x7fffffff;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      float ov = __shfl_xor_sync(0xffffffffu, best, o);
      int oi = __shfl_xor_sync(0xffffffffu, bi, o);
      if (ov > best || (ov == best && oi < bi)) { best = ov; bi = oi; }
    }
    if (threadIdx.x == 0) out_tokens[row] = bi;
  }
}

}  // namespace nrl

using namespace nrl;

extern "C" cudaError_t nrl_sample(const void* logits, int is_bf16, long row_stride, int rows, int V, float temperature,
                                  float top_p, unsigned long long seed, unsigned long long step, const int* row_ids,
                                  const int* row_steps, int* out_tokens, int impl_req, cudaStream_t s) {
  if (rows == 0) return cudaSuccess;
  if (temperature == 0.f) {
    if (is_bf16)
      argmax_kernel<__nv_bfloat16><<<rows, kSampThreads, 0, s>>>(static_cast<const __nv_bfloat16*>(logits), row_stride, V, out_tokens);
    else
      argmax_kernel<float><<<rows, kSampThreads, 0, s>>>(static_cast<const float*>(logits), row_stride, V, out_tokens);
  } else {
    float inv_t = 1.f / temperature;
    if (top_p >= 1.f) top_p = 2.f;     // keep everything
    const int hist_bytes = (kSampThreads / 32) * 64 * 32 * 4;
    static bool configured = false;
    if (!configured) {
      cudaFuncSetAttribute(sample_top_p_kernel<__nv_bfloat16>, cudaFuncAttributeMaxDynamicSharedMemorySize, hist_bytes);
      cudaFuncSetAttribute(sample_top_p_kernel<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, hist_bytes);
      configured = true;
    }
    // bf16 rows that slice into 16-byte aligned pieces of <= 160 KB: the shared-memory-resident cluster kernel
    // impl_req: 0 = automatic (NANORLHF_SAMPLER_KERNEL=stream forces the fallback), 1 = streaming kernel, 2 = cluster kernel or error
    static const int env_impl = [] { const char* e = std::getenv("NANORLHF_SAMPLER_KERNEL"); return e && std::string(e) == "stream" ? 1 : 0; }();
    // (21 / 22 / 24: the cluster kernel with 1 / 2 / 4 CTAs per row, for benchmarking)
    int impl = impl_req != 0 ? impl_req : env_impl;
    // slices of <= 100 KB let two CTAs share an SM (one row's load overlaps another row's passes: 235 vs 254 us at 1024 x 151 936)
    int cmin = 1;
    while (cmin < kMaxCluster && (static_cast<long>(V) * 2 + cmin - 1) / cmin > kSliceBytesMax) cmin *= 2;
    int csize = cmin;
    while (csize < kMaxCluster && (static_cast<long>(V) * 2 + csize - 1) / csize > 100 * 1024) csize *= 2;
    if (impl > 20) { csize = std::max(cmin, std::min(impl - 20, kMaxCluster)); impl = 2; }
    const int slice = ((V + csize - 1) / csize + 7) / 8 * 8;
    const bool smem_ok = is_bf16 && V % 8 == 0 && row_stride % 8 == 0 && (reinterpret_cast<uintptr_t>(logits) & 15) == 0 &&
                         static_cast<long>(slice) * 2 <= kSliceBytesMax && top_p >= 0.5f;     // (rejection rate ~ 1 - top_p)
    if (impl == 2 && !smem_ok) return cudaErrorInvalidValue;
    if (smem_ok && impl != 1) {
      const int smem = slice * 2;
      static bool smem_configured = false;
      if (!smem_configured) {
        cudaError_t e = cudaFuncSetAttribute(sample_top_p_smem_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSliceBytesMax);
        if (e != cudaSuccess) return e;
        smem_configured = true;
      }
      cudaLaunchConfig_t cfg{};
      cfg.gridDim = dim3(static_cast<unsigned>(rows) * csize, 1, 1);
      cfg.blockDim = dim3(kSmemThreads, 1, 1);
      cfg.dynamicSmemBytes = smem;
      cfg.stream = s;
      cudaLaunchAttribute attr[1];
      attr[0].id = cudaLaunchAttributeClusterDimension;
      attr[0].val.clusterDim.x = csize; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
      cfg.attrs = attr;
      cfg.numAttrs = 1;
      return cudaLaunchKernelEx(&cfg, sample_top_p_smem_kernel, static_cast<const __nv_bfloat16*>(logits), row_stride, V, slice, inv_t,
                                top_p, seed, step, row_ids, row_steps, out_tokens);
    }
    if (is_bf16)
      sample_top_p_kernel<__nv_bfloat16><<<rows, kSampThreads, hist_bytes, s>>>(static_cast<const __nv_bfloat16*>(logits), row_stride, V,
                                                                       inv_t, top_p, seed, step, row_ids, row_steps, out_tokens);
    else
      sample_top_p_kernel<float><<<rows, kSampThreads, hist_bytes, s>>>(static_cast<const float*>(logits), row_stride, V, inv_t,
                                                               top_p, seed, step, row_ids, row_steps, out_tokens);
  }
  return cudaGetLastError();
}
