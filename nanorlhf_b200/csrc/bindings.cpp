// Python bindings (pybind11 via torch extension headers) for the sm_100a kernels.
// Kernels live in plain .cu translation units with a C launch API (kernels.h, gemm_sm100.h); this file
// only validates tensors, builds TMA descriptors and forwards the current CUDA stream.
#include <ATen/cuda/CUDAContext.h>
#include <c10/cuda/CUDAGuard.h>
#include <torch/extension.h>

#include <cstdlib>
#include <tuple>
#include <vector>

#include "gemm_sm100.h"
#include "gemm_tc.h"
#include "kernels.h"
#include "runtime.h"
#include "tma_host.h"

namespace {

inline cudaStream_t cur_stream() { return at::cuda::getCurrentCUDAStream().stream(); }

inline void check(cudaError_t e, const char* what) {
  TORCH_CHECK(e == cudaSuccess, what, ": ", cudaGetErrorString(e));
}

inline void check_bf16_2d(const torch::Tensor& t, const char* name) {
  TORCH_CHECK(t.is_cuda() && t.scalar_type() == torch::kBFloat16 && t.dim() == 2, name, " must be a 2D CUDA bf16 tensor");
  TORCH_CHECK(t.stride(1) == 1 && (t.stride(0) * 2) % 16 == 0 && (reinterpret_cast<uintptr_t>(t.data_ptr()) & 15) == 0,
              name, " must be row-major with 16-byte aligned rows");
}

int num_sms() {
  static int n = 0;
  if (n == 0) {
    int dev;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
  }
  return n;
}

int pick_block_n(int64_t M, int64_t N, bool allow_192 = false) {
  // score = (SM occupancy of the last wave) x (useful fraction of the padded N) x (per-tile efficiency of the shape)
  const int sms = num_sms();
  const int64_t num_m = (M + 127) / 128;
  auto score = [&](int bn, double tile_eff) {
    const int64_t num_n = (N + bn - 1) / bn, tiles = num_m * num_n;
    const int64_t waves = (tiles + sms - 1) / sms;
    return tile_eff * static_cast<double>(tiles) / static_cast<double>(waves * sms) * static_cast<double>(N) /
           static_cast<double>(num_n * bn);
  };
  int best = 256;
  double best_s = score(256, 1.0);
  if (allow_192 && score(192, 0.95) > best_s) { best = 192; best_s = score(192, 0.95); }
  if (N <= 128 || score(128, 0.82) > best_s) best = 128;
  return best;
}

// D[M,N] = A[M,K] @ B[N,K]^T (+ bias[N])
torch::Tensor gemm_bf16(const torch::Tensor& a, const torch::Tensor& b, const c10::optional<torch::Tensor>& bias,
                        c10::optional<torch::Tensor> out_opt, int64_t block_n, int64_t act) {
  check_bf16_2d(a, "a");
  check_bf16_2d(b, "b");
  const int64_t M = a.size(0), K = a.size(1), N = b.size(0);
  TORCH_CHECK(b.size(1) == K, "inner dimensions differ");
  TORCH_CHECK(K % 8 == 0 && N % 8 == 0, "K and N must be multiples of 8");
  c10::cuda::CUDAGuard guard(a.device());
  torch::Tensor out = out_opt.has_value() ? *out_opt : torch::empty({M, N}, a.options());
  check_bf16_2d(out, "out");
  TORCH_CHECK(out.size(0) == M && out.size(1) == N, "out has the wrong shape");
  if (M == 0) return out;
  // block_n == 512 selects the experimental 2-CTA kernel (256 x 256 tile per CTA pair); NRL_GEMM_2CTA=1 makes it the
  // automatic choice for N >= 256.  Not the default until it has been validated on hardware.
  static const bool auto_2cta = getenv("NRL_GEMM_2CTA") && atoi(getenv("NRL_GEMM_2CTA")) == 1;
  if (block_n == 512 || (block_n == 0 && auto_2cta && N >= 256 && M >= 256)) {
    CUtensorMap tmA2 = nrl::make_tma_2d(a.data_ptr(), M, K, a.stride(0) * 2, 128, 64, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2);
    CUtensorMap tmB2 = nrl::make_tma_2d(b.data_ptr(), N, K, b.stride(0) * 2, 128, 64, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2);
    CUtensorMap tmD2 = nrl::make_tma_2d(out.data_ptr(), M, N, out.stride(0) * 2, 128, 64, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2);
    nrl::GemmParams p2{};
    p2.M = M; p2.N = N; p2.K = K; p2.n_splits = 1; p2.scale = 1.f;
    p2.act = static_cast<int>(act);
    if (bias.has_value()) {
      TORCH_CHECK(bias->is_cuda() && bias->scalar_type() == torch::kBFloat16 && bias->numel() == N && bias->is_contiguous());
      p2.bias = reinterpret_cast<const __nv_bfloat16*>(bias->data_ptr());
    }
    check(nrl_gemm_bf16_tn_2cta(&tmA2, &tmB2, &tmD2, &p2, 0, num_sms(), cur_stream()), "gemm_bf16_2cta");
    return out;
  }
  const int bn = block_n > 0 ? static_cast<int>(block_n) : pick_block_n(M, N, /*allow_192=*/true);
  CUtensorMap tmA = nrl::make_tma_2d(a.data_ptr(), M, K, a.stride(0) * 2, 128, 64, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2);
  CUtensorMap tmB = nrl::make_tma_2d(b.data_ptr(), N, K, b.stride(0) * 2, bn, 64, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2);
  CUtensorMap tmD = nrl::make_tma_2d(out.data_ptr(), M, N, out.stride(0) * 2, 128, 64, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2);
  nrl::GemmParams p{};
  p.M = M; p.N = N; p.K = K; p.n_splits = 1; p.scale = 1.f;
  p.act = static_cast<int>(act);
  if (bias.has_value()) {
    TORCH_CHECK(bias->is_cuda() && bias->scalar_type() == torch::kBFloat16 && bias->numel() == N && bias->is_contiguous(),
                "bias must be a contiguous CUDA bf16 [N] tensor");
    p.bias = reinterpret_cast<const __nv_bfloat16*>(bias->data_ptr());
  }
  check(nrl_gemm_bf16_tn(&tmA, &tmB, &tmD, &p, bn, nrl::EPI_STORE, num_sms(), cur_stream()), "gemm_bf16");
  return out;
}

// ---- general tcgen05 GEMM (gemm_tc.cu) ---------------------------------------------------------------------------
// Tile-shape choice: expected throughput of a candidate = (per-tile efficiency measured on B200, profiles/gemm_bench_r2.json)
// x (useful fraction of the padded M and N) x (tiles / (waves x units)).  CG = 2 halves the number of scheduling units.
struct TcChoice { int cg, bn; };
TcChoice pick_tc_tile(int64_t M, int64_t N, bool a_mn, bool b_mn, bool f32_out) {
  struct Cand { int cg, bn; double eff; };
  static const Cand all[] = {{2, 256, 1.00}, {2, 192, 0.95}, {2, 128, 0.82}, {1, 256, 0.86}, {1, 192, 0.80}, {1, 128, 0.72}, {1, 64, 0.45}};
  const int sms = num_sms();
  TcChoice best{1, 128};
  double best_s = -1.0;
  for (const Cand& c : all) {
    if (a_mn && !((c.cg == 1 && c.bn != 192) || (c.cg == 2 && c.bn == 256))) continue;      // instantiated wgrad shapes
    if (f32_out && !a_mn) continue;
    const int64_t tm = c.cg * 128;
    const int64_t num_m = (M + tm - 1) / tm, num_n = (N + c.bn - 1) / c.bn, tiles = num_m * num_n;
    const int64_t units = sms / c.cg, waves = (tiles + units - 1) / units;
    const double s = c.eff * (static_cast<double>(M) / (num_m * tm)) * (static_cast<double>(N) / (num_n * c.bn)) *
                     (static_cast<double>(tiles) / (waves * units));
    if (s > best_s) { best_s = s; best = {c.cg, c.bn}; }
  }
  return best;
}

inline void check_tc_operand(const torch::Tensor& t, const char* name) {
  TORCH_CHECK(t.is_cuda() && t.scalar_type() == torch::kBFloat16 && t.dim() == 2, name, " must be a 2D CUDA bf16 tensor");
  TORCH_CHECK(t.stride(1) == 1 && (t.stride(0) * 2) % 16 == 0 && (reinterpret_cast<uintptr_t>(t.data_ptr()) & 15) == 0,
              name, " must have unit inner stride and 16-byte aligned rows");
}

// D[M,N] = alpha * (opA(a) opB(b)^T + opA(a2) opB(b2)^T) (+bias) (GELU).
//   a_mn = false: a is [M, K] (contraction contiguous);  a_mn = true: a is [K, M] (contraction = rows).  Same for b / N.
//   Output: bf16 `out` [M, N] (allocated when absent) or, when `out_f32` is given, fp32 out_f32 (+)= result.
torch::Tensor gemm_tc(const torch::Tensor& a, const torch::Tensor& b, bool a_mn, bool b_mn, const c10::optional<torch::Tensor>& a2,
                      const c10::optional<torch::Tensor>& b2, const c10::optional<torch::Tensor>& bias, int64_t act, double alpha,
                      c10::optional<torch::Tensor> out_opt, c10::optional<torch::Tensor> out_f32, bool accumulate, int64_t cg_req,
                      int64_t bn_req, int64_t splitk_req) {
  check_tc_operand(a, "a");
  check_tc_operand(b, "b");
  const int64_t M = a_mn ? a.size(1) : a.size(0), K = a_mn ? a.size(0) : a.size(1);
  const int64_t N = b_mn ? b.size(1) : b.size(0);
  TORCH_CHECK((b_mn ? b.size(0) : b.size(1)) == K, "gemm_tc: contraction lengths of a and b differ");
  TORCH_CHECK(N % 8 == 0, "gemm_tc: N must be a multiple of 8");      // operand alignment is checked per tensor (row strides)
  int64_t K2 = 0;
  if (a2.has_value() || b2.has_value()) {
    TORCH_CHECK(a2.has_value() && b2.has_value(), "gemm_tc: a2 and b2 go together");
    check_tc_operand(*a2, "a2");
    check_tc_operand(*b2, "b2");
    K2 = a_mn ? a2->size(0) : a2->size(1);
    TORCH_CHECK((a_mn ? a2->size(1) : a2->size(0)) == M && (b_mn ? b2->size(1) : b2->size(0)) == N &&
                (b_mn ? b2->size(0) : b2->size(1)) == K2, "gemm_tc: second operand pair has the wrong shape");
  }
  c10::cuda::CUDAGuard guard(a.device());
  const bool f32 = out_f32.has_value();
  torch::Tensor out;
  if (f32) {
    TORCH_CHECK(out_f32->is_cuda() && out_f32->scalar_type() == torch::kFloat32 && out_f32->dim() == 2 && out_f32->size(0) == M &&
                out_f32->size(1) == N && out_f32->stride(1) == 1 && out_f32->stride(0) % 4 == 0 &&
                (reinterpret_cast<uintptr_t>(out_f32->data_ptr()) & 15) == 0, "gemm_tc: out_f32 must be fp32 [M, N], 16-byte aligned rows");
    TORCH_CHECK(!bias.has_value() && act == 0, "gemm_tc: the fp32 epilogue has no bias / activation");
    out = *out_f32;
  } else {
    out = out_opt.has_value() ? *out_opt : torch::empty({M, N}, a.options());
    check_tc_operand(out, "out");
    TORCH_CHECK(out.size(0) == M && out.size(1) == N, "gemm_tc: out has the wrong shape");
  }
  if (M == 0 || N == 0) return out;
  // ---- split-K: few output tiles, long contraction (rank-r projections, LoRA weight gradients) ----
  const bool bias_ok = !bias.has_value() || ((reinterpret_cast<uintptr_t>(bias->data_ptr()) & 15) == 0 && bias->is_contiguous());
  if (cg_req <= 0 && bn_req <= 0 && K2 == 0 && !f32 && bias_ok && act == 0 && (a_mn == b_mn || !a_mn) && splitk_req != 0) {
    const int bn = N <= 64 ? 64 : 128;
    const int64_t tiles = ((M + 127) / 128) * ((N + bn - 1) / bn), num_kb = (K + 63) / 64;
    const int sms = num_sms();
    // measured (bench/sampler_bench.py, profiles/skinny_gemm_split_sweep_r2.json): the last CTA's reduction reads `splits` partial tiles, so
    // beyond ~4 k-ranges the fix-up costs more than the shorter main loop saves, and work items past one wave (148) cost a second wave
    int64_t splits = splitk_req > 0 ? splitk_req : std::min<int64_t>(std::min<int64_t>(sms / std::max<int64_t>(tiles, 1), num_kb / 4), 4);
    // (any N: the adapter-input gradient dA = t'^T x is [r, K_in] -- 12 or 24 tiles of a 6.6 k-long contraction)
    if (splitk_req > 0 || (tiles * 2 <= sms && splits >= 2)) {
      splits = std::max<int64_t>(1, std::min<int64_t>(splits, num_kb));
      const int64_t per = (num_kb + splits - 1) / splits;
      splits = (num_kb + per - 1) / per;                                 // no empty k-range
      // Captured CUDA graphs (decode step, training micro-step) hold these addresses: a workspace that is outgrown is
      // retired, never freed.  (Automatic dispatch needs at most num_sms tiles of 128 x 128 floats = 9.7 MB < the 16 MB minimum.)
      static thread_local torch::Tensor ws, counters;
      static thread_local std::vector<torch::Tensor> retired;
      const int64_t need = splits * tiles * 128 * bn;
      if (!ws.defined() || ws.numel() < need || ws.device() != a.device()) {
        if (ws.defined()) retired.push_back(ws);
        ws = torch::empty({std::max<int64_t>(need, 1 << 22)}, a.options().dtype(torch::kFloat32));
      }
      if (!counters.defined() || counters.numel() < tiles || counters.device() != a.device()) {
        if (counters.defined()) retired.push_back(counters);
        counters = torch::zeros({std::max<int64_t>(tiles, 4096)}, a.options().dtype(torch::kInt32));
      }
      const auto BF = CU_TENSOR_MAP_DATA_TYPE_BFLOAT16;
      CUtensorMap maps2[2];
      maps2[0] = a_mn ? nrl::make_tma_2d(a.data_ptr(), a.size(0), a.size(1), a.stride(0) * 2, 64, 64, BF, 2)
                      : nrl::make_tma_2d(a.data_ptr(), a.size(0), a.size(1), a.stride(0) * 2, 128, 64, BF, 2);
      maps2[1] = b_mn ? nrl::make_tma_2d(b.data_ptr(), b.size(0), b.size(1), b.stride(0) * 2, 64, 64, BF, 2)
                      : nrl::make_tma_2d(b.data_ptr(), b.size(0), b.size(1), b.stride(0) * 2, bn, 64, BF, 2);
      nrl::tc::TcParams ps{};
      ps.M = static_cast<int>(M); ps.N = static_cast<int>(N); ps.K = static_cast<int>(K); ps.K2 = 0;
      ps.alpha = static_cast<float>(alpha);
      if (bias.has_value()) {
        TORCH_CHECK(bias->is_cuda() && bias->scalar_type() == torch::kBFloat16 && bias->numel() == N, "gemm_tc: bias must be bf16 [N]");
        ps.bias = reinterpret_cast<const __nv_bfloat16*>(bias->data_ptr());
      }
      ps.splits = static_cast<int>(splits);
      ps.ws = ws.data_ptr<float>();
      ps.counters = counters.data_ptr<int>();
      ps.out_bf16 = reinterpret_cast<__nv_bfloat16*>(out.data_ptr());
      ps.out_stride = out.stride(0);
      check(nrl_gemm_tc_splitk(maps2, &ps, bn, a_mn ? 1 : 0, b_mn ? 1 : 0, sms, cur_stream()), "gemm_tc_splitk");
      return out;
    }
  }
  TcChoice ch = pick_tc_tile(M, N, a_mn, b_mn, f32);
  if (cg_req > 0) ch.cg = static_cast<int>(cg_req);
  if (bn_req > 0) ch.bn = static_cast<int>(bn_req);
  const auto BF = CU_TENSOR_MAP_DATA_TYPE_BFLOAT16;
  auto map_a = [&](const torch::Tensor& t) {
    return a_mn ? nrl::make_tma_2d(t.data_ptr(), t.size(0), t.size(1), t.stride(0) * 2, 64, 64, BF, 2)
                : nrl::make_tma_2d(t.data_ptr(), t.size(0), t.size(1), t.stride(0) * 2, 128, 64, BF, 2);
  };
  auto map_b = [&](const torch::Tensor& t) {
    return b_mn ? nrl::make_tma_2d(t.data_ptr(), t.size(0), t.size(1), t.stride(0) * 2, 64, 64, BF, 2)
                : nrl::make_tma_2d(t.data_ptr(), t.size(0), t.size(1), t.stride(0) * 2, ch.bn / ch.cg, 64, BF, 2);
  };
  CUtensorMap maps[5];
  maps[0] = map_a(a);
  maps[1] = map_b(b);
  maps[2] = K2 > 0 ? map_a(*a2) : maps[0];
  maps[3] = K2 > 0 ? map_b(*b2) : maps[1];
  maps[4] = f32 ? maps[0] : nrl::make_tma_2d(out.data_ptr(), M, N, out.stride(0) * 2, 128, 64, BF, 2);
  nrl::tc::TcParams p{};
  p.M = static_cast<int>(M); p.N = static_cast<int>(N); p.K = static_cast<int>(K); p.K2 = static_cast<int>(K2);
  p.alpha = static_cast<float>(alpha);
  p.act = static_cast<int>(act);
  if (bias.has_value()) {
    TORCH_CHECK(bias->is_cuda() && bias->scalar_type() == torch::kBFloat16 && bias->numel() == N && bias->is_contiguous(),
                "bias must be a contiguous CUDA bf16 [N] tensor");
    p.bias = reinterpret_cast<const __nv_bfloat16*>(bias->data_ptr());
  }
  if (f32) {
    p.out_f32 = out.data_ptr<float>();
    p.out_f32_stride = out.stride(0);
    p.accumulate = accumulate ? 1 : 0;
  }
  check(nrl_gemm_tc(maps, &p, ch.cg, ch.bn, a_mn ? 1 : 0, b_mn ? 1 : 0, f32 ? nrl::tc::EPI_F32 : nrl::tc::EPI_BF16, num_sms(),
                    cur_stream()), "gemm_tc (no kernel instantiated for this cg / block_n / operand-major combination?)");
  return out;
}

// out[b] = a[b] . b[b]^T for b < B: a [B, M, K], b [B, N, K], out [B, M, N]; any batch / row strides (inner stride 1), e.g. the
// per-head slices of packed [T, H, d] projections (DeBERTa bias tables) -- one launch, 3D TMA, no per-head copies.
torch::Tensor gemm_tc_batched(const torch::Tensor& a, const torch::Tensor& b, c10::optional<torch::Tensor> out_opt, int64_t cg_req, int64_t bn_req) {
  auto chk = [](const torch::Tensor& t, const char* name) {
    TORCH_CHECK(t.is_cuda() && t.scalar_type() == torch::kBFloat16 && t.dim() == 3 && t.stride(2) == 1 && (t.stride(1) * 2) % 16 == 0 &&
                (t.stride(0) * 2) % 16 == 0 && (reinterpret_cast<uintptr_t>(t.data_ptr()) & 15) == 0, name,
                " must be a 3D CUDA bf16 tensor with unit inner stride and 16-byte aligned row / batch strides");
  };
  chk(a, "a");
  chk(b, "b");
  const int64_t B = a.size(0), M = a.size(1), K = a.size(2), N = b.size(1);
  TORCH_CHECK(b.size(0) == B && b.size(2) == K && N % 8 == 0, "gemm_tc_batched: shapes differ");
  c10::cuda::CUDAGuard guard(a.device());
  torch::Tensor out = out_opt.has_value() ? *out_opt : torch::empty({B, M, N}, a.options());
  chk(out, "out");
  TORCH_CHECK(out.size(0) == B && out.size(1) == M && out.size(2) == N, "gemm_tc_batched: out has the wrong shape");
  if (B == 0 || M == 0) return out;
  TcChoice ch = pick_tc_tile(M * B, N, false, false, false);
  if (ch.bn == 192 || ch.bn == 64) ch.bn = N > 128 ? 256 : 128;
  if (cg_req > 0) ch.cg = static_cast<int>(cg_req);
  if (bn_req > 0) ch.bn = static_cast<int>(bn_req);
  const auto BF = CU_TENSOR_MAP_DATA_TYPE_BFLOAT16;
  CUtensorMap maps[3];
  maps[0] = nrl::make_tma_3d(a.data_ptr(), B, M, K, a.stride(1) * 2, a.stride(0) * 2, 128, 64, BF, 2);
  maps[1] = nrl::make_tma_3d(b.data_ptr(), B, N, K, b.stride(1) * 2, b.stride(0) * 2, ch.bn / ch.cg, 64, BF, 2);
  maps[2] = nrl::make_tma_3d(out.data_ptr(), B, M, N, out.stride(1) * 2, out.stride(0) * 2, 128, 64, BF, 2);
  nrl::tc::TcParams p{};
  p.M = static_cast<int>(M); p.N = static_cast<int>(N); p.K = static_cast<int>(K); p.alpha = 1.f; p.batch = static_cast<int>(B);
  check(nrl_gemm_tc_batched(maps, &p, ch.cg, ch.bn, num_sms(), cur_stream()), "gemm_tc_batched");
  return out;
}

// act[M, F] = silu(x Wg^T) * (x Wu^T) on gemm_tc (cta_group 1 | 2): w_interleaved [2F, K], rows per 32 features [32 gate | 32 up]
torch::Tensor gemm_tc_swiglu(const torch::Tensor& a, const torch::Tensor& w_interleaved, c10::optional<torch::Tensor> out_opt, int64_t cg_req,
                             int64_t bn_req) {
  check_tc_operand(a, "a");
  check_tc_operand(w_interleaved, "w");
  const int64_t M = a.size(0), K = a.size(1), N2 = w_interleaved.size(0);
  TORCH_CHECK(w_interleaved.size(1) == K && N2 % 128 == 0, "gemm_tc_swiglu: 2F must be a multiple of 128");
  c10::cuda::CUDAGuard guard(a.device());
  torch::Tensor out = out_opt.has_value() ? *out_opt : torch::empty({M, N2 / 2}, a.options());
  check_tc_operand(out, "out");
  if (M == 0) return out;
  TcChoice ch = pick_tc_tile(M, N2, false, false, false);
  if (ch.bn == 192 || ch.bn == 64) ch.bn = (N2 % 256 == 0) ? 256 : 128;         // instantiated SwiGLU tiles
  if (N2 % ch.bn != 0) ch.bn = 128;
  if (cg_req > 0) ch.cg = static_cast<int>(cg_req);
  if (bn_req > 0) ch.bn = static_cast<int>(bn_req);
  const auto BF = CU_TENSOR_MAP_DATA_TYPE_BFLOAT16;
  CUtensorMap maps[5];
  maps[0] = nrl::make_tma_2d(a.data_ptr(), M, K, a.stride(0) * 2, 128, 64, BF, 2);
  maps[1] = nrl::make_tma_2d(w_interleaved.data_ptr(), N2, K, w_interleaved.stride(0) * 2, ch.bn / ch.cg, 64, BF, 2);
  maps[2] = maps[0];
  maps[3] = maps[1];
  maps[4] = nrl::make_tma_2d(out.data_ptr(), M, N2 / 2, out.stride(0) * 2, 128, 64, BF, 2);
  nrl::tc::TcParams p{};
  p.M = static_cast<int>(M); p.N = static_cast<int>(N2); p.K = static_cast<int>(K); p.alpha = 1.f;
  check(nrl_gemm_tc(maps, &p, ch.cg, ch.bn, 0, 0, nrl::tc::EPI_SWIGLU, num_sms(), cur_stream()), "gemm_tc_swiglu");
  return out;
}

// act[M, F] = silu(x Wg^T) * (x Wu^T) with W_gu_interleaved[2F, K] (rows: per 32 features [32 gate | 32 up])
torch::Tensor gemm_swiglu(const torch::Tensor& a, const torch::Tensor& w_interleaved, c10::optional<torch::Tensor> out_opt) {
  check_bf16_2d(a, "a");
  check_bf16_2d(w_interleaved, "w");
  const int64_t M = a.size(0), K = a.size(1), N2 = w_interleaved.size(0);
  TORCH_CHECK(w_interleaved.size(1) == K && N2 % 128 == 0 && K % 8 == 0, "2F must be a multiple of 128");
  c10::cuda::CUDAGuard guard(a.device());
  torch::Tensor out = out_opt.has_value() ? *out_opt : torch::empty({M, N2 / 2}, a.options());
  check_bf16_2d(out, "out");
  if (M == 0) return out;
  static const bool auto_2cta = getenv("NRL_GEMM_2CTA") && atoi(getenv("NRL_GEMM_2CTA")) == 1;     // experimental
  if (auto_2cta && M >= 256 && N2 % 256 == 0) {
    CUtensorMap tA = nrl::make_tma_2d(a.data_ptr(), M, K, a.stride(0) * 2, 128, 64, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2);
    CUtensorMap tB = nrl::make_tma_2d(w_interleaved.data_ptr(), N2, K, w_interleaved.stride(0) * 2, 128, 64, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2);
    CUtensorMap tD = nrl::make_tma_2d(out.data_ptr(), M, N2 / 2, out.stride(0) * 2, 128, 64, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2);
    nrl::GemmParams p2{};
    p2.M = M; p2.N = N2; p2.K = K; p2.n_splits = 1; p2.scale = 1.f;
    check(nrl_gemm_bf16_tn_2cta(&tA, &tB, &tD, &p2, 1, num_sms(), cur_stream()), "gemm_swiglu_2cta");
    return out;
  }
  const int bn = pick_block_n(M, N2);
  CUtensorMap tmA = nrl::make_tma_2d(a.data_ptr(), M, K, a.stride(0) * 2, 128, 64, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2);
  CUtensorMap tmB = nrl::make_tma_2d(w_interleaved.data_ptr(), N2, K, w_interleaved.stride(0) * 2, bn, 64, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2);
  CUtensorMap tmD = nrl::make_tma_2d(out.data_ptr(), M, N2 / 2, out.stride(0) * 2, 128, 64, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2);
  nrl::GemmParams p{};
  p.M = M; p.N = N2; p.K = K; p.n_splits = 1; p.scale = 1.f;
  check(nrl_gemm_bf16_tn(&tmA, &tmB, &tmD, &p, bn, nrl::EPI_SWIGLU, num_sms(), cur_stream()), "gemm_swiglu");
  return out;
}

// ---- fp8 rollout path ------------------------------------------------------------------------------------
std::tuple<torch::Tensor, torch::Tensor> quant_rows_e4m3(const torch::Tensor& x) {
  check_bf16_2d(x, "x");
  c10::cuda::CUDAGuard guard(x.device());
  const int64_t M = x.size(0), K = x.size(1);
  TORCH_CHECK(K % 16 == 0, "K must be a multiple of 16");
  torch::Tensor q = torch::empty({M, K}, x.options().dtype(torch::kUInt8));
  torch::Tensor sc = torch::empty({M}, x.options().dtype(torch::kFloat32));
  check(nrl_quant_rows_e4m3(x.data_ptr(), x.stride(0), q.data_ptr(), K, sc.data_ptr<float>(), M, K, cur_stream()), "quant_rows_e4m3");
  return {q, sc};
}

// D[M,N] (bf16) = (Aq[M,K] Bq[N,K]^T) * a_scale[M] * b_scale[N] (+bias); swiglu=true: Bq rows interleaved, D[M,N/2]
torch::Tensor gemm_fp8(const torch::Tensor& aq, const torch::Tensor& a_scale, const torch::Tensor& bq, const torch::Tensor& b_scale,
                       const c10::optional<torch::Tensor>& bias, bool swiglu) {
  TORCH_CHECK(aq.is_cuda() && aq.scalar_type() == torch::kUInt8 && aq.dim() == 2 && aq.stride(1) == 1 && aq.stride(0) % 16 == 0);
  TORCH_CHECK(bq.is_cuda() && bq.scalar_type() == torch::kUInt8 && bq.dim() == 2 && bq.stride(1) == 1 && bq.stride(0) % 16 == 0);
  const int64_t M = aq.size(0), K = aq.size(1), N = bq.size(0);
  TORCH_CHECK(bq.size(1) == K && K % 16 == 0 && N % 8 == 0);
  TORCH_CHECK(a_scale.scalar_type() == torch::kFloat32 && a_scale.numel() == M && b_scale.scalar_type() == torch::kFloat32 && b_scale.numel() == N);
  c10::cuda::CUDAGuard guard(aq.device());
  const int64_t No = swiglu ? N / 2 : N;
  torch::Tensor out = torch::empty({M, No}, aq.options().dtype(torch::kBFloat16));
  if (M == 0) return out;
  const int bn = pick_block_n(M, N);
  CUtensorMap tmA = nrl::make_tma_2d(aq.data_ptr(), M, K, aq.stride(0), 128, 128, CU_TENSOR_MAP_DATA_TYPE_UINT8, 1);
  CUtensorMap tmB = nrl::make_tma_2d(bq.data_ptr(), N, K, bq.stride(0), bn, 128, CU_TENSOR_MAP_DATA_TYPE_UINT8, 1);
  CUtensorMap tmD = nrl::make_tma_2d(out.data_ptr(), M, No, out.stride(0) * 2, 128, 64, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2);
  nrl::GemmParams p{};
  p.M = M; p.N = N; p.K = K; p.n_splits = 1; p.scale = 1.f;
  p.row_scale = a_scale.data_ptr<float>();
  p.col_scale = b_scale.data_ptr<float>();
  if (bias.has_value()) p.bias = reinterpret_cast<const __nv_bfloat16*>(bias->data_ptr());
  check(nrl_gemm_fp8_tn(&tmA, &tmB, &tmD, &p, bn, swiglu ? nrl::EPI_SWIGLU : nrl::EPI_STORE, num_sms(), cur_stream()), "gemm_fp8");
  return out;
}

// fp8 rollout GEMM on gemm_tc (cta_group 1 | 2): D (bf16) = (Aq Bq^T) * a_scale[M] * b_scale[N] (+bias); swiglu: interleaved Bq rows, D [M, N/2]
torch::Tensor gemm_tc_fp8(const torch::Tensor& aq, const torch::Tensor& a_scale, const torch::Tensor& bq, const torch::Tensor& b_scale,
                          const c10::optional<torch::Tensor>& bias, bool swiglu, int64_t cg_req, int64_t bn_req) {
  TORCH_CHECK(aq.is_cuda() && aq.scalar_type() == torch::kUInt8 && aq.dim() == 2 && aq.stride(1) == 1 && aq.stride(0) % 16 == 0);
  TORCH_CHECK(bq.is_cuda() && bq.scalar_type() == torch::kUInt8 && bq.dim() == 2 && bq.stride(1) == 1 && bq.stride(0) % 16 == 0);
  const int64_t M = aq.size(0), K = aq.size(1), N = bq.size(0);
  TORCH_CHECK(bq.size(1) == K && K % 16 == 0 && N % 8 == 0);
  TORCH_CHECK(a_scale.scalar_type() == torch::kFloat32 && a_scale.numel() == M && b_scale.scalar_type() == torch::kFloat32 && b_scale.numel() == N);
  c10::cuda::CUDAGuard guard(aq.device());
  const int64_t No = swiglu ? N / 2 : N;
  torch::Tensor out = torch::empty({M, No}, aq.options().dtype(torch::kBFloat16));
  if (M == 0) return out;
  TcChoice ch = pick_tc_tile(M, N, false, false, false);
  if (ch.bn == 64) ch.bn = 128;
  if (swiglu && ch.bn == 192) ch.bn = (N % 256 == 0) ? 256 : 128;
  if (swiglu && N % ch.bn != 0) ch.bn = 128;
  if (cg_req > 0) ch.cg = static_cast<int>(cg_req);
  if (bn_req > 0) ch.bn = static_cast<int>(bn_req);
  CUtensorMap maps[3];
  maps[0] = nrl::make_tma_2d(aq.data_ptr(), M, K, aq.stride(0), 128, 128, CU_TENSOR_MAP_DATA_TYPE_UINT8, 1);
  maps[1] = nrl::make_tma_2d(bq.data_ptr(), N, K, bq.stride(0), ch.bn / ch.cg, 128, CU_TENSOR_MAP_DATA_TYPE_UINT8, 1);
  maps[2] = nrl::make_tma_2d(out.data_ptr(), M, No, out.stride(0) * 2, 128, 64, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2);
  nrl::tc::TcParams p{};
  p.M = static_cast<int>(M); p.N = static_cast<int>(N); p.K = static_cast<int>(K); p.alpha = 1.f;
  p.row_scale = a_scale.data_ptr<float>();
  p.col_scale = b_scale.data_ptr<float>();
  if (bias.has_value()) p.bias = reinterpret_cast<const __nv_bfloat16*>(bias->data_ptr());
  check(nrl_gemm_tc_fp8(maps, &p, ch.cg, ch.bn, swiglu ? 1 : 0, num_sms(), cur_stream()), "gemm_tc_fp8");
  return out;
}

int pick_splits(int64_t M, int64_t N, int bn) {
  int64_t num_m = (M + 127) / 128, num_n = (N + bn - 1) / bn;
  int64_t s = (num_sms() + num_m - 1) / num_m;       // enough work items to cover the SMs
  if (num_m >= num_sms()) s = 1;
  s = std::max<int64_t>(1, std::min<int64_t>(s, std::min<int64_t>(num_n, 64)));
  return static_cast<int>(s);
}

// fused lm-head log-prob forward: returns (logp, entropy, lse), all fp32 [M]
std::tuple<torch::Tensor, torch::Tensor, torch::Tensor> lmhead_logprob_fwd(const torch::Tensor& hidden,
                                                                          const torch::Tensor& weight,
                                                                          const torch::Tensor& targets,
                                                                          double inv_temperature, int64_t n_splits) {
  check_bf16_2d(hidden, "hidden");
  check_bf16_2d(weight, "weight");
  const int64_t M = hidden.size(0), K = hidden.size(1), V = weight.size(0);
  TORCH_CHECK(weight.size(1) == K && K % 8 == 0);
  TORCH_CHECK(targets.is_cuda() && targets.scalar_type() == torch::kInt32 && targets.numel() == M && targets.is_contiguous(),
              "targets must be contiguous CUDA int32 [M]");
  c10::cuda::CUDAGuard guard(hidden.device());
  auto fopt = hidden.options().dtype(torch::kFloat32);
  torch::Tensor logp = torch::empty({M}, fopt), ent = torch::empty({M}, fopt), lse = torch::empty({M}, fopt);
  if (M == 0) return {logp, ent, lse};
  const int bn = 256;
  const int splits = n_splits > 0 ? static_cast<int>(n_splits) : pick_splits(M, V, bn);
  torch::Tensor partials = torch::empty({splits, M, 4}, fopt);
  CUtensorMap tmA = nrl::make_tma_2d(hidden.data_ptr(), M, K, hidden.stride(0) * 2, 128, 64, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2);
  CUtensorMap tmB = nrl::make_tma_2d(weight.data_ptr(), V, K, weight.stride(0) * 2, bn, 64, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2);
  nrl::GemmParams p{};
  p.M = M; p.N = V; p.K = K; p.n_splits = splits; p.scale = static_cast<float>(inv_temperature);
  p.targets = targets.data_ptr<int>();
  p.partials = partials.data_ptr<float>();
  check(nrl_gemm_bf16_tn(&tmA, &tmB, &tmA, &p, bn, nrl::EPI_LOGPROB, num_sms(), cur_stream()), "lmhead_logprob_fwd");
  check(nrl_lmhead_combine(partials.data_ptr<float>(), M, splits, logp.data_ptr<float>(), ent.data_ptr<float>(),
                           lse.data_ptr<float>(), cur_stream()), "lmhead_combine");
  return {logp, ent, lse};
}

// K-LP backward, stage 1: dZ[M,V] = (onehot - softmax) * grad_logp / T, recomputed through the tensor cores
torch::Tensor lmhead_dlogits(const torch::Tensor& hidden, const torch::Tensor& weight, const torch::Tensor& targets,
                             const torch::Tensor& lse, const torch::Tensor& grad_logp, double inv_temperature) {
  check_bf16_2d(hidden, "hidden");
  check_bf16_2d(weight, "weight");
  const int64_t M = hidden.size(0), K = hidden.size(1), V = weight.size(0);
  TORCH_CHECK(V % 8 == 0 && K % 8 == 0);
  TORCH_CHECK(targets.scalar_type() == torch::kInt32 && lse.scalar_type() == torch::kFloat32 &&
              grad_logp.scalar_type() == torch::kFloat32 && lse.is_contiguous() && grad_logp.is_contiguous());
  c10::cuda::CUDAGuard guard(hidden.device());
  torch::Tensor dz = torch::empty({M, V}, hidden.options());
  if (M == 0) return dz;
  const int bn = 256;
  CUtensorMap tmA = nrl::make_tma_2d(hidden.data_ptr(), M, K, hidden.stride(0) * 2, 128, 64, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2);
  CUtensorMap tmB = nrl::make_tma_2d(weight.data_ptr(), V, K, weight.stride(0) * 2, bn, 64, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2);
  CUtensorMap tmD = nrl::make_tma_2d(dz.data_ptr(), M, V, dz.stride(0) * 2, 128, 64, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2);
  nrl::GemmParams p{};
  p.M = M; p.N = V; p.K = K; p.n_splits = 1; p.scale = static_cast<float>(inv_temperature);
  p.targets = targets.data_ptr<int>();
  p.lse = lse.data_ptr<float>();
  p.grad_logp = grad_logp.data_ptr<float>();
  check(nrl_gemm_bf16_tn(&tmA, &tmB, &tmD, &p, bn, nrl::EPI_DLOGITS, num_sms(), cur_stream()), "lmhead_dlogits");
  return dz;
}

// K-BC: out[N_out, K_in] = W + scale * lora_B[N_out, r] @ lora_A[r, K_in], on the tcgen05 GEMM (contraction = r)
// mc_out_ptr != 0: `out` is this rank's view of a symmetric (NVLS-multicast-bound) arena and mc_out_ptr the multicast address of
// its first element -- the epilogue then stores with multimem.st into every rank's arena at once.
void lora_merge(const torch::Tensor& w, const torch::Tensor& lora_a, const torch::Tensor& lora_b, double scale, torch::Tensor out,
                int64_t mc_out_ptr) {
  check_bf16_2d(w, "w");
  check_bf16_2d(lora_b, "lora_B");
  check_bf16_2d(out, "out");
  TORCH_CHECK(lora_a.is_cuda() && lora_a.scalar_type() == torch::kBFloat16 && lora_a.dim() == 2);
  const int64_t N_out = w.size(0), K_in = w.size(1), r = lora_a.size(0);
  TORCH_CHECK(lora_b.size(0) == N_out && lora_b.size(1) == r && lora_a.size(1) == K_in && out.sizes() == w.sizes());
  TORCH_CHECK(r % 8 == 0 && K_in % 8 == 0);
  c10::cuda::CUDAGuard guard(w.device());
  torch::Tensor a_t = lora_a.t().contiguous();                      // [K_in, r]: the B operand, contraction-major
  const int bn = 256;
  CUtensorMap tmA = nrl::make_tma_2d(lora_b.data_ptr(), N_out, r, lora_b.stride(0) * 2, 128, 64, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2);
  CUtensorMap tmB = nrl::make_tma_2d(a_t.data_ptr(), K_in, r, a_t.stride(0) * 2, bn, 64, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2);
  CUtensorMap tmD = nrl::make_tma_2d(out.data_ptr(), N_out, K_in, out.stride(0) * 2, 128, 64, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2);
  nrl::GemmParams p{};
  p.M = N_out; p.N = K_in; p.K = r; p.n_splits = 1; p.scale = static_cast<float>(scale);
  p.addend = reinterpret_cast<const __nv_bfloat16*>(w.data_ptr());
  p.addend_stride = w.stride(0);
  if (mc_out_ptr != 0) {
    TORCH_CHECK(K_in % 8 == 0 && (mc_out_ptr & 15) == 0, "multicast merge needs 16-byte aligned rows");
    p.mc_out = reinterpret_cast<__nv_bfloat16*>(mc_out_ptr);
    p.mc_stride = out.stride(0);
  }
  check(nrl_gemm_bf16_tn(&tmA, &tmB, &tmD, &p, bn, nrl::EPI_MERGE, num_sms(), cur_stream()), "lora_merge");
}

// ---- elementwise --------------------------------------------------------------------------------
std::tuple<torch::Tensor, torch::Tensor, torch::Tensor> rmsnorm(const torch::Tensor& x, const torch::Tensor& w, double eps,
                                                                const c10::optional<torch::Tensor>& residual,
                                                                bool want_rstd) {
  TORCH_CHECK(x.is_cuda() && x.scalar_type() == torch::kBFloat16 && x.is_contiguous() && x.dim() == 2);
  TORCH_CHECK(w.scalar_type() == torch::kBFloat16 && w.is_contiguous() && w.numel() == x.size(1));
  c10::cuda::CUDAGuard guard(x.device());
  const int rows = x.size(0), d = x.size(1);
  torch::Tensor y = torch::empty_like(x);
  torch::Tensor res_out, rstd;
  const void* res_ptr = nullptr;
  void* res_out_ptr = nullptr;
  if (residual.has_value()) {
    TORCH_CHECK(residual->is_contiguous() && residual->sizes() == x.sizes() && residual->scalar_type() == torch::kBFloat16);
    res_out = torch::empty_like(x);
    res_ptr = residual->data_ptr();
    res_out_ptr = res_out.data_ptr();
  }
  if (want_rstd) rstd = torch::empty({rows}, x.options().dtype(torch::kFloat32));
  check(nrl_rmsnorm(x.data_ptr(), res_ptr, w.data_ptr(), y.data_ptr(), res_out_ptr,
                    want_rstd ? rstd.data_ptr<float>() : nullptr, rows, d, static_cast<float>(eps), cur_stream()), "rmsnorm");
  return {y, res_out, rstd};
}

torch::Tensor rmsnorm_bwd(const torch::Tensor& x, const torch::Tensor& w, const torch::Tensor& gy, const torch::Tensor& rstd) {
  TORCH_CHECK(x.is_contiguous() && gy.is_contiguous() && x.scalar_type() == torch::kBFloat16 && gy.scalar_type() == torch::kBFloat16);
  c10::cuda::CUDAGuard guard(x.device());
  torch::Tensor gx = torch::empty_like(x);
  check(nrl_rmsnorm_bwd(x.data_ptr(), w.data_ptr(), gy.data_ptr(), rstd.data_ptr<float>(), gx.data_ptr(), x.size(0), x.size(1), cur_stream()),
        "rmsnorm_bwd");
  return gx;
}

// x: [T, H, D] (last two dims contiguous, arbitrary token stride); cos/sin: [T, D/2] fp32
torch::Tensor rope(const torch::Tensor& x, const torch::Tensor& cos_t, const torch::Tensor& sin_t, double sin_sign, bool inplace) {
  TORCH_CHECK(x.is_cuda() && x.scalar_type() == torch::kBFloat16 && x.dim() == 3 && x.stride(2) == 1 && x.stride(1) == x.size(2));
  TORCH_CHECK(cos_t.scalar_type() == torch::kFloat32 && cos_t.is_contiguous() && sin_t.is_contiguous() &&
              cos_t.size(0) == x.size(0) && cos_t.size(1) * 2 == x.size(2));
  c10::cuda::CUDAGuard guard(x.device());
  torch::Tensor y = inplace ? x : torch::empty({x.size(0), x.size(1), x.size(2)}, x.options());
  check(nrl_rope(x.data_ptr(), y.data_ptr(), cos_t.data_ptr<float>(), sin_t.data_ptr<float>(), x.size(0), x.size(1), x.size(2),
                 x.stride(0), y.stride(0), static_cast<float>(sin_sign), cur_stream()), "rope");
  return y;
}

torch::Tensor swiglu(const torch::Tensor& gu) {
  TORCH_CHECK(gu.is_cuda() && gu.scalar_type() == torch::kBFloat16 && gu.is_contiguous() && gu.dim() == 2 && gu.size(1) % 16 == 0);
  c10::cuda::CUDAGuard guard(gu.device());
  const int64_t F = gu.size(1) / 2;
  torch::Tensor out = torch::empty({gu.size(0), F}, gu.options());
  auto base = reinterpret_cast<const __nv_bfloat16*>(gu.data_ptr());
  check(nrl_swiglu(base, base + F, 2 * F, out.data_ptr(), gu.size(0), F, cur_stream()), "swiglu");
  return out;
}

torch::Tensor swiglu_bwd(const torch::Tensor& gu, const torch::Tensor& gout) {
  TORCH_CHECK(gu.is_contiguous() && gout.is_contiguous() && gout.scalar_type() == torch::kBFloat16);
  c10::cuda::CUDAGuard guard(gu.device());
  const int64_t F = gu.size(1) / 2;
  torch::Tensor dgu = torch::empty_like(gu);
  auto base = reinterpret_cast<const __nv_bfloat16*>(gu.data_ptr());
  auto dbase = reinterpret_cast<__nv_bfloat16*>(dgu.data_ptr());
  check(nrl_swiglu_bwd(base, base + F, 2 * F, gout.data_ptr(), dbase, dbase + F, 2 * F, gu.size(0), F, cur_stream()), "swiglu_bwd");
  return dgu;
}

// separate gate / up tensors (no concatenation copy)
torch::Tensor swiglu_pair(const torch::Tensor& gate, const torch::Tensor& up) {
  TORCH_CHECK(gate.is_cuda() && gate.scalar_type() == torch::kBFloat16 && gate.is_contiguous() && up.is_contiguous() &&
              gate.dim() == 2 && gate.sizes() == up.sizes() && gate.size(1) % 8 == 0);
  c10::cuda::CUDAGuard guard(gate.device());
  torch::Tensor out = torch::empty_like(gate);
  check(nrl_swiglu(gate.data_ptr(), up.data_ptr(), gate.size(1), out.data_ptr(), gate.size(0), gate.size(1), cur_stream()), "swiglu_pair");
  return out;
}

std::tuple<torch::Tensor, torch::Tensor> swiglu_pair_bwd(const torch::Tensor& gate, const torch::Tensor& up, const torch::Tensor& gout) {
  TORCH_CHECK(gate.is_contiguous() && up.is_contiguous() && gout.is_contiguous());
  c10::cuda::CUDAGuard guard(gate.device());
  torch::Tensor dg = torch::empty_like(gate), du = torch::empty_like(up);
  check(nrl_swiglu_bwd(gate.data_ptr(), up.data_ptr(), gate.size(1), gout.data_ptr(), dg.data_ptr(), du.data_ptr(), gate.size(1),
                       gate.size(0), gate.size(1), cur_stream()), "swiglu_pair_bwd");
  return {dg, du};
}

// ---- RL kernels -----------------------------------------------------------------------------------
std::tuple<torch::Tensor, torch::Tensor> gae_scan(const torch::Tensor& rewards, const c10::optional<torch::Tensor>& values,
                                                  double gamma, double lam) {
  TORCH_CHECK(rewards.is_cuda() && rewards.scalar_type() == torch::kFloat32 && rewards.is_contiguous() && rewards.dim() == 2);
  c10::cuda::CUDAGuard guard(rewards.device());
  torch::Tensor adv = torch::empty_like(rewards), ret;
  const float* vptr = nullptr;
  float* rptr = nullptr;
  if (values.has_value()) {
    TORCH_CHECK(values->is_contiguous() && values->scalar_type() == torch::kFloat32 && values->sizes() == rewards.sizes());
    ret = torch::empty_like(rewards);
    vptr = values->data_ptr<float>();
    rptr = ret.data_ptr<float>();
  }
  check(nrl_gae_scan(rewards.data_ptr<float>(), vptr, adv.data_ptr<float>(), rptr, rewards.size(0), rewards.size(1),
                     static_cast<float>(gamma), static_cast<float>(lam), cur_stream()), "gae_scan");
  return {adv, ret};
}

// returns (grad_unnorm [n] fp32, acc [10] fp32)
std::tuple<torch::Tensor, torch::Tensor> policy_loss(const torch::Tensor& new_lp, const torch::Tensor& old_lp,
                                                     const torch::Tensor& adv, const torch::Tensor& mask,
                                                     const c10::optional<torch::Tensor>& ref_lp, double cliprange,
                                                     double kl_coef) {
  for (const auto* t : {&new_lp, &old_lp, &adv})
    TORCH_CHECK(t->is_cuda() && t->scalar_type() == torch::kFloat32 && t->is_contiguous());
  TORCH_CHECK(mask.scalar_type() == torch::kBool && mask.is_contiguous() && mask.numel() == new_lp.numel());
  c10::cuda::CUDAGuard guard(new_lp.device());
  torch::Tensor grad = torch::empty_like(new_lp);
  torch::Tensor acc = torch::zeros({10}, new_lp.options());
  const float* rp = nullptr;
  if (ref_lp.has_value()) {
    TORCH_CHECK(ref_lp->is_contiguous() && ref_lp->scalar_type() == torch::kFloat32);
    rp = ref_lp->data_ptr<float>();
  }
  check(nrl_policy_loss(new_lp.data_ptr<float>(), old_lp.data_ptr<float>(), adv.data_ptr<float>(),
                        reinterpret_cast<const uint8_t*>(mask.data_ptr<bool>()), rp, static_cast<float>(cliprange),
                        static_cast<float>(kl_coef), new_lp.numel(), grad.data_ptr<float>(), acc.data_ptr<float>(), cur_stream()),
        "policy_loss");
  return {grad, acc};
}

std::tuple<torch::Tensor, torch::Tensor> value_loss(const torch::Tensor& vpred, const torch::Tensor& vold,
                                                    const torch::Tensor& ret, const torch::Tensor& mask, double clip) {
  for (const auto* t : {&vpred, &vold, &ret})
    TORCH_CHECK(t->is_cuda() && t->scalar_type() == torch::kFloat32 && t->is_contiguous());
  TORCH_CHECK(mask.scalar_type() == torch::kBool && mask.is_contiguous());
  c10::cuda::CUDAGuard guard(vpred.device());
  torch::Tensor grad = torch::empty_like(vpred);
  torch::Tensor acc = torch::zeros({4}, vpred.options());
  check(nrl_value_loss(vpred.data_ptr<float>(), vold.data_ptr<float>(), ret.data_ptr<float>(),
                       reinterpret_cast<const uint8_t*>(mask.data_ptr<bool>()), static_cast<float>(clip), vpred.numel(),
                       grad.data_ptr<float>(), acc.data_ptr<float>(), cur_stream()), "value_loss");
  return {grad, acc};
}

void adamw_flat(torch::Tensor param, const torch::Tensor& grad, torch::Tensor m, torch::Tensor v, double lr, double beta1,
                double beta2, double eps, double wd, int64_t step, double grad_scale, c10::optional<torch::Tensor> master) {
  TORCH_CHECK(param.is_cuda() && param.scalar_type() == torch::kBFloat16 && grad.scalar_type() == torch::kBFloat16);
  TORCH_CHECK(param.is_contiguous() && grad.is_contiguous() && m.is_contiguous() && v.is_contiguous());
  TORCH_CHECK(m.scalar_type() == v.scalar_type() && (m.scalar_type() == torch::kFloat32 || m.scalar_type() == torch::kBFloat16));
  TORCH_CHECK(param.numel() == grad.numel() && param.numel() == m.numel() && param.numel() == v.numel());
  c10::cuda::CUDAGuard guard(param.device());
  nrl::AdamHyper h;
  h.lr = lr; h.beta1 = beta1; h.beta2 = beta2; h.eps = eps; h.wd = wd;
  h.step_size = static_cast<float>(lr / (1.0 - std::pow(beta1, static_cast<double>(step))));
  h.inv_bc2 = static_cast<float>(1.0 / (1.0 - std::pow(beta2, static_cast<double>(step))));
  h.grad_scale = grad_scale;
  float* mw = nullptr;
  if (master.has_value()) {
    TORCH_CHECK(master->is_cuda() && master->scalar_type() == torch::kFloat32 && master->is_contiguous() &&
                master->numel() == param.numel(), "master must be a contiguous fp32 copy of param");
    mw = master->data_ptr<float>();
  }
  check(nrl_adamw_flat(param.data_ptr(), grad.data_ptr(), m.data_ptr(), v.data_ptr(), mw, param.numel(),
                       m.scalar_type() == torch::kBFloat16 ? 1 : 0, h, cur_stream()), "adamw_flat");
}

// ---- sampler kernels ------------------------------------------------------------------------------
torch::Tensor sample(const torch::Tensor& logits, double temperature, double top_p, int64_t seed, int64_t step,
                     const c10::optional<torch::Tensor>& row_ids, const c10::optional<torch::Tensor>& row_steps,
                     c10::optional<torch::Tensor> out_opt, int64_t impl) {
  TORCH_CHECK(logits.is_cuda() && logits.dim() == 2 && logits.stride(1) == 1);
  const bool bf16 = logits.scalar_type() == torch::kBFloat16;
  TORCH_CHECK(bf16 || logits.scalar_type() == torch::kFloat32);
  TORCH_CHECK((logits.stride(0) * logits.element_size()) % 16 == 0, "logit rows must be 16-byte aligned");
  c10::cuda::CUDAGuard guard(logits.device());
  torch::Tensor out = out_opt.has_value() ? *out_opt : torch::empty({logits.size(0)}, logits.options().dtype(torch::kInt32));
  const int* rid = nullptr;
  if (row_ids.has_value()) rid = row_ids->data_ptr<int>();
  const int* rst = nullptr;
  if (row_steps.has_value()) rst = row_steps->data_ptr<int>();
  check(nrl_sample(logits.data_ptr(), bf16 ? 1 : 0, logits.stride(0), logits.size(0), logits.size(1),
                   static_cast<float>(temperature), static_cast<float>(top_p), static_cast<unsigned long long>(seed),
                   static_cast<unsigned long long>(step), rid, rst, out.data_ptr<int>(), static_cast<int>(impl), cur_stream()), "sample");
  return out;
}

void kv_cache_write(const torch::Tensor& k, const torch::Tensor& v, torch::Tensor k_cache, torch::Tensor v_cache,
                    const torch::Tensor& slot_mapping, const c10::optional<torch::Tensor>& src_index) {
  TORCH_CHECK(k.dim() == 3 && v.dim() == 3 && k.stride(2) == 1 && k.stride(1) == k.size(2) && v.stride(2) == 1 && v.stride(1) == v.size(2));
  TORCH_CHECK(k_cache.dim() == 4 && k_cache.is_contiguous() && v_cache.is_contiguous());
  TORCH_CHECK(slot_mapping.scalar_type() == torch::kInt32 && slot_mapping.is_contiguous());
  const int* src = nullptr;
  if (src_index.has_value()) {
    TORCH_CHECK(src_index->scalar_type() == torch::kInt32 && src_index->numel() == slot_mapping.numel());
    src = src_index->data_ptr<int>();
  } else {
    TORCH_CHECK(slot_mapping.numel() == k.size(0));
  }
  c10::cuda::CUDAGuard guard(k.device());
  check(nrl_kv_cache_write(k.data_ptr(), v.data_ptr(), k.stride(0), v.stride(0), k_cache.data_ptr(), v_cache.data_ptr(),
                           slot_mapping.data_ptr<int>(), src, slot_mapping.numel(), k.size(1), k.size(2), k_cache.size(2), cur_stream()),
        "kv_cache_write");
}

torch::Tensor paged_decode(const torch::Tensor& q, const torch::Tensor& k_cache, const torch::Tensor& v_cache,
                           const torch::Tensor& block_tables, const torch::Tensor& context_lens, double scale,
                           int64_t splits, c10::optional<torch::Tensor> out_opt) {
  TORCH_CHECK(q.is_cuda() && q.scalar_type() == torch::kBFloat16 && q.dim() == 3 && q.stride(2) == 1 && q.stride(1) == q.size(2));
  TORCH_CHECK(block_tables.scalar_type() == torch::kInt32 && block_tables.is_contiguous() && context_lens.scalar_type() == torch::kInt32);
  c10::cuda::CUDAGuard guard(q.device());
  const int S = q.size(0), Hq = q.size(1), D = q.size(2), Hkv = k_cache.size(1);
  torch::Tensor out = out_opt.has_value() ? *out_opt : torch::empty({S, Hq, D}, q.options());
  torch::Tensor po, pml;
  float *pop = nullptr, *pmlp = nullptr;
  if (splits > 1) {
    po = torch::empty({S, Hkv, splits, 8, D}, q.options().dtype(torch::kFloat32));
    pml = torch::empty({S, Hkv, splits, 8, 2}, q.options().dtype(torch::kFloat32));
    pop = po.data_ptr<float>();
    pmlp = pml.data_ptr<float>();
  }
  check(nrl_paged_decode(q.data_ptr(), q.stride(0), k_cache.data_ptr(), v_cache.data_ptr(), block_tables.data_ptr<int>(),
                         context_lens.data_ptr<int>(), out.data_ptr(), pop, pmlp, S, Hq, Hkv, D, k_cache.size(2),
                         block_tables.size(1), static_cast<int>(splits), static_cast<float>(scale), cur_stream()),
        "paged_decode");
  return out;
}

// ---- packed-varlen flash attention ----------------------------------------------------------------------
inline void check_thd(const torch::Tensor& t, const char* name) {
  TORCH_CHECK(t.is_cuda() && t.scalar_type() == torch::kBFloat16 && t.dim() == 3 && t.stride(2) == 1 && t.stride(1) == t.size(2),
              name, " must be [T, H, D] bf16 with contiguous heads");
}

std::tuple<torch::Tensor, torch::Tensor> attn_varlen_fwd(const torch::Tensor& q, const torch::Tensor& k, const torch::Tensor& v,
                                                         const torch::Tensor& cu_seqlens, int64_t max_seqlen, double scale,
                                                         bool causal, const c10::optional<torch::Tensor>& rel_a,
                                                         const c10::optional<torch::Tensor>& rel_b,
                                                         const c10::optional<torch::Tensor>& lut) {
  check_thd(q, "q"); check_thd(k, "k"); check_thd(v, "v");
  TORCH_CHECK(cu_seqlens.scalar_type() == torch::kInt32 && cu_seqlens.is_contiguous());
  c10::cuda::CUDAGuard guard(q.device());
  const int T = q.size(0), Hq = q.size(1), D = q.size(2), Hkv = k.size(1);
  torch::Tensor out = torch::empty({T, Hq, D}, q.options());
  torch::Tensor lse = torch::empty({Hq, T}, q.options().dtype(torch::kFloat32));
  const void *ra = nullptr, *rb = nullptr;
  const short* lp = nullptr;
  int center = 0, NB = 0;
  if (rel_a.has_value()) {
    TORCH_CHECK(rel_b.has_value() && lut.has_value() && rel_a->is_contiguous() && rel_b->is_contiguous());
    TORCH_CHECK(lut->scalar_type() == torch::kInt16 && lut->is_contiguous());
    ra = rel_a->data_ptr(); rb = rel_b->data_ptr();
    lp = reinterpret_cast<const short*>(lut->data_ptr());
    center = (lut->numel() - 1) / 2;
    NB = rel_a->size(2);
    TORCH_CHECK(center >= max_seqlen + 63, "bucket LUT is too short for max_seqlen (needs a 64-entry margin)");
    TORCH_CHECK(NB % 8 == 0);
  }
  check(nrl_attn_varlen_fwd(q.data_ptr(), k.data_ptr(), v.data_ptr(), out.data_ptr(), lse.data_ptr<float>(), q.stride(0),
                            k.stride(0), v.stride(0), out.stride(0), cu_seqlens.data_ptr<int>(), cu_seqlens.numel() - 1, T, Hq,
                            Hkv, D, static_cast<float>(scale), causal ? 1 : 0, ra, rb, lp, center, NB, cur_stream()),
        "attn_varlen_fwd");
  return {out, lse};
}

// DeBERTa disentangled attention (non-causal, D = 64), every tile TMA-fed.  rel_a / rel_b: [H, T, NB] bf16.
torch::Tensor deberta_attn_fwd(const torch::Tensor& q, const torch::Tensor& k, const torch::Tensor& v,
                               const torch::Tensor& cu_seqlens, int64_t max_seqlen, double scale, const torch::Tensor& rel_a,
                               const torch::Tensor& rel_b, const torch::Tensor& lut) {
  check_thd(q, "q"); check_thd(k, "k"); check_thd(v, "v");
  TORCH_CHECK(cu_seqlens.scalar_type() == torch::kInt32 && cu_seqlens.is_contiguous());
  TORCH_CHECK(rel_a.is_contiguous() && rel_b.is_contiguous() && rel_a.scalar_type() == torch::kBFloat16 &&
              rel_b.scalar_type() == torch::kBFloat16 && rel_a.sizes() == rel_b.sizes() && rel_a.dim() == 3);
  TORCH_CHECK(lut.scalar_type() == torch::kInt16 && lut.is_contiguous());
  c10::cuda::CUDAGuard guard(q.device());
  const int T = q.size(0), H = q.size(1), D = q.size(2), NB = rel_a.size(2);
  TORCH_CHECK(D == 64 && k.size(1) == H && rel_a.size(0) == H && rel_a.size(1) == T && NB % 8 == 0);
  const int center = (lut.numel() - 1) / 2;
  TORCH_CHECK(center >= max_seqlen + 63, "bucket LUT is too short for max_seqlen (needs a 64-entry margin)");
  torch::Tensor out = torch::empty({T, H, D}, q.options());
  auto bf = CU_TENSOR_MAP_DATA_TYPE_BFLOAT16;
  static const int bn = (getenv("NRL_DEBERTA_BN") && atoi(getenv("NRL_DEBERTA_BN")) == 32) ? 32 : 64;   // keys per block (64: +9 % at 1660 tokens)
  const int w = bn == 64 ? 136 : 112;                                                                    // window columns
  const CUtensorMap maps[5] = {
      nrl::make_tma_2d(q.data_ptr(), T, static_cast<uint64_t>(H) * D, q.stride(0) * 2, 64, 64, bf, 2),
      nrl::make_tma_2d(k.data_ptr(), T, static_cast<uint64_t>(H) * D, k.stride(0) * 2, bn, 64, bf, 2),
      nrl::make_tma_2d(v.data_ptr(), T, static_cast<uint64_t>(H) * D, v.stride(0) * 2, bn, 64, bf, 2),
      nrl::make_tma_2d_plain(rel_a.data_ptr(), static_cast<uint64_t>(H) * T, NB, static_cast<uint64_t>(NB) * 2, 64, w, bf, 2),
      nrl::make_tma_2d_plain(rel_b.data_ptr(), static_cast<uint64_t>(H) * T, NB, static_cast<uint64_t>(NB) * 2, bn, w, bf, 2)};
  check(nrl_deberta_attn_fwd(maps, out.data_ptr(), nullptr, out.stride(0), cu_seqlens.data_ptr<int>(), cu_seqlens.numel() - 1, T,
                             H, static_cast<float>(scale), reinterpret_cast<const short*>(lut.data_ptr()), center, NB, bn,
                             cur_stream()), "deberta_attn_fwd");
  return out;
}

// y = LayerNorm(x + residual) (residual optional); bf16 [rows, d], d <= 2048
torch::Tensor add_layernorm(const torch::Tensor& x, const c10::optional<torch::Tensor>& residual, const torch::Tensor& w,
                            const torch::Tensor& b, double eps) {
  check_bf16_2d(x, "x");
  TORCH_CHECK(x.is_contiguous() && w.is_contiguous() && b.is_contiguous() && w.scalar_type() == torch::kBFloat16 &&
              b.scalar_type() == torch::kBFloat16 && w.numel() == x.size(1) && b.numel() == x.size(1));
  const void* rp = nullptr;
  if (residual.has_value()) {
    check_bf16_2d(*residual, "residual");
    TORCH_CHECK(residual->is_contiguous() && residual->sizes() == x.sizes());
    rp = residual->data_ptr();
  }
  c10::cuda::CUDAGuard guard(x.device());
  torch::Tensor y = torch::empty_like(x);
  check(nrl_add_layernorm(x.data_ptr(), rp, w.data_ptr(), b.data_ptr(), y.data_ptr(), x.size(0), x.size(1),
                          static_cast<float>(eps), cur_stream()), "add_layernorm");
  return y;
}

// tcgen05 forward: causal, D = 128, bf16.  q/k/v may be strided views (row stride multiple of 8 elements).
std::tuple<torch::Tensor, torch::Tensor> attn_fwd_tc(const torch::Tensor& q, const torch::Tensor& k, const torch::Tensor& v,
                                                     const torch::Tensor& cu_seqlens, double scale,
                                                     const c10::optional<torch::Tensor>& prof) {
  check_thd(q, "q"); check_thd(k, "k"); check_thd(v, "v");
  TORCH_CHECK(cu_seqlens.scalar_type() == torch::kInt32 && cu_seqlens.is_contiguous());
  c10::cuda::CUDAGuard guard(q.device());
  const int T = q.size(0), Hq = q.size(1), D = q.size(2), Hkv = k.size(1);
  TORCH_CHECK(D == 128 && Hq % Hkv == 0, "attn_fwd_tc: head_dim 128 only");
  torch::Tensor out = torch::empty({T, Hq, D}, q.options());
  torch::Tensor lse = torch::empty({Hq, T}, q.options().dtype(torch::kFloat32));
  auto map_of = [&](const torch::Tensor& t, int H, int box_rows) {
    return nrl::make_tma_2d(t.data_ptr(), T, static_cast<uint64_t>(H) * D, t.stride(0) * 2, box_rows, 64,
                            CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2);
  };
  // Q tiles are 128 rows, K/V blocks 64 rows (attention_fwd_tc.cu)
  const CUtensorMap mq = map_of(q, Hq, 128), mk = map_of(k, Hkv, 64), mv = map_of(v, Hkv, 64);
  check(nrl_attn_fwd_tc(&mq, &mk, &mv, out.data_ptr(), lse.data_ptr<float>(), out.stride(0), cu_seqlens.data_ptr<int>(),
                        cu_seqlens.numel() - 1, T, Hq, Hkv, static_cast<float>(scale), cur_stream(),
                        prof.has_value() ? reinterpret_cast<long long*>(prof->data_ptr<int64_t>()) : nullptr), "attn_fwd_tc");
  return {out, lse};
}

std::tuple<torch::Tensor, torch::Tensor, torch::Tensor> attn_varlen_bwd(const torch::Tensor& dout, const torch::Tensor& q,
                                                                       const torch::Tensor& k, const torch::Tensor& v,
                                                                       const torch::Tensor& o, const torch::Tensor& lse,
                                                                       const torch::Tensor& cu_seqlens, int64_t max_seqlen,
                                                                       double scale) {
  check_thd(q, "q"); check_thd(k, "k"); check_thd(v, "v"); check_thd(o, "o"); check_thd(dout, "dout");
  TORCH_CHECK(o.is_contiguous() && dout.is_contiguous() && dout.sizes() == o.sizes());
  c10::cuda::CUDAGuard guard(q.device());
  const int T = q.size(0), Hq = q.size(1), D = q.size(2), Hkv = k.size(1);
  torch::Tensor dq = torch::empty({T, Hq, D}, q.options()), dk = torch::empty({T, Hkv, D}, q.options()),
                dv = torch::empty({T, Hkv, D}, q.options());
  torch::Tensor delta = torch::empty({Hq, T}, q.options().dtype(torch::kFloat32));
  // dq/dk/dv are written with the strides of fresh contiguous tensors; q/k/v may be strided views
  TORCH_CHECK(q.stride(0) == dq.stride(0) && k.stride(0) == dk.stride(0) && v.stride(0) == dv.stride(0),
              "attn_varlen_bwd expects contiguous q/k/v");
  check(nrl_attn_varlen_bwd(dout.data_ptr(), q.data_ptr(), k.data_ptr(), v.data_ptr(), o.data_ptr(), lse.data_ptr<float>(),
                            delta.data_ptr<float>(), dq.data_ptr(), dk.data_ptr(), dv.data_ptr(), q.stride(0), k.stride(0),
                            v.stride(0), o.stride(0), cu_seqlens.data_ptr<int>(), cu_seqlens.numel() - 1, T, Hq, Hkv, D,
                            static_cast<float>(scale), cur_stream()), "attn_varlen_bwd");
  return {dq, dk, dv};
}

// tcgen05 backward: causal, D = 128, bf16, contiguous q/k/v/o/dout.
std::tuple<torch::Tensor, torch::Tensor, torch::Tensor> attn_bwd_tc(const torch::Tensor& dout, const torch::Tensor& q,
                                                                   const torch::Tensor& k, const torch::Tensor& v,
                                                                   const torch::Tensor& o, const torch::Tensor& lse,
                                                                   const torch::Tensor& cu_seqlens, double scale) {
  check_thd(q, "q"); check_thd(k, "k"); check_thd(v, "v"); check_thd(o, "o"); check_thd(dout, "dout");
  TORCH_CHECK(o.is_contiguous() && dout.is_contiguous() && dout.sizes() == o.sizes());
  TORCH_CHECK(lse.scalar_type() == torch::kFloat32 && lse.is_contiguous());
  c10::cuda::CUDAGuard guard(q.device());
  const int T = q.size(0), Hq = q.size(1), D = q.size(2), Hkv = k.size(1);
  TORCH_CHECK(D == 128 && Hq % Hkv == 0, "attn_bwd_tc: head_dim 128 only");
  torch::Tensor dq = torch::empty({T, Hq, D}, q.options()), dk = torch::empty({T, Hkv, D}, q.options()),
                dv = torch::empty({T, Hkv, D}, q.options());
  torch::Tensor delta = torch::empty({Hq, T}, q.options().dtype(torch::kFloat32));
  auto map_of = [&](const torch::Tensor& t, int H, int box_rows) {
    return nrl::make_tma_2d(t.data_ptr(), T, static_cast<uint64_t>(H) * D, t.stride(0) * 2, box_rows, 64,
                            CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2);
  };
  const CUtensorMap maps[8] = {map_of(q, Hq, 64),  map_of(dout, Hq, 64),  map_of(k, Hkv, 128), map_of(v, Hkv, 128),
                               map_of(q, Hq, 128), map_of(dout, Hq, 128), map_of(k, Hkv, 64),  map_of(v, Hkv, 64)};
  check(nrl_attn_bwd_tc(maps, o.data_ptr(), dout.data_ptr(), lse.data_ptr<float>(), delta.data_ptr<float>(), dq.data_ptr(),
                        dk.data_ptr(), dv.data_ptr(), o.stride(0), dq.stride(0), dk.stride(0), cu_seqlens.data_ptr<int>(),
                        cu_seqlens.numel() - 1, T, Hq, Hkv, static_cast<float>(scale), cur_stream()), "attn_bwd_tc");
  return {dq, dk, dv};
}

// ---- K-AR: fused all-reduce + AdamW over symmetric memory ---------------------------------------------
nrl::AdamHyper make_hyper(double lr, double beta1, double beta2, double eps, double wd, int64_t step, double grad_scale) {
  nrl::AdamHyper h;
  h.lr = lr; h.beta1 = beta1; h.beta2 = beta2; h.eps = eps; h.wd = wd;
  h.step_size = static_cast<float>(lr / (1.0 - std::pow(beta1, static_cast<double>(step))));
  h.inv_bc2 = static_cast<float>(1.0 / (1.0 - std::pow(beta2, static_cast<double>(step))));
  h.grad_scale = grad_scale;
  return h;
}

void allreduce_adam(const std::vector<int64_t>& grad_ptrs, const std::vector<int64_t>& param_ptrs, int64_t grad_mc,
                    int64_t param_mc, torch::Tensor m, torch::Tensor v, int64_t lo, int64_t n, int64_t rank, double lr,
                    double beta1, double beta2, double eps, double wd, int64_t step, double grad_scale,
                    bool use_multicast, int64_t max_blocks, c10::optional<torch::Tensor> master) {
  TORCH_CHECK(grad_ptrs.size() == param_ptrs.size() && !grad_ptrs.empty());
  TORCH_CHECK(m.is_cuda() && m.is_contiguous() && v.is_contiguous() && m.numel() == n && v.numel() == n);
  c10::cuda::CUDAGuard guard(m.device());
  std::vector<const void*> g(grad_ptrs.size());
  std::vector<void*> p(param_ptrs.size());
  for (size_t i = 0; i < g.size(); ++i) {
    g[i] = reinterpret_cast<const void*>(grad_ptrs[i]);
    p[i] = reinterpret_cast<void*>(param_ptrs[i]);
  }
  float* mw = nullptr;
  if (master.has_value()) {
    TORCH_CHECK(master->is_cuda() && master->scalar_type() == torch::kFloat32 && master->is_contiguous() && master->numel() == n);
    mw = master->data_ptr<float>();
  }
  check(nrl_allreduce_adam(g.data(), p.data(), reinterpret_cast<const void*>(grad_mc), reinterpret_cast<void*>(param_mc),
                           m.data_ptr(), v.data_ptr(), mw, lo, n, static_cast<int>(g.size()), static_cast<int>(rank),
                           m.scalar_type() == torch::kBFloat16 ? 1 : 0, use_multicast ? 1 : 0,
                           make_hyper(lr, beta1, beta2, eps, wd, step, grad_scale), static_cast<int>(max_blocks), cur_stream()),
        "allreduce_adam");
}

void allreduce_sum(const std::vector<int64_t>& buf_ptrs, int64_t lo, int64_t n, int64_t rank, double scale, int64_t max_blocks) {
  std::vector<void*> p(buf_ptrs.size());
  for (size_t i = 0; i < p.size(); ++i) p[i] = reinterpret_cast<void*>(buf_ptrs[i]);
  check(nrl_allreduce_sum(p.data(), lo, n, static_cast<int>(p.size()), static_cast<int>(rank), static_cast<float>(scale),
                          static_cast<int>(max_blocks), cur_stream()), "allreduce_sum");
}

void kv_cache_write_fp8(const torch::Tensor& k, const torch::Tensor& v, torch::Tensor kq, torch::Tensor vq, torch::Tensor ks,
                        torch::Tensor vs, const torch::Tensor& slot_mapping, const c10::optional<torch::Tensor>& src_index) {
  TORCH_CHECK(k.dim() == 3 && v.dim() == 3 && k.stride(2) == 1 && k.stride(1) == k.size(2) && v.stride(2) == 1 && v.stride(1) == v.size(2));
  TORCH_CHECK(kq.scalar_type() == torch::kUInt8 && kq.dim() == 4 && kq.is_contiguous() && vq.is_contiguous());
  TORCH_CHECK(ks.scalar_type() == torch::kFloat32 && ks.is_contiguous() && vs.is_contiguous());
  TORCH_CHECK(slot_mapping.scalar_type() == torch::kInt32 && slot_mapping.is_contiguous());
  const int* src = nullptr;
  if (src_index.has_value()) {
    TORCH_CHECK(src_index->scalar_type() == torch::kInt32 && src_index->numel() == slot_mapping.numel());
    src = src_index->data_ptr<int>();
  } else {
    TORCH_CHECK(slot_mapping.numel() == k.size(0));
  }
  c10::cuda::CUDAGuard guard(k.device());
  check(nrl_kv_cache_write_fp8(k.data_ptr(), v.data_ptr(), k.stride(0), v.stride(0), kq.data_ptr(), vq.data_ptr(), ks.data_ptr<float>(),
                               vs.data_ptr<float>(), slot_mapping.data_ptr<int>(), src, slot_mapping.numel(), k.size(1), k.size(2),
                               kq.size(2), cur_stream()), "kv_cache_write_fp8");
}

// decode-step fusion: RoPE on the q and k heads of qkv [S, (Hq + 2 Hkv) * 128] in place + page write of (rotated k, v)
void rope_kv_write(torch::Tensor qkv, const torch::Tensor& cos_t, const torch::Tensor& sin_t, torch::Tensor k_cache, torch::Tensor v_cache,
                   const c10::optional<torch::Tensor>& k_scale, const c10::optional<torch::Tensor>& v_scale, const torch::Tensor& slot_mapping,
                   int64_t Hq, int64_t Hkv) {
  TORCH_CHECK(qkv.is_cuda() && qkv.scalar_type() == torch::kBFloat16 && qkv.dim() == 2 && qkv.stride(1) == 1 && qkv.size(1) == (Hq + 2 * Hkv) * 128);
  TORCH_CHECK(cos_t.scalar_type() == torch::kFloat32 && cos_t.is_contiguous() && sin_t.is_contiguous() && cos_t.size(0) == qkv.size(0) && cos_t.size(1) == 64);
  TORCH_CHECK(slot_mapping.scalar_type() == torch::kInt32 && slot_mapping.numel() == qkv.size(0) && slot_mapping.is_contiguous());
  const bool kv8 = k_cache.scalar_type() == torch::kUInt8;
  TORCH_CHECK(k_cache.is_contiguous() && v_cache.is_contiguous() && k_cache.size(1) == Hkv);
  TORCH_CHECK(!kv8 || (k_scale.has_value() && v_scale.has_value()), "fp8 pages need their scale tensors");
  c10::cuda::CUDAGuard guard(qkv.device());
  check(nrl_rope_kv_write(qkv.data_ptr(), qkv.stride(0), cos_t.data_ptr<float>(), sin_t.data_ptr<float>(), k_cache.data_ptr(), v_cache.data_ptr(),
                          kv8 ? k_scale->data_ptr<float>() : nullptr, kv8 ? v_scale->data_ptr<float>() : nullptr, slot_mapping.data_ptr<int>(),
                          static_cast<int>(qkv.size(0)), static_cast<int>(Hq), static_cast<int>(Hkv), 128, 16, kv8 ? 1 : 0, cur_stream()),
        "rope_kv_write");
}

torch::Tensor paged_decode_fp8(const torch::Tensor& q, const torch::Tensor& kq, const torch::Tensor& vq, const torch::Tensor& ks,
                               const torch::Tensor& vs, const torch::Tensor& block_tables, const torch::Tensor& context_lens,
                               double scale, int64_t splits) {
  TORCH_CHECK(q.is_cuda() && q.scalar_type() == torch::kBFloat16 && q.dim() == 3 && q.stride(2) == 1 && q.stride(1) == q.size(2));
  TORCH_CHECK(block_tables.scalar_type() == torch::kInt32 && block_tables.is_contiguous() && context_lens.scalar_type() == torch::kInt32);
  c10::cuda::CUDAGuard guard(q.device());
  const int S = q.size(0), Hq = q.size(1), D = q.size(2), Hkv = kq.size(1);
  torch::Tensor out = torch::empty({S, Hq, D}, q.options());
  torch::Tensor po, pml;
  float *pop = nullptr, *pmlp = nullptr;
  if (splits > 1) {
    po = torch::empty({S, Hkv, splits, 8, D}, q.options().dtype(torch::kFloat32));
    pml = torch::empty({S, Hkv, splits, 8, 2}, q.options().dtype(torch::kFloat32));
    pop = po.data_ptr<float>();
    pmlp = pml.data_ptr<float>();
  }
  check(nrl_paged_decode_fp8(q.data_ptr(), q.stride(0), kq.data_ptr(), vq.data_ptr(), ks.data_ptr<float>(), vs.data_ptr<float>(),
                             block_tables.data_ptr<int>(), context_lens.data_ptr<int>(), out.data_ptr(), pop, pmlp, S, Hq, Hkv, D,
                             kq.size(2), block_tables.size(1), static_cast<int>(splits), static_cast<float>(scale), cur_stream()),
        "paged_decode_fp8");
  return out;
}

}  // namespace

PYBIND11_MODULE(TORCH_EXTENSION_NAME, m) {
  m.doc() = "nanorlhf_b200 sm_100a kernels";
  m.def("gemm_bf16", &gemm_bf16, py::arg("a"), py::arg("b"), py::arg("bias") = py::none(), py::arg("out") = py::none(),
        py::arg("block_n") = 0, py::arg("act") = 0);
  m.def("gemm_tc", &gemm_tc, py::arg("a"), py::arg("b"), py::arg("a_mn") = false, py::arg("b_mn") = false, py::arg("a2") = py::none(),
        py::arg("b2") = py::none(), py::arg("bias") = py::none(), py::arg("act") = 0, py::arg("alpha") = 1.0, py::arg("out") = py::none(),
        py::arg("out_f32") = py::none(), py::arg("accumulate") = false, py::arg("cg") = 0, py::arg("block_n") = 0,
        py::arg("split_k") = -1);      // -1 = automatic, 0 = never, n = exactly n k-ranges
  m.def("gemm_tc_fp8", &gemm_tc_fp8, py::arg("aq"), py::arg("a_scale"), py::arg("bq"), py::arg("b_scale"), py::arg("bias") = py::none(),
        py::arg("swiglu") = false, py::arg("cg") = 0, py::arg("block_n") = 0);
  m.def("gemm_tc_batched", &gemm_tc_batched, py::arg("a"), py::arg("b"), py::arg("out") = py::none(), py::arg("cg") = 0, py::arg("block_n") = 0);
  m.def("gemm_tc_swiglu", &gemm_tc_swiglu, py::arg("a"), py::arg("w_interleaved"), py::arg("out") = py::none(), py::arg("cg") = 0,
        py::arg("block_n") = 0);
  m.def("add_layernorm", &add_layernorm);
  m.def("lmhead_logprob_fwd", &lmhead_logprob_fwd, py::arg("hidden"), py::arg("weight"), py::arg("targets"),
        py::arg("inv_temperature"), py::arg("n_splits") = 0);
  m.def("lmhead_dlogits", &lmhead_dlogits);
  m.def("lora_merge", &lora_merge, py::arg("w"), py::arg("lora_a"), py::arg("lora_b"), py::arg("scale"), py::arg("out"),
        py::arg("mc_out_ptr") = 0);
  m.def("quant_rows_e4m3", &quant_rows_e4m3);
  m.def("gemm_fp8", &gemm_fp8, py::arg("aq"), py::arg("a_scale"), py::arg("bq"), py::arg("b_scale"), py::arg("bias") = py::none(),
        py::arg("swiglu") = false);
  m.def("gemm_swiglu", &gemm_swiglu, py::arg("a"), py::arg("w_interleaved"), py::arg("out") = py::none());
  m.def("rmsnorm", &rmsnorm, py::arg("x"), py::arg("w"), py::arg("eps"), py::arg("residual") = py::none(),
        py::arg("want_rstd") = false);
  m.def("rmsnorm_bwd", &rmsnorm_bwd);
  m.def("rope", &rope, py::arg("x"), py::arg("cos"), py::arg("sin"), py::arg("sin_sign") = 1.0, py::arg("inplace") = false);
  m.def("swiglu", &swiglu);
  m.def("swiglu_bwd", &swiglu_bwd);
  m.def("swiglu_pair", &swiglu_pair);
  m.def("swiglu_pair_bwd", &swiglu_pair_bwd);
  m.def("gae_scan", &gae_scan, py::arg("rewards"), py::arg("values") = py::none(), py::arg("gamma") = 1.0, py::arg("lam") = 1.0);
  m.def("policy_loss", &policy_loss);
  m.def("value_loss", &value_loss);
  m.def("adamw_flat", &adamw_flat, py::arg("param"), py::arg("grad"), py::arg("m"), py::arg("v"), py::arg("lr"), py::arg("beta1"),
        py::arg("beta2"), py::arg("eps"), py::arg("wd"), py::arg("step"), py::arg("grad_scale") = 1.0, py::arg("master") = py::none());
  m.def("sample", &sample, py::arg("logits"), py::arg("temperature"), py::arg("top_p"), py::arg("seed"), py::arg("step"),
        py::arg("row_ids") = py::none(), py::arg("row_steps") = py::none(), py::arg("out") = py::none(),
        py::arg("impl") = 0);           // 0 = automatic, 1 = streaming kernel, 2 = shared-memory-resident cluster kernel
  m.def("kv_cache_write", &kv_cache_write, py::arg("k"), py::arg("v"), py::arg("k_cache"), py::arg("v_cache"),
        py::arg("slot_mapping"), py::arg("src_index") = py::none());
  m.def("paged_decode", &paged_decode, py::arg("q"), py::arg("k_cache"), py::arg("v_cache"), py::arg("block_tables"),
        py::arg("context_lens"), py::arg("scale"), py::arg("splits") = 1, py::arg("out") = py::none());
  m.def("attn_varlen_fwd", &attn_varlen_fwd, py::arg("q"), py::arg("k"), py::arg("v"), py::arg("cu_seqlens"), py::arg("max_seqlen"),
        py::arg("scale"), py::arg("causal") = true, py::arg("rel_a") = py::none(), py::arg("rel_b") = py::none(),
        py::arg("lut") = py::none());
  m.def("attn_varlen_bwd", &attn_varlen_bwd);
  m.def("attn_bwd_tc", &attn_bwd_tc);
  m.def("deberta_attn_fwd", &deberta_attn_fwd);
  m.def("attn_fwd_tc", &attn_fwd_tc, py::arg("q"), py::arg("k"), py::arg("v"), py::arg("cu_seqlens"), py::arg("scale"),
        py::arg("prof") = py::none());
  m.def("kv_cache_write_fp8", &kv_cache_write_fp8, py::arg("k"), py::arg("v"), py::arg("kq"), py::arg("vq"), py::arg("ks"), py::arg("vs"),
        py::arg("slot_mapping"), py::arg("src_index") = py::none());
  m.def("paged_decode_fp8", &paged_decode_fp8);
  m.def("rope_kv_write", &rope_kv_write);
  m.def("allreduce_adam", &allreduce_adam, py::arg("grad_ptrs"), py::arg("param_ptrs"), py::arg("grad_mc"), py::arg("param_mc"),
        py::arg("m"), py::arg("v"), py::arg("lo"), py::arg("n"), py::arg("rank"), py::arg("lr"), py::arg("beta1"), py::arg("beta2"),
        py::arg("eps"), py::arg("wd"), py::arg("step"), py::arg("grad_scale"), py::arg("use_multicast"), py::arg("max_blocks"),
        py::arg("master") = py::none());
  m.def("allreduce_sum", &allreduce_sum);
  nrl::bind_runtime(m);
}
