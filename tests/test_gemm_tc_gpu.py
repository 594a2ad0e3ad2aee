"""T2: the general tcgen05 GEMM (csrc/gemm_tc.cu) against a plain fp32 PyTorch product -- every operand-major
combination, CTA-group size, tile width, the dual-source-K (LoRA) form and the fp32-accumulate epilogue."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def _ext():
    from nanorlhf_b200.ops import native
    return native.ext()


def _check(got, want, rel=1e-2, k=1):
    got, want = got.float(), want.float()
    r = ((got - want).norm() / want.norm().clamp_min(1e-12)).item()
    assert r < rel, f"relative Frobenius error {r}"
    # a few wrong rows / columns hide in a Frobenius norm: bound the worst element by bf16 rounding of the result plus
    # fp32-accumulation noise over the contraction
    tol = 2.0 ** -7 * want.abs().max().item() + 1e-3 * k ** 0.5
    worst = (got - want).abs().max().item()
    assert worst < tol, f"max abs error {worst} (tol {tol})"


def _operands(M, N, K, a_mn, b_mn, seed=0):
    g = torch.Generator(device="cuda").manual_seed(seed)
    a = torch.randn((K, M) if a_mn else (M, K), device="cuda", generator=g).bfloat16()
    b = torch.randn((K, N) if b_mn else (N, K), device="cuda", generator=g).bfloat16()
    af = a.float().t() if a_mn else a.float()
    bf = b.float().t() if b_mn else b.float()
    return a, b, af @ bf.t()


SHAPES = [(128, 256, 64), (300, 520, 200), (2048, 1536, 1536), (1000, 64, 512), (6912, 1536, 8960 // 4)]


@pytest.mark.parametrize("cg,bn", [(1, 64), (1, 128), (1, 192), (1, 256), (2, 128), (2, 192), (2, 256)])
@pytest.mark.parametrize("M,N,K", SHAPES)
def test_tn(cg, bn, M, N, K):
    a, b, want = _operands(M, N, K, False, False)
    bias = torch.randn(N, device="cuda").bfloat16()
    got = _ext().gemm_tc(a, b, False, False, None, None, bias, 0, 1.0, None, None, False, cg, bn, 0)
    _check(got, want + bias.float(), k=K)


@pytest.mark.parametrize("cg,bn", [(1, 64), (1, 128), (1, 192), (1, 256), (2, 128), (2, 192), (2, 256)])
@pytest.mark.parametrize("M,N,K", SHAPES)
def test_b_mn_dgrad_form(cg, bn, M, N, K):
    a, b, want = _operands(M, N, K, False, True)
    got = _ext().gemm_tc(a, b, False, True, None, None, None, 0, 0.5, None, None, False, cg, bn, 0)
    _check(got, 0.5 * want, k=K)


@pytest.mark.parametrize("cg,bn", [(1, 64), (1, 128), (1, 256), (2, 256)])
@pytest.mark.parametrize("M,N,K", [(128, 256, 64), (1536, 64, 6912), (64, 1536, 3000), (1024, 264, 520)])
def test_both_mn_wgrad_form(cg, bn, M, N, K):
    a, b, want = _operands(M, N, K, True, True)
    got = _ext().gemm_tc(a, b, True, True, None, None, None, 0, 1.0, None, None, False, cg, bn, 0)
    _check(got, want, k=K)
    acc = torch.randn(M, N, device="cuda")
    base = acc.clone()
    _ext().gemm_tc(a, b, True, True, None, None, None, 0, 2.0, None, acc, True, cg, bn)
    _check(acc, base + 2.0 * want, k=K)
    _ext().gemm_tc(a, b, True, True, None, None, None, 0, 1.0, None, acc, False, cg, bn, 0)
    _check(acc, want, k=K)


@pytest.mark.parametrize("b_mn", [False, True])
@pytest.mark.parametrize("cg,bn", [(1, 128), (1, 256), (2, 128), (2, 192), (2, 256)])
def test_dual_source_k_lora(cg, bn, b_mn):
    """y = x W^T + t B^T as ONE GEMM (t = s x A^T is the second A operand): the LoRA forward / dgrad form."""
    M, N, K, r = 1000, 1536, 1024, 64
    a, b, want = _operands(M, N, K, False, b_mn, seed=1)
    a2, b2, want2 = _operands(M, N, r, False, b_mn, seed=2)
    got = _ext().gemm_tc(a, b, False, b_mn, a2, b2, None, 0, 1.0, None, None, False, cg, bn, 0)
    _check(got, want + want2, k=K + r)
    # the second operand pair may be a column slice of a wider buffer (q / k / v adapters share one projection)
    wide = torch.randn(M, 3 * r, device="cuda").bfloat16()
    if not b_mn:
        got = _ext().gemm_tc(a, b, False, False, wide[:, r:2 * r], b2, None, 0, 1.0, None, None, False, cg, bn, 0)
        _check(got, want + wide[:, r:2 * r].float() @ b2.float().t(), k=K + r)


def test_auto_dispatch_and_gelu():
    a, b, want = _operands(777, 4096, 1024, False, False)
    bias = torch.randn(4096, device="cuda").bfloat16()
    got = _ext().gemm_tc(a, b, False, False, None, None, bias, 1)
    _check(got, torch.nn.functional.gelu(want + bias.float()), k=1024)


@pytest.mark.parametrize("form,M,N,K", [("NT", 64, 1536, 6912), ("NT", 1536, 64, 6912), ("NT", 8960, 64, 3000), ("NT", 128, 1536, 6912),
                                        ("TN", 6912, 64, 1536), ("NN", 6912, 64, 8960), ("NN", 1000, 64, 1536), ("TN", 300, 128, 4096)])
@pytest.mark.parametrize("split_k", [-1, 1, 3, 16])
def test_split_k(form, M, N, K, split_k):
    """Small outputs with long contractions (LoRA wgrads, rank-r projections): split-K with the in-kernel ordered fix-up."""
    a_mn, b_mn = form == "NT", form in ("NN", "NT")
    a, b, want = _operands(M, N, K, a_mn, b_mn, seed=3)
    ext = _ext()
    for _ in range(2):                                # the tile counters must reset themselves between launches
        got = ext.gemm_tc(a, b, a_mn, b_mn, None, None, None, 0, 0.25, None, None, False, 0, 0, split_k)
        _check(got, 0.25 * want, k=K)
    # split order is fixed -> bitwise reproducible
    again = ext.gemm_tc(a, b, a_mn, b_mn, None, None, None, 0, 0.25, None, None, False, 0, 0, split_k)
    assert torch.equal(got, again)
    # bias rides in the fix-up (small-batch decode projections: few tiles, the weight streamed by all SMs)
    bias = torch.randn(N, device="cuda").bfloat16()
    got = ext.gemm_tc(a, b, a_mn, b_mn, None, None, bias, 0, 0.25, None, None, False, 0, 0, split_k)
    _check(got, 0.25 * want + bias.float(), k=K)


def test_small_batch_decode_shapes_take_split_k():
    """M = 64 rows against 7B-sized weights (4608 / 3584 outputs): automatic dispatch == explicit single-pass kernel."""
    ext = _ext()
    for (M, N, K, with_bias) in ((64, 4608, 3584, True), (64, 3584, 18944, False), (256, 3584, 3584, False), (40, 1536, 8960, False)):
        a, b, want = _operands(M, N, K, False, False, seed=5)
        bias = torch.randn(N, device="cuda").bfloat16() if with_bias else None
        auto = ext.gemm_tc(a, b, False, False, None, None, bias)
        one = ext.gemm_tc(a, b, False, False, None, None, bias, 0, 1.0, None, None, False, 0, 0, 0)
        _check(auto, want + (bias.float() if with_bias else 0), k=K)
        _check(one, want + (bias.float() if with_bias else 0), k=K)


@pytest.mark.parametrize("cg,bn", [(1, 128), (1, 256), (2, 128), (2, 256), (0, 0)])
def test_swiglu_epilogue(cg, bn):
    """gate_up projection with silu(gate) * up formed on the accumulator tile (interleaved weight rows)."""
    torch.manual_seed(0)
    for (M, F, K) in ((300, 512, 256), (1024, 8960, 1536), (77, 128, 128)):
        x = torch.randn(M, K, device="cuda").bfloat16()
        wg = (torch.randn(F, K, device="cuda") * 0.05).bfloat16()
        wu = (torch.randn(F, K, device="cuda") * 0.05).bfloat16()
        wi = torch.empty(2 * F, K, device="cuda", dtype=torch.bfloat16)
        wi.view(F // 32, 2, 32, K).copy_(torch.stack([wg.view(F // 32, 32, K), wu.view(F // 32, 32, K)], 1))
        out = _ext().gemm_tc_swiglu(x, wi, None, cg, bn)
        g, u = x.float() @ wg.float().t(), x.float() @ wu.float().t()
        _check(out, torch.nn.functional.silu(g) * u, rel=2e-2, k=K)


@pytest.mark.parametrize("cg,bn", [(1, 128), (1, 256), (2, 128), (2, 256), (0, 0)])
def test_batched_strided_heads(cg, bn):
    """out[h] = x[:, h] pos[h]^T with x a packed [T, H, dh] projection (the DeBERTa bias tables): 3D TMA, no copies."""
    torch.manual_seed(0)
    for (T, H, dh, NB) in ((1000, 16, 64, 512), (300, 4, 64, 136), (26560 // 8, 16, 64, 512)):
        x = torch.randn(T, H, dh, device="cuda").bfloat16()
        pos = torch.randn(NB, H, dh, device="cuda").bfloat16().transpose(0, 1)          # [H, NB, dh] view
        out = _ext().gemm_tc_batched(x.transpose(0, 1), pos, None, cg, bn)
        want = torch.bmm(x.float().transpose(0, 1), pos.float().transpose(1, 2))
        _check(out, want, k=dh)


@pytest.mark.parametrize("cg,bn", [(1, 128), (1, 256), (2, 128), (2, 192), (2, 256), (0, 0)])
def test_fp8_e4m3_gemm(cg, bn):
    """kind::f8f6f4 GEMM with per-token x per-channel scales (the fp8 rollout path), plain and SwiGLU epilogues."""
    ext = _ext()
    torch.manual_seed(0)
    for (M, N, K, sw) in ((300, 512, 256, False), (1024, 2048, 1536, False), (512, 1536, 8960, False), (256, 1024, 512, True), (1024, 17920, 1536, True)):
        if sw and bn == 192:
            continue
        x = torch.randn(M, K, device="cuda").bfloat16()
        w = (torch.randn(N, K, device="cuda") * 0.05).bfloat16()
        xq, xs = ext.quant_rows_e4m3(x)
        wq, ws = ext.quant_rows_e4m3(w)
        xd = xq.view(torch.float8_e4m3fn).float() * xs[:, None]
        wd = wq.view(torch.float8_e4m3fn).float() * ws[:, None]
        bias = torch.randn(N, device="cuda").bfloat16() if not sw else None
        out = ext.gemm_tc_fp8(xq, xs, wq, ws, bias, sw, cg, bn)
        z = xd @ wd.t()
        if sw:
            F = N // 2
            zz = z.view(M, F // 32, 2, 32)
            want = (torch.nn.functional.silu(zz[:, :, 0]) * zz[:, :, 1]).reshape(M, F)
        else:
            want = z + bias.float()
        _check(out, want, rel=1e-2, k=K)


@pytest.mark.parametrize("n_proj", [1, 2, 3])
def test_lora_linear_autograd_single_and_grouped(n_proj):
    """ops.lora_linear / ops.lora_linear_group (dual-source-K forward, MN-major backward, shared rank-r launches) against
    fp32 autograd through the plain formula y = x W^T + b + s (x A^T) B^T."""
    from nanorlhf_b200 import ops
    torch.manual_seed(0)
    T, K, r, s = 1000, 1536, 64, 0.5
    Ns = [1536, 256, 256][:n_proj]
    x = (torch.randn(T, K, device="cuda") * 0.5).bfloat16().requires_grad_()
    ws = [(torch.randn(N, K, device="cuda") * 0.03).bfloat16() for N in Ns]
    bs = [torch.randn(N, device="cuda").bfloat16() if i != 1 else None for i, N in enumerate(Ns)]
    As = [(torch.randn(r, K, device="cuda") * 0.05).bfloat16().requires_grad_() for _ in Ns]
    Bs = [(torch.randn(N, r, device="cuda") * 0.05).bfloat16().requires_grad_() for N in Ns]
    gs = [torch.randn(T, N, device="cuda").bfloat16() for N in Ns]
    if n_proj == 1:
        ys = [ops.lora_linear(x, ws[0], bs[0], As[0], Bs[0], s)]
    else:
        ys = ops.lora_linear_group(x, [(ws[i], bs[i], As[i], Bs[i]) for i in range(n_proj)], s)
    torch.autograd.backward(ys, gs)
    got = [x.grad] + [a.grad for a in As] + [b.grad for b in Bs]
    xf = x.detach().float().requires_grad_()
    Af = [a.detach().float().requires_grad_() for a in As]
    Bf = [b.detach().float().requires_grad_() for b in Bs]
    yf = [xf @ ws[i].float().t() + (bs[i].float() if bs[i] is not None else 0) + s * (xf @ Af[i].t()) @ Bf[i].t() for i in range(n_proj)]
    torch.autograd.backward(yf, [g.float() for g in gs])
    want = [xf.grad] + [a.grad for a in Af] + [b.grad for b in Bf]
    for y, w in zip(ys, yf):
        _check(y, w.detach(), k=K)
    for g, w in zip(got, want):
        r_ = ((g.float() - w).norm() / w.norm()).item()
        assert r_ < 1.5e-2, r_
