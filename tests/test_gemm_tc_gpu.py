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
    got = _ext().gemm_tc(a, b, False, False, None, None, bias, 0, 1.0, None, None, False, cg, bn)
    _check(got, want + bias.float(), k=K)


@pytest.mark.parametrize("cg,bn", [(1, 64), (1, 128), (1, 192), (1, 256), (2, 128), (2, 256)])
@pytest.mark.parametrize("M,N,K", SHAPES)
def test_b_mn_dgrad_form(cg, bn, M, N, K):
    a, b, want = _operands(M, N, K, False, True)
    got = _ext().gemm_tc(a, b, False, True, None, None, None, 0, 0.5, None, None, False, cg, bn)
    _check(got, 0.5 * want, k=K)


@pytest.mark.parametrize("cg,bn", [(1, 64), (1, 128), (1, 256), (2, 256)])
@pytest.mark.parametrize("M,N,K", [(128, 256, 64), (1536, 64, 6912), (64, 1536, 3000), (1024, 264, 520)])
def test_both_mn_wgrad_form(cg, bn, M, N, K):
    a, b, want = _operands(M, N, K, True, True)
    got = _ext().gemm_tc(a, b, True, True, None, None, None, 0, 1.0, None, None, False, cg, bn)
    _check(got, want, k=K)
    acc = torch.randn(M, N, device="cuda")
    base = acc.clone()
    _ext().gemm_tc(a, b, True, True, None, None, None, 0, 2.0, None, acc, True, cg, bn)
    _check(acc, base + 2.0 * want, k=K)
    _ext().gemm_tc(a, b, True, True, None, None, None, 0, 1.0, None, acc, False, cg, bn)
    _check(acc, want, k=K)


@pytest.mark.parametrize("b_mn", [False, True])
@pytest.mark.parametrize("cg,bn", [(1, 128), (1, 256), (2, 128), (2, 256)])
def test_dual_source_k_lora(cg, bn, b_mn):
    """y = x W^T + t B^T as ONE GEMM (t = s x A^T is the second A operand): the LoRA forward / dgrad form."""
    M, N, K, r = 1000, 1536, 1024, 64
    a, b, want = _operands(M, N, K, False, b_mn, seed=1)
    a2, b2, want2 = _operands(M, N, r, False, b_mn, seed=2)
    got = _ext().gemm_tc(a, b, False, b_mn, a2, b2, None, 0, 1.0, None, None, False, cg, bn)
    _check(got, want + want2, k=K + r)
    # the second operand pair may be a column slice of a wider buffer (q / k / v adapters share one projection)
    wide = torch.randn(M, 3 * r, device="cuda").bfloat16()
    if not b_mn:
        got = _ext().gemm_tc(a, b, False, False, wide[:, r:2 * r], b2, None, 0, 1.0, None, None, False, cg, bn)
        _check(got, want + wide[:, r:2 * r].float() @ b2.float().t(), k=K + r)


def test_auto_dispatch_and_gelu():
    a, b, want = _operands(777, 4096, 1024, False, False)
    bias = torch.randn(4096, device="cuda").bfloat16()
    got = _ext().gemm_tc(a, b, False, False, None, None, bias, 1)
    _check(got, torch.nn.functional.gelu(want + bias.float()), k=1024)
