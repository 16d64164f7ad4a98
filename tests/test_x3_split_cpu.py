"""CPU checks of the arithmetic behind fp32 mode on the tensor cores: the exact
split of fp32 values into bf16 planes done by prep_kernel<PL=3>, and the error of the plane-pair
segment chains the tcgen05 kernels accumulate (6 pairs for S = X W, 5 for dW / dX with a
two-plane G'').  The GPU kernels are tested against the oracle in test_head_gpu.py; this file
pins the bounds those tests rely on without a GPU."""
import numpy as np
import torch


def planes(x: torch.Tensor, n: int):
    """p0 = bf16(x), p1 = bf16(x - p0), ... (round to nearest even), as the kernels do it."""
    out, r = [], x.clone()
    for _ in range(n):
        p = r.to(torch.bfloat16).to(torch.float32)
        out.append(p)
        r = r - p
    return out


def test_three_bf16_planes_reconstruct_fp32_exactly():
    g = torch.Generator().manual_seed(0)
    x = torch.randn(1 << 16, generator=g) * torch.exp2(torch.randint(-40, 40, (1 << 16,), generator=g).float())
    p0, p1, p2 = planes(x, 3)
    assert torch.equal((p0 + p1) + p2, x)                     # 3 x 8 significand bits = fp32's 24
    assert torch.all(p1.abs() <= p0.abs() * 2.0 ** -8 + 1e-45)
    assert torch.all(p2.abs() <= p0.abs() * 2.0 ** -16 + 1e-45)
    # zeros and exact bf16 values have empty low planes
    z = torch.tensor([0.0, 1.0, -2.5, float(2 ** 127)])
    q0, q1, q2 = planes(z, 3)
    assert torch.equal(q0, z) and not q1.any() and not q2.any()


def test_six_segment_chain_matches_fp32_product_to_rounding_level():
    g = torch.Generator().manual_seed(1)
    X = torch.randn(64, 512, generator=g) * 3.0
    W = torch.randn(512, 96, generator=g) * 0.05
    xp, wp = planes(X, 3), planes(W, 3)
    exact = X.double() @ W.double()
    scale = (X.abs().double() @ W.abs().double())             # sum_k |x||w|: the natural error scale
    pairs6 = [(0, 0), (0, 1), (1, 0), (1, 1), (0, 2), (2, 0)]  # p + q <= 2 (asm_umma_gemm.cu kSegHi/kSegLo)
    chain = sum(xp[a].double() @ wp[b].double() for a, b in pairs6)
    assert float(((chain - exact).abs() / scale).max()) < 2.0 ** -22
    pairs3 = pairs6[:3]                                        # ASM_X3_SEGS=3
    chain3 = sum(xp[a].double() @ wp[b].double() for a, b in pairs3)
    assert float(((chain3 - exact).abs() / scale).max()) < 2.0 ** -14
    # plain bf16 mode for comparison: one plane each
    bf = xp[0].double() @ wp[0].double()
    assert float(((bf - exact).abs() / scale).max()) > 2.0 ** -12


def test_two_plane_g_against_three_plane_operand():
    g = torch.Generator().manual_seed(2)
    G = torch.randn(128, 80, generator=g) * 1e-3              # G'' [B, C]
    X = torch.randn(128, 64, generator=g) * 3.0               # X   [B, D]
    gp, xp = planes(G, 2), planes(X, 3)
    exact = G.double().T @ X.double()                         # dW^T [C, D]
    scale = G.abs().double().T @ X.abs().double()
    pairs5 = [(0, 0), (0, 1), (1, 0), (1, 1), (0, 2)]         # kSegG / kSegO
    chain = sum(gp[a].double().T @ xp[b].double() for a, b in pairs5)
    err = float(((chain - exact).abs() / scale).max())
    assert err < 2.0 ** -15, err                              # two G'' planes: ~2^-17 per product
