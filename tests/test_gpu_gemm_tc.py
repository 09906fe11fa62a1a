"""``gode_gemm_tc_f32`` -- the general tcgen05 3xTF32 GEMM (csrc/transform_tc.cu, k_gemm_tc) that carries the QC edge
encoder, the GAT projections and the input layers -- against fp64 on shapes that exercise every edge: partial tiles in M, N
and K, unaligned leading dimensions, K long enough for several accumulation groups, padding that must not leak."""
import ctypes as C

import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("M,N,K,pad", [(128, 128, 32, True), (128, 128, 128, True), (300, 200, 70, True), (1000, 533, 267, True),
                                       (257, 129, 33, False), (4096, 256, 1024, True), (5, 3, 1, False), (700, 384, 2667, True)])
@pytest.mark.parametrize("relu,use_bias", [(0, False), (1, True)])
def test_gemm_tc_matches_fp64(M, N, K, pad, relu, use_bias):
    import graph_odenet_b200  # noqa: F401
    from graph_odenet_b200 import ops
    from graph_odenet_b200._lib import lib, check
    dev = torch.device("cuda:0")
    g = torch.Generator(device=dev).manual_seed(M * 7 + N * 3 + K)
    ld = lambda n: (n + 3) // 4 * 4 if pad else n
    A = torch.zeros(M, ld(K), device=dev)
    Bt = torch.zeros(N, ld(K), device=dev)
    A[:, :K] = torch.randn(M, K, device=dev, generator=g)
    Bt[:, :K] = torch.randn(N, K, device=dev, generator=g)
    if pad:                                    # padding must not leak into the product
        A[:, K:] = 1e6
        Bt[:, K:] = -1e6
    bias = torch.randn(N, device=dev, generator=g) if use_bias else None
    Cm = torch.full((M, ld(N)), float("nan"), device=dev)
    check(lib.gode_gemm_tc_f32(M, N, K, ops._p(A), A.stride(0), ops._p(Bt), Bt.stride(0), ops._p(bias), relu, ops._p(Cm),
                               Cm.stride(0), 0, ops._stream()), "gode_gemm_tc_f32")
    torch.cuda.synchronize()
    want = A[:, :K].double() @ Bt[:, :K].double().t()
    if use_bias:
        want = want + bias.double()
    if relu:
        want = want.clamp_min(0)
    got = Cm[:, :N].double()
    scale = float(want.abs().max())
    assert torch.isfinite(got).all()
    # an fp32 SGEMM's own error grows like sqrt(K) * 6e-8; the bar is 2e-6 of the result's scale for every K tested
    err = float((got - want).abs().max())
    lib_err = float(((A[:, :K] @ Bt[:, :K].t() + (bias if use_bias else 0)).clamp_min(0 if relu else -float("inf")).double() - want).abs().max())
    print("gemm_tc M=%d N=%d K=%d: max err / scale %.2e (torch fp32 matmul: %.2e)" % (M, N, K, err / scale, lib_err / scale))
    assert err <= 2e-6 * scale, (err, scale)
    if ld(N) > N:
        assert torch.isnan(Cm[:, N:]).all()                      # columns beyond N are never written
