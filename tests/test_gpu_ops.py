"""-m gpu: the canonical CUDA building blocks against the CPU oracle, bit for bit (through the C ABI)."""
import pytest
import torch

gpu = pytest.mark.gpu


def _dev():
    return torch.device("cuda:0")


@gpu
@pytest.mark.parametrize("M,N,K,epi", [(1, 1024, 1024, 0), (3, 100, 256, 0), (9, 515, 768, 1), (64, 1024, 4096, 0),
                                        (257, 200, 1024, 1), (5, 8195, 1024, 0), (8, 64, 1536, 0)])
def test_linear_bit_exact(M, N, K, epi):
    from meshanything_b200 import capi
    from oracle import decoder as orc
    g = torch.Generator().manual_seed(M * 1000 + N)
    w = (torch.randn(N, K, generator=g) * 0.05).half()
    b = (torch.randn(N, generator=g) * 0.1).half()
    x = torch.randn(M, K, generator=g).half()
    ref = orc.linear(w, b, x, relu=bool(epi))
    got = capi.linear_f16(w.to(_dev()), b.to(_dev()), x.to(_dev()), epilogue=epi).cpu()
    assert torch.equal(got.view(torch.int16), ref.view(torch.int16))
    ref_nb = orc.linear(w, None, x, relu=False)
    got_nb = capi.linear_f16(w.to(_dev()), None, x.to(_dev())).cpu()
    assert torch.equal(got_nb.view(torch.int16), ref_nb.view(torch.int16))


@gpu
def test_linear_gelu_tolerance():
    from meshanything_b200 import capi
    g = torch.Generator().manual_seed(5)
    w = (torch.randn(300, 768, generator=g) * 0.05).half()
    b = (torch.randn(300, generator=g) * 0.1).half()
    x = torch.randn(17, 768, generator=g).half()
    got = capi.linear_f16(w.to(_dev()), b.to(_dev()), x.to(_dev()), epilogue=capi.EPI_GELU).cpu().float()
    pre = (x.double() @ w.double().T + b.double()).half().float()
    ref = torch.nn.functional.gelu(pre)
    assert (got - ref).abs().max() < 4e-3  # fp16 output rounding (1 ulp at |y| < 4)


@gpu
@pytest.mark.parametrize("W", [768, 1024])
def test_layernorm_bit_exact(W):
    from meshanything_b200 import capi
    from oracle import decoder as orc
    g = torch.Generator().manual_seed(W)
    x = torch.randn(11, W, generator=g) * 2
    r = torch.randn(11, W, generator=g).half()
    gamma = 1 + 0.1 * torch.randn(W, generator=g)
    beta = 0.1 * torch.randn(W, generator=g)
    for eps in (1e-5, 1e-12):
        ref32, ref16 = orc.layernorm(x, r, gamma, beta, eps)
        o32, o16 = capi.layernorm(x.to(_dev()), r.to(_dev()), gamma.to(_dev()), beta.to(_dev()), eps)
        assert torch.equal(o32.cpu().view(torch.int32), ref32.view(torch.int32))
        assert torch.equal(o16.cpu().view(torch.int16), ref16.view(torch.int16))
    ref32, _ = orc.layernorm(x, None, gamma, beta, 1e-5)
    o32, _ = capi.layernorm(x.to(_dev()), None, gamma.to(_dev()), beta.to(_dev()), 1e-5)
    assert torch.equal(o32.cpu().view(torch.int32), ref32.view(torch.int32))
    torch_ref = torch.nn.functional.layer_norm(x, (W,), gamma, beta, 1e-5)
    assert (o32.cpu() - torch_ref).abs().max() < 1e-5


@gpu
@pytest.mark.parametrize("H,T,nk", [(16, 300, [1, 2, 31, 32, 33, 255, 256, 257, 300]), (12, 1100, [1100, 513, 1024]),
                                     (2, 4096, [4096])])
def test_attention_bit_exact(H, T, nk):
    from meshanything_b200 import capi
    from oracle import decoder as orc
    g = torch.Generator().manual_seed(T)
    M = len(nk)
    q = torch.randn(M, H, 64, generator=g).half()
    k = torch.randn(H, T, 64, generator=g).half()
    v = torch.randn(H, T, 64, generator=g).half()
    ref = orc.attention(q, k, v, nk)
    d = _dev()
    slots = torch.zeros(M, dtype=torch.int32, device=d)
    nkeys = torch.tensor(nk, dtype=torch.int32, device=d)
    got = capi.attention_f16(q.to(d), k.unsqueeze(0).contiguous().to(d), v.unsqueeze(0).contiguous().to(d), nkeys,
                             slots).cpu()
    assert torch.equal(got.view(torch.int16), ref.view(torch.int16))
    # and against plain fp32 softmax attention (tolerance: fp16 P and output rounding)
    for m, n in enumerate(nk):
        s = torch.einsum("hd,htd->ht", q[m].float(), k[:, :n].float()) * 0.125
        o = torch.einsum("ht,htd->hd", torch.softmax(s, -1), v[:, :n].float())
        assert (got[m].float() - o).abs().max() < 3e-3


@gpu
@pytest.mark.parametrize("M,N,K,epi", [(128, 128, 256, 0), (257, 768, 768, 0), (4096, 1536, 768, 0), (130, 128, 256, 1),
                                        (1057, 3072, 768, 2), (300, 768, 3072, 0), (64, 1152, 768, 0)])
def test_linear_tensor_core(M, N, K, epi):
    """tcgen05/TMA GEMM (encoder / detokenizer): fp16 in, fp32 accumulate in the hardware's order -> compared with an
    fp64 product rounded once to fp16 (tolerance: one fp16 ulp + fp32 accumulation noise) and with the canonical kernel."""
    from meshanything_b200 import capi
    g = torch.Generator().manual_seed(M + N + K)
    w = (torch.randn(N, K, generator=g) * 0.05).half()
    b = (torch.randn(N, generator=g) * 0.1).half()
    x = torch.randn(M, K, generator=g).half()
    d = _dev()
    got = capi.linear_tc_f16(w.to(d), b.to(d), x.to(d), epilogue=epi).cpu()
    canon = capi.linear_f16(w.to(d), b.to(d), x.to(d), epilogue=epi).cpu()
    pre = x.double() @ w.double().T + b.double()
    if epi == 1:
        pre = torch.relu(pre)
    elif epi == 2:
        pre = torch.nn.functional.gelu(pre.half().double())
    tol = 2.0 ** -10 * pre.abs() + 2e-3
    assert ((got.double() - pre).abs() <= tol).all(), (got.double() - pre).abs().max()
    # same values as the canonical kernel up to the last fp16 bit in a small fraction of the entries
    diff = (got.float() - canon.float()).abs()
    assert (diff <= 2.0 ** -9 * canon.float().abs() + 1e-3).all()
    assert (diff == 0).float().mean() > 0.98
