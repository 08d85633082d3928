"""not gpu: the CPU oracle against golden vectors and against plain fp64/fp32 torch math."""
import os

import numpy as np
import torch

from tests.util import decoder_sd, random_prefix

HERE = os.path.dirname(os.path.abspath(__file__))


def test_oracle_vs_hf_golden():
    """Teacher-forced logits of the oracle vs transformers' own OPTDecoderLayer + the reference's embedding
    code in fp32 (tests/golden/make_golden.py).  Tolerance: the oracle rounds to fp16 at the autocast
    points (measured max |diff| 5.1e-3 on logits of std 1.6); the HF run does not round."""
    from oracle.decoder import OracleDecoder
    g = np.load(os.path.join(HERE, "golden", "decoder_hf_fp32.npz"))
    forced, steps = g["forced"].tolist(), g["steps"].tolist()
    ref = torch.from_numpy(g["logits"])
    nl = int(g["n_layers"])
    o = OracleDecoder(decoder_sd(nl), nl, 257 + len(forced))
    _, lg = o.generate(random_prefix(1, seed=int(g["prefix_seed"]))[0], len(forced), eos_id=-1, forced=forced,
                       keep_logits=True)
    got = torch.stack([lg[s] for s in steps]).float()
    assert (got - ref).abs().max() < 2e-2
    top2 = torch.topk(ref, 2, dim=1).values
    clear = (top2[:, 0] - top2[:, 1]) > 4e-2
    assert torch.equal(got.argmax(1)[clear], ref.argmax(1)[clear])


def test_oracle_vs_hf_golden_full_depth():
    """The same pin at full depth: 24 layers, 320 teacher-forced positions (contexts 257..576, i.e. across the second and
    third attention chunk, special tokens also at the chunk boundary), EVERY position compared: the oracle's fp16
    logits against transformers' fp32 OPTDecoderLayer stack on every 8th vocabulary entry, and the argmax wherever the
    reference's top-2 margin exceeds the tolerance.  This pins the autocast rounding points the oracle mirrors
    (fp16 Linear outputs, fp16 P, fp32 LayerNorm / residual) over a long dependent chain."""
    from oracle.decoder import OracleDecoder
    g = np.load(os.path.join(HERE, "golden", "decoder_hf_fp32_deep.npz"))
    forced, cols = g["forced"].tolist(), torch.from_numpy(g["cols"]).long()
    ref = torch.from_numpy(g["logits16"]).float()
    nl = int(g["n_layers"])
    o = OracleDecoder(decoder_sd(nl), nl, 257 + len(forced))
    _, lg = o.generate(random_prefix(1, seed=int(g["prefix_seed"]))[0], len(forced), eos_id=-1, forced=forced,
                       keep_logits=True)
    got = torch.stack(lg).float()
    diff = (got[:, cols] - ref).abs()
    # tolerance: fp16 rounding of the activations through 24 layers + the fp16 storage of the fixture (2^-11 relative);
    # measured: max 6.1e-3, mean 9.8e-4 on logits of std 1.6; the argmax agrees at all 320 positions
    assert diff.max() < 2e-2 and diff.mean() < 3e-3, (float(diff.max()), float(diff.mean()))
    top2 = torch.from_numpy(g["top2"])
    clear = (top2[:, 0] - top2[:, 1]) > 2e-2
    assert int(clear.sum()) > 280
    assert torch.equal(got.argmax(1)[clear], torch.from_numpy(g["argmax"]).long()[clear])


def test_oracle_linear_against_fp64():
    from oracle.decoder import linear
    g = torch.Generator().manual_seed(0)
    w = (torch.randn(130, 1024, generator=g) * 0.05).half()
    b = (torch.randn(130, generator=g) * 0.1).half()
    x = torch.randn(7, 1024, generator=g).half()
    y = linear(w, b, x).double()
    ref = x.double() @ w.double().T + b.double()
    assert ((y - ref).abs() <= 2.0 ** -11 * ref.abs() + 1e-6).all()       # one fp16 rounding
    yr = linear(w, b, x, relu=True)
    assert torch.equal(yr, torch.clamp(linear(w, b, x), min=0))


def test_oracle_layernorm_against_torch():
    from oracle.decoder import layernorm
    g = torch.Generator().manual_seed(1)
    for W in (768, 1024):
        x = torch.randn(5, W, generator=g) * 3
        r = torch.randn(5, W, generator=g).half()
        ga, be = 1 + 0.1 * torch.randn(W, generator=g), 0.1 * torch.randn(W, generator=g)
        y, y16 = layernorm(x, r, ga, be)
        ref = torch.nn.functional.layer_norm(x + r.float(), (W,), ga, be, 1e-5)
        assert (y - ref).abs().max() < 2e-5
        assert torch.equal(y16, y.half())


def test_oracle_attention_against_softmax():
    from oracle.decoder import attention
    g = torch.Generator().manual_seed(2)
    H, T = 4, 700
    q = torch.randn(3, H, 64, generator=g).half()
    k = torch.randn(H, T, 64, generator=g).half()
    v = torch.randn(H, T, 64, generator=g).half()
    nk = [1, 256, 700]
    out = attention(q, k, v, nk)
    for m, n in enumerate(nk):
        s = torch.einsum("hd,htd->ht", q[m].double(), k[:, :n].double()) * 0.125
        o = torch.einsum("ht,htd->hd", torch.softmax(s, -1), v[:, :n].double())
        assert (out[m].double() - o).abs().max() < 3e-3


def test_oracle_exp():
    from oracle.decoder import lib
    L = lib()
    xs = np.linspace(-79.9, 0.0, 4001).astype(np.float32).astype(np.float64)
    got = np.array([L.orc_exp(float(x)) for x in xs])
    rel = np.abs(got - np.exp(xs)) / np.exp(xs)
    assert rel.max() < 6e-6              # |x| * 2^-24 from the rounding of x*log2(e)
    assert rel[xs > -4.0].max() < 4e-7   # polynomial error 7e-8 + rounding
    assert L.orc_exp(0.0) == 1.0 and L.orc_exp(-81.0) == 0.0


def test_oracle_kv_cache_rows_are_the_projections():
    """the K/V rows the oracle caches for the prefix are the k_proj / v_proj outputs of layer 0."""
    from oracle.decoder import OracleDecoder, linear
    sd = decoder_sd(3)
    o = OracleDecoder(sd, 3, 300)
    prefix = random_prefix(1, seed=3)[0]
    o.prefill(prefix)
    P = "transformer.model.decoder"
    h0 = (prefix + sd[f"{P}.cond_embed.weight"][0]) + sd[f"{P}.embed_positions.weight"][2:259]
    k = linear(sd[f"{P}.layers.0.self_attn.k_proj.weight"], sd[f"{P}.layers.0.self_attn.k_proj.bias"], h0.half())
    assert torch.equal(o.get_kv(0, 0, 100), k[100])


def test_oracle_generate_semantics():
    """HF generate(): eos stops a single row; forced ids are echoed; step count = max_new_tokens."""
    from oracle.decoder import OracleDecoder
    o = OracleDecoder(decoder_sd(3), 3, 257 + 16)
    prefix = random_prefix(1, seed=9)[0]
    ids, _ = o.generate(prefix, 16)
    assert len(ids) == 16
    eos = ids[4]
    first = ids.index(eos)
    ids2, _ = o.generate(prefix, 16, eos_id=eos)
    assert ids2 == ids[:first + 1]


def test_oracle_results_do_not_depend_on_the_openmp_team_size():
    """bench.py picks the OpenMP team size that is fastest on the host (`orc_set_threads`); every reduction of the oracle
    lives inside one thread, so ids and logits must be the same bits for any team size."""
    from oracle import decoder as orc
    from oracle.decoder import OracleDecoder
    sd = decoder_sd(2)
    prefix = random_prefix(1, seed=12)[0]
    outs = []
    try:
        for t in (1, 3, max(1, orc.max_threads())):
            orc.set_threads(t)
            ids, logits = OracleDecoder(sd, 2, 257 + 12).generate(prefix, 12, keep_logits=True)
            outs.append((ids, torch.as_tensor(np.asarray(logits)).clone()))
    finally:
        orc.set_threads(max(1, (os.cpu_count() or 1)))
    for ids, lg in outs[1:]:
        assert ids == outs[0][0]
        assert torch.equal(lg.view(torch.int16) if lg.dtype == torch.float16 else lg, outs[0][1].view(torch.int16)
                           if outs[0][1].dtype == torch.float16 else outs[0][1])


def test_bench_usable_cpus():
    import bench
    n = bench._usable_cpus()
    assert 1 <= n <= (os.cpu_count() or 1)


def test_config1_chain_against_reference_modules():
    """BASELINE configs[0] (the reference's CPU-runnable case) along the whole chain, on the CPU, against
    tests/golden/config1_chain.npz (make_golden_config1.py): mouse.npy through the reference's own Dataset, the
    reference's own AlignedShapeLatentPerceiver, the decoder oracle (578 greedy tokens), and transformers' BertEncoder
    with meshanything.py's detokenizer code.
      * oracle/torch_ref.py's encoder agrees with the reference's modules (fixture stored in fp16: tolerance
        1e-3 * |x| + 1e-3);
      * the C oracle reproduces the committed ids from the committed prefix exactly (pins it across toolchains,
        OpenMP team sizes and refactors);
      * torch_ref's detokenizer gives the HF BertEncoder's coordinate bins (>= 97 % equal, all faces present)."""
    from meshanything_b200 import checkpoint as ck
    from oracle import torch_ref
    from oracle.decoder import OracleDecoder
    g = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
    pc = torch.from_numpy(np.load(os.path.join(g, "config1_mouse.npz"))["pc_normal"][None])
    fx = np.load(os.path.join(g, "config1_chain.npz"))
    sd = ck.synthetic_state_dict(0)
    with torch.no_grad():
        pf, prefix = torch_ref.encoder_forward(sd, pc)
    for got, key in ((pf[0], "point_feature"), (prefix[0], "prefix")):
        ref = torch.from_numpy(fx[key]).float()
        assert ((got - ref).abs() <= 1e-3 * ref.abs() + 1e-3).all(), (key, float((got - ref).abs().max()))
    n = 9 * 64 + 2
    ids, _ = OracleDecoder(sd, 24, 257 + n).generate(torch.from_numpy(fx["prefix"]).float(), n)
    assert ids == fx["ids"].astype(np.int64).tolist()
    with torch.no_grad():
        coords = torch_ref.detokenize(sd, torch_ref.postprocess_ids(torch.tensor([ids]), 64),
                                      torch.from_numpy(fx["point_feature"]).float()[None])
    ref_coords = torch.from_numpy(fx["bins"].astype(np.float32)).view(64, 3, 3) / 128 - 0.5
    assert fx["face_mask"].all() and not torch.isnan(coords).any()
    assert float((coords[0] == ref_coords).float().mean()) >= 0.97
