"""not gpu: the fp32 torch restatement (oracle/torch_ref.py) against goldens made by the reference's own
modules (tests/golden/make_golden_encoder.py)."""
import os

import numpy as np
import torch

from meshanything_b200 import checkpoint as ck
from meshanything_b200.inputs import synthetic_pc_normal

HERE = os.path.dirname(os.path.abspath(__file__))
_cache = {}


def _sd():
    if "sd" not in _cache:
        _cache["sd"] = ck.make_state_dict(ck.all_specs(1), 0)
    return _cache["sd"]


def _pf():
    if "pf" not in _cache:
        from oracle import torch_ref
        with torch.no_grad():
            _cache["pf"] = torch_ref.encoder_forward(_sd(), synthetic_pc_normal(2, first=0))
    return _cache["pf"]


def test_encoder_restatement_matches_reference_modules():
    g = np.load(os.path.join(HERE, "golden", "encoder_ref_modules.npz"))
    pf, prefix = _pf()
    # both fp32 on the CPU, same math, different op grouping: 1e-4 on O(1) values
    assert np.abs(pf[:, ::4, ::8].numpy() - g["point_feature"]).max() < 2e-4
    assert np.abs(prefix[:, ::4, ::8].numpy() - g["prefix"]).max() < 2e-4


def test_detokenizer_restatement_matches_hf_bert():
    from oracle import torch_ref
    g = np.load(os.path.join(HERE, "golden", "detok_hf_bert.npz"))
    ids = torch.from_numpy(g["ids"])
    pf, _ = _pf()
    with torch.no_grad():
        coords, logits = torch_ref.detokenize(_sd(), ids, pf, return_logits=True)
    assert np.abs(logits[:, :, :, ::4].numpy() - g["logits"]).max() < 2e-3
    mask = torch.from_numpy(g["face_mask"])
    bins = torch.from_numpy(g["bins"]).view(2, -1, 3, 3)
    exp = bins.float() / 128 - 0.5
    assert torch.equal(torch.isnan(coords[:, :, 0, 0]), ~mask)
    assert torch.equal(coords[mask], exp[mask])


def test_postprocess_ids():
    """meshanything.py:142,163-172 on a hand-made generate() result."""
    from oracle import torch_ref
    F = 2
    res = torch.tensor([[0, 10, 11, 12, 13, 14, 15, 16, 17, 18, 1, 2, 2]])          # bos, 9 tokens, eos, pad, pad
    out = torch_ref.postprocess_ids(res, F)
    assert out.shape == (1, 9 * F)
    assert out[0, :9].tolist() == [7, 8, 9, 10, 11, 12, 13, 14, 15]
    assert (out[0, 9:] == -1).all()


def test_checkpoint_has_reference_keys():
    """the synthetic checkpoint loads strictly into the reference's encoder (done by make_golden_encoder.py);
    here: key families and shapes of SURVEY.md 8b."""
    specs = ck.all_specs(24)
    assert specs["transformer.model.decoder.embed_positions.weight"][0] == (18261, 1024)
    assert specs["transformer.model.decoder.quantize_codebooks"][0] == (1, 8192, 1024)
    assert specs["transformer.lm_head.weight"][0] == (8195, 1024)
    assert specs["point_encoder.model.shape_model.encoder.query"][0] == (257, 768)
    assert specs["point_encoder.model.shape_model.encoder.input_proj.weight"][0] == (768, 54)
    assert specs["tokenizer.decoder.layer.5.in_proj_weight"][0] == (2304, 768)
    assert specs["tokenizer.to_coor_logits.0.weight"][0] == (1152, 768)
    assert specs["cond_proj.weight"][0] == (1024, 1536)
    n_dec = sum(int(np.prod(s[0])) for k, s in specs.items() if k.startswith("transformer.model.decoder.layers."))
    assert n_dec == 24 * 12596224          # per-layer params of SURVEY.md 8d
