"""Synthetic MeshAnything-350M checkpoint with the reference's exact state-dict keys and shapes.

No pretrained weights are reachable offline (`main.py:95-98` downloads them from the HF hub), so the
oracle, the tests and `bench.py` share a deterministic random checkpoint.  Key names/shapes follow
SURVEY.md section 8(b) "Weights", i.e. what `model.load_state_dict(tensors, strict=True)` at
/root/reference/main.py:104 expects:

  point_encoder.model.*   -- /root/reference/MeshAnything/miche/michelangelo/models/tsal/sal_perceiver.py
  tokenizer.*             -- /root/reference/MeshAnything/models/meshanything.py:12-41 (BERT keys in the
                             optimum-BetterTransformer spelling the published checkpoint was saved with)
  transformer.*           -- /root/reference/MeshAnything/models/shape_opt.py:188-235
  cond_head_proj, cond_proj -- meshanything.py:120-121

Every tensor is drawn from its own generator seeded by (seed, crc32(key)), so any subset of the
checkpoint (e.g. decoder only, or the first two layers) is bit-identical to the same tensors of the
full one.  Biases and LayerNorm affine parameters are deliberately non-trivial so that parity tests
are sensitive to them (the reference initialisers would leave them at 0 / 1).
"""
from __future__ import annotations

import zlib
from typing import Dict, Iterable, Optional, Tuple

import torch

from .config import DEC, ENC, TOK

Spec = Tuple[Tuple[int, ...], str, float]   # shape, kind, scale


def _linear(d: Dict[str, Spec], name: str, n_out: int, n_in: int, std: float, bias: bool = True):
    d[f"{name}.weight"] = ((n_out, n_in), "normal", std)
    if bias:
        d[f"{name}.bias"] = ((n_out,), "normal", 0.02)


def _ln(d: Dict[str, Spec], name: str, width: int, sep: str = "."):
    d[f"{name}{sep}weight"] = ((width,), "gamma", 0.05)
    d[f"{name}{sep}bias"] = ((width,), "normal", 0.02)


def decoder_specs(n_layers: int = DEC.n_layers) -> Dict[str, Spec]:
    d: Dict[str, Spec] = {}
    p = "transformer.model.decoder"
    h = DEC.hidden
    d[f"{p}.embed_tokens.weight"] = ((DEC.vocab, h), "normal", 0.02)          # unused (shape_opt.py:207)
    d[f"{p}.extra_embeds.weight"] = ((3, h), "normal", 0.5)
    _linear(d, f"{p}.input_layer", h, DEC.codebook_dim, 0.02)
    d[f"{p}.embed_positions.weight"] = ((DEC.n_positions + DEC.pos_offset, h), "normal", 0.5)
    d[f"{p}.token_embed_positions.weight"] = ((DEC.face_per_token + 3, h), "normal", 0.3)
    d[f"{p}.cond_embed.weight"] = ((2, h), "normal", 0.1)
    for i in range(n_layers):
        q = f"{p}.layers.{i}"
        for proj in ("q_proj", "k_proj", "v_proj"):
            _linear(d, f"{q}.self_attn.{proj}", h, h, 0.03)
        # small out_proj / fc2 keep the residual stream input-dependent through 24 post-LN layers, so
        # that greedy decoding of the random model does not collapse to one repeated token
        _linear(d, f"{q}.self_attn.out_proj", h, h, 0.01)
        _ln(d, f"{q}.self_attn_layer_norm", h)
        _linear(d, f"{q}.fc1", DEC.ffn, h, 0.03)
        _linear(d, f"{q}.fc2", h, DEC.ffn, 0.007)
        _ln(d, f"{q}.final_layer_norm", h)
    # meshanything.py:118 creates zeros; a trained checkpoint holds the VQ codebook here.
    d[f"{p}.quantize_codebooks"] = ((1, DEC.codebook_size, DEC.codebook_dim), "normal", 1.0)
    d["transformer.lm_head.weight"] = ((DEC.vocab, h), "normal", 0.05)
    _linear(d, "cond_head_proj", h, DEC.cond_dim, 0.03)
    _linear(d, "cond_proj", h, DEC.cond_dim * 2, 0.02)
    return d


def _miche_block(d: Dict[str, Spec], name: str, width: int):
    d[f"{name}.attn.c_qkv.weight"] = ((3 * width, width), "normal", 0.03)   # qkv_bias: false (yaml:18)
    _linear(d, f"{name}.attn.c_proj", width, width, 0.02)
    _ln(d, f"{name}.ln_1", width)
    _linear(d, f"{name}.mlp.c_fc", 4 * width, width, 0.03)
    _linear(d, f"{name}.mlp.c_proj", width, 4 * width, 0.015)
    _ln(d, f"{name}.ln_2", width)


def encoder_specs(include_unused: bool = True) -> Dict[str, Spec]:
    d: Dict[str, Spec] = {}
    w = ENC.width
    m = "point_encoder.model"
    s = f"{m}.shape_model"
    if include_unused:
        d[f"{m}.shape_projection"] = ((w, w), "zeros", 0.0)              # clip_asl_module.py:24 (unused)
    d[f"{s}.encoder.query"] = ((ENC.num_latents, w), "normal", 0.5)
    _linear(d, f"{s}.encoder.input_proj", w, ENC.fourier_dim + ENC.point_feats, 0.15)
    c = f"{s}.encoder.cross_attn"
    d[f"{c}.attn.c_q.weight"] = ((w, w), "normal", 0.03)
    d[f"{c}.attn.c_kv.weight"] = ((2 * w, w), "normal", 0.03)
    _linear(d, f"{c}.attn.c_proj", w, w, 0.02)
    _ln(d, f"{c}.ln_1", w)
    _ln(d, f"{c}.ln_2", w)
    _linear(d, f"{c}.mlp.c_fc", 4 * w, w, 0.03)
    _linear(d, f"{c}.mlp.c_proj", w, 4 * w, 0.015)
    _ln(d, f"{c}.ln_3", w)
    for i in range(ENC.enc_layers):
        _miche_block(d, f"{s}.encoder.self_attn.resblocks.{i}", w)
    _ln(d, f"{s}.encoder.ln_post", w)
    _linear(d, f"{s}.pre_kl", 2 * ENC.embed_dim, w, 0.03)
    _linear(d, f"{s}.post_kl", w, ENC.embed_dim, 0.1)
    for i in range(ENC.dec_layers):
        _miche_block(d, f"{s}.transformer.resblocks.{i}", w)
    if include_unused:
        # geo_decoder (sal_perceiver.py:115-159,228-240): never called by MeshAnything.forward, but
        # load_state_dict(strict=True) needs the keys.
        g = f"{s}.geo_decoder"
        _linear(d, f"{g}.query_proj", w, ENC.fourier_dim, 0.1)
        c = f"{g}.cross_attn_decoder"
        d[f"{c}.attn.c_q.weight"] = ((w, w), "normal", 0.03)
        d[f"{c}.attn.c_kv.weight"] = ((2 * w, w), "normal", 0.03)
        _linear(d, f"{c}.attn.c_proj", w, w, 0.02)
        _ln(d, f"{c}.ln_1", w)
        _ln(d, f"{c}.ln_2", w)
        _linear(d, f"{c}.mlp.c_fc", 4 * w, w, 0.03)
        _linear(d, f"{c}.mlp.c_proj", w, 4 * w, 0.015)
        _ln(d, f"{c}.ln_3", w)
        _ln(d, f"{g}.ln_post", w)
        _linear(d, f"{g}.output_proj", 1, w, 0.03)
    return d


def tokenizer_specs(n_layers: int = TOK.layers) -> Dict[str, Spec]:
    d: Dict[str, Spec] = {}
    w = TOK.width
    for i in range(n_layers):
        q = f"tokenizer.decoder.layer.{i}"
        d[f"{q}.in_proj_weight"] = ((3 * w, w), "normal", 0.03)
        d[f"{q}.in_proj_bias"] = ((3 * w,), "normal", 0.02)
        d[f"{q}.out_proj_weight"] = ((w, w), "normal", 0.02)
        d[f"{q}.out_proj_bias"] = ((w,), "normal", 0.02)
        d[f"{q}.linear1_weight"] = ((TOK.ffn, w), "normal", 0.03)
        d[f"{q}.linear1_bias"] = ((TOK.ffn,), "normal", 0.02)
        d[f"{q}.linear2_weight"] = ((w, TOK.ffn), "normal", 0.015)
        d[f"{q}.linear2_bias"] = ((w,), "normal", 0.02)
        _ln(d, f"{q}.norm1", w, sep="_")
        _ln(d, f"{q}.norm2", w, sep="_")
    d["tokenizer.pos_embedding.weight"] = ((TOK.max_faces, w), "normal", 0.1)
    _ln(d, "tokenizer.layernorm", w)
    _ln(d, "tokenizer.point_layernorm", w)
    d["tokenizer.point_pe.weight"] = ((TOK.cond_length, w), "normal", 0.1)
    _linear(d, "tokenizer.cond_proj", w, ENC.width, 0.03)
    _linear(d, "tokenizer.cond_head_proj", w, ENC.width, 0.03)
    _linear(d, "tokenizer.project_down_codebook", w, 3 * DEC.codebook_dim, 0.01)
    _linear(d, "tokenizer.to_coor_logits.0", TOK.discrete_num * 9, w, 0.05)
    return d


def all_specs(n_dec_layers: int = DEC.n_layers) -> Dict[str, Spec]:
    d = encoder_specs()
    d.update(tokenizer_specs())
    d.update(decoder_specs(n_dec_layers))
    return d


def _draw(key: str, spec: Spec, seed: int) -> torch.Tensor:
    shape, kind, scale = spec
    if kind == "zeros":
        return torch.zeros(shape, dtype=torch.float32)
    g = torch.Generator(device="cpu")
    g.manual_seed((seed * 1000003 + zlib.crc32(key.encode())) & 0x7FFFFFFFFFFFFFFF)
    t = torch.randn(shape, generator=g, dtype=torch.float32)
    if kind == "normal":
        return t * scale
    if kind == "gamma":
        return 1.0 + t * scale
    raise ValueError(kind)


def make_state_dict(specs: Dict[str, Spec], seed: int = 0,
                    keys: Optional[Iterable[str]] = None) -> Dict[str, torch.Tensor]:
    """Materialise (a subset of) a synthetic checkpoint as fp32 CPU tensors."""
    names = list(specs.keys()) if keys is None else list(keys)
    return {k: _draw(k, specs[k], seed) for k in names}


def synthetic_decoder_state_dict(seed: int = 0, n_layers: int = DEC.n_layers) -> Dict[str, torch.Tensor]:
    return make_state_dict(decoder_specs(n_layers), seed)


def synthetic_state_dict(seed: int = 0, n_dec_layers: int = DEC.n_layers) -> Dict[str, torch.Tensor]:
    return make_state_dict(all_specs(n_dec_layers), seed)


_FP32_PATTERNS = ("LayerNorm.", "ln_1.", "ln_2.", "ln_3.", "ln_post.", "layernorm.", "layer_norm.", "norm1_", "norm2_", "embed_positions",
                  "extra_embeds", "token_embed_positions", "cond_embed", "pos_embedding", "point_pe", ".query",
                  "quantize_codebooks", "shape_projection", "geo_decoder")


def consumed_as_fp16(name: str) -> bool:
    """True for the parameters the arenas keep as fp16 (`nn.Linear` weights and biases: exactly the values autocast
    would produce with `tensor.half()`); LayerNorm affine parameters, embedding tables, the Perceiver query and the VQ
    codebook stay fp32.  Used by parallel.broadcast_state_dict to ship fp16 where fp16 is what is kept."""
    return not any(p in name for p in _FP32_PATTERNS)
