"""fp32 PyTorch restatement of the floating-point stages around the decode loop -- TEST INFRASTRUCTURE ONLY.

  encoder     a1-a8 : FourierEmbedder              miche/michelangelo/models/modules/embedder.py:87-105
                      CrossAttentionEncoder._forward        models/tsal/sal_perceiver.py:74-99
                      ResidualCrossAttentionBlock / ResidualAttentionBlock / MLP / attention
                                                            models/modules/transformer_blocks.py:41-74,109-112,147-185,223-244
                      encode_latents / encode_kl_embed / decode   sal_perceiver.py:372-396,273-275
                      AlignedShapeAsLatentPLModule.encode_latents / to_shape_latents   asl_pl_module.py:145-157,182-185
                      MeshAnything.process_point_feature    MeshAnything/models/meshanything.py:125-132
  detokenizer a17-a18: MeshAnything.get_codes               meshanything.py:178-212
                      NoiseResistantDecoder.forward / process_point_feature   meshanything.py:42-80
                      undiscretize                          meshanything.py:214-223
                      BERT layer (bert-base-uncased, post-LN, GELU, eps 1e-12) in optimum's BetterTransformer
                      key spelling (external: transformers==4.39.3 / optimum==1.18.0)

Everything is computed in fp32 on the CPU (no autocast): this is the "torch fp32 reference" the
floating-point CUDA kernels are compared with under a stated tolerance.  It is pinned against the
reference's own modules by tests/golden/make_golden_encoder.py (encoder: the reference's
`AlignedShapeLatentPerceiver` imported from /root/reference; detokenizer: transformers' BertLayer).
"""
from __future__ import annotations

import math
from typing import Dict, Tuple

import torch
import torch.nn.functional as F

ENC = "point_encoder.model.shape_model"


def _lin(sd, name, x, bias=True):
    return F.linear(x, sd[f"{name}.weight"], sd.get(f"{name}.bias") if bias else None)


def _ln(sd, name, x, eps=1e-5, sep="."):
    return F.layer_norm(x, (x.shape[-1],), sd[f"{name}{sep}weight"], sd[f"{name}{sep}bias"], eps)


def fourier_embed(x: torch.Tensor, num_freqs: int = 8) -> torch.Tensor:
    freqs = 2.0 ** torch.arange(num_freqs, dtype=torch.float32)       # include_pi: false (yaml:11)
    emb = (x[..., None].contiguous() * freqs).view(*x.shape[:-1], -1)
    return torch.cat((x, emb.sin(), emb.cos()), dim=-1)


def _attend(q, k, v, heads):
    """q [B,n,heads,64], k/v [B,s,heads,64] -> [B,n,heads*64]; scale 64^-1/4 on both (transformer_blocks.py:59,67-72)."""
    scale = 1 / math.sqrt(math.sqrt(q.shape[-1]))
    w = torch.einsum("bthc,bshc->bhts", q * scale, k * scale)
    w = torch.softmax(w.float(), dim=-1)
    return torch.einsum("bhts,bshc->bthc", w, v).reshape(q.shape[0], q.shape[1], -1)


def _self_block(sd, name, x, heads=12):
    bs, n, width = x.shape
    qkv = F.linear(_ln(sd, f"{name}.ln_1", x), sd[f"{name}.attn.c_qkv.weight"])
    qkv = qkv.view(bs, n, heads, -1)
    q, k, v = torch.split(qkv, width // heads, dim=-1)
    x = x + _lin(sd, f"{name}.attn.c_proj", _attend(q, k, v, heads))
    h = _lin(sd, f"{name}.mlp.c_proj", F.gelu(_lin(sd, f"{name}.mlp.c_fc", _ln(sd, f"{name}.ln_2", x))))
    return x + h


def encode_latents(sd: Dict[str, torch.Tensor], pc_normal: torch.Tensor) -> torch.Tensor:
    """pc_normal [B,4096,6] -> point_feature [B,257,768] (asl_pl_module.py:145-157)."""
    pc, feats = pc_normal[..., :3].float(), pc_normal[..., 3:6].float()
    data = torch.cat([fourier_embed(pc), feats], dim=-1)
    data = _lin(sd, f"{ENC}.encoder.input_proj", data)
    bs = pc.shape[0]
    x = sd[f"{ENC}.encoder.query"][None].expand(bs, -1, -1)
    c = f"{ENC}.encoder.cross_attn"
    heads, width = 12, 768
    q = F.linear(_ln(sd, f"{c}.ln_1", x), sd[f"{c}.attn.c_q.weight"]).view(bs, -1, heads, width // heads)
    kv = F.linear(_ln(sd, f"{c}.ln_2", data), sd[f"{c}.attn.c_kv.weight"]).view(bs, data.shape[1], heads, -1)
    k, v = torch.split(kv, width // heads, dim=-1)
    x = x + _lin(sd, f"{c}.attn.c_proj", _attend(q, k, v, heads))
    x = x + _lin(sd, f"{c}.mlp.c_proj", F.gelu(_lin(sd, f"{c}.mlp.c_fc", _ln(sd, f"{c}.ln_3", x))))
    for i in range(8):
        x = _self_block(sd, f"{ENC}.encoder.self_attn.resblocks.{i}", x)
    return _ln(sd, f"{ENC}.encoder.ln_post", x)


def to_shape_latents(sd, latents: torch.Tensor) -> torch.Tensor:
    """[B,256,768] -> [B,256,768]: pre_kl -> mean (first 64) -> post_kl -> 16 blocks (asl_pl_module.py:182-185)."""
    moments = _lin(sd, f"{ENC}.pre_kl", latents)
    mean = moments[..., :64]                                          # DiagonalGaussianDistribution.mode()
    x = _lin(sd, f"{ENC}.post_kl", mean)
    for i in range(16):
        x = _self_block(sd, f"{ENC}.transformer.resblocks.{i}", x)
    return x


def process_point_feature(sd, point_feature: torch.Tensor) -> torch.Tensor:
    """meshanything.py:125-132 -> prefix [B,257,1024]."""
    head = _lin(sd, "cond_head_proj", point_feature[:, 0])
    lat = to_shape_latents(sd, point_feature[:, 1:])
    rest = _lin(sd, "cond_proj", torch.cat([point_feature[:, 1:], lat], dim=-1))
    return torch.cat([head[:, None], rest], dim=1)


def encoder_forward(sd, pc_normal: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
    pf = encode_latents(sd, pc_normal)
    return pf, process_point_feature(sd, pf)


# ---------------------------------------------------------------------------- detokenizer

def postprocess_ids(results: torch.Tensor, n_max_triangles: int) -> torch.Tensor:
    """meshanything.py:142,163-172: pad with eos, strip first/last, specials -> -1, others -3."""
    B = results.shape[0]
    gen_len = n_max_triangles * 9 + 2
    out = torch.ones(B, gen_len, dtype=torch.long)
    out[:, :results.shape[1]] = results
    out = out[:, 1:-1].clone()
    special = (out == 0) | (out == 1) | (out == 2)
    out[special] = -1
    out[~special] -= 3
    return out


def get_codes(sd, ids: torch.Tensor) -> torch.Tensor:
    cb = sd["transformer.model.decoder.quantize_codebooks"][0]
    B = ids.shape[0]
    idx = ids.view(B, -1, 3)
    mask = idx == -1
    codes = cb[idx.masked_fill(mask, 0)]
    codes = codes.masked_fill(mask[..., None], 0.0)
    return codes.sum(dim=2)                                            # [B, 3F, 1024]


def _bert_layer(sd, name, x, heads=12, eps=1e-12):
    B, n, w = x.shape
    qkv = F.linear(x, sd[f"{name}.in_proj_weight"], sd[f"{name}.in_proj_bias"])
    q, k, v = qkv.split(w, dim=-1)
    sh = lambda t: t.view(B, n, heads, w // heads).transpose(1, 2)
    a = F.scaled_dot_product_attention(sh(q), sh(k), sh(v))
    a = a.transpose(1, 2).reshape(B, n, w)
    x = F.layer_norm(x + F.linear(a, sd[f"{name}.out_proj_weight"], sd[f"{name}.out_proj_bias"]), (w,),
                     sd[f"{name}.norm1_weight"], sd[f"{name}.norm1_bias"], eps)
    h = F.linear(F.gelu(F.linear(x, sd[f"{name}.linear1_weight"], sd[f"{name}.linear1_bias"])),
                 sd[f"{name}.linear2_weight"], sd[f"{name}.linear2_bias"])
    return F.layer_norm(x + h, (w,), sd[f"{name}.norm2_weight"], sd[f"{name}.norm2_bias"], eps)


def detokenize(sd, ids: torch.Tensor, point_feature: torch.Tensor, n_layers: int = 6,
               return_logits: bool = False):
    """ids [B,9F] in {-1, 0..8191}; point_feature [B,257,768] -> coords [B,F,3,3] (NaN rows = no face)."""
    B = ids.shape[0]
    t = "tokenizer"
    pf = torch.zeros(B, 257, 768)
    pf[:, 0] = _lin(sd, f"{t}.cond_head_proj", point_feature[:, 0])
    pf[:, 1:] = _lin(sd, f"{t}.cond_proj", point_feature[:, 1:])
    pf = _ln(sd, f"{t}.point_layernorm", pf + sd[f"{t}.point_pe.weight"][None, :257])
    codes = get_codes(sd, ids)
    nf = codes.shape[1] // 3
    face = _lin(sd, f"{t}.project_down_codebook", codes.view(B, nf, 3 * 1024))
    face_mask = (ids.view(B, nf, 9) != -1).all(dim=-1)
    face = face.masked_fill(~face_mask[..., None], 0.0)
    face = _ln(sd, f"{t}.layernorm", face + sd[f"{t}.pos_embedding.weight"][None, :nf])
    x = torch.cat([pf, face], dim=1)
    for i in range(n_layers):
        x = _bert_layer(sd, f"{t}.decoder.layer.{i}", x)
    dec = x[:, 257:].masked_fill(~face_mask[..., None], 0.0)
    logits = _lin(sd, f"{t}.to_coor_logits.0", dec).view(B, nf, 9, 128)
    coords = logits.argmax(dim=-1).view(B, nf, 3, 3).float() / 128 * 1.0 - 0.5
    coords = coords.masked_fill(~face_mask[:, :, None, None], float("nan"))
    return (coords, logits) if return_logits else coords
