"""Golden vectors for the encoder and the detokenizer, produced by the reference's OWN code (dev container only).

encoder_ref_modules.npz
    /root/reference/MeshAnything/miche/michelangelo/models/tsal/sal_perceiver.py
    `AlignedShapeLatentPerceiver` (yaml params of shapevae-256.yaml:7-19) loaded with the synthetic
    checkpoint (strict) and run in fp32 on the CPU on a seeded synthetic (4096,6) point cloud;
    wrapped with asl_pl_module.py:145-157,182-185 and meshanything.py:125-132 (8 lines, restated here
    because those two modules need omegaconf/trimesh/optimum to import).
    Stored: point_feature[::4, ::8] and prefix[::4, ::8] (fp32) -- small strided views.

detok_hf_bert.npz
    transformers' BertEncoder (bert-base-uncased geometry, 6 layers, eps 1e-12, eager attention, fp32)
    loaded with the synthetic tokenizer weights mapped from the BetterTransformer spelling; everything
    around it follows meshanything.py:42-80,178-223 literally.  Stored: coordinates bins and a
    strided view of the 9x128 logits.
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from meshanything_b200 import checkpoint as ck  # noqa: E402
from meshanything_b200.inputs import synthetic_pc_normal  # noqa: E402


def reference_encoder(sd, pc_normal):
    sys.path.insert(0, "/root/reference")
    import warnings
    warnings.filterwarnings("ignore")
    from MeshAnything.miche.michelangelo.models.tsal.sal_perceiver import AlignedShapeLatentPerceiver
    m = AlignedShapeLatentPerceiver(device=None, dtype=None, num_latents=256, embed_dim=64, point_feats=3,
                                    num_freqs=8, include_pi=False, heads=12, width=768, num_encoder_layers=8,
                                    num_decoder_layers=16, use_ln_post=True, init_scale=0.25, qkv_bias=False,
                                    use_checkpoint=True).eval()
    pre = "point_encoder.model.shape_model."
    m.load_state_dict({k[len(pre):]: v for k, v in sd.items() if k.startswith(pre)}, strict=True)
    with torch.no_grad():
        pc, feats = pc_normal[..., :3].float(), pc_normal[..., 3:6].float()
        shape_embed, latents = m.encode_latents(pc, feats)                       # asl_pl_module.py:150-152
        pf = torch.cat([shape_embed.unsqueeze(1), latents], dim=1)              # :153-157
        kl, _ = m.encode_kl_embed(pf[:, 1:], sample_posterior=False)            # asl_pl_module.py:184
        shape_latents = m.decode(kl)                                            # :185
        B = pf.shape[0]
        prefix = torch.zeros(B, 257, 1024)                                      # meshanything.py:126-130
        prefix[:, 0] = torch.nn.functional.linear(pf[:, 0], sd["cond_head_proj.weight"], sd["cond_head_proj.bias"])
        prefix[:, 1:] = torch.nn.functional.linear(torch.cat([pf[:, 1:], shape_latents], dim=-1),
                                                   sd["cond_proj.weight"], sd["cond_proj.bias"])
    return pf, prefix


def reference_detok(sd, ids, point_feature):
    from transformers import BertConfig
    from transformers.models.bert.modeling_bert import BertEncoder
    from einops import rearrange, reduce
    cfg = BertConfig(hidden_size=768, num_hidden_layers=6, num_attention_heads=12, intermediate_size=3072,
                     hidden_act="gelu", layer_norm_eps=1e-12, hidden_dropout_prob=0.0,
                     attention_probs_dropout_prob=0.0)
    cfg._attn_implementation = "eager"
    enc = BertEncoder(cfg).eval()
    mp = {}
    for i in range(6):
        s, d = f"tokenizer.decoder.layer.{i}", f"layer.{i}"
        wq, wk, wv = sd[f"{s}.in_proj_weight"].split(768, 0)
        bq, bk, bv = sd[f"{s}.in_proj_bias"].split(768, 0)
        mp.update({f"{d}.attention.self.query.weight": wq, f"{d}.attention.self.query.bias": bq,
                   f"{d}.attention.self.key.weight": wk, f"{d}.attention.self.key.bias": bk,
                   f"{d}.attention.self.value.weight": wv, f"{d}.attention.self.value.bias": bv,
                   f"{d}.attention.output.dense.weight": sd[f"{s}.out_proj_weight"],
                   f"{d}.attention.output.dense.bias": sd[f"{s}.out_proj_bias"],
                   f"{d}.attention.output.LayerNorm.weight": sd[f"{s}.norm1_weight"],
                   f"{d}.attention.output.LayerNorm.bias": sd[f"{s}.norm1_bias"],
                   f"{d}.intermediate.dense.weight": sd[f"{s}.linear1_weight"],
                   f"{d}.intermediate.dense.bias": sd[f"{s}.linear1_bias"],
                   f"{d}.output.dense.weight": sd[f"{s}.linear2_weight"], f"{d}.output.dense.bias": sd[f"{s}.linear2_bias"],
                   f"{d}.output.LayerNorm.weight": sd[f"{s}.norm2_weight"],
                   f"{d}.output.LayerNorm.bias": sd[f"{s}.norm2_bias"]})
    enc.load_state_dict(mp, strict=True)
    lin = lambda n, x: torch.nn.functional.linear(x, sd[f"tokenizer.{n}.weight"], sd[f"tokenizer.{n}.bias"])
    ln = lambda n, x: torch.nn.functional.layer_norm(x, (768,), sd[f"tokenizer.{n}.weight"], sd[f"tokenizer.{n}.bias"])
    with torch.no_grad():
        B = ids.shape[0]
        # get_codes (meshanything.py:178-212)
        idx = rearrange(ids, "b (n q) -> b n q", q=3)
        mask = idx == -1
        codes = sd["transformer.model.decoder.quantize_codebooks"][0][idx.masked_fill(mask, 0)]
        codes = codes.permute(2, 0, 1, 3).masked_fill(rearrange(mask, "b n q -> q b n 1"), 0.0)
        code_embed = reduce(codes, "q ... -> ...", "sum")
        # NoiseResistantDecoder.forward (meshanything.py:50-80)
        pf = torch.zeros(B, 257, 768)
        pf[:, 0] = lin("cond_head_proj", point_feature[:, 0])
        pf[:, 1:] = lin("cond_proj", point_feature[:, 1:])
        pf = ln("point_layernorm", pf + sd["tokenizer.point_pe.weight"][None, :257])
        face = lin("project_down_codebook", rearrange(code_embed, "b (nf nv) d -> b nf (nv d)", nv=3))
        face_mask = reduce(ids != -1, "b (nf nv q) -> b nf", "all", nv=3, q=3)
        face[~face_mask] = 0
        face = ln("layernorm", face + sd["tokenizer.pos_embedding.weight"][None, :face.shape[1]])
        out = enc(hidden_states=torch.cat([pf, face], dim=1)).last_hidden_state
        dec = out[:, 257:].masked_fill(~face_mask.unsqueeze(-1), 0.0)
        logits = rearrange(lin("to_coor_logits.0", dec), "... (v c) -> ... v c", v=9)
        bins = logits.argmax(dim=-1)
    return bins, logits, face_mask


def main():
    sd = ck.make_state_dict(ck.all_specs(1), 0)
    pc = synthetic_pc_normal(2, first=0)                    # [2,4096,6] fp16
    pf, prefix = reference_encoder(sd, pc)
    np.savez_compressed(os.path.join(HERE, "encoder_ref_modules.npz"),
                        point_feature=pf[:, ::4, ::8].numpy(), prefix=prefix[:, ::4, ::8].numpy())
    print("encoder golden", pf.shape, prefix.shape, float(pf.std()), float(prefix.std()))
    # detokenizer: 12 faces, a special token inside face 7 and everything after face 9 absent
    g = torch.Generator().manual_seed(5)
    F = 12
    ids = torch.randint(0, 8192, (2, 9 * F), generator=g)
    ids[0, 9 * 7 + 4] = -1
    ids[0, 9 * 10:] = -1
    ids[1, 9 * 11 + 8] = -1
    bins, logits, mask = reference_detok(sd, ids, pf)
    np.savez_compressed(os.path.join(HERE, "detok_hf_bert.npz"), ids=ids.numpy(), bins=bins.numpy(),
                        face_mask=mask.numpy(), logits=logits[:, :, :, ::4].numpy())
    print("detok golden", bins.shape, logits.shape)


if __name__ == "__main__":
    main()
