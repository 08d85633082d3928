"""Generates the committed golden fixtures under tests/golden/.  Run ONCE in the dev container
(`python tests/golden/make_golden.py [decoder|greedy|encoder|all]`): it imports the reference from
/root/reference (read-only) and transformers' own OPT layer, neither of which exists on the GPU box.

decoder_hf_fp32.npz
    fp32 logits of a 3-layer ShapeOPT decoder built from the reference's and its dependency's OWN code:
      * transformers.models.opt.modeling_opt.OPTDecoderLayer / OPTLearnedPositionalEmbedding
        (installed 5.5.0; same math as the pinned 4.39.3 -- post-LN, ReLU, learned positions offset 2),
        eager attention, fp32, causal mask;
      * /root/reference/MeshAnything/models/shape_opt.py: `ShapeOPTDecoder.embed_with_vae` (:237-245,
        called unbound on a stand-in object because the class cannot be constructed under
        transformers 5.x) and `OPTFacePositionalEmbedding.forward` (:448-460).
    It pins the oracle's restatement of the decoder math (tests/test_oracle_vs_hf.py, tolerance: the
    oracle rounds to fp16 where CUDA autocast does, this reference does not round at all).

decoder_greedy_seed0_F64.json
    greedy ids of the 24-layer synthetic decoder for config 1's length (F=64 -> 578 new tokens),
    produced by the CPU oracle; regression fixture for the GPU test (which also re-derives it).
"""
import json
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from tests.util import decoder_sd, random_prefix  # noqa: E402

P = "transformer.model.decoder"


def hf_reference_logits(sd, n_layers, prefix, ids):
    """fp32 logits for every generated position, teacher-forced on `ids` (full-sequence recompute)."""
    sys.path.insert(0, "/root/reference")
    from transformers import OPTConfig
    from transformers.models.opt.modeling_opt import OPTDecoderLayer, OPTLearnedPositionalEmbedding
    import MeshAnything.models.shape_opt as so

    cfg = OPTConfig(hidden_size=1024, num_hidden_layers=n_layers, ffn_dim=4096, num_attention_heads=16,
                    do_layer_norm_before=False, word_embed_proj_dim=1024, activation_function="relu",
                    enable_bias=True, layer_norm_elementwise_affine=True, dropout=0.0, attention_dropout=0.0)
    cfg._attn_implementation = "eager"
    layers = []
    for i in range(n_layers):
        layer = OPTDecoderLayer(cfg, layer_idx=i).eval()
        layer.load_state_dict({k[len(f"{P}.layers.{i}."):]: v for k, v in sd.items()
                               if k.startswith(f"{P}.layers.{i}.")}, strict=True)
        layers.append(layer)
    npos = sd[f"{P}.embed_positions.weight"].shape[0]
    pos_emb = OPTLearnedPositionalEmbedding(npos - 2, 1024)
    pos_emb.load_state_dict({"weight": sd[f"{P}.embed_positions.weight"]})
    face_pos = so.OPTFacePositionalEmbedding(12, 1024)
    face_pos.load_state_dict({"weight": sd[f"{P}.token_embed_positions.weight"]})
    stand_in = types.SimpleNamespace(
        word_embed_proj_dim=1024,
        extra_embeds=torch.nn.Embedding.from_pretrained(sd[f"{P}.extra_embeds.weight"]),
        input_layer=torch.nn.Linear(1024, 1024),
        quantize_codebooks=sd[f"{P}.quantize_codebooks"],
    )
    stand_in.input_layer.load_state_dict({"weight": sd[f"{P}.input_layer.weight"], "bias": sd[f"{P}.input_layer.bias"]})
    cond = sd[f"{P}.cond_embed.weight"]
    n = len(ids)
    with torch.no_grad():
        rows = [prefix + cond[0]]                                     # shape_opt.py:331-337
        for i in range(1, n):                                         # token fed at step i is ids[i-1]
            tok = torch.tensor([[ids[i - 1]]])
            mask = torch.ones(1, 257 + i, dtype=torch.long)   # HF generate keeps a LongTensor mask
            e = so.ShapeOPTDecoder.embed_with_vae(stand_in, tok)      # shape_opt.py:321
            e = e + face_pos(mask[:, 257:], None, tok, 9)             # shape_opt.py:323-325
            e = e + cond[1]                                           # shape_opt.py:326-328
            rows.append(e[0])
        emb = torch.cat(rows, dim=0)[None]                            # [1, 257+n-1, 1024]
        S = emb.shape[1]
        hidden = emb + pos_emb(torch.ones(1, S, dtype=torch.long), 0)  # shape_opt.py:359-364
        causal = torch.full((S, S), float("-inf")).triu(1)[None, None]
        for layer in layers:
            hidden = layer(hidden, attention_mask=causal)
        logits = hidden[0, 256:] @ sd["transformer.lm_head.weight"].T  # shape_opt.py:155
    return logits.numpy()                                              # [n, vocab]


def make_decoder():
    from oracle.decoder import OracleDecoder
    NL, n = 3, 24
    sd = decoder_sd(NL)
    prefix = random_prefix(1, seed=3)[0]
    oracle = OracleDecoder(sd, NL, 257 + n)
    ids, _ = oracle.generate(prefix, n)
    forced = list(ids)
    forced[5], forced[6], forced[7] = 0, 1, 2          # exercise the special-token embedding path
    ref = hf_reference_logits(sd, NL, prefix, forced)
    steps = [0, 1, 2, 5, 6, 7, 8, 9, 10, 23]            # keeps the fixture small (10 x 8195 fp32)
    np.savez_compressed(os.path.join(HERE, "decoder_hf_fp32.npz"), forced=np.asarray(forced, dtype=np.int32),
                        steps=np.asarray(steps, dtype=np.int32), logits=ref[steps].astype(np.float32), n_layers=NL,
                        prefix_seed=3)
    print("decoder_hf_fp32.npz", ref.shape)


def make_decoder_deep():
    """decoder_hf_fp32_deep.npz: the same pin at full depth -- 24 layers, 320 teacher-forced positions (contexts 257..576:
    the second and third 256-key attention chunk), EVERY position stored: fp16 logits of every 8th vocabulary entry
    (+ the three special tokens), the fp32 argmax and the top-2 values."""
    from oracle.decoder import OracleDecoder
    NL, n = 24, 320
    sd = decoder_sd(NL)
    prefix = random_prefix(1, seed=4)[0]
    oracle = OracleDecoder(sd, NL, 257 + n)
    ids, _ = oracle.generate(prefix, n, eos_id=-1)
    forced = list(ids)
    for pos, t in ((5, 0), (6, 1), (7, 2), (100, 1), (255, 2), (256, 0), (300, 1)):   # special-token embeddings, also at
        forced[pos] = t                                                              # the chunk boundary
    ref = torch.from_numpy(hf_reference_logits(sd, NL, prefix, forced))
    cols = sorted(set(range(0, ref.shape[1], 8)) | {0, 1, 2})
    top2 = torch.topk(ref, 2, dim=1)
    np.savez_compressed(os.path.join(HERE, "decoder_hf_fp32_deep.npz"), forced=np.asarray(forced, dtype=np.int32),
                        cols=np.asarray(cols, dtype=np.int32), logits16=ref[:, cols].half().numpy(),
                        argmax=top2.indices[:, 0].numpy().astype(np.int32), top2=top2.values.numpy().astype(np.float32),
                        n_layers=NL, prefix_seed=4)
    print("decoder_hf_fp32_deep.npz", ref.shape, "std", float(ref.std()))


def make_greedy(faces=64, n_layers=24, eos_id=1):
    """greedy ids of the synthetic decoder from the CPU oracle: F=64 is config 1's length (578 tokens), F=800
    config 2's (7202 tokens, contexts up to 7458: ~2.5 minutes of oracle time on 8 cores), F=1600 config 5's
    (14402 tokens, contexts up to 14658 = 58 attention chunks; the first 4 layers only, to bound the oracle time)."""
    from oracle.decoder import OracleDecoder
    sd = decoder_sd(n_layers)
    n = faces * 9 + 2
    prefix = random_prefix(1, seed=1)[0]
    oracle = OracleDecoder(sd, n_layers, 257 + n)
    ids, _ = oracle.generate(prefix, n, eos_id=eos_id)
    json.dump({"ids": ids, "n_layers": n_layers, "prefix_seed": 1, "checkpoint_seed": 0, "faces": faces, "eos_id": eos_id},
              open(os.path.join(HERE, f"decoder_greedy_seed0_F{faces}.json"), "w"))
    print(f"decoder_greedy_seed0_F{faces}.json", len(ids), ids[:12], "distinct", len(set(ids)))


if __name__ == "__main__":
    what = sys.argv[1] if len(sys.argv) > 1 else "all"
    if what in ("decoder", "all"):
        make_decoder()
    if what in ("decoder_deep", "all"):
        make_decoder_deep()
    if what in ("greedy", "all"):
        make_greedy(64)
    if what in ("greedy800", "all"):
        make_greedy(800)
    if what in ("greedy1600", "all"):
        make_greedy(1600, n_layers=4, eos_id=-1)   # length coverage: eos disabled so that all 14402 steps run
    if what in ("encoder", "all"):
        from tests.golden import make_golden_encoder
        make_golden_encoder.main()
