"""Host side of the Michelangelo point-cloud encoder (a1-a8) and of the VQ detokenizer (a17-a18):
fp16/fp32 device copies of the weights behind the C structs of include/meshanything_b200.h.

State-dict keys: /root/reference/MeshAnything/miche/michelangelo/models/tsal/sal_perceiver.py (encoder),
/root/reference/MeshAnything/models/meshanything.py:12-41 (tokenizer, BERT in BetterTransformer spelling).
"""
from __future__ import annotations

import ctypes as C
from typing import Dict, Optional

import torch

from . import capi
from .config import ENC

_vp = C.c_void_p
_S = "point_encoder.model.shape_model"


class MicheBlock(C.Structure):
    _fields_ = [(n, _vp) for n in ("c_qkv_w", "c_proj_w", "c_proj_b", "ln1_g", "ln1_b", "ln2_g", "ln2_b",
                                   "fc_w", "fc_b", "proj_w", "proj_b")]


class EncoderWeights(C.Structure):
    _fields_ = ([(n, _vp) for n in ("input_proj_w", "input_proj_b", "query", "cq_w", "ckv_w", "cproj_w", "cproj_b",
                                    "ln1_g", "ln1_b", "ln2_g", "ln2_b", "ln3_g", "ln3_b",
                                    "fc_w", "fc_b", "proj_w", "proj_b")]
                + [("enc", MicheBlock * 8)]
                + [(n, _vp) for n in ("lnpost_g", "lnpost_b", "pre_kl_w", "pre_kl_b", "post_kl_w", "post_kl_b")]
                + [("dec", MicheBlock * 16)]
                + [(n, _vp) for n in ("cond_head_w", "cond_head_b", "cond_w", "cond_b")])


class BertLayer(C.Structure):
    _fields_ = [(n, _vp) for n in ("in_w", "in_b", "out_w", "out_b", "l1_w", "l1_b", "l2_w", "l2_b",
                                   "n1_g", "n1_b", "n2_g", "n2_b")]


class TokenizerWeights(C.Structure):
    _fields_ = ([("n_layers", C.c_int), ("layer", BertLayer * 8)]
                + [(n, _vp) for n in ("pos_embedding", "point_pe", "ln_g", "ln_b", "pln_g", "pln_b",
                                      "cond_w", "cond_b", "cond_head_w", "cond_head_b", "down_w", "down_b",
                                      "coor_w", "coor_b", "codebook")])


class _Arena:
    def __init__(self, device: torch.device):
        if device.type != "cuda":
            raise RuntimeError("meshanything_b200 needs a CUDA device (no CPU fallback)")
        self.device = device
        self.keep = []

    def h16(self, t: torch.Tensor, pad_cols: Optional[int] = None) -> int:
        t = t.detach().to(device=self.device, dtype=torch.float16)
        if pad_cols is not None and t.shape[1] < pad_cols:      # zero-pad K to a multiple of 256 (canonical Linear)
            p = torch.zeros((t.shape[0], pad_cols), dtype=torch.float16, device=self.device)
            p[:, :t.shape[1]] = t
            t = p
        t = t.contiguous()
        self.keep.append(t)
        return t.data_ptr()

    def f32(self, t: torch.Tensor) -> int:
        t = t.detach().to(device=self.device, dtype=torch.float32).contiguous()
        if t.data_ptr() % 256:          # a view into a packed buffer: kernels use 16-byte loads
            t = t.clone()
        self.keep.append(t)
        return t.data_ptr()


class EncoderArena(_Arena):
    def __init__(self, sd: Dict[str, torch.Tensor], device: torch.device):
        super().__init__(device)
        w = EncoderWeights()
        e = f"{_S}.encoder"
        w.input_proj_w = self.h16(sd[f"{e}.input_proj.weight"], pad_cols=256)
        w.input_proj_b = self.h16(sd[f"{e}.input_proj.bias"])
        w.query = self.f32(sd[f"{e}.query"])
        c = f"{e}.cross_attn"
        w.cq_w = self.h16(sd[f"{c}.attn.c_q.weight"])
        w.ckv_w = self.h16(sd[f"{c}.attn.c_kv.weight"])
        w.cproj_w = self.h16(sd[f"{c}.attn.c_proj.weight"])
        w.cproj_b = self.h16(sd[f"{c}.attn.c_proj.bias"])
        for i, n in ((1, "ln_1"), (2, "ln_2"), (3, "ln_3")):
            setattr(w, f"ln{i}_g", self.f32(sd[f"{c}.{n}.weight"]))
            setattr(w, f"ln{i}_b", self.f32(sd[f"{c}.{n}.bias"]))
        w.fc_w = self.h16(sd[f"{c}.mlp.c_fc.weight"])
        w.fc_b = self.h16(sd[f"{c}.mlp.c_fc.bias"])
        w.proj_w = self.h16(sd[f"{c}.mlp.c_proj.weight"])
        w.proj_b = self.h16(sd[f"{c}.mlp.c_proj.bias"])
        for i in range(ENC.enc_layers):
            self._block(w.enc[i], sd, f"{e}.self_attn.resblocks.{i}")
        w.lnpost_g = self.f32(sd[f"{e}.ln_post.weight"])
        w.lnpost_b = self.f32(sd[f"{e}.ln_post.bias"])
        w.pre_kl_w = self.h16(sd[f"{_S}.pre_kl.weight"])
        w.pre_kl_b = self.h16(sd[f"{_S}.pre_kl.bias"])
        w.post_kl_w = self.h16(sd[f"{_S}.post_kl.weight"], pad_cols=256)
        w.post_kl_b = self.h16(sd[f"{_S}.post_kl.bias"])
        for i in range(ENC.dec_layers):
            self._block(w.dec[i], sd, f"{_S}.transformer.resblocks.{i}")
        w.cond_head_w = self.h16(sd["cond_head_proj.weight"])
        w.cond_head_b = self.h16(sd["cond_head_proj.bias"])
        w.cond_w = self.h16(sd["cond_proj.weight"])
        w.cond_b = self.h16(sd["cond_proj.bias"])
        self.c = w
        self._ws = None

    def _block(self, b: MicheBlock, sd, name: str):
        b.c_qkv_w = self.h16(sd[f"{name}.attn.c_qkv.weight"])
        b.c_proj_w = self.h16(sd[f"{name}.attn.c_proj.weight"])
        b.c_proj_b = self.h16(sd[f"{name}.attn.c_proj.bias"])
        b.ln1_g = self.f32(sd[f"{name}.ln_1.weight"])
        b.ln1_b = self.f32(sd[f"{name}.ln_1.bias"])
        b.ln2_g = self.f32(sd[f"{name}.ln_2.weight"])
        b.ln2_b = self.f32(sd[f"{name}.ln_2.bias"])
        b.fc_w = self.h16(sd[f"{name}.mlp.c_fc.weight"])
        b.fc_b = self.h16(sd[f"{name}.mlp.c_fc.bias"])
        b.proj_w = self.h16(sd[f"{name}.mlp.c_proj.weight"])
        b.proj_b = self.h16(sd[f"{name}.mlp.c_proj.bias"])

    def forward(self, pc_normal: torch.Tensor):
        """pc_normal fp16 [B,4096,6] (device) -> (point_feature fp32 [B,257,768], prefix fp32 [B,257,1024])."""
        if not pc_normal.is_cuda:
            raise RuntimeError("pc_normal must be on the CUDA device")
        assert pc_normal.dim() == 3 and pc_normal.shape[1] == ENC.n_points and pc_normal.shape[2] == 6
        pc = pc_normal.to(torch.float16).contiguous()
        B = pc.shape[0]
        L = capi.lib()
        need = L.ma_encoder_workspace_bytes(B)
        if self._ws is None or self._ws.numel() < need:
            self._ws = torch.empty(need, dtype=torch.uint8, device=self.device)
        pf = torch.empty((B, 257, 768), dtype=torch.float32, device=self.device)
        prefix = torch.empty((B, 257, 1024), dtype=torch.float32, device=self.device)
        capi.check(L.ma_encoder_forward(C.byref(self.c), capi.ptr(pc), B, capi.ptr(pf), capi.ptr(prefix),
                                        capi.ptr(self._ws), capi.stream_ptr()), "ma_encoder_forward")
        return pf, prefix


class TokenizerArena(_Arena):
    def __init__(self, sd: Dict[str, torch.Tensor], device: torch.device):
        super().__init__(device)
        w = TokenizerWeights()
        n = 0
        while f"tokenizer.decoder.layer.{n}.in_proj_weight" in sd or \
                f"tokenizer.decoder.layer.{n}.attention.self.query.weight" in sd:
            n += 1
        w.n_layers = n
        for i in range(n):
            self._layer(w.layer[i], sd, f"tokenizer.decoder.layer.{i}")
        t = "tokenizer"
        w.pos_embedding = self.f32(sd[f"{t}.pos_embedding.weight"])
        w.point_pe = self.f32(sd[f"{t}.point_pe.weight"])
        w.ln_g, w.ln_b = self.f32(sd[f"{t}.layernorm.weight"]), self.f32(sd[f"{t}.layernorm.bias"])
        w.pln_g, w.pln_b = self.f32(sd[f"{t}.point_layernorm.weight"]), self.f32(sd[f"{t}.point_layernorm.bias"])
        w.cond_w, w.cond_b = self.h16(sd[f"{t}.cond_proj.weight"]), self.h16(sd[f"{t}.cond_proj.bias"])
        w.cond_head_w = self.h16(sd[f"{t}.cond_head_proj.weight"])
        w.cond_head_b = self.h16(sd[f"{t}.cond_head_proj.bias"])
        w.down_w = self.h16(sd[f"{t}.project_down_codebook.weight"])
        w.down_b = self.h16(sd[f"{t}.project_down_codebook.bias"])
        w.coor_w = self.h16(sd[f"{t}.to_coor_logits.0.weight"])
        w.coor_b = self.h16(sd[f"{t}.to_coor_logits.0.bias"])
        w.codebook = self.f32(sd["transformer.model.decoder.quantize_codebooks"][0])
        self.c = w
        self._ws = None

    def _layer(self, l: BertLayer, sd, name: str):
        if f"{name}.in_proj_weight" in sd:        # optimum BetterTransformer spelling (the published checkpoint)
            g = lambda k: sd[f"{name}.{k}"]
            in_w, in_b = g("in_proj_weight"), g("in_proj_bias")
            keys = dict(out_w="out_proj_weight", out_b="out_proj_bias", l1_w="linear1_weight", l1_b="linear1_bias",
                        l2_w="linear2_weight", l2_b="linear2_bias", n1_g="norm1_weight", n1_b="norm1_bias",
                        n2_g="norm2_weight", n2_b="norm2_bias")
        else:                                      # plain HF BertLayer spelling
            a = f"{name}.attention"
            in_w = torch.cat([sd[f"{a}.self.{k}.weight"] for k in ("query", "key", "value")], 0)
            in_b = torch.cat([sd[f"{a}.self.{k}.bias"] for k in ("query", "key", "value")], 0)
            sd = dict(sd)
            keys = dict(out_w="attention.output.dense.weight", out_b="attention.output.dense.bias",
                        l1_w="intermediate.dense.weight", l1_b="intermediate.dense.bias",
                        l2_w="output.dense.weight", l2_b="output.dense.bias",
                        n1_g="attention.output.LayerNorm.weight", n1_b="attention.output.LayerNorm.bias",
                        n2_g="output.LayerNorm.weight", n2_b="output.LayerNorm.bias")
            g = lambda k: sd[f"{name}.{k}"]
        l.in_w, l.in_b = self.h16(in_w), self.h16(in_b)
        for f, k in keys.items():
            setattr(l, f, (self.f32 if f.startswith("n") else self.h16)(g(k)))

    def detokenize(self, gen_ids: torch.Tensor, point_feature: torch.Tensor, n_max_triangles: int,
                   want_ids: bool = False):
        """gen_ids int32 [B, 9F+2] (raw generate() output), point_feature fp32 [B,257,768] ->
        coords fp32 [B,F,3,3] with NaN rows for absent faces (meshanything.py:163-176)."""
        B, F = gen_ids.shape[0], n_max_triangles
        assert gen_ids.shape[1] == 9 * F + 2 and gen_ids.dtype == torch.int32 and gen_ids.is_cuda
        L = capi.lib()
        need = L.ma_detokenize_workspace_bytes(B, F)
        if self._ws is None or self._ws.numel() < need:
            self._ws = torch.empty(need, dtype=torch.uint8, device=self.device)
        out = torch.empty((B, F, 3, 3), dtype=torch.float32, device=self.device)
        ids_out = torch.empty((B, 9 * F), dtype=torch.int32, device=self.device) if want_ids else None
        capi.check(L.ma_detokenize(C.byref(self.c), capi.ptr(gen_ids.contiguous()), 9 * F + 2, B, F,
                                   capi.ptr(point_feature.contiguous()), capi.ptr(out), capi.ptr(ids_out),
                                   capi.ptr(self._ws), capi.stream_ptr()), "ma_detokenize")
        return (out, ids_out) if want_ids else out
