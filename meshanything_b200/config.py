"""Model constants of the MeshAnything-350M hot path.

Values follow the reference:
  decoder  : /root/reference/MeshAnything/models/meshanything.py:83-123 (ShapeOPTConfig from
             facebook/opt-350m: hidden 1024, 24 layers, 16 heads, ffn 4096, ReLU, post-LN,
             word_embed_proj_dim forced to hidden at :112-113, vocab 8192+3, 18259 positions (+2 offset))
  encoder  : /root/reference/MeshAnything/miche/shapevae-256.yaml:7-19
  tokenizer: /root/reference/MeshAnything/models/meshanything.py:12-41 (bert-base-uncased, 6 layers)
"""
from dataclasses import dataclass


@dataclass(frozen=True)
class DecoderConfig:
    hidden: int = 1024
    n_layers: int = 24
    n_heads: int = 16
    head_dim: int = 64
    ffn: int = 4096
    codebook_size: int = 8192
    codebook_dim: int = 1024
    vocab: int = 8195            # codebook_size + bos/eos/pad  (meshanything.py:99)
    n_positions: int = 18259     # meshanything.py:97-98 ; table has +2 rows (HF OPT offset)
    pos_offset: int = 2
    cond_length: int = 257       # meshanything.py:91
    cond_dim: int = 768
    face_per_token: int = 9      # meshanything.py:89-90
    bos_id: int = 0
    eos_id: int = 1
    pad_id: int = 2

    def max_new_tokens(self, n_max_triangles: int) -> int:
        # meshanything.py:93,140 : generate_length = n_max_triangles*9 + 2
        return n_max_triangles * self.face_per_token + 2

    def max_context(self, n_max_triangles: int) -> int:
        return self.cond_length + self.max_new_tokens(n_max_triangles)


@dataclass(frozen=True)
class EncoderConfig:
    num_latents: int = 257       # 1 + 256  (sal_perceiver.py:332)
    width: int = 768
    heads: int = 12
    head_dim: int = 64
    enc_layers: int = 8
    dec_layers: int = 16
    embed_dim: int = 64
    num_freqs: int = 8
    point_feats: int = 3
    n_points: int = 4096
    fourier_dim: int = 51        # 3 * (2*8 + 1)   (embedder.py:77-81)


@dataclass(frozen=True)
class TokenizerConfig:
    width: int = 768
    heads: int = 12
    head_dim: int = 64
    layers: int = 6
    ffn: int = 3072
    max_faces: int = 18000
    discrete_num: int = 128
    cond_length: int = 257
    ln_eps: float = 1e-12        # bert-base-uncased layer_norm_eps


DEC = DecoderConfig()
ENC = EncoderConfig()
TOK = TokenizerConfig()
