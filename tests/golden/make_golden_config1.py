"""Golden for BASELINE configs[0] -- the reference's CPU-runnable case -- along the whole chain (dev container only).

  pc_normal   tests/golden/config1_mouse.npz: the reference's own Dataset on pc_examples/mouse.npy (make_golden_inputs.py)
  encoder     the reference's OWN AlignedShapeLatentPerceiver (sal_perceiver.py, strict load of the synthetic checkpoint,
              fp32, CPU) + the 8 wrapper lines of asl_pl_module.py / meshanything.py -- make_golden_encoder.reference_encoder
  decoder     oracle/decoder_oracle.c (greedy, 64-face cap -> 578 tokens) on that prefix ROUNDED TO fp16 (so the stored
              prefix reproduces the run exactly); the oracle is the only executable statement of the decoder here
              (ShapeOPT cannot be constructed under the installed transformers)
  detokenizer transformers' BertEncoder + meshanything.py:42-80,163-223 literally -- make_golden_encoder.reference_detok

Written to tests/golden/config1_chain.npz: prefix fp16 [257,1024], point_feature fp16 [257,768], ids int16 [578],
bins int16 [64,9], face_mask bool [64].  Tests: tests/test_oracle.py (CPU: torch_ref encoder vs the reference's prefix,
oracle ids reproduce, torch_ref detokenizer vs the HF one) and tests/test_gpu_pipeline.py (GPU: decoder ids from the
reference's prefix bit-exact, encoder / detokenizer within their tolerances).

usage: python tests/golden/make_golden_config1.py
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)

from meshanything_b200 import checkpoint as ck  # noqa: E402
from oracle import torch_ref  # noqa: E402
from oracle.decoder import OracleDecoder  # noqa: E402
import make_golden_encoder as mge  # noqa: E402

F = 64


def main():
    pc = torch.from_numpy(np.load(os.path.join(HERE, "config1_mouse.npz"))["pc_normal"][None])   # [1,4096,6] fp16
    sd = ck.synthetic_state_dict(0)
    pf, prefix = mge.reference_encoder(sd, pc)
    pf16, prefix16 = pf[0].half(), prefix[0].half()
    n = 9 * F + 2
    ids, _ = OracleDecoder(sd, 24, 257 + n).generate(prefix16.float(), n)
    assert len(ids) == n, "random weights never emit EOS"
    dec_ids = torch_ref.postprocess_ids(torch.tensor([ids]), F)                       # meshanything.py:163-172
    bins, _, face_mask = mge.reference_detok(sd, dec_ids, pf16.float()[None])
    out = os.path.join(HERE, "config1_chain.npz")
    np.savez_compressed(out, prefix=prefix16.numpy(), point_feature=pf16.numpy(), ids=np.asarray(ids, dtype=np.int16),
                        bins=bins[0].numpy().astype(np.int16), face_mask=face_mask[0].numpy())
    print("wrote", out, os.path.getsize(out), "bytes; first ids", ids[:8], "faces kept", int(face_mask.sum()))


if __name__ == "__main__":
    main()
