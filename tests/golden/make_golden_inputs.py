"""Golden INPUT fixture for BASELINE configs[0] (the reference's CPU-runnable plumbing case).

Runs the reference's OWN `Dataset` class (main.py:15-58) on pc_examples/mouse.npy under numpy seed 0 (what
`accelerate.utils.set_seed(args.seed)` leaves in numpy, main.py:97) -- the class source is cut out of
/root/reference/main.py with `ast` and executed on its own, because importing that module pulls in flash-attn,
accelerate and trimesh, none of which its pc_normal branch uses.  Written to tests/golden/config1_mouse.npz:
  raw        fp16 (N, 6)   the file's contents (the input a user passes with --input_path)
  pc_normal  fp16 (4096, 6) what Dataset.__getitem__(0) hands to the model (subsampled, centred, scaled, normals checked)
Only runs in the development container (needs /root/reference); the fixture travels with the repo.

usage: python tests/golden/make_golden_inputs.py
"""
import ast
import os

import numpy as np

REF = "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))


def reference_dataset_class():
    src = open(os.path.join(REF, "main.py")).read()
    tree = ast.parse(src)
    node = next(n for n in tree.body if isinstance(n, ast.ClassDef) and n.name == "Dataset")
    ns = {"np": np}                      # the pc_normal branch needs numpy only
    exec(compile(ast.Module(body=[node], type_ignores=[]), os.path.join(REF, "main.py"), "exec"), ns)
    return ns["Dataset"]


def main():
    path = os.path.join(REF, "pc_examples", "mouse.npy")
    raw = np.load(path)
    Dataset = reference_dataset_class()
    np.random.seed(0)
    ds = Dataset("pc_normal", [path])
    item = ds[0]
    assert item["uid"] == "mouse" and item["pc_normal"].shape == (4096, 6) and item["pc_normal"].dtype == np.float16
    out = os.path.join(HERE, "config1_mouse.npz")
    np.savez_compressed(out, raw=raw, pc_normal=item["pc_normal"])
    print("wrote", out, raw.shape, raw.dtype, os.path.getsize(out), "bytes")


if __name__ == "__main__":
    main()
