"""not gpu: host-side pre/post-processing that main.py shares with the reference (main.py:15-58,156-175; mesh_to_pc.py)."""
import os

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CUBE = """v 0 0 0\nv 1 0 0\nv 1 1 0\nv 0 1 0\nv 0 0 1\nv 1 0 1\nv 1 1 1\nv 0 1 1
f 1 2 3 4\nf 5 8 7 6\nf 1 5 6 2\nf 2 6 7 3\nf 3 7 8 4\nf 5 1 4 8\n"""


def test_numpy_mesh_sampler(tmp_path):
    import mesh_to_pc
    p = tmp_path / "cube.obj"
    p.write_text(CUBE)
    mesh = mesh_to_pc.SimpleMesh.load_obj(str(p))
    assert mesh.faces.shape == (12, 3)                      # quads are fan-triangulated
    np.random.seed(0)
    clouds, used = mesh_to_pc.process_mesh_to_pc([mesh])
    pc = clouds[0]
    assert pc.shape == (4096, 6) and pc.dtype == np.float16 and used[0] is mesh
    xyz, nrm = pc[:, :3].astype(np.float32), pc[:, 3:].astype(np.float32)
    assert np.allclose(np.linalg.norm(nrm, axis=1), 1.0, atol=1e-3)
    on_face = np.isclose(xyz, 0, atol=2e-3) | np.isclose(xyz, 1, atol=2e-3)
    assert on_face.any(axis=1).all()                        # every sample lies on a cube face
    # area weighting: the six faces get ~1/6 of the samples each
    axis = np.argmax(np.abs(nrm), axis=1)
    side = (np.take_along_axis(nrm, axis[:, None], 1)[:, 0] > 0).astype(int)
    counts = np.bincount(axis * 2 + side, minlength=6)
    assert counts.min() > 4096 / 6 * 0.8


def test_dataset_normalisation(tmp_path, monkeypatch):
    monkeypatch.syspath_prepend(ROOT)
    from meshanything_b200.inputs import normalize_pc_normal, synthetic_pc_normal
    rng = np.random.default_rng(0)
    xyz = rng.uniform(-3, 5, size=(5000, 3))
    nrm = rng.normal(size=(5000, 3))
    nrm /= np.linalg.norm(nrm, axis=1, keepdims=True)
    out = normalize_pc_normal(np.concatenate([xyz, nrm], axis=1))
    c = out[:, :3].astype(np.float64)
    assert out.dtype == np.float16 and abs(np.abs(c).max() - 0.9995) < 1e-3
    assert np.allclose(c.min(0) + c.max(0), 0, atol=2e-3) or np.abs(c).max() <= 1.0   # centred on the bounding box
    with pytest.raises(AssertionError):
        normalize_pc_normal(np.concatenate([xyz, nrm * 0.5], axis=1))            # main.py:54
    s = synthetic_pc_normal(2, first=0)
    assert s.shape == (2, 4096, 6) and s.dtype.is_floating_point


def test_obj_export_merges_vertices_and_faces(tmp_path, monkeypatch):
    monkeypatch.syspath_prepend(ROOT)
    import importlib
    main = importlib.import_module("main")
    tri = np.array([[[0, 0, 0], [1, 0, 0], [0, 1, 0]],
                    [[1, 0, 0], [0, 1, 0], [1, 1, 0]],
                    [[0, 0, 0], [1, 0, 0], [0, 1, 0]]], dtype=np.float32)      # third face duplicates the first
    path = tmp_path / "m.obj"
    n = main.export_obj(str(path), tri)
    txt = path.read_text().splitlines()
    assert n == 2 and sum(l.startswith("f ") for l in txt) == 2
    assert sum(l.startswith("v ") for l in txt) == 4                            # 9 corners -> 4 distinct vertices
