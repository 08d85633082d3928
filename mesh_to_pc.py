"""Drop-in for /root/reference/mesh_to_pc.py: mesh -> (4096, 6) fp16 point cloud with normals.

Uses trimesh / mesh2sdf / skimage when they are installed (same calls as the reference); otherwise a
small numpy implementation of area-weighted surface sampling (what `trimesh.Trimesh.sample` does) is
used and `marching_cubes=True` raises (mesh2sdf is required for the watertight conversion).
"""
import numpy as np

try:  # optional host-side dependencies of the reference
    import trimesh
except Exception:  # pragma: no cover
    trimesh = None


class SimpleMesh:
    """Minimal triangle mesh (vertices [V,3], faces [F,3]) with the two members the pipeline needs."""

    def __init__(self, vertices, faces):
        self.vertices = np.asarray(vertices, dtype=np.float64)
        self.faces = np.asarray(faces, dtype=np.int64)

    @property
    def face_normals(self):
        t = self.vertices[self.faces]
        n = np.cross(t[:, 1] - t[:, 0], t[:, 2] - t[:, 0])
        ln = np.linalg.norm(n, axis=1, keepdims=True)
        return n / np.where(ln > 0, ln, 1.0)

    def sample(self, count, return_index=False):
        t = self.vertices[self.faces]
        area = 0.5 * np.linalg.norm(np.cross(t[:, 1] - t[:, 0], t[:, 2] - t[:, 0]), axis=1)
        idx = np.searchsorted(np.cumsum(area), np.random.random(count) * area.sum())
        idx = np.minimum(idx, len(area) - 1)
        r = np.random.random((count, 2))
        flip = r.sum(axis=1) > 1.0
        r[flip] = 1.0 - r[flip]
        tri = t[idx]
        pts = tri[:, 0] + r[:, :1] * (tri[:, 1] - tri[:, 0]) + r[:, 1:] * (tri[:, 2] - tri[:, 0])
        return (pts, idx) if return_index else pts

    @staticmethod
    def load_obj(path):
        vs, fs = [], []
        with open(path) as f:
            for line in f:
                p = line.split()
                if not p:
                    continue
                if p[0] == "v":
                    vs.append([float(x) for x in p[1:4]])
                elif p[0] == "f":
                    ids = [int(x.split("/")[0]) for x in p[1:]]
                    ids = [i - 1 if i > 0 else len(vs) + i for i in ids]
                    for k in range(1, len(ids) - 1):          # fan triangulation
                        fs.append([ids[0], ids[k], ids[k + 1]])
        return SimpleMesh(vs, fs)


def load_mesh(path):
    if trimesh is not None:
        return trimesh.load(path)
    if not path.endswith(".obj"):
        raise ImportError("trimesh is needed to load non-OBJ meshes")
    return SimpleMesh.load_obj(path)


def normalize_vertices(vertices, scale=0.9):
    """Centre on the bounding box and scale its longest side to 2*scale; returns (vertices, centre, factor)."""
    lo, hi = vertices.min(0), vertices.max(0)
    centre = 0.5 * (lo + hi)
    factor = 2.0 * scale / (hi - lo).max()
    return (vertices - centre) * factor, centre, factor


def export_to_watertight(normalized_mesh, octree_depth: int = 7):
    """Watertight remesh used by `--mc` (reference mesh_to_pc.py:13-40): unsigned distance field on a 2^depth grid
    (mesh2sdf), marching cubes at iso level 2/size, mapped back to the input frame."""
    try:
        import mesh2sdf.core
        import skimage.measure
    except Exception as e:  # pragma: no cover
        raise ImportError("--mc needs mesh2sdf, scikit-image and trimesh") from e
    size = 2 ** octree_depth
    unit_vertices, centre, factor = normalize_vertices(normalized_mesh.vertices)
    field = np.abs(mesh2sdf.core.compute(unit_vertices, normalized_mesh.faces, size=size))
    verts, faces, normals, _ = skimage.measure.marching_cubes(field, 2 / size)
    verts = (verts / size * 2 - 1) / factor + centre
    return trimesh.Trimesh(verts, faces, normals=normals)


def process_mesh_to_pc(mesh_list, marching_cubes=False, sample_num=4096):
    """[mesh] -> ([fp16 (sample_num, 6) points + face normals], [mesh actually sampled])."""
    clouds, used = [], []
    for mesh in mesh_list:
        if marching_cubes:
            mesh = export_to_watertight(mesh)
            print("MC over!")
        pts, tri = mesh.sample(sample_num, return_index=True)
        clouds.append(np.concatenate([pts, mesh.face_normals[tri]], axis=-1, dtype=np.float16))
        used.append(mesh)
        print("process mesh success")
    return clouds, used
