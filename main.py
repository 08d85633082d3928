"""Drop-in for /root/reference/main.py (same flags, same outputs) on top of the B200-native MeshAnything.

    python main.py --input_type pc_normal --input_path pc_examples/mouse.npy --out_dir out [--sampling]
    torchrun --nproc-per-node 8 main.py --input_type pc_normal --input_dir pcs --batchsize_per_gpu 64

Differences forced by the environment: no accelerate / hf_hub (there is no network) -- one process per
GPU is launched with torchrun, batches are dealt round-robin to the ranks as accelerate's prepared
DataLoader does (main.py:146), and weights come from `--pretrained_weights` (a local safetensors file
with the published keys) or, with `--pretrained_weights synthetic`, from the seeded random checkpoint.
"""
import argparse
import datetime
import os
import time

import numpy as np
import torch

from MeshAnything.models.meshanything import MeshAnything
from mesh_to_pc import load_mesh, process_mesh_to_pc


def _subsample_points(path, n_points=4096):
    """`--input_type pc_normal`: an .npy of >= 4096 (xyz, normal) rows; a random 4096-subset without replacement
    (global numpy RNG, seeded by --seed as the reference does through accelerate.set_seed)."""
    cloud = np.load(path)
    assert cloud.shape[0] >= n_points, "input pc_normal should have at least 4096 points"
    keep = np.random.choice(cloud.shape[0], n_points, replace=False)
    return cloud[keep]


def _uid_of(path):
    return path.split('/')[-1].split('.')[0]


class Dataset:
    """Same contract as the reference's Dataset (main.py:15-58): items are {'pc_normal': fp16 (4096, 6), 'uid': str},
    coordinates centred on the bounding box and scaled to max |x| = 0.9995, unit normals asserted."""

    def __init__(self, input_type, input_list, mc=False):
        if input_type == 'pc_normal':
            clouds = [_subsample_points(p) for p in input_list]
        elif input_type == 'mesh':
            if mc:
                print("First Marching Cubes and then sample point cloud, need several minutes...")
            clouds, _ = process_mesh_to_pc([load_mesh(p) for p in input_list], marching_cubes=mc)
        else:
            raise ValueError(f"unknown input_type {input_type!r}")
        self.data = [{'pc_normal': c, 'uid': _uid_of(p)} for c, p in zip(clouds, input_list)]
        print(f"dataset total data samples: {len(self.data)}")

    def __len__(self):
        return len(self.data)

    def __getitem__(self, idx):
        from meshanything_b200.inputs import normalize_pc_normal
        entry = self.data[idx]
        return {'pc_normal': normalize_pc_normal(entry['pc_normal']), 'uid': entry['uid']}


_FLAGS = [  # (flag, default, type) -- the reference's command line (main.py:60-89)
    ('--llm', "facebook/opt-350m", str), ('--input_dir', None, str), ('--input_path', None, str),
    ('--out_dir', "inference_out", str), ('--pretrained_weights', "MeshAnything_350m.pth", str),
    ('--codebook_size', 8192, int), ('--codebook_dim', 1024, int), ('--n_max_triangles', 800, int),
    ('--batchsize_per_gpu', 1, int), ('--seed', 0, int),
]


def get_args():
    parser = argparse.ArgumentParser("MeshAnything", add_help=False)
    for flag, default, typ in _FLAGS:
        parser.add_argument(flag, default=default, type=typ)
    parser.add_argument('--input_type', choices=['mesh', 'pc_normal'], default='pc',
                        help="Type of the asset to process (default: pc)")
    for switch in ('--mc', '--sampling'):
        parser.add_argument(switch, default=False, action="store_true")
    # not in the reference: run this rank's shapes through `batchsize_per_gpu` decoder cache slots, refilling a slot as
    # soon as its mesh is complete, instead of padded batches (MeshAnything.forward_queue)
    parser.add_argument('--continuous_batching', default=False, action="store_true")
    return parser.parse_args()


def load_model(args, device=None):
    model = MeshAnything(args)
    print("load model over!!!")
    if args.pretrained_weights == "synthetic":
        from meshanything_b200 import checkpoint
        tensors = checkpoint.synthetic_state_dict(0)
    else:
        from safetensors import safe_open
        if not os.path.exists(args.pretrained_weights):
            raise FileNotFoundError(f"{args.pretrained_weights}: put the published MeshAnything_350m.pth (safetensors) "
                                    "here, or pass --pretrained_weights synthetic")
        tensors = {}
        with safe_open(args.pretrained_weights, framework="pt", device="cpu") as f:
            for k in f.keys():
                tensors[k] = f.get_tensor(k)
    model.load_state_dict(tensors, strict=True, device=device)
    print("load weights over!!!")
    return model


def fix_winding(vertices, tri):
    """What `trimesh.Trimesh.fix_normals` does to the faces (main.py:166 of the reference), in numpy: make the winding
    consistent across every shared edge (breadth-first over face adjacency), then flip each connected component whose
    signed volume is negative so that normals point outwards."""
    tri = np.array(tri, dtype=np.int64, copy=True)
    nf = len(tri)
    if nf == 0:
        return tri
    # undirected edge -> faces that use it, with the direction each face traverses it in
    edges = {}
    for f in range(nf):
        for k in range(3):
            a, b = int(tri[f, k]), int(tri[f, (k + 1) % 3])
            if a != b:
                edges.setdefault((min(a, b), max(a, b)), []).append((f, a < b))
    adj = [[] for _ in range(nf)]
    for users in edges.values():
        if len(users) == 2:                       # manifold edge: consistent iff traversed in opposite directions
            (f0, d0), (f1, d1) = users
            adj[f0].append((f1, d0 == d1))
            adj[f1].append((f0, d0 == d1))
    comp = -np.ones(nf, dtype=np.int64)
    flip = np.zeros(nf, dtype=bool)
    ncomp = 0
    for seed in range(nf):
        if comp[seed] >= 0:
            continue
        comp[seed] = ncomp
        queue = [seed]
        while queue:
            f = queue.pop()
            for g, same_dir in adj[f]:
                if comp[g] < 0:
                    comp[g] = ncomp
                    flip[g] = flip[f] ^ same_dir     # same direction on the shared edge = opposite orientation
                    queue.append(g)
        ncomp += 1
    tri[flip] = tri[flip][:, ::-1]
    t = np.asarray(vertices, dtype=np.float64)[tri]
    vol6 = np.einsum("ij,ij->i", t[:, 0], np.cross(t[:, 1], t[:, 2]))
    for c in range(ncomp):
        sel = comp == c
        if vol6[sel].sum() < 0:
            tri[sel] = tri[sel][:, ::-1]
    return tri


def export_obj(path, faces_xyz):
    """merge_vertices + unique_faces + fix_normals + orange face colour of main.py:161-174 (trimesh when available,
    the numpy equivalents otherwise)."""
    vertices = faces_xyz.reshape(-1, 3)
    triangles = np.arange(len(vertices)).reshape(-1, 3)
    try:
        import trimesh
        mesh = trimesh.Trimesh(vertices=vertices, faces=triangles, force="mesh", merge_primitives=True)
        mesh.merge_vertices()
        mesh.update_faces(mesh.unique_faces())
        mesh.fix_normals()
        mesh.visual.face_colors = np.tile(np.array([255, 165, 0, 255], dtype=np.uint8), (len(mesh.faces), 1))
        mesh.export(path)
        return len(mesh.faces)
    except ImportError:
        uniq, inv = np.unique(np.round(vertices, 8), axis=0, return_inverse=True)
        tri = inv.reshape(-1)[triangles]
        _, keep = np.unique(np.sort(tri, axis=1), axis=0, return_index=True)
        tri = fix_winding(uniq, tri[np.sort(keep)])
        with open(path, "w") as f:
            for v in uniq:
                f.write(f"v {v[0]:.8f} {v[1]:.8f} {v[2]:.8f} 1.00000000 0.64705882 0.00000000\n")
            for t in tri:
                f.write(f"f {t[0] + 1} {t[1] + 1} {t[2] + 1}\n")
        return len(tri)


if __name__ == "__main__":
    args = get_args()
    from meshanything_b200 import parallel
    rank, world, local = parallel.init_from_env()
    cur_time = datetime.datetime.now().strftime("%d_%H-%M-%S")
    checkpoint_dir = os.path.join(args.out_dir, cur_time)
    os.makedirs(checkpoint_dir, exist_ok=True)
    device = torch.device("cuda", local)
    torch.cuda.set_device(device)
    model = load_model(args, device)

    if args.input_dir is not None:
        input_list = sorted(os.listdir(args.input_dir))
        if args.input_type == 'pc_normal':
            input_list = [os.path.join(args.input_dir, x) for x in input_list if x.endswith('.npy')]
        else:
            input_list = [os.path.join(args.input_dir, x) for x in input_list
                          if x.endswith('.ply') or x.endswith('.obj') or x.endswith('.npy')]
    elif args.input_path is not None:
        input_list = [args.input_path]
    else:
        raise ValueError("input_dir or input_path must be provided.")
    np.random.seed(args.seed)
    torch.manual_seed(args.seed)
    dataset = Dataset(args.input_type, input_list, args.mc)

    bs = args.batchsize_per_gpu
    batches = [list(range(i, min(i + bs, len(dataset)))) for i in range(0, len(dataset), bs)]
    begin_time = time.time()
    print("Generation Start!!!")

    def save(item, recon_mesh):
        recon_mesh = recon_mesh[~torch.isnan(recon_mesh[:, 0, 0])]
        save_path = os.path.join(checkpoint_dir, f'{item["uid"]}_gen.obj')
        export_obj(save_path, recon_mesh.cpu().numpy())
        print(f"{save_path} Over!!")

    if args.continuous_batching:
        mine = [dataset[i] for i in range(rank, len(dataset), world)]   # shapes are independent: round-robin by rank
        outs = model.forward_queue((torch.from_numpy(it['pc_normal']) for it in mine), sampling=args.sampling,
                                   slots=max(1, bs))
        for it, recon_mesh in zip(mine, outs):
            save(it, recon_mesh)
        batches = []
    for bi, idxs in enumerate(batches):
        if bi % world != rank:
            continue
        items = [dataset[i] for i in idxs]
        pc = torch.from_numpy(np.stack([it['pc_normal'] for it in items]))
        outputs = model(pc, sampling=args.sampling)
        for batch_id, it in enumerate(items):
            save(it, outputs[batch_id])
    print(f"Total time: {time.time() - begin_time}")
