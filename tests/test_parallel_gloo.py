"""not gpu: the multi-process host logic (weight broadcast at init, batch sharding) with world_size 2 on gloo."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from meshanything_b200 import checkpoint as ck
from meshanything_b200 import parallel


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      LOCAL_RANK=str(rank))
    r, w, _ = parallel.init_from_env(backend="gloo")
    specs = {k: v for k, v in ck.decoder_specs(1).items() if "layers.0" in k or "cond_" in k}
    sd = ck.make_state_dict(specs, 0) if r == 0 else None
    out = parallel.broadcast_state_dict(sd, specs, torch.device("cpu"))
    ref = ck.make_state_dict(specs, 0)
    # Linear parameters travel as fp16 (what the arenas keep), everything else as fp32, bit for bit
    ok = all(torch.equal(out[k], ref[k].half() if ck.consumed_as_fp16(k) else ref[k]) for k in specs)
    ok = ok and any(out[k].dtype == torch.float16 for k in specs) and any(out[k].dtype == torch.float32 for k in specs)
    lo, hi = parallel.shard_range(11, r, w)
    q.put((r, ok, lo, hi))
    dist.barrier()
    dist.destroy_process_group()


def test_broadcast_and_sharding_world2():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    assert all(ok for _, ok, _, _ in res)
    assert [(lo, hi) for _, _, lo, hi in res] == [(0, 6), (6, 11)]


def test_shard_range_covers_everything():
    for n in (0, 1, 7, 64, 512):
        for w in (1, 2, 4, 8):
            spans = [parallel.shard_range(n, r, w) for r in range(w)]
            covered = [i for lo, hi in spans for i in range(lo, hi)]
            assert covered == list(range(n))
