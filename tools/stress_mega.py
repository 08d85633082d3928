"""Repeats full-length batch-1 generates with the persistent kernel, checking the poll-timeout flag, the time of every
run and that all runs produce identical ids (python tools/stress_mega.py [runs] [faces])."""
import sys, time, torch
import os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from meshanything_b200 import capi
from meshanything_b200.checkpoint import decoder_specs, make_state_dict
from meshanything_b200.decoder import DecoderArena, Generator
from bench import synthetic_prefix
runs = int(sys.argv[1]) if len(sys.argv) > 1 else 20
F = int(sys.argv[2]) if len(sys.argv) > 2 else 800
dev = torch.device('cuda:0')
arena = DecoderArena(make_state_dict(decoder_specs(24), 0), dev, n_layers=24)
n = 9 * F + 2
gen = Generator(arena, 1, 257 + n)
p = synthetic_prefix(1, 0).to(dev)
ref = None
bad = 0
for i in range(runs):
    torch.cuda.synchronize(); t = time.time()
    ids, _ = gen.generate(p, n, flags=capi.GEN_NO_EARLY_EXIT)
    torch.cuda.synchronize(); dt = time.time() - t
    e = gen.mega_error()
    same = True if ref is None else bool(torch.equal(ids, ref))
    if ref is None:
        ref = ids.clone()
    flag = '' if (e == 0 and same) else '  <-- PROBLEM'
    if flag: bad += 1
    print(f'run {i:3d}: {dt*1000:8.1f} ms  error={e} same_ids={same}{flag}', flush=True)
print('problems:', bad)
