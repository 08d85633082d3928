"""Phase timeline of the persistent decode kernel (CTA 0, globaltimer): python tools/trace_mega.py [new_tokens]"""
import sys, torch, collections
import os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from meshanything_b200 import capi
from meshanything_b200.checkpoint import decoder_specs, make_state_dict
from meshanything_b200.decoder import DecoderArena, Generator
from bench import synthetic_prefix
NL = 24
dev = torch.device('cuda:0')
arena = DecoderArena(make_state_dict(decoder_specs(NL), 0), dev, n_layers=NL)
n = int(sys.argv[1]) if len(sys.argv) > 1 else 30
gen = Generator(arena, 1, 257 + n + 8)
p = synthetic_prefix(1, 0).to(dev)
for _ in range(2):
    gen.generate(p, n, flags=capi.GEN_NO_EARLY_EXIT)
gen.generate(p, n, flags=capi.GEN_NO_EARLY_EXIT | capi.GEN_TRACE)
torch.cuda.synchronize()
print('error flag', gen.mega_error())
tr = gen.mega_trace(1280)
# per layer stamps: [qkv: prologue_done, weights_ready, compute+refill done, barrier done], attn, out, fc1, fc2 (1 each)
per_layer = 10
names = ['qkv.prologue', 'qkv.wait', 'qkv.gemv', 'qkv.sync', 'qkv.emit+refill', 'attn.items', 'attn.merge', 'out', 'fc1', 'fc2']
agg = collections.defaultdict(list)
i = 1
for L in range(NL):
    for k in range(per_layer):
        agg[names[k]].append((tr[i] - tr[i - 1]) / 1000.0)
        i += 1
for k in names:
    v = agg[k]
    print(f'{k:14s} avg {sum(v)/len(v):6.2f} us  min {min(v):6.2f} max {max(v):6.2f}')
print('lm+pick', (tr[i] - tr[i - 1]) / 1000.0, 'us; step total', (tr[i] - tr[0]) / 1000.0, 'us')

# per-CTA skew at (step 1, layer 12): time each CTA reaches the end of each phase, relative to the earliest CTA
tc = gen.mega_trace_cta(147)
import statistics
labels = ['x ready(qkv in)', 'qkv done', 'attn items done', 'attn merge done', 'out done', 'fc1 done', 'fc2 done']
base = min(r[0] for r in tc if r[0])
for k, lab in enumerate(labels):
    col = [(r[k] - base) / 1000.0 for r in tc if r[k]]
    srt = sorted(range(len(col)), key=lambda i: col[i])
    print(f'{lab:18s} min {min(col):6.2f} med {statistics.median(col):6.2f} max {max(col):6.2f} us   slowest CTAs {srt[-3:]} fastest {srt[:3]}')
