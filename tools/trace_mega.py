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
n = int(sys.argv[1]) if len(sys.argv) > 1 else 6
gen = Generator(arena, 1, 257 + n + 8)
p = synthetic_prefix(1, 0).to(dev)
for _ in range(2):
    gen.generate(p, n, flags=capi.GEN_NO_EARLY_EXIT)
gen.generate(p, n, flags=capi.GEN_NO_EARLY_EXIT | capi.GEN_TRACE)
torch.cuda.synchronize()
print('error flag', gen.mega_error())
tr = gen.mega_trace(1280)
# per layer stamps of CTA 0 (decode_mega.cu STAMP()): 10 phase ends
per_layer = 10
names = ['x in (LN2+wait yb)', 'qkv', 'attn items', 'merge (wait parts)', 'out partial', 'reduce A', 'LN1 (wait ya)',
         'fc1', 'fc2 partial (wait f)', 'reduce B']
step0 = 0
# trace the LAST complete step of the launch: find how many steps fit
per_step = 1 + NL * per_layer + 1
nsteps = sum(1 for i in range(0, 1270 - per_step, per_step) if tr[i + per_step - 1])
i0 = (nsteps - 1) * per_step
agg = collections.defaultdict(list)
i = i0 + 1
for L in range(NL):
    for k in range(per_layer):
        agg[names[k]].append((tr[i] - tr[i - 1]) / 1000.0)
        i += 1
tot = 0.0
for k in names:
    v = agg[k]
    tot += sum(v) / len(v)
    print(f'{k:22s} avg {sum(v)/len(v):6.2f} us  min {min(v):6.2f} max {max(v):6.2f}')
print(f'layer total {tot:.2f} us; lm+pick {(tr[i] - tr[i - 1]) / 1000.0:.2f} us; step total {(tr[i] - tr[i0]) / 1000.0:.1f} us (step {nsteps - 1} of the launch)')

# per-CTA skew at (step 1, layer 12): time each CTA reaches the end of each phase, relative to the earliest CTA
tc = gen.mega_trace_cta(144)
import statistics
base = min(r[0] for r in tc if r[0])
names2 = names + ['LN1: ya polled', 'fc1: A weights ready', 'fc2: f polled', 'fc2: B weights ready', 'x: yb polled (this layer)', 'merge: parts polled']
order = [14, 0, 1, 2, 15, 3, 4, 5, 10, 6, 11, 7, 12, 13, 8, 9]
for k in order:
    lab = names2[k]
    col = [(r[k] - base) / 1000.0 for r in tc if r[k]]
    srt = sorted(range(len(col)), key=lambda i: col[i])
    print(f'{lab:22s} min {min(col):6.2f} med {statistics.median(col):6.2f} max {max(col):6.2f} us   slowest CTAs {srt[-3:]} fastest {srt[:3]}')
