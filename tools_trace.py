"""Phase timeline of the persistent decode kernel (CTA 0, globaltimer): python tools_trace.py [faces]"""
import sys, torch
sys.path.insert(0, '.')
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
tr = gen.mega_trace(150)
t0 = tr[0]
names = ['qkv', 'attn', 'out', 'fc1', 'fc2']
# stamps: step start, then 5 per layer, then after pick
d = [(tr[i + 1] - tr[i]) / 1000.0 for i in range(0, 1 + 5 * NL)]
import collections
agg = collections.defaultdict(list)
for i in range(5 * NL):
    agg[names[i % 5]].append(d[i])
for k in names:
    v = agg[k]
    print(f'{k:5s} avg {sum(v)/len(v):6.2f} us  min {min(v):6.2f} max {max(v):6.2f}')
print('lm+pick', d[5 * NL], 'us; step total', (tr[1 + 5 * NL] - tr[0]) / 1000.0, 'us')
print('second step total', (tr[2 * (1 + 5 * NL) ] - tr[1 + 5 * NL]) / 1000.0 if len(tr) > 2 * (1 + 5 * NL) else None)
