"""Fine phase timeline of the persistent decode kernel (CTA 0, thread 0): python tools/trace_fine.py [new_tokens]"""
import sys, torch, collections, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from meshanything_b200 import capi
from meshanything_b200.checkpoint import decoder_specs, make_state_dict
from meshanything_b200.decoder import DecoderArena, Generator
from bench import synthetic_prefix
NL = 24
dev = torch.device('cuda:0')
arena = DecoderArena(make_state_dict(decoder_specs(NL), 0), dev, n_layers=NL)
n = int(sys.argv[1]) if len(sys.argv) > 1 else 4
gen = Generator(arena, 1, 257 + n + 8)
p = synthetic_prefix(1, 0).to(dev)
for _ in range(2):
    gen.generate(p, n, flags=capi.GEN_NO_EARLY_EXIT)
gen.generate(p, n, flags=capi.GEN_NO_EARLY_EXIT | capi.GEN_TRACE_FINE)
torch.cuda.synchronize()
print('error flag', gen.mega_error())
tr = gen.mega_trace(1280)
# layer 0 has no LN2 (x in = embedding): 2 stamps fewer
names0 = ['embed']
namesL = ['x: yb polled', 'x: ln2 params', 'x: LN2 done']
rest = ['qkv: weights', 'qkv: published', 'attn done', 'merge: parts polled', 'merge done', 'out: weights', 'out: published',
        'reduce A', 'LN1: ya polled', 'LN1: params', 'LN1 done', 'fc1: weights', 'fc1: published', 'fc2: f polled',
        'fc2: weights', 'fc2: published', 'reduce B']
per_step = 1 + (len(names0) + len(rest)) + (NL - 1) * (len(namesL) + len(rest)) + 3   # lm phase: 2 fine stamps + 1
step = 1 if tr[2 * per_step - 1] else 0
i = step * per_step + 1
agg = collections.defaultdict(list)
for L in range(NL):
    for k in (names0 if L == 0 else namesL) + rest:
        agg[k].append((tr[i] - tr[i - 1]) / 1000.0)
        i += 1
tot = 0.0
for k in namesL + rest:
    v = agg[k]
    tot += sum(v) / len(v)
    print(f'{k:22s} avg {sum(v)/len(v):6.2f} us  min {min(v):6.2f} max {max(v):6.2f}')
print(f'layer total {tot:.2f} us; lm+pick {(tr[i] - tr[i - 1]) / 1000.0:.2f} us; step {step}')
