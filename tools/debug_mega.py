"""Post-mortem of the persistent decode kernel: run a few generates; on a time-out print where every CTA stopped.
python tools/debug_mega.py [layers] [new_tokens] [repeats]"""
import os, sys, collections, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from meshanything_b200 import capi
from meshanything_b200.checkpoint import decoder_specs, make_state_dict
from meshanything_b200.decoder import DecoderArena, Generator
from bench import synthetic_prefix
NL = int(sys.argv[1]) if len(sys.argv) > 1 else 24
n = int(sys.argv[2]) if len(sys.argv) > 2 else 60
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 3
dev = torch.device('cuda:0')
arena = DecoderArena(make_state_dict(decoder_specs(NL), 0), dev, n_layers=NL)
capi.lib().ma_mega_set_debug(200_000_000, 0)   # 0.2 s per wait
gen = Generator(arena, 1, 257 + n)
p = synthetic_prefix(1, 0).to(dev)
names = ['x in', 'qkv', 'attn', 'merge', 'out', 'redA', 'LN1', 'fc1', 'fc2', 'redB']
for rep in range(reps):
    ids, lens = gen.generate(p, n, flags=capi.GEN_NO_EARLY_EXIT | capi.GEN_WHERE)
    torch.cuda.synchronize()
    err = gen.mega_error()
    print(f'rep {rep}: lens {int(lens[0])} error code {err & 0xff} cta {err >> 8}')
    if err:
        f = gen.mega_fail()
        print('  first time-out:', f)
        wh = gen.mega_where()
        wp = gen.mega_wprog()
        # the CTA(s) that are furthest behind
        key = lambda w: (w[0], w[1], w[2])
        slow = sorted(range(144), key=lambda c: key(wh[c]))[:3]
        for c in slow:
            print(f'  slowest CTA {c}: where (step {wh[c][0]-1}, L {wh[c][1]}, {names[wh[c][2]]}); per-warp marks (epoch, mark: 1-6 in x-in/LN (+16 = ln2), 7-12 merge):',
                  [(v >> 8, v & 255) for v in wp[c]])
        wh = gen.mega_where()
        cnt = collections.Counter((w[0] - 1, w[1], names[w[2]] if 0 <= w[2] < 10 else w[2]) for w in wh)
        for k, v in sorted(cnt.items()):
            ctas = [c for c, w in enumerate(wh) if (w[0] - 1, w[1], names[w[2]] if 0 <= w[2] < 10 else w[2]) == k]
            print('  last phase passed', k, 'x', v, 'CTAs', ctas[:24], '...' if len(ctas) > 24 else '')
        break
