"""Summarise a YOLACT_B200_PROFILE_DUMP per-layer CSV: time share and algorithmic TFLOP/s by layer shape."""
import collections, csv, sys
rows = list(csv.DictReader(open(sys.argv[1])))
agg = collections.OrderedDict()
for r in rows:
    a = agg.setdefault(int(r['op']), dict(r, ms=0.0, n=0))
    a['ms'] += float(r['ms']); a['n'] += 1
KIND = {0: 'stem(s2d+conv_tc)', 1: 'maxpool', 2: 'phase_split', 4: 'upsample_add', 5: 'upsample2x', 6: 'head_finalize'}
groups = collections.OrderedDict()
for a in agg.values():
    ms = a['ms'] / a['n']
    kind = int(a['kind'])
    if kind == 11:                                                     # OP_BNECK: conv3 (cin -> cout) + residual chained into the next conv1 (cout -> cin)
        key = 'fused 1x1 %4s->%4s + res, 1x1 ->%4s @%3s' % (a['cin'], a['cout'], a['cin'], a['h_out'])
    elif kind == 3:
        key = 'conv %sx%s s%s %4s->%4s @%3s' % (a['k'], a['k'], a['stride'], a['cin'], a['cout'], a['h_out'])
    else:
        key = KIND.get(kind, 'kind %d' % kind)
    g = groups.setdefault(key, dict(ms=0.0, n=0, gf=0.0, batch=a['batch']))
    g['ms'] += ms; g['n'] += 1; g['gf'] += float(a['gflop'])
tot = sum(g['ms'] for g in groups.values())
print(f'forward total {tot:.3f} ms over {len(agg)} layers')
for key, g in sorted(groups.items(), key=lambda kv: -kv[1]['ms']):
    print('%-42s n=%2d  %7.3f ms  %5.1f%%  %7.0f TFLOP/s' % (key, g['n'], g['ms'], 100 * g['ms'] / tot, g['gf'] / g['ms'] if g['ms'] else 0))
