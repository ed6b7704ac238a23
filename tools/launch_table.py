#!/usr/bin/env python
"""profiles/rNN_launches.md from the ncu launch list of the bench command.
usage: launch_table.py gpurun_out/r02_launches.csv > profiles/r02_launches.md"""
import collections
import csv
import sys

rows = [r for r in csv.reader(open(sys.argv[1])) if r]
hi = [i for i, r in enumerate(rows) if r[0] == 'ID'][0]
h = rows[hi]
ix = {k: i for i, k in enumerate(h)}
L = []
for r in rows[hi + 2:]:
    if len(r) < len(h):
        continue
    L.append((r[ix['Kernel Name']], r[ix['Grid Size']], r[ix['Block Size']], float(r[ix['Metric Value']].replace(',', '')) / 1e6))
print("# Round 2: launch list of the bench command\n")
print("`ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv python bench.py --steps 2 --warmup 1` on one B200")
print("(tools/round2_end.sh; %d launches captured; per-launch times under ncu are serialised and cold-cache, so only the" % len(L))
print("shares are meaningful -- the bench's own numbers come from CUDA events).\n")
big = [l for l in L if 'k_eval<8, 384, 0, 0>' in l[0] and l[3] > 10]
print("**Timed region of `value`** (W = 1 warm-up + K = 2 steps): %d launches of `pj5::k_eval<8, 384, 0, 0>` over 2^20 states," % len(big))
print("%s ms each -- one kernel per step, `gpu_launches` = K. Nothing else runs between the events.\n" % ', '.join('%.2f' % l[3] for l in big))
agg = collections.OrderedDict()
for k, g, b, ms in L:
    a = agg.setdefault((k, g, b), [0, 0.0])
    a[0] += 1
    a[1] += ms
tot = sum(v[1] for v in agg.values())
print("Whole command (the step above, the end-to-end legs through the host API in 256 MB chunks, the factored / consumer side records,")
print("the three side workloads):\n")
print("| kernel | grid | block | launches | total ms | share |\n|---|---|---|---|---|---|")
what = {'k_eval<8, 384, 0, 0>': 'eval_jacob, GRI-shaped', 'k_eval<8, 384, 3, 0>': 'factored record, GRI-shaped',
        'k_eval<16, 512, 0, 1>': 'eval_jacob, n-heptane-shaped, working set in global memory',
        'k_eval<8, 384, 0, 1>': 'eval_jacob, USC-II-shaped, working set in global memory',
        'k_eval<32, 384, 0, 0>': 'eval_jacob, H2/O2 PaSR states', 'k_newton': 'x = (I - gamma J)^-1 r'}
other = [0, 0.0]
for (k, g, b), v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    if k.startswith('void at::'):
        other[0] += v[0]
        other[1] += v[1]
        continue
    w = [v2 for k2, v2 in what.items() if k2 in k]
    print("| `%s` %s | %s | %s | %d | %.2f | %.1f %% |" % (k.split('(')[0], '(' + w[0] + ')' if w else '', g, b, v[0], v[1], 100 * v[1] / tot))
print("| torch copy / fill / compare kernels of the bench's own checks | | | %d | %.2f | %.1f %% |" % (other[0], other[1], 100 * other[1] / tot))
print("\nEvery compute kernel in the list is this repository's (`pj5::k_eval`, `pjc::k_newton`); the torch kernels are the bench's")
print("buffer fills and its device-side equality checks, outside every timed region.")
