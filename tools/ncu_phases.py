#!/usr/bin/env python
"""Aggregate an ncu source-page profile by source line ranges of one file.
usage: ncu_phases.py REPORT LIB KERNEL FILE name:lo-hi[,lo-hi] ...
Instructions inlined from other files (or from FILE lines outside every range, e.g. small
helpers) inherit the phase of the closest preceding instruction that has one."""
import csv, re, subprocess, sys, tempfile, os
from collections import defaultdict

def main():
    rep, so, kname, src = sys.argv[1:5]
    ranges = []
    for a in sys.argv[5:]:
        nm, rs = a.split(':')
        ranges.append((nm, [tuple(int(v) for v in r.split('-')) for r in rs.split(',')]))
    tmp = tempfile.mkdtemp()
    subprocess.run(['cuobjdump', '-xelf', 'all', os.path.abspath(so)], cwd=tmp, capture_output=True)
    cubin = [f for f in os.listdir(tmp) if f.endswith('.cubin')][0]
    dis = subprocess.run(['nvdisasm', '-gi', '-c', os.path.join(tmp, cubin)], capture_output=True, text=True).stdout
    line_of, cur, in_fn = {}, None, False
    for ln in dis.splitlines():
        if ln.startswith('\t.section') or ln.startswith('//-----'):
            in_fn = ('.text.' in ln and kname in ln)
        if not in_fn:
            continue
        m = re.search(r'//## File "([^"]+)", line (\d+)(.*)', ln)
        if m:
            fn, l = os.path.basename(m.group(1)), int(m.group(2))
            # for inlined library code use the outermost "inlined at" kernels.cuh line
            m2 = re.findall(r'File "([^"]+)", line (\d+)', ln)
            cur = (fn, l)
            # the kernel body is the last function of its file: of all lines of FILE on the inlining
            # chain take the largest one
            mine = [int(l2) for f2, l2 in m2 if os.path.basename(f2) == src]
            if mine:
                cur = (src, max(mine))
            continue
        m = re.match(r'\s*/\*([0-9a-f]+)\*/\s+(.*?);', ln)
        if m:
            line_of[int(m.group(1), 16)] = (cur, m.group(2).strip())
    out = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    start = [i for i, r in enumerate(rows) if r and r[0] == 'Kernel Name'][0]
    hdr = rows[start + 1]
    ix = {h: i for i, h in enumerate(hdr)}
    body = [r for r in rows[start + 2:] if r and r[0] != 'Kernel Name']
    base = int(body[0][0], 16)
    agg = defaultdict(lambda: [0.0, 0.0, 0.0, 0.0, 0.0])
    stalls = defaultdict(lambda: defaultdict(float))
    scols = [h for h in hdr if h.startswith('stall_') and 'Not Issued' not in h]
    tot = [0.0, 0.0]
    prev = 'other'
    for r in body:
        off = int(r[0], 16) - base
        key, sass = line_of.get(off, (None, ''))
        s = float(r[ix['# Samples']] or 0); ins = float(r[ix['Instructions Executed']] or 0)
        thr = float(r[ix['Thread Instructions Executed']] or 0)
        name = None
        if key and key[0] == src:
            for nm, rs in ranges:
                if any(lo <= key[1] <= hi for lo, hi in rs):
                    name = nm
                    break
        if name is None:
            name = prev
        prev = name
        a = agg[name]
        for h in scols:
            stalls[name][h] += float(r[ix[h]] or 0)
        a[0] += s; a[1] += ins; a[2] += thr
        op = sass.split()[0] if sass else ''
        if op.startswith('@'):
            op = sass.split()[1]
        if op[:2] in ('DF', 'DA', 'DM', 'DS') or op.startswith('DSETP'):
            a[3] += ins
        if op.startswith(('LDS', 'STS', 'LDG', 'STG', 'LD.', 'ST.', 'LDC')):
            a[4] += ins
        tot[0] += s; tot[1] += ins
    print('%-14s %8s %8s %8s %8s %8s' % ('phase', 'smp%', 'inst%', 'thr/inst', 'fp64%', 'mem%'))
    for nm, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print('%-14s %8.1f %8.1f %8.1f %8.1f %8.1f' % (nm, 100 * a[0] / tot[0], 100 * a[1] / tot[1],
              a[2] / max(a[1], 1), 100 * a[3] / max(a[1], 1), 100 * a[4] / max(a[1], 1)))
    print('total warp inst %.3e' % tot[1])
    for nm, a in sorted(agg.items(), key=lambda kv: -kv[1][0]):
        st = sorted(stalls[nm].items(), key=lambda kv: -kv[1])[:5]
        tt = sum(stalls[nm].values()) or 1
        print('%-10s' % nm, '  '.join('%s %.0f%%' % (k[6:], 100 * v / tt) for k, v in st))

if __name__ == '__main__':
    main()
