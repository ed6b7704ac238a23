#!/usr/bin/env python
"""Join an ncu SASS source page (csv) with nvdisasm line info -> per-source-line profile.

usage: ncu_lines.py REPORT.ncu-rep LIB.so KERNEL_SUBSTRING [min_pct]
"""
import csv
import re
import subprocess
import sys
import tempfile
import os
from collections import defaultdict


def main():
    rep, so, kname = sys.argv[1:4]
    min_pct = float(sys.argv[4]) if len(sys.argv) > 4 else 0.7
    tmp = tempfile.mkdtemp()
    subprocess.run(['cuobjdump', '-xelf', 'all', os.path.abspath(so)], cwd=tmp, capture_output=True)
    cubin = [f for f in os.listdir(tmp) if f.endswith('.cubin')][0]
    dis = subprocess.run(['nvdisasm', '-gi', '-c', os.path.join(tmp, cubin)], capture_output=True, text=True).stdout
    # map offset -> line for the kernel
    line_of = {}
    cur = None
    in_fn = False
    for ln in dis.splitlines():
        if ln.startswith('\t.section') or ln.startswith('//-----'):
            in_fn = ('.text.' in ln and kname in ln)
        if not in_fn:
            continue
        m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
        if m:
            cur = (os.path.basename(m.group(1)), int(m.group(2)))
            if 'inlined at' in ln:
                m2 = re.findall(r'line (\d+)', ln)
            continue
        m = re.match(r'\s*/\*([0-9a-f]+)\*/\s+(.*?);', ln)
        if m:
            line_of[int(m.group(1), 16)] = (cur, m.group(2).strip())
    out = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    # first kernel matching
    start = [i for i, r in enumerate(rows) if r and r[0] == 'Kernel Name']
    hdr = rows[start[0] + 1]
    ix = {h: i for i, h in enumerate(hdr)}
    body = []
    for r in rows[start[0] + 2:]:
        if not r or r[0] == 'Kernel Name':
            break
        body.append(r)
    base = int(body[0][0], 16)
    agg = defaultdict(lambda: [0.0, 0.0, 0.0, defaultdict(float)])
    stall_cols = [h for h in hdr if h.startswith('stall_') and 'Not Issued' not in h]
    tot_s = tot_i = 0.0
    for r in body:
        off = int(r[0], 16) - base
        key, _ = line_of.get(off, (None, ''))
        s = float(r[ix['# Samples']] or 0)
        ins = float(r[ix['Instructions Executed']] or 0)
        thr = float(r[ix['Thread Instructions Executed']] or 0)
        a = agg[key]
        a[0] += s; a[1] += ins; a[2] += thr
        for h in stall_cols:
            a[3][h] += float(r[ix[h]] or 0)
        tot_s += s; tot_i += ins
    print('total samples %.0f, warp instructions %.0f' % (tot_s, tot_i))
    src_cache = {}
    for key, a in sorted(agg.items(), key=lambda kv: -kv[1][0]):
        if a[0] / tot_s * 100 < min_pct and a[1] / tot_i * 100 < min_pct:
            continue
        text = ''
        if key:
            fn = key[0]
            if fn not in src_cache:
                for root in (os.path.dirname(os.path.abspath(so)) + '/../csrc', '.'):
                    p = os.path.join(root, fn)
                    if os.path.exists(p):
                        src_cache[fn] = open(p).read().splitlines()
                        break
                else:
                    src_cache[fn] = []
            L = src_cache[fn]
            text = L[key[1] - 1].strip()[:90] if 0 < key[1] <= len(L) else ''
        top = sorted(a[3].items(), key=lambda kv: -kv[1])[:3]
        print('%-22s %5.1f%% smp %5.1f%% inst thr/inst %4.1f  [%s] %s' % (
            '%s:%d' % key if key else '?', 100 * a[0] / tot_s, 100 * a[1] / tot_i, a[2] / max(a[1], 1),
            ' '.join('%s %.0f%%' % (h[6:], 100 * v / max(a[0], 1)) for h, v in top), text))


if __name__ == '__main__':
    main()
