#!/usr/bin/env python
"""Opcode histogram (weighted by executed warp instructions) of one source line range.
usage: ncu_ops.py REPORT LIB KERNEL FILE lo-hi [top]"""
import csv, re, subprocess, sys, tempfile, os
from collections import defaultdict

def main():
    rep, so, kname, src, rng = sys.argv[1:6]
    top = int(sys.argv[6]) if len(sys.argv) > 6 else 25
    lo, hi = (int(v) for v in rng.split('-'))
    tmp = tempfile.mkdtemp()
    subprocess.run(['cuobjdump', '-xelf', 'all', os.path.abspath(so)], cwd=tmp, capture_output=True)
    cubin = [f for f in os.listdir(tmp) if f.endswith('.cubin')][0]
    dis = subprocess.run(['nvdisasm', '-g', '-c', os.path.join(tmp, cubin)], capture_output=True, text=True).stdout
    line_of, cur, in_fn = {}, None, False
    for ln in dis.splitlines():
        if ln.startswith('\t.section') or ln.startswith('//-----'):
            in_fn = ('.text.' in ln and kname in ln)
        if not in_fn:
            continue
        m = re.search(r'//## File "([^"]+)", line (\d+)(.*)', ln)
        if m:
            cur = (os.path.basename(m.group(1)), int(m.group(2)))
            for f2, l2 in re.findall(r'File "([^"]+)", line (\d+)', ln):
                if os.path.basename(f2) == src:
                    cur = (src, int(l2))
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
    ops, lines, tot = defaultdict(float), defaultdict(float), 0.0
    for r in body:
        off = int(r[0], 16) - base
        key, sass = line_of.get(off, (None, ''))
        if not key or key[0] != src or not (lo <= key[1] <= hi):
            continue
        ins = float(r[ix['Instructions Executed']] or 0)
        t = sass.split()
        op = t[1] if t and t[0].startswith('@') else (t[0] if t else '')
        ops[op.split('.')[0] + ('.' + op.split('.')[1] if '.' in op else '')] += ins
        lines[key[1]] += ins
        tot += ins
    print('total %.3e' % tot)
    for op, v in sorted(ops.items(), key=lambda kv: -kv[1])[:top]:
        print('  %-16s %6.2f%%' % (op, 100 * v / tot))
    print('by line:')
    for l, v in sorted(lines.items(), key=lambda kv: -kv[1])[:top]:
        print('  line %4d %6.2f%%' % (l, 100 * v / tot))

if __name__ == '__main__':
    main()
