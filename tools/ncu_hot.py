#!/usr/bin/env python
"""Top SASS instructions by stall samples, with source line and dominant stall reason.
usage: ncu_hot.py REPORT LIB KERNEL [top]"""
import csv, re, subprocess, sys, tempfile, os

def main():
    rep, so, kname = sys.argv[1:4]
    top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
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
        m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
        if m:
            cur = (os.path.basename(m.group(1)), int(m.group(2)))
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
    scols = [h for h in hdr if h.startswith('stall_') and 'Not Issued' not in h]
    tot = sum(float(r[ix['# Samples']] or 0) for r in body)
    items = []
    for r in body:
        off = int(r[0], 16) - base
        key, sass = line_of.get(off, (None, ''))
        s = float(r[ix['# Samples']] or 0)
        st = max(scols, key=lambda h: float(r[ix[h]] or 0))
        items.append((s, off, key, sass, st[6:], float(r[ix['Instructions Executed']] or 0)))
    items.sort(reverse=True)
    for s, off, key, sass, st, ins in items[:top]:
        print('%5.2f%%  %06x  %-22s %-12s exec %.2e  %s' % (100 * s / tot, off, '%s:%d' % key if key else '?', st, ins, sass[:70]))

if __name__ == '__main__':
    main()
