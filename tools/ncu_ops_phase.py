"""Development: dynamic opcode histogram of a source-line range (outermost line of jac6.cuh) from an ncu source csv
and the matching nvdisasm -gi listing.  usage: ncu_ops_phase.py SRC.csv LISTING.txt LO HI [inner_file:lo-hi]"""
import collections, csv, re, sys
src_csv, listing, lo, hi = sys.argv[1], sys.argv[2], int(sys.argv[3]), int(sys.argv[4])
lines = open(listing).read().splitlines()
cur = None; out = []
for ln in lines:
    m = re.search(r'//## File "([^"]+)", line (\d+)(.*)', ln)
    if m:
        m2 = re.findall(r'File "([^"]+)", line (\d+)', ln)
        mine = [int(l) for f, l in m2 if f.endswith('jac6.cuh')]
        cur = (max(mine) if mine else None, m.group(1).split('/')[-1], int(m.group(2)))
        continue
    m = re.match(r'\s*/\*([0-9a-f]+)\*/\s+(.*?);', ln)
    if m:
        out.append((int(m.group(1), 16), cur, m.group(2)))
rows = list(csv.reader(open(src_csv)))
start = [i for i, r in enumerate(rows) if r and r[0] == 'Kernel Name'][0]
hdr = rows[start + 1]; ix = {h: i for i, h in enumerate(hdr)}
body = [r for r in rows[start + 2:] if r and r[0] != 'Kernel Name']
base = int(body[0][0], 16)
cnt = {int(r[0], 16) - base: float(r[ix['Instructions Executed']] or 0) for r in body}
ops = collections.Counter(); inner = collections.Counter(); tot = 0
for a, c, s in out:
    if c and c[0] and lo <= c[0] <= hi:
        n = cnt.get(a, 0)
        op = s.split()[0]
        if op.startswith('@'):
            op = s.split()[1]
        ops[op.split('.')[0]] += n
        inner[(c[1], c[2])] += n
        tot += n
print('total %.1fM' % (tot / 1e6))
print(' '.join('%s:%.1f' % (k, v / 1e6) for k, v in ops.most_common(40)))
print('by innermost line:')
for (f, l), n in inner.most_common(45):
    print('  %s:%d %.1fM' % (f, l, n / 1e6))
# per-phase shared-memory wavefronts and global tag requests (optional: pass 'wf' as 5th arg)
if len(sys.argv) > 5 and sys.argv[5] == 'wf':
    wf = {int(r[0], 16) - base: (float(r[ix['L1 Wavefronts Shared']] or 0), float(r[ix['L1 Wavefronts Shared Ideal']] or 0),
                                 float(r[ix['L1 Tag Requests Global']] or 0)) for r in body}
    tw = ti = tg = 0
    for a, c, s in out:
        if c and c[0] and lo <= c[0] <= hi:
            w_, i_, g_ = wf.get(a, (0, 0, 0)); tw += w_; ti += i_; tg += g_
    print('shared wavefronts %.1fM (ideal %.1fM), global tag requests %.1fM' % (tw / 1e6, ti / 1e6, tg / 1e6))
