"""Development: static count of selected SASS opcodes per source line (outermost line of FILE on the inlining chain).
usage: python tools/sass_lines.py lib.so KERNEL_SUBSTR FILE OPCODE[,OPCODE...]"""
import collections, os, re, subprocess, sys, tempfile
lib, kname, src, ops = sys.argv[1:5]
ops = tuple(ops.split(','))
tmp = tempfile.mkdtemp()
subprocess.run(['cuobjdump', '-xelf', 'all', os.path.abspath(lib)], cwd=tmp, capture_output=True)
cubin = [f for f in os.listdir(tmp) if f.endswith('.cubin')][0]
dis = subprocess.run(['nvdisasm', '-gi', '-c', os.path.join(tmp, cubin)], capture_output=True, text=True).stdout
cur, on = None, False
cnt = collections.Counter()
tot = collections.Counter()
for ln in dis.splitlines():
    if ln.startswith('\t.section') or ln.startswith('//-----'):
        on = '.text.' in ln and kname in ln
    if not on:
        continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
    if m:
        mine = [int(l) for f, l in re.findall(r'File "([^"]+)", line (\d+)', ln) if os.path.basename(f) == src]
        cur = max(mine) if mine else None
        continue
    m = re.match(r'\s*/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)', ln)
    if m:
        tot[cur] += 1
        if m.group(1).startswith(ops):
            cnt[cur] += 1
for k in sorted(cnt, key=lambda k: (k is None, k)):
    print('line %s: %d of %d' % (k, cnt[k], tot[k]))
