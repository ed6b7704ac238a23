"""Development: SASS of one kernel of a built library (substring match on the mangled name) and its
instruction mix.  usage: python tools/sass_fn.py lib.so k_jac6ILi8ELi384 [out.sass]"""
import collections
import re
import subprocess
import sys


def extract(lib, pat):
    txt = subprocess.run(['cuobjdump', '-sass', lib], capture_output=True, text=True).stdout
    out, on = [], False
    for line in txt.splitlines():
        if 'Function :' in line:
            on = pat in line
        if on:
            out.append(line)
    return out


def mix(lines):
    c = collections.Counter()
    for ln in lines:
        m = re.search(r'\*/\s+(?:@!?U?P\d\s+)?([A-Z][A-Z0-9_]*)', ln)
        if m:
            c[m.group(1)] += 1
    return c


if __name__ == '__main__':
    lines = extract(sys.argv[1], sys.argv[2])
    if len(sys.argv) > 3:
        open(sys.argv[3], 'w').write('\n'.join(lines) + '\n')
    c = mix(lines)
    print('instructions', sum(c.values()))
    print(' '.join('%s:%d' % kv for kv in c.most_common(45)))
