"""Opcode histogram of the sm_100a SASS of every translation unit (cuobjdump -sass on the
objects of spinterps_b200/lib): which instructions the shipped kernels actually contain --
DMMA (FP64 tensor), UBLKCP (cp.async.bulk), SYNCS (mbarrier), STG...EF (streaming stores) ...
    python scripts/sass_histogram.py  ->  profiles/r2_sass_<unit>.txt"""
import collections
import re
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
LIB = ROOT / 'spinterps_b200' / 'lib'
OUT = ROOT / 'profiles'
tag = sys.argv[1] if len(sys.argv) > 1 else 'r2'
for obj in sorted(LIB.glob('spx_*.o')):
    sass = subprocess.run(['cuobjdump', '-sass', str(obj)], capture_output=True, text=True).stdout
    per_fn = collections.OrderedDict()
    fn = None
    for line in sass.splitlines():
        m = re.match(r'\s*Function : (\S+)', line)
        if m:
            fn = subprocess.run(['c++filt', m.group(1)], capture_output=True, text=True).stdout.strip()
            fn = re.sub(r'\(.*', '', fn)[:90]
            per_fn[fn] = collections.Counter()
            continue
        m = re.match(r'\s*/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Za-z0-9_.]*)', line)
        if m and fn:
            per_fn[fn][m.group(1)] += 1
    total = collections.Counter()
    for c in per_fn.values():
        total.update(c)
    lines = ['# %s: sm_100a SASS opcode histogram (cuobjdump -sass), %d kernels, %d instructions'
             % (obj.name, len(per_fn), sum(total.values())), '']
    key = ['DMMA', 'UBLKCP', 'UTMALDG', 'UTMASTG', 'SYNCS', 'STG.E.EF', 'REDUX', 'LDGSTS', 'DFMA', 'DADD',
           'DMUL', 'MUFU', 'SHFL', 'BAR', 'ATOM', 'RED', 'LDS', 'STS']
    lines.append('## selected opcode families, whole unit')
    for k in key:
        n = sum(v for op, v in total.items() if op.startswith(k))
        if n:
            lines.append('%-10s %7d' % (k, n))
    lines.append('')
    lines.append('## per kernel: instructions, then the families above')
    for fn, c in per_fn.items():
        fam = ', '.join('%s %d' % (k, sum(v for op, v in c.items() if op.startswith(k)))
                        for k in key if sum(v for op, v in c.items() if op.startswith(k)))
        lines.append('%-92s %6d  %s' % (fn, sum(c.values()), fam))
    lines.append('')
    lines.append('## top 25 opcodes, whole unit')
    for op, n in total.most_common(25):
        lines.append('%-28s %7d' % (op, n))
    (OUT / ('%s_sass_%s.txt' % (tag, obj.stem))).write_text('\n'.join(lines) + '\n')
    print(obj.name, len(per_fn), 'kernels')
