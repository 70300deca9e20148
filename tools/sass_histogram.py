#!/usr/bin/env python
"""Opcode histogram per kernel of the shipped library (cuobjdump -sass): the Blackwell-native evidence (UTCHMMA = tcgen05.mma,
LDTM = tcgen05.ld, UTMALDG / UTMASTG = TMA, UTCBAR = tcgen05.commit, SYNCS = mbarrier, LDGSTS = cp.async, HMMA = legacy mma.sync).

    python tools/sass_histogram.py [path/to/libppyolo_b200.so] > profiles/r02_sass_opcodes.txt
"""
import collections
import os
import re
import subprocess
import sys

REPO = os.path.join(os.path.dirname(os.path.abspath(__file__)), '..')
lib = sys.argv[1] if len(sys.argv) > 1 else os.path.join(REPO, 'pytorch-ppyolo_b200', 'ppyolo_b200', 'libppyolo_b200.so')
KEYS = ('UTCHMMA', 'UTCQMMA', 'UTCMMA', 'LDTM', 'STTM', 'UTMALDG', 'UTMASTG', 'UTMAPF', 'UBLKCP', 'UTCBAR', 'UTCATOM', 'SYNCS', 'LDGSTS', 'HMMA',
        'FFMA', 'LDG', 'STG', 'LDS', 'STS', 'REDG', 'RED', 'ATOMG', 'ATOMS', 'ATOM', 'MUFU', 'SHFL', 'BAR', 'UCGABAR', 'ELECT', 'LDL', 'STL')
sass = subprocess.run(['cuobjdump', '-sass', lib], capture_output=True, text=True).stdout
demangle = lambda n: subprocess.run(['c++filt', n], capture_output=True, text=True).stdout.strip()
kernels = collections.OrderedDict()
cur = None
for line in sass.splitlines():
    m = re.search(r'Function : (\S+)', line)
    if m:
        cur = kernels.setdefault(m.group(1), collections.Counter())
        continue
    m = re.match(r'\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_]*(?:\.[A-Z0-9_]+)*)', line)
    if m and cur is not None:
        op = m.group(1)
        cur['_total'] += 1
        cur[op] += 1
print('# SASS opcode histogram per kernel of %s (cuobjdump -sass, sm_100a)' % os.path.basename(lib))
print('# tcgen05.mma -> UTCHMMA[.2CTA]; tcgen05.ld -> LDTM; TMA -> UTMALDG / UTMASTG; tcgen05.commit -> UTCBAR; mbarrier -> SYNCS\n')
for name, c in kernels.items():
    dn = demangle(name)
    dn = dn.replace('(anonymous namespace)::', '').replace('void ', '')
    dn = re.sub(r'\(.*', '', dn)
    groups = collections.OrderedDict()
    for op, n in sorted(c.items()):
        if op == '_total':
            continue
        for k in KEYS:
            if op == k or op.startswith(k + '.'):
                detail = op if k in ('UTCHMMA', 'UTMALDG', 'UTMASTG', 'UTCBAR', 'LDTM', 'UTCQMMA') else ('REDG.F32x4' if 'F32x4' in op else k)
                groups[detail] = groups.get(detail, 0) + n
                break
    print('%s   [%d instructions]' % (dn, c['_total']))
    print('    ' + '  '.join('%s=%d' % kv for kv in groups.items()))
