#!/usr/bin/env python
"""SASS instruction count and opcode histogram per kernel of the built library (cuobjdump -sass), plus the mnemonics that
evidence the design: python tools/sass_summary.py [kernel-name regex] > profiles/<tag>_sass_opcode_summary.txt"""
import collections, re, subprocess, sys
LIB = 'adorym_b200/libadorym_b200.so'
pat = re.compile(sys.argv[1]) if len(sys.argv) > 1 else None
out = subprocess.run(['cuobjdump', '-sass', LIB], capture_output=True, text=True).stdout
name, ops = None, None
kernels = []
for line in out.splitlines():
    m = re.match(r'\s*Function : (\S+)', line)
    if m:
        name = subprocess.run(['c++filt', m.group(1)], capture_output=True, text=True).stdout.strip()
        ops = collections.Counter()
        kernels.append((name, ops))
        continue
    m = re.match(r'\s*/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)', line)
    if m and ops is not None:
        ops[m.group(1)] += 1
EVID = ('SYNCS', 'UTMALDG', 'UTMAREDG', 'UBLKCP', 'UBLKPF', 'FADD2', 'FMUL2', 'FFMA2', 'REDG', 'UTMACMDFLUSH', 'LDL', 'STL')
print('cuobjdump -sass %s (sm_100a): SASS instruction count and opcode histogram per kernel.' % LIB)
print('Mnemonics that evidence the design: UTMALDG / UTMAREDG / UBLKCP (TMA tile load, TMA reduce-add, bulk copy), SYNCS (mbarrier), '
      'FADD2 / FMUL2 / FFMA2 (packed complex arithmetic), REDG (red.global.add), LDL / STL (spills).\n')
for name, ops in sorted(kernels, key=lambda k: -sum(k[1].values())):
    if pat is not None and not pat.search(name):
        continue
    print(name)
    print('  %d instructions: %s' % (sum(ops.values()), ', '.join('%s %d' % kv for kv in ops.most_common(18))))
    print('  evidence: %s' % {k: ops[k] for k in EVID if ops[k]})
