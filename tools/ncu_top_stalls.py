#!/usr/bin/env python
"""Top stalled SASS instructions of a kernel from an .ncu-rep: python tools/ncu_top_stalls.py rep kernel_regex [N]"""
import csv, io, subprocess, sys
rep, kre = sys.argv[1], sys.argv[2]
N = int(sys.argv[3]) if len(sys.argv) > 3 else 30
raw = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv', '--kernel-name', 'regex:' + kre], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr = None
data = []
for r in rows:
    if 'Address' in r and 'Source' in r:
        if hdr is not None:
            break          # first kernel instance only
        hdr = r
        continue
    if hdr is not None and len(r) == len(hdr):
        data.append(r)
ix = {h: i for i, h in enumerate(hdr)}
S = '# Samples'
tot = sum(int(r[ix[S]]) for r in data)
keys = [h for h in hdr if h.startswith('stall_') and 'Not Issued' not in h]
agg = {k[6:]: sum(int(r[ix[k]]) for r in data) for k in keys}
print('total samples', tot, 'instructions', len(data), 'warp-inst executed', sum(int(r[ix['Instructions Executed']]) for r in data))
print({k: round(v / tot, 3) for k, v in sorted(agg.items(), key=lambda x: -x[1]) if v > 0.01 * tot})
ops = {}
for r in data:
    op = r[ix['Source']].strip().split()[0] if not r[ix['Source']].strip().startswith('@') else r[ix['Source']].strip().split()[1]
    op = op.split('.')[0]
    o = ops.setdefault(op, [0, 0]); o[0] += int(r[ix['Instructions Executed']]); o[1] += int(r[ix[S]])
ti = sum(v[0] for v in ops.values())
print('opcode mix (share of executed warp-inst, share of samples):')
for k, v in sorted(ops.items(), key=lambda x: -x[1][0])[:16]:
    print('  %-10s %.3f %.3f' % (k, v[0] / ti, v[1] / tot))
for r in sorted(data, key=lambda r: -int(r[ix[S]]))[:N]:
    st = {k[6:]: int(r[ix[k]]) for k in keys if int(r[ix[k]]) > 0.15 * int(r[ix[S]])}
    print(r[ix[S]], r[ix['Address']][-5:], r[ix['Source']].strip()[:64], st)
