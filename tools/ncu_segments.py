#!/usr/bin/env python
"""Split a kernel's SASS at BAR.SYNC instructions and report, per segment (= one shared-memory pass of the slice loop),
executed warp instructions, stall samples and the dominant stall reasons.
usage: python tools/ncu_segments.py rep kernel_regex"""
import csv, io, subprocess, sys
rep, kre = sys.argv[1], sys.argv[2]
raw = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv', '--kernel-name', 'regex:' + kre], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, data = None, []
for r in rows:
    if 'Address' in r and 'Source' in r:
        if hdr is not None:
            break
        hdr = r
        continue
    if hdr is not None and len(r) == len(hdr):
        data.append(r)
ix = {h: i for i, h in enumerate(hdr)}
keys = [h for h in hdr if h.startswith('stall_') and 'Not Issued' not in h]
segs, cur = [], {'n': 0, 'inst': 0, 'samp': 0, 'st': {}, 'lds': 0, 'sts': 0, 'fp': 0, 'first': None}
tot_s = sum(int(r[ix['# Samples']]) for r in data)
tot_i = sum(int(r[ix['Instructions Executed']]) for r in data)
for r in data:
    src = r[ix['Source']].strip()
    op = (src.split()[1] if src.startswith('@') else src.split()[0]).split('.')[0]
    e = int(r[ix['Instructions Executed']])
    cur['n'] += 1; cur['inst'] += e; cur['samp'] += int(r[ix['# Samples']])
    if cur['first'] is None: cur['first'] = r[ix['Address']][-5:]
    if op == 'LDS': cur['lds'] += e
    if op == 'STS': cur['sts'] += e
    if op in ('FADD', 'FMUL', 'FFMA'): cur['fp'] += e
    for k in keys:
        cur['st'][k[6:]] = cur['st'].get(k[6:], 0) + int(r[ix[k]])
    if op == 'BAR':
        segs.append(cur)
        cur = {'n': 0, 'inst': 0, 'samp': 0, 'st': {}, 'lds': 0, 'sts': 0, 'fp': 0, 'first': None}
segs.append(cur)
print('total warp-inst %d samples %d' % (tot_i, tot_s))
for i, s in enumerate(segs):
    top = sorted(s['st'].items(), key=lambda x: -x[1])[:4]
    print('seg %2d @%s sass %4d  inst %5.1f%%  samples %5.1f%%  fp %4.1f%% lds %4.1f%% sts %4.1f%%  %s' % (
        i, s['first'], s['n'], 100. * s['inst'] / tot_i, 100. * s['samp'] / tot_s, 100. * s['fp'] / tot_i,
        100. * s['lds'] / tot_i, 100. * s['sts'] / tot_i, ' '.join('%s=%.1f%%' % (k, 100. * v / tot_s) for k, v in top)))
