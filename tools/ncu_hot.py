#!/usr/bin/env python3
"""Top stalled SASS instructions of the first kernel in an .ncu-rep (source page), with a per-reason total."""
import csv, subprocess, sys, io
rep = sys.argv[1]; topn = int(sys.argv[2]) if len(sys.argv) > 2 else 40
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"] + (sys.argv[3:] if len(sys.argv) > 3 else []), capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
# find header row
hi = next(i for i, r in enumerate(rows) if 'Source' in r and '# Samples' in r)
hdr = rows[hi]; ix = {h: i for i, h in enumerate(hdr)}
stalls = [h for h in hdr if h.startswith('stall_') and 'Not Issued' not in h]
items = []; agg = {h: 0 for h in stalls}; tot = 0
for n, r in enumerate(rows[hi + 1:]):
    if len(r) != len(hdr): break
    try: s = int(r[ix['# Samples']])
    except: continue
    tot += s
    for h in stalls:
        try: agg[h] += int(r[ix[h]])
        except: pass
    items.append((s, n, r))
print('samples', tot, 'instructions', len(items))
print(' '.join('%s=%.1f%%' % (k[6:], 100.0 * v / max(1, tot)) for k, v in sorted(agg.items(), key=lambda kv: -kv[1]) if v > 0.005 * tot))
# cumulative by region: print every instruction's samples compactly for top
for s, n, r in sorted(items, reverse=True)[:topn]:
    top = sorted([(int(r[ix[h]] or 0), h[6:]) for h in stalls], reverse=True)[:2]
    print('%6d %5d  %-60s %s' % (s, n, r[ix['Source']][:60], top))
