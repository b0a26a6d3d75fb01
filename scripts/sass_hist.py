"""Opcode histogram (executed instructions, stall samples) from `ncu --page source --csv` output."""
import collections, csv, sys
rows = list(csv.reader(open(sys.argv[1])))
hi = next(i for i, r in enumerate(rows) if 'Source' in r and 'Instructions Executed' in r)
hdr = rows[hi]
ai, ii, si = hdr.index('Source'), hdr.index('Instructions Executed'), hdr.index('# Samples')
data = []
for r in rows[hi + 1:]:
    try:
        data.append((r[ai].strip(), int(r[ii]), int(r[si])))
    except (ValueError, IndexError):
        pass
tot = sum(d[1] for d in data); ts = sum(d[2] for d in data)
print('total warp-inst', tot, 'samples', ts)
h = collections.Counter(); hs = collections.Counter()
for s, n, sm in data:
    parts = s.split()
    op = parts[1] if parts[0].startswith('@') else parts[0]
    op = op.split('.')[0]
    h[op] += n; hs[op] += sm
for op, n in h.most_common(int(sys.argv[2]) if len(sys.argv) > 2 else 24):
    print(f"{op:10s} {100*n/tot:5.1f}% inst   {100*hs[op]/max(ts,1):5.1f}% samples")
