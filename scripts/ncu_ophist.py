"""Opcode histogram (warp-level executed instructions, stall samples) from `ncu --page source --csv`."""
import csv, collections, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[1]
iS, iE, iSm = hdr.index('Source'), hdr.index('Instructions Executed'), hdr.index('# Samples')
tot = 0; ops = collections.Counter(); samp = collections.Counter()
for r in rows[2:]:
    if len(r) <= iE or not r[iE].isdigit(): continue
    e = int(r[iE]); tot += e
    t = r[iS].split()
    op = (t[1] if t[0].startswith('@') else t[0]).split('.')[0]
    ops[op] += e; samp[op] += int(r[iSm] or 0)
print('total warp instr', tot, 'sass lines', len(rows) - 2)
for k, v in ops.most_common(int(sys.argv[2]) if len(sys.argv) > 2 else 20):
    print(f"{k:10s} {v:12d} {100*v/tot:5.1f}%  samples {samp[k]}")
