"""Attribute `ncu --page source --csv` instruction counts to CUDA source lines using nvdisasm -g
output of the same cubin.  usage: ncu_lines.py src.csv dis.txt mangled_substring [top]"""
import csv, re, sys, collections
rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[1]
iE, iSm, iS = hdr.index('Instructions Executed'), hdr.index('# Samples'), hdr.index('Source')
inst = [(int(r[iE]), int(r[iSm] or 0), r[iS].strip()) for r in rows[2:] if len(r) > iE and r[iE].isdigit()]
lines = open(sys.argv[2]).read().split('\n')
start = next(i for i, l in enumerate(lines) if '.text.' in l and sys.argv[3] in l and l.startswith('//---'))
cur = None; seq = []
for l in lines[start + 1:]:
    if l.startswith('//---') and '.text.' in l: break
    m = re.search(r'//## File ".*?", line (\d+)', l)
    if m: cur = int(m.group(1)); continue
    if re.match(r'\s+/\*[0-9a-f]{4,}\*/', l): seq.append(cur)
print('sass in ncu', len(inst), 'sass in disasm', len(seq))
per = collections.Counter(); sm = collections.Counter()
for (e, s, _), ln in zip(inst, seq): per[ln] += e; sm[ln] += s
tot = sum(per.values())
src = open('floodplanet_code_b200/csrc/' + sys.argv[5]).read().split('\n') if len(sys.argv) > 5 else None
for ln, e in per.most_common(int(sys.argv[4]) if len(sys.argv) > 4 else 25):
    text = src[ln - 1].strip()[:90] if src and ln else ''
    print(f"{e:12d} {100*e/tot:5.1f}% samp {sm[ln]:6d}  L{ln}: {text}")
