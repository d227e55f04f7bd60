"""Summarise an `ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --csv`
launch list per kernel: usage ncu_launch_table.py launches.csv [--md]"""
import csv, sys, collections, re
rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 10]
hdr = rows[0]
iK, iM, iV, iU = hdr.index('Kernel Name'), hdr.index('Metric Name'), hdr.index('Metric Value'), hdr.index('Metric Unit')
iID = hdr.index('ID')
per = collections.defaultdict(lambda: collections.defaultdict(float)); cnt = collections.Counter(); seen = set()
def short(k):
    k = re.sub(r'\(.*$', '', k).replace('fp::', '').replace('void ', '')
    return k[:70]
lo, hi = -1, 1 << 60
if '--step' in sys.argv:      # trim to one training step: first ingest_kernel .. next repack_batch_kernel
    ids = {}
    for r in rows[1:]: ids.setdefault(int(r[iID]), r[iK])
    starts = [i for i in sorted(ids) if 'ingest_kernel' in ids[i]]
    lo = starts[0]
    ends = [i for i in sorted(ids) if 'repack_batch_kernel' in ids[i] and i > lo]
    hi = ends[0] if ends else max(ids)
for r in rows[1:]:
    if not (lo <= int(r[iID]) <= hi): continue
    k = short(r[iK]); v = float(r[iV].replace(',', '')); u = r[iU]
    if r[iM] == 'gpu__time_duration.sum':
        v = v / 1e6 if u in ('ns', 'nsecond') else (v / 1e3 if u in ('us', 'usecond') else v)   # -> ms
        per[k]['ms'] += v
        if (r[iID], k) not in seen: seen.add((r[iID], k)); cnt[k] += 1
    else:
        mult = {'byte': 1, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9}.get(u, 1)
        per[k]['bytes'] += v * mult
tot = sum(p['ms'] for p in per.values())
print(f"Total {tot:.2f} ms over {sum(cnt.values())} launches.\n")
print("| ms | share | launches | DRAM GB | DRAM TB/s | kernel |\n|---:|---:|---:|---:|---:|---|")
for k, p in sorted(per.items(), key=lambda kv: -kv[1]['ms']):
    if p['ms'] / tot < 0.001: continue
    print(f"| {p['ms']:.3f} | {100*p['ms']/tot:.1f}% | {cnt[k]} | {p['bytes']/1e9:.2f} | {p['bytes']/1e9/max(p['ms'],1e-9):.2f} | `{k}` |")
