"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: the launches of the last
search step (from the last coarse dense_exact launch with a DENSE output to the next one)."""
import csv, sys, collections
path = sys.argv[1]
which = int(sys.argv[2]) if len(sys.argv) > 2 else -1
rows = list(csv.reader(open(path, errors="ignore")))
h = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
hdr = rows[h]; ci = {k: i for i, k in enumerate(hdr)}
seq = []
for r in rows[h + 1:]:
    if len(r) < len(hdr): continue
    v = float(r[ci["Metric Value"]].replace(",", "")); u = r[ci["Metric Unit"]]
    ms = v / 1e6 if u in ("ns", "nsecond") else v / 1e3 if u.startswith("us") else v
    name = r[ci["Kernel Name"]].replace("auncel::", "").replace("void ", "")
    name = name.split("(")[0]
    seq.append((name, r[ci["Grid Size"]], ms))
starts = [i for i, s in enumerate(seq) if s[0].startswith("dense_exact_kernel<1, 0>") or s[0].startswith("dense_exact_kernel<0, 0>")]
a = starts[which]; b = starts[which + 1] if which + 1 < 0 and which != -1 else len(seq)
if which != -1 and which + 1 < len(starts): b = starts[which + 1]
step = seq[a:b]
agg = collections.OrderedDict()
for n, g, ms in step:
    d = agg.setdefault(n, [0, 0.0]); d[0] += 1; d[1] += ms
tot = sum(ms for _, _, ms in step)
print(f"launches {len(step)}  sum of kernel durations {tot:.3f} ms (serialised, cold cache under ncu)")
for n, (c, ms) in sorted(agg.items(), key=lambda x: -x[1][1]):
    print(f"{n:48s} launches={c:4d} total_ms={ms:9.3f} share={100*ms/tot:5.1f}%")
if "-v" in sys.argv:
    for n, g, ms in step:
        if ms > 0.03: print(f"  {n:46s} {g:>16s} {ms:8.3f}")
