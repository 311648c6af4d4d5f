"""Aggregate `ncu -i X.ncu-rep --page source --print-source cuda,sass --csv` per source line:
share of stall samples, of executed instructions, lanes per instruction and the dominant stall reasons.
Usage: python scripts/ncu_hotspots.py source.csv [top]"""
import csv, sys, collections
path = sys.argv[1]; top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
rows = list(csv.reader(open(path)))
cur_file = None; hdr = None; last_line = 0
agg = collections.defaultdict(lambda: collections.Counter()); src = {}
for r in rows:
    if len(r) == 2 and r[0] == "File Path": cur_file = r[1].split("/")[-1]; continue
    if len(r) == 2: continue
    if r[0] == "Line No": hdr = r; continue
    if hdr is None or len(r) != len(hdr): continue
    d = dict(zip(hdr, r))
    if r[0].strip():
        last_line = int(r[0]); src[(cur_file, last_line)] = r[1]
    key = (cur_file, last_line)
    a = agg[key]
    def num(k):
        try: return float(d.get(k, "0") or 0)
        except ValueError: return 0.0
    a["samples"] += num("# Samples"); a["inst"] += num("Instructions Executed"); a["thr"] += num("Thread Instructions Executed")
    a["loc"] += num("L2 Theoretical Sectors Local")
    for k in hdr:
        if k.startswith("stall_") and "Not Issued" not in k: a[k] += num(k)
S = sum(a["samples"] for a in agg.values()); I = sum(a["inst"] for a in agg.values()); T = sum(a["thr"] for a in agg.values())
print(f"total samples {S:.0f} instructions {I:.0f} lanes/instruction {T / max(I, 1):.2f}")
byfile = collections.Counter()
for (f, l), a in agg.items(): byfile[f] += a["samples"]
for f, v in byfile.most_common(): print(f"  {f}: {v / S:.3f} of samples")
stall = collections.Counter()
for a in agg.values():
    for k, v in a.items():
        if k.startswith("stall_"): stall[k] += v
st = sum(stall.values())
print("stall reasons:", " ".join(f"{k[6:]} {v / st:.3f}" for k, v in stall.most_common(8)))
print("hottest lines (share of samples | share of instructions | lanes/instruction | top stalls | local sectors):")
for key, a in sorted(agg.items(), key=lambda kv: -kv[1]["samples"])[:top]:
    ss = sorted(((k[6:], v) for k, v in a.items() if k.startswith("stall_") and v > 0), key=lambda kv: -kv[1])[:3]
    tot = sum(v for k, v in a.items() if k.startswith("stall_")) or 1
    print(f"  {key[0]}:{key[1]:<4d} {a['samples'] / S:.3f} {a['inst'] / I:.3f} {a['thr'] / max(a['inst'], 1):5.1f}  "
          + " ".join(f"{k} {v / tot:.2f}" for k, v in ss) + f"  loc {a['loc']:.0f} | {src.get(key, "").strip()[:90]}")
