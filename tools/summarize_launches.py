"""Turns an `ncu --metrics gpu__time_duration.sum --csv` launch list into a per-kernel table (markdown on stdout)."""
import csv, re, sys
from collections import defaultdict

path = sys.argv[1]
rows = [r for r in csv.reader(open(path, errors="replace")) if len(r) > 5]
hdr = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
H = rows[hdr]
kn, mv, mu = H.index("Kernel Name"), H.index("Metric Value"), H.index("Metric Unit")
tot = defaultdict(float); cnt = defaultdict(int)
for r in rows[hdr + 1:]:
    try:
        v = float(r[mv].replace(",", ""))
    except ValueError:
        continue
    v *= {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(r[mu], 1.0)
    name = re.sub(r"\(.*$", "", r[kn]).replace("void ", "").strip()
    tot[name] += v; cnt[name] += 1
S = sum(tot.values())
print("%d launches, %.2f ms summed\n" % (sum(cnt.values()), S / 1e3))
print("| kernel | launches | total us | avg us | share |\n|---|---|---|---|---|")
for k in sorted(tot, key=lambda k: -tot[k])[: int(sys.argv[2]) if len(sys.argv) > 2 else 40]:
    print("| %s | %d | %.1f | %.2f | %.3f |" % (k, cnt[k], tot[k], tot[k] / cnt[k], tot[k] / S))
