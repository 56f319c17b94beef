"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list (one bench round) into profiles/launches_top.md:
device-time share per kernel and the split between this repo's kernels (`hfl::`) and library kernels."""
import collections
import csv
import sys

src = sys.argv[1] if len(sys.argv) > 1 else "gpurun_out/launches.csv"
dst = sys.argv[2] if len(sys.argv) > 2 else "profiles/launches_top.md"
rows = []
with open(src, newline="") as f:
    lines = [ln for ln in f if not ln.startswith("==")]
rd = csv.reader(lines)
header = None
for r in rd:
    if header is None:
        if "Kernel Name" in r:
            header = r
        continue
    if len(r) != len(header):
        continue
    rec = dict(zip(header, r))
    if rec.get("Metric Name") != "gpu__time_duration.sum":
        continue
    val = float(rec["Metric Value"].replace(",", ""))
    unit = rec.get("Metric Unit", "ns")
    us = val * {"ns": 1e-3, "nsecond": 1e-3, "us": 1.0, "usecond": 1.0, "ms": 1e3, "msecond": 1e3}.get(unit, 1e-3)
    rows.append((rec["Kernel Name"], us))
cnt, tim = collections.Counter(), collections.Counter()
for name, us in rows:
    key = name.split("(")[0][:110]
    cnt[key] += 1
    tim[key] += us
total = sum(tim.values()) or 1.0
own = sum(v for k, v in tim.items() if "hfl::" in k)
own_n = sum(v for k, v in cnt.items() if "hfl::" in k)
out = ["# Device-time shares of the kernels of one benchmark round (ncu --metrics gpu__time_duration.sum, %d launches sampled)" % len(rows),
       "", "Serialised, cold-cache replay under ncu: compare SHARES, not absolutes (B200_PROFILING.md).  `hfl::` = this repo's kernels.",
       "", "**This repo's kernels: %.1f %% of the sampled device time, %d of %d launches (%.1f %%).**" % (
           100 * own / total, own_n, len(rows), 100.0 * own_n / max(len(rows), 1)), "",
       "| share | launches | total us | avg us | kernel |", "|---|---|---|---|---|"]
for k, v in sorted(tim.items(), key=lambda kv: -kv[1])[:40]:
    out.append("| %.1f %% | %d | %.0f | %.1f | `%s` |" % (100 * v / total, cnt[k], v, v / cnt[k], k))
open(dst, "w").write("\n".join(out) + "\n")
print("\n".join(out[:12]))
