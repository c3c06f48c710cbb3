"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel share of the summed device time.
(cold-cache, serialised times: compare SHARES, not absolutes)\n    python tools/summarize_launches.py launches.csv [--from-last KERNEL_NAME_SUBSTRING]"""
import collections
import csv
import gzip
import re
import sys

path = sys.argv[1]
f = gzip.open(path, "rt") if path.endswith(".gz") else open(path)
rows = [r for r in csv.reader(l for l in f if l.startswith('"'))]
hdr = rows[0]
ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
if len(sys.argv) > 3 and sys.argv[2] == "--from-last":
    # keep the launches from the LAST occurrence of a kernel on: a capture of N steps cut down to its last whole step
    # (every step starts with the VAE's first convolution, conv3x3_c3_fwd_kernel)
    idx = [i for i, r in enumerate(rows) if i > 0 and sys.argv[3] in r[ki]]
    if idx:
        print(f"# launches {idx[-1]}..{len(rows) - 1} of {len(rows) - 1} (from the last {sys.argv[3]})")
        rows = [hdr] + rows[idx[-1]:]
tot = collections.Counter()
cnt = collections.Counter()
for r in rows[1:]:
    name = re.sub(r"\(.*", "", r[ki])[:70]
    v = float(r[vi].replace(",", ""))
    v *= {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(r[ui], 1.0)     # -> microseconds
    tot[name] += v
    cnt[name] += 1
total = sum(tot.values())
print(f"{len(rows) - 1} launches, total {total / 1e3:.1f} ms")
for name, v in tot.most_common(40):
    print(f"{v / total * 100:6.2f}%  {v / 1e3:9.2f} ms  n={cnt[name]:5d}  avg={v / cnt[name]:9.1f} us  {name}")
