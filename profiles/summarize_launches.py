"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list into per-kernel launches / total ms / share.
usage: python profiles/summarize_launches.py gpurun_out/<launches>.csv "<command that was profiled>" > profiles/<name>.csv"""
import csv, sys, collections
rows = [r for r in csv.reader(l for l in open(sys.argv[1]) if not l.startswith("=="))]
h = rows[0]
ki, vi, ui = h.index("Kernel Name"), h.index("Metric Value"), h.index("Metric Unit")
tot, cnt = collections.Counter(), collections.Counter()
for r in rows[1:]:
    if len(r) <= vi:
        continue
    v = float(r[vi].replace(",", ""))
    ms = v / 1e6 if r[ui] in ("ns", "nsecond") else v / 1e3 if r[ui] in ("us", "usecond") else v
    tot[r[ki]] += ms
    cnt[r[ki]] += 1
T = sum(tot.values())
print(f"# ncu launch list summary: {sys.argv[2] if len(sys.argv) > 2 else ''}")
print("# ncu --metrics gpu__time_duration.sum --clock-control none  (cold-cache, serialised: compare SHARES, not absolutes)")
print(f"# {sum(cnt.values())} launches, total {T:.1f} ms")
print("kernel,launches,total_ms,share")
for k, v in tot.most_common():
    print(f"\"{k[:75]}\",{cnt[k]},{v:.3f},{v / T:.4f}")
