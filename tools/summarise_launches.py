"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per kernel count, total and mean time,
share of the library's own kernels (namespace bias::)."""
import csv, sys, collections, re
rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 14 and r[0].isdigit()]
agg = collections.OrderedDict()
for r in rows:
    name = re.sub(r"\(.*", "", r[4]).replace("void ", "")
    unit, val = r[13], float(r[14].replace(",", ""))
    us = val / 1e3 if unit.startswith("ns") else val * (1e3 if unit.startswith("ms") else 1.0)
    c, t = agg.get(name, (0, 0.0))
    agg[name] = (c + 1, t + us)
foreign = ("at::", "at_cuda", "nccl", "cub::", "thrust::", "cutlass", "cublas", "elementwise_kernel", "vectorized_", "Memset", "Memcpy")
has_ns = any(k.startswith("bias::") for k in agg)      # ncu prints the namespace only in some of its demangling modes
ours = {k: v for k, v in agg.items() if (k.startswith("bias::") if has_ns else not any(f in k for f in foreign))}
tot = sum(t for _, t in ours.values())
print(f"{len(rows)} launches captured; {sum(c for c, _ in ours.values())} of them are libbias_cuda kernels ({tot / 1e3:.3f} ms)")
print(f"{'kernel':70s} {'launches':>8s} {'total ms':>10s} {'mean us':>10s} {'share':>7s}")
for k, (c, t) in sorted(ours.items(), key=lambda kv: -kv[1][1]):
    print(f"{k[:70]:70s} {c:8d} {t / 1e3:10.3f} {t / c:10.1f} {100 * t / tot:6.1f}%")
other = sum(t for k, (c, t) in agg.items() if k not in ours)
print(f"(torch kernels of the synthetic-corpus generator and NCCL, not on the timed path: {other / 1e3:.3f} ms)")
