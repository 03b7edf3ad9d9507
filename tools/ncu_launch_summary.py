"""Summarise an ncu CSV launch list (`--metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --csv`) of
tools/step_once.py: the launches of the MEASURED step (after the first adam_kernel, up to and including the second), grouped by
kernel: count, summed / average duration, share of the step, DRAM bytes per launch.

    python tools/ncu_launch_summary.py gpurun_out/r02_launches.csv > profiles/r02_ncu_launch_list_b16_step.txt"""
import collections
import csv
import re
import sys

rows = collections.OrderedDict()
with open(sys.argv[1]) as f:
    lines = [ln for ln in f if ln.startswith('"')]
for r in csv.DictReader(lines):
    d = rows.setdefault(int(r["ID"]), {"name": r["Kernel Name"]})
    d[r["Metric Name"]] = float(r["Metric Value"].replace(",", ""))
    d[r["Metric Name"] + ".unit"] = r["Metric Unit"]
ids = sorted(rows)
adam = [i for i in ids if "adam_kernel" in rows[i]["name"]]
lo, hi = adam[0], adam[-1]
step = [rows[i] for i in ids if lo < i <= hi]


def short(n):
    n = re.sub(r"\(.*", "", n)
    n = n.replace("void ", "").replace("<unnamed>::", "")
    return n[:78]


def us(d):
    v, u = d["gpu__time_duration.sum"], d.get("gpu__time_duration.sum.unit", "ns")
    return v / 1e3 if u in ("ns", "nsecond") else v * 1e3 if u in ("ms", "msecond") else v


def mb(d, k):
    v, u = d.get(k, 0.0), d.get(k + ".unit", "byte")
    return v / 1e6 if u == "byte" else v / 1e3 if u == "Kbyte" else v if u == "Mbyte" else v * 1e3


agg = collections.defaultdict(lambda: [0, 0.0, 0.0, 0.0])
for d in step:
    a = agg[short(d["name"])]
    a[0] += 1
    a[1] += us(d)
    a[2] += mb(d, "dram__bytes_read.sum")
    a[3] += mb(d, "dram__bytes_write.sum")
tot = sum(a[1] for a in agg.values())
print("kernels in the measured step: %d, summed duration %.2f ms (per-launch times under ncu are serialised and cold-cache: compare SHARES)" % (len(step), tot / 1e3))
print("%10s %6s %6s %10s %12s %12s  kernel" % ("total ms", "share", "count", "avg us", "rd MB/launch", "wr MB/launch"))
for name, a in sorted(agg.items(), key=lambda kv: -kv[1][1])[:60]:
    print("%10.3f %5.1f%% %6d %10.1f %12.2f %12.2f  %s" % (a[1] / 1e3, 100 * a[1] / tot, a[0], a[1] / a[0], a[2] / a[0], a[3] / a[0], name))
g = [a for n, a in agg.items() if n.startswith("gemm_tc")]
if g:
    n = sum(a[0] for a in g)
    print("\nall gemm_tc* launches (ofa_gemm_bf16): %d launches, %.2f ms, DRAM read %.2f MB + write %.2f MB per launch" % (
        n, sum(a[1] for a in g) / 1e3, sum(a[2] for a in g) / n, sum(a[3] for a in g) / n))
