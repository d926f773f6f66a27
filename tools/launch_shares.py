"""Summarise an ncu launch list (`--metrics gpu__time_duration.sum --csv`) of bench.py: per-kernel share of ONE forward step.

The step boundary is found from the launch list itself: the (kernel name, grid) sequence of a forward repeats, so the
smallest period L with rows[-L:] == rows[-2L:-L] whose window holds EVERY distinct launch of the list's second half is
one step -- whatever the precision tier or workload launches (the ten identical inner blocks of a net repeat with a
shorter period, but their window lacks the full-grid launches of blocks 0 / 11 and of the encoder / decoder).
usage: python tools/launch_shares.py profiles/<launch list>.csv [--json]"""
import collections
import csv
import json
import re
import sys


def load(path):
    with open(path) as f:
        lines = [l for l in f if not l.startswith("==")]
    rows = list(csv.DictReader(lines))
    names = [r["Kernel Name"] for r in rows]
    dur = [float(r["Metric Value"].replace(",", "")) / 1e6 for r in rows]   # ns -> ms
    grid = [r["Grid Size"] for r in rows]
    return names, grid, dur


def short(name):
    k = re.sub(r"\(.*", "", name).replace("void ", "").replace("msfno::", "")
    return re.sub(r"at::native::|<unnamed>::", "", k)[:72]


def period(keys, lo=8):
    n = len(keys)
    everything = set(keys[n // 2:])
    for L in range(lo, n // 2 + 1):
        if keys[n - L:] == keys[n - 2 * L:n - L] and set(keys[n - L:]) == everything:
            return L
    return None


def shares(path):
    names, grid, dur = load(path)
    keys = list(zip(names, grid))
    L = period(keys)
    if L is None:
        raise SystemExit("no repeating step found in %s (%d launches): profile at least two steps" % (path, len(keys)))
    a, b = len(keys) - L, len(keys)
    agg = collections.defaultdict(lambda: [0, 0.0])
    for i in range(a, b):
        k = short(names[i])
        if "gemm_" in k or "fft" in k or "dft" in k:
            k += " grid" + grid[i].replace(" ", "")
        agg[k][0] += 1
        agg[k][1] += dur[i]
    tot = sum(v[1] for v in agg.values())
    by_kernel = collections.defaultdict(lambda: [0, 0.0])
    for i in range(a, b):
        k = re.sub(r"<.*", "", short(names[i]))
        by_kernel[k][0] += 1
        by_kernel[k][1] += dur[i]
    return {"launches_per_step": L, "serialized_cold_ms": tot,
            "by_kernel": {k: {"launches": v[0], "ms": v[1], "share": v[1] / tot} for k, v in sorted(by_kernel.items(), key=lambda kv: -kv[1][1])},
            "by_shape": {k: {"launches": v[0], "ms": v[1], "share": v[1] / tot} for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1])}}


def main():
    path = sys.argv[1]
    res = shares(path)
    if "--json" in sys.argv:
        print(json.dumps(res, indent=1))
        return
    print("launches in one forward: %d, serialized cold total %.3f ms" % (res["launches_per_step"], res["serialized_cold_ms"]))
    for k, v in res["by_kernel"].items():
        print("%7.3f ms %5.1f%% x%3d  %s" % (v["ms"], 100 * v["share"], v["launches"], k))
    print("--- by kernel and grid")
    for k, v in list(res["by_shape"].items())[:32]:
        print("%7.3f ms %5.1f%% x%3d  %s" % (v["ms"], 100 * v["share"], v["launches"], k))


if __name__ == "__main__":
    main()
