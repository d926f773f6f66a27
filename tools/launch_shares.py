"""Summarise an ncu launch list (gpu__time_duration.sum CSV) of bench.py: per-kernel share of ONE forward."""
import collections
import csv
import re
import sys


def main(path):
    with open(path) as f:
        lines = [l for l in f if not l.startswith("==")]
    rows = list(csv.DictReader(lines))
    names = [r["Kernel Name"] for r in rows]
    dur = [float(r["Metric Value"].replace(",", "")) / 1e6 for r in rows]
    grid = [r["Grid Size"] for r in rows]
    def is_full_fwd_fft(n, g):
        n = n.replace("msfno::", " ")
        return (" rfft_trunc" in n or " rfft2d_kernel" in n) and g.startswith("(23,")
    starts = [i for i, (n, g) in enumerate(zip(names, grid)) if is_full_fwd_fft(n, g)]
    lead = starts[0]  # kernels of one forward that precede the first full-grid FFT (encoder, stats)
    a, b = starts[1] - lead, starts[2] - lead
    agg = collections.defaultdict(lambda: [0, 0.0])
    for i in range(a, b):
        k = re.sub(r"\(.*", "", names[i]).replace("void ", "").replace("msfno::", "")[:64]
        if "gemm_" in names[i] or "rfft" in names[i] or "fft2d" in names[i]:
            k += " grid" + grid[i]
        agg[k][0] += 1
        agg[k][1] += dur[i]
    tot = sum(v[1] for v in agg.values())
    print("launches in one forward: %d, serialized cold total %.3f ms" % (b - a, tot))
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1])[:32]:
        print("%7.3f ms %5.1f%% x%3d  %s" % (v[1], 100 * v[1] / tot, v[0], k))


if __name__ == "__main__":
    main(sys.argv[1])
