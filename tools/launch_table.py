"""One K-loop iteration out of an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel totals and the
launch sequence.  Usage: python tools/launch_table.py launches.csv [marker-kernel-substring]"""
import collections
import csv
import sys


def main():
    path = sys.argv[1]
    marker = sys.argv[2] if len(sys.argv) > 2 else "neumann_update"
    rows = list(csv.reader(open(path)))
    hi = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
    hdr = rows[hi]
    kn, mv = hdr.index("Kernel Name"), hdr.index("Metric Value")
    seq = []
    for r in rows[hi + 1:]:
        if len(r) > mv:
            try:
                seq.append((r[kn], float(r[mv].replace(",", ""))))
            except ValueError:
                pass
    idx = [i for i, (n, _) in enumerate(seq) if marker in n]
    a, b = idx[-3] + 1, idx[-2] + 1
    it = seq[a:b]
    short = lambda n: n.split("(")[0].replace("void ", "").replace("<unnamed>::", "")[:64]
    print(f"launches per iteration: {len(it)}; sum of kernel durations: {sum(t for _, t in it) / 1e3:.1f} us\n")
    agg = collections.OrderedDict()
    for n, t in it:
        k = short(n)
        agg.setdefault(k, [0.0, 0])
        agg[k][0] += t
        agg[k][1] += 1
    tot = sum(t for _, t in it)
    print("| us | share | launches | kernel |\n|---|---|---|---|")
    for k, (t, c) in sorted(agg.items(), key=lambda kv: -kv[1][0]):
        print(f"| {t / 1e3:.1f} | {100 * t / tot:.1f}% | {c} | `{k}` |")
    print("\nsequence (us):")
    print(", ".join(f"{short(n).split('<')[0]} {t / 1e3:.0f}" for n, t in it))


if __name__ == "__main__":
    main()
