"""Summarise an ncu launch list (--metrics gpu__time_duration.sum --csv) into a markdown table.
usage: python tools/launch_summary.py gpurun_out/launches.csv [skip_first_n_steps_marker]"""
import csv, re, sys, collections

def rows(path):
    with open(path, newline="") as f:
        lines = [l for l in f if not l.startswith("==")]
    rd = csv.DictReader(lines)
    for r in rd:
        if r.get("Metric Name") == "gpu__time_duration.sum":
            v = float(r["Metric Value"].replace(",", ""))
            unit = r.get("Metric Unit", "ns")
            scale = {"ns": 1e-6, "us": 1e-3, "usecond": 1e-3, "nsecond": 1e-6, "ms": 1.0, "msecond": 1.0}.get(unit, 1e-6)
            yield int(r["ID"]), r["Kernel Name"], v * scale

def short(name):
    name = re.sub(r"<.*", "", name)
    name = re.sub(r"\(.*", "", name)
    return name.strip()

def main():
    path = sys.argv[1]
    rs = list(rows(path))
    # the timed step = launches after the LAST k_sgd-but-one ... simpler: split by k_sgd occurrences
    sgd = [i for i, r in enumerate(rs) if "k_sgd" in r[1]]
    if len(sgd) >= 2:
        rs = rs[sgd[-2] + 1: sgd[-1] + 1]
    agg = collections.OrderedDict()
    for _, n, ms in rs:
        k = short(n)
        a = agg.setdefault(k, [0.0, 0])
        a[0] += ms; a[1] += 1
    tot = sum(a[0] for a in agg.values())
    print(f"Kernels in the last step: {len(rs)}; sum of durations {tot:.2f} ms\n")
    print("| ms | share | launches | kernel |\n|---:|---:|---:|---|")
    for k, (ms, n) in sorted(agg.items(), key=lambda kv: -kv[1][0]):
        print(f"| {ms:.3f} | {100 * ms / tot:.1f}% | {n} | `{k}` |")
    ours = sum(ms for k, (ms, n) in agg.items() if "tp::" in k)
    print(f"\nOur kernels (tp::*): {ours:.2f} ms = {100 * ours / tot:.1f}% of the step")

if __name__ == "__main__":
    main()
