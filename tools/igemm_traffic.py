"""ncu CSV (igemm kernels, per-launch DRAM bytes / duration / tensor / L2 metrics) -> profiles/<name>.json

Capture (one GPU, never a timing run):
  ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,\
sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,lts__throughput.avg.pct_of_peak_sustained_elapsed \
      --clock-control none -k regex:k_igemm --csv --log-file gpurun_out/igemm_metrics.csv \
      python bench.py --steps 1 --warmup 1 --no-graph --no-cpu-baseline --no-e2e --no-topk
usage: python tools/igemm_traffic.py gpurun_out/igemm_metrics.csv profiles/r01_igemm_traffic.json [launches_per_step=170] [skip_tail=54]
(skip_tail: bench.py ends with one batch-2 eval forward that only measures tensor shapes — 54 tiny fprop launches after
the last train step)
"""
import collections, csv, json, re, sys

UNIT = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-6, "us": 1e-3, "usecond": 1e-3, "nsecond": 1e-6,
        "ms": 1.0, "msecond": 1.0, "%": 1.0}


def main():
    src, dst = sys.argv[1], sys.argv[2]
    per_step = int(sys.argv[3]) if len(sys.argv) > 3 else 170
    skip_tail = int(sys.argv[4]) if len(sys.argv) > 4 else 54
    lines = [l for l in open(src, newline="") if not l.startswith("==")]
    launches = collections.OrderedDict()
    for r in csv.DictReader(lines):
        if "k_igemm" not in r["Kernel Name"]:
            continue
        d = launches.setdefault(int(r["ID"]), {"name": re.sub(r"\(.*", "", r["Kernel Name"]).replace("tp::", "").strip()})
        d[r["Metric Name"]] = float(r["Metric Value"].replace(",", "")) * UNIT.get(r["Metric Unit"], 1.0)
    allv = list(launches.values())
    last = allv[len(allv) - skip_tail - per_step: len(allv) - skip_tail]
    agg = collections.OrderedDict()
    for d in last:
        a = agg.setdefault(d["name"], collections.defaultdict(float))
        ms = d.get("gpu__time_duration.sum", 0.0)
        a["launches"] += 1; a["ms"] += ms
        a["rd"] += d.get("dram__bytes_read.sum", 0.0); a["wr"] += d.get("dram__bytes_write.sum", 0.0)
        a["tw"] += ms * d.get("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", 0.0)
        a["lw"] += ms * d.get("lts__throughput.avg.pct_of_peak_sustained_elapsed", 0.0)
    total = sum(a["rd"] + a["wr"] for a in agg.values())
    out = {"source": f"ncu per-launch metrics over the last eager bench.py step (B=512, RN50 ERK-80), igemm kernels only; {src}",
           "launches_per_step": len(last), "dram_bytes_per_step": total, "dram_bytes_per_launch": total / max(len(last), 1),
           "kernels": {k: {"launches": int(a["launches"]), "ms": a["ms"], "dram_read_GB": a["rd"] / 1e9, "dram_write_GB": a["wr"] / 1e9,
                           "tensor_pct_time_weighted": a["tw"] / a["ms"] if a["ms"] else 0.0,
                           "lts_pct_time_weighted": a["lw"] / a["ms"] if a["ms"] else 0.0,
                           "dram_GBps": (a["rd"] + a["wr"]) / a["ms"] / 1e6 if a["ms"] else 0.0} for k, a in agg.items()}}
    json.dump(out, open(dst, "w"), indent=1)
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
