"""Turn an `ncu --metrics gpu__time_duration.sum --csv` launch list into the markdown table kept under
profiles/ (per kernel: launches, total time, share of the device time of the run).

    python profiles/launch_table.py gpurun_out/launches.csv [bench.json] > profiles/rNN_ncu_launches.md
"""
import collections
import csv
import json
import sys


def main():
    rows = []
    with open(sys.argv[1], newline="") as fh:
        lines = [line for line in fh if not line.startswith("==")]
    reader = csv.DictReader(lines)
    for r in reader:
        if r.get("Metric Name") != "gpu__time_duration.sum":
            continue
        value = float(r["Metric Value"].replace(",", ""))
        unit = r.get("Metric Unit", "ns")
        scale = {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(unit, 1e-6)
        rows.append((r["Kernel Name"], value * scale))
    total = sum(ms for _, ms in rows)
    agg = collections.OrderedDict()
    for name, ms in rows:
        n, t = agg.get(name, (0, 0.0))
        agg[name] = (n + 1, t + ms)
    print(f"Total device time over {len(rows)} launches: {total:.1f} ms\n")
    print("| kernel | launches | total ms | share |\n|---|---|---|---|")
    for name, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:28]:
        print(f"| `{name[:70]}` | {n} | {t:.2f} | {t / total:.3f} |")
    if len(sys.argv) > 2:
        j = json.loads(open(sys.argv[2]).read().strip().splitlines()[-1])
        print("\nShares from the CUDA-event table of the un-profiled bench run (`kernels`, separate 2-step "
              "pass) for comparison:\n")
        print("| C-ABI entry point | ms / launch | share of step |\n|---|---|---|")
        table = j.get("kernels") or {k: v for k, v in j.items()
                                     if isinstance(v, dict) and "ms_per_launch" in v}
        for k, v in sorted(table.items(), key=lambda kv: -kv[1]["share_of_step"]):
            print(f"| `{k}` | {v['ms_per_launch']} | {v['share_of_step']} |")


if __name__ == "__main__":
    main()
