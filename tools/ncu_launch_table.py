"""Summarise an ncu launch list (``--csv`` with gpu__time_duration.sum and friends).

    python tools/ncu_launch_table.py profiles/r02_launches_bench.csv [--last N]

Prints one row per kernel name: launches, mean duration, share of the summed
time, warp instructions and DRAM bytes per launch.  ``--last N`` keeps only the
final N launches (one timed step of bench.py).
"""
import argparse
import csv
import re
from collections import OrderedDict


def short_name(name: str) -> str:
    name = re.sub(r"\(.*$", "", name)
    name = re.sub(r"^void\s+", "", name)
    name = re.sub(r"vkb::|\(anonymous namespace\)::", "", name)
    return name


def load(path):
    with open(path, newline="") as handle:
        lines = [line for line in handle if not line.startswith("==")]
    rows = list(csv.DictReader(lines))
    launches = OrderedDict()
    for row in rows:
        key = int(row["ID"])
        item = launches.setdefault(key, {"name": short_name(row["Kernel Name"])})
        value = float(row["Metric Value"].replace(",", ""))
        unit = row["Metric Unit"]
        metric = row["Metric Name"]
        if metric == "gpu__time_duration.sum":
            scale = {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(unit, 1e-3)
            item["us"] = value * scale
        elif metric.startswith("dram__bytes"):
            scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(unit, 1.0)
            item["dram"] = item.get("dram", 0.0) + value * scale
        elif metric == "smsp__inst_executed.sum":
            item["inst"] = value
    return list(launches.values())


def main():
    parser = argparse.ArgumentParser()
    parser.add_argument("csv")
    parser.add_argument("--last", type=int, default=0)
    args = parser.parse_args()
    launches = load(args.csv)
    if args.last:
        launches = launches[-args.last:]
    table = OrderedDict()
    for item in launches:
        row = table.setdefault(item["name"], {"n": 0, "us": 0.0, "dram": 0.0, "inst": 0.0})
        row["n"] += 1
        row["us"] += item.get("us", 0.0)
        row["dram"] += item.get("dram", 0.0)
        row["inst"] += item.get("inst", 0.0)
    total = sum(row["us"] for row in table.values())
    print(f"{'kernel':<58} {'n':>3} {'us/launch':>10} {'share':>6} {'Minst':>8} {'dram MB':>9}")
    for name, row in table.items():
        n = row["n"]
        print(f"{name[:58]:<58} {n:>3} {row['us'] / n:>10.1f} {row['us'] / total:>6.1%} "
              f"{row['inst'] / n / 1e6:>8.2f} {row['dram'] / n / 1e6:>9.1f}")
    print(f"{'sum':<58} {len(launches):>3} {total:>10.1f}")


if __name__ == "__main__":
    main()
