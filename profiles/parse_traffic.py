"""ncu CSV (`--metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum -k regex:gemm_kernel` over
tests/gpu_l2_probe.py) -> profiles/gemm_traffic.json: DRAM bytes per launch of each GEMM member of THIS build.
bench.py copies the numbers into `roofline.traffic` / `roofline.traffic_members` (it never invents them).
Usage: python profiles/parse_traffic.py <ncu.csv> [out.json]"""
import csv
import json
import sys

ORDER = ["linear1", "linear2", "fc1", "linear1_f8", "linear2_f8"]  # launch order in tests/gpu_l2_probe.py (x2 rounds)
ALGO = {"linear1": 1.84e9, "linear2": 1.59e9, "fc1": 1.08e9, "linear1_f8": 1.70e9, "linear2_f8": 0.80e9}


def parse(path):
    with open(path) as f:
        lines = [l for l in f if l.startswith('"')]
    per = {}
    for row in csv.DictReader(lines):
        v = float(row["Metric Value"].replace(",", ""))
        unit = row["Metric Unit"]
        mul = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-6, "us": 1e-3, "ms": 1.0,
               "nsecond": 1e-6, "usecond": 1e-3, "msecond": 1.0, "second": 1e3}.get(unit, 1)
        per.setdefault(int(row["ID"]), {})[row["Metric Name"]] = v * mul
    ids = sorted(per)
    out = {}
    for i, lid in enumerate(ids):
        if i < len(ORDER):
            continue  # first round = warm-up
        name = ORDER[i % len(ORDER)]
        m = per[lid]
        rd, wr = m.get("dram__bytes_read.sum", 0.0), m.get("dram__bytes_write.sum", 0.0)
        out[name] = {"dram_bytes": rd + wr, "dram_read": rd, "dram_write": wr, "algorithmic_bytes": ALGO[name],
                     "amplification": (rd + wr) / ALGO[name], "ncu_ms": m.get("gpu__time_duration.sum")}
    return out


if __name__ == "__main__":
    res = parse(sys.argv[1])
    res["_source"] = sys.argv[1]
    dst = sys.argv[2] if len(sys.argv) > 2 else "profiles/gemm_traffic.json"
    with open(dst, "w") as f:
        json.dump(res, f, indent=1)
    print(json.dumps(res, indent=1))
