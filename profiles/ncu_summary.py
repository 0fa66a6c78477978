"""`ncu -i X.ncu-rep --page raw --csv` -> one line per captured launch with the roofline-relevant metrics
(duration, DRAM bytes and % of peak, tensor-pipe / FMA / XU activity, issue activity, registers).
Usage: python profiles/ncu_summary.py raw.csv > summary.txt"""
import csv
import re
import sys

WANT = [("gpu__time_duration.sum", "ms"), ("dram__bytes_read.sum", "MB rd"), ("dram__bytes_write.sum", "MB wr"),
        ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram%"),
        ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", "tensor%"),
        ("sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_elapsed", "xu%"),
        ("sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_elapsed", "fma%"),
        ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue%"),
        ("lts__t_bytes.sum", "MB L2"), ("launch__registers_per_thread", "regs"), ("launch__grid_size", "grid")]
SCALE = {"ms": {"ns": 1e-6, "us": 1e-3, "ms": 1, "s": 1e3, "nsecond": 1e-6, "usecond": 1e-3, "msecond": 1, "second": 1e3},
         "MB": {"byte": 1e-6, "Kbyte": 1e-3, "Mbyte": 1, "Gbyte": 1e3, "Tbyte": 1e6}}


def main(path):
    rows = list(csv.reader(open(path)))
    hdr, units = rows[0], rows[1]
    ix = {h: i for i, h in enumerate(hdr)}
    print("# " + " | ".join(["kernel"] + [n for _, n in WANT]))
    for r in rows[2:]:
        name = re.sub(r"\(.*", "", r[ix["Kernel Name"]])[:70]
        out = [f"{name:70s}"]
        for key, label in WANT:
            if key not in ix or r[ix[key]] in ("", "n/a"):
                out.append(f"{label} -")
                continue
            v = float(r[ix[key]].replace(",", ""))
            u = units[ix[key]]
            if label == "ms":
                v *= SCALE["ms"].get(u, 1)
            elif label.startswith("MB"):
                v *= SCALE["MB"].get(u, 1)
            out.append(f"{label} {v:.3f}" if label == "ms" else f"{label} {v:.1f}")
        print(" | ".join(out))


if __name__ == "__main__":
    main(sys.argv[1])
