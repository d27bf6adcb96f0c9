"""Reduce an `ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum --csv` capture of one training step to the average
DRAM bytes per gemm_kernel launch (bench.py's roofline.traffic).
    python profiles/gemm_dram_traffic.py profiles/r1f_gemm_dram_phoneme_only_step.csv > profiles/r1f_gemm_dram_traffic.json
"""
import csv
import json
import sys


def main(path):
    with open(path, newline="") as f:
        lines = [l for l in f if l.startswith('"')]
    per = {}
    for r in csv.DictReader(lines):
        if "gemm_kernel" not in r["Kernel Name"]:
            continue
        v = float(r["Metric Value"].replace(",", ""))
        u = r["Metric Unit"].lower()
        v *= {"byte": 1, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9}[u]
        per.setdefault(r["ID"], {})[r["Metric Name"]] = v
    n = len(per)
    rd = sum(p.get("dram__bytes_read.sum", 0.0) for p in per.values()) / n
    wr = sum(p.get("dram__bytes_write.sum", 0.0) for p in per.values()) / n
    print(json.dumps({"source": f"ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum over every gemm_kernel launch of one step ({path})",
                      "launches": n, "avg_dram_bytes_per_launch": rd + wr, "avg_read_bytes": rd, "avg_write_bytes": wr}, indent=1))


if __name__ == "__main__":
    main(sys.argv[1])
