"""Adds / refreshes an entry of profiles/ncu_traffic.json from an ncu report:

    python tools/ncu_traffic_update.py gpurun_out/r02_ncu_geglu_fp16.ncu-rep \
        --kernel gemm2_tcgen05_kernel --shape 86016 12288 1536 --epilogue 1 --dtype fp16 \
        --algorithmic-bytes 1359000000 --capture profiles/r02_ncu_geglu_fp16_summary.txt

The summary text of the same report (tools/ncu_summary.py) is written to --capture, which is
the file bench.py cites as `roofline.traffic_source`."""
import argparse
import csv
import io
import json
import os
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _bytes(value, unit):
    scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}
    return float(value.replace(",", "")) * scale[unit]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("report")
    ap.add_argument("--kernel", required=True)
    ap.add_argument("--shape", type=int, nargs=3, required=True)
    ap.add_argument("--epilogue", type=int, required=True)
    ap.add_argument("--dtype", required=True, choices=["fp16", "bf16"])
    ap.add_argument("--algorithmic-bytes", type=float, default=None)
    ap.add_argument("--capture", required=True)
    a = ap.parse_args()
    out = subprocess.run(["ncu", "-i", a.report, "--page", "raw", "--csv"], capture_output=True,
                         text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    row = [r for r in rows[2:] if a.kernel in r[hdr.index("Kernel Name")]][0]
    rd = _bytes(row[hdr.index("dram__bytes_read.sum")], units[hdr.index("dram__bytes_read.sum")])
    wr = _bytes(row[hdr.index("dram__bytes_write.sum")], units[hdr.index("dram__bytes_write.sum")])
    summary = subprocess.run(["python", os.path.join(ROOT, "tools", "ncu_summary.py"), a.report],
                             capture_output=True, text=True).stdout
    with open(os.path.join(ROOT, a.capture), "w") as f:
        f.write(summary)
    path = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    with open(path) as f:
        table = json.load(f)
    entry = {"kernel": row[hdr.index("Kernel Name")].split("(")[0].replace("void ", ""),
             "shape": a.shape, "epilogue": a.epilogue, "dtype": a.dtype,
             "dram_read_bytes": int(rd), "dram_write_bytes": int(wr), "capture": a.capture}
    if a.algorithmic_bytes:
        entry["algorithmic_bytes"] = int(a.algorithmic_bytes)
    table["kernels"] = [k for k in table["kernels"]
                        if not (k["shape"] == a.shape and k["epilogue"] == a.epilogue and
                                k.get("dtype") == a.dtype)] + [entry]
    with open(path, "w") as f:
        json.dump(table, f, indent=1)
    print(entry)


if __name__ == "__main__":
    main()
