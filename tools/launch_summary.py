"""Summarises an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel
totals of ONE denoise step (delimited by the patchify kernel that starts each DiT step, or
e.g. "end:cfg_ddim" for the UNet step).  usage: launch_summary.py csv [out] [delim] [detail]"""
import collections
import csv
import re
import sys


def main(path, out=None, delim="patchify", detail=False):
    rows = list(csv.reader(open(path)))
    hi = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
    hdr, data = rows[hi], rows[hi + 1:]
    ki, vi, gi = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Grid Size")
    names = [(r[ki], float(r[vi].replace(",", "")) / 1e6, r[gi]) for r in data if len(r) > vi]
    # delim "x": a step starts at each kernel whose name contains x; "end:x": ends at it
    end = delim.startswith("end:")
    key = delim[4:] if end else delim
    # (a run of consecutive delimiter kernels — the CFG-doubled step patchifies twice — is one mark)
    starts = [i + (1 if end else 0) for i, (n, _, _) in enumerate(names)
              if key in n and (end or i == 0 or key not in names[i - 1][0])]
    a, b = starts[-2], starts[-1]
    step = names[a:b]
    agg = collections.defaultdict(lambda: [0, 0.0])
    for n, ms, grid in step:
        k = re.sub(r"^void ", "", n.split("(")[0]).replace("dwm::", "")
        if detail or "attn_kernel" in k:
            k += " grid=" + grid.replace(" ", "")
        agg[k][0] += 1
        agg[k][1] += ms
    tot = sum(v[1] for v in agg.values())
    lines = ["one step (launches %d..%d of %d): %d launches, sum of kernel durations %.1f ms"
             % (a, b, len(names), len(step), tot)]
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        lines.append("%-96s n=%4d ms=%8.2f share=%.3f" % (k[:96], v[0], v[1], v[1] / tot))
    text = "\n".join(lines)
    print(text)
    if out:
        open(out, "w").write(text + "\n")


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2] if len(sys.argv) > 2 else None,
         sys.argv[3] if len(sys.argv) > 3 else "patchify", len(sys.argv) > 4)
