"""Turn the CSVs tools/gpu_validate.sh brings back into the tables of profiles/*.md.

  python tools/summarise_ncu.py launches gpurun_out/<tag>_launches.csv      # per-kernel totals of the last full bench step
  python tools/summarise_ncu.py raw gpurun_out/<tag>_gather_raw.csv         # ncu --set full: one row per captured launch
"""
import csv
import sys
from collections import OrderedDict


def _rows(path):
    with open(path) as fh:
        lines = [ln for ln in fh if not ln.startswith("==")]
    return list(csv.DictReader(lines))


def launches(path):
    rows = _rows(path)
    names = [r["Kernel Name"] for r in rows]
    marks = [i for i, n in enumerate(names) if "dot_kernel" in n]          # the loss's dot product: once per step
    a, b = marks[-2], marks[-1]
    step = rows[a:b]
    unit = step[0]["Metric Unit"]
    sc = 1e-6 if unit.startswith("n") else 1e-3
    agg = OrderedDict()
    for r in step:
        n = r["Kernel Name"].replace("void ", "").split("(")[0].replace("gode::", "")
        t = float(r["Metric Value"].replace(",", "")) * sc
        c, s = agg.get(n, (0, 0.0))
        agg[n] = (c + 1, s + t)
    total = sum(s for _, s in agg.values())
    print("%d launches, %.1f ms summed" % (len(step), total))
    print("| kernel | launches | total ms | share |\n|---|---:|---:|---:|")
    for n, (c, s) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print("| `%s` | %d | %.2f | %.1f %% |" % (n[:90], c, s, 100 * s / total))


RAW = [("gpu__time_duration.sum", "ms", 1.0), ("dram__bytes_read.sum", "DRAM read GB", 1.0), ("dram__bytes_write.sum", "DRAM write GB", 1.0),
       ("lts__t_sector_hit_rate.pct", "L2 hit %", 1.0), ("l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed", "L1 LSU pipe %", 1.0),
       ("sm__inst_issued.avg.pct_of_peak_sustained_active", "issue slots %", 1.0),
       ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor pipe %", 1.0),
       ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps active %", 1.0), ("launch__registers_per_thread", "regs", 1.0),
       ("launch__grid_size", "grid", 1.0)]


def raw(path):
    rows = list(csv.reader(open(path)))
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    cols = [(k, lab) for k, lab, _ in RAW if k in idx]
    print("| kernel | " + " | ".join(lab for _, lab in cols) + " |\n|---|" + "---:|" * len(cols))
    for r in rows[2:]:
        name = r[idx["Kernel Name"]].replace("void ", "").split("(")[0]
        vals = []
        for k, _ in cols:
            v, u = r[idx[k]], units[idx[k]]
            try:
                f = float(v.replace(",", ""))
                if u == "byte":
                    f /= 1e9
                elif u in ("Kbyte",):
                    f /= 1e6
                elif u in ("Mbyte",):
                    f /= 1e3
                elif u in ("us", "usecond"):
                    f /= 1e3
                elif u in ("ns", "nsecond"):
                    f /= 1e6
                vals.append("%.2f" % f if f < 1000 else "%.0f" % f)
            except ValueError:
                vals.append(v)
        print("| `%s` | " % name[:40] + " | ".join(vals) + " |")


if __name__ == "__main__":
    {"launches": launches, "raw": raw}[sys.argv[1]](sys.argv[2])
