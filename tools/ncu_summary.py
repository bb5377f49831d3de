"""Summarise gpurun_out/ncu/*.ncu-rep into profiles/<round>/ncu_summary.md (+ one key-metric CSV per report)."""
import csv
import glob
import io
import os
import subprocess
import sys

KEYS = [
    ("gpu__time_duration.sum", "duration"),
    ("dram__bytes_read.sum", "dram_read"),
    ("dram__bytes_write.sum", "dram_write"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram_pct"),
    ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor_pipe_pct"),
    ("sm__inst_executed_pipe_tensor.sum", "tensor_inst"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm_pct"),
    ("lts__t_sector_hit_rate.pct", "l2_hit_pct"),
    ("launch__registers_per_thread", "regs"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps_active_pct"),
    ("launch__grid_size", "grid"),
    ("launch__block_size", "block"),
]


def to_bytes(v, unit):
    mult = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}
    return float(v) * mult.get(unit, 1)


def to_us(v, unit):
    return float(v) * {"ns": 1e-3, "us": 1, "ms": 1e3, "s": 1e6}.get(unit, 1)


def main():
    src = sys.argv[1] if len(sys.argv) > 1 else "gpurun_out/ncu"
    dst = sys.argv[2] if len(sys.argv) > 2 else "profiles/r01"
    os.makedirs(dst, exist_ok=True)
    rows = []
    for rep in sorted(glob.glob(os.path.join(src, "*.ncu-rep"))):
        name = os.path.basename(rep)[:-8]
        out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
        rd = list(csv.reader(io.StringIO(out)))
        if len(rd) < 3:
            continue
        hdr, unit, val = rd[0], rd[1], rd[2]
        col = {h: i for i, h in enumerate(hdr)}
        rec = {"name": name, "kernel": val[col["Kernel Name"]][:70] if "Kernel Name" in col else ""}
        extra = [h for h in hdr if "tensor" in h and "pct_of_peak_sustained_active" in h and "avg" in h]
        with open(os.path.join(dst, f"ncu_{name}_keymetrics.csv"), "w") as f:
            for h in [k for k, _ in KEYS] + extra:
                if h in col:
                    f.write(f"{h},{unit[col[h]]},{val[col[h]]}\n")
        for k, short in KEYS:
            if k in col:
                v, u = val[col[k]].replace(",", ""), unit[col[k]]
                if short == "duration":
                    rec[short] = to_us(v, u)
                elif short in ("dram_read", "dram_write"):
                    rec[short] = to_bytes(v, u)
                else:
                    try:
                        rec[short] = float(v)
                    except ValueError:
                        rec[short] = v
        best = 0.0
        for h in extra:
            try:
                best = max(best, float(val[col[h]]))
            except ValueError:
                pass
        rec["tensor_any_pct"] = best
        rows.append(rec)
    with open(os.path.join(dst, "ncu_summary.md"), "w") as f:
        f.write("| capture | kernel | time (us) | DRAM read+write (MB) | DRAM GB/s | DRAM % of ncu peak | tensor pipe % | L2 hit % | regs | grid x block |\n")
        f.write("|---|---|---|---|---|---|---|---|---|---|\n")
        for r in rows:
            tr = (r.get("dram_read", 0) + r.get("dram_write", 0))
            gbs = tr / (r["duration"] * 1e-6) * 1e-9 if r.get("duration") else 0
            f.write(f"| {r['name']} | `{r['kernel']}` | {r.get('duration', 0):.1f} | {tr * 1e-6:.1f} | {gbs:.0f} | "
                    f"{r.get('dram_pct', 0):.1f} | {max(r.get('tensor_pipe_pct', 0) or 0, r['tensor_any_pct']):.1f} | {r.get('l2_hit_pct', 0):.1f} | "
                    f"{r.get('regs', 0):.0f} | {r.get('grid', 0):.0f} x {r.get('block', 0):.0f} |\n")
    print(open(os.path.join(dst, "ncu_summary.md")).read())


if __name__ == "__main__":
    main()
