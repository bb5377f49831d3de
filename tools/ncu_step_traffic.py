"""DRAM traffic per tcgen05 kernel class from an ncu launch list of whole bench steps.

    ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv \
        --log-file gpurun_out/ncu_step.csv python bench.py --steps 1 --warmup 3 --ramp-seconds 0 --no-e2e --no-eager \
        --no-latency --no-cpu-baseline --no-clocks
    python tools/ncu_step_traffic.py gpurun_out/ncu_step.csv gpurun_out/tc_launches.csv profiles/r02/tc_traffic.json [steps_in_capture=5]

The capture holds `steps_in_capture` identical steps (3 warm-up + 1 timed + 1 per-launch-event pass); the LAST one is
zipped, launch by launch, with the per-launch event CSV bench.py wrote for that same pass (same order), which names the
layer of every tcgen05 launch.  Output: launches per step, DRAM bytes per step, and per kernel class the launches per
step and mean DRAM bytes per launch - what bench.py's `roofline.traffic` reads (and refuses when the launch count of a
later build differs)."""
import csv
import json
import os
import re
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bench import kernel_class  # noqa: E402

TC = re.compile(r"conv_gemm_tc_kernel|bottleneck_tail_kernel|bottleneck_next_kernel|bottleneck_block_kernel")


def main():
    ncu_csv, ev_csv, out = sys.argv[1], sys.argv[2], sys.argv[3]
    nsteps = int(sys.argv[4]) if len(sys.argv) > 4 else 5
    launches = {}
    with open(ncu_csv) as f:
        rows = [ln for ln in f if ln.startswith('"')]
    for r in csv.DictReader(rows):
        d = launches.setdefault(int(r["ID"]), {"name": r["Kernel Name"]})
        v = float(r["Metric Value"].replace(",", ""))
        unit = r["Metric Unit"]
        if r["Metric Name"] == "gpu__time_duration.sum":
            d["us"] = v * {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(unit, 1.0)
        else:
            d[r["Metric Name"]] = v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(unit, 1)
    order = [launches[k] for k in sorted(launches) if "hmv::" in launches[k]["name"]]     # torch's own kernels (input generation, isfinite) are not part of a step
    assert len(order) % nsteps == 0, f"{len(order)} launches do not split into {nsteps} identical steps"
    per = len(order) // nsteps
    step = order[-per:]
    tc = [k for k in step if TC.search(k["name"])]
    with open(ev_csv) as f:
        ev = list(csv.DictReader(f))
    assert len(ev) == len(tc), f"event CSV has {len(ev)} launches, the ncu step has {len(tc)} tcgen05 launches"
    classes = {}
    for k, e in zip(tc, ev):
        c = classes.setdefault(kernel_class(e["layer"]), {"launches_per_step": 0, "dram_bytes": 0.0, "us": 0.0, "kernel": k["name"][:80]})
        c["launches_per_step"] += 1
        c["dram_bytes"] += k.get("dram__bytes_read.sum", 0.0) + k.get("dram__bytes_write.sum", 0.0)
        c["us"] += k["us"]
    for c in classes.values():
        c["dram_bytes_per_launch"] = c["dram_bytes"] / c["launches_per_step"]
        c["ncu_us_per_launch"] = c["us"] / c["launches_per_step"]
    tot = lambda ks: sum(k.get("dram__bytes_read.sum", 0.0) + k.get("dram__bytes_write.sum", 0.0) for k in ks)
    res = {"source": f"{ncu_csv} (ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none, "
                     f"last of {nsteps} B=64 steps)",
           "launches_per_step": per, "tc_launches_per_step": len(tc), "tc_dram_bytes_per_step": tot(tc), "all_dram_bytes_per_step": tot(step),
           "tc_ms_per_step_ncu": sum(k["us"] for k in tc) * 1e-3, "step_ms_ncu_serialised": sum(k["us"] for k in step) * 1e-3,
           "classes": classes}
    res["tc_share_of_step_ncu"] = res["tc_ms_per_step_ncu"] / res["step_ms_ncu_serialised"]
    with open(out, "w") as f:
        json.dump(res, f, indent=1)
    print(json.dumps({k: v for k, v in res.items() if k != "classes"}, indent=1))
    for n, c in sorted(classes.items(), key=lambda kv: -kv[1]["us"]):
        print(f"{n:40s} x{c['launches_per_step']:2d}  {c['ncu_us_per_launch']:8.1f} us  {c['dram_bytes_per_launch'] * 1e-6:9.1f} MB/launch")


if __name__ == "__main__":
    main()
