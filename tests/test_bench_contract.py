"""The bench.py JSON-line contract, checked on CPU through the reference arm (the only arm that runs without a GPU):
one line on stdout, the required keys, and the reference-arm extras (impl, cpu_baseline, zero-byte e2e)."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_one_contract_line():
    env = dict(os.environ, OMP_NUM_THREADS="4")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "3"],
                       capture_output=True, text=True, timeout=600, env=env, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [ln for ln in r.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1, r.stdout
    d = json.loads(lines[0])
    for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
              "vs_baseline", "dtype", "data", "config", "impl", "cpu_baseline", "e2e"):
        assert k in d, k
    assert d["impl"] == "reference" and d["unit"] == "poses/s" and d["higher_is_better"] is True
    assert d["steps"] == 1 and d["warmup"] == 3 and d["value"] > 0
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "workload" in d["config"] and "model" not in d["config"]


def test_own_arm_fails_loudly_without_a_gpu():
    """No CPU fallback: on a machine without CUDA the product arm must exit non-zero instead of printing a number."""
    import torch
    if torch.cuda.is_available():
        return
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--steps", "1", "--warmup", "3", "--no-cpu-baseline"],
                       capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert r.returncode != 0
    assert not [ln for ln in r.stdout.splitlines() if ln.strip().startswith("{")]


def test_roofline_report_groups_classes_and_rejects_stale_traffic(tmp_path, monkeypatch):
    """Host logic of the roofline leg: launches are grouped per kernel class, each class is put on the roofline that
    bounds it, the headline is the class with the largest share, and an ncu capture whose launch count differs from
    the timed run is reported as `traffic: null` with the reason."""
    sys.path.insert(0, ROOT)
    import bench
    csv = tmp_path / "tc.csv"
    rows = ["layer,M,N,K_real,bn,ms,gflop,tflops,mbytes"]
    for step in range(2):
        for b in range(6):
            rows.append(f"layer3.{b}.conv2,327680,256,2304,256,0.270,386.5,1431,336.7")          # tensor-bound class
        for b in range(5):
            rows.append(f"layer3.{b}.conv3+next.conv1,327680,1024,256,256,0.400,343.6,859,1678.0")   # HBM-bound class
    csv.write_text("\n".join(rows) + "\n")
    peaks = {"bf16_tflops_sustained": 1416.7, "hbm_gbs": 6536.4, "source": "measured"}
    monkeypatch.setattr(bench, "load_traffic", lambda *a: None)
    r = bench.roofline_report(str(csv), 2, 9.0, peaks, 5, 64, 64, 2 * (6 * 0.27 + 5 * 0.4), 2 * (6 * 386.5e9 + 5 * 343.6e9), 22, 700.0, {})
    assert r["kernel"] == "layer3.x.conv3+next.conv1" and r["bound"] == "hbm" and r["unit"] == "GB/s"
    assert abs(r["achieved"] - 1678.0 / 0.4) < 1 and abs(r["frac"] - (1678.0 / 6536.4) / 0.4) < 1e-6
    assert r["traffic"] is None and "no ncu capture" in r["traffic_note"]
    by = {c["kernel"]: c for c in r["classes"]}
    assert by["layer3.x.conv2"]["bound"] == "tensor" and by["layer3.x.conv2"]["launches_per_step"] == 6
    assert abs(r["all_tc"]["launches_per_step"] - 11) < 1e-9
    stale = {"source": "x.csv", "tc_launches_per_step": 43, "tc_dram_bytes_per_step": 1.0,
             "classes": {"layer3.x.conv3+next.conv1": {"launches_per_step": 5, "dram_bytes_per_launch": 1.65e9}}}
    monkeypatch.setattr(bench, "load_traffic", lambda *a: stale)
    r = bench.roofline_report(str(csv), 2, 9.0, peaks, 5, 64, 64, 1.0, 1.0, 22, 700.0, {})
    assert r["traffic"] is None and "stale" in r["traffic_note"]
    stale["tc_launches_per_step"] = 11
    r = bench.roofline_report(str(csv), 2, 9.0, peaks, 5, 64, 64, 1.0, 1.0, 22, 700.0, {})
    assert r["traffic"] == 1.65e9 and abs(r["traffic_over_algorithmic"] - 1.65e9 / 1678e6) < 1e-9
