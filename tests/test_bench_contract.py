"""bench.py contract checks that need no GPU: the reference arm prints ONE JSON line with the agreed keys, and the
workload table names BASELINE.json's configurations."""
import json
import os
import subprocess
import sys

from conftest import ROOT


def test_reference_arm_json_line():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--workload", "cfg1",
                          "--steps", "1", "--warmup", "1"], capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    for key in ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
                "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e"):
        assert key in d, key
    assert d["impl"] == "reference" and d["dtype"] == "f64" and d["higher_is_better"] is True and d["vs_baseline"] is None
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "workload" in d["config"] and d["value"] > 0


def test_workloads_match_baseline_json():
    sys.path.insert(0, ROOT)
    import bench
    base = json.load(open(os.path.join(ROOT, "BASELINE.json")))
    w4 = bench.WORKLOADS["cfg4"]
    assert (w4["N"], w4["d"], w4["encoding"], w4["kernel"], w4["q"], w4["layers"], w4["outer"], w4["agents"]) == \
        (65536, 4, "yz_cx", "projected", 8, 3, "gaussian", 8)
    assert "N=65536" in base["configs"][3] and "yz_cx" in base["configs"][3]
    w5 = bench.WORKLOADS["cfg5"]
    assert (w5["N"], w5["q"], w5["layers"], w5["agents"], w5["encoding"]) == (131072, 10, 4, 16, "kyriienko")
    w3 = bench.WORKLOADS["cfg3"]
    assert (w3["N"], w3["q"], w3["layers"], w3["agents"], w3["kernel"]) == (16384, 5, 2, 8, "fidelity")
    assert bench.flops_per_entry(w4) == 74 and bench.flops_per_entry(w3) == 259
