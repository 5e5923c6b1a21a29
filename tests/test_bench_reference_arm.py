"""The CPU arm of bench.py (`--impl reference`) runs without a GPU: one JSON line with the contract's keys."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_the_contract_line():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1",
                          "--warmup", "1", "--pairs", "4"], capture_output=True, text=True, timeout=900, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    line = json.loads(out.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["unit"] == "hypotheses/s" and line["higher_is_better"] is True
    assert line["value"] > 0 and line["steps"] == 1 and line["warmup"] == 1 and line["dtype"] == "f64"
    assert line["e2e"] == {"value": line["value"], "unit": "hypotheses/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    cb = line["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] == line["value"] and "C2" in cb["sample"]
    assert "workload" in line["config"]
    nat = line["cpu_baseline_native"]  # the -march=native build beside the reference-flags build
    assert nat["kind"] == "port" and nat["value"] > 0 and "-march=native" in nat["flags"]
    assert nat["same_trajectory_as_reference_flags_build"] == f"{line['config']['pairs_per_step']}/{line['config']['pairs_per_step']}"
    rs = line.get("reference_sources")
    if rs is not None:  # present when oracle/_ref/libplref2.so exists: the reference's own sources take the same path
        n = rs["problems"]
        assert rs["kind"] == "reference" and rs["value"] > 0 and rs["same_trajectory_as_port"] == f"{n}/{n}"
