"""bench.py contract checks that need no GPU: the reference arm prints ONE JSON line with the agreed keys."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_one_json_line():
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "1"],
                       capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, lines
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "rows/s" and d["higher_is_better"] is True
    assert d["metric"] == "RealNVP fit rows/sec (fwd+bwd+Adam)" and d["value"] > 0
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": "rows/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert d["config"]["workload"].startswith("c3")


def test_flops_per_row_matches_the_survey_table():
    sys.path.insert(0, ROOT)
    from bench import flops_per_row, WORKLOADS
    want = {"c2": (960, 2520), "c3": (327680, 909312), "c4": (983040, 2736128), "c5": (2621440, 7208960)}
    for name, (fwd, fit) in want.items():
        D, Cd, L, hidden, _, _ = WORKLOADS[name]
        assert flops_per_row(D, Cd, L, hidden[0]) == (fwd, fit), name
