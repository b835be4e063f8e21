"""bench.py --impl reference (the CPU arm the driver runs next to ours): one JSON line on stdout with the contract's
keys, exit code 0 -- also when the reference's pump threads are still alive at exit.  Small deque, one step: seconds."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_one_contract_line_and_exits_cleanly():
    cmd = [sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "1",
           "--cpu-entries", "4096", "--cpu-distinct", "256"]
    res = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert res.returncode == 0, res.stderr[-2000:]
    lines = [l for l in res.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, res.stdout[-2000:]
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "replay_sample_target_transitions_per_sec" and d["unit"] == "transitions/s"
    assert d["higher_is_better"] is True and d["n_gpus"] == 1 and d["steps"] == 1 and d["warmup"] == 1 and d["value"] > 0
    assert d["cpu_baseline"]["kind"] in ("reference", "port") and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["sample"]
    assert d["cpu_baseline"]["value"] == d["value"]
    e = d["e2e"]
    assert e["value"] == d["value"] and e["unit"] == d["unit"] and e["h2d_bytes_per_step"] == 0 and e["d2h_bytes_per_step"] == 0
    # the same config keys as our arm prints (bench.shared_config), so the driver's same_config check holds
    sys.path.insert(0, ROOT)
    import bench
    args = type("A", (), dict(workload="c51_b32", learner_steps=20, actions=4, cpu_entries=4096, ring=1_000_000, total_ring=0))()
    ours = bench.shared_config(bench.WORKLOADS["c51_b32"], 20, 4, 1_000_000, args)
    assert set(d["config"]) == set(ours) and d["config"]["workload"] == ours["workload"]
