"""bench.py's reference arm (`--impl reference`): the JSON line the driver parses, produced here on the CPU -- the unmodified
reference (oracle/_ref, compiled from /root/reference by oracle/Makefile) or, where it was not built, the oracle port, on the same
synthetic slots as the CUDA arm.  No GPU and nothing of libft8b200 is involved in this arm (bench.py:reference_arm)."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(extra_env=None, args=()):
    env = dict(os.environ)
    env.pop("RANK", None); env.pop("WORLD_SIZE", None)
    env.update(extra_env or {})
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "1", *args],
                          capture_output=True, text=True, timeout=600, env=env, cwd=ROOT)


def test_reference_arm_prints_one_json_line_with_the_contract_keys():
    p = _run()
    assert p.returncode == 0, p.stderr[-800:]
    lines = [l for l in p.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, "stdout must carry the JSON line only"
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and "unavailable" not in d
    assert d["metric"].startswith("FT8 15s-slots decoded/sec") and d["unit"] == "slots/s" and d["higher_is_better"] is True
    assert d["value"] > 0 and d["ms_per_step"] > 0 and d["n_gpus"] == 1 and d["steps"] == 1
    assert d["vs_baseline"] is None and d["scaling"] == "weak" and d["data"].startswith("synthetic")
    assert set(d["config"]) >= {"workload", "slots_per_gpu_per_step", "max_candidates", "max_messages", "ldpc_iterations"} and "model" not in d["config"]
    cb = d["cpu_baseline"]
    assert cb["kind"] in ("reference", "port") and cb["cores"] >= 1 and cb["value"] == d["value"] and cb["sample"]
    e = d["e2e"]
    assert e["value"] == d["value"] and e["unit"] == d["unit"] and e["h2d_bytes_per_step"] == 0 and e["d2h_bytes_per_step"] == 0
    if os.path.isdir("/root/reference"):   # built in this container by __graft_entry__.build(): the unmodified reference is what is timed
        assert cb["kind"] == "reference"


def test_reference_arm_runs_on_rank_0_only():
    p = _run({"RANK": "1", "WORLD_SIZE": "2", "LOCAL_RANK": "1", "MASTER_ADDR": "127.0.0.1", "MASTER_PORT": "29577"}, ("--gpus", "2"))
    assert p.returncode == 0 and p.stdout.strip() == "", (p.stdout[-200:], p.stderr[-400:])
