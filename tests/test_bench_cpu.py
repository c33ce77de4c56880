"""bench.py's reference arm (`--impl reference`) runs without a GPU: one JSON line on stdout with the contract's keys,
nothing from the other ranks under torchrun.  Tiny ad-hoc shape so the CPU suite stays short."""
import json
import os
import subprocess
import sys

from conftest import ROOT


def _run(extra_env):
    env = dict(os.environ, **extra_env)
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--shape", "600,800,10",
                           "--steps", "2", "--warmup", "1"], capture_output=True, text=True, env=env, timeout=300)


def test_reference_arm_prints_one_json_line():
    proc = _run({})
    assert proc.returncode == 0, proc.stderr[-2000:]
    lines = [l for l in proc.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, proc.stdout
    line = json.loads(lines[0])
    assert line["impl"] == "reference" and line["metric"] == "nmf_iterations_per_second" and line["unit"] == "iter/s"
    assert line["value"] > 0 and line["higher_is_better"] is True and line["steps"] == 2 and line["warmup"] == 1
    assert line["cpu_baseline"]["kind"] == "reference" and line["cpu_baseline"]["cores"] >= 1
    assert line["cpu_baseline"]["value"] == line["value"] and "sklearn" in line["cpu_baseline"]["sample"]
    assert line["e2e"] == {"value": line["value"], "unit": "iter/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert line["vs_baseline"] is None and line["data"] == "synthetic" and "workload" in line["config"]


def test_reference_arm_other_ranks_exit_quietly():
    proc = _run({"RANK": "1", "WORLD_SIZE": "2", "LOCAL_RANK": "1"})
    assert proc.returncode == 0 and proc.stdout.strip() == ""


def test_roofline_traffic_is_read_from_the_tracked_ncu_summary():
    """bench.py takes roofline.traffic (dram bytes of one X-streaming launch) from the newest tracked
    `profiles/*ncu_full_C3.csv`, not from a typed-in constant: ~4 m n bytes + the factor re-reads."""
    sys.path.insert(0, ROOT)
    import bench
    cap = bench.ncu_capture("C3")
    assert cap is not None and cap["source"].startswith("profiles/") and cap["source"].endswith("ncu_full_C3.csv")
    m, n, _k, _d = bench.WORKLOADS["C3"]
    assert 4.0 * m * n <= cap["traffic"] <= 1.05 * 4.0 * m * n
    assert 0.5 < cap["tensor_pipe_active"] <= 1.0 and cap["launches"] >= 1
    assert bench.ncu_capture("no-such-workload") is None


def test_c4_stream_subject_distribution_matches_run_locally():
    """The C4 stream hands subject s to rank s % world -- the distribution run_locally(rank, world) uses -- and its
    default subject counts are BASELINE's 100 at N > 1 and one rank's share of it at N = 1."""
    src = open(os.path.join(ROOT, "bench.py")).read()
    assert "mine = list(range(rank, n_subjects, world))" in src
    assert "100 if world > 1 else 13" in src
    for world in (2, 4, 8):
        shares = [len(range(r, 100, world)) for r in range(world)]
        assert sum(shares) == 100 and max(shares) - min(shares) <= 1
    assert len(range(0, 100, 8)) == 13
