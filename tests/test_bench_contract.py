"""bench.py contract checks that need no GPU: the reference arm prints exactly ONE JSON line on stdout with the keys the
driver reads, whatever the launcher exported into the environment."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_one_json_line():
    env = dict(os.environ, OMP_NUM_THREADS="1")     # what torchrun exports to its workers
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "1", "--steps", "1",
                        "--warmup", "0"], capture_output=True, text=True, env=env, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [ln for ln in r.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1, r.stdout
    d = json.loads(lines[0])
    for key in ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
                "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e"):
        assert key in d, key
    assert d["impl"] == "reference" and d["unit"] == "Mpix/s" and d["value"] > 0 and d["vs_baseline"] is None
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["value"] == d["value"]
    assert "workload" in d["config"]
    # the keys the driver compares between the two arms (same workload, same sizes)
    assert all(d["config"].get(k) == v for k, v in (("n_gaussians", 1_000_000), ("width", 1920), ("height", 1080), ("sh_k", 16)))
    assert "physical cores" in d["cpu_baseline"]["sample"]


def test_reference_arm_other_ranks_exit_quietly():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1",
                        "--warmup", "0"], capture_output=True, text=True, env=env, timeout=120)
    assert r.returncode == 0 and r.stdout.strip() == ""


def test_matrix_configs_follow_the_reference_bench_sizes():
    """crates/brush-bench-test/src/benches.rs:222-287: {0.5, 1, 2.5} M splats at 1080p, 2 M at four resolutions, SH degree 0."""
    import importlib.util
    spec = importlib.util.spec_from_file_location("bench_mod", os.path.join(ROOT, "bench.py"))
    bench = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(bench)
    sizes = {(n, w, h) for _, n, w, h in bench.MATRIX}
    assert {(500_000, 1920, 1080), (1_000_000, 1920, 1080), (2_500_000, 1920, 1080)} <= sizes
    assert {(2_000_000, 1024, 1024), (2_000_000, 1920, 1080), (2_000_000, 2560, 1440), (2_000_000, 3200, 1800)} <= sizes
    for i, (_, n, w, h) in enumerate(bench.MATRIX):
        c = bench.CONFIGS[100 + i]
        assert (c["n"], c["w"], c["h"], c["k"]) == (n, w, h, 1) and c["forward_only"] and c["isect_cap"] >= 16 * n
    # the BASELINE configs are untouched by the matrix entries
    assert bench.CONFIGS[1] == dict(n=1_000_000, w=1920, h=1080, seed=0xB2000001, shift=0.0)
    assert bench.base_config()["sh_k"] == 16
