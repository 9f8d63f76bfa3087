"""bench.py's reference arm runs on the CPU: it must print exactly ONE JSON line on stdout with the contract's keys
(the B200 arm prints the same keys plus roofline / clocks / gpu_launches; it is exercised on the GPU box)."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_one_json_line():
    env = dict(os.environ, W2C_BENCH_IMG="128")   # small image: this is a contract test, not a measurement
    p = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1",
                        "--warmup", "1"], capture_output=True, text=True, timeout=600, env=env, cwd=ROOT)
    assert p.returncode == 0, p.stderr[-2000:]
    lines = [l for l in p.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, p.stdout
    d = json.loads(lines[0])
    for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
              "vs_baseline", "dtype", "data", "config", "impl", "cpu_baseline", "e2e"):
        assert k in d, k
    assert d["impl"] == "reference" and d["unit"] == "agent-frames/s" and d["higher_is_better"] is True
    # the UNMODIFIED reference when it can be found (/root/reference or the staged oracle/_ref copy), else the port
    from oracle import ref_harness
    assert d["cpu_baseline"]["kind"] == ("reference" if ref_harness.available() else "port")
    assert d["value"] > 0 and d["cpu_baseline"]["cores"] >= 1
    assert d["config"]["precision"] == "fp32" and d["config"]["cuda_graph"] is False
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0
    assert "workload" in d["config"]


def test_reference_arm_other_ranks_exit_quietly():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", W2C_BENCH_IMG="128")
    p = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2"],
                       capture_output=True, text=True, timeout=600, env=env, cwd=ROOT)
    assert p.returncode == 0 and p.stdout.strip() == ""


def test_reference_arm_runs_from_the_staged_copy(tmp_path):
    """oracle/make_ref.py stages a byte-identical copy of the reference package under oracle/_ref/ (git-ignored, it
    travels to the GPU box); the reference arm must run from it when /root/reference is absent."""
    from oracle import make_ref
    staged = make_ref.stage()
    if staged is None:
        staged = os.path.join(ROOT, "oracle", "_ref")
        if not os.path.isdir(os.path.join(staged, "ptsemseg")):
            import pytest
            pytest.skip("no reference tree and no staged copy on this machine")
    env = dict(os.environ, W2C_BENCH_IMG="128", W2C_REFERENCE_ROOT=staged)
    p = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1",
                        "--warmup", "0", "--batch", "1"], capture_output=True, text=True, timeout=600, env=env, cwd=ROOT)
    assert p.returncode == 0, p.stderr[-2000:]
    d = json.loads(p.stdout.strip())
    assert d["cpu_baseline"]["kind"] == "reference" and "staged" in d["cpu_baseline"]["sample"]
