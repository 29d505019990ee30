"""`python3 model.py` UX parity (README.md:12-14 of the reference): zero-arg entry point, CPU branch."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_cli_cpu_small(tmp_path):
    env = dict(os.environ, CUDA_VISIBLE_DEVICES="")
    r = subprocess.run(
        [sys.executable, os.path.join(ROOT, "model.py"), "--seq-len", "512", "--num-heads", "4", "--json",
         "--log-file", str(tmp_path / "log.log")],
        capture_output=True, text=True, timeout=300, env=env, cwd=str(tmp_path),
    )
    assert r.returncode == 0, r.stderr[-2000:]
    line = [l for l in r.stdout.splitlines() if l.startswith("{")][-1]
    d = json.loads(line)
    assert d["world_size"] == 1 and d["max_abs_err"] < 1e-4
    assert "Computation completed in" in r.stderr
    assert (tmp_path / "log.log").exists()


def test_cli_quantised_kv_formats(tmp_path):
    """--kv-format: the same entry point over an fp8 KV cache (CPU: de-quantising fallback), checked against the oracle."""
    env = dict(os.environ, CUDA_VISIBLE_DEVICES="")
    for fmt in ("fp8", "mxfp8", "mxfp8-simt"):
        r = subprocess.run(
            [sys.executable, os.path.join(ROOT, "model.py"), "--seq-len", "300", "--num-heads", "4", "--num-kv-heads", "2",
             "--json", "--kv-format", fmt, "--log-file", str(tmp_path / "log.log")],
            capture_output=True, text=True, timeout=300, env=env, cwd=str(tmp_path),
        )
        assert r.returncode == 0, r.stderr[-2000:]
        d = json.loads([l for l in r.stdout.splitlines() if l.startswith("{")][-1])
        assert d["kv_format"] == fmt and d["max_abs_err"] < 1e-3
