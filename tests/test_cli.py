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


def test_cli_under_torchrun_two_cpu_ranks(tmp_path):
    """`torchrun model.py`: the workers must rendezvous on the LAUNCHER's store, not on the config's default port
    (a port override hung the 8-GPU `torchrun model.py --json` run of round 2 until its timeout)."""
    import socket

    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    env = dict(os.environ, CUDA_VISIBLE_DEVICES="")
    for k in ("RANK", "WORLD_SIZE", "LOCAL_RANK", "MASTER_ADDR", "MASTER_PORT", "TORCHELASTIC_RUN_ID"):
        env.pop(k, None)
    r = subprocess.run(
        [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
         "--master-port", str(port), os.path.join(ROOT, "model.py"), "--seq-len", "256", "--num-heads", "4", "--json",
         "--log-file", str(tmp_path / "log.log")],
        capture_output=True, text=True, timeout=240, env=env, cwd=str(tmp_path),
    )
    assert r.returncode == 0, r.stderr[-2000:]
    d = json.loads([l for l in r.stdout.splitlines() if l.startswith("{")][-1])
    assert d["world_size"] == 2 and d["shape"]["S_global"] == 512 and d["max_abs_err"] < 1e-4
