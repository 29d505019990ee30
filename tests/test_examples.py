"""The examples run end to end on CPU (gloo for world > 1) and check themselves against single-process references."""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(args, port):
    env = dict(os.environ, CUDA_VISIBLE_DEVICES="")
    r = subprocess.run([sys.executable] + args + ["--port", str(port)], capture_output=True, text=True, timeout=600, env=env,
                       cwd=ROOT)
    assert r.returncode == 0, (r.stdout + r.stderr)[-3000:]
    return r.stdout + r.stderr


def test_context_parallel_training_example(port):
    out = _run([os.path.join(ROOT, "examples", "train_context_parallel.py"), "--world", "2", "--steps", "2",
                "--q-tokens", "32", "--ctx-tokens", "96", "--embed", "128"], port)
    assert "gradients match (2 rank(s))" in out


def test_context_parallel_training_example_causal_zigzag(port):
    out = _run([os.path.join(ROOT, "examples", "train_context_parallel.py"), "--world", "2", "--steps", "2", "--causal",
                "--ctx-tokens", "64", "--embed", "128"], port)
    assert "gradients match (2 rank(s))" in out


def test_decode_server_example_quantised_cache(port):
    out = _run([os.path.join(ROOT, "examples", "decode_server.py"), "--tokens-per-rank", "200", "--steps", "3",
                "--kv-format", "mxfp8", "--kv-heads", "2"], port)
    assert "decode-attention steps" in out and "kv=mxfp8" in out
