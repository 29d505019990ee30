"""One regression test per row of SURVEY.md section 8 that tests/test_api_cpu.py does not already cover (D1, D2, D4-D7, D9 live
there): D3, D8, D10, D11, D12, D13, D14.  Each test asserts the NEW behaviour the table prescribes; `ref:` lines cite
/root/reference/model.py."""
import json
import os
import subprocess
import sys

import torch

from _dist_utils import run_distributed

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker_d3(rank, world):
    import tree_attention_b200 as ta
    from tree_attention_b200.ops import reference as ref

    q, k, v = ta.make_data((1, 2, 40, 16), rank, "cpu", dtype=torch.float32, log=False)
    out = ta.tree_decode(q, k, v, rank, world, torch.device("cpu"))        # ref: 105-116 raises here for world_size > 1
    import torch.distributed as dist

    ks, vs = [torch.empty_like(k) for _ in range(world)], [torch.empty_like(v) for _ in range(world)]
    dist.all_gather(ks, k)
    dist.all_gather(vs, v)
    exp, _ = ref.attention_ref(q, torch.cat(ks, 2), torch.cat(vs, 2), softmax_scale=1.0)
    assert torch.allclose(out.double(), exp, atol=1e-5)


def test_D3_distributed_branch_runs_for_world_size_2(port):
    """ref: 111-112 -- the 4-D `.unsqueeze(-1).expand_as(4-D)` makes the distributed branch raise; here it computes."""
    run_distributed(_worker_d3, 2, port)


def test_D8_collective_payload_is_O_plus_two_scalars_per_row(monkeypatch):
    """ref: 103, 108, 114, 115 -- lse expanded to |O| and three full-size all-reduces = 3|O| on the wire.  The kept
    three-all-reduce baseline moves |O| + 2 scalars per row; the packed all-gather |O| + 1."""
    import torch.distributed as dist
    from tree_attention_b200.parallel import tree

    moved = []
    monkeypatch.setattr(dist, "all_reduce", lambda t, op=None, group=None: moved.append(t.numel()))
    b, h, sq, d = 2, 4, 3, 16
    o, lse = torch.randn(b, h, sq, d), torch.randn(b, h, sq)
    tree._combine_allreduce3(o, lse, None)
    rows = b * h * sq
    assert sorted(moved) == sorted([rows, rows * d, rows])          # max, numerator, denominator
    assert sum(moved) == rows * (d + 2) < 3 * rows * d


def _run_cli(tmp_path, *extra):
    env = dict(os.environ, CUDA_VISIBLE_DEVICES="")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "model.py"), "--seq-len", "256", "--num-heads", "2", "--json",
                        "--log-file", str(tmp_path / "log.log"), *extra],
                       capture_output=True, text=True, timeout=300, env=env, cwd=str(tmp_path))
    assert r.returncode == 0, r.stderr[-2000:]
    return json.loads([ln for ln in r.stdout.splitlines() if ln.startswith("{")][-1]), r


def test_D10_cli_times_after_warmup_and_reports_per_step_latency(tmp_path):
    """ref: 149-151 -- one cold, unsynchronised call timed with time.time().  Here: W warm-up steps, K timed steps, the
    latency reported per step (CUDA events on GPUs, a synchronised host clock on CPU)."""
    d, _ = _run_cli(tmp_path, "--steps", "3", "--warmup", "2")
    assert d["steps"] == 3 and d["warmup"] == 2 and d["timer"] == "host_clock"
    assert d["latency_us"] > 0 and d["kv_tokens_per_s"] > 0


def _worker_d11(rank, world):
    import torch.distributed as dist

    assert dist.get_backend() == "gloo" and dist.get_world_size() == world     # ref: 19-22 -- NCCL only, CPU cannot be multi-rank


def test_D11_cpu_multi_rank_over_gloo_with_a_configurable_port(port):
    """ref: 19-22, 105 -- backend, port (12355) and device gating are hard-coded.  `_dist_utils` calls
    `setup(rank, world, master_addr=..., master_port=<free port>)` on CPU ranks."""
    assert port != 12355
    run_distributed(_worker_d11, 2, port)


def test_D12_cli_validates_its_output_against_the_oracle(tmp_path):
    """ref: 150 -- the result is never looked at.  The CLI checks it against the fp64 oracle (and can be told not to)."""
    d, _ = _run_cli(tmp_path)
    assert d["max_abs_err"] is not None and d["max_abs_err"] < 1e-4
    d2, _ = _run_cli(tmp_path, "--no-check")
    assert d2.get("max_abs_err") is None


def test_D13_declared_dependencies_are_minimal_and_sufficient():
    """ref: requirements.txt:1-3 lists swarms / zetascale (unused) and omits loguru (imported at model.py:5)."""
    req = open(os.path.join(ROOT, "requirements.txt")).read().lower()
    deps = [ln.split("#")[0].strip() for ln in req.splitlines()]
    deps = [d for d in deps if d]
    assert not any(d.startswith(("swarms", "zetascale")) for d in deps)
    assert any(d.startswith("torch") for d in deps)
    # loguru is optional: the package imports (and logs) without it
    code = ("import sys; sys.modules['loguru'] = None\n"
            "import tree_attention_b200 as ta\n"
            "from tree_attention_b200.utils.logging import logger\n"
            "logger.info('ok'); print(type(logger).__name__)")
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=300, cwd=ROOT,
                       env=dict(os.environ, CUDA_VISIBLE_DEVICES=""))
    assert r.returncode == 0 and "_StdLogger" in r.stdout, r.stderr[-1500:]


def test_D14_nothing_is_logged_from_inside_the_op(capfd):
    """ref: 82, 116, 122 -- three log lines per call from inside flash_res_lse / tree_decode (host I/O on the hot path)."""
    import tree_attention_b200 as ta

    q, k, v = ta.make_data((1, 2, 64, 16), 0, "cpu", dtype=torch.float32, log=False)
    capfd.readouterr()
    ta.flash_res_lse(q, k, v)
    ta.tree_decode(q, k, v, 0, 1, torch.device("cpu"))
    ta.tree_attention(q, k, v, causal=True)
    out, err = capfd.readouterr()
    assert out == "" and err == ""
