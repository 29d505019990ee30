"""Where does the end-to-end step's host overhead go?  Variants of [pinned q -> device | decode step | result -> pinned host | sync]."""
import os, sys, time, json, statistics
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import tree_attention_b200 as ta
from tree_attention_b200.models.decoder import TreeDecodeSession

def med(f, n=60):
    for _ in range(5): f()
    xs = []
    for _ in range(n):
        t0 = time.perf_counter(); f(); xs.append((time.perf_counter() - t0) * 1e6)
    return round(statistics.median(xs), 1)

def main():
    S = int(sys.argv[1]) if len(sys.argv) > 1 else 131072
    dev = torch.device("cuda")
    ta.setup(0, 1)
    _, k, v = ta.make_data((1, 32, S, 128), 0, dev, dtype=torch.bfloat16, log=False)
    q = torch.randn(1, 32, 1, 128, device=dev).bfloat16()
    sess = TreeDecodeSession([(k, v)], softmax_scale=0.088, pdl=True)
    sess.step_device(q, 0); torch.cuda.synchronize()
    qh = q.cpu().pin_memory(); oh = torch.empty_like(qh).pin_memory()
    st = torch.cuda.current_stream()
    res = {"seq": S}
    # device-only: back-to-back prepared launches
    def dev_loop():
        for i in range(50): sess.step_device(None, 0)
        torch.cuda.synchronize()
    res["dev_us_per_step_backtoback"] = round(med(dev_loop, 10) / 50, 1)
    res["step_e2e_graph"] = med(lambda: sess.step(qh, oh, 0))
    def graph_only():
        sess.graphs[0].replay(); st.synchronize()
    res["kernel_graph_replay_sync"] = med(graph_only)
    def launch_only():
        sess._steps[0].launch(0); st.synchronize()
    res["kernel_prepared_launch_sync"] = med(launch_only)
    def eager_copies():
        sess.q_static.copy_(qh, non_blocking=True); sess._steps[0].launch(0); oh.copy_(sess.out_static[0], non_blocking=True); st.synchronize()
    res["eager_h2d_launch_d2h_sync"] = med(eager_copies)
    def e2e_graph_raw():
        sess.e2e_graphs[0].replay(); st.synchronize()
    res["e2e_graph_replay_sync_only"] = med(e2e_graph_raw)
    ev = torch.cuda.Event()
    def e2e_graph_event():
        sess.e2e_graphs[0].replay(); ev.record(); ev.synchronize()
    res["e2e_graph_replay_event_sync"] = med(e2e_graph_event)
    def null_sync():
        st.synchronize()
    res["null_stream_sync"] = med(null_sync)
    print(json.dumps(res))
    ta.cleanup()

main()
