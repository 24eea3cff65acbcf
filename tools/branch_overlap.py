#!/usr/bin/env python
"""How much of the text branch does the fuse step hide beside the audio branch? (one GPU, CUDA-graph replay)

Times, on the same box and the same inputs as bench.py: the audio branch alone (LN-split -> GEMM -> GRU -> dropout-split
-> GEMM -> GRU with the time-sum epilogue), the text branch alone (split -> GEMMs -> BiLSTM x2 + the text half of the
head), and the whole step on one stream and on two. `whole - audio` is what the text branch still costs.

    python tools/branch_overlap.py [--steps 200]
"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "icassp2022-depression_b200"))
os.environ.setdefault("OMP_NUM_THREADS", "1")

import torch  # noqa: E402

import b200rnn  # noqa: E402
import bench  # noqa: E402


def timed(fn, steps, warmup=20):
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        for _ in range(3):
            fn(0)
    torch.cuda.current_stream().wait_stream(side)
    torch.cuda.synchronize()
    graphs = []
    pool = None
    for i in range(bench.N_ROTATE):
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, pool=pool):
            fn(i)
        pool = g.pool()
        graphs.append(g)
    for i in range(warmup):
        graphs[i % len(graphs)].replay()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for i in range(steps):
        graphs[i % len(graphs)].replay()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / steps


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=200)
    args = ap.parse_args()
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(dev)
    torch.set_num_threads(1)
    torch.manual_seed(0)
    model = b200rnn.fusion_net(**bench.FUSE_ARGS).to(dev)
    for p in model.parameters():
        p.requires_grad = False
    model.fc_final[0].weight.requires_grad = True
    model.train()
    host = [bench._synthetic(bench.B_PER_GPU, 1234 + i) for i in range(bench.N_ROTATE)]
    dev_in = [(a.to(dev), t.to(dev), y.to(dev)) for a, t, y in host]
    out = {}
    for name, concurrent in (("two_streams", True), ("one_stream", False)):
        fused = b200rnn.FusedFuseStep(model, lr=bench.LR, exchange="none", concurrent_branches=concurrent)
        out[f"whole_step_{name}_ms"] = timed(
            lambda i, f=fused: f(b200rnn.FuseBatch(dev_in[i][0], dev_in[i][1]), dev_in[i][2]), args.steps)
    fused = b200rnn.FusedFuseStep(model, lr=bench.LR, exchange="none", concurrent_branches=False)
    with torch.no_grad():
        out["audio_branch_alone_ms"] = timed(
            lambda i: fused._audio_branch(b200rnn.FuseBatch(dev_in[i][0], dev_in[i][1])), args.steps)
        tf = torch.empty(bench.B_PER_GPU, bench.H_TEXT, device=dev)

        def text_only(i):
            seq, h_n = fused._text_branch(b200rnn.FuseBatch(dev_in[i][0], dev_in[i][1]))
            a0 = fused._args(seq, h_n, None, tf, None)
            a0.rng_state = fused.rng_state.data_ptr()
            from b200rnn import _lib
            from b200rnn.functional import _on, _stream_ptr
            import ctypes
            with _on(dev):
                _lib.check(_lib.load().b200rnn_fuse_head(ctypes.byref(a0), _stream_ptr(dev)), "text stage")

        out["text_branch_alone_ms"] = timed(text_only, args.steps)
    out["text_cost_beside_audio_ms"] = out["whole_step_two_streams_ms"] - out["audio_branch_alone_ms"]
    out = {k: round(v, 4) for k, v in out.items()}
    print(json.dumps(out))


if __name__ == "__main__":
    main()
