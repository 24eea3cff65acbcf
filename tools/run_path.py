#!/usr/bin/env python
"""Tiny driver for profiling: runs one piece of the hot path a few times (for ncu / compute-sanitizer).

    python tools/run_path.py gru_fwd|lstm_fwd|gru_train|lstm_train|gemm [--reps N] [--B 128]
"""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "icassp2022-depression_b200"))
import torch  # noqa: E402

import b200rnn  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("what")
ap.add_argument("--reps", type=int, default=3)
ap.add_argument("--B", type=int, default=128)
ap.add_argument("--T", type=int, default=0)
ap.add_argument("--time", action="store_true")
args = ap.parse_args()
dev = torch.device("cuda:0")
torch.manual_seed(0)
if args.what.startswith("gru"):
    m = b200rnn.GRU(256, 256, num_layers=2, dropout=0.3, batch_first=True).to(dev)
    x = torch.randn(args.B, args.T or 120, 256, device=dev)
elif args.what.startswith("lstm"):
    H = 256 if "256" in args.what else 128
    m = b200rnn.LSTM(1024, H, num_layers=2, dropout=0.3, bidirectional=True).to(dev)
    x = torch.randn(args.T or 30, args.B, 1024, device=dev)
else:
    m = None
if args.what == "gemm":
    a = torch.randn(15360, 256, device=dev)
    w = torch.randn(768, 256, device=dev)
    for _ in range(args.reps):
        b200rnn.gemm(a, w)
elif args.what.endswith("fwd"):
    m.train()
    with torch.no_grad():
        for _ in range(args.reps):
            m(x)
else:
    m.train()
    x.requires_grad_(True)
    for _ in range(args.reps):
        m.zero_grad()
        y = m(x)[0]
        y.sum().backward()
torch.cuda.synchronize()
if args.time:
    from b200rnn import _lib
    _lib.profile(True)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n = 20
    e0.record()
    if args.what.endswith("fwd"):
        with torch.no_grad():
            for _ in range(n):
                m(x)
    elif args.what != "gemm":
        for _ in range(n):
            m.zero_grad()
            m(x)[0].sum().backward()
    e1.record()
    torch.cuda.synchronize()
    out = {"what": args.what, "B": args.B, "ms_per_iter": e0.elapsed_time(e1) / n}
    for name, kind in (("rec_fwd", 0), ("rec_bwd", 1), ("gemm", 2)):
        ms, cnt = _lib.profile_read(kind)
        out[name] = f"{ms / max(cnt, 1) * 1e3:.1f} us x{cnt // n}/iter"
    _lib.profile(False)
    print(out)
print("done", args.what)
