#!/usr/bin/env python
"""Layer-forward time (input projection + recurrence, CUDA events) of the FFMA vs tensor-core recurrence by batch size.
Run twice: B200RNN_REC_TC=0 and B200RNN_REC_TC=1 (the switch is read once per process)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "icassp2022-depression_b200"))
import torch, b200rnn
dev = torch.device("cuda:0")
tag = os.environ.get("B200RNN_REC_TC", "0")
for kind in ("gru256", "lstm128bi"):
    for B in (4, 8, 15, 16, 30, 32, 45, 64, 128):
        if kind == "gru256":
            m = b200rnn.GRU(256, 256, num_layers=1, batch_first=True).to(dev).eval(); x = torch.randn(B, 120, 256, device=dev)
        else:
            m = b200rnn.LSTM(1024, 128, num_layers=1, bidirectional=True, batch_first=True).to(dev).eval(); x = torch.randn(B, 30, 1024, device=dev)
        with torch.no_grad():
            for _ in range(3):
                m(x)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(20):
                m(x)
            e1.record(); torch.cuda.synchronize()
        print(f"REC_TC={tag} {kind} B={B}: {e0.elapsed_time(e1) / 20 * 1e3:.1f} us/layer-forward", flush=True)
