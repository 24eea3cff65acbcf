#!/usr/bin/env python
"""Small shapes through every recurrence variant, for compute-sanitizer (memcheck / racecheck):
    compute-sanitizer --tool memcheck python tools/sanitize_paths.py
    B200RNN_REC_TC=1 compute-sanitizer --tool racecheck python tools/sanitize_paths.py
"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "icassp2022-depression_b200"))
import torch, b200rnn
from torch.nn.utils.rnn import pack_padded_sequence
dev = torch.device("cuda:0")
torch.manual_seed(0)
for kind, I, H, L, bi in (("gru", 64, 256, 2, False), ("lstm", 64, 128, 2, True), ("gru", 32, 128, 1, True), ("lstm", 32, 256, 1, False)):
    cls = b200rnn.GRU if kind == "gru" else b200rnn.LSTM
    m = cls(I, H, num_layers=L, bidirectional=bi, batch_first=True, dropout=0.3 if L > 1 else 0.0).to(dev).train()
    B, T = 9, 6
    x = torch.randn(B, T, I, device=dev, requires_grad=True)
    y = m(x)[0]
    y.sum().backward()
    lengths = torch.tensor([6, 1, 3, 6, 2, 5, 4, 6, 1])
    xp = pack_padded_sequence(x.detach().requires_grad_(True), lengths, batch_first=True, enforce_sorted=False)
    yp = m(xp)[0]
    yp.data.sum().backward()
    with torch.no_grad():
        m.eval()(x)
    torch.cuda.synchronize()
    print(kind, H, "ok", flush=True)
