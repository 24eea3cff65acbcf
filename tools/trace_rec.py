#!/usr/bin/env python
"""Debug: per-step phase timeline (SM clock cycles) of the forward GRU recurrence, CTA 0 / warp 0."""
import ctypes, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "icassp2022-depression_b200"))
import torch, b200rnn
from b200rnn import _lib
lib = _lib.load()
lib.b200rnn_debug_set_trace.argtypes = [ctypes.c_void_p]
kind = sys.argv[1] if len(sys.argv) > 1 else "gru"
dev = torch.device("cuda:0")
if kind == "gru":
    m = b200rnn.GRU(256, 256, num_layers=1, batch_first=True).to(dev).eval(); x = torch.randn(128, 120, 256, device=dev); T = 120
else:
    m = b200rnn.LSTM(1024, 128, num_layers=1, bidirectional=True).to(dev).eval(); x = torch.randn(30, 128, 1024, device=dev); T = 30
with torch.no_grad():
    m(x)
    buf = torch.zeros(T, 8, dtype=torch.int64, device=dev)
    lib.b200rnn_debug_set_trace(buf.data_ptr())
    m(x)
    torch.cuda.synchronize()
    lib.b200rnn_debug_set_trace(None)
t = buf.cpu()
names = ["top->lastwait", "lastchunk", "fold+reduce", "gates", "allgather", "tail->next top"]
for s in range(8, 16):
    r = t[s]; nxt = t[s + 1][0]
    d = [int(r[1] - r[0]), int(r[2] - r[1]), int(r[3] - r[2]), int(r[4] - r[3]), int(r[5] - r[4]), int(nxt - r[5])]
    print(f"step {s}: total {int(nxt - r[0])}  " + "  ".join(f"{n}={v}" for n, v in zip(names, d)))
