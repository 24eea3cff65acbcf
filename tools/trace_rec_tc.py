#!/usr/bin/env python
"""Debug: per-step phase timeline (SM clock cycles) of the tensor-core GRU recurrence, CTA 0 / thread 0.
Run with B200RNN_REC_TC=1 [B200RNN_TC_DBG=n]."""
import ctypes, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "icassp2022-depression_b200"))
import torch, b200rnn
from b200rnn import _lib
lib = _lib.load()
lib.b200rnn_debug_set_trace.argtypes = [ctypes.c_void_p]
B = int(sys.argv[1]) if len(sys.argv) > 1 else 128
dev = torch.device("cuda:0")
m = b200rnn.GRU(256, 256, num_layers=1, batch_first=True).to(dev).eval(); x = torch.randn(B, 120, 256, device=dev); T = 120
with torch.no_grad():
    m(x)
    buf = torch.zeros(T, 8, dtype=torch.int64, device=dev)
    lib.b200rnn_debug_set_trace(buf.data_ptr())
    m(x)
    torch.cuda.synchronize()
    lib.b200rnn_debug_set_trace(None)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        m(x)
    e1.record(); torch.cuda.synchronize()
t = buf.cpu()
names = ["recv wait", "split+sync", "mma issue", "mma done wait", "ldtm+pre+sync", "cell+send"]
print(f"B={B} dbg={os.environ.get('B200RNN_TC_DBG', '0')}: layer fwd (gemm+rec) {e0.elapsed_time(e1) / 10 * 1e3:.1f} us")
for s in range(8, 12):
    r = t[s]; nxt = t[s + 1][0]
    d = [int(r[i + 1] - r[i]) for i in range(6)]
    print(f"step {s}: total {int(nxt - r[0])}  " + "  ".join(f"{n}={v}" for n, v in zip(names, d)))
