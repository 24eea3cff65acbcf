"""Bring-up check of the tensor-core recurrence (B200RNN_REC_TC=1) against stock torch CPU.

Prints, per case, the max abs error of y / h_n and where the worst entries sit (time step, batch row, unit), which is
what tells a swizzle / descriptor / exchange bug apart from rounding. Run on a GPU box:
    B200RNN_REC_TC=1 python tools/tc_rec_check.py
"""
import os
import sys
import time

import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "icassp2022-depression_b200"))
import b200rnn  # noqa: E402

dev = torch.device("cuda:0")
CASES = [
    # kind, B, T, I, H, L, bi
    ("gru", 5, 1, 64, 256, 1, False),
    ("gru", 5, 2, 64, 256, 1, False),
    ("gru", 5, 3, 64, 256, 1, False),
    ("gru", 16, 8, 64, 256, 1, False),
    ("gru", 128, 120, 256, 256, 2, False),
    ("gru", 7, 9, 32, 128, 1, True),
    ("lstm", 5, 3, 64, 128, 1, False),
    ("lstm", 128, 30, 1024, 128, 2, True),
    ("lstm", 64, 30, 1024, 256, 2, True),
]
only = sys.argv[1:]
for kind, B, T, I, H, L, bi in CASES:
    if only and kind not in only:
        continue
    torch.manual_seed(0)
    cls = torch.nn.GRU if kind == "gru" else torch.nn.LSTM
    ref = cls(I, H, num_layers=L, bidirectional=bi, batch_first=True).eval()
    mine = b200rnn.from_torch(ref).to(dev).eval()
    x = torch.randn(B, T, I)
    with torch.no_grad():
        yr, sr = ref(x)
        ym, sm = mine(x.to(dev))
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(5):
            mine(x.to(dev))
        torch.cuda.synchronize()
        dt = (time.perf_counter() - t0) / 5
    hr = sr[0] if isinstance(sr, tuple) else sr
    hm = sm[0] if isinstance(sm, tuple) else sm
    err = (ym.cpu() - yr).abs()
    print(f"{kind} B={B} T={T} I={I} H={H} L={L} bi={bi}: y err {err.max().item():.3e}  h_n err "
          f"{(hm.cpu() - hr).abs().max().item():.3e}  ({dt * 1e3:.3f} ms/fwd eager)", flush=True)
    if err.max().item() > 1e-5:
        bad = (err > 1e-5)
        print("   bad fraction", bad.float().mean().item())
        print("   per-time max ", [f"{v:.1e}" for v in err.amax(dim=(0, 2))[:8].tolist()])
        print("   per-batch max", [f"{v:.1e}" for v in err.amax(dim=(1, 2))[:16].tolist()])
        pu = err.amax(dim=(0, 1))
        print("   per-unit max (first 64)", [f"{v:.0e}" for v in pu[:64].tolist()])
        print("   units bad:", int((pu > 1e-5).sum()), "of", pu.numel())
