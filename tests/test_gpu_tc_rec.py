"""The opt-in tensor-core forward recurrence (csrc/rnn_rec_tc.cu, B200RNN_REC_TC=1) holds the same parity bar as the
default FFMA kernel. The switch is read once per process, so the check runs in a child process."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu

CHILD = r"""
import sys, torch
sys.path.insert(0, sys.argv[1])
import b200rnn
dev = torch.device("cuda:0")
worst = 0.0
for kind, B, T, I, H, L, bi in [("gru", 5, 3, 64, 256, 1, False), ("gru", 37, 40, 256, 256, 2, False),
                                ("gru", 7, 9, 32, 128, 1, True), ("lstm", 33, 30, 1024, 128, 2, True)]:
    torch.manual_seed(0)
    cls = torch.nn.GRU if kind == "gru" else torch.nn.LSTM
    ref = cls(I, H, num_layers=L, bidirectional=bi, batch_first=True).eval()
    mine = b200rnn.from_torch(ref).to(dev).eval()
    x = torch.randn(B, T, I)
    xr = x.clone().requires_grad_(True); xm = x.clone().to(dev).requires_grad_(True)
    mine.train(); ref.train()            # dropout = 0: train mode only switches on the saved-for-backward stores
    yr, _ = ref(xr); ym, _ = mine(xm)
    worst = max(worst, (ym.detach().cpu() - yr.detach()).abs().max().item())
    yr.sum().backward(); ym.sum().backward()
    g = (xm.grad.cpu() - xr.grad).abs().max().item() / xr.grad.abs().max().item()
    assert g < 1e-4, (kind, "dx", g)
print("WORST", worst)
assert worst < 1e-5, worst
"""


def test_tensor_core_recurrence_matches_torch_cpu_in_a_child_process():
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    env = dict(os.environ, B200RNN_REC_TC="1", B200RNN_DEBUG="1")
    out = subprocess.run([sys.executable, "-c", CHILD, os.path.join(root, "icassp2022-depression_b200")], env=env,
                         capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-4000:]
    assert "tc fwd mode=" in out.stderr, "the tensor-core kernel did not take the launch"
    assert "WORST" in out.stdout
