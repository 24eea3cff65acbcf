"""GPU parity tests: the CUDA path (through the C-ABI) against the oracle on identical seeded inputs.

Oracle = stock torch.nn.GRU / nn.LSTM on CPU (the dependency that holds the reference's arithmetic) and the
float64 numpy restatement (oracle/rnn_numpy.py). Tolerances (SURVEY.md §8c error budget; fp32 vs fp64 oracle is
~3e-7 on outputs, ~2e-6 on grads): outputs <= 1e-5 abs, logits <= 1e-4 abs (north_star), grads <= 1e-4 relative
to the largest gradient entry. Everything runs in eval() / dropout=0 except the dropout tests.
"""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

OUT_TOL = 1e-5
GRAD_RTOL = 1e-4


def _dev():
    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    return torch.device("cuda:0")


def _mk(kind, I, H, L, bi, batch_first, seed=0):
    import b200rnn

    torch.manual_seed(seed)
    cls = torch.nn.GRU if kind == "gru" else torch.nn.LSTM
    ref = cls(I, H, num_layers=L, bidirectional=bi, batch_first=batch_first)
    ref.eval()
    mine = b200rnn.from_torch(ref).to(_dev()).eval()
    return ref, mine


def _relmax(a, b):
    return ((a - b).abs().max() / b.abs().max().clamp_min(1e-30)).item()


def _run_pair(ref, mine, x_cpu, permuted_view=False, use_states=True, seed=1):
    """Forward + backward on both; returns dict of errors."""
    g = torch.Generator().manual_seed(seed)
    xr = x_cpu.clone().requires_grad_(True)
    xm_base = x_cpu.clone().to(_dev()).requires_grad_(True)
    xin_r, xin_m = xr, xm_base
    if permuted_view:  # the reference feeds a permuted NON-contiguous view (text_bilstm_whole.py:103)
        xin_r, xin_m = xr.permute(1, 0, 2), xm_base.permute(1, 0, 2)
    out_r = ref(xin_r)
    out_m = mine(xin_m)
    yr, ym = out_r[0], out_m[0]
    sr = out_r[1] if isinstance(out_r[1], tuple) else (out_r[1],)
    sm = out_m[1] if isinstance(out_m[1], tuple) else (out_m[1],)
    errs = {"y": (ym.cpu() - yr).abs().max().item()}
    for i, (a, b) in enumerate(zip(sm, sr)):
        errs[f"state{i}"] = (a.cpu() - b).abs().max().item()
    w = torch.randn(yr.shape, generator=g)
    loss_r = (yr * w).sum()
    loss_m = (ym * w.to(_dev())).sum()
    if use_states:
        for a, b in zip(sm, sr):
            ws = torch.randn(b.shape, generator=g)
            loss_r = loss_r + (b * ws).sum()
            loss_m = loss_m + (a * ws.to(_dev())).sum()
    loss_r.backward()
    loss_m.backward()
    torch.cuda.synchronize()
    errs["dx"] = _relmax(xm_base.grad.cpu(), xr.grad)
    for (n, pr), (_, pm) in zip(ref.named_parameters(), mine.named_parameters()):
        errs["d" + n] = _relmax(pm.grad.cpu(), pr.grad)
    return errs


def _assert_errs(errs, where=""):
    for k, v in errs.items():
        tol = OUT_TOL if (k == "y" or k.startswith("state")) else GRAD_RTOL
        assert v <= tol, f"{where}: {k} error {v:.3e} > {tol:.0e}  (all: {errs})"


CASES = [
    # kind, B, T, I, H, L, bi, batch_first
    ("gru", 4, 5, 256, 256, 1, False, True),
    ("gru", 6, 9, 256, 256, 2, False, True),      # audio_gru_whole shape family
    ("gru", 1, 3, 256, 256, 2, False, True),      # B=1, true EATD T=3
    ("gru", 5, 7, 96, 128, 2, True, False),       # bidirectional GRU, H=128, I not a tile multiple
    ("gru", 13, 4, 37, 128, 1, False, False),     # I not a multiple of 4 (scalar-load path), ragged batch
    ("lstm", 4, 5, 1024, 128, 1, True, False),
    ("lstm", 7, 6, 1024, 128, 2, True, False),    # text_bilstm_whole shape family (H=128)
    ("lstm", 3, 4, 1024, 256, 2, True, False),    # BASELINE config c3 family (H=256)
    ("lstm", 1, 32, 1024, 128, 2, True, False),   # BASELINE configs[0] shape
    ("lstm", 9, 3, 64, 128, 2, False, True),      # unidirectional, batch_first LSTM
    ("lstm", 2, 1, 40, 256, 1, True, False),      # T=1
]


@pytest.mark.parametrize("kind,B,T,I,H,L,bi,bf", CASES)
def test_rnn_forward_backward_matches_torch_cpu(kind, B, T, I, H, L, bi, bf):
    ref, mine = _mk(kind, I, H, L, bi, bf)
    torch.manual_seed(123)
    x = torch.randn((B, T, I) if bf else (T, B, I))
    errs = _run_pair(ref, mine, x)
    _assert_errs(errs, f"{kind} B{B} T{T} I{I} H{H} L{L} bi{bi} bf{bf}")


def test_permuted_noncontiguous_input_is_consumed_in_place():
    ref, mine = _mk("lstm", 1024, 128, 2, True, False)
    torch.manual_seed(5)
    x = torch.randn(6, 8, 1024)  # [B,T,E]; model permutes to [T,B,E] view
    errs = _run_pair(ref, mine, x, permuted_view=True)
    _assert_errs(errs, "permuted")


@pytest.mark.parametrize("kind,B,T,H,L,bi", [
    ("gru", 64, 120, 256, 2, False),    # BASELINE c2 (forward/backward vs torch CPU at full size)
    ("lstm", 64, 30, 256, 2, True),     # BASELINE c3
    ("gru", 130, 33, 256, 2, False),    # > one wave of batch slices, ragged tail
    ("lstm", 130, 12, 128, 2, True),
])
def test_full_size_against_torch_cpu(kind, B, T, H, L, bi):
    I = 256 if kind == "gru" else 1024
    ref, mine = _mk(kind, I, H, L, bi, kind == "gru")
    torch.manual_seed(7)
    x = torch.randn((B, T, I) if kind == "gru" else (T, B, I))
    errs = _run_pair(ref, mine, x)
    _assert_errs(errs, f"full {kind} B{B} T{T} H{H}")


@pytest.mark.parametrize("kind", ["gru", "lstm"])
def test_against_float64_numpy_oracle(kind):
    from oracle.rnn_numpy import NumpyRNN

    bi = kind == "lstm"
    I, H, L, B, T = 48, 128, 2, 5, 11
    ref, mine = _mk(kind, I, H, L, bi, False, seed=3)
    torch.manual_seed(11)
    x = torch.randn(T, B, I)
    w64 = [p.detach().double().numpy() for p in ref.parameters()]
    orc = NumpyRNN(kind, w64, L, bi)
    out = orc.forward(x.double().numpy())
    xm = x.to(_dev()).requires_grad_(True)
    res = mine(xm)
    ym = res[0]
    assert np.abs(ym.detach().cpu().double().numpy() - out[0]).max() < OUT_TOL
    dy = torch.randn(ym.shape, generator=torch.Generator().manual_seed(2))
    (ym * dy.to(_dev())).sum().backward()
    dx64, dparams64 = orc.backward(dy.double().numpy())
    assert np.abs(xm.grad.cpu().double().numpy() - dx64).max() / np.abs(dx64).max() < GRAD_RTOL
    for p, g64 in zip(mine.parameters(), dparams64):
        assert np.abs(p.grad.cpu().double().numpy() - g64).max() / max(np.abs(g64).max(), 1e-30) < GRAD_RTOL


def test_batch_rows_are_independent_and_deterministic():
    """Size-independent property at BASELINE size: any sub-batch reproduces its rows bit-for-bit-close, and a
    rerun is bit-identical (fixed reduction orders, no atomics)."""
    ref, mine = _mk("gru", 256, 256, 2, False, True)
    torch.manual_seed(9)
    x = torch.randn(128, 120, 256, device=_dev())
    with torch.no_grad():
        y_full, h_full = mine(x)
        y_again, _ = mine(x)
        y_half, h_half = mine(x[32:96])
    assert torch.equal(y_full, y_again)
    assert (y_full[32:96] - y_half).abs().max().item() <= 2e-6
    assert (h_full[:, 32:96] - h_half).abs().max().item() <= 2e-6


def test_reverse_direction_is_forward_on_flipped_time():
    """Property: the reverse half of a bidirectional layer equals a forward scan of the time-flipped input."""
    import b200rnn

    torch.manual_seed(4)
    bi = torch.nn.LSTM(64, 128, num_layers=1, bidirectional=True)
    uni = torch.nn.LSTM(64, 128, num_layers=1)
    with torch.no_grad():
        for n in ("weight_ih_l0", "weight_hh_l0", "bias_ih_l0", "bias_hh_l0"):
            getattr(uni, n).copy_(getattr(bi, n + "_reverse"))
    mb = b200rnn.from_torch(bi).to(_dev()).eval()
    mu = b200rnn.from_torch(uni).to(_dev()).eval()
    x = torch.randn(30, 128, 64, device=_dev())
    with torch.no_grad():
        yb, _ = mb(x)
        yu, _ = mu(x.flip(0))
    assert (yb[:, :, 128:] - yu.flip(0)).abs().max().item() <= 2e-6


def test_dropout_train_mode_statistics_and_backward_mask():
    """Inter-layer dropout (rnn.py:1233-1236): keep-rate ~ 1-p, kept values scaled 1/(1-p), fresh mask per call,
    and backward applies the same mask as forward."""
    import b200rnn

    p = 0.5
    torch.manual_seed(0)
    m = b200rnn.GRU(256, 256, num_layers=2, dropout=p, batch_first=True).to(_dev()).train()
    # make layer 1 the identity on its input sum: zero recurrent part is not possible, so test via gradients:
    x = torch.randn(16, 10, 256, device=_dev(), requires_grad=True)
    y1, _ = m(x)
    y2, _ = m(x)
    assert not torch.equal(y1, y2), "train-mode dropout must draw a fresh mask on every call"
    m.eval()
    y3, _ = m(x)
    y4, _ = m(x)
    assert torch.equal(y3, y4)
    # mask statistics through the public intermediate: compare layer-0 output (eval, 1 layer) with what layer 1 sees
    m1 = b200rnn.GRU(256, 256, num_layers=1, batch_first=True).to(_dev())
    with torch.no_grad():
        for n in ("weight_ih_l0", "weight_hh_l0", "bias_ih_l0", "bias_hh_l0"):
            getattr(m1, n).copy_(getattr(m, n))
    # linear probe: d(loss)/d(layer-0 output) is masked; use finite T=1 so dropout input == h_1 of layer 0
    m.train()
    xs = torch.randn(64, 1, 256, device=_dev(), requires_grad=True)
    ys, _ = m(xs)
    ys.sum().backward()
    assert torch.isfinite(xs.grad).all()
    # gradient w.r.t. the layer-1 input weights only sees kept units: columns of dW_ih_l1 for dropped units are not
    # all zero across the batch, but the expected keep-rate shows up in the zero pattern of a single sample
    xs1 = torch.randn(1, 1, 256, device=_dev())
    m.zero_grad()
    y1s, _ = m(xs1)
    y1s.sum().backward()
    col_zero = (m.weight_ih_l1.grad.abs().sum(0) == 0).float().mean().item()
    assert 0.3 < col_zero < 0.7, f"fraction of dropped layer-0 units {col_zero} not ~ p"
    with torch.no_grad():
        h0, _ = m1(xs1)
    kept = m.weight_ih_l1.grad.abs().sum(0) != 0
    # for kept units, dW_ih_l1[:, j] = dgi * h0_j/(1-p)  =>  ratio of two columns equals ratio of scaled inputs
    g = m.weight_ih_l1.grad
    jk = torch.nonzero(kept).flatten()[:2]
    ratio_g = (g[:, jk[0]] / g[:, jk[1]]).median().item()
    ratio_h = (h0[0, 0, jk[0]] / h0[0, 0, jk[1]]).item()
    assert abs(ratio_g - ratio_h) <= 1e-3 * abs(ratio_h) + 1e-5


def test_no_grad_train_mode_forward_matches_eval_when_p_is_zero():
    """fuse_net_whole.py:337 runs the encoders under no_grad in train() mode; with p=0 that equals eval()."""
    import b200rnn

    torch.manual_seed(1)
    m = b200rnn.LSTM(1024, 128, num_layers=2, dropout=0.0, bidirectional=True).to(_dev())
    x = torch.randn(30, 16, 1024, device=_dev())
    m.train()
    with torch.no_grad():
        a, _ = m(x)
    m.eval()
    with torch.no_grad():
        b, _ = m(x)
    assert torch.equal(a, b)


def test_gradient_accumulates_over_two_backward_calls():
    ref, mine = _mk("gru", 256, 256, 2, False, True)
    torch.manual_seed(2)
    x = torch.randn(4, 6, 256)
    for _ in range(2):
        ref(x)[0].sum().backward()
        mine(x.to(_dev()))[0].sum().backward()
    for (n, pr), (_, pm) in zip(ref.named_parameters(), mine.named_parameters()):
        assert _relmax(pm.grad.cpu(), pr.grad) <= GRAD_RTOL, n
