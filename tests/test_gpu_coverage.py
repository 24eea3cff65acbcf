"""Every recurrence instantiation in the library is launched by some test: the wide-batch fallback configs
(C = 8 / BS = 8 clusters, taken when ceil(B / 4) clusters of the tuned config do not fit the chip) and their
per-sequence-length twins are reached here with batches beyond one wave - forward + backward against stock torch CPU.
Also the LayerNorm-prologue widths other than the reference's 256."""
import pytest
import torch

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


@pytest.mark.parametrize("kind,H,B", [("gru", 256, 160), ("gru", 128, 300), ("lstm", 256, 152), ("lstm", 128, 300)])
@pytest.mark.parametrize("ragged", [False, True])
def test_wide_batch_fallback_configs_match_torch_cpu(kind, H, B, ragged):
    import b200rnn

    torch.manual_seed(5)
    T, I = 4, 32
    cls = torch.nn.GRU if kind == "gru" else torch.nn.LSTM
    ref = cls(I, H, num_layers=1, batch_first=True)
    mine = b200rnn.from_torch(ref).to(DEV)
    x = torch.randn(B, T, I)
    xr, xm = x.clone().requires_grad_(True), x.clone().to(DEV).requires_grad_(True)
    if ragged:
        lens = torch.randint(1, T + 1, (B,))
        lens[0] = T
        pk = lambda t: torch.nn.utils.rnn.pack_padded_sequence(t, lens, batch_first=True, enforce_sorted=False)  # noqa: E731
        yr = torch.nn.utils.rnn.pad_packed_sequence(ref(pk(xr))[0], batch_first=True, total_length=T)[0]
        ym = torch.nn.utils.rnn.pad_packed_sequence(mine(pk(xm))[0], batch_first=True, total_length=T)[0]
    else:
        yr, ym = ref(xr)[0], mine(xm)[0]
    w = torch.randn_like(yr)
    (yr * w).sum().backward()
    (ym * w.to(DEV)).sum().backward()
    torch.cuda.synchronize()
    assert (ym.detach().cpu() - yr.detach()).abs().max().item() < 1e-5
    assert (xm.grad.cpu() - xr.grad).abs().max().item() <= 1e-4 * xr.grad.abs().max().item()
    gmax = max(p.grad.abs().max().item() for p in ref.parameters())
    for (n, p), (_, q) in zip(mine.named_parameters(), ref.named_parameters()):
        assert (p.grad.cpu() - q.grad).abs().max().item() <= 1e-4 * gmax, n


@pytest.mark.parametrize("I", [128, 512, 1024])
def test_layernorm_prologue_widths(I):
    """LayerNorm folded around the layer-0 GEMMs (forward and backward) for the other instantiated feature widths."""
    import b200rnn

    torch.manual_seed(6)
    B, T, H = 5, 6, 128
    gru_r, ln_r = torch.nn.GRU(I, H, num_layers=1, batch_first=True), torch.nn.LayerNorm(I)
    with torch.no_grad():
        ln_r.weight.uniform_(0.5, 1.5)
        ln_r.bias.uniform_(-0.5, 0.5)
    gru_m = b200rnn.from_torch(gru_r).to(DEV)
    ln_m = torch.nn.LayerNorm(I).to(DEV)
    ln_m.load_state_dict(ln_r.state_dict())
    x = torch.randn(B, T, I)
    xr, xm = x.clone().requires_grad_(True), x.clone().to(DEV).requires_grad_(True)
    pr = gru_r(ln_r(xr))[0].sum(dim=1)
    pm = gru_m.forward_ln_sum(xm, ln_m)
    w = torch.randn_like(pr)
    (pr * w).sum().backward()
    (pm * w.to(DEV)).sum().backward()
    torch.cuda.synchronize()
    assert (pm.detach().cpu() - pr.detach()).abs().max().item() < 1e-4
    assert (xm.grad.cpu() - xr.grad).abs().max().item() <= 1e-4 * xr.grad.abs().max().item()
    assert (ln_m.weight.grad.cpu() - ln_r.weight.grad).abs().max().item() <= 1e-4 * ln_r.weight.grad.abs().max().item()
    assert (ln_m.bias.grad.cpu() - ln_r.bias.grad).abs().max().item() <= 1e-4 * ln_r.bias.grad.abs().max().item()
    with torch.no_grad():   # the no-grad fused path (pooled only, nothing saved)
        assert (gru_m.forward_ln_sum(xm.detach(), ln_m).cpu() - pr.detach()).abs().max().item() < 1e-4


def test_recurrence_is_bitwise_deterministic_over_repeated_runs():
    """The own-slice delivery of the persistent kernels is ordered by mbarrier arrive / try_wait only (no block
    barrier): a missing ordering would show as run-to-run differences. 40 repetitions of forward + backward, GRU and
    BiLSTM, must be bit-identical (the split-K and bias reductions are fixed-order as well)."""
    import b200rnn

    torch.manual_seed(9)
    for kind, B, T, I, H, bi in (("gru", 128, 60, 256, 256, False), ("lstm", 64, 30, 256, 128, True)):
        cls = b200rnn.GRU if kind == "gru" else b200rnn.LSTM
        m = cls(I, H, num_layers=2, bidirectional=bi, batch_first=True).to(DEV)
        x = torch.randn(B, T, I, device=DEV, requires_grad=True)
        ref = None
        for _ in range(40):
            m.zero_grad()
            x.grad = None
            y = m(x)[0]
            y.square().sum().backward()
            got = [y.detach().clone(), x.grad.clone()] + [p.grad.clone() for p in m.parameters()]
            if ref is None:
                ref = got
            else:
                for a, b in zip(got, ref):
                    assert torch.equal(a, b), kind
