"""GPU parity for ragged batches (torch PackedSequence semantics) — SURVEY.md §8f "variable-length sequences".

The DAIC feature extraction (DAICFeatureExtarction/feature_extraction.py:45-64) yields a different number of
responses per participant; the reference pads them, stock torch users pack them. Oracle: stock torch.nn.GRU / LSTM on
CPU fed the same PackedSequence. The kernels get the padded [T,B,*] block plus a device int32 lengths array: past its
length a sequence keeps its state and emits zeros. Tolerances as tests/test_gpu_parity.py.
"""
import pytest
import torch
from torch.nn.utils.rnn import pack_padded_sequence, pad_packed_sequence, pack_sequence

pytestmark = pytest.mark.gpu

OUT_TOL = 1e-5
GRAD_RTOL = 1e-4


def _dev():
    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    return torch.device("cuda:0")


def _pair(kind, I, H, L, bi, bf, seed=0):
    import b200rnn

    torch.manual_seed(seed)
    cls = torch.nn.GRU if kind == "gru" else torch.nn.LSTM
    ref = cls(I, H, num_layers=L, bidirectional=bi, batch_first=bf).eval()
    return ref, b200rnn.from_torch(ref).to(_dev()).eval()


def _states(out):
    return out[1] if isinstance(out[1], tuple) else (out[1],)


CASES = [
    # kind, B, T, I, H, L, bi, batch_first
    ("gru", 5, 9, 24, 128, 1, False, True),
    ("gru", 7, 17, 256, 256, 2, False, True),     # audio_gru_whole stack, ragged
    ("gru", 6, 11, 40, 256, 2, True, False),
    ("lstm", 5, 9, 24, 128, 1, True, True),
    ("lstm", 9, 13, 1024, 128, 2, True, False),   # text_bilstm_whole stack, ragged
    ("lstm", 4, 10, 64, 256, 2, False, True),
]


@pytest.mark.parametrize("kind,B,T,I,H,L,bi,bf", CASES)
def test_packed_sequence_forward_backward_matches_torch_cpu(kind, B, T, I, H, L, bi, bf):
    ref, mine = _pair(kind, I, H, L, bi, bf)
    g = torch.Generator().manual_seed(3)
    lengths = torch.randint(1, T + 1, (B,), generator=g)
    lengths[int(torch.randint(0, B, (1,), generator=g))] = T        # at least one full-length row
    x = torch.randn((B, T, I) if bf else (T, B, I), generator=g)
    xr = x.clone().requires_grad_(True)
    xm = x.clone().to(_dev()).requires_grad_(True)
    pr = pack_padded_sequence(xr, lengths, batch_first=bf, enforce_sorted=False)
    pm = pack_padded_sequence(xm, lengths, batch_first=bf, enforce_sorted=False)
    out_r, out_m = ref(pr), mine(pm)
    assert isinstance(out_m[0], torch.nn.utils.rnn.PackedSequence)
    assert torch.equal(out_m[0].batch_sizes, out_r[0].batch_sizes)
    assert torch.equal(out_m[0].sorted_indices.cpu(), out_r[0].sorted_indices)
    assert torch.equal(out_m[0].unsorted_indices.cpu(), out_r[0].unsorted_indices)
    assert (out_m[0].data.cpu() - out_r[0].data).abs().max().item() <= OUT_TOL
    for a, b in zip(_states(out_m), _states(out_r)):
        assert a.shape == b.shape
        assert (a.cpu() - b).abs().max().item() <= OUT_TOL

    # backward through outputs and final states
    w = torch.randn(out_r[0].data.shape, generator=g)
    loss_r = (out_r[0].data * w).sum()
    loss_m = (out_m[0].data * w.to(_dev())).sum()
    for a, b in zip(_states(out_m), _states(out_r)):
        ws = torch.randn(b.shape, generator=g)
        loss_r = loss_r + (b * ws).sum()
        loss_m = loss_m + (a * ws.to(_dev())).sum()
    loss_r.backward()
    loss_m.backward()
    scale = xr.grad.abs().max().clamp_min(1e-30)
    assert ((xm.grad.cpu() - xr.grad).abs().max() / scale).item() <= GRAD_RTOL
    for (n, p_r), (_, p_m) in zip(ref.named_parameters(), mine.named_parameters()):
        scale = p_r.grad.abs().max().clamp_min(1e-30)
        assert ((p_m.grad.cpu() - p_r.grad).abs().max() / scale).item() <= GRAD_RTOL, n


def test_padded_positions_are_zero_and_untouched_by_padding_values():
    """Garbage in the padding must not reach any output, state or gradient."""
    ref, mine = _pair("gru", 32, 128, 2, True, True)
    g = torch.Generator().manual_seed(5)
    B, T = 6, 12
    lengths = torch.tensor([12, 3, 7, 1, 12, 5])
    x = torch.randn(B, T, 32, generator=g)
    x2 = x.clone()
    for b in range(B):
        x2[b, lengths[b]:] = 1e3 * torch.randn(T - int(lengths[b]), 32, generator=g)
    outs = []
    for xx in (x, x2):
        xm = xx.to(_dev()).requires_grad_(True)
        pm = pack_padded_sequence(xm, lengths, batch_first=True, enforce_sorted=False)
        y, h = mine(pm)
        yp, _ = pad_packed_sequence(y, batch_first=True, total_length=T)
        (yp.sum() + h.sum()).backward()
        outs.append((yp.detach().cpu(), h.detach().cpu(), xm.grad.cpu()))
    for a, b in zip(outs[0], outs[1]):
        assert torch.equal(a, b)
    yp, _, dx = outs[0]
    for b in range(B):
        assert yp[b, lengths[b]:].abs().max().item() == 0 if lengths[b] < T else True
        assert dx[b, lengths[b]:].abs().max().item() == 0 if lengths[b] < T else True


def test_pack_sequence_sorted_input_and_lengths_equal_to_T_match_dense_path():
    ref, mine = _pair("lstm", 48, 128, 2, True, False)
    g = torch.Generator().manual_seed(9)
    seqs = [torch.randn(n, 48, generator=g) for n in (8, 8, 8)]
    with torch.no_grad():
        dense = mine(torch.stack(seqs, dim=1).to(_dev()))
        packed = mine(pack_sequence([s.to(_dev()) for s in seqs]))   # enforce_sorted=True: no index tensors
    assert packed[0].sorted_indices is None
    yp, _ = pad_packed_sequence(packed[0])
    assert (yp - dense[0]).abs().max().item() <= 1e-6
    assert (packed[1][0] - dense[1][0]).abs().max().item() <= 1e-6
    assert (packed[1][1] - dense[1][1]).abs().max().item() <= 1e-6
