"""Host-side logic of b200rnn.FlatAdamW that needs no GPU: the reference's parameter grouping
(audio_gru_whole.py:247-255 `get_param_group`: names containing 'ln' get weight_decay 0, everything else 1e-5),
re-homing of parameters / gradients into flat buffers, and autograd accumulating into the flat gradient."""
import pytest
import torch

import b200rnn


class _Tiny(torch.nn.Module):
    def __init__(self):
        super().__init__()
        self.ln = torch.nn.LayerNorm(6)
        self.fc = torch.nn.Linear(6, 4)
        self.fc_out = torch.nn.Linear(4, 2)

    def forward(self, x):
        return self.fc_out(torch.relu(self.fc(self.ln(x))))


def _reference_groups(model):   # restated from audio_gru_whole.py:247-255
    nd, rest = [], []
    for name, p in model.named_parameters():
        (nd if "ln" in name else rest).append(p)
    return [{"params": rest, "weight_decay": 1e-5}, {"params": nd, "weight_decay": 0}]


def test_like_reference_groups_like_get_param_group():
    torch.manual_seed(0)
    m = _Tiny()
    want = _reference_groups(m)
    opt = b200rnn.FlatAdamW.like_reference(m, lr=1e-3)
    assert [g.weight_decay for g in opt.groups] == [1e-5, 0.0]
    for g, w in zip(opt.groups, want):
        assert [id(p) for p in g.params] == [id(p) for p in w["params"]]


def test_parameters_and_gradients_live_in_the_flat_buffers():
    torch.manual_seed(0)
    m = _Tiny()
    before = {n: p.detach().clone() for n, p in m.named_parameters()}
    opt = b200rnn.FlatAdamW.like_reference(m, lr=1e-3)
    for n, p in m.named_parameters():
        assert torch.equal(p.detach(), before[n]), n            # values survive the move
    for g in opt.groups:
        lo, hi = g.flat_p.data_ptr(), g.flat_p.data_ptr() + 4 * g.flat_p.numel()
        off = 0
        for p in g.params:
            off = (off + 63) // 64 * 64     # packed in order, every view on a 256-byte boundary of the flat buffer
            assert lo <= p.data_ptr() < hi and p.data_ptr() == lo + 4 * off
            assert p.grad.data_ptr() == g.flat_g.data_ptr() + 4 * off
            off += p.numel()
        assert off == g.flat_p.numel()
    x = torch.randn(5, 6)
    m(x).square().sum().backward()
    ref = _Tiny()
    ref.load_state_dict({n: v for n, v in before.items()})
    ref(x).square().sum().backward()
    for g in opt.groups:
        assert g.flat_g.abs().sum() > 0
    for (n, p), (_, q) in zip(m.named_parameters(), ref.named_parameters()):
        assert torch.allclose(p.grad, q.grad, atol=1e-6), n     # autograd accumulated into the flat views
    opt.zero_grad()
    assert all(float(g.flat_g.abs().sum()) == 0.0 for g in opt.groups)
    assert all(float(p.grad.abs().sum()) == 0.0 for p in m.parameters())


def test_frozen_parameters_are_left_out_and_step_has_no_cpu_path():
    m = _Tiny()
    m.fc.weight.requires_grad = False
    opt = b200rnn.FlatAdamW.like_reference(m, lr=1e-3)
    assert all(id(m.fc.weight) != id(p) for g in opt.groups for p in g.params)
    assert (sum(p.numel() for g in opt.groups for p in g.params) ==
            sum(p.numel() for p in m.parameters() if p.requires_grad))
    with pytest.raises(b200rnn.B200RNNError, match="no CPU path"):
        opt.step()      # the update is a CUDA kernel (b200rnn_adamw); there is no CPU fallback
