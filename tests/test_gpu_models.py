"""GPU parity of the mirrored model classes (b200rnn.AudioBiLSTM / TextBiLSTM / fusion_net / MyLoss) against the
golden fixtures recorded from the REFERENCE's own classes (oracle/make_golden.py). Logits <= 1e-4 (north_star),
gradients <= 1e-4 relative to the largest entry of each tensor."""
import numpy as np
import pytest
import torch

from conftest import load_golden
from oracle import params

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _close(summary, tensor, rtol, what):
    got = params.summarize(tensor)
    assert set(got) == set(summary), what
    if "full" in summary:
        scale = max(np.abs(summary["full"]).max(), 1e-30)
        assert np.abs(got["full"] - summary["full"]).max() <= rtol * scale, what
        return
    scale = max(float(summary["absmax"][0]), 1e-30)
    assert np.abs(got["head"] - summary["head"]).max() <= rtol * scale, what
    assert np.abs(got["sample"] - summary["sample"]).max() <= rtol * scale, what
    assert abs(float(got["abssum"][0]) - float(summary["abssum"][0])) <= rtol * float(summary["abssum"][0]) + 1e-12, what


@pytest.mark.parametrize("case,cls_name,regression", [
    ("audio_clf_b3_t5", "AudioBiLSTM", False),
    ("text_clf_b3_t6", "TextBiLSTM", False),
    ("audio_reg_b2_t3", "AudioBiLSTM", True),
    ("text_reg_b2_t3", "TextBiLSTM", True),
    ("c1_text_b1_t32", "TextBiLSTM", False),       # BASELINE.json configs[0]
])
def test_single_modal_models_match_reference_goldens(case, cls_name, regression):
    import b200rnn

    arrays, meta = load_golden(case)
    model = getattr(b200rnn, cls_name)(meta["cfg"], regression=regression)
    params.fill_module(model)
    model = model.to(DEV).eval()
    x = params.inputs_for(case, meta["shape"]).to(DEV).requires_grad_(True)
    out = model(x)
    B = meta["shape"][0]
    if meta["loss"] == "ce":
        loss = torch.nn.CrossEntropyLoss()(out, params.labels_for(case, B).to(DEV))
    elif meta["loss"] == "l1":
        loss = torch.nn.L1Loss()(out, (params.inputs_for(case + ":target", (B, 1)).abs() * 10).to(DEV))
    else:
        loss = torch.nn.SmoothL1Loss()(out, (params.inputs_for(case + ":target", (B, 1)).abs() * 10).to(DEV))
    loss.backward()
    torch.cuda.synchronize()
    assert np.abs(out.detach().cpu().numpy() - arrays["out"]).max() < 1e-4
    assert abs(loss.item() - float(arrays["loss"][0])) < 1e-4
    _close(arrays["dx"], x.grad, 1e-4, "dx")
    for n, p in model.named_parameters():
        key = "grad:" + n
        if key in arrays:
            _close(arrays[key], p.grad, 1e-4, key)


@pytest.mark.parametrize("case,kind", [("gru_boundary_b5_t24", "gru"), ("lstm_boundary_b5_t17", "lstm")])
def test_rnn_boundary_matches_reference_instance(case, kind):
    """The exact drop-in boundary: what the reference model's own nn.GRU / nn.LSTM instance produced."""
    import b200rnn

    arrays, meta = load_golden(case)
    cls = b200rnn.AudioBiLSTM if kind == "gru" else b200rnn.TextBiLSTM
    model = cls(meta["cfg"])
    params.fill_module(model)
    rnn = getattr(model, meta["attr"]).to(DEV).eval()
    x = params.inputs_for(case, meta["shape"]).to(DEV).requires_grad_(True)
    res = rnn(x.permute(1, 0, 2) if meta["time_major"] else x)
    y = res[0]
    hs = res[1] if isinstance(res[1], tuple) else (res[1],)
    w = params.inputs_for(case + ":w", y.shape).to(DEV)
    loss = (y * w).sum()
    for i, h in enumerate(hs):
        loss = loss + (h * params.inputs_for(case + f":wh{i}", h.shape).to(DEV)).sum()
    loss.backward()
    torch.cuda.synchronize()
    _close(arrays["y"], y, 1e-5, "y")
    for i, h in enumerate(hs):
        _close(arrays[f"state{i}"], h, 1e-5, f"state{i}")
    _close(arrays["dx"], x.grad, 1e-4, "dx")
    for n, p in rnn.named_parameters():
        _close(arrays["grad:" + n], p.grad, 1e-4, n)


@pytest.mark.parametrize("case,regression", [("fuse_clf_b3_t3", False), ("fuse_reg_b3_t3", True)])
def test_fusion_step_matches_reference_goldens(case, regression):
    """fuse_net_whole train-step semantics from the reference's list-of-pairs input (fuse_net_whole.py:429-456)."""
    import b200rnn

    arrays, meta = load_golden(case)
    cfg = meta["cfg"]
    model = b200rnn.fusion_net(cfg["text_embed_size"], cfg["text_hidden_dims"], cfg["rnn_layers"], cfg["dropout"],
                               cfg["num_classes"], cfg["audio_hidden_dims"], cfg["audio_embed_size"],
                               regression=regression)
    params.fill_module(model)
    model = model.to(DEV).eval()
    B, T = meta["B"], meta["T"]
    audio = params.inputs_for(case + ":audio", (B, T, cfg["audio_embed_size"])).numpy()
    text = params.inputs_for(case + ":text", (B, T, cfg["text_embed_size"])).numpy()
    x = [[audio[i], text[i]] for i in range(B)]
    tf, af = model.pretrained_feature(x)
    out = model(torch.cat((tf, af), dim=1))
    loss = b200rnn.MyLoss(cfg["text_hidden_dims"], regression=regression)(tf, af, arrays["target"].tolist(), model)
    loss.backward()
    torch.cuda.synchronize()
    assert np.abs(tf.cpu().numpy() - arrays["text_feature"]).max() < 1e-4
    assert np.abs(af.cpu().numpy() - arrays["audio_feature"]).max() < 1e-4
    assert np.abs(out.detach().cpu().numpy() - arrays["out"]).max() < 1e-4
    assert abs(loss.item() - float(arrays["loss"][0])) < 1e-4
    _close(arrays["grad:fc_final.0.weight"], model.fc_final[0].weight.grad, 1e-4, "fc_final grad")
    assert [n for n, p in model.named_parameters() if p.grad is not None] == ["fc_final.0.weight"]


def test_data_parallel_shards_reproduce_full_batch_gradient():
    """Single-process DP equivalence: 8 shard gradients, mean-reduced, equal the full-batch gradient."""
    import b200rnn

    torch.manual_seed(0)
    cfg = dict(num_classes=2, dropout=0.0, rnn_layers=2, embedding_size=256, hidden_dims=256)
    model = b200rnn.AudioBiLSTM(cfg).to(DEV).eval()
    x = torch.randn(32, 12, 256, device=DEV)
    y = torch.randint(0, 2, (32,), device=DEV)
    torch.nn.functional.cross_entropy(model(x), y).backward()
    full = torch.cat([p.grad.reshape(-1) for p in model.parameters() if p.grad is not None]).clone()
    model.zero_grad()
    bucket = b200rnn.GradBucket(model)     # wgrad kernels accumulate straight into the flat bucket
    for r in range(8):
        sl = b200rnn.shard_batch(32, r, 8)
        (torch.nn.functional.cross_entropy(model(x[sl]), y[sl]) / 8).backward()
    # compare parameter by parameter (unused attention_layer params stay zero in the bucket)
    off = 0
    k = 0
    for p in model.parameters():
        n = p.numel()
        g = bucket.flat[off:off + n]
        off += n
        if g.abs().sum() == 0:
            continue
        ref = full[k:k + n]
        k += n
        assert (g - ref).abs().max().item() <= 1e-4 * ref.abs().max().item() + 1e-9
    assert k == full.numel()


@pytest.mark.parametrize("use_ln,train_mode", [(True, False), (False, False), (True, True)])
def test_fused_ln_gru_sum_matches_unfused(use_ln, train_mode):
    """b200rnn_forward_fused (LayerNorm prologue + time-sum epilogue) == ln -> GRU -> sum(dim=1)."""
    import b200rnn

    torch.manual_seed(3)
    gru = b200rnn.GRU(256, 256, num_layers=2, dropout=0.0, batch_first=True).to(DEV)
    ln = torch.nn.LayerNorm(256).to(DEV)
    with torch.no_grad():
        ln.weight.uniform_(0.5, 1.5)
        ln.bias.uniform_(-0.2, 0.2)
    gru.train(train_mode)
    x = torch.randn(9, 14, 256, device=DEV) * 3 + 1
    with torch.no_grad():
        fused = gru.forward_ln_sum(x, ln if use_ln else None)
        ref = gru(ln(x) if use_ln else x)[0].sum(dim=1)
    assert fused.shape == ref.shape == (9, 256)
    assert (fused - ref).abs().max().item() <= 2e-5   # sum of 14 outputs, each within 1e-6
    # CPU oracle of the same expression
    cpu = torch.nn.GRU(256, 256, num_layers=2, batch_first=True)
    cpu.load_state_dict(gru.state_dict())
    ln_cpu = torch.nn.LayerNorm(256)
    ln_cpu.load_state_dict(ln.state_dict())
    with torch.no_grad():
        xc = x.cpu()
        oracle = cpu(ln_cpu(xc) if use_ln else xc)[0].sum(dim=1)
    assert (fused.cpu() - oracle).abs().max().item() <= 5e-5


def test_fused_path_falls_back_under_autograd():
    import b200rnn

    gru = b200rnn.GRU(256, 256, num_layers=1, batch_first=True).to(DEV)
    ln = torch.nn.LayerNorm(256).to(DEV)
    x = torch.randn(3, 4, 256, device=DEV, requires_grad=True)
    out = gru.forward_ln_sum(x, ln)
    out.sum().backward()
    assert x.grad is not None and torch.isfinite(x.grad).all()


def test_frozen_weight_cache_is_bitwise_neutral_and_follows_weight_updates():
    """b200rnn_prepare_weights: frozen encoders split weight_ih once; the cached path must be bit-identical to the
    split-on-the-fly path, must refresh after an in-place weight edit, and must be off for trainable modules."""
    import b200rnn

    torch.manual_seed(3)
    gru = b200rnn.GRU(256, 256, num_layers=2, batch_first=True).to(DEV).eval()
    ln = torch.nn.LayerNorm(256).to(DEV)
    x = torch.randn(9, 14, 256, device=DEV)
    with torch.no_grad():
        assert gru.frozen_weight_cache() is None            # trainable: no cache
        base = gru.forward_ln_sum(x, ln).clone()
        for p in gru.parameters():
            p.requires_grad = False
        cache = gru.frozen_weight_cache()
        assert cache is not None and gru.frozen_weight_cache() is cache     # built once
        assert torch.equal(gru.forward_ln_sum(x, ln), base)
        gru.weight_ih_l1.mul_(1.5)                          # in-place edit bumps the version counter
        cache2 = gru.frozen_weight_cache()
        assert cache2 is not cache
        got = gru.forward_ln_sum(x, ln).clone()
        for p in gru.parameters():
            p.requires_grad = True
        assert gru.frozen_weight_cache() is None
        assert torch.equal(gru.forward_ln_sum(x, ln), got)  # uncached result with the edited weights
        assert not torch.equal(got, base)


@pytest.mark.parametrize("B,T,reg", [(64, 120, False), (5, 7, False), (6, 9, True)])
def test_fused_ln_gru_pool_under_autograd_matches_cpu_oracle(B, T, reg):
    """SURVEY.md 8f rank 1 for the TRAINING path (audio_gru_whole.py:103-108 + loss.backward() :190): LayerNorm folded
    into the layer-0 projection forward and backward, time pooling in the recurrence epilogue, pooled gradient
    broadcast inside the BPTT kernel - against oracle.ref_models.RefAudio (stock torch, CPU) at BASELINE c2 size."""
    import b200rnn
    from oracle import ref_models

    torch.manual_seed(1)
    cfg = dict(num_classes=1 if reg else 2, dropout=0.0, rnn_layers=2, embedding_size=256, hidden_dims=256)
    ref = ref_models.RefAudio(cfg, regression=reg).train()
    mine = b200rnn.AudioBiLSTM(cfg, regression=reg)
    mine.load_state_dict(ref.state_dict())
    mine = mine.to(DEV).train()
    x = torch.randn(B, T, 256)
    w = torch.randn(B, cfg["num_classes"])
    xr = x.clone().requires_grad_(True)
    xm = x.clone().to(DEV).requires_grad_(True)
    out_r = ref(xr)
    (out_r * w).sum().backward()
    out_m = mine(xm)
    (out_m * w.to(DEV)).sum().backward()
    torch.cuda.synchronize()
    assert (out_m.cpu() - out_r.detach()).abs().max().item() < 1e-4 * max(1.0, out_r.abs().max().item())
    gx = xr.grad
    assert (xm.grad.cpu() - gx).abs().max().item() <= 1e-4 * gx.abs().max().item()
    ref_g = dict(ref.named_parameters())
    gmax = max(p.grad.abs().max().item() for p in ref.parameters() if p.grad is not None)
    n = 0
    for name, p in mine.named_parameters():
        if ref_g[name].grad is None:
            continue
        assert p.grad is not None, name
        err = (p.grad.cpu() - ref_g[name].grad).abs().max().item()
        assert err <= 1e-4 * gmax, (name, err, gmax)
        n += 1
    assert n >= (14 if not reg else 12)
