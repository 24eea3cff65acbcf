"""GPU parity of the fused model-shell kernels against the PyTorch expressions they
replace (which tests/test_gpu_models.py in turn pins to the reference's goldens)."""
import pytest
import torch

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _fuse_model(train=False):
    import b200rnn

    torch.manual_seed(0)
    m = b200rnn.fusion_net(1024, 128, 2, 0.3, 2, 256, 256).to(DEV)
    for p in m.parameters():
        p.requires_grad = False
    m.fc_final[0].weight.requires_grad = True
    m.train(train)
    return m


def test_attention_pool_kernel_matches_torch():
    import b200rnn
    from b200rnn import fused_head

    torch.manual_seed(1)
    layer = torch.nn.Sequential(torch.nn.Linear(128, 128), torch.nn.ReLU(inplace=True)).to(DEV)
    seq = torch.randn(30, 17, 256, device=DEV)          # [T,B,2H] time-major
    h_n = torch.randn(4, 17, 128, device=DEV)
    got = fused_head.attention_pool(seq, h_n, layer)
    with torch.no_grad():
        ref = b200rnn.attention_pool(layer, seq.permute(1, 0, 2), h_n.permute(1, 0, 2))
    assert (got - ref).abs().max().item() < 1e-5   # values O(1), softmax over N(0,1)*sqrt(H) scores


def test_mlp_dropout_kernel_eval_and_train():
    from b200rnn import fused_head, _lib

    torch.manual_seed(2)
    lin = torch.nn.Linear(256, 256).to(DEV)
    x = torch.randn(64, 256, device=DEV)
    got = fused_head.mlp_dropout(x, lin, 0.3, False, None, 0)
    with torch.no_grad():
        ref = torch.relu(lin(x))
    assert (got - ref).abs().max().item() < 1e-5
    hdr = torch.tensor([1234, 0], dtype=torch.int64, device=DEV)
    a = fused_head.mlp_dropout(x, lin, 0.3, True, hdr, 0)
    b = fused_head.mlp_dropout(x, lin, 0.3, True, hdr, 0)
    assert torch.equal(a, b), "same {seed, offset} and stream => same masks"
    hdr2 = torch.tensor([1234, 999], dtype=torch.int64, device=DEV)
    c = fused_head.mlp_dropout(x, lin, 0.3, True, hdr2, 0)
    assert not torch.equal(a, c)
    # output dropout: zeros where ReLU was positive occur at rate ~p
    base = fused_head.mlp_dropout(x, lin, 0.0, True, hdr, 0)
    pos = base > 0
    with torch.no_grad():
        x1 = torch.ones(4096, 256, device=DEV)
        lin1 = torch.nn.Linear(256, 256).to(DEV)
        lin1.weight.fill_(1.0 / 256)
        lin1.bias.fill_(1.0)
    d = fused_head.mlp_dropout(x1, lin1, 0.3, True, hdr, 0)
    drop_rate = (d == 0).float().mean().item()
    assert 0.28 < drop_rate < 0.32, drop_rate
    kept = d[d != 0]
    # kept outputs are (bias + mean of kept&scaled inputs) / (1-p): mean ~ (1 + 1) / 0.7
    assert abs(kept.mean().item() - 2.0 / 0.7) < 0.02
    assert pos.any()


def test_fuse_loss_grad_and_adam_match_torch():
    import b200rnn
    from b200rnn import _lib

    lib = _lib.load()
    torch.manual_seed(3)
    B, Ht, Ha = 128, 128, 256
    tf = torch.randn(B, Ht, device=DEV)
    af = torch.randn(B, Ha, device=DEV)
    y = torch.randint(0, 2, (B,), device=DEV)
    W = (torch.randn(2, Ht + Ha, device=DEV) * 0.05).requires_grad_(True)
    Wk = W.detach().clone()

    class _M:  # what MyLoss reads
        fc_final = [type("L", (), {"weight": W})()]

    opt = torch.optim.Adam([W], lr=8e-6)
    grad = torch.zeros(2 * (Ht + Ha), device=DEV)
    m = torch.zeros_like(grad)
    v = torch.zeros_like(grad)
    step = torch.zeros((), device=DEV)
    loss_k = torch.zeros((), device=DEV)
    probs = torch.empty(B, 2, device=DEV)
    stream = torch.cuda.current_stream().cuda_stream
    for it in range(4):
        opt.zero_grad()
        loss = b200rnn.MyLoss(Ht)(tf, af, y, _M)
        loss.backward()
        ref_probs = torch.softmax(torch.cat((tf, af), 1) @ W.detach().t(), dim=1)
        _lib.check(lib.b200rnn_fuse_loss_grad(tf.data_ptr(), Ht, af.data_ptr(), Ha, y.data_ptr(), B, Wk.data_ptr(),
                                              grad.data_ptr(), 0, loss_k.data_ptr(), probs.data_ptr(), stream), "loss")
        assert abs(loss_k.item() - loss.item()) < 2e-6
        assert (grad.view(2, -1) - W.grad).abs().max().item() < 1e-6 * max(1.0, W.grad.abs().max().item())
        assert (probs - ref_probs).abs().max().item() < 1e-6
        opt.step()
        _lib.check(lib.b200rnn_adam(Wk.data_ptr(), grad.data_ptr(), m.data_ptr(), v.data_ptr(), step.data_ptr(),
                                    Wk.numel(), 8e-6, 0.9, 0.999, 1e-8, stream), "adam")
        assert (Wk - W.detach()).abs().max().item() < 1e-7
    assert step.item() == 4.0


def test_fused_fuse_step_equals_generic_path_in_eval():
    """Dropout off (eval): the fused step must reproduce the generic PyTorch-shell step (which the golden tests pin to
    the reference): features, probabilities, loss and the updated fc_final weight."""
    import b200rnn

    m1 = _fuse_model(train=False)
    m2 = _fuse_model(train=False)
    m2.load_state_dict(m1.state_dict())
    torch.manual_seed(5)
    audio = torch.randn(16, 20, 256, device=DEV)
    text = torch.randn(16, 6, 1024, device=DEV)
    y = torch.randint(0, 2, (16,), device=DEV)
    batch = b200rnn.FuseBatch(audio, text)
    # generic
    opt = torch.optim.Adam([m1.fc_final[0].weight], lr=1e-3)
    tf, af = m1.pretrained_feature(batch)
    out = m1(torch.cat((tf, af), 1))
    loss = b200rnn.MyLoss(128)(tf, af, y, m1)
    loss.backward()
    opt.step()
    # fused
    step = b200rnn.FusedFuseStep(m2, lr=1e-3)
    tf2, af2 = step.features(batch)
    probs, loss2 = step(batch, y)
    torch.cuda.synchronize()
    assert (tf2 - tf).abs().max().item() < 1e-5 and (af2 - af).abs().max().item() < 1e-4
    assert (probs - out.detach()).abs().max().item() < 1e-5
    assert abs(loss2.item() - loss.item()) < 1e-5
    assert (m2.fc_final[0].weight - m1.fc_final[0].weight).abs().max().item() < 1e-6


def test_fused_fuse_step_train_mode_runs_and_varies():
    import b200rnn

    m = _fuse_model(train=True)
    step = b200rnn.FusedFuseStep(m)
    batch = b200rnn.FuseBatch(torch.randn(8, 12, 256, device=DEV), torch.randn(8, 5, 1024, device=DEV))
    y = torch.randint(0, 2, (8,), device=DEV)
    _, l1 = step(batch, y)
    a = l1.item()
    _, l2 = step(batch, y)
    assert a != l2.item(), "train-mode dropout must change the features between steps"
    assert torch.isfinite(m.fc_final[0].weight).all()


def test_flat_adamw_matches_torch_adamw_with_reference_grouping():
    """audio_gru_whole.py:247-255, 307: AdamW(lr 6e-6; wd 1e-5, and 0 for 'ln'); three full training steps."""
    import copy

    import b200rnn

    torch.manual_seed(11)
    cfg = dict(num_classes=2, dropout=0.0, rnn_layers=2, embedding_size=256, hidden_dims=256)
    m1 = b200rnn.AudioBiLSTM(cfg).to(DEV).train()
    m2 = copy.deepcopy(m1)
    named = list(m1.named_parameters())
    groups = [{"params": [p for n, p in named if "ln" not in n], "weight_decay": 1e-2},
              {"params": [p for n, p in named if "ln" in n], "weight_decay": 0.0}]
    lr = 1e-3   # large enough that three steps move the weights measurably
    opt1 = torch.optim.AdamW(groups, lr=lr)
    opt2 = b200rnn.FlatAdamW.like_reference(m2, lr=lr, weight_decay=1e-2)
    x = torch.randn(8, 10, 256, device=DEV)
    y = torch.randint(0, 2, (8,), device=DEV)
    crit = torch.nn.CrossEntropyLoss()
    for _ in range(3):
        opt1.zero_grad()
        crit(m1(x), y).backward()
        opt1.step()
        opt2.zero_grad()
        crit(m2(x), y).backward()
        opt2.step()
    torch.cuda.synchronize()
    for (n, a), (_, b) in zip(m1.named_parameters(), m2.named_parameters()):
        if a.grad is None:
            continue
        assert (a - b).abs().max().item() <= 2e-6 + 1e-4 * lr, n   # 3 steps of size <= lr each
    assert opt2.step_count.item() == 3.0
    # the parameters really live in the flat buffers now
    g0 = opt2.groups[0]
    assert all(p.data_ptr() >= g0.flat_p.data_ptr() for p in g0.params)


def test_fuse_head_train_mode_equals_the_unfused_kernel_chain_with_the_same_rng():
    """Train mode (dropout 0.3): the single-launch head draws the same Philox streams as attention_pool +
    mlp_dropout(stream 0) + mlp_dropout(stream 2) run one after the other on the same {seed, offset}, so the features
    must agree to rounding although dropout is on."""
    import b200rnn
    from b200rnn import fused_head
    from b200rnn.functional import rnn_forward_fused

    m = _fuse_model(train=True)
    m.lstm_net.eval()            # keep the encoders deterministic: only the head dropout is under test here
    m.lstm_net_audio.eval()
    step = b200rnn.FusedFuseStep(m)
    torch.manual_seed(8)
    batch = b200rnn.FuseBatch(torch.randn(33, 12, 256, device=DEV), torch.randn(33, 7, 1024, device=DEV))
    state0 = step.rng_state.clone()
    tf, af = step.features(batch)
    hdr = step.rng_hdr.clone()
    assert hdr[0].item() == state0[0].item() and hdr[1].item() == state0[1].item()
    assert step.rng_state[1].item() > state0[1].item(), "the device-side RNG offset must advance"
    with torch.no_grad():
        seq, h_n, _ = rnn_forward_fused(batch.text.permute(1, 0, 2), m.lstm_net._flat_weights, m.lstm_net._config(),
                                        m.lstm_net._rng_state)
        ctx = fused_head.attention_pool(seq, h_n, m.attention_layer)
        tf_ref = fused_head.mlp_dropout(ctx, m.fc_out[1], 0.3, True, hdr, 0)
        pooled = m.lstm_net_audio.forward_ln_sum(batch.audio, m.ln)
        af_ref = fused_head.mlp_dropout(pooled, m.fc_audio[1], 0.3, True, hdr, 2)
    torch.cuda.synchronize()
    assert (tf == 0).float().mean().item() > 0.2, "output dropout must zero ~30 % (plus ReLU zeros)"
    assert ((tf == 0) == (tf_ref == 0)).all() and ((af == 0) == (af_ref == 0)).all(), "same masks"
    assert (tf - tf_ref).abs().max().item() < 1e-5
    assert (af - af_ref).abs().max().item() < 2e-4   # |pooled| ~ 120 * |h|: rounding of a differently ordered dot


@pytest.mark.parametrize("train", [False, True])
def test_fused_fuse_step_regression_flavour_matches_cpu_oracle(train):
    """Regression/fuse_net.py:345-366, 373-412: sigmoid(modal_attn) gate + ReLU output, two-head SmoothL1 MyLoss, Adam on
    fc_final.0.weight - three steps against oracle.ref_models.RefFusion(regression=True) on CPU (dropout 0 in train)."""
    import b200rnn
    from oracle import ref_models

    args = dict(text_embed_size=1024, text_hidden_dims=128, rnn_layers=2, dropout=0.0 if train else 0.3, num_classes=1,
                audio_hidden_dims=256, audio_embed_size=256)
    torch.manual_seed(0)
    ref = ref_models.RefFusion(regression=True, **args)
    mine = b200rnn.fusion_net(regression=True, **args)
    mine.load_state_dict(ref.state_dict())
    mine = mine.to(DEV)
    ref.train(train)
    mine.train(train)
    lr = 1e-3
    opt = torch.optim.Adam([ref.fc_final[0].weight], lr=lr)
    step = b200rnn.FusedFuseStep(mine, lr=lr)
    g = torch.Generator().manual_seed(17)
    for _ in range(3):
        audio, text = torch.randn(20, 15, 256, generator=g), torch.randn(20, 5, 1024, generator=g)
        y = torch.rand(20, generator=g) * 3.0          # PHQ-like targets around the SmoothL1 knee
        opt.zero_grad()
        tf_r, af_r = ref.pretrained_feature_tensors(audio, text)
        out_r = ref(torch.cat((tf_r, af_r), dim=1))
        loss_r = ref_models.ref_fusion_loss(tf_r, af_r, y.view(-1, 1), ref)
        loss_r.backward()
        opt.step()
        out_m, loss_m = step(b200rnn.FuseBatch(audio.to(DEV), text.to(DEV)), y.to(DEV))
        torch.cuda.synchronize()
        assert (out_m.cpu() - out_r.detach()).abs().max().item() < 1e-4
        assert abs(loss_m.item() - loss_r.item()) < 1e-5 * max(1.0, abs(loss_r.item()))
        assert (mine.fc_final[0].weight.detach().cpu() - ref.fc_final[0].weight.detach()).abs().max().item() < 2e-6


def test_fused_fuse_step_rejects_bad_labels():
    import b200rnn

    m = _fuse_model(train=False)
    step = b200rnn.FusedFuseStep(m)
    batch = b200rnn.FuseBatch(torch.randn(4, 6, 256, device=DEV), torch.randn(4, 3, 1024, device=DEV))
    with pytest.raises(ValueError):
        step(batch, torch.zeros(3, dtype=torch.int64, device=DEV))
    _, loss = step(batch, torch.tensor([0, 1, 1, 0], dtype=torch.int32, device=DEV))   # int32 is converted, not misread
    assert torch.isfinite(loss)
    _, loss = step(batch, torch.tensor([0, 1, 2, 0], device=DEV))                      # class 2 does not exist
    assert torch.isnan(loss)


@pytest.mark.parametrize("T,B,H", [(30, 64, 256), (6, 5, 128), (1, 3, 128)])
def test_attention_pool_autograd_function_matches_torch(T, B, H):
    """b200rnn_attention_pool / _bwd as one autograd Function vs the PyTorch expression of text_bilstm_whole.py:74-99:
    context, d seq (both halves), d h_n, d attention_layer weight and bias."""
    import b200rnn
    from b200rnn import fused_head

    torch.manual_seed(21)
    layer = torch.nn.Sequential(torch.nn.Linear(H, H), torch.nn.ReLU(inplace=True)).to(DEV)
    seq = (torch.randn(T, B, 2 * H, device=DEV) * 0.5).requires_grad_(True)
    h_n = (torch.randn(4, B, H, device=DEV) * 0.5).requires_grad_(True)
    w = torch.randn(B, H, device=DEV)
    got = fused_head.attention_pool_tm(layer, seq, h_n)
    (got * w).sum().backward()
    g1 = [seq.grad.clone(), h_n.grad.clone(), layer[0].weight.grad.clone(), layer[0].bias.grad.clone()]
    for t in (seq, h_n, layer[0].weight, layer[0].bias):
        t.grad = None
    ref = b200rnn.attention_pool(layer, seq.permute(1, 0, 2), h_n.permute(1, 0, 2))
    (ref * w).sum().backward()
    g2 = [seq.grad, h_n.grad, layer[0].weight.grad, layer[0].bias.grad]
    assert (got - ref).abs().max().item() < 1e-5
    for a, b, name in zip(g1, g2, ("dseq", "dh_n", "dW", "db")):
        assert (a - b).abs().max().item() <= 2e-5 * max(1.0, b.abs().max().item()), name
