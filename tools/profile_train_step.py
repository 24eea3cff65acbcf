#!/usr/bin/env python
"""Run a few EAGER train steps of one BASELINE module config so that `ncu` can list their kernels:
    ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file out.csv \
        python tools/profile_train_step.py c2|c3|ft [steps]
c2 = audio_gru_whole train (B=64, T=120), c3 = text_bilstm_whole train (B=64, T=30, H=256), ft = fuse fine-tune (B=128)."""
import os
import sys

import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "icassp2022-depression_b200"))
import b200rnn  # noqa: E402

kind = sys.argv[1] if len(sys.argv) > 1 else "c2"
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
dev = torch.device("cuda:0")
torch.manual_seed(0)
if kind == "ft":
    model = b200rnn.fusion_net(1024, 128, 2, 0.3, 2, 256, 256).to(dev).train()
    opt = b200rnn.FlatAdamW([{"params": list(model.parameters()), "weight_decay": 0.0}], lr=8e-6, model=model)
    ts = b200rnn.FuseFineTuneStep(model, opt, 128, 120, 30, use_graph=False)
    args = (torch.randn(128, 120, 256, device=dev), torch.randn(128, 30, 1024, device=dev),
            torch.randint(0, 2, (128,), device=dev))
else:
    if kind == "c2":
        cfg = dict(num_classes=2, dropout=0.5, rnn_layers=2, embedding_size=256, hidden_dims=256)
        model, shape = b200rnn.AudioBiLSTM(cfg).to(dev).train(), (64, 120, 256)
    else:
        cfg = dict(num_classes=2, dropout=0.5, rnn_layers=2, embedding_size=1024, hidden_dims=256, bidirectional=True)
        model, shape = b200rnn.TextBiLSTM(cfg).to(dev).train(), (64, 30, 1024)
    opt = b200rnn.FlatAdamW.like_reference(model, lr=6e-6, weight_decay=1e-5)
    ts = b200rnn.TrainStep(model, opt, shape, use_graph=False)
    args = (torch.randn(*shape, device=dev), torch.randint(0, 2, (shape[0],), device=dev))
for _ in range(steps):
    ts.step(*args)
torch.cuda.synchronize()
print("done", kind, float(ts.loss_value))
