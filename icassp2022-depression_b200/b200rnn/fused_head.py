"""Fused model-shell kernels of the fuse step, Python side.

``FusedFuseStep`` runs one ``fuse_net_whole`` train step in reference semantics (fuse_net_whole.py:421-465: encoders
under no_grad with train-mode dropout, only ``fc_final.0.weight`` trainable, ``MyLoss``, Adam) with ~20 launches of
this library instead of ~90 framework launches: attention pooling, the two Dropout-Linear-ReLU-Dropout heads, the
two-head cross entropy with its weight gradient and the fused softmax output, and the Adam update are one kernel each.
It is a drop-in for ``pretrained_feature`` + ``forward`` + ``MyLoss`` + ``backward`` + ``optimizer.step`` when the
model is the classification ``fusion_net`` with two classes; anything else keeps the generic PyTorch path.
"""
from __future__ import annotations

import ctypes
from typing import Optional

import torch

from . import _lib
from .functional import _on, rnn_forward_fused
from .staging import FuseBatch


def _stream(device=None) -> int:
    return torch.cuda.current_stream(device).cuda_stream


def _require_cuda(*named) -> None:
    """Raw pointers go straight to CUDA kernels: a host tensor must fail here, loudly (no CPU path)."""
    for name, t in named:
        if t is not None and not t.is_cuda:
            raise _lib.B200RNNError(f"b200rnn: {name} is on {t.device}; the fused shell kernels run on CUDA only "
                                    "and have no CPU path")


@torch.no_grad()
def attention_pool(seq_tm: torch.Tensor, h_n: torch.Tensor, attention_layer: torch.nn.Module) -> torch.Tensor:
    """``attention_net_with_w`` on the time-major LSTM output ``seq_tm`` [T,B,2H] and ``h_n`` [L*D,B,H] -> [B,H]."""
    lib = _lib.load()
    _require_cuda(("seq", seq_tm), ("h_n", h_n), ("attention weight", attention_layer[0].weight))
    T, B, H2 = seq_tm.shape
    H = H2 // 2
    lin = attention_layer[0]
    ctx = torch.empty(B, H, dtype=torch.float32, device=seq_tm.device)
    h_n = h_n.contiguous()
    assert seq_tm.stride(2) == 1
    with _on(seq_tm.device):
        rc = lib.b200rnn_attention_pool(seq_tm.data_ptr(), seq_tm.stride(0), seq_tm.stride(1), h_n.data_ptr(),
                                        h_n.shape[0], B, T, H, lin.weight.data_ptr(), lin.bias.data_ptr(),
                                        ctx.data_ptr(), _stream(seq_tm.device))
    _lib.check(rc, "b200rnn_attention_pool")
    return ctx


@torch.no_grad()
def mlp_dropout(x: torch.Tensor, linear: torch.nn.Linear, p: float, training: bool, rng_hdr: Optional[torch.Tensor],
                stream_id: int) -> torch.Tensor:
    """``Dropout(p) -> linear -> ReLU -> Dropout(p)`` for a square ``linear`` (fc_out / fc_audio of fusion_net)."""
    lib = _lib.load()
    _require_cuda(("x", x), ("linear.weight", linear.weight), ("rng header", rng_hdr))
    B, n = x.shape
    assert linear.weight.shape == (n, n)
    out = torch.empty_like(x)
    with _on(x.device):
        rc = lib.b200rnn_mlp_dropout(x.data_ptr(), B, n, linear.weight.data_ptr(), linear.bias.data_ptr(),
                                     out.data_ptr(), int(training), float(p),
                                     rng_hdr.data_ptr() if rng_hdr is not None else None, stream_id, _stream(x.device))
    _lib.check(rc, "b200rnn_mlp_dropout")
    return out


class FusedFuseStep:
    """One reference-semantics ``fuse_net_whole`` train step on fused kernels (see module docstring)."""

    def __init__(self, model, lr: float = 8e-6, betas=(0.9, 0.999), eps: float = 1e-8, bucket=None):
        if getattr(model, "regression", False) or model.num_classes != 2:
            raise NotImplementedError("FusedFuseStep covers the 2-class classification fusion_net")
        self.model = model
        self.lr, self.betas, self.eps = float(lr), (float(betas[0]), float(betas[1])), float(eps)
        w = model.fc_final[0].weight
        _require_cuda(("model (fc_final.0.weight)", w))
        dev = w.device
        self.w = w
        self.bucket = bucket  # b200rnn.dp.GradBucket over the trainable parameter(s), or None (single process)
        self.grad = bucket.flat if bucket is not None else torch.zeros(w.numel(), device=dev)
        assert self.grad.numel() == w.numel(), "reference semantics: only fc_final.0.weight is trainable"
        self.m = torch.zeros(w.numel(), device=dev)
        self.v = torch.zeros(w.numel(), device=dev)
        self.step_count = torch.zeros((), device=dev)
        self.loss = torch.zeros((), device=dev)
        self.rng_hdr = torch.zeros(2, dtype=torch.int64, device=dev)
        self.rng_state = torch.tensor([(torch.initial_seed() * 2654435761 + 12345) & 0x7FFFFFFFFFFFFFFF, 0],
                                      dtype=torch.int64, device=dev)

    @torch.no_grad()
    def features(self, batch: FuseBatch):
        m = self.model
        lib = _lib.load()
        seq, h_n, _ = rnn_forward_fused(batch.text.permute(1, 0, 2), m.lstm_net._flat_weights, m.lstm_net._config(),
                                        m.lstm_net._rng_state)
        ctx = attention_pool(seq, h_n, m.attention_layer)
        training, p = m.training, m.dropout
        if training and p > 0:
            B = batch.text.shape[0]
            consume = (B * max(m.text_hidden_dims, m.audio_hidden_dims) + 3) // 4
            _lib.check(lib.b200rnn_rng_next(self.rng_hdr.data_ptr(), self.rng_state.data_ptr(), consume, _stream()),
                       "b200rnn_rng_next")
        text_feature = mlp_dropout(ctx, m.fc_out[1], p, training, self.rng_hdr, 0)
        pooled = m.lstm_net_audio.forward_ln_sum(batch.audio, m.ln)
        audio_feature = mlp_dropout(pooled, m.fc_audio[1], p, training, self.rng_hdr, 2)
        return text_feature, audio_feature

    @torch.no_grad()
    def __call__(self, batch: FuseBatch, labels: torch.Tensor):
        """Runs the step; returns (probs [B,2], loss scalar tensor)."""
        lib = _lib.load()
        m = self.model
        _require_cuda(("labels", labels), ("batch.audio", batch.audio), ("batch.text", batch.text))
        tf, af = self.features(batch)
        B = tf.shape[0]
        # the kernel reads `const int64_t labels[B]`: anything else (int32 from numpy, float, a strided view) would be
        # silently misread, so it is converted here; the values must be class indices 0/1 (checked on the device)
        if labels.dtype != torch.int64 or not labels.is_contiguous():
            labels = labels.to(torch.int64).contiguous()
        if labels.numel() != B:
            raise ValueError(f"FusedFuseStep: {labels.numel()} labels for a batch of {B}")
        probs = torch.empty(B, 2, dtype=torch.float32, device=tf.device)
        _lib.check(lib.b200rnn_fuse_loss_grad(tf.data_ptr(), tf.shape[1], af.data_ptr(), af.shape[1], labels.data_ptr(), B,
                                              self.w.data_ptr(), self.grad.data_ptr(), 0, self.loss.data_ptr(),
                                              probs.data_ptr(), _stream()), "b200rnn_fuse_loss_grad")
        if self.bucket is not None:
            self.bucket.allreduce()
        _lib.check(lib.b200rnn_adam(self.w.data_ptr(), self.grad.data_ptr(), self.m.data_ptr(), self.v.data_ptr(),
                                    self.step_count.data_ptr(), self.w.numel(), self.lr, self.betas[0], self.betas[1],
                                    self.eps, _stream()), "b200rnn_adam")
        return probs, self.loss
