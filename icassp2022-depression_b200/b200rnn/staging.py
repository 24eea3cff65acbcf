"""Input staging for the fuse path.

The reference converts a python list of numpy arrays with ``torch.tensor(list)`` on every step
(fuse_net_whole.py:343) — 70 % of its CPU step time at B=128. Here the list is packed once into pinned host
buffers (one per modality) and shipped with two asynchronous H2D copies on the current stream.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Sequence

import numpy as np
import torch


@dataclass
class FuseBatch:
    """Device-resident batch of the fuse model: audio [B,T_a,E_a], text [B,T_t,E_t] (float32)."""

    audio: torch.Tensor
    text: torch.Tensor


class PinnedStager:
    """Reusable pinned host buffers + device buffers for a fixed batch shape (double-buffer friendly)."""

    def __init__(self, audio_shape, text_shape, device):
        self.device = torch.device(device)
        pin = self.device.type == "cuda"
        self.h_audio = torch.empty(audio_shape, dtype=torch.float32, pin_memory=pin)
        self.h_text = torch.empty(text_shape, dtype=torch.float32, pin_memory=pin)
        self.d_audio = torch.empty(audio_shape, dtype=torch.float32, device=self.device)
        self.d_text = torch.empty(text_shape, dtype=torch.float32, device=self.device)

    @property
    def h2d_bytes(self) -> int:
        return self.h_audio.numel() * 4 + self.h_text.numel() * 4

    def fill_host(self, pairs: Sequence) -> None:
        a, t = self.h_audio.numpy(), self.h_text.numpy()
        for i, ele in enumerate(pairs):
            a[i] = ele[0]
            t[i] = ele[1]

    def to_device(self) -> FuseBatch:
        """Enqueue the two H2D copies on the current stream (async w.r.t. the host when pinned)."""
        self.d_audio.copy_(self.h_audio, non_blocking=True)
        self.d_text.copy_(self.h_text, non_blocking=True)
        return FuseBatch(self.d_audio, self.d_text)


def stage_fuse_batch(pairs: Sequence, device) -> FuseBatch:
    """One-shot staging of the reference's ``[[audio(T,Ea), text(T,Et)], ...]`` list."""
    first_a, first_t = np.asarray(pairs[0][0]), np.asarray(pairs[0][1])
    B = len(pairs)
    st = PinnedStager((B, *first_a.shape), (B, *first_t.shape), device)
    st.fill_host(pairs)
    batch = st.to_device()
    if torch.device(device).type == "cuda":
        torch.cuda.current_stream().synchronize()  # the pinned buffers die with `st`
    return batch
