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


def bind_host_thread_to_gpu_numa_node(device) -> dict:
    """Pin the calling process to the CPUs of the NUMA node the GPU hangs off, BEFORE the pinned staging buffers are
    allocated (first-touch places them on that node). On an 8-GPU box every rank otherwise allocates wherever the
    launcher happened to start it, and the H2D copies of the remote-node ranks cross the inter-socket link (measured
    at 8 ranks: 0.70 ms per 31.5 MB batch on the local ranks, 0.87 ms on the others). Best effort: returns what it did."""
    import os

    info = {"bound": False}
    try:
        dev = torch.device(device)
        idx = dev.index if dev.index is not None else torch.cuda.current_device()
        bus = torch.cuda.get_device_properties(idx).pci_bus_id if hasattr(torch.cuda.get_device_properties(idx), "pci_bus_id") else None
        domain = getattr(torch.cuda.get_device_properties(idx), "pci_domain_id", 0)
        devid = getattr(torch.cuda.get_device_properties(idx), "pci_device_id", 0)
        if bus is None:
            return info
        path = f"/sys/bus/pci/devices/{domain:04x}:{bus:02x}:{devid:02x}.0/numa_node"
        with open(path) as fh:
            node = int(fh.read().strip())
        info["numa_node"] = node
        if node < 0:
            return info
        with open(f"/sys/devices/system/node/node{node}/cpulist") as fh:
            cpus = set()
            for part in fh.read().strip().split(","):
                a, _, b = part.partition("-")
                cpus.update(range(int(a), int(b or a) + 1))
        allowed = cpus & set(os.sched_getaffinity(0))
        if allowed:
            os.sched_setaffinity(0, allowed)
            info.update(bound=True, cpus=len(allowed))
    except Exception as exc:  # noqa: BLE001 - a sandbox without sysfs / affinity rights must not break the run
        info["error"] = f"{type(exc).__name__}: {exc}"
    return info


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
