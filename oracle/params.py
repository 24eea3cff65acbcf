"""ORACLE — test infrastructure only. Deterministic, name-keyed parameter values.

Golden fixtures would be several MB if they carried the models' weights, so they carry none: every parameter is
regenerated from its NAME (crc32 -> numpy PCG64 stream, stable across numpy versions) both when the fixture is
made from the reference classes and when a test fills the oracle / CUDA model. Independent of torch's RNG and of
parameter creation order.
"""
from __future__ import annotations

import zlib

import numpy as np
import torch


def tensor_for(name: str, shape, scale: float | None = None) -> torch.Tensor:
    rng = np.random.default_rng(zlib.crc32(name.encode("utf-8")))
    shape = tuple(shape)
    if scale is None:
        fan = shape[-1] if len(shape) > 1 else shape[0]
        scale = 1.0 / np.sqrt(max(fan, 1))
    if "ln" in name.split(".")[0] and name.endswith("weight"):
        arr = 1.0 + rng.uniform(-0.1, 0.1, size=shape)  # LayerNorm gains stay near 1
    else:
        arr = rng.uniform(-scale, scale, size=shape)
    return torch.from_numpy(arr.astype(np.float32))


@torch.no_grad()
def fill_module(module: torch.nn.Module) -> None:
    """Overwrite every parameter of ``module`` with its name-keyed value."""
    for name, p in module.named_parameters():
        p.copy_(tensor_for(name, p.shape).to(p.device))


def inputs_for(tag: str, shape) -> torch.Tensor:
    rng = np.random.default_rng(zlib.crc32(("input:" + tag).encode("utf-8")))
    return torch.from_numpy(rng.standard_normal(size=tuple(shape)).astype(np.float32))


def labels_for(tag: str, n: int, num_classes: int = 2) -> torch.Tensor:
    rng = np.random.default_rng(zlib.crc32(("label:" + tag).encode("utf-8")))
    return torch.from_numpy(rng.integers(0, num_classes, size=n).astype(np.int64))


def summarize(t: torch.Tensor, full_limit: int = 4096) -> dict:
    """Compact, order-stable fingerprint of a tensor for the golden fixtures (full copy when small)."""
    a = t.detach().cpu().double().reshape(-1).numpy()
    if a.size <= full_limit:
        return {"full": a.astype(np.float32)}
    stride = max(1, a.size // 509)
    return {
        "sum": np.array([a.sum()], dtype=np.float64),
        "abssum": np.array([np.abs(a).sum()], dtype=np.float64),
        "head": a[:64].astype(np.float32),
        "sample": a[::stride][:512].astype(np.float32),
        "absmax": np.array([np.abs(a).max()], dtype=np.float64),
    }
