"""Data-parallel path on CPU: world_size-2 gloo processes. Checks the one-allreduce flat gradient bucket and the
batch sharding: shard gradients, summed over ranks and averaged, equal the full-batch gradient."""
import os
import socket
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import PKG, ROOT


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, out):
    for p in (ROOT, PKG):
        if p not in sys.path:
            sys.path.insert(0, p)
    import b200rnn
    from oracle import ref_models

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        torch.manual_seed(100 + rank)      # deliberately different init per rank ...
        cfg = dict(num_classes=2, dropout=0.0, rnn_layers=1, embedding_size=16, hidden_dims=8, bidirectional=False)
        model = ref_models.RefAudio(cfg).eval()   # stock torch.nn on CPU: the DP plumbing is device-agnostic
        b200rnn.broadcast_parameters(model)       # ... made identical by one flat broadcast from rank 0
        bucket = b200rnn.GradBucket(model)
        g = torch.Generator().manual_seed(7)
        x = torch.randn(8, 5, 16, generator=g)
        y = torch.randint(0, 2, (8,), generator=g)
        sl = b200rnn.shard_batch(8, rank, world)
        bucket.zero()
        loss = torch.nn.functional.cross_entropy(model(x[sl]), y[sl])   # mean over the shard
        loss.backward()
        assert all(p.grad.data_ptr() == v.data_ptr() for p, v in bucket._bound)
        bucket.allreduce(average=True)
        if rank == 0:
            ref = ref_models.RefAudio(cfg).eval()
            ref.load_state_dict(model.state_dict())
            torch.nn.functional.cross_entropy(ref(x), y).backward()
            # views are 256-byte aligned inside the flat bucket (padding stays zero): compare view by view;
            # unused params (attention_layer) stay zero
            gmax = max(p.grad.abs().max().item() for p in ref.parameters() if p.grad is not None)
            err = 0.0
            for (p, v), q in zip(bucket._bound, ref.parameters()):
                gr = q.grad if q.grad is not None else torch.zeros_like(q)
                err = max(err, (v - gr).abs().max().item() / gmax)
                assert (v.data_ptr() - bucket.flat.data_ptr()) % 256 == 0
            torch.save({"err": err, "numel": bucket.numel, "nparams": len(bucket.params)}, out)
        dist.barrier()
    finally:
        dist.destroy_process_group()


def test_two_rank_gloo_bucket_allreduce_equals_full_batch(tmp_path):
    out = str(tmp_path / "res.pt")
    port = _free_port()
    mp.spawn(_worker, args=(2, port, out), nprocs=2, join=True)
    res = torch.load(out)
    assert res["err"] < 1e-5, res
    assert res["numel"] > 0


def test_shard_batch_partitions_exactly():
    import b200rnn

    for n in (1, 7, 128, 1024):
        for world in (1, 2, 3, 8):
            rows = []
            for r in range(world):
                s = b200rnn.shard_batch(n, r, world)
                rows.extend(range(s.start, s.stop))
            assert rows == list(range(n))


def test_bucket_survives_zero_grad_set_to_none():
    """The reference loops call optimizer.zero_grad() (audio_gru_whole.py:183), which drops .grad since torch 2.0.
    The bucket re-attaches its views: two steps through torch's zero_grad give the same bucket as bucket.zero()."""
    import b200rnn
    from oracle import ref_models

    torch.manual_seed(5)
    cfg = dict(num_classes=2, dropout=0.0, rnn_layers=1, embedding_size=16, hidden_dims=8, bidirectional=False)
    model = ref_models.RefAudio(cfg).eval()
    bucket = b200rnn.GradBucket(model)
    opt = torch.optim.SGD(model.parameters(), lr=0.0)
    x, y = torch.randn(6, 5, 16), torch.randint(0, 2, (6,))
    results = []
    for zero in (lambda: opt.zero_grad(), lambda: opt.zero_grad(set_to_none=False), bucket.zero, bucket.zero_grad):
        for _ in range(2):   # second pass would double the gradient if stale values survived
            zero()
            torch.nn.functional.cross_entropy(model(x), y).backward()
            bucket.allreduce()
        assert all(p.grad.data_ptr() == v.data_ptr() for p, v in bucket._bound)
        results.append(bucket.flat.clone())
    for r in results[1:]:
        assert torch.equal(r, results[0])
    assert results[0].abs().max().item() > 0
