#!/usr/bin/env python
"""Multi-GPU check of the in-kernel NVLink gradient exchange of b200rnn.FusedFuseStep (run under torchrun, N >= 2):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 tools/dp_exchange_check.py

Every rank runs the same 5 eval-mode steps on its own shard twice - once with exchange="peer" (peer stores + flags
inside b200rnn_fuse_head) and once with exchange="nccl" (ncclAllReduce + b200rnn_adamw) - from identical weights, and
rank 0 additionally replays the GLOBAL batch on one GPU (gradient of the mean over the global batch). Checks:
  * all ranks end with bit-identical fc_final.0.weight under "peer" (same summation order everywhere),
  * "peer" == "nccl" == single-GPU global batch to 1e-7,
  * a CUDA-graph capture of the peer step replays correctly (device-resident step counter / parity),
  * exchange="peer_async" (gradient sent in step s, summed + applied at the start of step s+1, flush() at the end)
    ends with the same weights, eagerly and as a replayed graph.
Prints one JSON line on rank 0; exit code 0 = all checks passed.
"""
import copy
import json
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "icassp2022-depression_b200"))
import b200rnn  # noqa: E402


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    B, steps, lr = 16, 5, 1e-3
    torch.manual_seed(0)
    base = b200rnn.fusion_net(1024, 128, 2, 0.3, 2, 256, 256).to(dev).eval()
    b200rnn.broadcast_parameters(base)
    g = torch.Generator().manual_seed(77)
    audio = torch.randn(steps, B * world, 20, 256, generator=g)
    text = torch.randn(steps, B * world, 6, 1024, generator=g)
    labels = torch.randint(0, 2, (steps, B * world), generator=g)
    sl = slice(rank * B, (rank + 1) * B)

    def run(exchange, use_graph=False):
        m = copy.deepcopy(base)
        st = b200rnn.FusedFuseStep(m, lr=lr, exchange=exchange)
        assert st.exchange == exchange, (st.exchange, exchange)
        flush = st.flush
        if not use_graph:
            for s in range(steps):
                st(b200rnn.FuseBatch(audio[s, sl].to(dev), text[s, sl].to(dev)), labels[s, sl].to(dev))
        else:
            a, t, y = audio[0, sl].to(dev), text[0, sl].to(dev), labels[0, sl].to(dev)
            st(b200rnn.FuseBatch(a, t), y)                   # eager warm-up = step 0
            gr = torch.cuda.CUDAGraph()
            side = torch.cuda.Stream()
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):
                pass
            with torch.cuda.graph(gr):
                st(b200rnn.FuseBatch(a, t), y)
            for s in range(1, steps):
                a.copy_(audio[s, sl]); t.copy_(text[s, sl]); y.copy_(labels[s, sl])
                gr.replay()
            # the capture itself does not execute: steps 1..4 ran => 5 steps in total with the eager one
            del gr
        flush()                       # peer_async: the last step's update is still pending
        torch.cuda.synchronize()
        w = m.fc_final[0].weight.detach().clone()
        st.close()
        return w

    w_peer = run("peer")
    w_nccl = run("nccl")
    w_graph = run("peer", use_graph=True)
    w_async = run("peer_async")
    w_async_graph = run("peer_async", use_graph=True)
    gathered = [torch.empty_like(w_peer) for _ in range(world)]
    dist.all_gather(gathered, w_peer)
    res = {"world": world}
    res["peer_identical_across_ranks"] = all(torch.equal(gathered[0], x) for x in gathered)
    res["peer_vs_nccl_max_abs"] = (w_peer - w_nccl).abs().max().item()
    res["graph_vs_eager_max_abs"] = (w_graph - w_peer).abs().max().item()
    res["async_vs_sync_max_abs"] = (w_async - w_peer).abs().max().item()
    res["async_graph_vs_sync_max_abs"] = (w_async_graph - w_peer).abs().max().item()
    if rank == 0:
        m = copy.deepcopy(base)
        st = b200rnn.FusedFuseStep(m, lr=lr, exchange="none")      # single-GPU replay of the global batch
        for s in range(steps):
            st(b200rnn.FuseBatch(audio[s].to(dev), text[s].to(dev)), labels[s].to(dev))
        torch.cuda.synchronize()
        res["peer_vs_single_gpu_global_batch_max_abs"] = (w_peer - m.fc_final[0].weight.detach()).abs().max().item()
        res["weight_moved"] = (w_peer - base.fc_final[0].weight.detach()).abs().max().item()
        ok = (res["peer_identical_across_ranks"] and res["peer_vs_nccl_max_abs"] <= 1e-7 and
              res["graph_vs_eager_max_abs"] <= 1e-7 and res["async_vs_sync_max_abs"] <= 1e-7 and
              res["async_graph_vs_sync_max_abs"] <= 1e-7 and res["peer_vs_single_gpu_global_batch_max_abs"] <= 1e-7 and
              res["weight_moved"] > 1e-4)
        res["ok"] = bool(ok)
        print(json.dumps(res), flush=True)
    dist.barrier()
    dist.destroy_process_group()
    if rank == 0 and not res["ok"]:
        sys.exit(1)


if __name__ == "__main__":
    main()
