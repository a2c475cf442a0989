#!/usr/bin/env python
"""Row-partitioned propagation of every layer family at config-5 scale (run under torchrun on N GPUs of one box):
GCN, LightGCN, GraphSage(mean), GAT, 3 layers x 128, peer exchange.  One JSON line per family from rank 0:
ms per propagation (max over ranks, CUDA events) and aggregate edges/s = K * nnz / t."""
import json
import os
import sys

import torch
import torch.distributed as dist

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
from deep_cbrs_amar_renaissance_b200 import ops  # noqa: E402
from deep_cbrs_amar_renaissance_b200.distributed import RowPartition  # noqa: E402
from deep_cbrs_amar_renaissance_b200.graph import DeviceGraph  # noqa: E402
from deep_cbrs_amar_renaissance_b200.keras_like import set_seed  # noqa: E402
from deep_cbrs_amar_renaissance_b200.models import gnn  # noqa: E402


def main():
    scale = sys.argv[1] if len(sys.argv) > 1 else "c5"
    n_users, n_items, n_edges = {"c5": (10_000_000, 1_000_000, 1_000_000_000), "c5-tenth": (1_000_000, 100_000, 100_000_000)}[scale]
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    n = n_users + n_items
    for name, view in (("GCN", "norm"), ("LightGCN", "norm"), ("GraphSage", "raw"), ("GAT", "raw")):
        row, col = ops.synth_bipartite(n_users, n_items, n_edges, 42, dev)
        graph = DeviceGraph(row, col, None, n)
        del row, col
        set_seed(42)
        kw = dict(embedding_dim=128, final_node="concatenation")
        model = getattr(gnn, name)(graph, n_layers=3, **kw) if name == "LightGCN" else getattr(gnn, name)(graph, n_hiddens=[128] * 3, **kw)
        seq = model.gnn_layers
        nnz = getattr(graph, view).nnz
        part = RowPartition([n_users, n_items], final_types=[1]).attach(seq)
        part.csr_slices(view, graph)
        part.release_full_views(graph)
        torch.cuda.empty_cache()
        for _ in range(2):
            model(None)
        dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        steps = 3
        for _ in range(steps):
            model(None)
        e1.record()
        dist.barrier()
        torch.cuda.synchronize()
        ms = torch.tensor([e0.elapsed_time(e1) / steps], device=dev, dtype=torch.float64)
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        part.heap.check()
        if rank == 0:
            edges = nnz + (n if name == "GAT" else 0)
            print(json.dumps({"family": name, "scale": scale, "n_gpus": world, "layers": 3, "dim": 128, "edges_per_layer": edges,
                              "ms_per_propagation": ms.item(), "edges_per_s": 3 * edges / (ms.item() * 1e-3),
                              "exchange": part.exchange, "pipeline": part.pipeline}), flush=True)
        part.close()
        seq.partition = None
        del model, seq, graph, part
        torch.cuda.empty_cache()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
