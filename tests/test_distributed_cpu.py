"""CPU, world_size 2 over gloo: the N>1 host logic of the train step — batch sharding, the
1/B_global loss/gradient scaling and the flat-buffer all-reduce (loss riding in the last slot).
The per-rank arithmetic is done by the ORACLE here (there is no GPU in this container); on the
GPU the same scaling feeds ddfa_graph_label_bce and the NCCL all-reduce (trainer.py)."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from deepdfa_b200 import batched_graph as G
from deepdfa_b200 import synth
from oracle import ggnn_oracle as O

FEAT = "_ABS_DATAFLOW_api_all_limitall_1000_limitsubkeys_1000"


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        torch.manual_seed(0)
        model = O.OracleFlowGNNGGNN(FEAT, 40, 8, 3, 2, concat_all_absdf=True, positive_weight=2.0).double()
        full = synth.make_batch(10, 12, seed=4, variable=True, input_dim=40, vuln_rate=0.5)
        shard = G.split_batch(full, world)[rank]
        b_global = full.batch_size
        # local SUM of per-graph BCE terms scaled by 1/B_global  (== trainer.py's loss_scale / grad_scale)
        label = model.get_label(shard)
        out = model(shard)
        out = out.unsqueeze(0) if out.dim() == 0 else out
        lw = 1 + (2.0 - 1) * label
        terms = (1 - label) * out + lw * (torch.log1p(torch.exp(-out.abs())) + torch.clamp(-out, min=0))
        loss_local = terms.sum() / b_global
        loss_local.backward()
        flat = torch.cat([p.grad.reshape(-1) for p in model.parameters()] + [loss_local.detach().reshape(1)])
        dist.all_reduce(flat, op=dist.ReduceOp.SUM)
        if rank == 0:
            q.put(flat)
    finally:
        dist.destroy_process_group()


def test_sharded_step_equals_global_step():
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    flat = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    torch.manual_seed(0)
    model = O.OracleFlowGNNGGNN(FEAT, 40, 8, 3, 2, concat_all_absdf=True, positive_weight=2.0).double()
    full = synth.make_batch(10, 12, seed=4, variable=True, input_dim=40, vuln_rate=0.5)
    loss, _ = model.training_loss(full)          # reference: mean BCE over the global batch (base_module.py:74,183)
    loss.backward()
    ref = torch.cat([p.grad.reshape(-1) for p in model.parameters()] + [loss.detach().reshape(1)])
    assert torch.allclose(flat, ref, atol=1e-12)
