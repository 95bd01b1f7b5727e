"""GPU: FusedTrainer behaviour beyond one step — training trajectories and decisions against the oracle trained on the same
stream (BASELINE configs[2] "F1 vs reference"), shape bucketing on a shuffled variable-size stream (datamodule.py:123-129),
captured-graph lifetime under workspace growth, run-ahead of the host over the arena id staging, unequal shards."""
import copy

import numpy as np
import pytest
import torch

import deepdfa_b200 as D
from deepdfa_b200 import batched_graph as G
from deepdfa_b200 import synth
from oracle import ggnn_oracle as O

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
FEAT = "_ABS_DATAFLOW_api_all_limitall_1000_limitsubkeys_1000"


learnable_batch = synth.make_learnable_batch


def f1_at_half(prob, label):
    from sklearn.metrics import f1_score
    return float(f1_score(label.astype(int), (prob > 0.5).astype(int), zero_division=0))


@pytest.mark.parametrize("engine", ["simt", "tcgen05"])
def test_training_decisions_and_f1_follow_the_oracle(engine):
    """BASELINE configs[2] ("F1 vs reference"), at a size the CPU oracle trains in seconds: 200 optimisation steps of both arms
    on ONE synthetic stream (same initial weights, same batches, Adam lr 1e-3 / wd 1e-2, BCE pos_weight), then both classify the
    same 768 held-out graphs.  Reported: loss curves, decision agreement and the F1 of either arm at the 0.5 threshold
    (base_module.py:186,348-383)."""
    torch.manual_seed(0)
    o = O.OracleFlowGNNGGNN(FEAT, 1002, 32, 8, 2, concat_all_absdf=True, positive_weight=1.5)
    m = D.FlowGNNGGNNModule(FEAT, 1002, 32, 8, 2, concat_all_absdf=True, positive_weight=1.5, engine=engine)
    m.load_state_dict(copy.deepcopy(o.state_dict()))
    m.to(DEV)
    tr = D.FusedTrainer(m)
    opt = O.make_optimizer(o)
    stream = [learnable_batch(24, 40, seed=300 + i) for i in range(40)]
    worst_loss = 0.0
    first = last = None
    for step in range(200):
        b = stream[step % len(stream)]
        opt.zero_grad()
        loss_ref, _ = o.training_loss(b)
        loss_ref.backward()
        opt.step()
        loss = float(tr.step(b))
        worst_loss = max(worst_loss, abs(loss - float(loss_ref)) / max(1.0, abs(float(loss_ref))))
        if step == 0:
            first = float(loss_ref)
        last = float(loss_ref)
    held = [learnable_batch(256, 40, seed=900 + i) for i in range(3)]
    probs_o, probs_m, labels = [], [], []
    with torch.no_grad():
        for b in held:
            probs_o.append(torch.sigmoid(o(b)).numpy())
            _, p, lab = m.validation_step((b, {}), 0)
            probs_m.append(p.cpu().numpy())
            labels.append(lab.cpu().numpy())
            assert np.array_equal(lab.cpu().numpy(), o.get_label(b).numpy().astype(np.int32))
    po, pm, y = np.concatenate(probs_o), np.concatenate(probs_m), np.concatenate(labels)
    agree = float(((po > 0.5) == (pm > 0.5)).mean())
    f1_o, f1_m = f1_at_half(po, y), f1_at_half(pm, y)
    print(f"train 200 steps engine={engine}: loss {first:.4f} -> {last:.4f}, worst rel loss gap {worst_loss:.2e}; held-out 768 graphs: "
          f"decision agreement {agree:.4f}, F1 oracle {f1_o:.4f} vs ours {f1_m:.4f}, max|dprob| {float(np.abs(po - pm).max()):.2e}")
    assert last < 0.7 * first, "the stream is learnable: the loss must fall"
    assert f1_o > 0.8, "the oracle arm must have learned the task for the comparison to mean anything"
    # the curves separate as rounding differences pass through 200 Adam steps (measured r03a: simt 1.4e-3, tcgen05 3.7e-2 at one
    # step of the steep part of the curve; gradient error 1e-6 vs 2e-5) — what must agree are the decisions and the F1 below
    assert worst_loss < (5e-3 if engine == "simt" else 8e-2)
    assert agree >= 0.99 and abs(f1_o - f1_m) <= 0.02


def _run_stream(trainer, batches, order, prefetch=False):
    out = []
    for k, i in enumerate(order):
        loss = trainer.step(batches[i])
        if prefetch and k + 1 < len(order):      # stage the next batch (host padding + H2D on the copy stream) while this step runs
            trainer.prefetch(batches[order[k + 1]])
        out.append(float(loss))
    return out


@pytest.mark.parametrize("engine", ["simt", "tcgen05"])
def test_bucketed_variable_stream_replays_graphs_and_matches_eager(engine):
    """A shuffled stream where every batch has its own (N, E): with bucket_nodes / bucket_edges the trainer pads each batch
    with one zero-weight dummy graph to a bucket shape and replays a few captured graphs; losses follow the eager, unpadded
    trainer (the dummy graph contributes nothing), and the parameters end up the same."""
    batches = [synth.make_batch(12, 30, seed=400 + i, variable=True, vuln_rate=0.4) for i in range(10)]
    shapes = {(b.num_nodes(), b.num_edges()) for b in batches}
    assert len(shapes) >= 8
    order = list(range(10)) * 4
    results = {}
    for mode in ("eager", "bucketed", "bucketed+prefetch"):
        torch.manual_seed(3)
        m = D.FlowGNNGGNNModule(FEAT, 1002, 32, 4, 2, concat_all_absdf=True, positive_weight=2.0, engine=engine).to(DEV)
        if mode == "eager":
            tr = D.FusedTrainer(m)
        else:
            tr = D.FusedTrainer(m, use_cuda_graph=True, bucket_nodes=256, bucket_edges=512, bucket_min_pad_nodes=8, max_graph_shapes=16)
        results[mode] = (_run_stream(tr, [b.pin_memory() for b in batches] if "prefetch" in mode else batches, order, prefetch="prefetch" in mode),
                         [p.detach().clone() for p in m.param_list()], tr)
    tr = results["bucketed"][2]
    nshapes = tr.num_bucket_shapes()
    assert 1 <= nshapes <= 4, nshapes                      # ten distinct shapes collapse onto a few bucket shapes
    captured = sum(1 for slot in tr._stream_slots.values() for st in slot["sets"] if st["graph"] is not None)
    assert captured >= nshapes                             # ... and those are replayed as CUDA graphs
    for mode in ("bucketed", "bucketed+prefetch"):
        for a, b in zip(results["eager"][0], results[mode][0]):
            assert abs(a - b) < 2e-5 * max(1.0, abs(a)), (mode, results["eager"][0][:6], results[mode][0][:6])
        for p, q in zip(results["eager"][1], results[mode][1]):
            assert float((p - q).abs().max()) < 5e-4       # 40 Adam steps at lr 1e-3; sign-level noise only


def test_captured_graphs_survive_workspace_growth():
    """order small (eager), small (capture), BIG (workspace regrows), small (replay).  The replayed graph bakes in the
    old workspace pointers: they must still be valid (retired, not freed) and the step must still be right."""
    small = [synth.make_batch(8, 30, seed=500 + i, vuln_rate=0.4).to(DEV) for i in range(2)]
    big = synth.make_batch(64, 120, seed=510, vuln_rate=0.4).to(DEV)
    order = [("s", 0), ("s", 0), ("s", 1), ("s", 1), ("b", 0), ("s", 0), ("s", 1), ("b", 0), ("s", 0)]
    losses = {}
    for mode in ("eager", "graph"):
        torch.manual_seed(4)
        m = D.FlowGNNGGNNModule(FEAT, 1002, 32, 3, 2, concat_all_absdf=True, engine="tcgen05").to(DEV)
        tr = D.FusedTrainer(m, use_cuda_graph=(mode == "graph"))
        cur = []
        for kind, i in order:
            cur.append(float(tr.step(small[i] if kind == "s" else big)))
            if mode == "graph" and kind == "b":
                junk = [torch.full((1 << 20,), float("nan"), device=DEV) for _ in range(8)]   # would land in freed blocks
                del junk
        losses[mode] = cur
        if mode == "graph":
            assert tr.ws.generation > 0 and tr.ws.retired_bytes() > 0      # the big batch did regrow the workspace
            assert len(tr._graphs) >= 2
    for a, b in zip(losses["eager"], losses["graph"]):
        assert abs(a - b) < 1e-5 * max(1.0, abs(a)), (losses["eager"], losses["graph"])


def test_step_ids_host_may_run_ahead_of_the_device():
    """step_ids with CUDA graphs lets the host enqueue many steps ahead; the pinned id staging must not be rewritten
    before its H2D copy ran.  Every graph of the arena has the same size, so all id lists share one shape slot."""
    graphs = [synth.make_batch(1, 24, seed=600 + i, vuln_rate=0.5) for i in range(40)]
    rng = np.random.default_rng(0)
    id_lists = [rng.integers(0, 40, 8) for _ in range(30)]
    finals = {}
    for mode in ("synced", "run_ahead"):
        torch.manual_seed(6)
        m = D.FlowGNNGGNNModule(FEAT, 1002, 32, 3, 2, concat_all_absdf=True, engine="tcgen05").to(DEV)
        tr = D.FusedTrainer(m, use_cuda_graph=True)
        arena = D.GraphArena.from_graphs(graphs, DEV)
        hist = torch.zeros(len(id_lists), device=DEV)
        for i, ids in enumerate(id_lists):
            hist[i:i + 1].copy_(tr.step_ids(arena, ids))       # device-side copy in stream order: no host sync
            if mode == "synced":
                torch.cuda.synchronize()
        torch.cuda.synchronize()
        finals[mode] = hist.cpu()
    # float atomics make two runs differ in the last bits; a torn or overwritten id list trains on other graphs (loss off by ~1e-1)
    assert (finals["synced"] - finals["run_ahead"]).abs().max() < 1e-3, (finals["synced"], finals["run_ahead"])
    assert finals["synced"].std() > 1e-2


def test_unequal_shards_need_and_use_the_global_batch():
    """Shards balanced by nodes hold different numbers of graphs: the loss / gradient scale is 1 / B_global, never 1 / B_local.
    A single-rank trainer given one shard + the global batch size reproduces that shard's share of the global mean loss."""
    g = synth.make_batch(sizes=[200, 10, 10, 10, 10, 10, 10, 10], seed=7, vuln_rate=0.5)
    parts = G.split_batch(g, 2)
    assert parts[0].batch_size != parts[1].batch_size
    torch.manual_seed(8)
    o = O.OracleFlowGNNGGNN(FEAT, 1002, 32, 3, 2, concat_all_absdf=True)
    total = 0.0
    for part in parts:
        m = D.FlowGNNGGNNModule(FEAT, 1002, 32, 3, 2, concat_all_absdf=True, engine="simt")
        m.load_state_dict(o.state_dict())
        m.to(DEV)
        tr = D.FusedTrainer(m, distributed=False)
        total += float(tr.step(part.to(DEV), global_batch=g.batch_size))
    loss_ref, _ = o.training_loss(g)
    assert abs(total - float(loss_ref)) < 1e-5
    tr.world = 2                                   # a multi-rank trainer refuses to guess
    with pytest.raises(D.DdfaError):
        tr.step(parts[0].to(DEV))
