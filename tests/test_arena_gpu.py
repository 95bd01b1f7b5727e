"""Batch producer: device-side batch assembly from a resident graph arena vs the collate path."""
import numpy as np
import pytest
import torch

import deepdfa_b200 as D
from deepdfa_b200 import synth
from deepdfa_b200._lib import DdfaError
from deepdfa_b200.engine import prepare_graph

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
FEAT = "_ABS_DATAFLOW_datatype_all_limitall_1000_limitsubkeys_1000"


def _dataset():
    # ragged sizes incl. a 1-node graph; 41 graphs in 3 "files"
    parts = [synth.make_batch(sizes=[1, 7, 150, 33, 2, 64, 19], seed=1, vuln_rate=0.3),
             synth.make_batch(16, 40, seed=2, variable=True, vuln_rate=0.2),
             synth.make_batch(18, 25, seed=3, variable=True)]
    singles = [g for p in parts for g in D.unbatch(p)]
    return singles


def test_arena_batch_is_bit_identical_to_collate_plus_csr_build():
    singles = _dataset()
    arena = D.GraphArena.from_graphs(singles, DEV)
    assert arena.num_graphs == len(singles)
    rng = np.random.default_rng(0)
    for ids in ([0], [3, 3, 0], list(range(len(singles))), rng.permutation(len(singles))[:17].tolist(), rng.integers(0, len(singles), 300).tolist()):
        ab = arena.batch(ids)
        ab.check()
        ref = D.batch([singles[i] for i in ids]).to(DEV)
        rdg = prepare_graph(ref, DEV)
        adg = prepare_graph(ab, DEV)              # the attached CSR, no build
        assert (adg.num_nodes, adg.num_edges, adg.batch_size) == (rdg.num_nodes, rdg.num_edges, rdg.batch_size)
        for name in ("indptr", "indptr_t", "graph_ptr"):
            assert torch.equal(getattr(adg, name), getattr(rdg, name)), name
        E_ = rdg.num_edges
        assert torch.equal(adg.indices[:E_], rdg.indices[:E_]) and torch.equal(adg.indices_t[:E_], rdg.indices_t[:E_])
        for k, v in ref.ndata.items():
            assert torch.equal(ab.ndata[k].to(v.dtype), v), k
        assert torch.equal(ab.batch_num_nodes(), ref.batch_num_nodes())
        assert torch.equal(ab.batch_num_edges().cpu(), ref.batch_num_edges().cpu())
        # edges(): same multiset of (src, dst) as the collated batch
        s1, d1 = ab.edges(); s2, d2 = ref.edges()
        k1 = torch.sort(d1 * ref.num_nodes() + s1).values; k2 = torch.sort(d2.to(DEV) * ref.num_nodes() + s2.to(DEV)).values
        assert torch.equal(k1, k2)
    with pytest.raises(IndexError):
        arena.batch([len(singles)])
    with pytest.raises(ValueError):
        arena.batch([])


def test_module_and_trainer_on_arena_batches():
    singles = _dataset()
    arena = D.GraphArena.from_graphs(singles, DEV)
    torch.manual_seed(0)
    m = D.FlowGNNGGNNModule(FEAT, 1002, 32, 4, 2, concat_all_absdf=True, engine="tcgen05").to(DEV)
    ids = [5, 2, 9, 30, 11, 2]
    with torch.no_grad():
        out_a = m(arena.batch(ids), {})
        out_r = m(D.batch([singles[i] for i in ids]), {})
    assert torch.equal(out_a, out_r)
    assert torch.equal(m.get_label(arena.batch(ids)), m.get_label(D.batch([singles[i] for i in ids])))
    # training: id lists through the CUDA-graph path == eager steps on collated host batches
    id_lists = [list(range(i, i + 8)) for i in (0, 8, 16)] * 3
    losses = {}
    for mode in ("collate_eager", "arena_graph", "arena_eager"):
        torch.manual_seed(1)
        mm = D.FlowGNNGGNNModule(FEAT, 1002, 32, 4, 2, concat_all_absdf=True, engine="tcgen05").to(DEV)
        tr = D.FusedTrainer(mm, use_cuda_graph=(mode == "arena_graph"))
        cur = []
        for il in id_lists:
            if mode == "collate_eager":
                cur.append(float(tr.step(D.batch([singles[i] for i in il]))))
            else:
                cur.append(float(tr.step_ids(arena, il)))
        losses[mode] = cur
    for mode in ("arena_graph", "arena_eager"):
        for a, b in zip(losses["collate_eager"], losses[mode]):
            assert abs(a - b) < 1e-5 * max(1.0, abs(a)), (mode, losses)
