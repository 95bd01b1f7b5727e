"""the DGL-free reader of the reference's processed files (schema of dbize.py / dbize_absdf.py output)."""
import numpy as np
import pandas as pd
import pytest
import torch

import deepdfa_b200 as D
from deepdfa_b200 import bigvul_io as IO

FEAT = "_ABS_DATAFLOW_datatype_all_limitall_1000_limitsubkeys_1000"
TAIL = "_all_limitall_1000_limitsubkeys_1000"


def _write_dataset(root, seed=0, n_graphs=7):
    """Files as dbize.py:104-105 / dbize_absdf.py write them: an unnamed index column first, rows grouped by graph."""
    rng = np.random.default_rng(seed)
    folder = root / "bigvul"
    folder.mkdir(parents=True)
    node_rows, edge_rows, truth = [], [], {}
    for gid in rng.permutation(np.arange(100, 100 + 3 * n_graphs, 3))[:n_graphs]:        # non-contiguous, unsorted graph ids
        n = int(rng.integers(2, 12))
        node_ids = rng.permutation(np.arange(1000, 1000 + 4 * n, 4))[:n]                 # Joern node ids: arbitrary
        # CFG-like edges in dgl ids; make sure the highest id appears (dgl.graph infers N = max id + 1)
        src = rng.integers(0, n, size=2 * n); dst = rng.integers(0, n, size=2 * n)
        src[0], dst[0] = n - 1, 0
        vul = rng.random(n) < 0.3
        feats = {k: rng.integers(0, 1002, n) for k in D.batched_graph.ABS_DATAFLOW_SUBKEYS}
        feats["main"] = feats["datatype"]        # FEAT is the datatype file itself (config_default.yaml: feat = ..._datatype_all_...)
        for i in range(n):
            node_rows.append(dict(graph_id=gid, node_id=int(node_ids[i]), dgl_id=i, vuln=int(vul[i]), code=f"x = {i};", _label="CALL"))
        for s_, d_ in zip(src, dst):
            edge_rows.append(dict(graph_id=gid, innode=int(s_), outnode=int(d_)))
        truth[int(gid)] = dict(n=n, src=src, dst=dst, vul=vul.astype(np.int32), feats=feats, node_ids=node_ids)
    nodes = pd.DataFrame(node_rows); edges = pd.DataFrame(edge_rows)
    nodes.to_csv(folder / "nodes.csv"); edges.to_csv(folder / "edges.csv")
    def feat_file(stem, key):
        rows = [dict(graph_id=g, node_id=int(t["node_ids"][i]), **{stem: int(t["feats"][key][i])}) for g, t in truth.items() for i in range(t["n"])]
        df = pd.DataFrame(rows).sample(frac=1.0, random_state=1)      # feature files need not be in node order: it is a merge
        df.to_csv(folder / f"nodes_feat_{stem}_fixed.csv")
    for sub in D.batched_graph.ABS_DATAFLOW_SUBKEYS:
        feat_file(f"_ABS_DATAFLOW_{sub}{TAIL}", sub)
    return truth


def test_reader_reproduces_dbize_graphs_and_graphmogrifier(tmp_path):
    truth = _write_dataset(tmp_path)
    graphs = IO.load_graphs(tmp_path, "bigvul", FEAT, concat_all_absdf=True)
    assert sorted(graphs) == sorted(truth)
    for gid, t in truth.items():
        g = graphs[gid]
        n = t["n"]
        assert g.num_nodes() == n and g.batch_size == 1
        src, dst = g.edges()
        # dbize_graphs.py:24-25: dgl.graph((innode, outnode)) then add_self_loop -> the file's edges, then one v->v per node
        assert src.tolist() == t["src"].tolist() + list(range(n)) and dst.tolist() == t["dst"].tolist() + list(range(n))
        assert g.ndata["_ABS_DATAFLOW"].dtype == torch.int64 and g.ndata["_ABS_DATAFLOW"].tolist() == t["feats"]["main"].tolist()
        for sub in D.batched_graph.ABS_DATAFLOW_SUBKEYS:
            assert g.ndata[f"_ABS_DATAFLOW_{sub}"].tolist() == t["feats"][sub].tolist()
        assert g.ndata["_VULN"].dtype == torch.int32 and g.ndata["_VULN"].tolist() == t["vul"].tolist()
    # collate in the reference's order and run the oracle label rule (base_module.py:83-95): max of _VULN per graph
    batch = D.batch([graphs[i] for i in sorted(graphs)])
    assert batch.batch_size == len(truth) and batch.num_nodes() == sum(t["n"] for t in truth.values())


def test_reader_rejects_inconsistent_files(tmp_path):
    _write_dataset(tmp_path)
    nodes = pd.read_csv(tmp_path / "bigvul" / "nodes.csv", index_col=0)
    nodes.iloc[:-1].to_csv(tmp_path / "bigvul" / "nodes.csv")          # one node row missing: DGL would refuse the ndata too
    with pytest.raises(ValueError, match="node rows"):
        IO.load_graphs(tmp_path, "bigvul", FEAT, concat_all_absdf=True)


@pytest.mark.parametrize("undersample,oversample", [("v1.0", None), ("v2.5", None), (0.25, None), (None, 2.0), ("v1.0", 1.5), (None, None)])
def test_epoch_indices_follow_dclass(undersample, oversample):
    rng0 = np.random.default_rng(3)
    df = pd.DataFrame({"id": np.arange(500) * 7, "vul": (rng0.random(500) < 0.12).astype(int)})

    def reference_rule(df, rng):               # restatement of dclass.py:84-105 with the reference's own pandas calls
        index = df.index
        if undersample is not None or oversample is not None:
            vul = df[df.vul == 1]; nonvul = df[df.vul == 0]
            if undersample is not None:
                if str(undersample).startswith("v"):
                    nonvul = nonvul.sample(int(len(vul) * float(str(undersample)[1:])), replace=False, random_state=rng)
                else:
                    nonvul = nonvul.sample(int(len(nonvul) * undersample), replace=False, random_state=rng)
            if oversample is not None:
                vul = vul.sample(int(len(vul) * oversample), replace=True, random_state=rng)
            index = pd.concat([vul, nonvul]).index
        return index

    a, b = np.random.RandomState(0), np.random.RandomState(0)
    for _epoch in range(3):                    # the RandomState persists across epochs
        assert IO.epoch_indices(df, undersample, oversample, a).tolist() == reference_rule(df, b).tolist()
    if undersample == "v1.0" and oversample is None:
        idx = IO.epoch_indices(df, undersample, None, np.random.RandomState(1))
        assert (df.loc[idx].vul == 1).sum() == (df.vul == 1).sum() == (df.loc[idx].vul == 0).sum()
