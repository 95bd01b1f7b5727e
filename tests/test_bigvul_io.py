"""the DGL-free reader of the reference's processed files (schema of dbize.py / dbize_absdf.py output)."""
import numpy as np
import pandas as pd
import pytest
import torch

import deepdfa_b200 as D
from deepdfa_b200 import bigvul_io as IO

from bigvul_fixture import FEAT, write_dataset as _write_dataset


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


def test_reader_matches_reference_loaders(tmp_path):
    """tests/golden/reference_io_golden.pt holds what the reference's OWN graphmogrifier.get_nodes_df / get_graphs and
    BigVulDataset.get_epoch_indices produced on these files (make_reference_io_golden.py); the reader must reproduce it."""
    import os
    golden = torch.load(os.path.join(os.path.dirname(__file__), "golden", "reference_io_golden.pt"), weights_only=False)
    _write_dataset(tmp_path, seed=0, n_graphs=7)
    nodes = IO.read_nodes(tmp_path, "bigvul", FEAT, concat_all_absdf=True)
    assert list(nodes.columns) == golden["nodes_columns"]
    for col, ref in golden["nodes_records"].items():
        assert nodes[col].tolist() == ref, col
    graphs = IO.load_graphs(tmp_path, "bigvul", FEAT, concat_all_absdf=True)
    assert list(graphs) == golden["graph_ids"]                      # groupby order = ascending graph id
    for gid in golden["graph_ids"]:
        assert graphs[gid].num_nodes() == golden["num_nodes"][gid]
        assert sorted(graphs[gid].ndata) == sorted(golden["ndata"][gid])
        for name, ref in golden["ndata"][gid].items():
            got = graphs[gid].ndata[name]
            assert got.dtype == ref.dtype and torch.equal(got, ref), (gid, name)
    rng0 = np.random.default_rng(3)
    df = pd.DataFrame({"id": np.arange(500) * 7, "vul": (rng0.random(500) < 0.12).astype(int)})
    for key, ref_epochs in golden["epoch_indices"].items():
        us, os_ = key.split("|")
        us = None if us == "None" else (us if us.startswith("v") else float(us))
        os_ = None if os_ == "None" else float(os_)
        rng = np.random.RandomState(0)
        for ref in ref_epochs:
            assert IO.epoch_indices(df, us, os_, rng).tolist() == ref, key
