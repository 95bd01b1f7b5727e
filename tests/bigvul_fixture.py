"""Synthetic processed-dataset files in the reference's schema (dbize.py:104-105, dbize_absdf.py) — shared by tests/test_bigvul_io.py and
tests/golden/make_reference_io_golden.py so that the fixture and the test see the same files."""
import numpy as np
import pandas as pd

import deepdfa_b200 as D

FEAT = "_ABS_DATAFLOW_datatype_all_limitall_1000_limitsubkeys_1000"
TAIL = "_all_limitall_1000_limitsubkeys_1000"


def write_dataset(root, seed=0, n_graphs=7):
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


