"""DGL-free reader of the reference's processed Big-Vul files -> graphs -> :class:`GraphArena`.

What the reference does, and where:

* ``DDFA/sastvd/scripts/dbize.py:40-60,104-105`` writes ``nodes.csv`` (one row per CFG node: ``graph_id, node_id, dgl_id, vuln,
  code, _label``; ``dgl_id`` = row number inside the graph) and ``edges.csv`` (``graph_id, innode, outnode`` in dgl ids).
* ``DDFA/sastvd/scripts/dbize_graphs.py:17-33`` turns every graph's edge rows into ``dgl.graph((innode, outnode))`` — so the
  message source is ``innode`` and the destination ``outnode``, and the node count is ``max id + 1`` — then
  ``dgl.add_self_loop`` (one ``v -> v`` edge per node, appended), and saves ``graphs.bin``.
* ``DDFA/sastvd/scripts/dbize_absdf.py`` writes ``nodes_feat_<feat>_fixed.csv`` (``graph_id, node_id, <feat>``): the
  abstract-dataflow vocabulary index of every node (0 = not a definition, 1 = unknown, 2.. = known).
* ``DDFA/sastvd/linevd/graphmogrifier.py:20-40`` left-merges the feature file(s) onto ``nodes.csv`` by ``(graph_id, node_id)``
  (``concat_all_absdf``: the four ``_ABS_DATAFLOW_{api,datatype,literal,operator}...`` files, renamed to
  ``_ABS_DATAFLOW_<subkey>``); ``:59-95`` attaches, per graph in ``groupby("graph_id")`` order and in FILE ROW ORDER inside
  the graph, ``ndata["_ABS_DATAFLOW"]``, the four subkey vectors and ``ndata["_VULN"]``; graphs without node rows are dropped.
* ``DDFA/sastvd/helpers/dclass.py:84-105`` draws the per-epoch index set (``undersample="v1.0"``: all vulnerable examples plus
  as many non-vulnerable ones, sampled without replacement from a persistent ``RandomState``).

This module restates exactly that with pandas (the ``.bin`` container itself is DGL's and is not read: it holds nothing that
``edges.csv`` does not).  The dataset is not shipped with the reference, so the tests build small files in the same schema —
and the reference's OWN ``get_nodes_df`` / ``get_graphs`` / ``get_epoch_indices`` were run on those files in the build container
(``tests/golden/make_reference_io_golden.py``); ``tests/test_bigvul_io.py::test_reader_matches_reference_loaders`` holds the
reader to their output.  Only the three DGL calls of ``dbize_graphs.py`` stay a restatement.
"""
from __future__ import annotations

from pathlib import Path
from typing import Dict, Optional, Tuple

import numpy as np
import pandas as pd
import torch

from .batched_graph import ABS_DATAFLOW_SUBKEYS, BatchedCFG, add_self_loop, graph


def _sample_text(sample_mode: bool) -> str:
    return "_sample" if sample_mode else ""


def read_edge_graphs(edges_csv) -> Dict[int, BatchedCFG]:
    """``dbize_graphs.py:17-27``: one graph per ``graph_id`` (in ``groupby`` = ascending id order), self-loops appended."""
    df = pd.read_csv(edges_csv, index_col=0, usecols=["Unnamed: 0", "graph_id", "innode", "outnode"])
    out: Dict[int, BatchedCFG] = {}
    for graph_id, group in df.groupby("graph_id"):
        g = graph((group["innode"].tolist(), group["outnode"].tolist()))     # src = innode, dst = outnode, N = max id + 1
        out[int(graph_id)] = add_self_loop(g)
    return out


def read_nodes(processed_dir, dsname: str = "bigvul", feat: Optional[str] = None, concat_all_absdf: bool = False,
               sample_mode: bool = False, load_features: bool = True) -> pd.DataFrame:
    """Node table with the requested feature columns merged in — the semantics of ``graphmogrifier.get_nodes_df`` (:20-40):
    every feature file is LEFT-joined on ``(graph_id, node_id)``, so node order is that of ``nodes.csv``."""
    folder = Path(processed_dir) / dsname
    tag = _sample_text(sample_mode)
    table = pd.read_csv(folder / f"nodes{tag}.csv", index_col=0, dtype={"code": str}, na_values=[],
                        usecols=["Unnamed: 0", "graph_id", "node_id", "dgl_id", "vuln", "code", "_label"]).reset_index(drop=True)
    table["code"] = table["code"].astype(str)
    if not load_features:
        return table
    joins = []                                           # (file stem, column rename or None)
    if feat is not None:
        joins.append((feat, None))
    if concat_all_absdf:
        tail = feat[feat.index("_all"):]                 # e.g. "_all_limitall_1000_limitsubkeys_1000"
        joins += [(f"_ABS_DATAFLOW_{sub}{tail}", f"_ABS_DATAFLOW_{sub}") for sub in ABS_DATAFLOW_SUBKEYS]
    for stem, new_name in joins:
        extra = pd.read_csv(folder / f"nodes_feat_{stem}_fixed{tag}.csv", index_col=0)
        if new_name is not None:
            first = [c for c in extra.columns if c.startswith("_ABS_DATAFLOW")][0]
            extra = extra.rename(columns={first: new_name})
        table = table.merge(extra, how="left", on=["graph_id", "node_id"])
    return table


def attach_node_data(graphs_by_id: Dict[int, BatchedCFG], nodes: pd.DataFrame, feat: str, concat_all_absdf: bool = False
                     ) -> Dict[int, BatchedCFG]:
    """``graphmogrifier.get_graphs`` (:59-95): ndata per graph in file row order; graphs without node rows are dropped."""
    out: Dict[int, BatchedCFG] = {}
    for graph_id, group in nodes.groupby("graph_id"):
        g = graphs_by_id[int(graph_id)]
        if len(group) != g.num_nodes():
            # DGL raises here too (ndata length must equal the node count)
            raise ValueError(f"graph {graph_id}: {len(group)} node rows but the edge list implies {g.num_nodes()} nodes")
        ndata = {"_ABS_DATAFLOW": torch.LongTensor(group[feat].tolist())}
        if concat_all_absdf:
            for other in ABS_DATAFLOW_SUBKEYS:
                ndata[f"_ABS_DATAFLOW_{other}"] = torch.LongTensor(group[f"_ABS_DATAFLOW_{other}"].tolist())
        ndata["_VULN"] = torch.Tensor(group["vuln"].tolist()).int()
        src, dst = g.edges()
        out[int(graph_id)] = BatchedCFG(src, dst, g.batch_num_nodes(), ndata, g.batch_num_edges(), num_nodes=g.num_nodes())
    return out


def load_graphs(processed_dir, dsname: str = "bigvul", feat: Optional[str] = None, concat_all_absdf: bool = False,
                sample_mode: bool = False) -> Dict[int, BatchedCFG]:
    """edges.csv + nodes.csv + feature files -> {graph_id: single graph with ndata}, as the reference's dataset holds them."""
    base = Path(processed_dir) / dsname
    graphs = read_edge_graphs(base / f"edges{_sample_text(sample_mode)}.csv")
    nodes = read_nodes(processed_dir, dsname, feat, concat_all_absdf, sample_mode)
    return attach_node_data(graphs, nodes, feat, concat_all_absdf)


def load_arena(processed_dir, device="cuda", **kw) -> Tuple["object", np.ndarray]:
    """The whole processed dataset as a device-resident :class:`deepdfa_b200.GraphArena`; returns (arena, graph_ids) with
    ``graph_ids[i]`` = the reference's id of arena graph ``i`` (ascending)."""
    from .arena import GraphArena
    graphs = load_graphs(processed_dir, **kw)
    ids = np.array(sorted(graphs), dtype=np.int64)
    return GraphArena.from_graphs([graphs[int(i)] for i in ids], device), ids


def epoch_indices(df: pd.DataFrame, undersample=None, oversample=None, rng: Optional[np.random.RandomState] = None) -> pd.Index:
    """Index set of one epoch — ``BigVulDataset.get_epoch_indices`` (dclass.py:84-105).  ``df``: one row per example with a
    0/1 column ``vul``; ``rng``: the dataset's persistent ``np.random.RandomState(seed)`` (successive epochs draw different
    subsets).  ``undersample="v<f>"`` keeps ``int(f * #vulnerable)`` non-vulnerable rows, a float keeps that fraction of them;
    ``oversample=<f>`` redraws ``int(f * #vulnerable)`` vulnerable rows with replacement.  Draw order (non-vulnerable first,
    then vulnerable) and the result order (vulnerable rows, then non-vulnerable) follow the reference, so the same seed gives
    the same epochs."""
    if undersample is None and oversample is None:
        return df.index
    positives, negatives = df[df.vul == 1], df[df.vul == 0]
    if undersample is not None:
        spec = str(undersample)
        keep = int(len(positives) * float(spec[1:])) if spec.startswith("v") else int(len(negatives) * undersample)
        negatives = negatives.sample(keep, replace=False, random_state=rng)
    if oversample is not None:
        positives = positives.sample(int(len(positives) * oversample), replace=True, random_state=rng)
    return positives.index.append(negatives.index)
