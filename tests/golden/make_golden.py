"""Generates tests/golden/ggnn_golden.pt with the ORACLE (oracle/ggnn_oracle.py) on the CPU.

The reference ships no golden vectors for this path and cannot run here (no dgl / Lightning), so
these fixtures are produced by the restatement; the torch-delegated parts (Embedding, Linear,
GRUCell, BCEWithLogitsLoss, Adam) are the reference's own dependencies executing.  Run from the
repo root:  python tests/golden/make_golden.py
Each case stores the inputs, the init seed + per-tensor checksums of the state_dict (to detect RNG
drift), logits in fp32 and fp64, the loss, per-parameter gradient L2 norms (fp64) and, for the
'tiny' case, the full state_dict and the parameters after 3 Adam steps.
"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from deepdfa_b200 import synth  # noqa: E402  (data generator only)
from oracle.ggnn_oracle import OracleFlowGNNGGNN, make_optimizer  # noqa: E402

FEAT = "_ABS_DATAFLOW_api_all_limitall_1000_limitsubkeys_1000"


def graph_blob(g):
    src, dst = g.edges()
    return {"src": src.clone(), "dst": dst.clone(), "batch_num_nodes": g.batch_num_nodes().clone(),
            "ndata": {k: v.clone() for k, v in g.ndata.items()}}


def checksums(sd):
    return {k: (float(v.double().sum()), float(v.double().abs().sum())) for k, v in sd.items()}


def run_case(name, g, *, input_dim, hidden_dim, n_steps, layers, concat, seed, encoder_mode=False, pos_weight=None,
             keep_state=False, adam_steps=0):
    torch.manual_seed(seed)
    m32 = OracleFlowGNNGGNN(FEAT, input_dim, hidden_dim, n_steps, layers, concat_all_absdf=concat,
                            encoder_mode=encoder_mode, positive_weight=pos_weight)
    sd = {k: v.clone() for k, v in m32.state_dict().items()}
    m64 = OracleFlowGNNGGNN(FEAT, input_dim, hidden_dim, n_steps, layers, concat_all_absdf=concat,
                            encoder_mode=encoder_mode, positive_weight=pos_weight).double()
    m64.load_state_dict({k: v.double() for k, v in sd.items()})
    case = {"name": name, "graph": graph_blob(g), "seed": seed, "checksums": checksums(sd),
            "ctor": dict(feat=FEAT, input_dim=input_dim, hidden_dim=hidden_dim, n_steps=n_steps,
                         num_output_layers=layers, concat_all_absdf=concat, encoder_mode=encoder_mode,
                         positive_weight=pos_weight)}
    with torch.no_grad():
        case["out_fp32"] = m32(g).clone()
        case["out_fp64"] = m64(g).clone()
    case["labels"] = m32.get_label(g).clone()
    if not encoder_mode:
        loss, _ = m64.training_loss(g)
        loss.backward()
        case["loss_fp64"] = float(loss)
        case["grad_norm_fp64"] = {k: float(p.grad.norm()) for k, p in m64.named_parameters()}
        case["grads_fp64"] = {k: p.grad.clone() for k, p in m64.named_parameters()} if keep_state else None
    if keep_state:
        case["state_dict"] = sd
    if adam_steps:
        opt = make_optimizer(m32)
        losses = []
        for _ in range(adam_steps):
            opt.zero_grad()
            loss, _ = m32.training_loss(g)
            loss.backward()
            opt.step()
            losses.append(float(loss))
        case["adam_losses"] = losses
        case["state_after_adam"] = {k: v.clone() for k, v in m32.state_dict().items()}
    return case


def main():
    cases = []
    # tiny, fully self-contained (state_dict stored): H=8 x4 -> D=32, V=20
    g = synth.make_batch(sizes=[1, 2, 17, 9, 33], input_dim=20, seed=11)
    cases.append(run_case("tiny_T3_L2", g, input_dim=20, hidden_dim=8, n_steps=3, layers=2, concat=True, seed=3,
                          pos_weight=2.5, keep_state=True, adam_steps=3))
    g = synth.make_batch(sizes=[5, 12, 7], input_dim=20, seed=12)
    cases.append(run_case("tiny_single_T4_L1", g, input_dim=20, hidden_dim=16, n_steps=4, layers=1, concat=False, seed=4,
                          keep_state=True))
    cases.append(run_case("tiny_encoder_T2", g, input_dim=20, hidden_dim=8, n_steps=2, layers=3, concat=True, seed=5,
                          encoder_mode=True, keep_state=True))
    # reference-shaped (D=128, V=1002): seeds x {T5 L3 (shipped config), T8 L2, T8 L3}
    for seed in (0, 1, 2):
        g = synth.make_batch(num_graphs=12, nodes_per_graph=60, variable=True, seed=100 + seed)
        for (T, L) in ((5, 3), (8, 2), (8, 3)):
            cases.append(run_case(f"ref_s{seed}_T{T}_L{L}", g, input_dim=1002, hidden_dim=32, n_steps=T, layers=L,
                                  concat=True, seed=seed))
    g = synth.make_edge_cases()
    cases.append(run_case("edge_cases_T5_L3", g, input_dim=1002, hidden_dim=32, n_steps=5, layers=3, concat=True, seed=9))
    out = os.path.join(ROOT, "tests", "golden", "ggnn_golden.pt")
    torch.save({"torch_version": torch.__version__, "cases": cases}, out)
    print(out, os.path.getsize(out), "bytes,", len(cases), "cases")


if __name__ == "__main__":
    main()
