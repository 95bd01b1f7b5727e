"""Runs single tcgen05-engine kernels on a C0-sized batch, for `ncu -k regex:...` captures.
   python scripts/kernel_only.py fwd|bwd [graphs]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from deepdfa_b200 import synth
from deepdfa_b200._lib import ENGINE_TCGEN05, lib
from deepdfa_b200.engine import _p, _stream_ptr, prepare_graph

DEV, D = "cuda:0", 128
which = sys.argv[1] if len(sys.argv) > 1 else "fwd"
graphs = int(sys.argv[2]) if len(sys.argv) > 2 else 256
L = lib()
g = synth.make_batch(graphs, 150, seed=1)
dg = prepare_graph(g, DEV)
N = g.num_nodes()
torch.manual_seed(0)
k = 1.0 / D ** 0.5
wf = (torch.rand(3 * D, D, device=DEV) * 2 - 1) * k
whh = (torch.rand(3 * D, D, device=DEV) * 2 - 1) * k
bf, bih, bhh = [(torch.rand(3 * D, device=DEV) * 2 - 1) * k for _ in range(3)]
h = torch.tanh(torch.randn(N, D, device=DEV))
ib = L.call("ddfa_act_image_bytes", N)
s_img = torch.zeros(ib, dtype=torch.uint8, device=DEV); h_img = torch.zeros(ib, dtype=torch.uint8, device=DEV)
o_img = torch.zeros(ib, dtype=torch.uint8, device=DEV)
out = torch.empty(N, D, device=DEV); gates = torch.rand(4, N, D, device=DEV)
st = _stream_ptr()
L.call("ddfa_act_to_image", _p(h), N, D, _p(h_img), st)
L.call("ddfa_gather_sum_image", _p(dg.indptr), _p(dg.indices), _p(h), N, D, _p(s_img), None, st)
if which == "fwd":
    wsb = L.call("ddfa_gru_step_workspace_bytes", 0, D, ENGINE_TCGEN05)
    ws = torch.zeros(wsb, dtype=torch.uint8, device=DEV)
    L.call("ddfa_gru_step_prepare", _p(wf), _p(bf), _p(bih), _p(whh), _p(bhh), D, ENGINE_TCGEN05, _p(ws), wsb, st)
    for i in range(6):
        train = i % 2 == 1
        L.call("ddfa_gru_step_fwd_image", _p(s_img), _p(h_img), _p(h), _p(dg.indptr), N, D, _p(out), _p(o_img) if train else None,
               _p(gates) if train else None, _p(ws), wsb, st)
else:
    wsb = L.call("ddfa_gru_step_bwd_workspace_bytes", N, D, ENGINE_TCGEN05)
    ws = torch.zeros(wsb, dtype=torch.uint8, device=DEV)
    L.call("ddfa_gru_step_prepare_bwd", _p(wf), _p(whh), D, ENGINE_TCGEN05, _p(ws), wsb, st)
    dh_o = torch.randn(N, D, device=DEV); ds = torch.empty(N, D, device=DEV); dh = torch.empty(N, D, device=DEV)
    acc = [torch.zeros(3 * D, D, device=DEV), torch.zeros(3 * D, device=DEV), torch.zeros(3 * D, device=DEV),
           torch.zeros(3 * D, D, device=DEV), torch.zeros(3 * D, device=DEV)]
    for i in range(4):
        L.call("ddfa_gru_step_bwd_image", _p(dh_o), _p(h), _p(h_img), _p(s_img), _p(gates), _p(dg.indptr), N, D, _p(ds), _p(dh), _p(acc[0]), _p(acc[1]),
               _p(acc[2]), _p(acc[3]), _p(acc[4]), _p(ws), wsb, 1 if i == 0 else 2, st)
torch.cuda.synchronize()
print("done", which, N)
