"""Fused data-parallel train step for ``FlowGNNGGNNModule`` on B200.

One process per GPU.  Per step: forward (T x {gather, GRU}) -> readout+MLP -> labels+BCE ->
hand-written backward -> ONE all-reduce of the flat gradient buffer (NCCL over NVLink/NVSwitch via
``torch.distributed``; the loss rides in the buffer's last element) -> fused Adam over the flat
parameter buffer.  Graphs never exchange messages, so the batch shards across ranks with no
data-path collective; the gradient all-reduce is the only exchange step.

Replaces, for the hot path only, Lightning's ``Trainer.fit`` loop around
``BaseModule.training_step`` (base_module.py:171-199) + ``torch.optim.Adam`` (config_default.yaml:43-47).
"""
from __future__ import annotations

from typing import Optional

import os

import torch
import torch.distributed as dist

from . import _lib
from . import engine as E
from .batched_graph import BatchedCFG, as_batched_cfg
from .module import FlowGNNGGNNModule, _ENGINES

_ALIGN = 64  # elements; keeps every parameter 256-byte aligned inside the flat buffers


class FusedTrainer:
    def __init__(self, module: FlowGNNGGNNModule, lr: float = 1e-3, betas=(0.9, 0.999), eps: float = 1e-8,
                 weight_decay: float = 1e-2, process_group=None, use_cuda_graph: bool = False, max_graph_shapes: int = 8,
                 max_resident_graphs: int = 64, distributed: bool = True, bucket_nodes: int = 0, bucket_edges: int = 0,
                 bucket_min_pad_nodes: int = 64, overlap_allreduce: bool = True, exchange: str = "auto"):
        """``distributed=False`` makes this a single-rank trainer even inside an initialised process group (no all-reduce).
        ``bucket_nodes`` / ``bucket_edges`` > 0 switch on shape bucketing for HOST batches under ``use_cuda_graph``: every batch
        is padded with ONE dummy graph of isolated nodes up to the next multiple of ``bucket_nodes`` nodes (at least
        ``bucket_min_pad_nodes`` of them, which also carry the padding edges as self loops) and ``bucket_edges`` edges, so that a
        shuffled stream of ever-new ``(N, E)`` (the reference reshuffles every epoch, datamodule.py:123-129) replays a handful of
        captured graphs.  The dummy graph has zero loss weight (``ddfa_graph_label_bce_valid``): no gradient comes from it."""
        if module.device.type != "cuda":
            raise _lib.DdfaError("FusedTrainer needs the module on a CUDA device (no CPU fallback)")
        if module.hparams.label_style != "graph" or module.hparams.encoder_mode:
            raise NotImplementedError("FusedTrainer fuses the shipped configuration (label_style='graph', a classifier head); train "
                                      "label_style='node' / encoder_mode modules through module.training_step + torch.optim")
        self.module = module
        self.device = module.device
        self.lr, self.betas, self.eps, self.weight_decay = lr, betas, eps, weight_decay
        self.pg = process_group
        self.world = dist.get_world_size(process_group) if (distributed and dist.is_available() and dist.is_initialized()) else 1
        self.bucket_nodes, self.bucket_edges, self.bucket_min_pad_nodes = int(bucket_nodes), int(bucket_edges), int(bucket_min_pad_nodes)
        # The gradient exchange is split in two: everything except the four GGNN weight matrices' gradients is final before the
        # batched weight-gradient GEMM starts, so those ranges are all-reduced on a side stream WHILE that launch runs; only
        # [w_msg, b_msg, w_ih, w_hh] (0.46 MB of the 1.5 MB) is reduced after it.  False: one all-reduce of the whole buffer.
        self.overlap_allreduce = bool(overlap_allreduce)
        self._ar_stream = None
        # exchange = "p2p": no NCCL call in the step — the flat parameter and gradient buffers live in symmetric (peer-mapped)
        # memory and ONE kernel per rank does reduce-scatter + Adam + all-gather over NVLink (ddfa_allreduce_adam_p2p); optimizer
        # moments are sharded (each rank keeps them for its 1/R slice only).  Single node.  "nccl": all-reduce + ddfa_adam_flat.
        # "auto" (default): "p2p" when every rank of the group is on this node and the symmetric-memory set-up succeeds on ALL
        # ranks, else "nccl" (the reason is kept in ``exchange_note``).  Measured at N = 2 / 4 / 8: profiles/r03k, r03n, r03o.
        if exchange not in ("auto", "nccl", "p2p"):
            raise ValueError(f"exchange must be 'auto', 'nccl' or 'p2p', got {exchange!r}")
        self.exchange_note = None
        auto = exchange == "auto"
        if auto:
            exchange = "p2p"
            local = int(os.environ.get("LOCAL_WORLD_SIZE", "0") or 0)
            if self.world > 1 and (dist.get_backend(process_group) != "nccl" or (local and local != self.world)
                                   or torch.cuda.device_count() < self.world):
                exchange, self.exchange_note = "nccl", "auto: ranks span more than this node (or a non-NCCL group): NCCL all-reduce"
        self.exchange = exchange if self.world > 1 else "nccl"
        self.use_cuda_graph = use_cuda_graph
        # a captured graph bakes in the batch SHAPE (and, for resident batches, the batch object): cap how many are kept so a
        # stream of ever-new shapes (un-bucketed real data) degrades to eager launches instead of growing without bound
        self.max_graph_shapes = max_graph_shapes
        self.max_resident_graphs = max_resident_graphs
        plist = module.param_list()
        offs, total = [], 0
        for p in plist:
            offs.append(total)
            total += (p.numel() + _ALIGN - 1) // _ALIGN * _ALIGN
        self.numel = total
        ntab = len(module._tables())
        self._gemm_grad_range = (offs[ntab], offs[ntab + 4])     # flat offsets of [w_msg, b_msg, w_ih, w_hh]
        with torch.cuda.device(self.device):
            if self.exchange == "p2p":
                try:
                    self._setup_p2p(total)
                    ok, why = 1, None
                except _lib.DdfaError as exc:
                    if not auto:
                        raise
                    ok, why = 0, str(exc)
                # all ranks take the same path: one failed set-up sends every rank to NCCL.  (The set-up itself contains collectives —
                # symmetric-memory rendezvous — so this covers failures every rank sees alike: the module missing, peer access
                # unavailable, an allocation refused; a rank that dies alone is a job failure either way.)
                if auto:
                    flag = torch.tensor([ok], dtype=torch.int32, device=self.device)
                    dist.all_reduce(flag, op=dist.ReduceOp.MIN, group=process_group)
                    if int(flag.item()) == 0:
                        self.exchange = "nccl"
                        self.exchange_note = "auto: symmetric-memory set-up failed on a rank (" + (why or "another rank") + "): NCCL all-reduce"
            if self.exchange != "p2p":
                self.flat_p = torch.zeros(total, dtype=torch.float32, device=self.device)
                self.flat_g = torch.zeros(total + _ALIGN, dtype=torch.float32, device=self.device)  # [+ loss slot]
            self.exp_avg = torch.zeros(total, dtype=torch.float32, device=self.device)
            self.exp_avg_sq = torch.zeros(total, dtype=torch.float32, device=self.device)
            self.step_count = torch.zeros(1, dtype=torch.int32, device=self.device)
        gviews = []
        for p, o in zip(plist, offs):
            view = self.flat_p[o:o + p.numel()].view_as(p)
            view.copy_(p.data)
            p.data = view                      # module parameters now alias the flat buffer
            gviews.append(self.flat_g[o:o + p.numel()].view_as(p))
        K, nl = len(module._tables()), module._num_layers
        self.params = E.ParamPack.from_flat_list([p.data for p in plist], K, nl)
        self.grads = E.ParamPack.from_flat_list(gviews, K, nl)
        self.loss_slot = self.flat_g[total:total + 1]          # this rank's share of the loss goes here (the kernels' loss_out)
        if self.exchange == "p2p":                                # ... and the global loss into a local word (peers read the slot above)
            self._loss_local = self.loss_slot
            self.loss_slot = torch.zeros(1, dtype=torch.float32, device=self.device)
        self.ws = E.Workspace(self.device)
        self._graphs = {}
        self._stream_slots = {}
        self._copy_stream = None
        self._warm_shapes = set()

    # ------------------------------------------------------------------------------------
    def _setup_p2p(self, total: int):
        """Symmetric allocations + rendezvous (torch.distributed._symmetric_memory): every rank gets device pointers to every
        rank's parameter / gradient / flag buffers."""
        try:
            import torch.distributed._symmetric_memory as symm
            group = self.pg if self.pg is not None else dist.group.WORLD
            self.flat_p = symm.empty(total, dtype=torch.float32, device=self.device)
            self.flat_g = symm.empty(total + _ALIGN, dtype=torch.float32, device=self.device)
            self._flags = symm.empty(64, dtype=torch.int32, device=self.device)
            self.flat_p.zero_(); self.flat_g.zero_(); self._flags.zero_()
            torch.cuda.synchronize(self.device)
            hp, hg, hf = (symm.rendezvous(t, group) for t in (self.flat_p, self.flat_g, self._flags))
            self._peer_ptrs = tuple([int(x) for x in h.buffer_ptrs] for h in (hp, hg, hf))
            self._symm_handles = (hp, hg, hf)
            self._p2p_rank = dist.get_rank(group)
        except Exception as exc:       # no silent downgrade to NCCL: the caller asked for the peer-memory exchange
            raise _lib.DdfaError(f"exchange='p2p': symmetric-memory setup failed ({type(exc).__name__}: {exc})") from exc
        if len(self._peer_ptrs[0]) != self.world or 2 * self.world > 64:
            raise _lib.DdfaError("exchange='p2p': unexpected symmetric-memory world size")
        self._ticket = torch.zeros(1, dtype=torch.int32, device=self.device)
        dist.barrier(group)            # every rank's flag words are zero before anyone's first kernel can write one

    def _global_batch(self, global_batch: Optional[int], local_graphs: int) -> int:
        """The divisor of the mean BCE (base_module.py:74,183).  Ranks generally hold different numbers of graphs
        (batched_graph.split_batch balances by nodes), so with more than one rank the caller must say what the global batch is."""
        if global_batch is not None:
            return int(global_batch)
        if self.world > 1:
            raise _lib.DdfaError("FusedTrainer.step: global_batch is required when world_size > 1 (shards hold different numbers of graphs; "
                                 "the loss is the mean over the GLOBAL batch)")
        return int(local_graphs)

    def _enqueue(self, g, dg, idx, vuln, global_batch: int, num_valid: Optional[int] = None):
        m = self.module
        eng = _ENGINES[m.engine]
        pw = 1.0 if m.hparams.positive_weight is None else float(m.hparams.positive_weight)
        self.flat_g.zero_()
        _, logits, saved = E.forward(self.params, dg, idx, m.hparams.n_steps, training=True, engine=eng, alloc=self.ws)
        _, _, dlogits = E.graph_label_bce(dg, vuln, logits, pw, 1.0 / global_batch, 1.0 / global_batch, True,
                                          alloc=self.ws, loss_out=self._loss_local if self.exchange == "p2p" else self.loss_slot,
                                          num_valid=num_valid)
        if self.exchange == "p2p":
            E.backward(self.params, dg, saved, self.grads, dlogits=dlogits, engine=eng, alloc=self.ws)
            pp, pg_, pf = self._peer_ptrs
            _lib.lib().call("ddfa_allreduce_adam_p2p", _lib.ptr_array(pp), _lib.ptr_array(pg_), _lib.ptr_array(pf), self._p2p_rank, self.world,
                            self.exp_avg.data_ptr(), self.exp_avg_sq.data_ptr(), self.step_count.data_ptr(), self.numel, self.numel,
                            self.loss_slot.data_ptr(), self._ticket.data_ptr(), self.lr, self.betas[0], self.betas[1], self.eps,
                            self.weight_decay, torch.cuda.current_stream().cuda_stream)
            return
        split = self.world > 1 and self.overlap_allreduce
        E.backward(self.params, dg, saved, self.grads, dlogits=dlogits, engine=eng, alloc=self.ws,
                   on_small_grads_ready=self._reduce_small_grads if split else None)
        if split:
            lo, hi = self._gemm_grad_range
            dist.all_reduce(self.flat_g[lo:hi], op=dist.ReduceOp.SUM, group=self.pg)
            torch.cuda.current_stream().wait_stream(self._ar_stream)
        elif self.world > 1:
            dist.all_reduce(self.flat_g, op=dist.ReduceOp.SUM, group=self.pg)
        L = _lib.lib()
        L.call("ddfa_adam_flat", self.flat_p.data_ptr(), self.flat_g.data_ptr(), self.exp_avg.data_ptr(),
               self.exp_avg_sq.data_ptr(), self.step_count.data_ptr(), self.numel, self.lr, self.betas[0], self.betas[1],
               self.eps, self.weight_decay, torch.cuda.current_stream().cuda_stream)

    def _reduce_small_grads(self):
        """All-reduce of the embedding / bias / readout / MLP gradients and the loss slot on a side stream (engine.backward
        calls this before the weight-gradient launch)."""
        lo, hi = self._gemm_grad_range
        main = torch.cuda.current_stream()
        if self._ar_stream is None:
            self._ar_stream = torch.cuda.Stream(device=self.device)
        self._ar_stream.wait_stream(main)
        with torch.cuda.stream(self._ar_stream):
            dist.all_reduce(self.flat_g[:lo], op=dist.ReduceOp.SUM, group=self.pg)
            dist.all_reduce(self.flat_g[hi:], op=dist.ReduceOp.SUM, group=self.pg)      # ... b_ih, b_hh, gate, MLP, [loss]

    # ------------------------------------------------------------------------------------
    # ---- host batches through per-shape static buffers + captured graphs -------------------------------------------------
    def _bucket_shape(self, N: int, Eg: int):
        """Padded (nodes, edges) of a batch under shape bucketing, or None when bucketing is off."""
        if self.bucket_nodes <= 0:
            return None
        bn, be = self.bucket_nodes, max(self.bucket_edges, 1)
        Nb = (N + max(self.bucket_min_pad_nodes, 1) + bn - 1) // bn * bn
        Eb = (Eg + be - 1) // be * be
        return Nb, Eb

    def num_bucket_shapes(self) -> int:
        return sum(1 for k in self._stream_slots if k[0] == "bucket")

    def _stream_slot(self, g, global_batch: Optional[int]):
        N, Eg, B = g.num_nodes(), g.num_edges(), g.batch_size
        gb = self._global_batch(global_batch, B)
        bucket = self._bucket_shape(N, Eg)
        key = ("bucket", bucket[0], bucket[1], B, gb) if bucket else ("exact", N, Eg, B, gb)
        slot = self._stream_slots.get(key)
        if slot is None:
            if len(self._stream_slots) >= self.max_graph_shapes:
                return None
            src, dst = g.edges()
            dev = self.device
            Ns, Es, Bs = (bucket[0], bucket[1], B + 1) if bucket else (N, Eg, B)

            def new_set():
                st = {"src": torch.empty(Es, dtype=src.dtype, device=dev), "dst": torch.empty(Es, dtype=dst.dtype, device=dev),
                      "bnn": torch.empty(Bs, dtype=torch.int64, device=dev),
                      "ndata": {k: torch.zeros((Ns,) + tuple(v.shape[1:]), dtype=v.dtype, device=dev) for k, v in g.ndata.items()},
                      "graph": None, "keep": None, "free": None, "ready": None}
                return st
            # two input-buffer sets: while the graph of one set runs, the next batch is copied into the other (prefetch)
            slot = {"sets": [new_set(), new_set()], "next": 0, "staged": None, "warm": False, "N": Ns, "gb": gb,
                    "valid": B if bucket else None,
                    "iota": torch.arange(Es, dtype=src.dtype, device=dev) if bucket else None}
            self._stream_slots[key] = slot
        return slot

    def _stage(self, slot, g, stream):
        """Copies the host batch ``g`` into the slot's next buffer set on ``stream``; returns the set index.  Under bucketing
        the tails are (re)written too: padding nodes get feature index 0 / _VULN 0, the padding edges become self loops spread
        round-robin over the padding nodes, and the dummy graph's node count goes into the last ``batch_num_nodes`` entry."""
        i = slot["next"]
        slot["next"] = 1 - i
        st = slot["sets"][i]
        with torch.cuda.stream(stream):
            if st["free"] is not None:
                stream.wait_event(st["free"])           # the graph that last read this set has finished
            src, dst = g.edges()
            N, Eg, B = g.num_nodes(), g.num_edges(), g.batch_size
            st["src"][:Eg].copy_(src, non_blocking=True)
            st["dst"][:Eg].copy_(dst, non_blocking=True)
            st["bnn"][:B].copy_(g.batch_num_nodes(), non_blocking=True)
            for k, v in g.ndata.items():
                st["ndata"][k][:N].copy_(v, non_blocking=True)
            if slot["valid"] is not None:
                Nb, Eb = slot["N"], st["src"].shape[0]
                pad_nodes = Nb - N
                st["bnn"][B:].fill_(pad_nodes)
                for k in st["ndata"]:
                    st["ndata"][k][N:].zero_()
                if Eb > Eg:
                    torch.remainder(slot["iota"][: Eb - Eg], pad_nodes, out=st["src"][Eg:])
                    st["src"][Eg:].add_(N)
                    st["dst"][Eg:].copy_(st["src"][Eg:])
            ev = torch.cuda.Event()
            ev.record(stream)
            st["ready"] = ev
        return i

    def prefetch(self, batch, global_batch: Optional[int] = None) -> None:
        """Starts the host->device copy of a (pinned) host batch on a side stream so that it overlaps the step that is running;
        the following ``step(batch)`` with the SAME batch object picks the staged copy up.  No-op without ``use_cuda_graph`` or
        for device batches."""
        if not self.use_cuda_graph:
            return
        g = as_batched_cfg(batch)
        if g.device.type != "cpu":
            return
        with torch.cuda.device(self.device):
            if self._copy_stream is None:
                self._copy_stream = torch.cuda.Stream(device=self.device)
            slot = self._stream_slot(g, global_batch)
            if slot is not None:
                slot["staged"] = (id(batch), self._stage(slot, g, self._copy_stream))

    def _step_streamed(self, batch, g, global_batch: Optional[int]) -> torch.Tensor:
        """Host batch + use_cuda_graph: the batch's arrays are copied into device buffers that are STATIC per shape
        (num_nodes, num_edges, batch_size — or per BUCKET shape with ``bucket_nodes`` / ``bucket_edges``) and one captured
        CUDA graph per buffer set covers the whole step including the device CSR build — a new batch of a known shape costs
        its H2D copies (overlappable: ``prefetch``) plus one graph launch.  The first visit of a shape runs eagerly
        (workspace growth), the next two capture."""
        m = self.module
        with torch.cuda.device(self.device):
            slot = self._stream_slot(g, global_batch)
            if slot is None:          # more shapes than max_graph_shapes: same kernels, launched eagerly
                return self._step_eager(batch, global_batch)
            N, gb = slot["N"], slot["gb"]
            main = torch.cuda.current_stream()
            staged = slot["staged"]
            slot["staged"] = None
            if staged is not None and staged[0] == id(batch):
                i = staged[1]
                main.wait_event(slot["sets"][i]["ready"])
            else:
                i = self._stage(slot, g, main)
            st = slot["sets"][i]

            def enqueue():
                gs = BatchedCFG(st["src"], st["dst"], st["bnn"], dict(st["ndata"]), num_nodes=N)   # no cached device CSR
                g_, dg, idx = m._prepare(gs)
                vuln = gs.ndata["_VULN"]
                if vuln.dtype != torch.int32:
                    vuln = vuln.to(torch.int32)
                self._enqueue(g_, dg, idx, vuln.contiguous(), gb, num_valid=slot["valid"])
                return (gs, dg, idx, vuln)

            if not slot["warm"]:
                st["keep"] = enqueue()
                slot["warm"] = True
            else:
                if st["graph"] is None:
                    torch.cuda.synchronize(self.device)
                    cg = torch.cuda.CUDAGraph()
                    with torch.cuda.graph(cg):
                        st["keep"] = enqueue()            # tensors allocated during capture live in the graph's pool
                    st["graph"] = cg
                st["graph"].replay()
            ev = torch.cuda.Event()
            ev.record(main)
            st["free"] = ev
        return self.loss_slot

    def step_ids(self, arena, ids, global_batch: Optional[int] = None) -> torch.Tensor:
        """One optimisation step on the graphs ``ids`` of a device-resident :class:`deepdfa_b200.arena.GraphArena` (SURVEY.md §8
        f1: the batch producer).  With ``use_cuda_graph`` the batch is assembled into static per-shape buffers by
        ``ddfa_arena_batch`` inside one captured graph, so a step costs the H2D copy of the id list plus one graph launch;
        otherwise it is ``step(arena.batch(ids))``."""
        if not self.use_cuda_graph:
            return self.step(arena.batch(ids), global_batch)
        import numpy as np
        m = self.module
        ids_np = np.asarray(ids.cpu() if isinstance(ids, torch.Tensor) else ids, dtype=np.int64).reshape(-1)
        if ids_np.size == 0 or ids_np.min() < 0 or ids_np.max() >= arena.num_graphs:
            raise IndexError("step_ids: empty id list or graph id out of range")
        B = int(ids_np.shape[0])
        N = int(arena.nodes_per_graph[ids_np].sum())
        Eg = int(arena.edges_per_graph[ids_np].sum())
        gb = self._global_batch(global_batch, B)
        key = ("arena", id(arena), N, Eg, B, gb)
        slot = self._stream_slots.get(key)
        with torch.cuda.device(self.device):
            if slot is None:
                if len(self._stream_slots) >= self.max_graph_shapes:
                    return self._step_eager(arena.batch(ids), global_batch)
                # a ring of pinned id stages: the host may run several steps ahead of the device (that is what the captured
                # graph is for), so a stage is rewritten only after the H2D copy that last read it has completed
                slot = {"out": arena.alloc_outputs(B, N, Eg), "stages": [torch.empty(B, dtype=torch.int32).pin_memory() for _ in range(4)],
                        "stage_done": [None] * 4, "turn": 0, "steps": 0,
                        "graph": None, "warm": False, "keep": None, "arena": arena}     # the arena stays alive with its graph
                self._stream_slots[key] = slot
            k = slot["turn"]
            slot["turn"] = (k + 1) % len(slot["stages"])
            if slot["stage_done"][k] is not None:
                slot["stage_done"][k].synchronize()
            slot["stages"][k].copy_(torch.from_numpy(ids_np.astype(np.int32)))
            slot["out"]["ids"].copy_(slot["stages"][k], non_blocking=True)
            ev = torch.cuda.Event()
            ev.record()
            slot["stage_done"][k] = ev
            slot["steps"] += 1
            if slot["keep"] is not None and slot["steps"] % 256 == 0:
                slot["keep"][0].check()      # the assembler's device error counter (bad id / totals mismatch): one sync every 256 steps

            def enqueue():
                g = arena._assemble(slot["out"]["ids"], B, N, Eg, slot["out"])
                g_, dg, idx = m._prepare(g)
                self._enqueue(g_, dg, idx, g.ndata["_VULN"], gb)
                return (g, dg, idx)

            if not slot["warm"]:
                slot["keep"] = enqueue()
                slot["warm"] = True
            else:
                if slot["graph"] is None:
                    torch.cuda.synchronize(self.device)
                    cg = torch.cuda.CUDAGraph()
                    with torch.cuda.graph(cg):
                        slot["keep"] = enqueue()
                    slot["graph"] = cg
                slot["graph"].replay()
        return self.loss_slot

    def step(self, batch, global_batch: Optional[int] = None) -> torch.Tensor:
        """One optimisation step on this rank's shard.  Returns the device tensor holding the
        global mean loss (valid after the step's stream work completes)."""
        m = self.module
        if self.use_cuda_graph:
            gb_ = as_batched_cfg(batch)
            if gb_.device.type == "cpu":
                return self._step_streamed(batch, gb_, global_batch)
        return self._step_eager(batch, global_batch)

    def _step_eager(self, batch, global_batch: Optional[int] = None) -> torch.Tensor:
        """Device-resident batch objects (one captured graph per object when ``use_cuda_graph``), or plain eager launches."""
        m = self.module
        g, dg, idx = m._prepare(batch)
        vuln = g.ndata["_VULN"]
        if vuln.device != self.device or vuln.dtype != torch.int32:
            key = "vuln_dev"
            cached = g._cache.get(key)
            if cached is None:
                cached = vuln.to(self.device, non_blocking=True).to(torch.int32).contiguous()
                g._cache[key] = cached
            vuln = cached
        global_batch = self._global_batch(global_batch, dg.batch_size)
        with torch.cuda.device(self.device):
            shape_key = (dg.num_nodes, dg.num_edges, dg.batch_size)
            capturable = self.use_cuda_graph and as_batched_cfg(batch).device.type == "cuda" and \
                (id(g) in self._graphs or len(self._graphs) < self.max_resident_graphs)
            if not capturable or shape_key not in self._warm_shapes:
                # eager step; also the warm-up (workspace growth, lazy CUDA module init) before any capture
                self._enqueue(g, dg, idx, vuln, global_batch)
                self._warm_shapes.add(shape_key)
            else:
                # one captured CUDA graph per resident batch object (its device pointers are baked in)
                entry = self._graphs.get(id(g))
                if entry is None:
                    torch.cuda.synchronize(self.device)
                    cg = torch.cuda.CUDAGraph()
                    with torch.cuda.graph(cg):
                        self._enqueue(g, dg, idx, vuln, global_batch)
                    entry = (cg, g, idx, vuln)     # keep the captured tensors alive
                    self._graphs[id(g)] = entry
                entry[0].replay()
        return self.loss_slot

    # ------------------------------------------------------------------------------------
    @staticmethod
    def dp_self_check(engine: str, device, rank: int, world: int, steps: int = 5, graphs_per_rank: int = 24, nodes: int = 60,
                      exchange: str = "auto") -> dict:
        """On-hardware data-parallel parity (SURVEY.md §8(e) "Determinism"): ``steps`` optimisation steps of a global batch
        sharded over the ``world`` ranks (node-balanced shards of different sizes, NCCL all-reduce) against the same steps of
        the UNSHARDED batch on this rank alone, same seeds.  fp32 summation order is the only difference.  Collective: every
        rank must call it.  Returns the loss curves and the largest parameter difference after the last step."""
        from . import synth
        from .batched_graph import split_batch
        feat = "_ABS_DATAFLOW_api_all_limitall_1000_limitsubkeys_1000"

        def make(distributed):
            torch.manual_seed(4321)
            m = FlowGNNGGNNModule(feat, 1002, 32, 8, 2, concat_all_absdf=True, positive_weight=4.0, engine=engine).to(device)
            return m, FusedTrainer(m, distributed=distributed, exchange=exchange if distributed else "nccl")
        m_dp, tr_dp = make(True)
        m_1, tr_1 = make(False)
        l_dp, l_1 = [], []
        for i in range(steps):
            b = synth.make_batch(graphs_per_rank * world, nodes, seed=900 + i, variable=True, vuln_rate=0.3)
            shard = split_batch(b, world)[rank]
            l_dp.append(float(tr_dp.step(shard.to(device), global_batch=b.batch_size)))
            l_1.append(float(tr_1.step(b.to(device), global_batch=b.batch_size)))
        dparam = max(float((p.data - q.data).abs().max()) for p, q in zip(m_dp.param_list(), m_1.param_list()))
        return {"steps": steps, "world": world, "global_batch": graphs_per_rank * world, "loss_sharded": l_dp, "loss_single_rank": l_1,
                "max_abs_loss_diff": max(abs(a - b) for a, b in zip(l_dp, l_1)), "max_abs_param_diff": dparam,
                "shard_sizes_differ": True, "exchange": exchange}
