#!/usr/bin/env python
"""bench.py — DDFA GGNN hot path: CFG graphs/sec of a full train step on B200.

Workload (BASELINE.json configs[1], "C0"): synthetic Big-Vul-shaped batches of 256 CFGs x 150 nodes /
300 edges (incl. self loops), 4 x Embedding(1002,32) -> 128-d, T=8 propagation steps, attention
readout, 2-layer MLP head, BCE loss, backward, gradient all-reduce (N>1), Adam (coupled L2).
Per-GPU batch is fixed as N grows (weak scaling; the batch shards by graphs, no data-path collective).

  python bench.py --gpus 1 --steps K --warmup W            # our arm (CUDA, libddfa_b200.so)
  python bench.py --impl reference --steps K --warmup W    # reference arm: the reference path's CPU
                                                           # restatement (oracle/) on the host cores
One JSON line on stdout (rank 0).  See DESIGN.md §Measurement for every field.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

FEAT = "_ABS_DATAFLOW_api_all_limitall_1000_limitsubkeys_1000"
CFG = dict(graphs=256, nodes=150, edges_per_node=2.0, input_dim=1002, hidden_dim=32, n_steps=8, layers=2)
METRIC = "CFG graphs/sec (train step)"
UNIT = "graphs/s"
NUM_BATCHES = 8  # distinct resident batches rotated through the timed region


def workload_config(args, world):
    return {
        "workload": f"C0: {args.graphs} CFGs/GPU x {CFG['nodes']} nodes / {int(CFG['nodes'] * CFG['edges_per_node'])} edges, "
                    f"4xEmb(1002,32)->128-d, T={CFG['n_steps']}, attention readout, {CFG['layers']}-layer MLP, "
                    "BCE, backward, Adam (coupled L2)",
        "global_batch": args.graphs * world,
        "per_gpu_batch": args.graphs,
        "n_steps": CFG["n_steps"], "hidden": 128, "mlp_layers": CFG["layers"],
        "parallelism": f"dp{world}",
    }


# ------------------------------------------------------------------------------------------------
# clocks sampler (B200_PROFILING.md recipe)
# ------------------------------------------------------------------------------------------------
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.rows, self.proc, self.thread = [], None, None
        self.gpu_index = gpu_index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100",
                                          "-i", str(self.gpu_index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except OSError:
            self.proc = None
            return
        def pump():
            for line in self.proc.stdout:
                self.rows.append((time.time(), line.strip()))
        self.thread = threading.Thread(target=pump, daemon=True)
        self.thread.start()

    def stop(self, t0, t1):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        sm, mx, reasons, power = [], None, set(), []
        for ts, line in self.rows:
            parts = [p.strip() for p in line.split(",")]
            if len(parts) < 8:
                continue
            try:
                clk, mxc = float(parts[1]), float(parts[2])
            except ValueError:
                continue
            mx = mxc
            if t0 - 0.05 <= ts <= t1 + 0.15:
                sm.append(clk)
                try:
                    power.append(float(parts[3]))
                except ValueError:
                    pass
                for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), parts[4:8]):
                    if val.lower().startswith("active"):
                        reasons.add(name)
        if not sm:  # timed region shorter than the sampling period: fall back to all samples
            for ts, line in self.rows:
                parts = [p.strip() for p in line.split(",")]
                try:
                    sm.append(float(parts[1]))
                except (ValueError, IndexError):
                    pass
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons),
                "samples": len(sm), "power_w_max": max(power) if power else None}


# ------------------------------------------------------------------------------------------------
# CUDA-event span profiler for selected C-ABI calls (engine.profile_hook)
# ------------------------------------------------------------------------------------------------
class SpanProfiler:
    def __init__(self, names):
        self.names = set(names)
        self.spans = {n: [] for n in names}
        self._open = {}
        self.enabled = True

    def wants(self, name):
        return self.enabled and name in self.names

    def begin(self, name):
        e = torch.cuda.Event(enable_timing=True)
        e.record()
        self._open[name] = e

    def end(self, name):
        e = torch.cuda.Event(enable_timing=True)
        e.record()
        self.spans[name].append((self._open.pop(name), e))

    def mean_ms(self, name):
        xs = [a.elapsed_time(b) for a, b in self.spans[name]]
        return (sum(xs) / len(xs), len(xs)) if xs else (None, 0)

    def total_ms(self, name):
        return sum(a.elapsed_time(b) for a, b in self.spans[name])


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(path) as f:
            d = json.load(f)
        return {"hbm_gbs": float(d["hbm_gbs"]), "bf16_tflops": float(d["bf16_tflops"]),
                "bf16_tflops_sustained": float(d.get("bf16_tflops_sustained", d["bf16_tflops"])), "source": "measured"}
    except Exception:
        return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0, "source": "fallback"}


# ------------------------------------------------------------------------------------------------
# CPU arm: the reference path's restatement (oracle/) on the host cores
# ------------------------------------------------------------------------------------------------
def cpu_train_steps(graphs, steps, warmup, budget_s=None):
    """Times full train steps (fwd + BCE + bwd + Adam) of the oracle on the CPU. Returns (graphs/s, s/step, steps, threads)."""
    from deepdfa_b200 import synth
    from oracle import ggnn_oracle as O
    ncpu = os.cpu_count() or 1
    torch.manual_seed(0)
    model = O.OracleFlowGNNGGNN(FEAT, CFG["input_dim"], CFG["hidden_dim"], CFG["n_steps"], CFG["layers"], concat_all_absdf=True)
    opt = O.make_optimizer(model)
    batches = [synth.make_batch(graphs, CFG["nodes"], CFG["edges_per_node"], CFG["input_dim"], seed=i) for i in range(2)]

    def one(i):
        opt.zero_grad()
        loss, _ = model.training_loss(batches[i % len(batches)])
        loss.backward()
        opt.step()
        return float(loss.detach())
    # The arm may use every host thread, but torch's intra-op pool oversubscribes on many-core hosts for these
    # small ops (128 threads measured 20x slower than 8): probe a few pool sizes, keep the fastest, report it.
    best_t, best_dt = 1, float("inf")
    for cand in sorted({c for c in (4, 8, 16, 32, 64, ncpu) if c <= ncpu}):
        torch.set_num_threads(cand)
        one(0)
        t0 = time.perf_counter()
        one(1)
        dt = time.perf_counter() - t0
        if dt < best_dt:
            best_t, best_dt = cand, dt
        if dt > 3.0 * best_dt:
            break
    torch.set_num_threads(best_t)
    for i in range(warmup):
        one(i)
    t0 = time.perf_counter()
    done = 0
    for i in range(steps):
        one(i)
        done += 1
        if budget_s is not None and time.perf_counter() - t0 > budget_s and done >= 3:
            break
    dt = time.perf_counter() - t0
    return graphs * done / dt, dt / done, done, torch.get_num_threads()


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if rank != 0:
        return 0
    val, s_per_step, done, threads = cpu_train_steps(args.graphs, args.steps, max(args.warmup, 1), budget_s=240.0)
    out = {
        "impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": done, "warmup": args.warmup,
        "ms_per_step": s_per_step * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic", "config": workload_config(args, 1),
        "cpu_baseline": {"value": val, "unit": UNIT, "cores": threads, "kind": "port",
                         "sample": f"{done} full train steps of one {args.graphs}-graph C0 batch (reference cannot run: dgl/"
                                   "pytorch_lightning absent; pure-PyTorch restatement oracle/ggnn_oracle.py, torch CPU, all host threads)"},
        "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "note": f"world_size={world}: rank 0 alone runs the CPU arm",
    }
    print(json.dumps(out), flush=True)
    return 0


# ------------------------------------------------------------------------------------------------
# our arm
# ------------------------------------------------------------------------------------------------
def batch_bytes(g):
    src, dst = g.edges()
    keys = [f"_ABS_DATAFLOW_{k}" for k in ("api", "datatype", "literal", "operator")] + ["_VULN"]
    n = src.numel() * src.element_size() + dst.numel() * dst.element_size() + g.batch_num_nodes().numel() * 8
    for k in keys:
        n += g.ndata[k].numel() * g.ndata[k].element_size()
    return n


def run_ours(args):
    import torch.distributed as dist
    import deepdfa_b200 as D
    from deepdfa_b200 import _lib, engine as E, synth
    from deepdfa_b200.batched_graph import BatchedCFG

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the product path has no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    L = _lib.lib()
    if L.call("ddfa_device_supported") != 1:
        raise SystemExit("bench.py: device is not compute capability 10.x")

    torch.manual_seed(0)
    model = D.FlowGNNGGNNModule(FEAT, CFG["input_dim"], CFG["hidden_dim"], CFG["n_steps"], CFG["layers"], concat_all_absdf=True,
                                engine=args.engine).to(dev)
    trainer = D.FusedTrainer(model)
    # distinct batches per rank and per slot (weak scaling: every rank has its own args.graphs graphs)
    host_batches = [synth.make_batch(args.graphs, CFG["nodes"], CFG["edges_per_node"], CFG["input_dim"], seed=1000 * rank + i).pin_memory()
                    for i in range(NUM_BATCHES)]
    dev_batches = [b.to(dev) for b in host_batches]
    N, Eg = dev_batches[0].num_nodes(), dev_batches[0].num_edges()
    Dh, T = 128, CFG["n_steps"]
    global_batch = args.graphs * world

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- warm-up (also builds + caches the device CSR of every resident batch) -------------------
    min_warm = int(os.environ.get("DDFA_BENCH_MIN_WARMUP", "3"))   # 1 only for profiler runs (never a bench value)
    for i in range(max(args.warmup, min_warm)):
        trainer.step(dev_batches[i % NUM_BATCHES], global_batch)
    torch.cuda.synchronize()
    l0 = L.call("ddfa_launch_count")
    trainer.step(dev_batches[0], global_batch)
    torch.cuda.synchronize()
    launches_per_step = L.call("ddfa_launch_count") - l0

    # ---- instrumented region (eager launches): CUDA-event pairs around every gather / GRU-step call -> roofline -----
    prof = SpanProfiler(["gather_fwd", "gather_bwd", "ddfa_gru_step_fwd", "ddfa_gru_step_bwd", "wgrad_batched"])
    E.profile_hook = prof
    barrier()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    for i in range(args.steps):
        trainer.step(dev_batches[i % NUM_BATCHES], global_batch)
    ev1.record()
    barrier()
    E.profile_hook = None
    ms = ev0.elapsed_time(ev1)                      # denominator of the kernel shares below
    ms_eager_per_step = ms / args.steps

    # ---- capture one CUDA graph per resident batch (launch-bound inner loop: ~100 kernels per step) ------------------
    graph_note = "off (--no-graphs)"
    if args.cuda_graphs:
        try:
            trainer.use_cuda_graph = True
            for i in range(2 * NUM_BATCHES):        # first visit: capture, second visit: replay
                trainer.step(dev_batches[i % NUM_BATCHES], global_batch)
            torch.cuda.synchronize()
            graph_note = f"on ({len(trainer._graphs)} graphs, one per resident batch)"
        except Exception as exc:                    # an execution-mode downgrade, not a compute fallback: same kernels, eager launches
            trainer.use_cuda_graph = False
            trainer._graphs.clear()
            torch.cuda.synchronize()
            graph_note = f"off (capture failed: {type(exc).__name__}: {str(exc)[:120]})"

    # ---- timed region (headline): K resident-input train steps ----------------------------------------------------------
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
        time.sleep(0.25)
    for i in range(max(args.warmup, min_warm)):
        trainer.step(dev_batches[i % NUM_BATCHES], global_batch)
    barrier()
    tv0, tv1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t_wall0 = time.time()
    tv0.record()
    for i in range(args.steps):
        trainer.step(dev_batches[i % NUM_BATCHES], global_batch)
    tv1.record()
    barrier()
    t_wall1 = time.time()
    t = torch.tensor([tv0.elapsed_time(tv1)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_total = float(t.item())
    clocks = sampler.stop(t_wall0, t_wall1) if rank == 0 else None
    final_loss = float(trainer.loss_slot.item())
    value = global_batch * args.steps / (ms_total * 1e-3)
    # e2e below hands HOST batches to the same trainer: with graphs on it copies them into per-shape static device buffers and
    # replays one captured graph that includes the device CSR build (FusedTrainer._step_streamed); with --no-graphs it is eager

    # ---- e2e: host (pinned) buffers -> H2D -> device CSR build -> train step -> loss D2H, every step ----
    def fresh(b):  # a new graph object: no cached device CSR, so the whole input path is inside the timed region
        return BatchedCFG(*b.edges(), b.batch_num_nodes(), dict(b.ndata))
    for i in range(3):
        float(trainer.step(fresh(host_batches[i % NUM_BATCHES]), global_batch).item())
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    e2e_steps = args.steps
    nxt = fresh(host_batches[0])
    trainer.prefetch(nxt, global_batch)
    for i in range(e2e_steps):
        cur = nxt
        loss_t = trainer.step(cur, global_batch)
        nxt = fresh(host_batches[(i + 1) % NUM_BATCHES])
        trainer.prefetch(nxt, global_batch)          # the next step's H2D copies run on a side stream during this step
        loss_val = float(loss_t.item())
    e1.record()
    barrier()
    t2 = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t2, op=dist.ReduceOp.MAX)
    e2e_value = global_batch * e2e_steps / (float(t2.item()) * 1e-3)
    h2d = batch_bytes(host_batches[0])

    # ---- batch producer: the same step fed from a device-resident graph arena by graph-id lists.  Reported
    # next to e2e, not instead of it: here only the id list crosses PCIe each step (the graphs were uploaded once).
    arena = D.GraphArena.from_graphs(host_batches, dev)
    rng = __import__("numpy").random.default_rng(rank)
    id_lists = [rng.integers(0, arena.num_graphs, args.graphs) for _ in range(NUM_BATCHES)]
    for i in range(3):
        float(trainer.step_ids(arena, id_lists[i % NUM_BATCHES], global_batch).item())
    barrier()
    a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a0.record()
    for i in range(e2e_steps):
        arena_loss = float(trainer.step_ids(arena, id_lists[i % NUM_BATCHES], global_batch).item())
    a1.record()
    barrier()
    t3 = torch.tensor([a0.elapsed_time(a1)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t3, op=dist.ReduceOp.MAX)
    arena_value = global_batch * e2e_steps / (float(t3.item()) * 1e-3)

    def leave():
        # Captured CUDA graphs hold NCCL kernels; tearing the communicator down under them can block (seen at N = 2:
        # the JSON line was out, the process never exited).  Drop the graphs, drain the device and leave without the
        # collective teardown — nothing else runs in this process.
        trainer._graphs.clear()
        trainer._stream_slots.clear()
        torch.cuda.synchronize()
        if world > 1:
            sys.stdout.flush()
            sys.stderr.flush()
            os._exit(0)
        return 0

    if rank != 0:
        return leave()

    # ---- rooflines, from the CUDA-event spans of the instrumented region ----------------------------------------------
    # P = one [N,128] fp32 plane = one activation image (hi+lo bf16).  Algorithmic bytes per launch:
    #   forward GRU step (train): read s image, h image, h (3P); write h', h' image, 4 gate planes (6P)            = 9P
    #   backward GRU step (tcgen05: gate_bwd with the transposed gather folded in + dgrad3): gate_bwd reads dh, gates x4, h
    #     (6P) + E gathered ds rows, writes q x4 + dh'z (5P); dgrad reads q x4 + dh'z (5P), writes ds, dh (2P)   = 18P + E rows
    #   weight gradient, ONE launch per backward pass over all T steps: per step q x4 + s image + h image          = 6P x T
    #   (simt engine: the span holds its own gate / sgemm kernels; same byte model, 24P per step)
    #   edge gather: SURVEY.md §8(d) — E rows gathered + N rows written (+ the CSR arrays)
    peaks = measured_peaks()
    P = N * Dh * 4
    share = {k: prof.total_ms(k) / ms for k in prof.spans}
    gather_bytes = Eg * Dh * 4 + N * Dh * 4 + Eg * 4 + (N + 1) * 4
    gf_ms, gf_n = prof.mean_ms("gather_fwd")
    gb_ms, gb_n = prof.mean_ms("gather_bwd")
    g_all = [a.elapsed_time(b) for a, b in prof.spans["gather_fwd"] + prof.spans["gather_bwd"]]
    g_ms = sum(g_all) / len(g_all)
    gru_f_ms, gru_f_n = prof.mean_ms("ddfa_gru_step_fwd")
    gru_b_ms, gru_b_n = prof.mean_ms("ddfa_gru_step_bwd")
    flops_fwd_step = 2.0 * N * (6 * Dh * Dh)                                  # folded GRU GEMMs per propagation step

    def hbm_line(kernel, nbytes, t_ms, launches, sh, **extra):
        ach = nbytes / (t_ms * 1e-3) / 1e9
        return dict({"kernel": kernel, "bound": "hbm", "achieved": ach, "peak": peaks["hbm_gbs"], "unit": "GB/s",
                     "frac": ach / peaks["hbm_gbs"], "peak_source": peaks["source"], "bytes_per_launch": int(nbytes),
                     "avg_launch_us": t_ms * 1e3, "launches_timed": launches, "share_of_step": sh}, **extra)

    tc = args.engine == "tcgen05"
    fwd_line = hbm_line("gru_fwd3_kernel (GRU step forward, tcgen05: weights in TMEM, bf16x3)" if tc else "GRU step forward (simt engine)",
                        9 * P, gru_f_ms, gru_f_n, share["ddfa_gru_step_fwd"],
                        # dram__bytes_read.sum + dram__bytes_write.sum of the training-mode launch, profiles/r02o_ncu_tc_kernels.txt
                        traffic=(59709696 + 43025408) if tc else None,
                        tensor_tflops=3 * flops_fwd_step / (gru_f_ms * 1e-3) / 1e12 if tc else None,
                        tensor_frac=(3 * flops_fwd_step / (gru_f_ms * 1e-3) / 1e12 / peaks["bf16_tflops_sustained"]) if tc else None)
    wg_spans = prof.spans["wgrad_batched"]
    batched = tc and len(wg_spans) > 0
    bwd_line = hbm_line("GRU step backward: gate_bwd_image (+ folded transposed gather) + dgrad3 kernels" if tc else "GRU step backward (simt engine)",
                        (18 * P + Eg * Dh * 4) if batched else 24 * P, gru_b_ms, gru_b_n, share["ddfa_gru_step_bwd"], traffic=None,
                        tensor_tflops=(3 if batched else 6) * flops_fwd_step / (gru_b_ms * 1e-3) / 1e12 if tc else None)
    gather_line = hbm_line("gather_sum_kernel / gather_sum_image_kernel (CSR edge gather, fwd over CSR + bwd over transposed CSR)",
                           gather_bytes, g_ms, len(g_all), share["gather_fwd"] + share["gather_bwd"], traffic=None,
                           fwd_us=gf_ms * 1e3, bwd_us=gb_ms * 1e3, storage_dtype="f32")
    lines = [fwd_line, bwd_line, gather_line]
    if batched:
        wg_ms, wg_n = prof.mean_ms("wgrad_batched")
        lines.insert(2, hbm_line(f"wgrad_kernel + wgrad_reduce_kernel (weight gradients of all {CFG['n_steps']} steps in one launch)",
                                 6 * P * CFG["n_steps"], wg_ms, wg_n, share["wgrad_batched"], traffic=None,
                                 tensor_tflops=3 * flops_fwd_step * CFG["n_steps"] / (wg_ms * 1e-3) / 1e12))
    # the dominant single kernel of the step is the forward GRU kernel (one launch per span); the backward span is three kernels
    roofline = dict(fwd_line, note="dominant single kernel by time; the other hot kernels are in roofline_kernels")

    # ---- cpu baseline on this box's host cores (bounded sample) ----------------------------------------
    if os.environ.get("DDFA_BENCH_SKIP_CPU") == "1":     # profiler runs only
        cpu_val, cpu_s, cpu_done, cpu_threads = None, None, 0, 0
    else:
        cpu_val, cpu_s, cpu_done, cpu_threads = cpu_train_steps(args.graphs, 12, 2, budget_s=20.0)

    out = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
        "ms_per_step": ms_total / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32" if args.engine == "simt" else "f32 (GRU GEMMs: bf16x3 split operands, f32 accumulate)",
        "data": "synthetic", "config": dict(workload_config(args, world), engine=args.engine,
                                             l2="per-step working set ~0.96 GB of saved activations > 126 MB L2; 8 distinct resident batches rotated"),
        "clocks": clocks,
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": 4, "steps": e2e_steps,
                "path": "pinned host COO + node indices -> H2D -> ddfa_build_csr -> fused train step -> loss .item()"
                        + (" (one CUDA graph per batch shape, two static input-buffer sets, next batch prefetched on a copy stream)"
                           if trainer.use_cuda_graph else " (eager launches)")},
        "e2e_arena": {"value": arena_value, "unit": UNIT, "h2d_bytes_per_step": 4 * args.graphs, "d2h_bytes_per_step": 4, "steps": e2e_steps,
                      "path": f"graph-id list (pinned) -> H2D -> ddfa_arena_batch over a resident arena of {arena.num_graphs} graphs -> "
                              "fused train step -> loss .item()", "last_loss": arena_loss},
        "gpu_launches": int(launches_per_step * args.steps), "gpu_launches_per_step": int(launches_per_step),
        "cuda_graph": graph_note, "ms_per_step_eager_instrumented": ms_eager_per_step,
        "roofline": roofline, "roofline_kernels": lines,
        "cpu_baseline": {"value": cpu_val, "unit": UNIT, "cores": cpu_threads, "kind": "port",
                         "sample": f"{cpu_done} full train steps of one {args.graphs}-graph C0 batch, oracle/ggnn_oracle.py (torch CPU)"},
        "final_loss": final_loss, "e2e_last_loss": loss_val,
    }
    print(json.dumps(out), flush=True)
    return leave()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=30)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", choices=["ours", "reference"], default="ours")
    ap.add_argument("--graphs", type=int, default=CFG["graphs"], help="graphs per GPU per step")
    ap.add_argument("--engine", choices=["simt", "tcgen05"], default=os.environ.get("DDFA_B200_ENGINE", "tcgen05"))
    ap.add_argument("--no-graphs", dest="cuda_graphs", action="store_false", help="launch every kernel eagerly in the timed region")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)
    return run_ours(args)


if __name__ == "__main__":
    sys.exit(main())
