#!/usr/bin/env python
"""bench.py — headline benchmark of the retrieval hot path (BASELINE.json: queries/sec at 10M subsessions,
d=128, top-100; fused per-session subsession-max + top-k; configs[2]).

    python bench.py --gpus N --steps K --warmup W            # this repo (CUDA, sm_100a)
    python bench.py --impl reference --steps K --warmup W     # the reference's CPU path (oracle port)

A "step" answers one batch of nq=1000 query sessions against the whole database and returns the top-100
sessions per query.  With N>1 ranks the database rows are sharded row-wise at session boundaries (strong
scaling: the 10M-row database is fixed), every rank searches its shard, and the per-rank (score, id)
candidates are merged after ONE NCCL all-gather.  Prints one JSON line on rank 0.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "queries/sec at 10M subsessions d=128 top-100"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--rows", type=int, default=10_000_000)
    ap.add_argument("--nq", type=int, default=1000)
    ap.add_argument("--k", type=int, default=100)
    ap.add_argument("--d", type=int, default=128)
    ap.add_argument("--mode", default="exact", choices=["exact", "bf16", "fp32"])
    ap.add_argument("--reduce", default="max", choices=["max", "sum", "none"])
    ap.add_argument("--metric", default="cos", choices=["cos", "l2"],
                    help="cos = the headline (inner product over normalised rows); l2 = squared L2 over the same rows")
    ap.add_argument("--cpu-sample-rows", type=int, default=1_000_000)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extra", action="store_true")
    ap.add_argument("--shard", default=None,
                    help="diagnostics: R/W = build and search only shard R of a W-way row sharding on ONE GPU (no collective)")
    return ap.parse_args()


def session_lengths(total_rows, seed):
    """1 + Poisson(7) subsession rows per session until total_rows are used"""
    rng = np.random.default_rng(seed)
    lens = []
    left = total_rows
    while left > 0:
        blk = 1 + rng.poisson(7, size=max(1024, left // 8 + 1))
        c = np.cumsum(blk)
        cut = int(np.searchsorted(c, left, side="left"))
        if cut < len(blk):
            blk = blk[:cut + 1].copy()
            blk[-1] -= int(c[cut] - left)
            lens.append(blk[blk > 0])
            left = 0
        else:
            lens.append(blk)
            left -= int(c[-1])
    lens = np.concatenate(lens).astype(np.int64)
    assert lens.sum() == total_rows
    return lens


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region (B200_PROFILING.md)"""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.idx = gpu_index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.idx), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "20"], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        for l in self.lines:
            f = [x.strip() for x in l.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                mx.append(float(f[2]))
            except ValueError:
                continue
            for name, v in zip(["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"], f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        j = json.load(open(p))
        return j, "measured"
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0}, "fallback"


def host_sample(rows, d, seed):
    """host copy of the workload generator for the CPU arm (same distribution, numpy RNG)"""
    lens = session_lengths(rows, seed)
    rng = np.random.default_rng(seed + 1)
    base = rng.standard_normal((len(lens), d), dtype=np.float32)
    db = np.repeat(base, lens, axis=0)
    db += 0.3 * rng.standard_normal(db.shape, dtype=np.float32)
    seg = np.concatenate([[0], np.cumsum(lens)]).astype(np.int64)
    return db, seg, base


def run_reference(a):
    """The reference's CPU path for this workload (faiss IndexFlatIP over normalize(emb),
    test_amazon_filterd.py:207-214,578, restated by oracle.search_blas since faiss is not installable):
    all host threads, a bounded row sample per step, extrapolated linearly in rows."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import search_oracle as so
    import torch
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    sample = min(a.cpu_sample_rows, a.rows)
    db, seg, base = host_sample(sample, a.d, 1234)
    dbn = so.normalize_util_numpy(db)
    rng = np.random.default_rng(99)
    q = base[rng.integers(0, len(base), size=a.nq)] + 0.3 * rng.standard_normal((a.nq, a.d), dtype=np.float32)
    red = {"max": so.REDUCE_MAX, "sum": so.REDUCE_SUM, "none": so.REDUCE_NONE}[a.reduce]
    for _ in range(a.warmup):
        so.search_blas(dbn, so.normalize_util_numpy(q), a.k, seg_off=seg, reduce=red, threads=cores)
    t0 = time.perf_counter()
    for _ in range(a.steps):
        so.search_blas(dbn, so.normalize_util_numpy(q), a.k, seg_off=seg, reduce=red, threads=cores)
    dt = (time.perf_counter() - t0) / max(a.steps, 1)
    scale = a.rows / float(sample)
    value = a.nq / (dt * scale)
    out = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": "queries/s", "n_gpus": a.gpus, "steps": a.steps,
        "warmup": a.warmup, "ms_per_step": dt * scale * 1e3, "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(a),
        "cpu_baseline": {"value": value, "unit": "queries/s", "cores": cores, "kind": "port",
                         "sample": "%d of %d rows per step (%.3f s/step measured), time scaled linearly in rows"
                                   % (sample, a.rows, dt)},
        "e2e": {"value": value, "unit": "queries/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    emit(out)


def encoder_bench(device, no_cpu=False, n_sessions=2000, batch=200):
    """config 5, encoder leg: the reference's model shape (768 -> 3 x 800 -> 3168 -> 1600, pretrain_filtered_amazon.py:
    262-287) on synthetic Amazon-filtered-shaped sessions, eval batch size 200 (test_amazon_filterd.py:488); features
    are precomputed (the private text model is outside the CUDA scope).  CPU figure: the encoder oracle, one batch."""
    import torch
    sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
    import encoder_common as ec
    import sessionsimilaritysearch_b200 as sss
    from sessionsimilaritysearch_b200 import graph, sessions, synth
    in_dim, hidden, n_layers, out_dim, msl = 768, 800, 3, 1600, 20
    P = ec.make_params(in_dim, hidden, n_layers, out_dim, msl, 11)
    enc = sss.SessionEncoder(P, in_dim=in_dim, hidden=hidden, n_layers=n_layers, out_dim=out_dim, max_seq_len=msl,
                             device=device)
    _, graphs = ec.make_graphs(n_sessions, in_dim, 17, sessions.sequence_to_graph)
    batches = [graph.collate(graphs[i:i + batch]).to("cuda:%d" % device) for i in range(0, n_sessions, batch)]
    enc.set_math("fp32")  # first figure: cuBLAS pedantic sgemm (the bit-faithful arithmetic)
    for b in batches[:2]:
        enc(b)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for b in batches:
        enc(b)
    e1.record()
    torch.cuda.synchronize()
    out = {"encoder_sessions_per_s": n_sessions / (e0.elapsed_time(e1) * 1e-3),
           "encoder_nodes_per_batch": int(batches[0]['query'].x.shape[0] + batches[0]['product'].x.shape[0])}
    try:  # dense linears on this library's split-bf16 tcgen05 GEMM
        enc.set_math("bf16x3")
        for b in batches[:2]:
            enc(b)
        e0.record()
        for b in batches:
            enc(b)
        e1.record()
        torch.cuda.synchronize()
        out["encoder_bf16x3_sessions_per_s"] = n_sessions / (e0.elapsed_time(e1) * 1e-3)
    except RuntimeError as e:
        out["encoder_bf16x3_sessions_per_s"] = "error: %s" % str(e)[:120]
    try:  # dense linears on the bf16 tensor cores (cuBLAS fp32 emulation), when the loaded cuBLAS has it
        enc.set_math("bf16x9")
        for b in batches[:2]:
            enc(b)
        e0.record()
        for b in batches:
            enc(b)
        e1.record()
        torch.cuda.synchronize()
        out["encoder_bf16x9_sessions_per_s"] = n_sessions / (e0.elapsed_time(e1) * 1e-3)
    except RuntimeError as e:
        out["encoder_bf16x9_sessions_per_s"] = "unavailable: %s" % str(e)[:80]
    # host featuriser: the reference's per-session Python (mirrored by sessions.sequence_to_graph + graph.collate)
    # against the native batched sss_featurize_batch on the same sessions
    from sessionsimilaritysearch_b200 import featurize
    sess_all = synth.make_sessions(n_sessions, 17)
    tok = synth.HashTokenizer()
    t0 = time.perf_counter()
    for i in range(0, 200, 200):
        graph.collate([sessions.sequence_to_graph(0, s, s[:1], tok, 20) for s in sess_all[i:i + 200]])
    out["featurizer_python_sessions_per_s"] = 200 / (time.perf_counter() - t0)
    t0 = time.perf_counter()
    vocab = featurize.QueryVocab()
    flat = featurize.flatten(sess_all, vocab)
    t1 = time.perf_counter()
    for i in range(0, n_sessions, batch):
        featurize.featurize_arrays(flat.slice(i, min(n_sessions, i + batch)))
    t2 = time.perf_counter()
    out["featurizer_native_sessions_per_s"] = n_sessions / (t2 - t0)
    out["featurizer_native_split"] = "flatten (python) %.1f us + native %.1f us per session" % (
        (t1 - t0) / n_sessions * 1e6, (t2 - t1) / n_sessions * 1e6)
    if not no_cpu:
        from oracle import encoder_oracle as eo
        cb = eo.batch_from_pyg(graph.collate(graphs[:batch]))
        t0 = time.perf_counter()
        eo.encoder_forward(P, cb, n_layers)
        out["encoder_cpu_sessions_per_s"] = batch / (time.perf_counter() - t0)
        out["encoder_cpu_threads"] = torch.get_num_threads()
    return out


def workload_config(a):
    return {"workload": "configs[2]: %d subsession rows (sessions of 1+Poisson(7) contiguous rows), d=%d, nq=%d, "
                        "fused per-session %s + top-%d, %s" % (a.rows, a.d, a.nq, a.reduce, a.k,
                                                               "cosine" if a.metric == "cos" else "squared L2"),
            "rows": a.rows, "d": a.d, "nq": a.nq, "k": a.k, "reduce": a.reduce, "mode": a.mode,
            "sharding": "rows/%d at session boundaries + 1 NCCL all-gather merge" % a.gpus if a.gpus > 1 else "none",
            "l2": "inputs larger than L2 (database %.2f GB bf16 per pass)" % (a.rows * a.d * 2 / 1e9)}


_JSON_FD = None


def quiet_stdout():
    """stdout carries exactly ONE JSON line: everything libraries print there (NCCL's version banner, cuBLAS or
    driver notices) is sent to stderr; emit() writes the line to the real stdout."""
    global _JSON_FD
    if _JSON_FD is None:
        sys.stdout.flush()
        _JSON_FD = os.dup(1)
        os.dup2(2, 1)


def emit(obj):
    line = (json.dumps(obj) + "\n").encode()
    if _JSON_FD is None:
        sys.stdout.write(line.decode())
        sys.stdout.flush()
    else:
        os.write(_JSON_FD, line)


def main():
    a = parse()
    quiet_stdout()
    if a.impl == "reference":
        return run_reference(a)
    import torch
    import torch.distributed as dist
    import sessionsimilaritysearch_b200 as sss
    from sessionsimilaritysearch_b200.dist import ShardedIndex

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    # ---- synthetic database, generated on the device shard by shard -------------------------------
    lens_all = session_lengths(a.rows, 1234)
    n_sess_all = len(lens_all)
    shard_r, shard_w = (int(x) for x in a.shard.split("/")) if a.shard else (rank, world)
    s_lo = n_sess_all * shard_r // shard_w
    s_hi = n_sess_all * (shard_r + 1) // shard_w
    lens = lens_all[s_lo:s_hi]
    row_off = int(lens_all[:s_lo].sum())
    g = torch.Generator(device=dev)
    g.manual_seed(4321 + shard_r)
    index_cls = sss.IndexFlatL2 if a.metric == "l2" else sss.IndexFlatIP
    inner = index_cls(a.d, device=local_rank, id_offset=(s_lo if a.reduce != "none" else row_off), mode=a.mode)
    lens_t = torch.from_numpy(lens).to(dev)
    host_rows = []
    chunk = 131072
    base_for_q = None
    for c0 in range(0, len(lens), chunk):
        l = lens_t[c0:c0 + chunk]
        base = torch.randn((l.numel(), a.d), generator=g, device=dev)
        if base_for_q is None:
            base_for_q = base[:8192].clone()
        rows = torch.repeat_interleave(base, l, dim=0)
        rows += 0.3 * torch.randn(rows.shape, generator=g, device=dev)
        inner.add(rows, norm=sss.NORM_UTIL)
        if rank == 0 and not a.no_cpu_baseline and sum(x.shape[0] for x in host_rows) < a.cpu_sample_rows:
            host_rows.append(rows.cpu().numpy())
        del rows, base
    seg = np.concatenate([[0], np.cumsum(lens)]).astype(np.int64)
    if a.reduce != "none":
        inner.set_segments(seg, a.reduce)
    index = ShardedIndex(inner, world_size=world, rank=rank, metric=(1 if a.metric == "l2" else 0)) if world > 1 else inner

    # ---- queries: noisy copies of database sessions (a query subsession resembles its session) ------
    gq = torch.Generator(device=dev)
    gq.manual_seed(99)  # same queries on every rank
    if world > 1:
        dist.broadcast(base_for_q, src=0)
    elif a.shard and shard_r != 0:  # the sharded run's queries come from rank 0's first sessions: regenerate those
        g0 = torch.Generator(device=dev)
        g0.manual_seed(4321)
        n0 = min(chunk, n_sess_all // shard_w)
        base_for_q = torch.randn((n0, a.d), generator=g0, device=dev)[:8192].clone()
    pick = torch.randint(0, base_for_q.shape[0], (a.nq,), generator=gq, device=dev)
    q_dev = sss.normalize(base_for_q[pick] + 0.3 * torch.randn((a.nq, a.d), generator=gq, device=dev))
    q_host = torch.empty((a.nq, a.d), dtype=torch.float32).pin_memory()
    q_host.copy_(q_dev)
    q_np = q_host.numpy()
    torch.cuda.synchronize()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            every = torch.empty(world, device=dev)
            dist.all_gather_into_tensor(every, ms)
            per_rank_ms.append([round(float(x) / steps, 4) for x in every])
            ms = every.max()
        return float(ms.item())

    per_rank_ms = []  # ms per step of every rank, one list per timed region (diagnostics: which rank sets the max)

    def step_dev():
        return index.search(q_dev, a.k)

    def step_e2e():
        return index.search(q_np, a.k)

    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()  # sampled from the warm-up on: every sample is under load
    for _ in range(a.warmup):
        step_dev()
    inner.set_profiling(True)
    scan_ns = [0, 0]

    def step_dev_prof():
        r = step_dev()
        st = inner.stats()
        scan_ns[0] += st["scan_ns"]
        scan_ns[1] += st["scan_launches"]
        return r

    ms_total = timed(step_dev_prof, a.steps)
    inner.set_profiling(False)
    refine_stats = {k: v for k, v in inner.stats().items() if k.startswith("refine_")}
    kernels_per_step = inner.stats()["kernels"] + (1 if world > 1 else 0)
    waves = inner.stats()["waves"]
    reruns = inner.stats()["reruns"]
    overflow_reason = inner.stats()["overflow_reason"]
    if world > 1:  # a rerun on ANY rank sets the step time: report the maximum over ranks
        t = torch.tensor([reruns, overflow_reason], device=dev, dtype=torch.int64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        reruns, overflow_reason = int(t[0]), int(t[1])
    scan_kernel_name = {"ss": "scan_bf16_kernel", "ts": "scan_bf16_ts_kernel", "2cta": "scan_bf16_2cta_kernel",
                        "kloop": "scan_bf16_kloop_kernel", "fp32": "scan_fp32_kernel"}[inner.stats()["scan_variant"]]
    for _ in range(max(1, a.warmup // 2)):
        step_e2e()
    ms_e2e = timed(step_e2e, a.steps)

    ms_step = ms_total / a.steps
    value = a.nq / (ms_step * 1e-3)
    e2e_value = a.nq / (ms_e2e / a.steps * 1e-3)

    extra = {}
    if not a.no_extra and world == 1:
        for mode, steps in (("bf16", 5), ("fp32", 2)):
            if mode == a.mode:
                continue
            try:
                fn = lambda: inner.search(q_dev, a.k, mode=mode)
                fn()
                extra[mode + "_qps"] = a.nq / (timed(fn, steps) / steps * 1e-3)
            except RuntimeError as e:
                extra[mode + "_qps"] = "error: %s" % e
        # agreement of the headline mode with the bit-faithful fp32 mode on this very workload
        try:
            D1, I1 = inner.search(q_dev, a.k)
            D2, I2 = inner.search(q_dev, a.k, mode="fp32")
            extra["ids_equal_fp32_mode"] = bool(torch.equal(I1, I2))
            extra["scores_equal_fp32_mode"] = bool(torch.equal(D1, D2))
        except RuntimeError as e:
            extra["ids_equal_fp32_mode"] = "error: %s" % e
        # DB-stream-bound regime: one query per pass
        try:
            fn = lambda: inner.search(q_dev[:1], a.k)
            fn()
            ms1 = timed(fn, 10) / 10
            extra["nq1_ms"] = ms1
            extra["nq1_db_stream_gbs"] = a.rows * a.d * 2 / (ms1 * 1e-3) / 1e9
        except RuntimeError as e:
            extra["nq1_ms"] = "error: %s" % e

    # (the sampler also covers the other search modes above: more samples, all of them under this process' load)
    clocks = sampler.stop() if rank == 0 else None

    if not a.no_extra and world == 1:
        try:
            extra.update(encoder_bench(local_rank, no_cpu=a.no_cpu_baseline))
        except Exception as e:  # the headline number must not depend on the encoder leg
            extra["encoder_error"] = str(e)[:200]

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    peaks, peak_kind = measured_peaks()
    rows_local = int(seg[-1])
    if a.reduce == "sum":  # linearity: the scan runs over the per-session summed rows
        rows_local = len(seg) - 1
    flops_per_step = 2.0 * a.nq * rows_local * a.d
    scan_s = scan_ns[0] * 1e-9 / a.steps
    tensor_mode = a.mode in ("exact", "bf16")
    ridge = float(peaks.get("bf16_tflops_sustained", 1400.0)) * 1e12 / (float(peaks.get("hbm_gbs", 6650.0)) * 1e9)
    if tensor_mode and a.nq < ridge:
        # fewer resident queries than the ridge (flop per DB byte): the pass is bound by the DB stream
        d_pad = (a.d + 63) // 64 * 64
        nq_pad = (a.nq + 127) // 128 * 128
        bytes_per_step = rows_local * d_pad * 2 + nq_pad * d_pad * 2
        peak = float(peaks.get("hbm_gbs", 6650.0))
        achieved = bytes_per_step / scan_s / 1e9 if scan_s > 0 else 0.0
        roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                    "traffic": None, "peak_kind": peak_kind + " copy bandwidth",
                    "kernel": scan_kernel_name, "launches_per_step": scan_ns[1] / a.steps,
                    "kernel_ms_per_step": scan_s * 1e3, "kernel_share_of_step": scan_s * 1e3 / ms_step,
                    "algorithmic": "rows*d_pad*2 + nq_pad*d_pad*2 B = %.3e per step (nq %d < ridge %.0f)"
                                   % (bytes_per_step, a.nq, ridge),
                    "whole_step_gbs": bytes_per_step / (ms_step * 1e-3) / 1e9}
    elif tensor_mode:
        peak = float(peaks.get("bf16_tflops_sustained", peaks.get("bf16_tflops", 1400.0)))
        achieved = flops_per_step / scan_s / 1e12 if scan_s > 0 else 0.0
        roofline = {"bound": "tensor", "achieved": achieved, "peak": peak, "unit": "TFLOP/s",
                    "frac": achieved / peak, "traffic": None, "peak_kind": peak_kind + " sustained bf16",
                    "kernel": scan_kernel_name, "launches_per_step": scan_ns[1] / a.steps,
                    "kernel_ms_per_step": scan_s * 1e3, "kernel_share_of_step": scan_s * 1e3 / ms_step,
                    "algorithmic": "2*nq*rows*d flop = %.3e per step" % flops_per_step}
    else:
        peak = 2 * 148 * 128 * 1.965e9 / 1e12  # fp32 FMA peak of the CUDA cores (nominal clocks)
        achieved = flops_per_step / scan_s / 1e12 if scan_s > 0 else 0.0
        roofline = {"bound": "tensor", "achieved": achieved, "peak": peak, "unit": "TFLOP/s", "frac": achieved / peak,
                    "traffic": None, "peak_kind": "nominal fp32 CUDA-core FMA peak (no tensor path in fp32 mode)",
                    "kernel": "scan_fp32_kernel"}
    tfile = os.path.join(ROOT, "profiles", "scan_traffic_bytes_per_step.json")
    if os.path.exists(tfile) and roofline["bound"] == "tensor" and a.rows == 10000000 and a.nq == 1000:
        try:
            roofline["traffic"] = json.load(open(tfile)).get("bytes_per_step")
        except Exception:
            pass

    cpu_baseline = None
    if not a.no_cpu_baseline and world == 1 and a.metric == "cos":
        from oracle import search_oracle as so
        cores = os.cpu_count() or 1
        hdb = np.concatenate(host_rows, axis=0)
        n_s = int(np.searchsorted(seg, min(a.cpu_sample_rows, hdb.shape[0]), side="right") - 1)
        n_rows_s = int(seg[n_s])
        hdb = so.normalize_util_numpy(hdb[:n_rows_s])
        red = {"max": so.REDUCE_MAX, "sum": so.REDUCE_SUM, "none": so.REDUCE_NONE}[a.reduce]
        so.search_blas(hdb[:65536], q_np, a.k, threads=cores)  # thread-pool warm-up
        t0 = time.perf_counter()
        Dc, Ic = so.search_blas(hdb, q_np, a.k, seg_off=seg[:n_s + 1], reduce=red, threads=cores)
        dt = time.perf_counter() - t0
        cpu_baseline = {"value": a.nq / (dt * a.rows / n_rows_s), "unit": "queries/s", "cores": cores, "kind": "port",
                        "sample": "first %d of %d rows (%d sessions), %.2f s measured, time scaled linearly in rows"
                                  % (n_rows_s, a.rows, n_s, dt)}

    out = {
        "metric": METRIC, "value": value, "unit": "queries/s", "n_gpus": world, "steps": a.steps, "warmup": a.warmup,
        "ms_per_step": ms_step, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
        "dtype": "bf16" if tensor_mode else "f32", "data": "synthetic",
        "config": workload_config(a),
        "e2e": {"value": e2e_value, "unit": "queries/s", "h2d_bytes_per_step": int(a.nq * a.d * 4),
                "d2h_bytes_per_step": int(a.nq * a.k * 12), "ms_per_step": ms_e2e / a.steps},
        "gpu_launches": int(kernels_per_step * a.steps),
        "clocks": clocks,
        "roofline": roofline,
        "cpu_baseline": cpu_baseline,
        "waves_per_step": waves, "overflow_reruns": reruns, "overflow_reason": overflow_reason, "per_rank_ms_per_step": per_rank_ms, "refine_volumes_last_step": refine_stats,
        "extra": extra,
    }
    emit(out)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
