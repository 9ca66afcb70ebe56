#!/usr/bin/env python
"""bench.py — headline benchmark of the retrieval hot path (BASELINE.json: queries/sec at 10M subsessions,
d=128, top-100; fused per-session subsession-max + top-k; configs[2]) plus the two larger configurations as extra
blocks of the same JSON line (configs[3]: 100M rows sharded; configs[4]: sessions -> encoder -> sharded search).

    python bench.py --gpus N --steps K --warmup W            # this repo (CUDA, sm_100a)
    python bench.py --impl reference --steps K --warmup W     # the reference's CPU path (oracle port)
    python bench.py --configs headline                        # headline only (default: headline,1m,100m,e2e,binary)

A "step" answers one batch of nq=1000 query sessions against the whole database and returns the top-100
sessions per query; every step uses a different query batch, whose targets are spread over all shards.  With N>1
ranks the database rows are sharded row-wise at session boundaries (strong scaling: the 10M-row database is
fixed), every rank searches its shard, and the per-rank packed (id, score) candidates are merged after ONE NCCL
all-gather.  Outside the timed regions the answers are checked against the CPU oracle (`parity`); a mismatch fails
the run.  Prints one JSON line on rank 0.
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
N_QUERY_BATCHES = 8      # distinct query batches cycled through the timed steps
N_PARITY_QUERIES = 16    # of batch 0, checked against the oracle over ALL rows


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
    ap.add_argument("--configs", default="headline,1m,100m,e2e,binary",
                    help="comma list of: headline (always run), 1m (configs[1]), 100m (configs[3]), e2e (configs[4]), "
                         "binary (Hamming search, N = 1 only)")
    ap.add_argument("--rows-100m", type=int, default=100_000_000)
    ap.add_argument("--e2e-sessions", type=int, default=100_000)
    ap.add_argument("--cpu-sample-rows", type=int, default=10_000_000,
                    help="rows of the CPU baseline leg (default: the whole 10M-row database, measured not scaled)")
    ap.add_argument("--ref-budget-s", type=float, default=100.0, help="time budget of the reference arm's timed region")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extra", action="store_true")
    ap.add_argument("--no-parity", action="store_true")
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


def workload_config(a, rows, world, name="configs[2]"):
    return {"workload": "%s: %d subsession rows (sessions of 1+Poisson(7) contiguous rows), d=%d, nq=%d, "
                        "fused per-session %s + top-%d, %s" % (name, rows, a.d, a.nq, a.reduce, a.k,
                                                               "cosine" if a.metric == "cos" else "squared L2"),
            "rows": rows, "d": a.d, "nq": a.nq, "k": a.k, "reduce": a.reduce, "mode": a.mode,
            "query_batches": "%d distinct batches cycled over the steps, targets drawn from all shards" % N_QUERY_BATCHES,
            "sharding": "rows/%d at session boundaries + 1 NCCL all-gather of packed candidates + merge" % world
                        if world > 1 else "none",
            "l2": "inputs larger than L2 (database %.2f GB bf16 per pass)" % (rows * a.d * 2 / 1e9)}


# ---- the reference arm ------------------------------------------------------------------------------------------

def host_database(rows, d, seed):
    """host copy of the workload generator for the CPU arm (same distribution, numpy RNG)"""
    lens = session_lengths(rows, seed)
    rng = np.random.default_rng(seed + 1)
    base = rng.standard_normal((len(lens), d), dtype=np.float32)
    db = np.repeat(base, lens, axis=0)
    for lo in range(0, db.shape[0], 1 << 20):  # in place, chunked: no second 5 GB temporary
        db[lo:lo + (1 << 20)] += 0.3 * rng.standard_normal((min(1 << 20, db.shape[0] - lo), d), dtype=np.float32)
    seg = np.concatenate([[0], np.cumsum(lens)]).astype(np.int64)
    return db, seg, base


def run_reference(a):
    """The reference's CPU path for this workload (faiss IndexFlatIP over normalize(emb),
    test_amazon_filterd.py:207-214,578, restated by oracle.search_blas since faiss is not installable): all host
    threads.  A step searches as many rows as the time budget allows — the whole database when it fits — and the time
    is scaled linearly in rows otherwise (stated in `sample`)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import search_oracle as so
    import torch
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    red = {"max": so.REDUCE_MAX, "sum": so.REDUCE_SUM, "none": so.REDUCE_NONE}[a.reduce]
    # probe: one step on 1M rows tells how many rows a step may hold
    probe_rows = min(1_000_000, a.rows)
    db, seg, base = host_database(probe_rows, a.d, 1234)
    dbn = so.normalize_util_numpy(db)
    rng = np.random.default_rng(99)
    q = base[rng.integers(0, len(base), size=a.nq)] + 0.3 * rng.standard_normal((a.nq, a.d), dtype=np.float32)
    so.search_blas(dbn[:65536], so.normalize_util_numpy(q), a.k, threads=cores)
    t0 = time.perf_counter()
    so.search_blas(dbn, so.normalize_util_numpy(q), a.k, seg_off=seg, reduce=red, threads=cores)
    per_row = (time.perf_counter() - t0) / probe_rows
    n_steps = max(1, a.steps + a.warmup)
    sample = int(min(a.rows, max(probe_rows, a.ref_budget_s / (per_row * n_steps))))
    if sample > probe_rows:
        del db, dbn
        db, seg, base = host_database(sample, a.d, 1234)
        dbn = so.normalize_util_numpy(db)
        q = base[rng.integers(0, len(base), size=a.nq)] + 0.3 * rng.standard_normal((a.nq, a.d), dtype=np.float32)
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
        "config": workload_config(a, a.rows, 1),
        "cpu_baseline": {"value": value, "unit": "queries/s", "cores": cores, "kind": "port",
                         "sample": ("all %d rows per step (%.3f s/step measured)" % (sample, dt)) if sample == a.rows else
                                   ("%d of %d rows per step (%.3f s/step measured), time scaled linearly in rows"
                                    % (sample, a.rows, dt))},
        "e2e": {"value": value, "unit": "queries/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    emit(out)


# ---- stdout discipline ------------------------------------------------------------------------------------------
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


# ---- this repo's arm --------------------------------------------------------------------------------------------

class Env:
    def __init__(self):
        import torch
        import torch.distributed as dist
        self.torch, self.dist = torch, dist
        self.rank = int(os.environ.get("RANK", "0"))
        self.local_rank = int(os.environ.get("LOCAL_RANK", "0"))
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        torch.cuda.set_device(self.local_rank)
        self.dev = torch.device("cuda", self.local_rank)
        if self.world > 1:
            dist.init_process_group("nccl", device_id=self.dev)

    def barrier(self):
        if self.world > 1:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def timed(self, fn, steps, per_rank=None):
        """K calls of fn(step) between barrier + synchronize on both sides, CUDA events, MAX over ranks (ms total)"""
        torch = self.torch
        self.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(steps):
            fn(i)
        e1.record()
        self.barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], device=self.dev)
        if self.world > 1:
            every = torch.empty(self.world, device=self.dev)
            self.dist.all_gather_into_tensor(every, ms)
            if per_rank is not None:
                per_rank.append([round(float(x) / steps, 4) for x in every])
            ms = every.max()
        return float(ms.item())


def shard_chunks(env, a, lens, seed, chunk=131072):
    """deterministic generator of one shard's rows, chunk by chunk at session boundaries: yields
    (rows [n, d] on the device, lengths of the chunk's sessions, every 997th session base).  Called twice with the same
    arguments it yields the same rows (index build, then the oracle's parity pass)."""
    torch = env.torch
    g = torch.Generator(device=env.dev)
    g.manual_seed(seed)
    lens_t = torch.from_numpy(lens).to(env.dev)
    for c0 in range(0, len(lens), chunk):
        l = lens_t[c0:c0 + chunk]
        base = torch.randn((l.numel(), a.d), generator=g, device=env.dev)
        rows = torch.repeat_interleave(base, l, dim=0)
        rows += 0.3 * torch.randn(rows.shape, generator=g, device=env.dev)
        yield rows, lens[c0:c0 + chunk], base[::997]


def run_search_config(env, a, rows, steps, warmup, name, want_extra, want_cpu, sampler=None):
    """build the (sharded) index of `rows` rows, time the search, check it against the oracle; returns a dict"""
    import sessionsimilaritysearch_b200 as sss
    from sessionsimilaritysearch_b200.dist import ShardedIndex
    torch, dist = env.torch, env.dist
    rank, world, dev = env.rank, env.world, env.dev
    lens_all = session_lengths(rows, 1234)
    n_sess_all = len(lens_all)
    shard_r, shard_w = (int(x) for x in a.shard.split("/")) if a.shard else (rank, world)
    s_lo = n_sess_all * shard_r // shard_w
    s_hi = n_sess_all * (shard_r + 1) // shard_w
    lens = lens_all[s_lo:s_hi]
    row_off = int(lens_all[:s_lo].sum())
    seed = 4321 + shard_r
    index_cls = sss.IndexFlatL2 if a.metric == "l2" else sss.IndexFlatIP
    id_off = s_lo if a.reduce != "none" else row_off
    inner = index_cls(a.d, device=env.local_rank, id_offset=id_off, mode=a.mode)
    t_build = time.perf_counter()
    pool = []
    for rows_c, _, bases in shard_chunks(env, a, lens, seed):
        inner.add(rows_c, norm=sss.NORM_UTIL)
        pool.append(bases.clone())
        del rows_c
    seg = np.concatenate([[0], np.cumsum(lens)]).astype(np.int64)
    if a.reduce != "none":
        inner.set_segments(seg, a.reduce)
    torch.cuda.synchronize()
    t_build = time.perf_counter() - t_build
    index = ShardedIndex(inner, world_size=world, rank=rank, metric=(1 if a.metric == "l2" else 0)) if world > 1 else inner

    # ---- queries: noisy copies of database sessions drawn from ALL shards; N_QUERY_BATCHES distinct batches
    pool = torch.cat(pool)[:1024].contiguous()
    if pool.shape[0] < 1024:
        pool = torch.cat([pool, pool.new_zeros(1024 - pool.shape[0], a.d)])
    if world > 1:
        every = torch.empty((world * 1024, a.d), device=dev)
        dist.all_gather_into_tensor(every, pool)
        pool = every
    pool = pool[pool.abs().sum(1) > 0]
    gq = torch.Generator(device=dev)
    gq.manual_seed(99)
    pick = torch.randint(0, pool.shape[0], (N_QUERY_BATCHES * a.nq,), generator=gq, device=dev)
    q_all = sss.normalize(pool[pick] + 0.3 * torch.randn((N_QUERY_BATCHES * a.nq, a.d), generator=gq, device=dev))
    if world > 1:
        dist.broadcast(q_all, src=0)  # one set of queries for every rank
    q_dev = [q_all[b * a.nq:(b + 1) * a.nq].contiguous() for b in range(N_QUERY_BATCHES)]
    q_host = []
    for qd in q_dev:
        h = torch.empty((a.nq, a.d), dtype=torch.float32).pin_memory()
        h.copy_(qd)
        q_host.append(h.numpy())
    torch.cuda.synchronize()

    per_rank_ms = []
    step_dev = lambda i: index.search(q_dev[i % N_QUERY_BATCHES], a.k)
    step_e2e = lambda i: index.search(q_host[i % N_QUERY_BATCHES], a.k)

    for i in range(max(warmup, N_QUERY_BATCHES if steps > 2 else warmup)):   # every batch once: graphs captured, paths warm
        step_dev(i)
    # A. the headline: device-resident queries, captured graph replay, no profiling
    ms_total = env.timed(step_dev, steps, per_rank_ms)
    st0 = inner.stats()
    kernels_per_step = st0["kernels"] + (1 if world > 1 else 0)
    waves, used_graph = st0["waves"], st0["graph"]
    # B. the same steps with CUDA events around every scan launch (plain launches): the roofline's kernel time
    inner.set_profiling(True)
    scan_ns = [0, 0]

    def step_prof(i):
        r = step_dev(i)
        st = inner.stats()
        scan_ns[0] += st["scan_ns"]
        scan_ns[1] += st["scan_launches"]
        return r

    step_prof(0)
    scan_ns = [0, 0]
    ms_prof = env.timed(step_prof, steps)
    st1 = inner.stats()
    inner.set_profiling(False)
    refine_stats = {k: v for k, v in st1.items() if k.startswith("refine_")}
    reruns, overflow_reason = st1["reruns"], st1["overflow_reason"]
    if world > 1:  # a rerun on ANY rank sets the step time: report the maximum over ranks
        t = torch.tensor([reruns, overflow_reason], device=dev, dtype=torch.int64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        reruns, overflow_reason = int(t[0]), int(t[1])
    scan_kernel_name = {"ss": "scan_bf16_kernel", "ts": "scan_bf16_ts_kernel", "2cta": "scan_bf16_2cta_kernel",
                        "kloop": "scan_bf16_kloop_kernel", "fp32": "scan_fp32_kernel"}[st1["scan_variant"]]
    # C. end to end: pinned host queries in, host results out, through the same public call
    for i in range(max(1, warmup // 2)):
        step_e2e(i)
    ms_e2e = env.timed(step_e2e, steps)

    ms_step = ms_total / steps
    res = {"value": a.nq / (ms_step * 1e-3), "ms_per_step": ms_step, "profiled_ms_per_step": ms_prof / steps,
           "e2e_value": a.nq / (ms_e2e / steps * 1e-3), "e2e_ms_per_step": ms_e2e / steps,
           "kernels_per_step": kernels_per_step, "waves": waves, "graph_replay": bool(used_graph), "reruns": reruns,
           "overflow_reason": overflow_reason, "per_rank_ms": per_rank_ms, "refine_stats": refine_stats,
           "build_s": t_build, "rows_local": int(seg[-1]), "sessions_local": len(seg) - 1,
           "scan_kernel": scan_kernel_name, "scan_s": scan_ns[0] * 1e-9 / steps, "scan_launches": scan_ns[1] / steps}

    extra = {}
    if want_extra and world == 1:
        for mode, k_steps in (("bf16", 5), ("fp32", 2)):
            if mode == a.mode:
                continue
            try:
                fn = lambda i: inner.search(q_dev[i % N_QUERY_BATCHES], a.k, mode=mode)
                fn(0)
                extra[mode + "_qps"] = a.nq / (env.timed(fn, k_steps) / k_steps * 1e-3)
            except RuntimeError as e:
                extra[mode + "_qps"] = "error: %s" % e
        try:  # agreement of the tensor-core modes with the bit-faithful fp32 mode on this very workload, all queries
            D1, I1 = inner.search(q_dev[0], a.k)
            D2, I2 = inner.search(q_dev[0], a.k, mode="fp32")
            extra["ids_equal_fp32_mode"] = bool(torch.equal(I1, I2))
            extra["scores_equal_fp32_mode"] = bool(torch.equal(D1, D2))
            D3, I3 = inner.search(q_dev[0], a.k, mode="bf16")
            i2, i3 = I2.cpu().numpy(), I3.cpu().numpy()
            extra["bf16_mode_recall_vs_fp32"] = float(np.mean([len(set(i3[r]) & set(i2[r])) / float(a.k)
                                                               for r in range(a.nq)]))
            extra["bf16_mode_max_abs_score_diff"] = float((D3 - D2).abs().max())
        except RuntimeError as e:
            extra["ids_equal_fp32_mode"] = "error: %s" % e
        try:  # DB-stream-bound regime: one query per pass
            fn = lambda i: inner.search(q_dev[i % N_QUERY_BATCHES][:1], a.k)
            fn(0)
            ms1 = env.timed(fn, 10) / 10
            extra["nq1_ms"] = ms1
            extra["nq1_db_stream_gbs"] = rows * a.d * 2 / (ms1 * 1e-3) / 1e9
        except RuntimeError as e:
            extra["nq1_ms"] = "error: %s" % e
    res["extra"] = extra
    if sampler is not None and rank == 0:
        res["clocks"] = sampler.stop()

    # ---- parity against the CPU oracle over ALL rows (outside the timed regions): every rank runs the fixed-order
    # oracle O2 over a regenerated host copy of ITS shard, chunk by chunk; rank 0 merges and compares bit for bit
    parity = None
    host_rows = [] if (want_cpu and world == 1 and a.metric == "cos") else None
    if not a.no_parity:
        from oracle import search_oracle as so
        t0 = time.perf_counter()
        Dg, Ig = index.search(q_dev[0], a.k)
        Dg, Ig = Dg.cpu().numpy()[:N_PARITY_QUERIES], Ig.cpu().numpy()[:N_PARITY_QUERIES]
        qs = q_host[0][:N_PARITY_QUERIES]
        red = {"max": so.REDUCE_MAX, "sum": so.REDUCE_SUM, "none": so.REDUCE_NONE}[a.reduce]
        metric = so.METRIC_L2 if a.metric == "l2" else so.METRIC_IP
        Dl, Il = None, None
        s_off = r_off = 0
        for rows_c, lens_c, _ in shard_chunks(env, a, lens, seed):
            h = rows_c.cpu().numpy()
            if host_rows is not None and sum(x.shape[0] for x in host_rows) < a.cpu_sample_rows:
                host_rows.append(h)
            hn = so.normalize(h, so.NORM_UTIL)
            seg_c = np.concatenate([[0], np.cumsum(lens_c)]).astype(np.int64)
            Dc, Ic = so.search_flat(hn, qs, a.k, metric=metric, seg_off=seg_c if a.reduce != "none" else None, reduce=red,
                                    id_offset=id_off + (s_off if a.reduce != "none" else r_off))
            if Dl is None:
                Dl, Il = Dc, Ic
            else:
                Dl, Il = so.topk_merge(np.stack([Dl, Dc]), np.stack([Il, Ic]), metric)
            s_off += len(lens_c)
            r_off += h.shape[0]
            del rows_c
        if a.reduce == "sum":  # a session never spans chunks, so chunk-wise sums are whole-session sums
            pass
        if world > 1:
            td = torch.from_numpy(Dl).to(dev)
            ti = torch.from_numpy(Il).to(dev)
            ad = torch.empty((world,) + td.shape, device=dev)
            ai = torch.empty((world,) + ti.shape, dtype=torch.int64, device=dev)
            dist.all_gather_into_tensor(ad, td)
            dist.all_gather_into_tensor(ai, ti)
            Dl, Il = so.topk_merge(ad.cpu().numpy(), ai.cpu().numpy(), metric)
        ids_ok = bool(np.array_equal(Il, Ig))
        sc_ok = bool(np.array_equal(Dl.view(np.uint32), Dg.view(np.uint32)))
        parity = {"ok": ids_ok and sc_ok, "checked_queries": N_PARITY_QUERIES, "ids_equal": ids_ok,
                  "scores_bit_equal": sc_ok, "rows_checked": rows, "n_gpus": world,
                  "against": "oracle O2 (fixed-order fp32, ties by id) over every row of every shard, merged",
                  "seconds": round(time.perf_counter() - t0, 2)}
    res["parity"] = parity

    # ---- CPU baseline beside it (rank 0, N=1): the oracle's BLAS port of the reference path on the host cores
    cpu_baseline = None
    if host_rows:
        from oracle import search_oracle as so
        cores = os.cpu_count() or 1
        hdb = np.concatenate(host_rows, axis=0) if len(host_rows) > 1 else host_rows[0]
        del host_rows
        n_s = int(np.searchsorted(seg, min(a.cpu_sample_rows, hdb.shape[0]), side="right") - 1)
        n_rows_s = int(seg[n_s])
        hdb = so.normalize_util_numpy(hdb[:n_rows_s])
        red = {"max": so.REDUCE_MAX, "sum": so.REDUCE_SUM, "none": so.REDUCE_NONE}[a.reduce]
        so.search_blas(hdb[:65536], q_host[0], a.k, threads=cores)  # thread-pool warm-up
        t0 = time.perf_counter()
        Dc, Ic = so.search_blas(hdb, q_host[0], a.k, seg_off=seg[:n_s + 1], reduce=red, threads=cores)
        dt = time.perf_counter() - t0
        full = n_rows_s == rows
        cpu_baseline = {"value": a.nq / (dt * rows / n_rows_s), "unit": "queries/s", "cores": cores, "kind": "port",
                        "sample": ("all %d rows (%d sessions), one step of %d queries: %.2f s measured" % (n_rows_s, n_s, a.nq, dt))
                                  if full else ("first %d of %d rows (%d sessions), %.2f s measured, time scaled linearly in rows"
                                                % (n_rows_s, rows, n_s, dt))}
        if full:  # cross-check against the reference's own arithmetic (BLAS summation order): tie-tolerant
            Dg, Ig = index.search(q_dev[0], a.k)
            Ig = Ig.cpu().numpy()
            cpu_baseline["recall_of_ours_vs_blas_path"] = float(np.mean([len(set(Ig[r]) & set(Ic[r])) / float(a.k)
                                                                         for r in range(a.nq)]))
        del hdb
    res["cpu_baseline"] = cpu_baseline
    del index, inner
    torch.cuda.empty_cache()
    return res


def roofline_of(a, res, rows_local, peaks, peak_kind):
    flops = 2.0 * a.nq * rows_local * a.d
    scan_s, ms_step = res["scan_s"], res["ms_per_step"]
    tensor_mode = a.mode in ("exact", "bf16")
    sustained = float(peaks.get("bf16_tflops_sustained", peaks.get("bf16_tflops", 1400.0)))
    burst = float(peaks.get("bf16_tflops", sustained))
    hbm = float(peaks.get("hbm_gbs", 6650.0))
    ridge = sustained * 1e12 / (hbm * 1e9)
    common = {"kernel": res["scan_kernel"], "launches_per_step": res["scan_launches"],
              "kernel_ms_per_step": scan_s * 1e3,
              "kernel_share_of_step": scan_s * 1e3 / res["profiled_ms_per_step"] if res["profiled_ms_per_step"] else None,
              "kernel_time_from": "CUDA events around every scan launch on the launching stream, in a second timed "
                                  "region of the same K steps (plain launches; the headline region replays a graph)",
              "traffic": None}
    if tensor_mode and a.nq < ridge:
        d_pad = (a.d + 63) // 64 * 64
        nq_pad = (a.nq + 127) // 128 * 128
        bytes_step = rows_local * d_pad * 2 + nq_pad * d_pad * 2
        ach = bytes_step / scan_s / 1e9 if scan_s > 0 else 0.0
        r = {"bound": "hbm", "achieved": ach, "peak": hbm, "unit": "GB/s", "frac": ach / hbm,
             "peak_kind": peak_kind + " copy bandwidth", "step_frac": bytes_step / (ms_step * 1e-3) / 1e9 / hbm,
             "algorithmic": "rows*d_pad*2 + nq_pad*d_pad*2 B = %.3e per step (nq %d < ridge %.0f)" % (bytes_step, a.nq, ridge)}
    elif tensor_mode:
        ach = flops / scan_s / 1e12 if scan_s > 0 else 0.0
        step_tf = flops / (ms_step * 1e-3) / 1e12
        r = {"bound": "tensor", "achieved": ach, "peak": sustained, "unit": "TFLOP/s", "frac": ach / sustained,
             "peak_kind": peak_kind + " sustained bf16 (the kernel is timed inside a long step)",
             "peak_burst": burst, "frac_burst": ach / burst,
             "step_achieved": step_tf, "step_frac": step_tf / sustained, "step_frac_burst": step_tf / burst,
             "algorithmic": "2*nq*rows*d flop = %.3e per step" % flops}
        tfile = os.path.join(ROOT, "profiles", "scan_traffic_bytes_per_step.json")
        if os.path.exists(tfile) and rows_local == 10000000 and a.nq == 1000:
            try:
                r_traffic = json.load(open(tfile)).get("bytes_per_step")
                common["traffic"] = r_traffic
                common["traffic_source"] = "static: one ncu --set full capture committed under profiles/ (not measured in this run)"
            except Exception:
                pass
    else:
        peak = 2 * 148 * 128 * 1.965e9 / 1e12  # fp32 FMA peak of the CUDA cores (nominal clocks)
        ach = flops / scan_s / 1e12 if scan_s > 0 else 0.0
        r = {"bound": "tensor", "achieved": ach, "peak": peak, "unit": "TFLOP/s", "frac": ach / peak,
             "peak_kind": "nominal fp32 CUDA-core FMA peak (no tensor path in fp32 mode)"}
    r.update(common)
    return r


def run_e2e(env, a):
    """configs[4]: sessions -> native featuriser -> GNN encoder (768 -> 3 x 800 -> 3168 -> 1600, data-parallel over the
    ranks) -> row-sharded cosine search with per-session max over the subsession embeddings -> top-100 sessions."""
    import sessionsimilaritysearch_b200 as sss
    from sessionsimilaritysearch_b200 import featurize, pipeline, synth
    sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
    import encoder_common as ec
    torch, dist = env.torch, env.dist
    n = a.e2e_sessions
    in_dim, hidden, n_layers, out_dim, msl = 768, 800, 3, 1600, 20
    enc = sss.SessionEncoder(ec.make_params(in_dim, hidden, n_layers, out_dim, msl, 11), in_dim=in_dim, hidden=hidden,
                             n_layers=n_layers, out_dim=out_dim, max_seq_len=msl, device=env.local_rank)
    t0 = time.perf_counter()
    sessions_all = synth.make_sessions(n, 17)
    t_gen = time.perf_counter() - t0
    # one vocabulary / item table for every rank: text features are keyed by string / item id
    vocab = featurize.QueryVocab()
    items = set([0])
    for s in sessions_all:
        for act in s:
            if act[1] == 's':
                vocab(act[2])
            else:
                items.add(act[-1])
    item_ids = np.asarray(sorted(items), dtype=np.int64)
    g = torch.Generator().manual_seed(3)
    cache = featurize.FeatureCache(torch.randn((len(vocab), in_dim), generator=g), item_ids,
                                   torch.randn((len(item_ids), in_dim), generator=g), env.local_rank)
    pipe = pipeline.SessionSearchPipeline(enc, cache, vocab, rank=env.rank, world=env.world, mode=a.mode)
    warm = featurize.flatten(sessions_all[:200], vocab)
    enc(featurize.featurize_batch(warm, cache))
    env.barrier()
    t0 = time.perf_counter()
    pipe.build(sessions_all)
    env.barrier()
    t_build = time.perf_counter() - t0
    rows_total = torch.tensor([pipe.n_rows_local], device=env.dev, dtype=torch.int64)
    if env.world > 1:
        dist.all_reduce(rows_total)
    rows_total = int(rows_total.item())
    queries = [s[:max(2, (2 * len(s)) // 3)] for s in sessions_all]
    pipe.search(queries[:2048 * env.world], a.k)   # warm-up: graphs, arenas
    env.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    e0.record()
    D, I = pipe.search(queries, a.k)
    e1.record()
    env.barrier()
    t_query = time.perf_counter() - t0
    tq = torch.tensor([t_query], device=env.dev)
    if env.world > 1:
        dist.all_reduce(tq, op=dist.ReduceOp.MAX)
    t_query = float(tq.item())
    own = I[:, 0].cpu().numpy() == np.arange(n)
    # encoder arithmetic: 2 * (rows x in x out) over every linear of the forward, from the node counts of one batch
    b = featurize.featurize_batch(featurize.flatten(queries[:200], vocab), cache)
    nq_, np_, ne_ = b["query"].x.shape[0], b["product"].input_ids.shape[0], b["product"].pos_emb_id.shape[0]
    zd = in_dim + n_layers * hidden
    fl = 0.0
    for l in range(n_layers):
        cin = in_dim if l == 0 else hidden
        fl += 2.0 * nq_ * cin * 2 * hidden + 2.0 * np_ * cin * 2 * hidden + 2.0 * np_ * hidden * hidden
        fl += 2.0 * np_ * hidden * 3 * hidden + 2.0 * np_ * cin * 3 * hidden
    fl += 2.0 * (nq_ + np_) * zd * (out_dim - msl) + 2.0 * (nq_ + ne_) * out_dim * out_dim + 2.0 * 200 * out_dim * out_dim
    torch.cuda.synchronize()
    e0.record()
    for _ in range(20):
        enc(b)
    e1.record()
    torch.cuda.synchronize()
    enc_ms = e0.elapsed_time(e1) / 20
    # parity of the wide-row (K-loop) shard search against the oracle on this rank's own rows
    parity = None
    qe_par = None
    if not a.no_parity:  # (a collective when world > 1: every rank takes part, rank 0 uses the result)
        qe_par = sss.normalize(pipe.encode_queries(queries[:8 * env.world])[:8].contiguous())
    if not a.no_parity and env.rank == 0:
        from oracle import search_oracle as so
        inner = pipe.index.inner if env.world > 1 else pipe.index
        lo, hi = pipeline.rank_slice(n, env.rank, env.world)
        subs, seg = pipeline.subsessions(sessions_all[lo:min(hi, lo + 4000)])
        emb, _ = pipeline.encode_sessions(enc, featurize.flatten(subs, vocab), cache)
        small = sss.IndexFlatIP(out_dim, device=env.local_rank, mode=a.mode)
        small.add(emb, norm=sss.NORM_UTIL)
        small.set_segments(seg, "max")
        qe = qe_par
        Ds, Is = small.search(qe, a.k)
        Do, Io = so.search_flat(so.normalize(emb.cpu().numpy(), so.NORM_UTIL), qe.cpu().numpy(), a.k, seg_off=seg,
                                reduce=so.REDUCE_MAX)
        parity = {"ok": bool(np.array_equal(Is.cpu().numpy(), Io) and
                             np.array_equal(Ds.cpu().numpy().view(np.uint32), Do.view(np.uint32))),
                  "checked": "8 queries x %d encoded subsession rows x 1600 of rank 0 (K-loop tensor scan + fp32 "
                             "re-score) against oracle O2, ids and scores bit for bit" % int(seg[-1]),
                  "own_session_first": float(own.mean())}
    tm = pipe.timings
    return {"workload": "configs[4]: %d sessions -> %d subsession rows x 1600 (session max), %d query sessions "
                        "(2/3 prefixes), top-%d sessions; encoder 768 -> 3 x 800 -> 3168 -> 1600, batches of 200"
                        % (n, rows_total, n, a.k),
            "n_gpus": env.world, "session_generation_s": t_gen,
            "db_build_s": t_build, "db_subsessions_per_s": rows_total / t_build,
            "db_encode_share": {"flatten_s": tm["db_flatten_s"], "featurize_host_s": tm["db_featurize_s"],
                                "encode_s": tm["db_encode_s"], "index_s": tm["db_index_s"]},
            "query_sessions_per_s_end_to_end": n / t_query, "query_path_s": t_query,
            "query_path_share": {"encode_s": tm.get("q_encode_s"), "host_featurize_s": tm.get("q_featurize_s"),
                                 "search_s": tm.get("q_search_s"),
                                 "search_batch_ms": [round(x, 1) for x in tm.get("q_search_batch_ms", [])]},
            "query_path": "flatten + native featuriser + encoder (data-parallel) + all-gather of embeddings + "
                          "row-sharded K-loop search + merge; wall clock, max over ranks",
            "encoder": {"ms_per_batch_200": enc_ms, "sessions_per_s_per_gpu": 200.0 / (enc_ms * 1e-3),
                        "launches_per_forward": enc.launches, "flop_per_batch": fl,
                        "algorithmic_tflops": fl / (enc_ms * 1e-3) / 1e12,
                        "issued_bf16_tflops": 3.0 * fl / (enc_ms * 1e-3) / 1e12,
                        "roofline": {"bound": "tensor", "unit": "TFLOP/s", "achieved": 3.0 * fl / (enc_ms * 1e-3) / 1e12,
                                     "peak": float(measured_peaks()[0].get("bf16_tflops", 1660.3)),
                                     "frac": 3.0 * fl / (enc_ms * 1e-3) / 1e12 / float(measured_peaks()[0].get("bf16_tflops", 1660.3)),
                                     "note": "three bf16 products per fp32-accurate multiply (split-bf16 GEMM); at batch 200 the "
                                             "forward is one or two tiles per SM per launch: operand stream and exposed "
                                             "epilogues bound it, not the tensor pipe (profiles/r02_encoder.md)"}},
            "own_session_first": float(own.mean()), "parity": parity}


def main():
    a = parse()
    quiet_stdout()
    if a.impl == "reference":
        return run_reference(a)
    env = Env()
    torch, dist = env.torch, env.dist
    rank, world = env.rank, env.world
    configs = set(x.strip() for x in a.configs.split(",") if x.strip())
    sampler = ClockSampler(env.local_rank)
    if rank == 0:
        sampler.start()  # sampled from the warm-up on: every sample is under load
    res = run_search_config(env, a, a.rows, a.steps, a.warmup, "configs[2]", not a.no_extra,
                            not a.no_cpu_baseline, sampler)
    blocks = {}
    failures = []
    if res["parity"] is not None and not res["parity"]["ok"]:
        failures.append("headline parity")
    if "1m" in configs and world == 1:
        try:  # configs[1]: 1M session embeddings, 1k query batch, cosine top-100, GEMM + top-k only (no session reduce)
            import copy
            a1 = copy.copy(a)
            a1.reduce = "none"
            r1 = run_search_config(env, a1, 1_000_000, max(2, min(a.steps, 10)), 3, "configs[1]", False, False)
            peaks, peak_kind = measured_peaks()
            blocks["1m"] = {"workload": workload_config(a1, 1_000_000, world, "configs[1]")["workload"],
                            "value": r1["value"], "unit": "queries/s", "ms_per_step": r1["ms_per_step"],
                            "e2e_value": r1["e2e_value"], "waves": r1["waves"],
                            "roofline": roofline_of(a1, r1, r1["rows_local"], peaks, peak_kind), "parity": r1["parity"]}
            if r1["parity"] is not None and not r1["parity"]["ok"]:
                failures.append("1m parity")
        except Exception as e:
            blocks["1m"] = {"error": str(e)[:300]}
    if "100m" in configs:
        try:
            steps_b = max(2, min(a.steps, 5))
            r2 = run_search_config(env, a, a.rows_100m, steps_b, 2, "configs[3]", False, False)
            peaks, peak_kind = measured_peaks()
            blocks["100m"] = {"workload": workload_config(a, a.rows_100m, world, "configs[3]")["workload"],
                              "n_gpus": world, "rows_per_gpu": r2["rows_local"], "value": r2["value"],
                              "unit": "queries/s", "ms_per_step": r2["ms_per_step"], "steps": steps_b,
                              "e2e_value": r2["e2e_value"], "build_s": r2["build_s"], "waves": r2["waves"],
                              "reruns": r2["reruns"], "scaling": "strong (100M rows fixed, rows/GPU = 100M / N)",
                              "roofline": roofline_of(a, r2, r2["rows_local"], peaks, peak_kind) if rank == 0 else None,
                              "parity": r2["parity"]}
            if r2["parity"] is not None and not r2["parity"]["ok"]:
                failures.append("100m parity")
        except Exception as e:  # the headline must not depend on the larger configurations
            blocks["100m"] = {"error": str(e)[:300]}
    if "e2e" in configs:
        try:
            blocks["e2e"] = run_e2e(env, a)
            p = blocks["e2e"].get("parity")
            if p is not None and not p["ok"]:
                failures.append("e2e parity")
        except Exception as e:
            blocks["e2e"] = {"error": str(e)[:300]}
    if "binary" in configs and world == 1:
        try:
            blocks["binary"] = run_binary(env, a)
        except Exception as e:
            blocks["binary"] = {"error": str(e)[:300]}

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    peaks, peak_kind = measured_peaks()
    rows_local = res["rows_local"] if a.reduce != "sum" else res["sessions_local"]
    roofline = roofline_of(a, res, rows_local, peaks, peak_kind)
    tensor_mode = a.mode in ("exact", "bf16")
    out = {
        "metric": METRIC, "value": res["value"], "unit": "queries/s", "n_gpus": world, "steps": a.steps,
        "warmup": a.warmup, "ms_per_step": res["ms_per_step"], "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "bf16" if tensor_mode else "f32", "data": "synthetic",
        "config": workload_config(a, a.rows, world),
        "e2e": {"value": res["e2e_value"], "unit": "queries/s", "h2d_bytes_per_step": int(a.nq * a.d * 4),
                "d2h_bytes_per_step": int(a.nq * a.k * 12), "ms_per_step": res["e2e_ms_per_step"]},
        "gpu_launches": int(res["kernels_per_step"] * a.steps),
        "clocks": res.get("clocks"),
        "roofline": roofline,
        "cpu_baseline": res["cpu_baseline"],
        "parity": res["parity"],
        "graph_replay": res["graph_replay"], "profiled_ms_per_step": res["profiled_ms_per_step"],
        "waves_per_step": res["waves"], "overflow_reruns": res["reruns"], "overflow_reason": res["overflow_reason"],
        "per_rank_ms_per_step": res["per_rank_ms"], "refine_volumes_last_step": res["refine_stats"],
        "build_s": res["build_s"],
        "extra": res["extra"],
        "configs": blocks,
    }
    if failures:
        out["failed"] = failures
    emit(out)
    if world > 1:
        dist.destroy_process_group()
    if failures:
        sys.stderr.write("bench.py: PARITY FAILURE: %s\n" % ", ".join(failures))
        sys.exit(1)


def run_binary(env, a, rows=None, nq_list=(1000, 128, 8, 1), k=100):
    """Hamming top-k over 256-bit codes (what fine_tune_ours.test() executes as committed: code_len 250
    -> 256-bit codes -> faiss.IndexBinaryFlat, fine_tune_ours.py:826,839-843,871-876): +-1 E4M3 tensor-core scan above 16
    queries per call, popcount scan over the packed codes below."""
    import sessionsimilaritysearch_b200 as sss
    torch = env.torch
    rows = int(rows or a.rows_100m)
    nbits, nbytes = 256, 32
    g = torch.Generator(device=env.dev).manual_seed(7)
    ix = sss.IndexBinaryFlat(nbits, device=env.local_rank)
    pool = None
    for lo in range(0, rows, 4_000_000):
        n = min(4_000_000, rows - lo)
        c = torch.randint(0, 256, (n, nbytes), generator=g, device=env.dev, dtype=torch.uint8)
        if pool is None:
            pool = c[:4096].clone()
        ix.add(c)
        del c
    out = {"workload": "%d codes x %d bit (IndexBinaryFlat), top-%d, queries = database codes with 8%% of the bits "
                       "flipped" % (rows, nbits, k), "rows": rows}
    peaks_hbm = None
    for nq in nq_list:
        flip = (torch.rand((nq, nbytes, 8), generator=g, device=env.dev) < 0.08)
        w = (2 ** torch.arange(7, -1, -1, device=env.dev)).to(torch.int32)
        q = pool[:nq] ^ (flip.to(torch.int32) * w).sum(-1).to(torch.uint8)
        for _ in range(3):
            D, I = ix.search(q, k)
        ms = env.timed(lambda i: ix.search(q, k), 5) / 5
        st = ix.stats()
        ix.set_profiling(True)
        ix.search(q, k)
        sp = ix.stats()
        ix.set_profiling(False)
        tensor = st["scan_variant"] in ("ts", "2cta")
        row_bytes = 256 if tensor else nbytes
        blk = {"ms_per_search": ms, "queries_per_s": nq / (ms * 1e-3), "scan_variant": st["scan_variant"],
               "waves": st["waves"], "graph_replay": bool(st["graph"]),
               "scan_kernel_ms": sp["scan_ns"] * 1e-6,
               "db_stream_gbs_as_stored": rows * row_bytes / (ms * 1e-3) / 1e9,
               "db_stream_gbs_packed_equivalent": rows * nbytes / (ms * 1e-3) / 1e9,
               "pairs_per_s": nq * rows / (ms * 1e-3),
               "tensor_tflops_fp8": (2.0 * nq * rows * nbits / (ms * 1e-3) / 1e12) if tensor else None}
        # parity on a sample: the popcount oracle over the first 2M codes against a 2M-code index
        out["nq%d" % nq] = blk
    if env.rank == 0 and not a.no_parity:
        from oracle import search_oracle as so
        n_s = min(rows, 2_000_000)
        gs = torch.Generator(device=env.dev).manual_seed(7)
        c = torch.randint(0, 256, (min(4_000_000, rows), nbytes), generator=gs, device=env.dev, dtype=torch.uint8)[:n_s]
        small = sss.IndexBinaryFlat(nbits, device=env.local_rank)
        small.add(c)
        q = pool[:64] ^ 1
        Ds, Is = small.search(q, k)
        Do, Io = so.search_hamming(c.cpu().numpy(), q.cpu().numpy(), k)
        D4, I4 = small.search(q[:4].contiguous(), k)       # the popcount path (<= 16 queries)
        out["parity"] = {"ok": bool(np.array_equal(Ds.cpu().numpy(), Do) and np.array_equal(Is.cpu().numpy(), Io) and
                                    np.array_equal(D4.cpu().numpy(), Do[:4]) and np.array_equal(I4.cpu().numpy(), Io[:4])),
                         "checked": "64 queries (tensor path) and 4 queries (popcount path) x %d codes against the "
                                    "popcount oracle, distances and ids" % n_s}
    return out


if __name__ == "__main__":
    main()
