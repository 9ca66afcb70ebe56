"""Encoder timing at the reference's model shape (768 -> 3 x 800 -> 3168 -> 1600) and eval batch size (200 sessions).
Run plain for CUDA-event timings, or under `ncu --metrics gpu__time_duration.sum` for the per-kernel launch list."""
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
import encoder_common as ec  # noqa: E402
import sessionsimilaritysearch_b200 as sss  # noqa: E402
from sessionsimilaritysearch_b200 import graph, sessions  # noqa: E402


def main():
    n_batches = int(sys.argv[1]) if len(sys.argv) > 1 else 6
    batch = 200
    in_dim, hidden, n_layers, out_dim, msl = 768, 800, 3, 1600, 20
    P = ec.make_params(in_dim, hidden, n_layers, out_dim, msl, 11)
    enc = sss.SessionEncoder(P, in_dim=in_dim, hidden=hidden, n_layers=n_layers, out_dim=out_dim, max_seq_len=msl, device=0)
    _, graphs = ec.make_graphs(batch * n_batches, in_dim, 17, sessions.sequence_to_graph)
    batches = [graph.collate(graphs[i:i + batch]).to("cuda:0") for i in range(0, len(graphs), batch)]
    for b in batches[:2]:
        enc(b)
    torch.cuda.synchronize()
    for b in batches:
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter()
        e0.record()
        enc(b)
        e1.record()
        torch.cuda.synchronize()
        t1 = time.perf_counter()
        print("batch: %d query + %d product nodes, device %.3f ms, wall %.3f ms, %d launches" % (
            b['query'].x.shape[0], b['product'].x.shape[0], e0.elapsed_time(e1), (t1 - t0) * 1e3, enc.launches))
    # back to back: the pipeline's form (deferred flag check, no host sync per batch)
    reps = 5
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    t0 = time.perf_counter()
    for _ in range(reps):
        for b in batches:
            enc(b, defer_check=True)
    t1 = time.perf_counter()
    e1.record()
    torch.cuda.synchronize()
    enc.check_flags()
    ms = e0.elapsed_time(e1) / (reps * len(batches))
    print("back to back: %.3f ms per forward = %.0f sessions/s (host enqueue %.3f ms per forward)"
          % (ms, batch / ms * 1e3, (t1 - t0) * 1e3 / (reps * len(batches))))
    print("inside the C call: %.3f ms of host time per forward" % (enc.host_ns / 1e6))


if __name__ == "__main__":
    main()
