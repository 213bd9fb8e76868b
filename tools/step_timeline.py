#!/usr/bin/env python
"""Where does the time of one training step go?  Brackets every C-ABI call of FusedStep with one-thread %globaltimer
stamp kernels (codae_debug_stamp) on the stream the call is issued on, captures the instrumented step into a CUDA graph,
replays it, and prints a per-call timeline (start / end in us since the first stamp, stream, duration) plus the gaps.
The stamps are ordinary stream work: they serialise programmatic dependent launches and add ~1-2 us each, so read the
output for structure (what overlaps what, which stream waits for which), not for absolute step time.

    python tools/step_timeline.py [--workload embedding|modanet|polyvore] [--chain] [--replays 5]
"""
import argparse
import ctypes
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "mui-deepautoencoder_b200"))
import torch  # noqa: E402

import bench  # noqa: E402
from codae import _C  # noqa: E402
from codae.model import EmbeddingDenoisingAutoencoder  # noqa: E402
from codae.tool import Corrupter, FusedStep  # noqa: E402

CALLS = ["corrupt_fwd", "linear_fwd", "linear_chain", "mse_loss_fwd_bwd", "linear_wgrad", "linear_wgrad_sq", "linear_dgrad",
         "grad_sqnorm", "counter_add", "adam_step", "adam_step_partials", "clip_adam_step"]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="embedding", choices=sorted(bench.WORKLOADS))
    ap.add_argument("--chain", action="store_true")
    ap.add_argument("--replays", type=int, default=5)
    args = ap.parse_args()
    w = bench.WORKLOADS[args.workload]
    dev = torch.device("cuda", 0)
    io = w["S"] * w["E"]
    data = bench.synthetic_rows(min(w["N"], 32768), io, w["seed"], dev)
    arch = [dict(name=str(i), size=w["E"], type="regression", position=i * w["E"]) for i in range(w["S"])]
    model = EmbeddingDenoisingAutoencoder(io, w["z"], w["E"], w["nin"], w["nout"], False)
    model.set_compute_dtype(w["dtype"])
    model.to(dev)
    cor = Corrupter(data.shape[0], arch, w["k_max"], dev, seed=w["seed"])
    fs = FusedStep(model, cor, data, lr=w["lr"], weight_decay=w["wd"], clip=w["clip"], use_graph=False,
                   chain_forward=True if args.chain else None, chain_backward=True if args.chain else None)
    idx = torch.randint(0, data.shape[0], (w["B"],), device=dev)
    for _ in range(3):
        fs.step(idx)
    torch.cuda.synchronize()

    lib = _C.lib()
    lib.codae_debug_stamp.argtypes = [ctypes.c_void_p, ctypes.c_void_p]
    stamps = torch.zeros(4096, dtype=torch.int64, device=dev)
    events = []          # (name, stream id, slot of the start stamp)

    def stamp(slot):
        lib.codae_debug_stamp(ctypes.c_void_p(stamps.data_ptr() + 8 * slot), _C.stream())

    orig = {n: getattr(_C, n) for n in CALLS}

    def wrap(n):
        def f(*a, **k):
            slot = 2 * len(events)
            events.append((n, torch.cuda.current_stream().cuda_stream, slot))
            stamp(slot)
            r = orig[n](*a, **k)
            stamp(slot + 1)
            return r
        return f

    try:
        for n in CALLS:
            setattr(_C, n, wrap(n))
        fs.step(idx)                                   # eager, instrumented: records the call list
        torch.cuda.synchronize()
        n_calls = len(events)
        del events[:]
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            fs._enqueue(w["B"], fs._buffers(w["B"]), 0, w["B"], fs.data, fs._buffers(w["B"])["idx"], cor.device_tables()[0])
    finally:
        for n in CALLS:
            setattr(_C, n, orig[n])
    assert len(events) == n_calls
    for _ in range(args.replays):
        g.replay()
    torch.cuda.synchronize()
    t = stamps.cpu().tolist()
    t0 = min(t[e[2]] for e in events)
    streams = {}
    print("%-20s %-6s %10s %10s %9s" % ("call", "stream", "start us", "end us", "dur us"))
    last_end = {}
    for name, sid, slot in events:
        tag = streams.setdefault(sid, "s%d" % len(streams))
        a, b = (t[slot] - t0) / 1e3, (t[slot + 1] - t0) / 1e3
        gap = a - last_end.get(tag, a)
        last_end[tag] = b
        print("%-20s %-6s %10.2f %10.2f %9.2f   (gap on its stream %.2f)" % (name, tag, a, b, b - a, gap))
    end = max(t[e[2] + 1] for e in events)
    print("instrumented step: %.1f us from the first to the last stamp; %d calls" % ((end - t0) / 1e3, len(events)))


if __name__ == "__main__":
    main()
