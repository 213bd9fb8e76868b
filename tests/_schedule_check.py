# Data-race check of FusedStep's multi-stream schedules on the CPU (run by test_schedule_races.py in a subprocess: it
# monkeypatches torch.cuda).  Every C-ABI call becomes a node with the byte intervals it reads and writes; fake streams and
# events record exactly the ordering the real ones would impose (stream order, wait_event, wait_stream).  For every pair of
# calls that touch overlapping bytes with at least one write, the earlier one must HAPPEN BEFORE the later one through
# those edges -- otherwise the two kernels could run concurrently on the GPU and the schedule has a race.
# Three consecutive steps are issued per scenario so that hazards across step boundaries are covered as well.
# Programmatic dependent launch is modelled too: codae_linear_fwd / codae_linear_dgrad request their WEIGHT tiles before
# griddepcontrol.wait, i.e. as early as the previous kernel of their stream STARTS (event / stream waits issued in between are
# conservatively assumed not to hold that prefetch back) -- unless the launch was made a full dependency by the library's
# weights-written bookkeeping (csrc/common.cuh: codae_mark_weights_written / codae_pdl_allowed), which is replayed here.
import ctypes
import os
import sys

_R = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, _R)
sys.path.insert(0, os.path.join(_R, "mui-deepautoencoder_b200"))
import torch  # noqa: E402

torch.Tensor.is_cuda = property(lambda self: True)
from codae import _C  # noqa: E402

NODES = []          # dict(name, stream, reads, writes, preds)


class FakeStream:
    count = 0

    def __init__(self, *a, **k):
        self.last = None
        self.last_kernel = None
        FakeStream.count += 1
        self.name = "s%d" % FakeStream.count
        self.cuda_stream = FakeStream.count

    def _node(self, name, reads=(), writes=(), extra=()):
        preds = set(extra)
        if self.last is not None:
            preds.add(self.last)
        NODES.append(dict(name=name, stream=self.name, reads=list(reads), writes=list(writes), preds=preds))
        self.last = len(NODES) - 1
        return self.last

    def wait_event(self, e):
        if e.node is not None:
            self._node("wait_event", extra=[e.node])

    def wait_stream(self, s):
        if s.last is not None:
            self._node("wait_stream", extra=[s.last])

    def synchronize(self):
        pass


class FakeEvent:
    def __init__(self, *a, **k):
        self.node = None

    def record(self, stream=None):
        s = stream if stream is not None else CUR[-1]
        self.node = s.last


MAIN = FakeStream()
CUR = [MAIN]


class _Use:
    def __init__(self, s):
        self.s = s

    def __enter__(self):
        CUR.append(self.s)
        return self.s

    def __exit__(self, *a):
        CUR.pop()
        return False


torch.cuda.Stream = FakeStream
torch.cuda.Event = FakeEvent
torch.cuda.current_stream = lambda *a, **k: CUR[-1]
torch.cuda.stream = lambda s: _Use(s)
torch.cuda.synchronize = lambda *a: None
torch.cuda.current_device = lambda: 0


def span(t):
    """Byte interval [lo, hi) a (possibly strided) tensor view can touch."""
    if t is None:
        return None
    es = t.element_size()
    lo = t.untyped_storage().data_ptr() + t.storage_offset() * es
    ext = 1
    for n, st in zip(t.shape, t.stride()):
        if n == 0:
            return None
        ext += (n - 1) * st
    return (lo, lo + ext * es)


META = {}                    # extra attributes for the next node (critical-path estimate only)
DIRTY = [False, None]        # codae_ctx::weights_dirty, dirty_stream
WHOLE = [0]                  # number of parameters of the current model (a whole-buffer update vs a per-layer one)
USE_MARKS = [True]           # False: ignore explicit codae_weights_written calls (mutation test)


def mark_written(stream=None):
    DIRTY[0], DIRTY[1] = True, (stream or CUR[-1])


def pdl_allowed():
    if DIRTY[0] and DIRTY[1] is CUR[-1]:
        DIRTY[0] = False
        return False
    return True


def op(name, reads, writes, pdl=False, prefetch=None, writes_weights=False):
    """pdl: the entry point launches through launch_pdl / the programmatic attribute (consumes the dirty flag);
    prefetch: weight operand requested ahead of griddepcontrol.wait when the launch is a programmatic dependent."""
    st = CUR[-1]
    programmatic = pdl_allowed() if pdl else False
    extra = []
    if programmatic and prefetch is not None and st.last_kernel is not None:
        # may begin as soon as the previous KERNEL of this stream begins: ordered only after that kernel's own predecessors;
        # it is over when the kernel that consumes the tiles is over (the kernel node depends on it)
        NODES.append(dict(name=name + ":weight-prefetch", stream=st.name, reads=[span(prefetch)], writes=[],
                          preds=set(NODES[st.last_kernel]["preds"])))
        extra = [len(NODES) - 1]
    st._node(name, [s for s in map(span, reads) if s], [s for s in map(span, writes) if s], extra)
    NODES[st.last].update(META)
    META.clear()
    st.last_kernel = st.last
    if writes_weights:
        mark_written(st)


# ---- the C ABI wrappers, as access declarations (argument order: codae/_C.py) ------------------------------------------
def corrupt_fwd(data, batch_idx, B, mask_table, run, mask_bits, col_var, io, out_cx, out_x=None, out_mask_id=None):
    op("corrupt_fwd", [data, batch_idx, mask_table, mask_bits, col_var], [out_cx, out_x, out_mask_id], pdl=True)


def linear_fwd(X, W, bias, Y, M, N, K, act, dtype):
    op("linear_fwd", [X[:M, :K], W[:N, :K], bias], [Y[:M, :N]], pdl=True, prefetch=W[:N, :K] if dtype == 1 else None)


def linear_dgrad(dY, W, A_prev, dX, M, N, K, dtype):
    op("linear_dgrad", [dY[:M, :N], W[:N, :K], None if A_prev is None else A_prev[:M, :K]], [dX[:M, :K]], pdl=True,
       prefetch=W[:N, :K] if dtype == 1 else None)


def linear_wgrad(dY, X, dW, db, M, N, K, dtype):
    op("linear_wgrad", [dY[:M, :N], X[:M, :K]], [dW[:N, :K], db], pdl=True)


def linear_wgrad_sq(dY, X, dW, M, N, K, dtype, sq):
    op("linear_wgrad_sq", [dY[:M, :N], X[:M, :K]], [dW[:N, :K], sq], pdl=True)


def mse_loss_fwd_bwd(x, batch_idx, y, mask_id, mask_bits, col_var, B, io, grad_scale, dy, acc, ws):
    op("mse_loss_fwd_bwd", [x, batch_idx, y, mask_id, acc], [dy, acc, ws], pdl=True)


def counter_add(counter, delta):
    op("counter_add", [counter], [counter], pdl=True)


def grad_sqnorm(g, out, ws):
    op("grad_sqnorm", [g], [out, ws], pdl=True)


def adam_step(pf, g, m, v, p_bf16, lr, beta1, beta2, eps, wd, step, max_norm, sqnorm, grad_scale, step_dev=None):
    META["whole"] = pf.numel() == WHOLE[0]
    op("adam_step", [pf, g, m, v, sqnorm, step_dev], [pf, m, v, p_bf16], writes_weights=True)


def adam_step_partials(pf, g, m, v, p_bf16, lr, beta1, beta2, eps, wd, step, max_norm, sq_partials, sqnorm_out, grad_scale, step_dev=None):
    META["whole"] = pf.numel() == WHOLE[0]
    op("adam_step_partials", [pf, g, m, v, sq_partials, step_dev], [pf, m, v, p_bf16, sqnorm_out], writes_weights=True)


def clip_adam_step(pf, g, m, v, p_bf16, lr, beta1, beta2, eps, wd, step, max_norm, sqnorm_out, ws, grad_scale, step_dev=None):
    META["whole"] = pf.numel() == WHOLE[0]
    op("clip_adam_step", [pf, g, m, v, step_dev], [pf, m, v, p_bf16, sqnorm_out, ws], writes_weights=True)


def dp_adam_step(peers, pf, m, v, w_dtype, lr, beta1, beta2, eps, wd, step, max_norm, sqnorm_out, ws, grad_scale, step_dev=None):
    # reads every rank's gradients (this rank's buffer stands for all of them), rewrites its own shard in place, writes the
    # weight buffer of every rank (the local one is PEER_BUFS["w"]) and the local master / moments
    META["whole"] = True
    op("dp_adam_step", [PEER_BUFS["g"], pf, m, v, step_dev], [PEER_BUFS["g"], pf, m, v, PEER_BUFS["w"], sqnorm_out, ws], writes_weights=True)


PEER_BUFS = {}


def cast_bf16(src, dst):
    op("cast_bf16", [src], [dst], writes_weights=True)


for _n in ("corrupt_fwd", "linear_fwd", "linear_dgrad", "linear_wgrad", "linear_wgrad_sq", "mse_loss_fwd_bwd", "counter_add",
           "grad_sqnorm", "adam_step", "adam_step_partials", "clip_adam_step", "cast_bf16", "dp_adam_step"):
    setattr(_C, _n, globals()[_n])
_C.linear_engine = lambda device, dtype, M, N, K: 1 if (dtype == 1 and N >= 32 and K >= 32) else 0
_C.linear_wgrad_sq_slots = lambda device, M, N, K, dtype: 7 if dtype == 1 else 0
_C.sqnorm_workspace = lambda device: torch.zeros(64, dtype=torch.uint8)
_C.loss_workspace = lambda device: torch.zeros(64, dtype=torch.uint8)
_C.ctx = lambda device=None: ctypes.c_void_p(1)
_C.weights_written = lambda device: mark_written() if USE_MARKS[0] else None

import codae.model._flat_mlp as fm  # noqa: E402


def _to(self, *a, **k):
    self._flatten(torch.device("cpu"))
    return self


fm.FlatMLP.to = _to
import codae.tool.data_tool as dtl  # noqa: E402

dtl.Corrupter._cuda_device = lambda self: torch.device("cpu")
_C.mask_table_philox = lambda seed, first, n, nb_run, device: torch.zeros((n, nb_run), dtype=torch.int16)

import torch.distributed as dist  # noqa: E402


def _all_reduce(t, op=None, group=None):
    globals()["op"]("all_reduce", [t], [t])


dist.all_reduce = _all_reduce


def _all_gather_into_tensor(out, inp, group=None):
    globals()["op"]("all_gather", [inp], [out])


dist.all_gather_into_tensor = _all_gather_into_tensor
dist.get_rank = lambda group=None: 0
dist.barrier = lambda group=None: None


class _FakeGroup:
    WORLD = object()


dist.group = _FakeGroup
import types  # noqa: E402

_symm = types.ModuleType("torch.distributed._symmetric_memory")
_symm_allocs = []


def _symm_empty(n, dtype=None, device=None):
    t = torch.zeros(n, dtype=dtype)
    _symm_allocs.append(t)
    if len(_symm_allocs) % 3 == 1:
        PEER_BUFS["g"] = t
    elif len(_symm_allocs) % 3 == 2:
        PEER_BUFS["w"] = t
    return t


class _Handle:
    def __init__(self, t):
        self.buffer_ptrs = [t.data_ptr(), t.data_ptr() + 0]


_symm.empty = _symm_empty
_symm.rendezvous = lambda t, group: _Handle(t)
sys.modules["torch.distributed._symmetric_memory"] = _symm
dist._symmetric_memory = _symm
_C.dp_shard_elems = lambda n, world: ((n + world - 1) // world + 7) // 8 * 8
_C.dp_workspace = lambda device: torch.zeros(64, dtype=torch.uint8)
_C.dp_peers = lambda world, rank, g, w, s, **kw: (world, rank, g, w, s)

from codae.dataset import ConcatenatedEmbeddingDataset  # noqa: E402
from codae.model import EmbeddingDenoisingAutoencoder  # noqa: E402
from codae.tool import Corrupter, FusedStep  # noqa: E402


def overlap(a, b):
    return a[0] < b[1] and b[0] < a[1]


def check(tag):
    n = len(NODES)
    reach = [0] * n            # bitset of ancestors
    for j in range(n):
        r = 0
        for i in NODES[j]["preds"]:
            r |= reach[i] | (1 << i)
        reach[j] = r
    races = []
    for j in range(n):
        nj = NODES[j]
        if not (nj["reads"] or nj["writes"]):
            continue
        for i in range(j):
            ni = NODES[i]
            if (reach[j] >> i) & 1:
                continue
            hit = any(overlap(w, x) for w in ni["writes"] for x in nj["reads"] + nj["writes"]) or \
                any(overlap(r_, w) for r_ in ni["reads"] for w in nj["writes"])
            if hit:
                races.append((i, ni["name"], ni["stream"], j, nj["name"], nj["stream"]))
    kernels = sum(1 for x in NODES if x["reads"] or x["writes"])
    streams = len({x["stream"] for x in NODES})
    print("%-34s %3d calls on %d streams, %d unordered conflicting pairs" % (tag, kernels, streams, len(races)))
    for r in races[:8]:
        print("   RACE: #%d %s (%s)  <->  #%d %s (%s)" % r)
    return len(races)


# Optional: critical path of ONE steady-state step through the recorded DAG, with the per-launch durations measured on a B200
# for the embedding.yaml shapes (profiles/r01_bench_embedding_bf16_final.json; per-layer optimizer launches = 1/10 of the
# one-launch kernel + 1.5 us).  No resource contention, no launch gaps: a lower bound that ranks schedules, not a prediction.
DUR = {"corrupt_fwd": 2.8, "linear_fwd": 7.2, "linear_dgrad": 7.5, "linear_wgrad": 6.4, "linear_wgrad_sq": 6.9,
       "mse_loss_fwd_bwd": 5.7, "counter_add": 1.0, "grad_sqnorm": 20.0, "all_reduce": 80.0}
FULL_UPDATE = {"adam_step": 110.0, "adam_step_partials": 124.0, "clip_adam_step": 138.0, "dp_adam_step": 100.0}


def critical_path(n_layers, first, last):
    """Longest path (us) through nodes [first, last) given that everything before `first` is done at time 0."""
    end = {}
    for j in range(first, last):
        nd = NODES[j]
        name = nd["name"].split(":")[0]
        if nd["name"].endswith(":weight-prefetch") or not (nd["reads"] or nd["writes"]):
            d = 0.0
        elif name in FULL_UPDATE:
            whole = sum(b - a for a, b in nd["writes"][:1]) > 0 and nd.get("whole", False)
            d = FULL_UPDATE[name] if nd.get("whole") else FULL_UPDATE[name] / n_layers + 1.5
        else:
            d = DUR.get(name, 1.0)
        start = max([end.get(i, 0.0) for i in nd["preds"] if i >= first] + [0.0])
        end[j] = start + d
    return max(end.values()) if end else 0.0


def scenario(tag, dtype="bf16", clip=True, world=1, steps=3, batches=(8, 8, 8), expect=None, **kw):
    del NODES[:]
    MAIN.last = MAIN.last_kernel = None
    DIRTY[0], DIRTY[1] = False, None
    ds = ConcatenatedEmbeddingDataset.from_tensors([torch.rand(64, 32) for _ in range(3)])
    m = EmbeddingDenoisingAutoencoder(96, 96, 32, 4, 4, False)       # 10 Linear layers, like config/embedding.yaml
    m.set_compute_dtype(dtype)
    m.to(torch.device("cpu"))
    cor = Corrupter(64, ds.arch, 2, torch.device("cpu"))
    fs = FusedStep(m, cor, ds.data, 1e-3, 1e-4, clip=clip, world_size=world, **kw)
    WHOLE[0] = m.flat.numel()
    marks = []
    for B in batches:
        marks.append(len(NODES))
        fs.step(torch.arange(B))
    marks.append(len(NODES))
    if expect is not None:
        # the launches of one steady-state step, per stream in issue order (guards the default schedule against accidental edits)
        got = {}
        for nd in NODES[marks[-2]:marks[-1]]:
            if (nd["reads"] or nd["writes"]) and not nd["name"].endswith(":weight-prefetch"):
                got.setdefault(nd["stream"], []).append(nd["name"])
        seqs = sorted(got.values(), key=len, reverse=True)
        assert seqs == expect, (tag, seqs)
    if os.environ.get("CODAE_SCHEDULE_ESTIMATE") == "1" and len(set(batches)) == 1:
        # steady state: the last of the identical steps, everything issued before it taken as complete
        print("%-34s critical path of a steady-state step: %6.1f us (no contention, no launch gaps)"
              % (tag, critical_path(len(m.dims), marks[-2], marks[-1])))
    fs.evaluate(torch.arange(4))
    fs.step(torch.arange(batches[0]))
    fs.flush()
    return check(tag)


bad = 0
MAIN_DEFAULT = ["corrupt_fwd"] + ["linear_fwd"] * 10 + ["mse_loss_fwd_bwd"] + ["linear_dgrad"] * 9
bad += scenario("default (norm-free update)", expect=[MAIN_DEFAULT + ["counter_add", "adam_step_partials"], ["linear_wgrad_sq"] * 10])
bad += scenario("cooperative clip+Adam", wgrad_sqnorm=False,
                expect=[MAIN_DEFAULT + ["counter_add", "clip_adam_step"], ["linear_wgrad"] * 10])
bad += scenario("separate norm + Adam kernels", wgrad_sqnorm=False, fused_clip_adam=False)
bad += scenario("fp32 engine (one stream)", dtype="fp32")
bad += scenario("un-clipped", clip=False)
bad += scenario("data parallel, peer kernel", world=2, dp_mode="peer",
                expect=[MAIN_DEFAULT + ["counter_add", "dp_adam_step"], ["linear_wgrad"] * 10])
bad += scenario("data parallel, peer kernel, fp32 engine", world=2, dp_mode="peer", dtype="fp32")
bad += scenario("data parallel, overlapped all-reduce", world=2, dp_mode="nccl",
                expect=[MAIN_DEFAULT + ["counter_add", "clip_adam_step"], ["linear_wgrad"] * 10, ["all_reduce"]])
bad += scenario("data parallel, one all-reduce", world=2, dp_mode="nccl", overlap_allreduce=False)

# the checker itself: a schedule with a known race must be flagged
del NODES[:]
MAIN.last = MAIN.last_kernel = None
side = FakeStream()
a, b_ = torch.zeros(16), torch.zeros(16)
op("writer", [], [a])
with torch.cuda.stream(side):
    op("reader without a wait", [a], [b_])
assert check("self-test: missing wait_event") == 1
del NODES[:]
MAIN.last = MAIN.last_kernel = None
op("writer", [], [a])
ev = FakeEvent()
ev.record(MAIN)
side.wait_event(ev)
with torch.cuda.stream(side):
    op("reader after wait_event", [a], [b_])
assert check("self-test: with wait_event") == 0
print("SCHEDULES OK" if bad == 0 else "SCHEDULE RACES: %d" % bad)
sys.exit(0 if bad == 0 else 1)
