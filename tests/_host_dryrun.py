# Host-orchestration dry run with a FAKE backend (kernels replaced by argument-checking no-ops, CPU tensors posing as
# CUDA ones): catches Python-level errors in the kernel sequencing before any GPU time is spent. Run by test_host_dryrun.py.
import sys, types, ctypes
import os; _R = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, _R); sys.path.insert(0, os.path.join(_R, "mui-deepautoencoder_b200"))
import torch, numpy as np
torch.Tensor.is_cuda = property(lambda self: True)
from codae import _C
class FakeLib:
    def __getattr__(self, name):
        if name.endswith("workspace_bytes"):
            return lambda *a: 1 << 16
        if name == "codae_linear_engine":
            return lambda c, dtype, M, N, K: 1 if (dtype == 1 and N >= 32 and K >= 32) else 0   # csrc/linear.cu: tc_shape_ok
        if name == "codae_linear_wgrad_sq_slots":
            return lambda c, M, N, K, dtype: 7 if dtype == 1 else 0
        def f(*a):
            exp = _C.SIGNATURES[name][1]
            assert len(a) == len(exp), (name, len(a), len(exp))
            for v, t in zip(a, exp):
                if t is _C._vp:
                    assert v is None or isinstance(v, (ctypes.c_void_p, int)), (name, type(v))
                elif t in (_C._i, _C._i64, _C._sz, _C._u64):
                    assert isinstance(v, (int, np.integer)) and not isinstance(v, bool), (name, type(v), v)
                elif t in (_C._f, _C._d):
                    assert isinstance(v, (int, float)), (name, type(v))
            return 0
        return f
_C.lib = lambda: FakeLib()
_C.ctx = lambda device=None: ctypes.c_void_p(1)
_C.stream = lambda: ctypes.c_void_p(0)
_C._dev_check = lambda *a: None
orig_device = torch.device
import codae.model._flat_mlp as fm
def to(self, *a, **k):
    self._flatten(torch.device("cpu")); return self
fm.FlatMLP.to = to
import codae.tool.data_tool as dtl
dtl.Corrupter._cuda_device = lambda self: torch.device("cpu")
class _FakeStream:
    def __init__(self, *a, **k): pass
    def wait_event(self, e): pass
    def wait_stream(self, s): pass
    def __enter__(self): return self
    def __exit__(self, *a): return False
class _FakeEvent:
    def __init__(self, *a, **k): pass
    def record(self, s=None): pass
torch.cuda.Stream = _FakeStream
torch.cuda.Event = _FakeEvent
torch.cuda.current_stream = lambda *a, **k: _FakeStream()
torch.cuda.stream = lambda s: s
torch.cuda.synchronize = lambda *a: None
torch.cuda.current_device = lambda: 0

from codae.dataset import ConcatenatedEmbeddingDataset, MixedVariableDataset
from codae.model import EmbeddingDenoisingAutoencoder, MixedVariableDenoisingAutoencoder
from codae.tool import Corrupter, FusedStep, RankingLoss, CombinedCriterion
from codae.tool.inference import ComplementarityScorer, SwapScorer, predict_slot

dev = torch.device("cpu")
for dtype in ("fp32", "bf16"):
    ds = ConcatenatedEmbeddingDataset.from_tensors([torch.rand(64, 32) for _ in range(3)])
    m = EmbeddingDenoisingAutoencoder(96, 40, 32, 3, 3, False); m.set_compute_dtype(dtype); m.to(dev)
    cor = Corrupter(64, ds.arch, 2, dev)
    fs = FusedStep(m, cor, ds.data, 1e-3, 1e-4, clip=True)
    fs.step(torch.arange(8)); fs.step(torch.arange(5)); print(dtype, "launches", fs.kernel_launches, fs.last_loss(8))
    fs.evaluate(torch.arange(4)); print(fs.read_monitors())
    st = (torch.rand(8, 96), torch.zeros(8, cor.nb_run, dtype=torch.int16)); fs.step(None, staged=st)
    if dtype == "bf16":   # norm-free clipped step: one slot range per layer, Adam from the partials
        fq = FusedStep(m, cor, ds.data, 1e-3, 1e-4, clip=True, wgrad_sqnorm=True)
        fq.step(torch.arange(8)); fq.step(torch.arange(5)); fq.step(torch.arange(0))
        assert fq._bufs[8]["sq_partials"].numel() == 7 * len(m.dims) and fq.kernel_launches == 2
        fq.step(torch.arange(8)); print("wgrad_sqnorm launches", fq.kernel_launches)
    else:
        try:
            FusedStep(m, cor, ds.data, 1e-3, 1e-4, clip=True, wgrad_sqnorm=True); raise AssertionError("fp32 engine accepted wgrad_sqnorm")
        except RuntimeError:
            pass
    # legacy
    masks, fmask = cor.get_masks((1, 2, 3), 0)
    x = ds.data[[1, 2, 3]]
    y = m(m.corrupt(input_data=x, mask=fmask)); loss = torch.nn.MSELoss()(x, y); loss.backward()
    print("legacy ok", y.shape, m.linears()[0].weight.grad is not None, m.encode(x).shape, m.decode(torch.rand(3, 40)).shape)
    rl = RankingLoss(ds, list(range(20, 40)), dev)
    c1 = Corrupter(64, ds.arch, 1, dev); _, f1 = c1.get_masks((1, 2, 3), 0)
    f1 = torch.ones(3, 96); f1[0, :32] = 0; f1[1, 32:64] = 0; f1[2, 64:] = 0
    print("rank", rl.get(y.detach(), f1, (1, 2, 3)))
    p = predict_slot(m, x, 1, 32); sc = ComplementarityScorer(ds.data_per_category[1], 32, k=5); print(sc.topk(p)[1].shape)
    sw = SwapScorer(m, ds.data_per_category[1], 32, k=5, chunk=16); print("swap", sw.topk(x[0], 1)[1].shape)
# abalone
arch = [dict(name="Sex", size=3, type="classification", position=0)] + [dict(name=str(i), size=1, type="regression", position=3 + i) for i in range(8)]
ds = MixedVariableDataset.from_arch(arch, torch.rand(100, 11))
m = MixedVariableDenoisingAutoencoder(arch, 11, 4, dev, 2, 2, False); m.to(dev)
cor = Corrupter(100, arch, 3, dev)
ft = FusedStep(m, cor, ds.data, 1e-3, 1e-4, clip=True, tiny_mlp=True,
               mixed=dict(arch=arch, weight=[0.4] + [1] * 8, norm_scale=torch.rand(8), norm_min=torch.rand(8), norm_first=3))
ft.step(torch.arange(16), run=5); print("tiny mlp launches", ft.kernel_launches)      # corrupt, fwd, loss, monitor, bwd, counter, clip+Adam
assert ft.kernel_launches == 7
ft.evaluate(torch.arange(7), run=2)
fs = FusedStep(m, cor, ds.data, 1e-3, 1e-4, clip=True, mixed=dict(arch=arch, weight=[0.4] + [1] * 8, norm_scale=torch.rand(8), norm_min=torch.rand(8), norm_first=3))
fs.step(torch.arange(16), run=5); fs.evaluate(torch.arange(7), run=2); print({k: (v.shape if hasattr(v, "shape") else v) for k, v in fs.read_monitors().items()})
crit = CombinedCriterion(arch, 3, dev, torch.tensor([0, 0, 0] + [1] * 8), weight=[0.4] + [1] * 8, reduction="mean")
x = ds.data[:6]; masks, fmask = cor.get_masks(tuple(range(6)), 1)
y = m(m.corrupt(x, fmask)); l = crit(x=x, y=y); l.backward(); print("abalone legacy", float(l))
mc = CombinedCriterion(arch, 3, dev, torch.tensor([0, 0, 0] + [1] * 8), reduction="none"); print(mc(x, y.detach(), as_numpy=True).shape)
print("DRY RUN OK")
