"""Flat-buffer MLP shared by the two denoising autoencoders.

The nn.Linear / nn.ReLU modules only carry the structure (same `state_dict` keys and `print(model)` as
the reference); on a CUDA device every Parameter is re-pointed into ONE flat fp32 buffer and all arithmetic goes
through libcodae_b200's C ABI.

Layout: layer l is stored as the AUGMENTED matrix W'[out, ld] = [ W | 0.. | b | 0.. ] with the bias in column
bcol = round_up(in, 8) and ld = round_up(bcol + 1, 64): rows start on 128-byte boundaries in the bf16 shadow, so TMA
boxes and epilogue stores touch whole 32-byte sectors (a pitch of in+8 measured 20-30 % slower at 4096 wide).  Activations carry a matching constant-1 column,
so  Y = X' . W'^T  adds the bias inside the contraction and  dW' = dY^T . X'  yields the bias gradient in column bcol:
no bias epilogue, no separate column-sum kernel.  `weight` / `bias` Parameters are strided views of W'.
"""
import torch

from codae import _C


def _round_up(x, m):
    return (x + m - 1) // m * m


class _MLPFunction(torch.autograd.Function):
    """Legacy-API bridge: `model(x)` + `loss.backward()` route through the CUDA kernels
    (reference call sites: embedding_denoising_autoencoder.py:166,183; train_dae_on_embedding.py:210)."""

    @staticmethod
    def forward(ctx, x, mlp, first, last, *params):
        acts = mlp.forward_layers(x, first, last)
        ctx.mlp, ctx.first, ctx.last, ctx.acts = mlp, first, last, acts
        out_w = mlp.dims[last - 1][1]
        return acts[-1][:, :out_w].float() if acts[-1].dtype != torch.float32 else acts[-1][:, :out_w]

    @staticmethod
    def backward(ctx, dy):
        mlp = ctx.mlp
        gflat = torch.zeros_like(mlp.flat)
        dx = mlp.backward_layers(ctx.acts, dy, ctx.first, ctx.last, gflat, need_dx=ctx.needs_input_grad[0])
        grads = []
        for l in range(ctx.first, ctx.last):
            grads.append(mlp.weight_view(gflat, l))
            grads.append(mlp.bias_view(gflat, l))
        return (dx, None, None, None) + tuple(grads)


class FlatMLP(torch.nn.Module):

    def _finish_init(self, dims, relu, nb_encoder):
        """dims: [(in, out)], relu: [bool] per Linear; encoder = first nb_encoder Linears."""
        self.dims = list(dims)
        self.relu = list(relu)
        self.nb_encoder = nb_encoder
        self.flat = None          # fp32 [P_padded] on the CUDA device
        self.flat_bf16 = None     # bf16 shadow (tensor-core engine)
        self.flat_x3 = None       # [3, P_padded] bf16 planes hi / mid / lo of flat (fp32-parity tensor-core engine)
        self.compute_dtype = "fp32"
        self._layout = None

    # ---- structure ---------------------------------------------------------------------------
    def linears(self):
        return [m for m in list(self.input_layer) + list(self.output_layer) if isinstance(m, torch.nn.Linear)]

    def layout(self):
        """[(w_off, ld, bcol)] per layer and the padded total; every row starts 16-byte aligned in both the fp32
        buffer and its bf16 shadow."""
        if self._layout is None:
            off, out = 0, []
            for (i, o) in self.dims:
                bcol = _round_up(i, 8)
                ld = _round_up(bcol + 1, 64)
                out.append((off, ld, bcol))
                off += o * ld
            self._layout = (out, off)
        return self._layout

    def aug_view(self, flat, l):
        """W'[out, bcol + 1] (pitch ld): the operand the GEMMs contract over (weights, zero pad, bias column).  `flat` may be
        the [3, n] plane shadow of the fp32-parity engine: the view is then [3, out, bcol + 1]."""
        (w_off, ld, bcol), (i, o) = self.layout()[0][l], self.dims[l]
        if flat.dim() == 2:
            return flat[:, w_off:w_off + o * ld].view(3, o, ld)[:, :, :bcol + 1]
        return flat[w_off:w_off + o * ld].view(o, ld)[:, :bcol + 1]

    def weight_view(self, flat, l):
        (w_off, ld, _), (i, o) = self.layout()[0][l], self.dims[l]
        if flat.dim() == 2:
            return flat[:, w_off:w_off + o * ld].view(3, o, ld)[:, :, :i]
        return flat[w_off:w_off + o * ld].view(o, ld)[:, :i]

    def bias_view(self, flat, l):
        (w_off, ld, bcol), (_, o) = self.layout()[0][l], self.dims[l]
        return flat[w_off:w_off + o * ld].view(o, ld)[:, bcol]

    @staticmethod
    def act_width(w):
        """Pitch of an activation buffer of logical width w: data, zero pad to 8, the constant-1 column, zero pad to a
        multiple of 64 elements (128-byte rows in bf16)."""
        return _round_up(_round_up(w, 8) + 1, 64)

    @staticmethod
    def new_activation(B, w, dtype, device, width=None):
        """dtype: torch.float32 / torch.bfloat16, or "x3" for the three bf16 planes of the fp32-parity engine ([3, B, pitch];
        the constant 1 is (1, 0, 0))."""
        if dtype == "x3":
            a = _C.new_x3((B, width or FlatMLP.act_width(w)), device)
            a[0, :, _round_up(w, 8)] = 1
            return a
        a = torch.zeros((B, width or FlatMLP.act_width(w)), dtype=dtype, device=device)
        a[:, _round_up(w, 8)] = 1
        return a

    def nb_parameters(self):
        return sum(i * o + o for i, o in self.dims)

    # ---- device placement ------------------------------------------------------------------------
    def to(self, *args, **kwargs):
        """Like the reference's override (embedding_denoising_autoencoder.py:214-223), then flatten."""
        self = super().to(*args, **kwargs)
        self.input_layer = self.input_layer.to(*args, **kwargs)
        self.output_layer = self.output_layer.to(*args, **kwargs)
        dev = next(self.parameters()).device
        if dev.type == "cuda":
            self._flatten(dev)
        return self

    def _flatten(self, dev):
        lay, total = self.layout()
        flat = torch.zeros(total, dtype=torch.float32, device=dev)
        for l, lin in enumerate(self.linears()):
            self.weight_view(flat, l).copy_(lin.weight.data)
            self.bias_view(flat, l).copy_(lin.bias.data)
            lin.weight.data = self.weight_view(flat, l)
            lin.bias.data = self.bias_view(flat, l)
        self.flat = flat
        self.flat_bf16 = None
        self.flat_x3 = None

    def rebind_flat(self, new_flat):
        """Moves the flat fp32 parameter buffer into `new_flat` (same layout; e.g. a symmetric-memory allocation the peers of a
        data-parallel run can write) and re-points every Parameter view."""
        new_flat.copy_(self.flat)
        for l, lin in enumerate(self.linears()):
            lin.weight.data = self.weight_view(new_flat, l)
            lin.bias.data = self.bias_view(new_flat, l)
        self.flat = new_flat

    def set_compute_dtype(self, name):
        """"fp32" (default, the reference's precision): tcgen05 tensor cores on bf16 triples of the fp32 values (six MMAs per
        k-step, fp32 accumulation; tracks torch's fp32 nn.Linear to fp32 rounding) for every width the tensor-core path can tile;
        tabular widths (abalone) run the exact-fp32 FFMA engine.  "bf16": tcgen05 tensor cores on bf16 operands with fp32
        accumulation and fp32 master weights (1e-2 mode).  "fp32_simt": the FFMA engine regardless of width (A/B and tests).
        Nothing falls back silently: the engine is decided here, up front, from the layer widths."""
        if name not in ("fp32", "bf16", "fp32_simt"):
            raise Exception("Unknown compute dtype.")
        self.compute_dtype = name
        return self

    def engine_dtype(self):
        """_C.BF16 / _C.F32X3 (tensor cores) or _C.F32 (FFMA)."""
        if self.flat is not None and self.compute_dtype in ("bf16", "fp32"):
            want = _C.BF16 if self.compute_dtype == "bf16" else _C.F32X3
            eng = _C.ENGINE_TCGEN05_BF16 if want == _C.BF16 else _C.ENGINE_TCGEN05_F32X3
            if all(_C.linear_engine(self.flat.device, want, 128, o, i) == eng for i, o in self.dims):
                return want
        return _C.F32

    @staticmethod
    def act_dtype(eng):
        """Activation buffer type of an engine for new_activation()."""
        return {_C.BF16: torch.bfloat16, _C.F32X3: "x3"}.get(eng, torch.float32)

    def refresh_shadow(self):
        """Rewrites the copy of the weights the tensor-core GEMMs read from the fp32 master (after a load_state_dict, an
        external optimizer step, ...); the fused optimizer kernels keep it current by themselves."""
        eng = self.engine_dtype()
        if eng == _C.BF16:
            if self.flat_bf16 is None:
                self.flat_bf16 = torch.empty(self.flat.numel(), dtype=torch.bfloat16, device=self.flat.device)
            _C.cast_bf16(self.flat, self.flat_bf16)
        elif eng == _C.F32X3:
            if self.flat_x3 is None:
                self.flat_x3 = torch.empty((3, self.flat.numel()), dtype=torch.bfloat16, device=self.flat.device)
            _C.split_x3(self.flat, self.flat_x3)

    def gemm_weights(self, eng=None):
        """The buffer the GEMMs of engine `eng` read: flat (FFMA), flat_bf16 or flat_x3."""
        eng = self.engine_dtype() if eng is None else eng
        return {_C.BF16: self.flat_bf16, _C.F32X3: self.flat_x3}.get(eng, self.flat)

    def sync_weights(self):
        """A data-parallel trainer with a sharded update (FusedStep, dp_mode="peer") registers its flush() here: every entry
        point of the model that reads the fp32 master weights first gathers the other ranks' shards (a COLLECTIVE: call it on
        all ranks).  No-op otherwise."""
        hook = getattr(self, "_flush_hook", None)
        if hook is not None:
            hook()

    def state_dict(self, *args, **kwargs):
        self.sync_weights()
        return super().state_dict(*args, **kwargs)

    def load_state_dict(self, *args, **kwargs):
        """Loads into the flat buffer (the Parameters are views of it) and rewrites the copy the tensor-core GEMMs read."""
        out = super().load_state_dict(*args, **kwargs)
        if self.flat is not None and (self.flat_bf16 is not None or self.flat_x3 is not None):
            self.refresh_shadow()
        return out

    def _require_cuda(self):
        self.sync_weights()
        if self.flat is None:
            raise RuntimeError("codae: the model must be moved to a CUDA device (model.to(device)); "
                               "the B200 path has no CPU fallback")

    # ---- compute -----------------------------------------------------------------------------------
    def forward_layers(self, x, first, last, acts=None):
        """Runs Linear[first:last] (+ReLU).  Returns the activation list [input, out_first, ...]; buffers are
        [B, round_up(width, 8)] with zero padding.  The last Linear of the chain writes fp32."""
        self._require_cuda()
        eng = self.engine_dtype()
        B = x.shape[0]
        dev = self.flat.device
        adt = self.act_dtype(eng)
        self.refresh_shadow()
        wflat = self.gemm_weights(eng)
        in_w = self.dims[first][0]
        if eng == _C.F32X3:
            t = self.new_activation(B, in_w, torch.float32, dev)
            t[:, :in_w] = x
            a0 = _C.new_x3(t.shape, dev)
            _C.split_x3(t, a0)
        else:
            a0 = self.new_activation(B, in_w, adt, dev)
            a0[:, :in_w] = x
        out = [a0]
        for l in range(first, last):
            i, o = self.dims[l]
            is_last = l == last - 1
            y = self.new_activation(B, o, torch.float32 if is_last else adt, dev)
            _C.linear_fwd(out[-1], self.aug_view(wflat, l), None, y, B, o, _round_up(i, 8) + 1,
                          _C.ACT_RELU if self.relu[l] else _C.ACT_NONE, eng)
            out.append(y)
        return out

    def backward_layers(self, acts, dy, first, last, gflat, need_dx=False):
        """dW/db of Linear[first:last] into `gflat` (flat layout) and optionally dL/dx."""
        eng = self.engine_dtype()
        dev = self.flat.device
        adt = self.act_dtype(eng)
        wflat = self.gemm_weights(eng)
        B = dy.shape[0]
        o_last = self.dims[last - 1][1]
        if eng == _C.F32X3:
            t = torch.zeros((B, _round_up(o_last, 8)), dtype=torch.float32, device=dev)
            t[:, :o_last] = dy
            g = _C.new_x3(t.shape, dev)
            _C.split_x3(t, g)
        else:
            g = torch.zeros((B, _round_up(o_last, 8)), dtype=adt, device=dev)
            g[:, :o_last] = dy
        # the last Linear of a chain may itself be followed by ReLU (encode() alone never is; decode() neither)
        dx = None
        for l in range(last - 1, first - 1, -1):
            i, o = self.dims[l]
            a_in = acts[l - first]
            if eng != _C.F32X3 and a_in.dtype != adt:
                a_in = a_in.to(adt)
            _C.linear_wgrad(g, a_in, self.aug_view(gflat, l), None, B, o, _round_up(i, 8) + 1, eng)
            if l > first or need_dx:
                gp = _C.new_x3((B, _round_up(i, 8)), dev) if eng == _C.F32X3 else torch.zeros((B, _round_up(i, 8)), dtype=adt, device=dev)
                prev_relu = l > 0 and self.relu[l - 1] and l > first
                _C.linear_dgrad(g, self.weight_view(wflat, l), a_in if prev_relu else None, gp, B, o, i, eng)
                g = gp
                if l == first:
                    dx = _C.x3_to_f32(gp)[:, :i] if eng == _C.F32X3 else gp[:, :i].float()
        return dx

    def _run(self, x, first, last):
        self._require_cuda()
        if not x.is_cuda:
            raise RuntimeError("codae: input is not on a CUDA device; the B200 path has no CPU fallback")
        params = []
        for lin in self.linears()[first:last]:
            params += [lin.weight, lin.bias]
        return _MLPFunction.apply(x, self, first, last, *params)

    def forward(self, x):
        """y = decode(encode(x))  (embedding_denoising_autoencoder.py:137-151)."""
        return self._run(x, 0, len(self.dims))

    def encode(self, x):
        """(embedding_denoising_autoencoder.py:155-168)"""
        return self._run(x, 0, self.nb_encoder)

    def decode(self, z):
        """(embedding_denoising_autoencoder.py:171-185)"""
        return self._run(z, self.nb_encoder, len(self.dims))

    def corrupt_dense(self, input_data, mask):
        """input_data.clone() * mask as one kernel (embedding_denoising_autoencoder.py:239)."""
        if not input_data.is_cuda:
            raise RuntimeError("codae: input is not on a CUDA device; the B200 path has no CPU fallback")
        x = input_data.contiguous()
        m = mask.to(torch.float32).expand_as(x).contiguous()
        out = torch.empty_like(x)
        _C.mul_mask(x, m, out)
        return out

    def init_weight_general_rule(self, m):
        """xavier_uniform_ on every Linear (embedding_denoising_autoencoder.py:188-197)."""
        if m.__class__.__name__.find('Linear') != -1:
            torch.nn.init.xavier_uniform_(m.weight)

    def init_bias_zero(self, m):
        """zero bias (embedding_denoising_autoencoder.py:200-211)."""
        if m.__class__.__name__.find('Linear') != -1:
            m.bias.data.fill_(0)
