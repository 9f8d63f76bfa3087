"""Host-side execution engine of the accelerated forward path.

A model (models/agents.py) describes its forward once per (batch, height, width) as a straight-line *program* of
libw2c launches over preallocated NHWC buffers; the engine
  * folds BatchNorm + bias and packs conv weights once per weight version (setup-time, on device),
  * records every launch as a pre-bound ctypes call (no per-step Python tensor work),
  * optionally captures the whole program in a CUDA graph and replays it per step.
PyTorch supplies device memory, streams and graphs only.
"""
import ctypes
import gc
import os
import threading
import weakref

import torch

from . import _lib, ops

# "fp16": IEEE-half activations and weights on the same kernels - the speed and bytes of "bf16" with three more
# mantissa bits (logits ~2e-3 of the range from the fp32 reference instead of ~2e-2); needs activations < 65504
# "fp16x3": two fp16 planes (hi | lo), every product as hi*hi + hi*lo + lo*hi: the bf16x3 scheme on half operands
# "mixed": the fp16x3 storage with a per-layer pass plan (MIXED_ONE_PASS): the layers listed there run ONE pass
#          (plain fp16 arithmetic on the hi planes, a third of the MMA work), everything else three. The plan comes
#          from the per-layer error attribution in profiles/r2_precision_attribution.md: the policy net decides the
#          softmax weights and tolerates no rounding (fp16 there alone costs 6e-3 of the logit range), decoder layers
#          sit too close to the output; the wide middle of the feature encoder is where one pass is affordable within
#          the 1e-3 bound.
PRECISIONS = {"bf16": ops.ACT_BF16, "bf16x3": ops.ACT_BF16X2, "fp16": ops.ACT_FP16, "fp16x3": ops.ACT_FP16X2,
              "mixed": ops.ACT_FP16X2}
# stack -> 1-based layer numbers (conv<i> of n_segnet_encoder, backbone.py:19-39) that run one pass under "mixed".
#   "encoder_fused"  the value-map encoder of a model whose decoder input is an attention-weighted SUM of several
#                    agents' maps (MIMOcom, MIMOcomWho, LearnWhen2Com): the sum averages the per-map rounding noise
#   "encoder"        the value-map encoder whose map reaches the decoder directly / concatenated (Single_agent,
#                    All_agents, MIMO_All_agents, LearnWho2Com's own map): less headroom, fewer one-pass layers
# Measured on B200 (bench.py `parity`, tests/test_parity_gpu.py): each one-pass layer adds 2.5-4e-4 (in quadrature) to
# the ~1e-4 floor of the three-pass path; the layers listed are the ones with the most MMA time per unit of error
# (conv6: 0.99 ms per 40 frames for 0.66e-7 of squared error, conv9: 0.84 / 1.5e-7, conv8: 0.49 / 0.96e-7). Seven
# one-pass layers (conv1-3, 5, 6, 8, 9) measured 7.5e-4 on the 5-agent bench scene but 1.10e-3 on the 8-agent scene
# of the sharded runs: the plan keeps a factor of ~1.6 in hand instead.
MIXED_ONE_PASS = {"encoder_fused": (6, 8, 9), "encoder": (6, 9)}
BN_EPS_DEFAULT = 1e-5


def default_precision():
    p = os.environ.get("W2C_PRECISION", "bf16")
    if p not in PRECISIONS:
        raise ValueError("W2C_PRECISION must be one of %s (got %r)" % (sorted(PRECISIONS), p))
    return p


# Program.side_stream records the two encoder chains as parallel CUDA-graph branches. W2C_TWO_STREAMS = 1 always,
# 0 never, unset: only where the builder asks for it (auto=True). Measured on B200: no gain for the n_segnet pair
# (3399 / 3386 vs 3419 agent-frames/s serial - that step runs at the 1 kW power cap, where overlapping kernels cannot
# add throughput) but +7 % for the resnet18 pair, whose 55 launches of 20-90 us each leave SMs idle (12310 -> 13208
# agent-frames/s, two runs each): the models switch it on for resnet backbones.
# n_segnet steps of at most this many input pixels per rank also fork (measured: see profiles/r2_bench.md, B = 1 rows)
TWO_STREAM_MAX_PIXELS = int(os.environ.get("W2C_TWO_STREAM_MAX_PIXELS", str(8 * 512 * 512)))
_TS = os.environ.get("W2C_TWO_STREAMS")
TWO_STREAMS = None if _TS is None else _TS == "1"


# Fused encoder head (csrc/enc_head.cu: conv1 + conv2 of n_segnet_encoder in one kernel, the 64-channel full-resolution
# map stays in shared memory). bench.py --no-fused-ends / tests flip this for the A/B against the two-kernel path.
FUSE_ENCODER_HEAD = os.environ.get("W2C_FUSE_ENDS", "1") != "0"


# Experiment, off by default (W2C_GRAPH_COLLECTIVE=1): capture the one collective of a sharded forward
# (sharding.all_gather_slots) INSIDE the step's CUDA graph instead of splitting the program into two graphs around it.
# Measured on 2 x B200: the same throughput as the split (7123 vs 7129 agent-frames/s - the host gap is hidden by the
# asynchronous launch queue) and destroy_process_group() then hangs at exit, so the split stays.
CAPTURE_COLLECTIVES = os.environ.get("W2C_GRAPH_COLLECTIVE", "0") == "1"


# CUDA-graph capture is serialised across host threads (one model per device per thread is the nn.DataParallel shape,
# train.py:177) and runs in thread-local error mode: another thread's allocations or synchronisations would otherwise
# invalidate a capture in flight (seen on 2 GPUs: "The CUDA Graph is empty", stale outputs on replay).
_CAPTURE_LOCK = threading.Lock()


def use_graphs_default():
    return os.environ.get("W2C_CUDA_GRAPH", "1") != "0"


class ActMap:
    """Handle of an NHWC activation map living in (a channel slice of) a bf16 buffer."""
    __slots__ = ("buf", "n", "h", "w", "c", "cstride", "coffset")

    def __init__(self, buf, n, h, w, c, cstride=None, coffset=0):
        self.buf, self.n, self.h, self.w, self.c = buf, n, h, w, c
        self.cstride = cstride if cstride is not None else c
        self.coffset = coffset

    def slice(self, coffset, c):
        return ActMap(self.buf, self.n, self.h, self.w, c, self.cstride, self.coffset + coffset)

    def images(self, first, count):
        """Sub-range of images [first, first+count) as a view (agent-major layouts make agents contiguous)."""
        return ActMap(self.buf[first:first + count], count, self.h, self.w, self.c, self.cstride, self.coffset)


class PackedConv:
    __slots__ = ("w", "scale", "shift", "cin", "cout", "kind", "relu", "subsample")


def _stride2_view(pc):
    """The same packed conv with the stride-4 post-subsampling switched off (Program.conv runs it first)."""
    v = PackedConv()
    for k in PackedConv.__slots__:
        setattr(v, k, getattr(pc, k))
    v.subsample = 1
    return v


class WeightCache:
    """Device-side packed operands for one (model, device, precision). Rebuilt when the model invalidates it.

    live=prog (train mode): the operands are re-derived from the module's parameters INSIDE the program, every run -
    the pack / fold launches are recorded into `prog` ahead of the first launch that reads them - because an optimizer
    step changes the parameters in place between two forwards (trainer.py:668-670). The parameters must then live on
    the program's device as contiguous fp32 tensors (their storage is read by the recorded launches)."""

    def __init__(self, device, act, live=None):
        self.device = device
        self.act = act
        # (a weak reference: program -> cache -> program would make every train-mode program cyclic garbage, freed
        # only when the collector happens to run - e.g. in the middle of a later CUDA-graph capture, which the
        # destruction of the old program's graphs then invalidates)
        self.live = weakref.proxy(live) if live is not None else None
        self._convs = {}
        self._misc = {}

    def _param(self, t):
        """fp32 device tensor of a parameter: a copy at setup time, the parameter's own storage in live mode."""
        if t is None:
            return None
        if self.live is None:
            return t.detach().to(self.device, torch.float32).contiguous()
        d = t.detach()
        if d.device != self.device or d.dtype != torch.float32 or not d.is_contiguous():
            raise RuntimeError("train mode reads the parameters in place: they must be contiguous fp32 tensors on %s"
                               % self.device)
        return d

    def _emit(self, fn, *args):
        """Run a setup launch now (eval) or record it into the live program (train)."""
        if self.live is None:
            _lib.check(fn(*args, ops._stream()), getattr(fn, "__name__", "setup launch"))
        else:
            self.live._record(fn, *args)

    def _emit_torch(self, fn):
        if self.live is None:
            fn()
        else:
            self.live.calls.append((lambda _s, fn=fn: fn() or 0, None, self.live._sid))

    # conv / transposed conv (+ optional BatchNorm) -> packed bf16 weight + fp32 scale/shift
    def conv(self, conv, bn, relu, key=None):
        key = key or id(conv)
        pc = self._convs.get(key)
        if pc is not None:
            return pc
        transposed = isinstance(conv, torch.nn.ConvTranspose2d)
        w = self._param(conv.weight)
        cin_real = conv.in_channels
        cin = (cin_real + 63) // 64 * 64
        kh, kw = conv.kernel_size
        stride = conv.stride[0]
        if transposed:
            kind = ops.DECONV3X3_S2
        elif (kh, kw) == (3, 3):
            # stride 4 (img_encoder.squeezer with feat_squeezer=4, agent.py:51-52): out4[m] = sum_k x[4m + k - 1] is the
            # stride-2 conv at the even output positions, out2[2m]: run the stride-2 kernel, keep every other pixel
            kind = {1: ops.CONV3X3_S1, 2: ops.CONV3X3_S2, 4: ops.CONV3X3_S2}.get(stride)
        elif (kh, kw) == (1, 1):
            kind = {1: ops.CONV1X1_S1, 2: ops.CONV1X1_S2}.get(stride)
        else:
            kind = None
        if kind is None:
            raise NotImplementedError("no sm_100a kernel for conv k=%s stride=%s" % ((kh, kw), stride))
        pc = PackedConv()
        lib = _lib.load()
        ntaps = kh * kw
        pc.w = torch.empty(lib.w2c_packed_weight_bytes(conv.out_channels, cin, ntaps, self.act) // 2,
                           dtype=torch.bfloat16, device=self.device)
        self.keep = getattr(self, "keep", [])
        self.keep.append(w)
        self._emit(lib.w2c_pack_conv_weight, w.data_ptr(), conv.out_channels, cin_real, cin, ntaps, int(transposed),
                   self.act, pc.w.data_ptr())
        pc.cin, pc.cout, pc.kind, pc.relu = cin, conv.out_channels, kind, bool(relu)
        pc.subsample = 2 if (not transposed and stride == 4) else 1
        pc.scale, pc.shift = self._fold(conv, bn)
        self._convs[key] = pc
        return pc

    def _fold(self, conv, bn):
        cout = conv.out_channels
        scale = torch.empty(cout, dtype=torch.float32, device=self.device)
        shift = torch.empty(cout, dtype=torch.float32, device=self.device)
        ts = [self._param(conv.bias)]
        eps = BN_EPS_DEFAULT
        if bn is not None:
            ts += [self._param(t) for t in (bn.weight, bn.bias, bn.running_mean, bn.running_var)]
            eps = bn.eps
        else:
            ts += [None] * 4
        self.keep = getattr(self, "keep", [])
        self.keep.append(ts)
        p = lambda t: t.data_ptr() if t is not None else None
        self._emit(_lib.load().w2c_fold_bn, p(ts[0]), p(ts[1]), p(ts[2]), p(ts[3]), p(ts[4]), float(eps), cout,
                   scale.data_ptr(), shift.data_ptr())
        return scale, shift

    def dgrad(self, conv):
        """Operand of the DATA-gradient conv of `conv` (include/w2c.h, w2c_pack_conv_weight_ex): the same weight tensor
        re-indexed so that the forward tensor-core kernels compute dL/dx from dL/dy. Returns a PackedConv whose cin /
        cout are the gradient conv's (cin = the forward layer's padded cout)."""
        key = ("dgrad", id(conv))
        pc = self._convs.get(key)
        if pc is not None:
            return pc
        transposed = isinstance(conv, torch.nn.ConvTranspose2d)
        w = self._param(conv.weight)
        kh, kw = conv.kernel_size
        ntaps = kh * kw
        stride = conv.stride[0]
        fin, fout = conv.in_channels, conv.out_channels      # forward channel counts
        fout_pad = (fout + 63) // 64 * 64
        if fin % 64:
            raise NotImplementedError("data gradient of a conv with %d input channels" % fin)
        if transposed:
            kind, tr, flip = ops.CONV3X3_S2, 0, 0
        elif (kh, kw) == (3, 3) and stride == 1:
            kind, tr, flip = ops.CONV3X3_S1, 1, 1
        elif (kh, kw) == (3, 3) and stride == 2:
            kind, tr, flip = ops.DECONV3X3_S2, 1, 0
        elif (kh, kw) == (1, 1) and stride in (1, 2):
            kind, tr, flip = ops.CONV1X1_S1, 1, 0               # (stride 2: followed by the zero interleave)
        else:
            raise NotImplementedError("no data-gradient kernel for conv k=%s stride=%s" % ((kh, kw), stride))
        lib = _lib.load()
        pc = PackedConv()
        pc.w = torch.empty(lib.w2c_packed_weight_bytes(fin, fout_pad, ntaps, self.act) // 2, dtype=torch.bfloat16,
                           device=self.device)
        self.keep = getattr(self, "keep", [])
        self.keep.append(w)
        self._emit(lib.w2c_pack_conv_weight_ex, w.data_ptr(), fin, fout, fout_pad, ntaps, tr, flip, self.act,
                   pc.w.data_ptr())
        pc.cin, pc.cout, pc.kind, pc.relu, pc.subsample = fout_pad, fin, kind, False, 1
        pc.scale = torch.ones(fin, dtype=torch.float32, device=self.device)
        pc.shift = torch.zeros(fin, dtype=torch.float32, device=self.device)
        self._convs[key] = pc
        return pc

    # 3-input-channel stem: fp32 [cout][k] weight + scale/shift
    def conv_raw(self, conv):
        """The conv alone (scale = 1, shift = bias, no ReLU): what train-mode BatchNorm takes its statistics from."""
        return self.conv(conv, None, False, key=("raw", id(conv)))

    def stem_raw(self, conv):
        key = ("stem_raw", id(conv))
        st = self._misc.get(key)
        if st is None:
            w = self._param(conv.weight).reshape(conv.out_channels, -1)   # (a view: the parameter is contiguous)
            scale, shift = self._fold(conv, None)
            st = (w, scale, shift)
            self._misc[key] = st
        return st

    def stem_raw_pair(self, conv_a, conv_b):
        """Two raw first layers concatenated on the output-channel axis; refreshed per run in live mode."""
        key = ("stem_raw_pair", id(conv_a), id(conv_b))
        st = self._misc.get(key)
        if st is None:
            parts = [self.stem_raw(conv_a), self.stem_raw(conv_b)]
            st = tuple(torch.cat((parts[0][i], parts[1][i]), 0).contiguous() for i in range(3))

            def refresh(parts=parts, st=st):
                for i in range(3):
                    n0 = parts[0][i].shape[0]
                    st[i][:n0].copy_(parts[0][i])
                    st[i][n0:].copy_(parts[1][i])
            if self.live is not None:
                self._emit_torch(refresh)
            self._misc[key] = st
        return st

    def stem(self, conv, bn):
        key = ("stem", id(conv))
        st = self._misc.get(key)
        if st is None:
            w = self._param(conv.weight).reshape(conv.out_channels, -1).contiguous()
            scale, shift = self._fold(conv, bn)
            st = (w, scale, shift)
            self._misc[key] = st
        return st

    def stem_pair(self, conv_a, bn_a, conv_b, bn_b):
        """Two 3 -> 64 first layers concatenated on the output-channel axis (one fused 3 -> 128 stem)."""
        key = ("stem_pair", id(conv_a), id(conv_b))
        st = self._misc.get(key)
        if st is None:
            wa, sa, ha = self.stem(conv_a, bn_a)
            wb, sb, hb = self.stem(conv_b, bn_b)
            st = (torch.cat((wa, wb), 0).contiguous(), torch.cat((sa, sb)).contiguous(),
                  torch.cat((ha, hb)).contiguous())
            self._misc[key] = st
        return st

    # key/query head: fc.0 permuted to NHWC flatten order
    def mlp(self, fc, spatial):
        key = ("mlp", id(fc))
        m = self._misc.get(key)
        if m is None:
            f32 = self._param
            w0_src = self._param(fc[0].weight)
            n_feat = w0_src.shape[1]
            if n_feat != 256 * spatial * spatial:
                raise ValueError("key/query head expects %d input features, the policy map provides %d"
                                 % (n_feat, 256 * spatial * spatial))
            w0 = torch.empty((256, n_feat), dtype=torch.float32, device=self.device)

            def refresh(w0=w0, w0_src=w0_src):
                w0.view(256, spatial, spatial, 256).copy_(w0_src.view(256, 256, spatial, spatial).permute(0, 2, 3, 1))
            self._emit_torch(refresh)
            m = (w0, f32(fc[0].bias), f32(fc[2].weight), f32(fc[2].bias), f32(fc[4].weight), f32(fc[4].bias))
            self._misc[key] = m
        return m

    def tensor(self, t, key):
        v = self._misc.get(key)
        if v is None:
            v = self._param(t)
            self._misc[key] = v
        return v


class Program:
    """A recorded straight-line sequence of libw2c launches for one input shape."""

    def __init__(self, weights, device, act):
        self.weights = weights
        self.device = device
        self.act = act
        self.planes = ops.planes_of(act)
        self.calls = []      # (fn, args tuple, stream id) with the stream appended at run time; see side_stream()
        self._sid = 0        # stream id new calls are recorded on (0 = the caller's stream, 1 = the side stream)
        self._side = None    # torch.cuda.Stream of the side chain, created on first use
        self._capture_stream = None
        self.keep = []       # keeps ctypes structs / tensors alive
        self.graph = None
        self._batched = False   # _batch_setup_calls() ran
        self._tape_side = False  # a side-stream region is on the tape and not joined yet
        self.n_launches = 0  # kernel launches per run (counted from the library's counter)
        self._lib = _lib.load()
        # program-level I/O options (models/agents.py sets them before building):
        #   input_u8   the static input is the loader's raw uint8 RGB HWC frames [b, agents, h, w, 3]; lut = the
        #              loader-transform table (ops.loader_lut)
        #   labels     the decoder also writes a uint8 label map; logits=False drops the fp32 logits write
        self.input_u8 = False
        self.lut = None
        self.want_labels = False
        self.want_logits = True
        self.labels_out = None
        self.pass_plan = None   # {"stack": (layer numbers)} that run ONE MMA pass (two-plane formats; see PRECISIONS)
        self.train = False      # train-mode forward: BatchNorm from batch statistics + running-stat update (bn_train.cu)
        # grad: the train-mode forward also records its backward pass: every op appends a closure to
        # `tape`; finish_backward() replays the tape in reverse into `bprog`, a second Program of launches that turns
        # the gradient of the logits (copied into `dlogits` by the autograd bridge, models/agents.py) into parameter
        # gradients. Gradient maps mirror the forward buffers one to one (same shape, same views; see grad_map).
        self.grad = False
        self.tape = []
        self.bprog = None
        self.dlogits = None
        self.param_grads = []   # [(parameter, fn() -> gradient tensor in the parameter's layout)]
        self._grad_roots = {}
        self._f32_grads = {}
        self._grad_claimed = set()
        self._zero_list = []

    # ---- buffers
    def act_buf(self, n, h, w, c):
        t = torch.empty((n, h, w, self.planes * c), dtype=torch.bfloat16, device=self.device)
        self.keep.append(t)
        return ActMap(t, n, h, w, c)

    def f32_buf(self, *shape, dtype=torch.float32, zero=False):
        t = (torch.zeros if zero else torch.empty)(shape, dtype=dtype, device=self.device)
        self.keep.append(t)
        return t

    # ---- recorded launches
    def _record(self, fn, *args):
        self.calls.append((fn, args, self._sid))

    # ---- two concurrent chains
    # u_encoder and query_key_net are independent kernel chains between the stem and the attention (agent.py:1111 and
    # 1124 read the same views). Recorded on two streams they become parallel branches of the CUDA graph, so one
    # chain's CTAs fill the other's tails: persistent kernels with 4.3 waves of tiles, ~5 us ramp-up per launch and
    # the sub-wave layers at 16x16 otherwise leave SMs idle at every kernel boundary.
    _FORK, _JOIN = "fork", "join"
    _SIDE_BEGIN, _SIDE_END = "side-begin", "side-end"   # tape markers: the backward pass mirrors the two chains

    def side_stream(self, auto=False):
        """Context manager: calls recorded inside run on the side stream, which first waits for everything recorded
        so far. join() makes the main chain wait for them. auto: the builder's own preference, used when
        W2C_TWO_STREAMS is unset."""
        prog = self
        enabled = auto if TWO_STREAMS is None else TWO_STREAMS

        class _Ctx:
            def __enter__(self):
                if enabled:
                    prog.calls.append((Program._FORK, None, 0))
                    prog._sid = 1
                    if prog.grad:
                        prog.tape.append(Program._SIDE_BEGIN)
                        prog._tape_side = True
                return enabled

            def __exit__(self, *exc):
                prog._sid = 0
                if enabled and prog.grad:
                    prog.tape.append(Program._SIDE_END)
                return False

        return _Ctx()

    def join(self):
        # (a join without a preceding fork is a no-op at run time)
        self.calls.append((Program._JOIN, None, 0))
        if self.grad and self._tape_side:
            self.tape.append(Program._JOIN)
            self._tape_side = False

    def passes_for(self, stack, layer):
        """MMA passes of layer number `layer` of `stack` under this program's precision plan (0 = format default)."""
        if self.pass_plan and layer in self.pass_plan.get(stack, ()):
            return 1
        return 0

    def conv(self, x, pc, out=None, residual=None, nchw_out=None, block_n=0, labels=None, passes=0, bn_sums=None):
        """x: ActMap -> ActMap (or the fp32 NCHW tensor when nchw_out is given). labels: uint8 [n, h, w] tensor that
        receives the arg-max class of the fp32 NCHW logits (nchw_out may then be the string 'none': labels only).
        passes: MMA passes over the operand planes (0 = the format's default; 1 = hi*hi only). bn_sums: fp64 [2 * cout]
        tensor the conv epilogue adds sum(z) | sum(z^2) of its output to, where the launch can (w2c_conv_fuses_bn_sums);
        self.conv_fused_sums says afterwards whether it does."""
        self.conv_fused_sums = False
        if x.c != pc.cin:
            raise ValueError("conv expects %d input channels, got %d" % (pc.cin, x.c))
        if pc.subsample == 2:
            if nchw_out is not None or residual is not None or x.h % 4 or x.w % 4:
                raise ValueError("stride-4 conv: NHWC output, no residual, H and W divisible by 4")
            full = self.conv(x, _stride2_view(pc), passes=passes)       # (h/2, w/2) map
            if out is None:
                out = self.act_buf(x.n, x.h // 4, x.w // 4, pc.cout)
            elif (out.n, out.h, out.w, out.c) != (x.n, x.h // 4, x.w // 4, pc.cout):
                raise ValueError("conv (stride 4): the caller's output map does not match the layer's output")
            planes = self.planes

            def run(_stream, full=full, out=out):
                for pl in range(planes):
                    d = out.buf[..., pl * out.cstride + out.coffset: pl * out.cstride + out.coffset + out.c]
                    sv = full.buf[:, ::2, ::2, pl * full.cstride + full.coffset: pl * full.cstride + full.coffset + full.c]
                    d.copy_(sv)
                return 0
            self.calls.append((run, None, self._sid))
            return out
        if pc.kind in (ops.CONV3X3_S2, ops.CONV1X1_S2):
            ho, wo = x.h // 2, x.w // 2
        elif pc.kind == ops.DECONV3X3_S2:
            ho, wo = x.h * 2, x.w * 2
        else:
            ho, wo = x.h, x.w
        if isinstance(nchw_out, str):
            if labels is None:
                raise ValueError("conv: logits can only be dropped when a label map is written")
            y_ptr, out_fmt, ycs, yco, ret = None, ops.OUT_NCHW_F32, 0, 0, labels
        elif nchw_out is not None:
            y_ptr, out_fmt, ycs, yco, ret = nchw_out.data_ptr(), ops.OUT_NCHW_F32, 0, 0, nchw_out
        else:
            if out is None:
                out = self.act_buf(x.n, ho, wo, pc.cout)
            elif (out.n, out.h, out.w, out.c) != (x.n, ho, wo, pc.cout):
                raise ValueError("conv: the caller's output map is %dx%dx%dx%d but the layer produces %dx%dx%dx%d"
                                 % (out.n, out.h, out.w, out.c, x.n, ho, wo, pc.cout))
            y_ptr, out_fmt, ycs, yco, ret = out.buf.data_ptr(), ops.OUT_NHWC, out.cstride, out.coffset, out
        a = _lib.ConvArgs(x=x.buf.data_ptr(), w=pc.w.data_ptr(), scale=pc.scale.data_ptr(), shift=pc.shift.data_ptr(),
                          residual=residual.buf.data_ptr() if residual is not None else None, y=y_ptr, n=x.n,
                          h_in=x.h, w_in=x.w, cin=pc.cin, cout=pc.cout, x_cstride=x.cstride, x_coffset=x.coffset,
                          y_cstride=ycs, y_coffset=yco, kind=pc.kind, relu=int(pc.relu), act=self.act,
                          out_fmt=out_fmt, impl=ops.IMPL_TCGEN05, block_n=block_n,
                          labels=labels.data_ptr() if labels is not None else None, passes=passes)
        if bn_sums is not None and self._lib.w2c_conv_fuses_bn_sums(ctypes.byref(a)) == 1:
            a.bn_sums = bn_sums.data_ptr()
            self.conv_fused_sums = True
        self.keep.append(a)
        self._record(self._lib.w2c_conv_bnrelu_fwd, ctypes.byref(a))
        return ret

    # ---- conv + BatchNorm (+ residual) (+ ReLU) as the model sees it: folded in eval mode, batch statistics in train mode
    def conv_bn(self, x, conv, bn, relu, out=None, residual=None, nchw_out=None, labels=None, passes=0):
        if not self.train:
            return self.conv(x, self.weights.conv(conv, bn, relu), out=out, residual=residual, nchw_out=nchw_out,
                             labels=labels, passes=passes)
        if labels is not None or isinstance(nchw_out, str):
            raise ValueError("train mode produces logits, not label maps")
        if bn is None:
            # plain Conv2d (+ReLU) of simple_decoder.pred: bias and ReLU stay in the conv epilogue
            y = self.conv(x, self.weights.conv(conv, None, relu), out=out, residual=residual, nchw_out=nchw_out)
            if self.grad:
                if residual is not None:
                    raise NotImplementedError("backward of a residual conv without BatchNorm")
                self.tape.append(lambda: self._bwd_conv_unit(x, conv, None, relu, y, None, None, nchw=nchw_out is not None))
            return y
        pc = self.weights.conv_raw(conv)
        if nchw_out is not None:
            if not self.grad:
                z = self.conv(x, pc, nchw_out=nchw_out)
                self._bn_train_nchw(z, z, bn, relu, None)
                return z
            z = self.conv(x, pc, nchw_out=self.f32_buf(*nchw_out.shape))
            stats = self.f32_buf(2 * z.shape[1])
            self._bn_train_nchw(z, nchw_out, bn, relu, stats)
            self.tape.append(lambda: self._bwd_conv_unit(x, conv, bn, relu, nchw_out, z, stats, nchw=True))
            return nchw_out
        # the batch statistics come out of the conv epilogue where the launch can provide them (persistent kernel,
        # whole 64-channel groups, one storage plane); otherwise _bn_train runs its own pass over z
        ws = self._bn_ws(pc.cout)
        if not self.grad:
            z = self.conv(x, pc, out=out, bn_sums=ws[0])
            self._bn_train(z, z, bn, relu, residual, None, ws=ws, sums_ready=self.conv_fused_sums)
            return z
        z = self.conv(x, pc, bn_sums=ws[0])   # the raw conv output is kept: the backward pass normalises it again
        fused = self.conv_fused_sums
        y = out if out is not None else self.act_buf(z.n, z.h, z.w, z.c)
        if (y.n, y.h, y.w, y.c) != (z.n, z.h, z.w, z.c):
            raise ValueError("conv: the caller's output map does not match the layer's output")
        stats = self.f32_buf(2 * z.c)
        ss = self._bn_train(z, y, bn, relu, residual, stats, ws=ws, sums_ready=fused)
        self.tape.append(lambda: self._bwd_conv_unit(x, conv, bn, relu, y, z, stats, residual=residual, fwd_affine=ss))
        return y

    def _bn_ws(self, c):
        return (self.f32_buf(2 * c, dtype=torch.float64, zero=True), self.f32_buf(c), self.f32_buf(c))

    @staticmethod
    def _bn_ptrs(bn, device):
        for t in (bn.weight, bn.bias, bn.running_mean, bn.running_var, bn.num_batches_tracked):
            if t is not None and (t.device != device or not t.is_contiguous()):
                raise RuntimeError("train mode updates the BatchNorm buffers in place: the module must live on %s" % device)
        p = lambda t: t.data_ptr() if t is not None else None
        return (p(bn.weight), p(bn.bias), p(bn.running_mean), p(bn.running_var), p(bn.num_batches_tracked),
                float(bn.eps), float(bn.momentum if bn.momentum is not None else 0.1))

    def _bn_train(self, z, y, bn, relu, residual=None, stats=None, ws=None, sums_ready=False):
        """z: ActMap holding the raw conv output -> y (may be z: in place) = normalised with batch statistics
        (+residual) (+ReLU). stats: fp32 [2c] tensor receiving mean | invstd for the backward pass. ws: the (sums,
        scale, shift) work buffers; sums_ready: the conv that wrote z already added this batch's sums to ws[0]."""
        gamma, beta, rm, rv, nbt, eps, mom = self._bn_ptrs(bn, self.device)
        sums, scale, shift = ws if ws is not None else self._bn_ws(z.c)
        inplace = y is z
        fn = self._lib.w2c_bn_train_from_sums_fwd if sums_ready else self._lib.w2c_bn_train_fwd
        self._record(fn, z.buf.data_ptr(), residual.buf.data_ptr() if residual is not None else None,
                     z.n * z.h * z.w, z.c, z.cstride, z.coffset, self.act, int(bool(relu)), gamma, beta, eps, mom, rm, rv,
                     nbt, sums.data_ptr(), scale.data_ptr(), shift.data_ptr(), None if inplace else y.buf.data_ptr(),
                     0 if inplace else y.cstride, 0 if inplace else y.coffset,
                     stats.data_ptr() if stats is not None else None)
        if residual is not None and not inplace and (residual.cstride, residual.coffset) != (y.cstride, y.coffset):
            raise ValueError("bn_train: the residual must be laid out like the output")
        return scale, shift

    def _bn_train_nchw(self, z, y, bn, relu, stats=None):
        gamma, beta, rm, rv, nbt, eps, mom = self._bn_ptrs(bn, self.device)
        n, c, h, w = z.shape
        sums, scale, shift = self._bn_ws(c)
        self._record(self._lib.w2c_bn_train_nchw_fwd, z.data_ptr(), n, c, h * w, int(bool(relu)), gamma, beta, eps, mom, rm,
                     rv, nbt, sums.data_ptr(), scale.data_ptr(), shift.data_ptr(), None if y is z else y.data_ptr(),
                     stats.data_ptr() if stats is not None else None)

    # ---- backward-pass recording (grad mode) ---------------------------------------------------------------------
    def grad_map(self, a):
        """The gradient map of forward map `a`: the same view (shape, strides, channel slice) of a zero-initialised
        mirror of a's underlying buffer, so that slices / image ranges of one forward buffer are slices of one
        gradient buffer. Gradients use the program's own storage format (train mode runs bf16 / bf16x3)."""
        st = a.buf.untyped_storage()
        root = self._grad_roots.get(st.data_ptr())
        if root is None:
            root = torch.zeros(st.nbytes() // 2, dtype=torch.bfloat16, device=self.device)
            self._grad_roots[st.data_ptr()] = root
        gv = torch.as_strided(root, a.buf.shape, a.buf.stride(), a.buf.storage_offset())
        return ActMap(gv, a.n, a.h, a.w, a.c, a.cstride, a.coffset)

    def f32_grad(self, t):
        """Gradient buffer of an fp32 NCHW tensor of the forward program (the logits; simple_decoder's small map)."""
        st = t.untyped_storage()
        root = self._f32_grads.get(st.data_ptr())
        if root is None:
            root = torch.zeros(st.nbytes() // 4, dtype=torch.float32, device=self.device)
            self._f32_grads[st.data_ptr()] = root
        return torch.as_strided(root, t.shape, t.stride(), t.storage_offset())

    def _claim(self, g):
        """True for the first writer of gradient map g (later writers must accumulate)."""
        key = (g.buf.data_ptr(), g.n, g.coffset, g.c)
        first = key not in self._grad_claimed
        self._grad_claimed.add(key)
        return first

    def _pgrad(self, *shape):
        t = torch.zeros(shape, dtype=torch.float32, device=self.device)
        self._zero_list.append(t)
        return t

    def _bwd_conv_unit(self, x, conv, bn, relu, y, z, stats, residual=None, nchw=False, stem=None, fwd_affine=None):
        """Backward of one conv (+BatchNorm) (+residual) (+ReLU) unit, recorded into self.bprog: gradient of y ->
        gradient of the raw conv output (BatchNorm / ReLU backward, csrc/bn_bwd.cu) -> weight gradient (csrc/wgrad.cu)
        and data gradient (the forward conv kernels on the re-indexed weight). x: input ActMap, or None for a first
        layer (`stem` = (x_nchw, b, n_agents, h, w, c_first, ksize): its weight gradient reads the fp32 views)."""
        bp = self.bprog
        lib = self._lib
        transposed = isinstance(conv, torch.nn.ConvTranspose2d)
        cout = conv.out_channels
        cpad = (cout + 63) // 64 * 64
        g_act = self.act
        p = lambda t: t.data_ptr() if t is not None else None
        # ---- 1. gradient of the raw conv output
        dgamma = self._pgrad(cout) if bn is not None else None
        dbeta = self._pgrad(cout)          # BatchNorm bias, or the conv bias when there is no BatchNorm
        sums = bp.f32_buf(2 * cout, dtype=torch.float64, zero=True)
        coef = bp.f32_buf(3 * cout)
        gamma_t = bn.weight if bn is not None else None
        if nchw:
            n, _, hh, ww = y.shape
            dy_t = self.f32_grad(y)
            # the NHWC gradient map is padded to the 64 channels the weight / data gradient kernels read; the kernel
            # writes the 8-channel groups that hold real channels only (11 classes: 16 of 64), the rest is zeroed
            # ONCE here and never written again (0.34 GB of scattered 16-byte stores per 10 frames otherwise)
            dz = bp.act_buf(n, hh, ww, cpad)
            dz.buf.zero_()
            c8 = (cout + 7) // 8 * 8
            bp._record(lib.w2c_bn_train_nchw_bwd, dy_t.data_ptr(), y.data_ptr(), p(z), dz.buf.data_ptr(), n, cout, hh * ww,
                       c8, dz.cstride, 0, g_act, int(bool(relu)), p(gamma_t), p(stats), p(dgamma), dbeta.data_ptr(),
                       sums.data_ptr(), coef.data_ptr())
        else:
            gy = self.grad_map(y)
            if z is None:
                dz = bp.act_buf(y.n, y.h, y.w, y.c)
            else:
                dz = self.grad_map(z)
            dres = None
            if residual is not None:
                dres = self.grad_map(residual)
                if not self._claim(dres):
                    raise NotImplementedError("residual gradient must be the first gradient of its map")
            a = _lib.BnBwdArgs(dy=gy.buf.data_ptr(), y=y.buf.data_ptr(), z=z.buf.data_ptr() if z is not None else None,
                               dz=dz.buf.data_ptr(), dres=dres.buf.data_ptr() if dres is not None else None,
                               n_px=y.n * y.h * y.w, c=y.c, dy_cstride=gy.cstride, dy_coffset=gy.coffset,
                               y_cstride=y.cstride, y_coffset=y.coffset, z_cstride=z.cstride if z is not None else 0,
                               z_coffset=z.coffset if z is not None else 0, dz_cstride=dz.cstride, dz_coffset=dz.coffset,
                               dres_cstride=dres.cstride if dres is not None else 0,
                               dres_coffset=dres.coffset if dres is not None else 0, act_f=self.act, act_g=g_act,
                               relu=int(bool(relu)), gamma=p(gamma_t), stats=p(stats), dgamma=p(dgamma),
                               dbeta=dbeta.data_ptr(), sums_ws=sums.data_ptr(), coef_ws=coef.data_ptr())
            if fwd_affine is not None and residual is None and relu and z is not None:
                # the ReLU mask from z * scale + shift (what the forward evaluated): y is not read (csrc/bn_bwd.cu)
                a.fwd_scale, a.fwd_shift = fwd_affine[0].data_ptr(), fwd_affine[1].data_ptr()
            bp.keep.append(a)
            bp._record(lib.w2c_bn_train_bwd, ctypes.byref(a))
        if bn is not None:
            self.param_grads.append((bn.weight, lambda: dgamma))
            self.param_grads.append((bn.bias, lambda: dbeta))
            if conv.bias is not None:   # a bias in front of a train-mode BatchNorm has no gradient (the mean removes it)
                zb = torch.zeros(cout, dtype=torch.float32, device=self.device)
                self.param_grads.append((conv.bias, lambda: zb))
        elif conv.bias is not None:
            self.param_grads.append((conv.bias, lambda: dbeta))
        # ---- 2. weight gradient
        kh, kw = conv.kernel_size
        if stem is not None:
            x_nchw, b, n_agents, h, w, c_first, ksize = stem
            dw = self._pgrad(cout, 3, ksize, ksize)
            bp._record(lib.w2c_stem_conv_wgrad, x_nchw.data_ptr(), dz.buf.data_ptr(), dw.data_ptr(), ksize, b, n_agents,
                       x_nchw.shape[1], c_first, h, w, cout, dz.cstride, dz.coffset, g_act)
            self.param_grads.append((conv.weight, lambda: dw))
            return
        stride = conv.stride[0]
        kind = self.weights.conv_raw(conv).kind if bn is not None else self.weights.conv(conv, None, relu).kind
        if stride == 4:
            raise NotImplementedError("backward of the stride-4 squeezer (feat_squeezer=4)")
        d0, d1 = (x.c, cpad) if transposed else (cpad, x.c)
        dwk = self._pgrad(d0, kh * kw, d1)
        wa = _lib.WgradArgs(x=x.buf.data_ptr(), dy=dz.buf.data_ptr(), dw=dwk.data_ptr(), n=x.n, h_in=x.h, w_in=x.w,
                            cin=x.c, cout=cpad, x_cstride=x.cstride, x_coffset=x.coffset, dy_cstride=dz.cstride,
                            dy_coffset=dz.coffset, kind=kind, act_x=self.act, act_dy=g_act, passes=0)
        bp.keep.append(wa)
        bp._record(lib.w2c_conv_wgrad, ctypes.byref(wa))

        def weight_grad(dwk=dwk, d0=d0, d1=d1):
            gq = dwk.view(d0, kh, kw, d1).permute(0, 3, 1, 2)
            return gq[:, :cout] if transposed else gq[:cout]
        self.param_grads.append((conv.weight, weight_grad))
        # ---- 3. data gradient (not for maps nothing upstream differentiates through)
        gx = self.grad_map(x)
        first = self._claim(gx)
        pcd = bp.weights.dgrad(conv)
        if (kh, kw) == (1, 1) and stride == 2:
            small = bp.conv(dz, pcd)
            bp._record(lib.w2c_upsample_zero2, small.buf.data_ptr(), None if first else gx.buf.data_ptr(), gx.buf.data_ptr(),
                       x.n, x.h, x.w, x.c, g_act)
            if gx.cstride != gx.c:
                raise NotImplementedError("stride-2 1x1 data gradient into a channel slice")
        else:
            bp.conv(dz, pcd, out=gx, residual=None if first else gx)

    def stem_bn(self, x_in, conv, bn, b, n_agents, h, w, c_first=0):
        """First layer (3x3 s1 of n_segnet_encoder or 7x7 s2 of resnet18) + BatchNorm + ReLU."""
        k7 = tuple(conv.kernel_size) == (7, 7)
        if not self.train:
            st = self.weights.stem(conv, bn)
            return (self.stem7x7 if k7 else self.stem3x3)(x_in, st, b, n_agents, h, w, c_first)
        z = (self.stem7x7 if k7 else self.stem3x3)(x_in, self.weights.stem_raw(conv), b, n_agents, h, w, c_first, raw=True)
        return self._stem_bn_tail(x_in, conv, bn, z, b, n_agents, h, w, c_first, 7 if k7 else 3)

    def _stem_bn_tail(self, x_in, conv, bn, z, b, n_agents, h, w, c_first, ksize):
        if not self.grad:
            self._bn_train(z, z, bn, True)
            return z
        y = self.act_buf(z.n, z.h, z.w, z.c)
        stats = self.f32_buf(2 * z.c)
        ss = self._bn_train(z, y, bn, True, None, stats)
        self.tape.append(lambda: self._bwd_conv_unit(None, conv, bn, True, y, z, stats, fwd_affine=ss,
                                                     stem=(x_in, b, n_agents, h, w, c_first, ksize)))
        return y

    def stem_pair_bn(self, x_in, conv_a, bn_a, conv_b, bn_b, b, n_agents, h, w):
        """Two encoders' first layers as one 3 -> 128 stem writing two dense 64-channel maps (+ BatchNorm + ReLU)."""
        k7 = tuple(conv_a.kernel_size) == (7, 7)
        fn = self.stem7x7 if k7 else self.stem3x3
        if not self.train:
            return fn(x_in, self.weights.stem_pair(conv_a, bn_a, conv_b, bn_b), b, n_agents, h, w, split=True)
        st = self.weights.stem_raw_pair(conv_a, conv_b)
        za, zb = fn(x_in, st, b, n_agents, h, w, split=True, raw=True)
        ks = 7 if k7 else 3
        return (self._stem_bn_tail(x_in, conv_a, bn_a, za, b, n_agents, h, w, 0, ks),
                self._stem_bn_tail(x_in, conv_b, bn_b, zb, b, n_agents, h, w, 0, ks))

    def can_fuse_head(self, stack):
        """True when conv1 + conv2 of an n_segnet encoder in `stack` may run as the fused head kernel: always in the
        one-plane formats; in a two-plane format only where the precision plan runs BOTH layers in one pass (the
        conv1 map inside the kernel is a single plane)."""
        if not FUSE_ENCODER_HEAD or self.train:
            return False
        if self.planes == 1:
            return True
        return self.passes_for(stack, 1) == 1 and self.passes_for(stack, 2) == 1

    def enc_head(self, x_in, st1, pc2, b, n_agents, h, w, c_first=0):
        """conv1 (stem operands st1 = (w [64][27], scale, shift)) + conv2 (PackedConv pc2) -> ActMap (h/2, w/2, 64)."""
        w1, scale1, shift1 = st1
        if w1.shape[0] != 64 or pc2.cin != 64 or pc2.cout != 64 or pc2.kind != ops.CONV3X3_S2 or not pc2.relu:
            raise ValueError("enc_head covers conv1 3->64 followed by conv2 64->64 stride 2 with ReLU")
        out = self.act_buf(b * n_agents, h // 2, w // 2, 64)
        a = _lib.EncHeadArgs(x=x_in.data_ptr(), lut=self.lut.data_ptr() if self.input_u8 else None, w1=w1.data_ptr(),
                             scale1=scale1.data_ptr(), shift1=shift1.data_ptr(), w2=pc2.w.data_ptr(),
                             scale2=pc2.scale.data_ptr(), shift2=pc2.shift.data_ptr(), y=out.buf.data_ptr(),
                             x_u8=int(self.input_u8), b=b, n_agents=n_agents, c_total=x_in.shape[1],
                             c_first=c_first // 3 if self.input_u8 else c_first, h=h, w=w, act=self.act,
                             y_cstride=out.cstride, y_coffset=out.coffset)
        self.keep.append(a)
        self._record(self._lib.w2c_enc_head_fwd, ctypes.byref(a))
        return out

    def stem3x3(self, x_nchw, st, b, n_agents, h, w, c_first=0, split=False, raw=False):
        """split=True (a fused pair of 64-channel first layers): returns TWO dense 64-channel maps. raw=True: no ReLU
        (the conv output train-mode BatchNorm starts from)."""
        wt, scale, shift = st
        cout = wt.shape[0]
        n_split = 2 if split else 1
        if split:
            if cout != 128:
                raise ValueError("a split stem needs 128 output channels")
            buf = torch.empty((2, b * n_agents, h, w, self.planes * 64), dtype=torch.bfloat16, device=self.device)
            self.keep.append(buf)
            out = tuple(ActMap(buf[i], b * n_agents, h, w, 64) for i in range(2))
            y_ptr = buf.data_ptr()
        else:
            out = self.act_buf(b * n_agents, h, w, cout)
            y_ptr = out.buf.data_ptr()
        if self.input_u8:  # x_nchw is the uint8 frame buffer [b, agents_total, h, w, 3]; c_first counts channels
            self._record(self._lib.w2c_stem_conv3x3_u8_fwd, x_nchw.data_ptr(), self.lut.data_ptr(), wt.data_ptr(),
                         scale.data_ptr(), shift.data_ptr(), y_ptr, b, n_agents, x_nchw.shape[1], c_first // 3, h, w,
                         cout, self.act, n_split)
            return out
        fn = self._lib.w2c_stem_conv3x3_raw_fwd if raw else self._lib.w2c_stem_conv3x3_fwd
        self._record(fn, x_nchw.data_ptr(), wt.data_ptr(), scale.data_ptr(),
                     shift.data_ptr(), y_ptr, b, n_agents, x_nchw.shape[1], c_first, h, w, cout, self.act, n_split)
        return out

    def stem7x7(self, x_nchw, st, b, n_agents, h, w, c_first=0, split=False, raw=False):
        """resnet18 first layer (7x7 s2). split=True (a fused pair of 64-channel first layers): two dense maps."""
        wt, scale, shift = st
        cout = wt.shape[0]
        n_split = 2 if split else 1
        if split:
            if cout != 128:
                raise ValueError("a split stem needs 128 output channels")
            buf = torch.empty((2, b * n_agents, h // 2, w // 2, self.planes * 64), dtype=torch.bfloat16,
                              device=self.device)
            self.keep.append(buf)
            out = tuple(ActMap(buf[i], b * n_agents, h // 2, w // 2, 64) for i in range(2))
            y_ptr = buf.data_ptr()
        else:
            out = self.act_buf(b * n_agents, h // 2, w // 2, cout)
            y_ptr = out.buf.data_ptr()
        if self.input_u8:
            self._record(self._lib.w2c_stem_conv7x7s2_u8_fwd, x_nchw.data_ptr(), self.lut.data_ptr(), wt.data_ptr(),
                         scale.data_ptr(), shift.data_ptr(), y_ptr, b, n_agents, x_nchw.shape[1], c_first // 3, h, w,
                         cout, self.act, n_split)
            return out
        fn = self._lib.w2c_stem_conv7x7s2_raw_fwd if raw else self._lib.w2c_stem_conv7x7s2_fwd
        self._record(fn, x_nchw.data_ptr(), wt.data_ptr(), scale.data_ptr(),
                     shift.data_ptr(), y_ptr, b, n_agents, x_nchw.shape[1], c_first, h, w, cout, self.act, n_split)
        return out

    def maxpool(self, x):
        out = self.act_buf(x.n, x.h // 2, x.w // 2, x.c)
        self._record(self._lib.w2c_maxpool3x3s2_fwd, x.buf.data_ptr(), out.buf.data_ptr(), x.n, x.h, x.w, x.c, self.act)
        if self.grad:
            def bwd():
                if x.cstride != x.c or x.coffset:
                    raise NotImplementedError("max-pool backward on a channel slice")
                gx, gy = self.grad_map(x), self.grad_map(out)
                if not self._claim(gx):
                    raise NotImplementedError("max-pool gradient must be the first gradient of its input")
                self.bprog._record(self._lib.w2c_maxpool3x3s2_bwd, x.buf.data_ptr(), gy.buf.data_ptr(), gx.buf.data_ptr(),
                                   x.n, x.h, x.w, x.c, self.act, self.act)
            self.tape.append(bwd)
        return out

    def bilinear(self, x_nchw, factor):
        n, c, h, w = x_nchw.shape
        out = self.f32_buf(n, c, h * factor, w * factor)
        self._record(self._lib.w2c_bilinear_up_fwd, x_nchw.data_ptr(), out.data_ptr(), n, c, h, w, factor)
        if self.grad:
            self.tape.append(lambda: self.bprog._record(self._lib.w2c_bilinear_up_bwd, self.f32_grad(out).data_ptr(),
                                                        self.f32_grad(x_nchw).data_ptr(), n, c, h, w, factor))
        return out

    def bilinear_argmax(self, x_nchw, factor, labels):
        n, c, h, w = x_nchw.shape
        self._record(self._lib.w2c_bilinear_argmax_fwd, x_nchw.data_ptr(), labels.data_ptr(), n, c, h, w, factor)
        return labels

    def argmax_labels(self, logits, labels):
        n, c, h, w = logits.shape
        self._record(self._lib.w2c_argmax_labels_fwd, logits.data_ptr(), labels.data_ptr(), n, c, h * w)
        return labels

    def kq_mlp(self, feat, mlp, out_dim, out=None):
        w0, b0, w1, b1, w2, b2 = mlp
        m = feat.n
        n_feat = feat.h * feat.w * feat.c
        if out is None:
            out = self.f32_buf(m, out_dim)
        elif tuple(out.shape) != (m, out_dim) or out.dtype != torch.float32 or not out.is_contiguous():
            raise ValueError("kq_mlp: out must be a contiguous fp32 [%d, %d] tensor" % (m, out_dim))
        ws = self.f32_buf(m * 384)
        self._record(self._lib.w2c_kq_mlp_fwd, feat.buf.data_ptr(), self.act, m, n_feat, w0.data_ptr(), b0.data_ptr(),
                     w1.data_ptr(), b1.data_ptr(), w2.data_ptr(), b2.data_ptr(), out_dim, out.data_ptr(), ws.data_ptr())
        return out

    def kq_mlp_heads(self, feat, heads, fcs=None):
        """heads: [(mlp weights tuple, out_dim, out tensor or None), ...] (1 or 2) over the same feature map, run as
        ONE fc0 + ONE fc12 launch. Returns the output tensors. fcs: the nn.Sequential of each head (grad mode: where
        the parameter gradients go)."""
        m = feat.n
        n_feat = feat.h * feat.w * feat.c
        arr = (_lib.MlpHead * len(heads))()
        outs = []
        for i, (mlp, out_dim, out) in enumerate(heads):
            w0, b0, w1, b1, w2, b2 = mlp
            if out is None:
                out = self.f32_buf(m, out_dim)
            elif tuple(out.shape) != (m, out_dim) or out.dtype != torch.float32 or not out.is_contiguous():
                raise ValueError("kq_mlp: out must be a contiguous fp32 [%d, %d] tensor" % (m, out_dim))
            arr[i] = _lib.MlpHead(w0.data_ptr(), b0.data_ptr(), w1.data_ptr(), b1.data_ptr(), w2.data_ptr(),
                                  b2.data_ptr(), out.data_ptr(), out_dim)
            outs.append(out)
        ws = self.f32_buf(len(heads) * m * 256)
        self.keep.append(arr)
        self._record(self._lib.w2c_kq_mlp_heads_fwd, feat.buf.data_ptr(), self.act, m, n_feat, arr, len(heads),
                     ws.data_ptr())
        if self.grad:
            self.tape.append(lambda: self._bwd_mlp_heads(feat, heads, fcs, arr, outs, ws, m, n_feat))
        return outs

    def _bwd_mlp_heads(self, feat, heads, fcs, arr, outs, ws_fwd, m, n_feat):
        if fcs is None:
            raise NotImplementedError("kq_mlp_heads: the modules are needed to route the parameter gradients")
        if feat.cstride != feat.c or feat.coffset or feat.c != 256:
            raise NotImplementedError("MLP-head backward needs a dense 256-channel policy map")
        bp = self.bprog
        garr = (_lib.MlpHeadGrad * len(heads))()
        side = feat.h
        for i, ((mlp, out_dim, _), fc) in enumerate(zip(heads, fcs)):
            dout = self.f32_grad(outs[i])
            dw0, db0 = self._pgrad(256, n_feat), self._pgrad(256)
            dw1, db1 = self._pgrad(128, 256), self._pgrad(128)
            dw2, db2 = self._pgrad(out_dim, 128), self._pgrad(out_dim)
            garr[i] = _lib.MlpHeadGrad(dout.data_ptr(), dw0.data_ptr(), db0.data_ptr(), dw1.data_ptr(), db1.data_ptr(),
                                       dw2.data_ptr(), db2.data_ptr())
            # fc.0.weight is stored NCHW-flattened, the kernel works in NHWC flatten order (WeightCache.mlp)
            self.param_grads.append((fc[0].weight, lambda dw0=dw0: dw0.view(256, side, side, 256).permute(0, 3, 1, 2)
                                     .reshape(256, n_feat)))
            for prm, gt in ((fc[0].bias, db0), (fc[2].weight, dw1), (fc[2].bias, db1), (fc[4].weight, dw2),
                            (fc[4].bias, db2)):
                self.param_grads.append((prm, lambda gt=gt: gt))
        gfeat = self.grad_map(feat)
        if not self._claim(gfeat):
            raise NotImplementedError("MLP-head gradient must be the first gradient of the policy map")
        ws = bp.f32_buf(len(heads) * m * 512 + 256)
        bp.keep.append(garr)
        bp._record(self._lib.w2c_kq_mlp_heads_bwd, feat.buf.data_ptr(), self.act, m, n_feat, arr, garr, len(heads),
                   ws_fwd.data_ptr(), gfeat.buf.data_ptr(), self.act, ws.data_ptr())

    def attn(self, keys, queries, wq, bq, val, fused, prob, coef, action, connect, *, b_sz, n_k, n_q, k_dim, q_dim,
             mode, sparse=False, mask_self=False, temperature=1.0, diag_bias=0.0, thresh=0.2, q_first=0, q_count=0,
             agents_per_rank=0, keys_rank_stride=0, queries_rank_stride=0, val_rank_stride=0, attn_module=None):
        p = lambda t: t.data_ptr() if t is not None else None
        a = _lib.AttnArgs(keys=p(keys), queries=p(queries), wq=p(wq), bq=p(bq), val=val.buf.data_ptr(),
                          fused=fused.buf.data_ptr(), prob_out=p(prob), coef_out=p(coef), action=p(action),
                          connect=p(connect), b_sz=b_sz, n_k=n_k, n_q=n_q, k_dim=k_dim, q_dim=q_dim,
                          hw=val.h * val.w, c=val.c, fused_cstride=fused.cstride, fused_coffset=fused.coffset,
                          act=self.act, mode=mode, sparse=int(bool(sparse)), mask_self=int(bool(mask_self)),
                          temperature=float(temperature), diag_bias=float(diag_bias), thresh=float(thresh),
                          q_first=q_first, q_count=q_count, agents_per_rank=agents_per_rank,
                          keys_rank_stride=keys_rank_stride, queries_rank_stride=queries_rank_stride,
                          val_rank_stride=val_rank_stride)
        if val.cstride != val.c or val.coffset != 0:
            raise ValueError("attention values must be a dense NHWC map")
        self.keep.append(a)
        self._record(self._lib.w2c_attn_fuse_fwd, ctypes.byref(a))
        if self.grad:
            if mode != ops.FUSE_SOFTMAX or agents_per_rank or coef is None:
                raise NotImplementedError("attention backward: dense agent-major softmax mode with coef_out only")
            self.tape.append(lambda: self._bwd_attn(keys, queries, wq, bq, val, fused, coef, b_sz, n_k, n_q, k_dim, q_dim,
                                                    sparse, temperature, attn_module))

    def _bwd_attn(self, keys, queries, wq, bq, val, fused, coef, b_sz, n_k, n_q, k_dim, q_dim, sparse, temperature,
                  attn_module):
        bp = self.bprog
        p = lambda t: t.data_ptr() if t is not None else None
        gval = self.grad_map(val)
        first = self._claim(gval)
        gfused = self.grad_map(fused)
        dkeys, dqueries = self.f32_grad(keys), self.f32_grad(queries)
        dwq = dbq = None
        if wq is not None:
            if attn_module is None:
                raise NotImplementedError("attention backward needs the projection module for its parameter gradients")
            dwq, dbq = self._pgrad(*wq.shape), self._pgrad(*bq.shape)
            self.param_grads.append((attn_module.linear.weight, lambda: dwq))
            self.param_grads.append((attn_module.linear.bias, lambda: dbq))
        dp = bp.f32_buf(b_sz * n_k * n_q, zero=True)
        a = _lib.AttnBwdArgs(keys=p(keys), queries=p(queries), wq=p(wq), bq=p(bq), val=val.buf.data_ptr(),
                             dfused=gfused.buf.data_ptr(), prob=p(coef), dval=gval.buf.data_ptr(), dkeys=p(dkeys),
                             dqueries=p(dqueries), dwq=p(dwq), dbq=p(dbq), dp_ws=p(dp), b_sz=b_sz, n_k=n_k, n_q=n_q,
                             k_dim=k_dim, q_dim=q_dim, hw=val.h * val.w, c=val.c, dfused_cstride=gfused.cstride,
                             dfused_coffset=gfused.coffset, act_f=self.act, act_g=self.act, sparse=int(bool(sparse)),
                             dval_accumulate=0 if first else 1, temperature=float(temperature))
        bp.keep.append(a)
        bp._record(self._lib.w2c_attn_fuse_bwd, ctypes.byref(a))

    def host_op(self, fn, capturable=False):
        """A torch-side step between kernel launches (the torch.distributed collective of a sharded forward); fn() is
        called on the current stream. capturable (and engine.CAPTURE_COLLECTIVES): it is recorded like a launch and
        ends up inside the CUDA graph; otherwise it splits the program into separately captured segments."""
        if self._sid:
            raise RuntimeError("host ops cannot be recorded on the side stream")
        if capturable and CAPTURE_COLLECTIVES:
            def run(_stream):
                fn()
                return 0
            self.calls.append((run, None, 0))
            return
        self.calls.append((None, fn, 0))

    def gather_images(self, src, dst, b, n_groups, sel=None):
        """dst image group g (b images), channel slice of `dst`  <-  src image group sel[g] (device int32 tensor; None =
        identity), channel slice of `src`."""
        if src.c != dst.c or (src.h, src.w) != (dst.h, dst.w):
            raise ValueError("gather_images: source and destination maps differ in shape")
        if self.grad:
            raise NotImplementedError("backward of the indexed image gather (selection / ComNet baselines)")
        self._record(self._lib.w2c_gather_images_fwd, src.buf.data_ptr(), dst.buf.data_ptr(),
                     sel.data_ptr() if sel is not None else None, n_groups, b, src.h, src.w, src.c, src.cstride,
                     src.coffset, dst.cstride, dst.coffset, self.act)

    def copy_channels(self, src, dst):
        """dst[..., slice] = src (device-to-device strided copy through torch; used for concat inputs only)."""
        if self.grad:
            def bwd():
                gs, gd = self.grad_map(src), self.grad_map(dst)
                first = self._claim(gs)
                self.bprog._record(self._lib.w2c_grad_add, gd.buf.data_ptr(), gd.cstride, gd.coffset,
                                   None if first else gs.buf.data_ptr(), gs.cstride, gs.coffset, gs.buf.data_ptr(),
                                   gs.cstride, gs.coffset, src.n * src.h * src.w, src.c, self.act)
            self.tape.append(bwd)

        def run(_stream):
            for pl in range(self.planes):
                d = dst.buf[..., pl * dst.cstride + dst.coffset: pl * dst.cstride + dst.coffset + dst.c]
                s = src.buf[..., pl * src.cstride + src.coffset: pl * src.cstride + src.coffset + src.c]
                d.copy_(s)
            return 0
        self.calls.append((run, None, self._sid))

    def memset(self, t):
        def run(_stream):
            t.zero_()
            return 0
        self.calls.append((run, None, self._sid))

    # ---- backward program
    def begin_backward(self):
        """Switch the program into grad mode (before any op is recorded): a second Program collects the backward
        launches, with its own live operand cache (the data-gradient weights are re-packed every backward)."""
        if self.act not in (ops.ACT_BF16, ops.ACT_BF16X2):
            raise NotImplementedError(
                "the backward pass runs in the 'bf16' / 'bf16x3' precisions (gradients need the fp32 exponent range, and "
                "the tensor cores take one element type per MMA): model.set_precision('bf16x3') for training")
        self.grad = True
        self.bprog = Program(None, self.device, self.act)
        self.bprog.weights = WeightCache(self.device, self.act, live=self.bprog)

    def finish_backward(self, pred):
        """Replay the tape in reverse into self.bprog. pred: the fp32 logits the loss is taken on."""
        self.dlogits = self.f32_grad(pred)
        bp = self.bprog
        zl = self._zero_list

        def zero(_s):
            if zl:
                torch._foreach_zero_(zl)
            return 0
        bp.calls.append((zero, None, 0))
        # The two encoder chains of the forward (side_stream / join) are independent in the backward pass too: replayed in
        # reverse, the forward's join is where the backward forks (the side stream waits for the decoder / attention
        # backward), the policy chain's backward stays on the main stream, the feature encoder's runs beside it, and the
        # forward's fork is where they join again. (Gradient buffers, work buffers and weight-gradient targets are
        # per layer; the operand packs sit at the head of the program, before the fork.)
        for fn in reversed(self.tape):
            if fn is Program._JOIN:
                bp.calls.append((Program._FORK, None, 0))
            elif fn is Program._SIDE_END:
                bp._sid = 1
            elif fn is Program._SIDE_BEGIN:
                bp._sid = 0
                bp.join()
            else:
                fn()
        self.tape = None
        # the prologue was recorded before the list was complete; it reads zl at run time

    def param_gradients(self):
        """{parameter: fp32 gradient tensor in the parameter's layout} after a run of self.bprog (views of the static
        gradient buffers: copy before the next backward)."""
        out = {}
        for prm, fn in self.param_grads:
            g = fn()
            out[prm] = g if prm not in out else out[prm] + g   # a module used twice in one forward
        return out

    # ---- execution
    def _segments(self):
        """Split the call list at host ops: [(calls, host_fn_or_None), ...]."""
        segs, cur = [], []
        for fn, args, sid in self.calls:
            if fn is None:
                segs.append((cur, args))
                cur = []
            else:
                cur.append((fn, args, sid))
        segs.append((cur, None))
        return segs

    def _run_calls(self, calls):
        main = torch.cuda.current_stream(self.device)
        streams = [main, None]
        ptrs = [ctypes.c_void_p(main.cuda_stream), None]
        forked = False
        for fn, args, sid in calls:
            if fn is Program._FORK:
                if self._side is None:
                    self._side = torch.cuda.Stream(self.device)
                streams[1], ptrs[1] = self._side, ctypes.c_void_p(self._side.cuda_stream)
                self._side.wait_stream(main)
                forked = True
                continue
            if fn is Program._JOIN:
                if forked:
                    main.wait_stream(self._side)
                    forked = False
                continue
            if sid and not forked:
                sid = 0  # (a sub-program that dropped the fork marker)
            if args is None:  # torch-side op (strided copy / memset): runs on the stream that is current
                if sid:
                    with torch.cuda.stream(streams[1]):
                        rc = fn(ptrs[1])
                else:
                    rc = fn(ptrs[0])
            else:
                rc = fn(*args, ptrs[sid])
            if rc != 0:
                _lib.check(rc, getattr(fn, "__name__", "launch"))
        if forked:
            main.wait_stream(self._side)

    def _batch_setup_calls(self):
        """Live (train-mode) programs re-derive every layer's packed operand and folded bias from the parameters on
        every run: 43-47 layers x (pack, fold) tiny launches in the forward program, as many data-gradient packs in
        the backward one. They depend on the parameters only, so they are merged into ONE launch per kind at the head
        of the program (w2c_pack_conv_weights_batch / w2c_fold_bn_batch: bit-identical results)."""
        if self._batched:
            return
        self._batched = True
        lib = self._lib
        packs, folds, rest = [], [], []
        for call in self.calls:
            fn, args, _sid = call
            if fn is lib.w2c_pack_conv_weight:
                w, cout, cin_real, cin, ntaps, tr, act, packed = args
                packs.append((act, _lib.PackItem(w=w, packed=packed, cout=cout, cin_real=cin_real, cin=cin, ntaps=ntaps,
                                                 transposed=tr, flip=0)))
            elif fn is lib.w2c_pack_conv_weight_ex:
                w, cout, cin_real, cin, ntaps, tr, flip, act, packed = args
                packs.append((act, _lib.PackItem(w=w, packed=packed, cout=cout, cin_real=cin_real, cin=cin, ntaps=ntaps,
                                                 transposed=tr, flip=flip)))
            elif fn is lib.w2c_fold_bn:
                bias, gamma, beta, mean, var, eps, cout, scale, shift = args
                folds.append(_lib.FoldItem(conv_bias=bias, gamma=gamma, beta=beta, mean=mean, var=var, scale=scale,
                                           shift=shift, eps=eps, cout=cout))
            else:
                rest.append(call)
        if len(packs) + len(folds) < 3 or len({a for a, _ in packs}) > 1:
            return
        head = []
        if packs:
            arr = (_lib.PackItem * len(packs))(*[it for _, it in packs])
            self.keep.append(arr)
            head.append((lib.w2c_pack_conv_weights_batch, (arr, len(packs), packs[0][0]), 0))
        if folds:
            arr = (_lib.FoldItem * len(folds))(*folds)
            self.keep.append(arr)
            head.append((lib.w2c_fold_bn_batch, (arr, len(folds)), 0))
        self.calls = head + rest

    def _run_eager(self):
        for calls, host in self._segments():
            self._run_calls(calls)
            if host is not None:
                host()

    def run(self, use_graph):
        self._batch_setup_calls()
        if not use_graph:
            if not self.n_launches:
                before = ops.launch_count()
                self._run_eager()
                self.n_launches = ops.launch_count() - before
            else:
                self._run_eager()
            return
        if self.graph is None:
            # first run eagerly (sets kernel attributes, surfaces errors), then capture each segment
            before = ops.launch_count()
            self._run_eager()
            self.n_launches = ops.launch_count() - before
            torch.cuda.synchronize(self.device)
            graphs = []
            for calls, host in self._segments():
                g = None
                if calls:
                    g = torch.cuda.CUDAGraph()
                    # an explicit capture stream on THIS program's device: torch.cuda.graph's default capture stream
                    # is created once per process, on whichever device captured first
                    if self._capture_stream is None:
                        self._capture_stream = torch.cuda.Stream(self.device)
                    # no garbage collection while a capture is open: freeing an older program (its CUDA graphs, its
                    # private memory pool) from inside the capturing thread invalidates the capture
                    gc_was_on = gc.isenabled()
                    gc.collect()
                    gc.disable()
                    try:
                        with _CAPTURE_LOCK, torch.cuda.graph(g, stream=self._capture_stream,
                                                             capture_error_mode="thread_local"):
                            self._run_calls(calls)
                    finally:
                        if gc_was_on:
                            gc.enable()
                graphs.append((g, host))
            self.graph = graphs
            return  # results of the eager pass are already in the buffers
        for g, host in self.graph:
            if g is not None:
                g.replay()
            if host is not None:
                host()

    def conv_only_program(self):
        """A program replaying just the tensor-core conv launches of this one (same buffers): used by bench.py to
        time the dominant kernel in isolation from the stems / attention / layout kernels."""
        sub = Program(self.weights, self.device, self.act)
        sub.keep = self.keep
        convs = (self._lib.w2c_conv_bnrelu_fwd, self._lib.w2c_enc_head_fwd)
        sub.calls = [c for c in self.calls if c[0] in convs or c[0] in (Program._FORK, Program._JOIN)]
        return sub
