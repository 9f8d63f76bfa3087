"""The When2com model family behind the reference's constructor / forward() signatures.

Each class mirrors one model of ptsemseg/models/agent.py: same constructor keywords, same state_dict keys, same
forward arguments and return tuples. forward() in eval mode compiles (once per input shape) a program of sm_100a
kernel launches through engine.Program and runs it; there is no eager PyTorch path.

  Single_agent      agent.py:375-395        All_agents        agent.py:399-469
  LearnWho2Com      agent.py:472-673        LearnWhen2Com     agent.py:676-889
  MIMO_All_agents   agent.py:892-981        MIMOcom           agent.py:983-1204
  MIMOcomWho        agent.py:1207-1423
"""
import random

import torch
import torch.nn as nn

from .. import engine, ops
from ..lazy import DeviceScalar
from .layers import (_Container, conv2DBatchNormRelu, deconv2DBatchNormRelu, get_decoder, get_encoder,
                     n_segnet_decoder, n_segnet_encoder, resnet_encoder, simple_decoder)

FEATURE_CHANNELS = 512


# ------------------------------------------------------------------------------------------- sub-module containers
class img_encoder(_Container):
    """agent.py:39-60: backbone + squeezer conv (stride by feat_squeezer)."""

    def __init__(self, n_classes=21, in_channels=3, feat_channel=512, feat_squeezer=-1,
                 enc_backbone="n_segnet_encoder"):
        super().__init__()
        self.feature_backbone = get_encoder(enc_backbone)(n_classes=n_classes, in_channels=in_channels)
        self.feat_squeezer = feat_squeezer
        stride = {2: 2, 4: 4}.get(feat_squeezer, 1)
        self.squeezer = conv2DBatchNormRelu(512, feat_channel, 3, stride, 1)


class img_decoder(_Container):
    """agent.py:63-89."""

    def __init__(self, n_classes=21, in_channels=512, agent_num=5, feat_squeezer=-1, dec_backbone="n_segnet_decoder"):
        super().__init__()
        self.feat_squeezer = feat_squeezer
        dec = get_decoder(dec_backbone)
        if feat_squeezer == 2:
            self.desqueezer = deconv2DBatchNormRelu(in_channels, in_channels)
            self.output_decoder = dec(n_classes=n_classes, in_channels=in_channels)
        elif feat_squeezer == 4:
            self.desqueezer1 = deconv2DBatchNormRelu(in_channels, 512)
            self.desqueezer2 = deconv2DBatchNormRelu(512, 512)
            self.output_decoder = dec(n_classes=n_classes, in_channels=512)
        else:
            self.output_decoder = dec(n_classes=n_classes, in_channels=in_channels)


class policy_net4(_Container):
    """agent.py:114-142: a private img_encoder followed by five conv-BN-ReLU (strides 1,1,2,1,2)."""
    SPEC = ((512, 512, 1), (512, 256, 1), (256, 256, 2), (256, 256, 1), (256, 256, 2))

    def __init__(self, n_classes=21, in_channels=512, input_feat_sz=32, enc_backbone="n_segnet_encoder"):
        super().__init__()
        self.img_encoder = img_encoder(n_classes=n_classes, in_channels=in_channels, enc_backbone=enc_backbone)
        for i, (cin, cout, stride) in enumerate(self.SPEC, 1):
            setattr(self, "conv%d" % i, conv2DBatchNormRelu(cin, cout, 3, stride, 1))


class km_generator(_Container):
    """agent.py:145-159 (and the identical `linear`, agent.py:162-178): 3-layer MLP head."""

    def __init__(self, out_size=128, input_feat_sz=32):
        super().__init__()
        side = input_feat_sz // 4
        self.n_feat = int(256 * side * side)
        self.fc = nn.Sequential(nn.Linear(self.n_feat, 256), nn.ReLU(inplace=True), nn.Linear(256, 128),
                                nn.ReLU(inplace=True), nn.Linear(128, out_size))


linear = km_generator


class _DotAttention(_Container):
    """Parameter holder of the *GeneralDotProductAttention modules (agent.py:242-368): one Linear(query -> key)."""

    def __init__(self, query_size, key_size):
        super().__init__()
        self.linear = nn.Linear(query_size, key_size)


class _ScaledAttention(_Container):
    """ScaledDotProductAttention, agent.py:194-213: parameter-free, temperature 128**0.5 (agent.py:523,725)."""

    def __init__(self, temperature):
        super().__init__()
        self.temperature = temperature


class _AdditiveAttention(_Container):
    """AdditiveAttentin, agent.py:215-240: score_i = linear_out(linear_feat(k_i) + linear_context(q)). The query term
    is the same for every supporter i, and softmax / sparsemax over i are shift-invariant, so the attention weights
    are those of a dot product of the keys with ONE fixed vector u = linear_feat.weight^T linear_out.weight^T: the
    attention kernel runs it as a projection-free dot attention whose "query" rows all hold u."""

    def __init__(self):
        super().__init__()
        self.linear_feat = nn.Linear(128, 128)
        self.linear_context = nn.Linear(128, 128)
        self.linear_out = nn.Linear(128, 1)


def _make_attention(attention, query_size, key_size):
    if attention == "general":
        return _DotAttention(query_size, key_size)
    if attention == "additive":
        if key_size != 128:
            raise ValueError("AdditiveAttentin has fixed 128-wide projections (agent.py:221-223); key_size=%d"
                             % key_size)
        return _AdditiveAttention()
    return _ScaledAttention(128 ** 0.5)


# ------------------------------------------------------------------------------------------- program builders
def _build_encoder(prog, enc, key, x_nchw, b, n_agents, h, w, c_first=0, out=None, stem=None, stack="encoder"):
    """img_encoder.forward (agent.py:56-60) on agents [c_first/3, ...) of the fp32 NCHW batch -> ActMap.
    stem: output of an already-issued (fused) first layer to start from instead of running conv1 here.
    stack: name under which the precision plan lists this encoder's layers (engine.MIXED_ONE_PASS)."""
    wc = prog.weights
    bb = enc.feature_backbone
    if isinstance(bb, n_segnet_encoder):
        if h % 32 or w % 32:
            raise ValueError("n_segnet_encoder needs H and W divisible by 32 (got %dx%d)" % (h, w))
        units = bb.units()
        first = 1   # index into units of the next layer to run
        if stem is not None:
            a = stem
        elif prog.can_fuse_head(stack):
            # conv1 + conv2 in one kernel: the 64-channel full-resolution map never reaches HBM (csrc/enc_head.cu)
            a = prog.enc_head(x_nchw, wc.stem(units[0].conv, units[0].bn), wc.conv(units[1].conv, units[1].bn, True),
                              b, n_agents, h, w, c_first)
            first = 2
        else:
            a = prog.stem_bn(x_nchw, units[0].conv, units[0].bn, b, n_agents, h, w, c_first)
        for i, u in enumerate(units[first:], first + 1):
            a = prog.conv_bn(a, u.conv, u.bn, True, passes=prog.passes_for(stack, i))
    elif isinstance(bb, resnet_encoder):
        if h % 32 or w % 32:
            raise ValueError("resnet_encoder needs H and W divisible by 32 (got %dx%d)" % (h, w))
        fb = bb.feature_backbone
        a = stem if stem is not None else prog.stem_bn(x_nchw, fb.conv1, fb.bn1, b, n_agents, h, w, c_first)
        a = prog.maxpool(a)
        for li in range(1, 5):
            for blk in getattr(fb, "layer%d" % li):
                idt = a
                if blk.downsample is not None:
                    idt = prog.conv_bn(a, blk.downsample[0], blk.downsample[1], False)
                y = prog.conv_bn(a, blk.conv1, blk.bn1, True)
                a = prog.conv_bn(y, blk.conv2, blk.bn2, True, residual=idt)
    else:
        raise ValueError("unknown encoder backbone %r" % type(bb).__name__)
    return prog.conv_bn(a, enc.squeezer.conv, enc.squeezer.bn, True, out=out)


def _build_decoder(prog, dec, a):
    """img_decoder.forward (agent.py:80-89) -> fp32 NCHW logits tensor."""
    wc = prog.weights
    if dec.feat_squeezer == 2:
        a = prog.conv_bn(a, dec.desqueezer.conv, dec.desqueezer.bn, True)
    elif dec.feat_squeezer == 4:
        a = prog.conv_bn(a, dec.desqueezer1.conv, dec.desqueezer1.bn, True)
        a = prog.conv_bn(a, dec.desqueezer2.conv, dec.desqueezer2.bn, True)
    od = dec.output_decoder
    if isinstance(od, n_segnet_decoder):
        units = od.units()
        for u in units[:-1]:
            a = prog.conv_bn(a, u.conv, u.bn, True)
        last = units[-1]
        labels = prog.f32_buf(a.n, a.h, a.w, dtype=torch.uint8) if prog.want_labels else None
        prog.labels_out = labels
        if labels is not None and not prog.want_logits:
            # label map only: the arg-max is taken on the accumulators, the fp32 logits never reach HBM
            return prog.conv_bn(a, last.conv, last.bn, True, nchw_out="none", labels=labels)
        logits = prog.f32_buf(a.n, last.conv.out_channels, a.h, a.w)
        prog.conv_bn(a, last.conv, last.bn, True, nchw_out=logits,  # logits pass BN+ReLU too, backbone.py:124
                  labels=labels)
        return logits
    if isinstance(od, simple_decoder):
        y = prog.conv_bn(a, od.pred[0], None, True)
        small = prog.f32_buf(a.n, od.pred[2].out_channels, a.h, a.w)
        prog.conv_bn(y, od.pred[2], None, False, nchw_out=small)
        prog.labels_out = None
        if prog.want_labels and not prog.want_logits:
            # label map only: up-sample and take the arg-max per output pixel in one kernel, nothing else is written
            prog.labels_out = prog.bilinear_argmax(small, 32, prog.f32_buf(a.n, a.h * 32, a.w * 32, dtype=torch.uint8))
            return prog.labels_out
        up = prog.bilinear(small, 32)
        if prog.want_labels:
            prog.labels_out = prog.argmax_labels(up, prog.f32_buf(up.shape[0], up.shape[2], up.shape[3],
                                                                   dtype=torch.uint8))
        return up
    raise ValueError("unknown decoder backbone %r" % type(od).__name__)


def _fused_stems(prog, enc_a, enc_b, x_nchw, b, n_agents, h, w):
    """u_encoder and query_key_net.img_encoder convolve the same pixels with different weights (agent.py:1111,1124):
    run both first layers as ONE 3 -> 128 stem so the image is read and im2col'd once. Returns the two 64-channel
    slices of the shared output buffer, or (None, None) when the backbones are not both n_segnet."""
    ba, bb = enc_a.feature_backbone, enc_b.feature_backbone
    if isinstance(ba, resnet_encoder) and isinstance(bb, resnet_encoder):
        fa, fb = ba.feature_backbone, bb.feature_backbone
        return prog.stem_pair_bn(x_nchw, fa.conv1, fa.bn1, fb.conv1, fb.bn1, b, n_agents, h, w)
    if not (isinstance(ba, n_segnet_encoder) and isinstance(bb, n_segnet_encoder)):
        return None, None
    ua, ub = ba.units()[0], bb.units()[0]
    # two dense 64-channel maps (not one interleaved 128-channel map: each encoder's stride-2 conv would read half of
    # every 256-byte pixel)
    return prog.stem_pair_bn(x_nchw, ua.conv, ua.bn, ub.conv, ub.bn, b, n_agents, h, w)


def _build_policy(prog, pol, x_nchw, b, n_agents, h, w, stem=None):
    """policy_net4.forward (agent.py:134-142) -> ActMap (n_agents*b, s, s, 256)."""
    a = _build_encoder(prog, pol.img_encoder, "img_encoder", x_nchw, b, n_agents, h, w, stem=stem, stack="policy")
    if a.h % 4 or a.w % 4:
        raise ValueError("policy_net4 needs the %dx%d feature map divisible by 4" % (a.h, a.w))
    for i in range(1, 6):
        u = getattr(pol, "conv%d" % i)
        a = prog.conv_bn(a, u.conv, u.bn, True)
    return a


class _Compiled:
    """One compiled forward: the program, its static input and its output buffers."""
    __slots__ = ("prog", "x", "out", "generation")


class _BackwardBridge(torch.autograd.Function):
    """Hands the program's logits to autograd: `loss.backward()` (trainer.py:669) reaches backward() below with the
    gradient of the logits, which is copied into the backward program's static input; the program's launches
    (engine.Program.bprog: BatchNorm / ReLU backward, tensor-core weight and data gradients, attention and MLP-head
    backward) then produce every parameter gradient, returned to autograd so that .grad accumulates as usual."""

    @staticmethod
    def forward(ctx, model, compiled, pred, *params):
        ctx.model, ctx.compiled, ctx.generation = model, compiled, compiled.generation
        ctx.params = params
        return pred.view_as(pred)

    @staticmethod
    def backward(ctx, gpred):
        c = ctx.compiled
        if c.generation != ctx.generation:
            raise RuntimeError("backward() after another forward of the same input shape: the program's buffers hold "
                               "the newer step (call backward before the next forward)")
        prog = c.prog
        with torch.cuda.device(prog.device), torch.no_grad():
            prog.dlogits.copy_(gpred)
            prog.bprog.run(ctx.model._w2c["graphs"])
            grads = prog.param_gradients()
        out = []
        for prm in ctx.params:
            g = grads.get(prm)
            out.append(None if g is None else g.to(prm.dtype).reshape(prm.shape).clone())
        return (None, None, None) + tuple(out)


# ------------------------------------------------------------------------------------------- model base
class _W2CModel(nn.Module):
    def __init__(self):
        super().__init__()
        self._w2c = {"precision": engine.default_precision(), "graphs": engine.use_graphs_default(),
                     "programs": {}, "weights": {}, "clone_outputs": True, "last": None,
                     "io": {"u8": False, "mean": ops.LOADER_MEAN_BGR, "norm": True, "labels": False, "logits": True}}

    # ---- configuration
    def set_precision(self, name):
        """'bf16' / 'fp16' (throughput), 'bf16x3' / 'fp16x3' (fp32-grade parity precision: hi/lo planes, three MMA
        passes) or 'mixed' (the fp16x3 storage with one pass where the error attribution allows it), see
        engine.PRECISIONS."""
        if name not in engine.PRECISIONS:
            raise ValueError("precision must be one of %s" % sorted(engine.PRECISIONS))
        self._w2c["precision"] = name
        return self

    def set_cuda_graphs(self, enabled):
        self._w2c["graphs"] = bool(enabled)
        self._w2c["programs"].clear()
        return self

    def set_clone_outputs(self, enabled):
        """False returns views of the engine's static output buffers (valid until the next forward)."""
        self._w2c["clone_outputs"] = bool(enabled)
        return self

    def set_input_format(self, fmt="f32_nchw", mean_bgr=ops.LOADER_MEAN_BGR, img_norm=True):
        """'f32_nchw' (the reference's forward() input: views already transformed and concatenated on the channel
        axis, (B, 3N, H, W) float32) or 'u8_hwc': the loader's RAW frames, (B, N, H, W, 3) uint8 RGB; the loader
        transform (airsim_loader.py:515-527) and the trainer's cat (trainer.py:651) are then fused into the first
        conv's gather, bit-identical to transforming on the host. n_segnet encoders only."""
        if fmt not in ("f32_nchw", "u8_hwc"):
            raise ValueError("input format must be 'f32_nchw' or 'u8_hwc'")
        self._w2c["io"].update(u8=fmt == "u8_hwc", mean=tuple(float(m) for m in mean_bgr), norm=bool(img_norm))
        self._w2c["programs"].clear()
        return self

    def set_label_output(self, enabled=True, logits=True):
        """Also produce the uint8 arg-max label map (`outputs.data.max(1)[1]`, trainer.py:804) on the device, read with
        last_labels(). logits=False drops the fp32 logits altogether: forward() then returns the (N*B, H, W) uint8
        label map in place of `pred` (what an evaluation loop consumes)."""
        self._w2c["io"].update(labels=bool(enabled), logits=bool(logits) or not enabled)
        self._w2c["programs"].clear()
        return self

    def last_labels(self):
        c = self._w2c["last"]
        if c is None or c.prog.labels_out is None:
            raise RuntimeError("no label map: call set_label_output(True) before forward()")
        return self._ret(c.prog.labels_out)

    def invalidate(self):
        """Drop packed weights and compiled programs (call after mutating parameters in place)."""
        self._w2c["programs"].clear()
        self._w2c["weights"].clear()

    def load_state_dict(self, *args, **kwargs):
        r = super().load_state_dict(*args, **kwargs)
        self.invalidate()
        return r

    def _apply(self, fn, *args, **kwargs):
        r = super()._apply(fn, *args, **kwargs)
        if "_w2c" in self.__dict__:
            self.invalidate()
        return r

    def kernel_launches_per_forward(self):
        """Launches of libw2c kernels in the most recently compiled programs (for bench accounting)."""
        return {k: c.prog.n_launches for k, c in self._w2c["programs"].items()}

    # ---- compile / run
    def _compiled(self, inputs, tag, builder, pre_run=None):
        # self.training (model.train(), trainer.py:659): the TRAIN-MODE FORWARD - every BatchNorm2d uses batch statistics
        # and updates its running statistics in place (csrc/bn_train.cu). It is a forward only: the outputs carry no
        # autograd graph, so loss.backward() on them raises; the backward pass is not part of this path.
        train = bool(self.training)
        # grad: also record the backward pass (the second half of SURVEY 8 f-1) when autograd is listening
        grad = train and torch.is_grad_enabled() and any(p.requires_grad for p in self.parameters())
        if train:
            # train-mode programs read the live parameters; the packed operands of the EVAL programs go stale with the
            # first optimizer step, so the next eval forward (the trainers' validation pass, trainer.py:687-700)
            # re-packs them
            self._w2c["stale_eval"] = True
        elif self._w2c.get("stale_eval"):
            self.invalidate()
            self._w2c["stale_eval"] = False
        if train and (self._w2c["io"]["u8"] or self._w2c["io"]["labels"]):
            raise RuntimeError("train mode takes the fp32 views and returns logits (set_input_format / "
                               "set_label_output are evaluation-path options)")
        if not (torch.is_tensor(inputs) and inputs.is_cuda):
            raise RuntimeError("%s.forward needs a CUDA tensor: there is no CPU fallback" % type(self).__name__)
        io = self._w2c["io"]
        if io["u8"]:
            if inputs.dim() != 5 or inputs.dtype != torch.uint8 or inputs.shape[-1] != 3:
                raise ValueError("expected (B, N, H, W, 3) uint8 frames (input format 'u8_hwc'), got %s %s"
                                 % (inputs.dtype, tuple(inputs.shape)))
        elif inputs.dim() != 4:
            raise ValueError("expected (B, 3*N, H, W) input, got shape %s" % (tuple(inputs.shape),))
        dev = inputs.device
        precision = self._w2c["precision"]
        act = engine.PRECISIONS[precision]
        key = (dev, precision, tuple(inputs.shape), tag, io["u8"], io["mean"], io["norm"], io["labels"], io["logits"], train,
               grad)
        c = self._w2c["programs"].get(key)
        if c is None:
            wkey = (dev, act)
            wc = self._w2c["weights"].get(wkey)
            if wc is None and not train:
                wc = engine.WeightCache(dev, act)
                self._w2c["weights"][wkey] = wc
            with torch.cuda.device(dev), torch.no_grad():
                prog = engine.Program(wc, dev, act)
                if train:
                    # the optimizer changes the parameters between two forwards: a train-mode program re-derives its
                    # packed operands from the live parameters on every run (engine.WeightCache, live mode)
                    prog.weights = engine.WeightCache(dev, act, live=prog)
                if grad:
                    prog.begin_backward()
                prog.pass_plan = engine.MIXED_ONE_PASS if (precision == "mixed" and not train) else None
                prog.train = train
                prog.want_labels, prog.want_logits = io["labels"], io["logits"]
                if io["u8"]:
                    prog.input_u8 = True
                    prog.lut = ops.loader_lut(io["mean"], io["norm"], dev)
                c = _Compiled()
                c.prog = prog
                c.x = prog.f32_buf(*inputs.shape, dtype=torch.uint8 if io["u8"] else torch.float32)
                c.out = builder(prog, c.x)
                c.generation = 0
                if grad:
                    prog.finish_backward(c.out["pred"])
            self._w2c["programs"][key] = c
        with torch.cuda.device(dev), torch.no_grad():
            c.x.copy_(inputs)
            if pre_run is not None:
                pre_run(c)  # per-call device state of a static program (e.g. the drawn selection indices)
            c.prog.run(self._w2c["graphs"])
        c.generation += 1
        self._w2c["last"] = c
        return c

    def _ret(self, t):
        return t.clone() if self._w2c["clone_outputs"] else t

    def _ret_pred(self, c):
        """The logits of a compiled forward; in train mode with autograd listening they carry the backward program."""
        pred = self._ret(c.out["pred"])
        if c.prog.grad and torch.is_grad_enabled():
            params = [p for p in self.parameters() if p.requires_grad]
            pred = _BackwardBridge.apply(self, c, pred, *params)
        return pred


def _bhw(inputs):
    """(B, H, W) of a (B, 3N, H, W) float batch or of (B, N, H, W, 3) uint8 frames."""
    return inputs.shape[0], inputs.shape[2], inputs.shape[3]  # H, W sit at dims 2, 3 in both layouts


def _check_views(inputs, n_agents):
    if inputs.dim() == 5:
        if inputs.shape[1] != n_agents:
            raise ValueError("expected frames of %d agents, got %d" % (n_agents, inputs.shape[1]))
        return
    if inputs.shape[1] != 3 * n_agents:
        raise ValueError("expected %d channels (3 per agent x %d agents), got %d" % (3 * n_agents, n_agents,
                                                                                      inputs.shape[1]))


# ------------------------------------------------------------------------------------------- models
class Single_agent(_W2CModel):
    def __init__(self, n_classes=21, in_channels=3, feat_channel=512, enc_backbone="n_segnet_encoder",
                 dec_backbone="n_segnet_decoder", feat_squeezer=-1):
        super().__init__()
        self.in_channels = in_channels
        self.encoder = img_encoder(n_classes=n_classes, in_channels=in_channels, feat_channel=feat_channel,
                                   feat_squeezer=feat_squeezer, enc_backbone=enc_backbone)
        self.decoder = img_decoder(n_classes=n_classes, in_channels=feat_channel, feat_squeezer=feat_squeezer,
                                   dec_backbone=dec_backbone)

    def forward(self, inputs):
        _check_views(inputs, 1)
        b, h, w = _bhw(inputs)

        def build(prog, x):
            feat = _build_encoder(prog, self.encoder, "encoder", x, b, 1, h, w)
            return {"pred": _build_decoder(prog, self.decoder, feat)}

        return self._ret_pred(self._compiled(inputs, "fwd", build))


class _AttentionModel(_W2CModel):
    """Shared constructor plumbing of the four learned-communication models."""
    # precision-plan stack of the value-map encoder (engine.MIXED_ONE_PASS): "encoder_fused" where the decoder sees
    # only an attention-weighted sum of maps, "encoder" where an agent's own map reaches the decoder as it is
    _value_stack = "encoder"

    def _init_common(self, n_classes, in_channels, feat_channel, feat_squeezer, attention, has_query, sparse,
                     shared_img_encoder, image_size, key_size, query_size, enc_backbone, dec_backbone, head_cls,
                     attention_module, decoder_in):
        self.in_channels = in_channels
        self.feature_map_channel = FEATURE_CHANNELS
        self.key_size = key_size
        self.query_size = query_size
        self.shared_img_encoder = shared_img_encoder
        self.has_query = has_query
        self.sparse = sparse
        self.attention = attention
        self.image_size = image_size
        mk = lambda: img_encoder(n_classes=n_classes, in_channels=in_channels, feat_channel=feat_channel,
                                 feat_squeezer=feat_squeezer, enc_backbone=enc_backbone)
        if shared_img_encoder == "unified" or isinstance(self, MIMOcom):
            if shared_img_encoder != "unified" and self.who:
                raise ValueError("Incorrect shared_img_encoder flag")  # agent.py:1229-1230
            self.u_encoder = mk()                   # (MIMOcom builds it unconditionally, agent.py:1001)
        elif shared_img_encoder == "only_normal_agents":   # agent.py:494-498,697-701 (srms_who2com.yml)
            self.degarded_encoder = mk()
            self.normal_encoder = mk()
        else:                                       # one encoder per agent, agent.py:500-510,703-713
            for i in range(1, 6):
                setattr(self, "encoder%d" % i, mk())
        self.key_net = head_cls(out_size=key_size, input_feat_sz=image_size / 32)
        self.attention_net = attention_module
        self.query_key_net = policy_net4(n_classes=n_classes, in_channels=in_channels, enc_backbone=enc_backbone)
        if has_query:
            self.query_net = head_cls(out_size=query_size, input_feat_sz=image_size / 32)
        self.decoder = img_decoder(n_classes=n_classes, in_channels=decoder_in, feat_squeezer=feat_squeezer,
                                   dec_backbone=dec_backbone)

    # parameter groups the reference trainers read (agent.py:1017-1030)
    @property
    def attention_paras(self):
        return list(self.attention_net.parameters())

    def _value_encoders(self):
        """[(encoder module, first agent, number of agents)] producing the feature maps, in agent order."""
        if hasattr(self, "u_encoder"):
            return None
        if hasattr(self, "degarded_encoder"):
            return [(self.degarded_encoder, 0, 1), (self.normal_encoder, 1, 4)]
        return [(getattr(self, "encoder%d" % (i + 1)), i, 1) for i in range(5)]

    @property
    def img_net_paras(self):
        extra = list(self.argmax_decoder.parameters()) if hasattr(self, "argmax_decoder") else []
        encs = self._value_encoders()
        enc_p = (list(self.u_encoder.parameters()) if encs is None
                 else [p for e, _, _ in encs for p in e.parameters()])
        return enc_p + list(self.decoder.parameters()) + extra

    @property
    def policy_net_paras(self):
        p = list(self.query_key_net.parameters()) + list(self.key_net.parameters()) + self.attention_paras
        if self.has_query:
            p = p + list(self.query_net.parameters())
        return p

    @property
    def all_paras(self):
        return self.img_net_paras + self.policy_net_paras

    def _attn_weights(self, prog):
        if isinstance(self.attention_net, _DotAttention):
            wq = prog.weights.tensor(self.attention_net.linear.weight, ("attn_w", id(self.attention_net)))
            bq = prog.weights.tensor(self.attention_net.linear.bias, ("attn_b", id(self.attention_net)))
            return wq, bq, 1.0
        if isinstance(self.attention_net, _AdditiveAttention):
            return None, None, 1.0
        return None, None, float(self.attention_net.temperature)

    def _additive_queries(self, prog, rows):
        """[rows, 128] fp32 buffer whose every row is u = linear_feat.weight^T linear_out.weight^T (see
        _AdditiveAttention): the stand-in query matrix of the additive attention."""
        a = self.attention_net
        if prog.grad:
            raise NotImplementedError("backward of the additive attention (its projections fold into one fixed query "
                                      "vector here): train with attention='general' or under torch.no_grad()")
        key = ("additive_u", id(a), rows)
        u = prog.weights._misc.get(key)
        if u is None:
            u = torch.empty((rows, 128), dtype=torch.float32, device=prog.device)
            w_out, w_feat = prog.weights._param(a.linear_out.weight), prog.weights._param(a.linear_feat.weight)

            def refresh(u=u, w_out=w_out, w_feat=w_feat):
                u.copy_((w_out @ w_feat).expand(rows, 128))
            prog.weights._emit_torch(refresh)       # (train mode: re-derived from the live parameters every run)
            prog.weights._misc[key] = u
        return u

    def _keys_queries(self, prog, x, b, n, h, w, dst=None, after_values=None):
        """u_encoder features, key and query vectors for the n agents in x (agent.py:1111-1148). dst = (keys,
        queries, val) tensors to write into (the rank's slot of the exchange buffer when agents are sharded).
        after_values(): recorded right after the feature encoder when that chain is NOT forked onto the side stream
        (the sharded forward starts its feature-map all-gather there); returns True if it did."""
        prog.values_hook_ran = False
        val_out = None
        if dst is not None:
            val_out = engine.ActMap(dst[2], n * b, dst[2].shape[1], dst[2].shape[2], dst[2].shape[3] // prog.planes)
        encs = self._value_encoders()
        if encs is None:
            segnet = isinstance(self.u_encoder.feature_backbone, n_segnet_encoder)
            if segnet and prog.can_fuse_head(self._value_stack):
                # each encoder starts with its own fused conv1 + conv2 head kernel (under the "mixed" precision only
                # the value encoder does: the policy net keeps three passes and runs its own first layer)
                stem_u = stem_p = None
            else:
                stem_u, stem_p = _fused_stems(prog, self.u_encoder, self.query_key_net.img_encoder, x, b, n, h, w)
            # the feature encoder runs beside the policy net + heads (independent chains); worth it for the resnet
            # pair's many small launches, not for the n_segnet pair (engine.TWO_STREAMS)
            # ... and for small steps (latency points: at <= 8 agent-frames the <= 64x64 layers of either chain give
            # a 148-SM part 2-32 tiles each, so the two chains fill each other's idle SMs)
            # ... and for training steps: the element-wise BatchNorm passes, the per-launch tails of the weight-gradient
            # kernels and the sub-wave layers of one chain leave room the other chain fills (n_segnet pair, 10 frames:
            # 19.3 -> 18.2 ms per step; the eval-mode step of the same pair loses with two chains)
            small_kernels = (isinstance(self.u_encoder.feature_backbone, resnet_encoder)
                             or n * b * h * w <= engine.TWO_STREAM_MAX_PIXELS or prog.train)
            with prog.side_stream(auto=small_kernels) as forked:
                val = _build_encoder(prog, self.u_encoder, "u_encoder", x, b, n, h, w, out=val_out, stem=stem_u,
                                     stack=self._value_stack)
            if after_values is not None and not forked:
                after_values()
                prog.values_hook_ran = True
        else:
            # separate encoders per agent group (agent.py:579-594,823-838), each writing its agents' images of the
            # agent-major feature buffer
            stem_p = None
            sq = {2: 2, 4: 4}.get(encs[0][0].feat_squeezer, 1)
            val = prog.act_buf(n * b, h // 32 // sq, w // 32 // sq, encs[0][0].squeezer.conv.out_channels)
            for enc, first, count in encs:
                _build_encoder(prog, enc, "enc%d" % first, x, b, count, h, w, c_first=3 * first,
                               out=val.images(first * b, count * b), stack=self._value_stack)
        qk = _build_policy(prog, self.query_key_net, x, b, n, h, w, stem=stem_p)
        if qk.h != qk.w:
            raise ValueError("square inputs only (the reference derives n_feat from image_size alone)")
        heads = [(prog.weights.mlp(self.key_net.fc, qk.h), self.key_size, None if dst is None else dst[0])]
        if self.has_query:
            heads.append((prog.weights.mlp(self.query_net.fc, qk.h), self.query_size,
                          None if dst is None else dst[1]))
            keys, queries = prog.kq_mlp_heads(qk, heads, fcs=(self.key_net.fc, self.query_net.fc))  # one pair of launches
        else:
            keys, = prog.kq_mlp_heads(qk, heads, fcs=(self.key_net.fc,))
            queries = prog.f32_buf(n * b, self.query_size) if dst is None else dst[1]
            queries.fill_(1.0)  # torch.ones(batch, 1, query_size), agent.py:1144
        prog.join()
        return val, keys, queries


_MODES = {"softmax": ops.FUSE_SOFTMAX, "activated": ops.FUSE_ACTIVATED, "argmax_test": ops.FUSE_ARGMAX}


class MIMOcom(_AttentionModel):
    who = False
    _value_stack = "encoder_fused"

    def __init__(self, n_classes=21, in_channels=3, feat_channel=512, feat_squeezer=-1, attention="additive",
                 has_query=True, sparse=False, agent_num=5, shuffle_flag=False, image_size=512,
                 shared_img_encoder=False, key_size=128, query_size=128, enc_backbone="n_segnet_encoder",
                 dec_backbone="n_segnet_decoder"):
        super().__init__()
        self.agent_num = agent_num
        self.shuffle_flag = shuffle_flag
        if agent_num > 8:
            raise ValueError("agent_num must be <= 8")
        head = linear if self.who else km_generator
        self._init_common(n_classes, in_channels, feat_channel, feat_squeezer, attention, has_query, sparse,
                          shared_img_encoder, image_size, key_size, query_size, enc_backbone, dec_backbone, head,
                          _DotAttention(query_size, key_size),
                          FEATURE_CHANNELS * 2 if self.who else FEATURE_CHANNELS)

    def shard_agents(self, group=None, rank=None, world=None):
        """Shard the agents over the ranks of a torch.distributed process group (one process per GPU). After this,
        forward() takes only this rank's views, (B, 3*agent_num/world, H, W) for agents
        [rank*apr, (rank+1)*apr), and returns `pred` for those agents; prob_action / action / num_connect are
        complete and identical on every rank. Pass world=1 (or call unshard_agents) to undo."""
        import torch.distributed as dist
        if world is None:
            world = dist.get_world_size(group)
        if rank is None:
            rank = dist.get_rank(group)
        if self.agent_num % world:
            raise ValueError("agent_num=%d is not divisible by %d ranks" % (self.agent_num, world))
        self._w2c["shard"] = None if world == 1 else (group, rank, world)
        self._w2c["programs"].clear()
        return self

    def unshard_agents(self):
        self._w2c["shard"] = None
        self._w2c["programs"].clear()
        return self

    def forward(self, inputs, training=True, MO_flag=False, inference="argmax"):
        if self.shared_img_encoder != "unified":
            raise ValueError("Incorrect encoder")  # agent.py:1116-1118,1335-1337: only the unified encoder exists
        if self._w2c.get("shard") is not None:
            if self.training and torch.is_grad_enabled():
                raise NotImplementedError("the backward pass is not sharded by agent: train on replicas "
                                          "(DistributedDataParallel over scenes) or under torch.no_grad()")
            return self._forward_sharded(inputs, training, MO_flag, inference)
        n = self.agent_num
        _check_views(inputs, n)
        if training or inference == "softmax":
            mode = "softmax"
        elif inference in _MODES:
            mode = inference
        else:
            raise ValueError("Incorrect inference mode")
        if not MO_flag:
            raise ValueError("MO_flag=False is not runnable in the reference either (agent.py:1153,1164-1165); "
                             "pass MO_flag=True (multiple_output: True in the shipped mrms configs)")
        b, h, w = _bhw(inputs)

        def build(prog, x):
            val, keys, queries = self._keys_queries(prog, x, b, n, h, w)
            wq, bq, temp = self._attn_weights(prog)
            prob = prog.f32_buf(b, n, n)
            coef = prog.f32_buf(b, n, n)
            action = prog.f32_buf(b, n, dtype=torch.int64)
            connect = prog.f32_buf(1, dtype=torch.int32, zero=True)
            if self.who:
                # decoder input = cat(fused, own features) on channels, agent.py:1382
                cat = prog.act_buf(n * b, val.h, val.w, 2 * val.c)
                fused = cat.slice(0, val.c)
                prog.copy_channels(val, cat.slice(val.c, val.c))
            else:
                cat = fused = prog.act_buf(n * b, val.h, val.w, val.c)
            prog.memset(connect)
            prog.attn(keys, queries, wq, bq, val, fused, prob, coef, action, connect, b_sz=b, n_k=n, n_q=n,
                      k_dim=self.key_size, q_dim=self.query_size, mode=_MODES[mode], mask_self=self.who,
                      temperature=temp, diag_bias=0.0 if self.who else 0.001, attn_module=self.attention_net)
            return {"pred": _build_decoder(prog, self.decoder, cat), "prob": prob, "action": action,
                    "connect": connect}

        c = self._compiled(inputs, mode, build)
        out = c.out
        prob = self._ret(out["prob"])
        if self.who:
            action = torch.argmax(prob, dim=1)  # agent.py:1397,1412: always from prob_action
        else:
            action = self._ret(out["action"])
        if mode == "softmax":
            num_connect = n - 1
        else:  # counted by the attention kernel; stays on the device until somebody reads it (lazy.DeviceScalar)
            num_connect = DeviceScalar.from_count(out["connect"], n * b)
        return self._ret_pred(c), prob, action, num_connect


    def _forward_sharded(self, inputs, training, MO_flag, inference):
        from .. import sharding
        group, rank, world = self._w2c["shard"]
        n = self.agent_num
        apr = n // world
        _check_views(inputs, apr)
        if training or inference == "softmax":
            mode = "softmax"
        elif inference in _MODES:
            mode = inference
        else:
            raise ValueError("Incorrect inference mode")
        if not MO_flag:
            raise ValueError("MO_flag=False is not runnable in the reference either; pass MO_flag=True")
        b, h, w = _bhw(inputs)

        def build(prog, x):
            # feature-map geometry as the encoder really produces it: the squeezer may stride (feat_squeezer 2 / 4,
            # agent.py:49-54) and sets the channel count (feat_channel)
            sq = {2: 2, 4: 4}.get(self.u_encoder.feat_squeezer, 1)
            fh, fw = h // 32 // sq, w // 32 // sq
            fc = self.u_encoder.squeezer.conv.out_channels
            lay = sharding.AgentShardLayout(n, world, rank, b, self.key_size, self.query_size, fh, fw, fc,
                                            prog.planes)
            exchange = lay.allocate(prog.device)
            prog.keep.append(exchange)
            k_loc, q_loc, v_loc = lay.views(exchange)
            # The exchange: the feature maps (99 % of the bytes) are gathered asynchronously right after the feature
            # encoder and travel over NVLink while the policy net runs; only the key / query gather sits before the
            # attention. (Chains forked onto two streams - the resnet pair - keep the single gather at the end.)
            pending = {}

            def start_values():
                prog.host_op(lambda: pending.__setitem__(
                    "work", sharding.all_gather_values(exchange, lay, group, async_op=True)))

            self._keys_queries(prog, x, b, apr, h, w, dst=(k_loc, q_loc, v_loc), after_values=start_values)
            if prog.values_hook_ran:
                def finish():
                    sharding.all_gather_keys_queries(exchange, lay, group)
                    pending.pop("work").wait()
                prog.host_op(finish)
            else:
                prog.host_op(lambda: sharding.all_gather_slots(exchange, lay, group), capturable=True)
            k0, q0, v0 = lay.views(exchange, 0)
            val = engine.ActMap(v0, apr * b, fh, fw, fc)
            wq, bq, temp = self._attn_weights(prog)
            prob = prog.f32_buf(b, n, n)
            coef = prog.f32_buf(b, n, n)
            action = prog.f32_buf(b, n, dtype=torch.int64)
            connect = prog.f32_buf(1, dtype=torch.int32, zero=True)
            if self.who:
                cat = prog.act_buf(apr * b, fh, fw, 2 * fc)
                fused = cat.slice(0, fc)
                prog.copy_channels(engine.ActMap(v_loc, apr * b, fh, fw, fc),
                                   cat.slice(fc, fc))
            else:
                cat = fused = prog.act_buf(apr * b, fh, fw, fc)
            prog.memset(connect)
            prog.attn(k0, q0, wq, bq, val, fused, prob, coef, action, connect, b_sz=b, n_k=n, n_q=n,
                      k_dim=self.key_size, q_dim=self.query_size, mode=_MODES[mode], mask_self=self.who,
                      temperature=temp, diag_bias=0.0 if self.who else 0.001, q_first=lay.first_agent, q_count=apr,
                      agents_per_rank=apr, keys_rank_stride=lay.keys_rank_stride,
                      queries_rank_stride=lay.queries_rank_stride, val_rank_stride=lay.val_rank_stride)
            return {"pred": _build_decoder(prog, self.decoder, cat), "prob": prob, "action": action,
                    "connect": connect}

        out = self._compiled(inputs, ("shard", rank, world, mode), build).out
        prob = self._ret(out["prob"])
        action = torch.argmax(prob, dim=1) if self.who else self._ret(out["action"])
        num_connect = n - 1 if mode == "softmax" else DeviceScalar.from_count(out["connect"], n * b)
        return self._ret(out["pred"]), prob, action, num_connect


class MIMOcomWho(MIMOcom):
    who = True
    _value_stack = "encoder"   # decoder input = cat(fused, own map), agent.py:1382


class LearnWhen2Com(_AttentionModel):
    _value_stack = "encoder_fused"

    def __init__(self, n_classes=21, in_channels=3, feat_channel=512, feat_squeezer=-1, attention="additive",
                 has_query=True, sparse=False, aux_agent_num=4, shuffle_flag=False, image_size=512,
                 shared_img_encoder=False, key_size=128, query_size=128, enc_backbone="n_segnet_encoder",
                 dec_backbone="n_segnet_decoder"):
        super().__init__()
        self.aux_agent_num = aux_agent_num
        self.shuffle_flag = shuffle_flag
        self._init_common(n_classes, in_channels, feat_channel, feat_squeezer, attention, has_query, sparse,
                          shared_img_encoder, image_size, key_size, query_size, enc_backbone, dec_backbone, linear,
                          _make_attention(attention, query_size, key_size), FEATURE_CHANNELS)
        # registered (and checkpointed) by the reference but never used by its forward, agent.py:734-735
        self.argmax_decoder = img_decoder(n_classes=n_classes, in_channels=FEATURE_CHANNELS,
                                          agent_num=aux_agent_num + 1, dec_backbone=dec_backbone)

    def forward(self, inputs, training=True, inference="argmax"):
        n = 5  # divide_num hard-coded, agent.py:763
        _check_views(inputs, n)
        if training or inference == "softmax":
            mode = "softmax"
        elif inference in _MODES:
            mode = inference
        else:
            raise ValueError("Incorrect inference mode")
        b, h, w = _bhw(inputs)

        def build(prog, x):
            val, keys, queries = self._keys_queries(prog, x, b, n, h, w)
            wq, bq, temp = self._attn_weights(prog)
            prob = prog.f32_buf(b, n, 1)
            coef = prog.f32_buf(b, n, 1)
            action = prog.f32_buf(b, 1, dtype=torch.int64)
            connect = prog.f32_buf(1, dtype=torch.int32, zero=True)
            fused = prog.act_buf(b, val.h, val.w, val.c)
            prog.memset(connect)
            q_dim = self.query_size
            if isinstance(self.attention_net, _AdditiveAttention):
                queries, q_dim = self._additive_queries(prog, queries.shape[0]), 128
            # one requester (agent 0: the first b rows of the agent-major query matrix), all five supporters
            prog.attn(keys, queries, wq, bq, val, fused, prob, coef, action, connect, b_sz=b, n_k=n, n_q=1,
                      k_dim=self.key_size, q_dim=q_dim, mode=_MODES[mode], sparse=self.sparse,
                      temperature=temp, diag_bias=0.0, attn_module=self.attention_net)
            return {"pred": _build_decoder(prog, self.decoder, fused), "prob": prob, "coef": coef,
                    "action": action, "connect": connect}

        c = self._compiled(inputs, mode, build)
        out = c.out
        pred = self._ret_pred(c)
        prob = out["prob"].transpose(1, 2).clone()  # (B, 1, 5) like attn_orig.transpose(2, 1), agent.py:368
        action = torch.argmax(prob, dim=2)
        if training:
            return pred, prob, action
        if mode == "softmax":
            return pred, prob, action, 4
        num_connect = DeviceScalar.from_count(out["connect"], b)
        if mode == "activated":
            return pred, prob, out["coef"].transpose(1, 2).clone(), num_connect  # action = thresholded weights
        return pred, prob, action, num_connect


class LearnWho2Com(_AttentionModel):
    def __init__(self, n_classes=21, in_channels=3, feat_channel=512, feat_squeezer=-1, attention="additive",
                 has_query=True, sparse=False, aux_agent_num=4, shuffle_flag=False, image_size=512,
                 shared_img_encoder=False, key_size=128, query_size=128, enc_backbone="n_segnet_encoder",
                 dec_backbone="n_segnet_decoder"):
        super().__init__()
        self.aux_agent_num = aux_agent_num
        self.shuffle_flag = shuffle_flag
        self._init_common(n_classes, in_channels, feat_channel, feat_squeezer, attention, has_query, sparse,
                          shared_img_encoder, image_size, key_size, query_size, enc_backbone, dec_backbone, linear,
                          _make_attention(attention, query_size, key_size), FEATURE_CHANNELS * 2)

    def forward(self, inputs, training=True, inference="argmax"):
        n = 5
        _check_views(inputs, n)
        if training or inference == "softmax":
            mode = "softmax"
        elif inference == "argmax_test":
            mode = "argmax_test"
        else:
            raise ValueError("Incorrect inference mode")  # 'argmax_train' needs an undefined argmax_decoder
        b, h, w = _bhw(inputs)

        def build(prog, x):
            val, keys, queries = self._keys_queries(prog, x, b, n, h, w)
            wq, bq, temp = self._attn_weights(prog)
            prob = prog.f32_buf(b, n - 1, 1)
            coef = prog.f32_buf(b, n - 1, 1)
            action = prog.f32_buf(b, 1, dtype=torch.int64)
            cat = prog.act_buf(b, val.h, val.w, 2 * val.c)  # cat(own, aux) on channels, agent.py:623
            prog.copy_channels(val.images(0, b), cat.slice(0, val.c))
            # supporters are agents 1..4 (agent.py:603-614): skip the first b rows / images
            q_dim = self.query_size
            if isinstance(self.attention_net, _AdditiveAttention):
                queries, q_dim = self._additive_queries(prog, queries.shape[0]), 128
            prog.attn(keys[b:], queries, wq, bq, val.images(b, (n - 1) * b), cat.slice(val.c, val.c), prob, coef,
                      action, None, b_sz=b, n_k=n - 1, n_q=1, k_dim=self.key_size, q_dim=q_dim,
                      mode=_MODES[mode], sparse=self.sparse, temperature=temp, diag_bias=0.0,
                      attn_module=self.attention_net)
            return {"pred": _build_decoder(prog, self.decoder, cat), "prob": prob}

        c = self._compiled(inputs, mode, build)
        prob = c.out["prob"].transpose(1, 2).clone()
        return self._ret_pred(c), prob, torch.argmax(prob, dim=2)


class MIMO_All_agents(_W2CModel):
    def __init__(self, n_classes=21, in_channels=3, feat_channel=512, aux_agent_num=4, shuffle_flag=False,
                 enc_backbone="n_segnet_encoder", dec_backbone="n_segnet_decoder", feat_squeezer=-1):
        super().__init__()
        self.agent_num = aux_agent_num
        self.in_channels = in_channels
        self.shuffle_flag = shuffle_flag
        self.encoder = img_encoder(n_classes=n_classes, in_channels=in_channels, feat_channel=feat_channel,
                                   feat_squeezer=feat_squeezer, enc_backbone=enc_backbone)
        width = 2 if shuffle_flag in ("selection", "ComNet") else self.agent_num
        self.decoder = img_decoder(n_classes=n_classes, in_channels=feat_channel * width,
                                   feat_squeezer=feat_squeezer, dec_backbone=dec_backbone)

    def forward(self, inputs):
        n = self.agent_num
        _check_views(inputs, n)
        b, h, w = _bhw(inputs)
        if self.shuffle_flag == "selection":
            return self._forward_selection(inputs, n, b, h, w)

        def build(prog, x):
            feat = _build_encoder(prog, self.encoder, "encoder", x, b, n, h, w)
            if self.shuffle_flag == "ComNet":
                # cat(own, mean of the other agents' maps), agent.py:948-961: the attention kernel with all-equal
                # scores and the diagonal masked gives exactly the 1/(n-1) weights
                cat = prog.act_buf(n * b, feat.h, feat.w, 2 * feat.c)
                prog.gather_images(feat, cat.slice(0, feat.c), b, n)
                zeros = prog.f32_buf(n * b, 8, zero=True)
                prob = prog.f32_buf(b, n, n)
                action = prog.f32_buf(b, n, dtype=torch.int64)
                prog.attn(zeros, zeros, None, None, feat, cat.slice(feat.c, feat.c), prob, None, action, None, b_sz=b,
                          n_k=n, n_q=n, k_dim=8, q_dim=8, mode=ops.FUSE_SOFTMAX, mask_self=True)
                return {"pred": _build_decoder(prog, self.decoder, cat)}
            cat = prog.act_buf(n * b, feat.h, feat.w, n * feat.c)
            for i in range(n):          # agent i decodes cat_j feat[(i + j) % n], agent.py:963-971
                for j in range(n):
                    src = feat.images(((i + j) % n) * b, b)
                    prog.copy_channels(src, cat.images(i * b, b).slice(j * feat.c, feat.c))
            return {"pred": _build_decoder(prog, self.decoder, cat)}

        return self._ret_pred(self._compiled(inputs, "fwd", build))

    def _forward_selection(self, inputs, n, b, h, w):
        """Random-selection baseline, agent.py:934-947: agent i decodes cat(own map, map of a randomly drawn agent).
        The draws come from Python's `random` exactly like the reference (same call order, so the same seed gives the
        same selection); they are copied into a device index array that the captured program reads."""
        picks = [random.randint(0, n - 1) for _ in range(n)]

        def build(prog, x):
            feat = _build_encoder(prog, self.encoder, "encoder", x, b, n, h, w)
            cat = prog.act_buf(n * b, feat.h, feat.w, 2 * feat.c)
            sel = prog.f32_buf(n, dtype=torch.int32)
            prog.gather_images(feat, cat.slice(0, feat.c), b, n)
            prog.gather_images(feat, cat.slice(feat.c, feat.c), b, n, sel=sel)
            return {"pred": _build_decoder(prog, self.decoder, cat), "sel": sel}

        def pre_run(c):
            c.out["sel"].copy_(torch.tensor(picks, dtype=torch.int32))

        out = self._compiled(inputs, "selection", build, pre_run).out
        action = torch.tensor(picks, dtype=torch.long, device=inputs.device).view(1, n).expand(b, n).contiguous()
        return self._ret(out["pred"]), action


class All_agents(_W2CModel):
    def __init__(self, n_classes=21, in_channels=3, feat_channel=512, aux_agent_num=4, shuffle_flag=False,
                 enc_backbone="n_segnet_encoder", dec_backbone="n_segnet_decoder", feat_squeezer=-1):
        super().__init__()
        self.agent_num = aux_agent_num
        self.in_channels = in_channels
        self.shuffle_flag = shuffle_flag
        for i in range(1, 6):
            setattr(self, "encoder%d" % i, img_encoder(n_classes=n_classes, in_channels=in_channels,
                                                       feat_channel=feat_channel, feat_squeezer=feat_squeezer,
                                                       enc_backbone=enc_backbone))
        width = 2 if shuffle_flag == "selection" else self.agent_num
        self.decoder = img_decoder(n_classes=n_classes, in_channels=feat_channel * width,
                                   feat_squeezer=feat_squeezer, dec_backbone=dec_backbone)

    def forward(self, inputs):
        _check_views(inputs, 5)  # divide_num hard-coded, agent.py:433
        b, h, w = _bhw(inputs)
        fc = self.encoder1.squeezer.conv.out_channels
        if self.shuffle_flag == "selection":
            return self._forward_selection(inputs, b, h, w, fc)
        used = 2 if self.shuffle_flag == "fixed2" else 5
        if self.decoder.output_decoder.in_channels != used * fc and self.decoder.feat_squeezer not in (2, 4):
            raise ValueError("decoder expects %d input channels but %d feature maps of %d channels are concatenated"
                             % (self.decoder.output_decoder.in_channels, used, fc))

        def build(prog, x):
            hh, ww = h // 32, w // 32
            sq = {2: 2, 4: 4}.get(self.encoder1.feat_squeezer, 1)
            hh, ww = hh // sq, ww // sq
            cat = prog.act_buf(b, hh, ww, used * fc)
            for i in range(used):       # five separate encoders write straight into their concat slice
                _build_encoder(prog, getattr(self, "encoder%d" % (i + 1)), "encoder%d" % (i + 1), x, b, 1, h, w,
                               c_first=3 * i, out=cat.slice(i * fc, fc))
            return {"pred": _build_decoder(prog, self.decoder, cat)}

        return self._ret_pred(self._compiled(inputs, "fwd", build))

    def _forward_selection(self, inputs, b, h, w, fc):
        """Random-selection baseline, agent.py:447-452,466-467: the requester decodes cat(own map, map of ONE randomly
        drawn agent - possibly itself); one `random.randint(0, 4)` per forward like the reference."""
        aux_id = random.randint(0, 4)

        def build(prog, x):
            hh, ww = h // 32, w // 32
            sq = {2: 2, 4: 4}.get(self.encoder1.feat_squeezer, 1)
            hh, ww = hh // sq, ww // sq
            feats = prog.act_buf(5 * b, hh, ww, fc)          # agent-major: all five encoders run, like the reference
            for i in range(5):
                _build_encoder(prog, getattr(self, "encoder%d" % (i + 1)), "encoder%d" % (i + 1), x, b, 1, h, w,
                               c_first=3 * i, out=feats.images(i * b, b))
            cat = prog.act_buf(b, hh, ww, 2 * fc)
            sel = prog.f32_buf(1, dtype=torch.int32)
            prog.gather_images(feats, cat.slice(0, fc), b, 1)
            prog.gather_images(feats, cat.slice(fc, fc), b, 1, sel=sel)
            return {"pred": _build_decoder(prog, self.decoder, cat), "sel": sel}

        def pre_run(c):
            c.out["sel"].fill_(aux_id)

        out = self._compiled(inputs, "selection", build, pre_run).out
        action = torch.full((b,), aux_id, dtype=torch.long, device=inputs.device)
        return self._ret(out["pred"]), action
