"""CPU ORACLE — test infrastructure, NOT part of the product.

A functional fp32 restatement (torch CPU ops, no nn.Module, no CUDA) of the When2com forward path of
GT-RIPL/MultiAgentPerception, driven directly from a reference-format state_dict. Every function cites the
reference file:line it follows (paths are into /root/reference/ptsemseg/models/).

Pinning: tests/golden/make_golden.py imports the UNMODIFIED reference modules in the build container, runs them on
seeded inputs/weights and commits the outputs under tests/golden/; tests/test_oracle.py checks this restatement
against those vectors (and, when /root/reference is present, against the live reference). The reference ships no
tests or golden vectors of its own, so that is the only pin available.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this module.
"""
import math
import random

import numpy as np
import torch
import torch.nn.functional as F

BN_EPS = 1e-5  # nn.BatchNorm2d default, utils.py:112


# --------------------------------------------------------------------------------------------- blocks
# Train-mode BatchNorm (model.train(), trainer.py:659): forward(..., train_stats={}) switches every BatchNorm2d to batch
# statistics (biased variance for the normalisation, utils.py:110-114 -> nn.BatchNorm2d) and collects the buffers the
# module would hold afterwards - running_mean / running_var updated with momentum 0.1 and the UNBIASED variance,
# num_batches_tracked + 1 - under their state_dict keys.
_TRAIN_STATS = None


def _bn(x, sd, p):
    if _TRAIN_STATS is None:
        return F.batch_norm(x, sd[p + ".running_mean"], sd[p + ".running_var"], sd[p + ".weight"], sd[p + ".bias"],
                            False, 0.1, BN_EPS)
    rm = _TRAIN_STATS.get(p + ".running_mean", sd[p + ".running_mean"]).clone()
    rv = _TRAIN_STATS.get(p + ".running_var", sd[p + ".running_var"]).clone()
    y = F.batch_norm(x, rm, rv, sd[p + ".weight"], sd[p + ".bias"], True, 0.1, BN_EPS)
    _TRAIN_STATS[p + ".running_mean"], _TRAIN_STATS[p + ".running_var"] = rm, rv
    _TRAIN_STATS[p + ".num_batches_tracked"] = _TRAIN_STATS.get(p + ".num_batches_tracked", 0) + 1
    return y


def cbr(x, sd, p, stride=1):
    """conv2DBatchNormRelu.forward, utils.py:87-120: Conv2d(k3, bias) -> BatchNorm2d(eval) -> ReLU."""
    y = F.conv2d(x, sd[p + ".cbr_unit.0.weight"], sd[p + ".cbr_unit.0.bias"], stride=stride, padding=1)
    return F.relu(_bn(y, sd, p + ".cbr_unit.1"))


def dcbr(x, sd, p):
    """deconv2DBatchNormRelu.forward, utils.py:148-168: ConvTranspose2d(k3 s2 p1 op1) -> BN -> ReLU."""
    y = F.conv_transpose2d(x, sd[p + ".dcbr_unit.0.weight"], sd[p + ".dcbr_unit.0.bias"], stride=2, padding=1,
                           output_padding=1)
    return F.relu(_bn(y, sd, p + ".dcbr_unit.1"))


# --------------------------------------------------------------------------------------------- backbones
_SEGNET_ENC_STRIDES = (1, 2, 1, 2, 1, 1, 2, 1, 1, 2, 1, 1, 2)  # backbone.py:19-39


def n_segnet_encoder(x, sd, p):
    """backbone.py:41-55."""
    for i, s in enumerate(_SEGNET_ENC_STRIDES):
        x = cbr(x, sd, "%s.conv%d" % (p, i + 1), s)
    return x


def _basic_block(x, sd, p, stride, down):
    """torchvision BasicBlock as instantiated by pretrainedmodels.resnet18 (backbone.py:63)."""
    idt = x
    y = F.conv2d(x, sd[p + ".conv1.weight"], None, stride=stride, padding=1)
    y = F.relu(_bn(y, sd, p + ".bn1"))
    y = F.conv2d(y, sd[p + ".conv2.weight"], None, padding=1)
    y = _bn(y, sd, p + ".bn2")
    if down:
        idt = _bn(F.conv2d(x, sd[p + ".downsample.0.weight"], None, stride=stride), sd, p + ".downsample.1")
    return F.relu(y + idt)


def resnet_encoder(x, sd, p):
    """backbone.py:72-96: conv1 -> (bn1, relu, maxpool, layer1) -> layer2 -> layer3 -> layer4 of resnet18."""
    q = p + ".feature_backbone"
    x = F.conv2d(x, sd[q + ".conv1.weight"], None, stride=2, padding=3)
    x = F.relu(_bn(x, sd, q + ".bn1"))
    x = F.max_pool2d(x, 3, 2, 1)
    for li in range(1, 5):
        for bi in range(2):
            first = bi == 0 and li > 1
            x = _basic_block(x, sd, "%s.layer%d.%d" % (q, li, bi), 2 if first else 1, first)
    return x


def img_encoder(x, sd, p, enc_backbone, feat_squeezer=-1):
    """img_encoder.forward, agent.py:56-60 (squeezer stride per agent.py:49-54)."""
    if enc_backbone == "n_segnet_encoder":
        x = n_segnet_encoder(x, sd, p + ".feature_backbone")
    elif enc_backbone == "resnet_encoder":
        x = resnet_encoder(x, sd, p + ".feature_backbone")
    else:
        raise ValueError("Encoder {} not available".format(enc_backbone))
    stride = {2: 2, 4: 4}.get(feat_squeezer, 1)
    y = F.conv2d(x, sd[p + ".squeezer.cbr_unit.0.weight"], sd[p + ".squeezer.cbr_unit.0.bias"], stride=stride, padding=1)
    return F.relu(_bn(y, sd, p + ".squeezer.cbr_unit.1"))


_SEGNET_DEC = ("d", "c", "c", "d", "c", "c", "d", "c", "d", "c", "d", "c")  # backbone.py:105-124


def n_segnet_decoder(x, sd, p):
    """backbone.py:126-140. Note the logits layer deconv12 is conv+BN+ReLU too (backbone.py:124)."""
    for i, kind in enumerate(_SEGNET_DEC):
        name = "%s.deconv%d" % (p, i + 1)
        x = dcbr(x, sd, name) if kind == "d" else cbr(x, sd, name)
    return x


def simple_decoder(x, sd, p):
    """backbone.py:156-164: conv3x3 -> ReLU -> conv3x3 -> bilinear x32 (align_corners=False)."""
    y = F.relu(F.conv2d(x, sd[p + ".pred.0.weight"], sd[p + ".pred.0.bias"], padding=1))
    y = F.conv2d(y, sd[p + ".pred.2.weight"], sd[p + ".pred.2.bias"], padding=1)
    return F.interpolate(y, size=(x.shape[2] * 32, x.shape[3] * 32), mode="bilinear", align_corners=False)


def img_decoder(x, sd, p, dec_backbone, feat_squeezer=-1):
    """img_decoder.forward, agent.py:80-89."""
    if feat_squeezer == 2:
        x = dcbr(x, sd, p + ".desqueezer")
    elif feat_squeezer == 4:
        x = dcbr(dcbr(x, sd, p + ".desqueezer1"), sd, p + ".desqueezer2")
    if dec_backbone == "n_segnet_decoder":
        return n_segnet_decoder(x, sd, p + ".output_decoder")
    if dec_backbone == "simple_decoder":
        return simple_decoder(x, sd, p + ".output_decoder")
    raise ValueError("Decoder {} not available".format(dec_backbone))


def policy_net4(x, sd, p, enc_backbone):
    """policy_net4.forward, agent.py:134-142: its own img_encoder, then 5 conv-BN-ReLU (strides 1,1,2,1,2)."""
    y = img_encoder(x, sd, p + ".img_encoder", enc_backbone)
    for i, s in enumerate((1, 1, 2, 1, 2)):
        y = cbr(y, sd, "%s.conv%d" % (p, i + 1), s)
    return y


def km_generator(x, sd, p):
    """km_generator.forward / linear.forward, agent.py:157-159,176-178 (flatten is NCHW order)."""
    n_feat = sd[p + ".fc.0.weight"].shape[1]
    y = x.reshape(-1, n_feat)
    y = F.relu(F.linear(y, sd[p + ".fc.0.weight"], sd[p + ".fc.0.bias"]))
    y = F.relu(F.linear(y, sd[p + ".fc.2.weight"], sd[p + ".fc.2.bias"]))
    return F.linear(y, sd[p + ".fc.4.weight"], sd[p + ".fc.4.bias"])


# --------------------------------------------------------------------------------------------- attention
def sparsemax_dim1(x):
    """Sparsemax(dim=1).forward, utils.py:834-877, for input (B, K, 1) as the single-request models call it."""
    orig = x.shape
    z = x.reshape(-1, x.shape[1])
    z = z - z.max(dim=1, keepdim=True)[0]
    zs = torch.sort(z, dim=1, descending=True)[0]
    rng = torch.arange(1, z.shape[1] + 1, dtype=z.dtype).view(1, -1)
    is_gt = (1 + rng * zs > torch.cumsum(zs, 1)).to(z.dtype)
    k = (is_gt * rng).max(dim=1, keepdim=True)[0]
    tau = ((is_gt * zs).sum(dim=1, keepdim=True) - 1) / k
    return torch.clamp(z - tau, min=0).reshape(orig)


def fuse(coef, vals):
    """out[b, j] = sum_i coef[b, i, j] * vals[b, i]  (agent.py:278-284)."""
    return torch.einsum("bij,bichw->bjchw", coef, vals)


def mimo_attention(q, k, v, sd, p, mask_self=False):
    """MIMOGeneralDotProductAttention.forward, agent.py:252-286 (mask_self: MIMOWho..., agent.py:306-343)."""
    qt = F.linear(q, sd[p + ".linear.weight"], sd[p + ".linear.bias"])
    s = torch.bmm(k, qt.transpose(2, 1))  # (B, n_key, n_query), no scaling
    if mask_self:
        eye = torch.eye(s.shape[1], s.shape[2], dtype=torch.bool)
        pmat = torch.softmax(s.masked_fill(eye, float("-inf")), dim=1)
    else:
        pmat = torch.softmax(s, dim=1)
    return fuse(pmat, v), pmat


def single_request_attention(q, k, v, sd, p, attention, sparse):
    """GeneralDotProductAttention (agent.py:355-368) / ScaledDotProductAttention (agent.py:203-213), one query."""
    if attention == "general":
        qt = F.linear(q, sd[p + ".linear.weight"], sd[p + ".linear.bias"])
        s = torch.bmm(k, qt.transpose(2, 1))
    elif attention == "additive":
        # AdditiveAttentin.forward, agent.py:225-240: linear_out(linear_feat(k) + linear_context(q))
        t1 = F.linear(k, sd[p + ".linear_feat.weight"], sd[p + ".linear_feat.bias"])
        t2 = F.linear(q, sd[p + ".linear_context.weight"], sd[p + ".linear_context.bias"])
        s = F.linear(t1 + t2, sd[p + ".linear_out.weight"], sd[p + ".linear_out.bias"])  # (B, n_key, 1)
    else:
        s = torch.bmm(k, q.transpose(2, 1)) / (128 ** 0.5)
    a = sparsemax_dim1(s) if sparse else torch.softmax(s, dim=1)  # (B, n_key, 1)
    return fuse(a, v)[:, 0], a.transpose(2, 1)


# --------------------------------------------------------------------------------------------- models
def _cfg(cfg):
    m = cfg["model"]
    return m, m["enc_backbone"], m["dec_backbone"]


def _split_agents(x, n):
    """divide_inputs + cat(dim 0): (B, 3n, H, W) -> (n*B, 3, H, W), agent-major (agent.py:1088-1108)."""
    return torch.cat([x[:, 3 * i:3 * i + 3] for i in range(n)], 0)


def single_agent_forward(sd, cfg, x):
    """Single_agent.forward, agent.py:392-395."""
    m, enc, dec = _cfg(cfg)
    f = img_encoder(x, sd, "encoder", enc, m["feat_squeezer"])
    return img_decoder(f, sd, "decoder", dec, m["feat_squeezer"])


def _activated(p, thres=0.2):
    return p * (p > thres).to(p.dtype)


def _connect(coef):
    """off-diagonal non-zeros / (N * B), agent.py:1052-1056,1073-1077."""
    b, nk, nq = coef.shape
    off = ~torch.eye(nk, nq, dtype=torch.bool)
    return int(((coef != 0) & off).sum()) / (nk * b)


def mimocom_forward(sd, cfg, x, training=True, MO_flag=True, inference="argmax", who=False):
    """MIMOcom.forward (agent.py:1098-1204) and MIMOcomWho.forward (agent.py:1317-1423)."""
    m, enc, dec = _cfg(cfg)
    n = m["agent_num"]
    if m["shared_img_encoder"] != "unified":
        raise ValueError("Incorrect encoder")
    if not MO_flag:
        raise ValueError("MO_flag=False is broken in the reference (SURVEY appendix C); not a parity target")
    b = x.shape[0]
    imgs = _split_agents(x, n)
    feat = img_encoder(imgs, sd, "u_encoder", enc)
    val = feat.view(n, b, *feat.shape[1:]).transpose(0, 1)  # (B, N, C, h, w)
    qk = policy_net4(imgs, sd, "query_key_net", enc)
    keys = km_generator(qk, sd, "key_net")
    key_mat = keys.view(n, b, -1).transpose(0, 1)
    if m["query"]:
        query_mat = km_generator(qk, sd, "query_net").view(n, b, -1).transpose(0, 1)
    else:
        query_mat = torch.ones(b, n, m["query_size"])
    fused, prob = mimo_attention(query_mat, key_mat, val, sd, "attention_net", mask_self=who)

    def decode(f):
        if who:
            f = torch.cat((f, val), dim=2)  # agent.py:1382
        fm = f.transpose(0, 1).reshape(n * b, *f.shape[2:])  # agents2batch, agent.py:1080-1086
        return img_decoder(fm, sd, "decoder", dec)

    if not who:
        prob = prob + 0.001 * torch.eye(n).view(1, n, n)  # agent.py:1164-1167 (after the fusion)
    if training or inference == "softmax":
        return decode(fused), prob, torch.argmax(prob, dim=1), n - 1
    if inference == "argmax_test":
        coef = F.one_hot(prob.max(dim=1)[1], n).to(prob.dtype).transpose(1, 2)
        action = torch.argmax(prob if who else coef, dim=1)
        return decode(fuse(coef, val)), prob, action, _connect(coef)
    if inference == "activated":
        coef = _activated(prob)
        action = torch.argmax(prob if who else coef, dim=1)
        return decode(fuse(coef, val)), prob, action, _connect(coef)
    raise ValueError("Incorrect inference mode")


def _value_maps(x, sd, m, enc, n, b):
    """Feature maps of the n agents, (B, N, C, h, w), by encoder-sharing mode (agent.py:569-594,814-838)."""
    mode = m["shared_img_encoder"]
    if mode == "unified":
        feat = img_encoder(_split_agents(x, n), sd, "u_encoder", enc, m["feat_squeezer"])
        return feat.view(n, b, *feat.shape[1:]).transpose(0, 1)
    if mode == "only_normal_agents":
        f0 = img_encoder(x[:, 0:3], sd, "degarded_encoder", enc, m["feat_squeezer"])
        rest = img_encoder(torch.cat([x[:, 3 * i:3 * i + 3] for i in range(1, n)], 0), sd, "normal_encoder", enc,
                           m["feat_squeezer"])
        return torch.cat((f0.unsqueeze(1), rest.view(n - 1, b, *rest.shape[1:]).transpose(0, 1)), 1)
    feats = [img_encoder(x[:, 3 * i:3 * i + 3], sd, "encoder%d" % (i + 1), enc, m["feat_squeezer"]) for i in range(n)]
    return torch.stack(feats, 1)


def learn_when2com_forward(sd, cfg, x, training=True, inference="argmax"):
    """LearnWhen2Com.forward, agent.py:811-889 (5 agents hard-coded, agent.py:763)."""
    m, enc, dec = _cfg(cfg)
    n = 5
    b = x.shape[0]
    imgs = _split_agents(x, n)
    val = _value_maps(x, sd, m, enc, n, b)
    qk = policy_net4(imgs, sd, "query_key_net", enc)
    keys = km_generator(qk, sd, "key_net").view(n, b, -1).transpose(0, 1)
    if m["query"]:
        query = km_generator(qk, sd, "query_net").view(n, b, -1).transpose(0, 1)[:, :1]
    else:
        query = torch.ones(b, 1, m["query_size"])
    aux, prob = single_request_attention(query, keys, val, sd, "attention_net", m["attention"], m["sparse"])
    if training:
        return img_decoder(aux, sd, "decoder", dec), prob, torch.argmax(prob, dim=2)
    if inference == "softmax":
        return img_decoder(aux, sd, "decoder", dec), prob, torch.argmax(prob, dim=2), 4
    if inference == "argmax_test":
        action = torch.argmax(prob, dim=2)
        sel = val[torch.arange(b), action[:, 0]]
        return img_decoder(sel, sd, "decoder", dec), prob, action, float((action[:, 0] != 0).sum()) / b
    if inference == "activated":
        act = _activated(prob)  # (B, 1, 5)
        f = fuse(act.transpose(1, 2), val)[:, 0]
        return img_decoder(f, sd, "decoder", dec), prob, act, int((act[:, :, 1:] != 0).sum()) / b
    raise ValueError("Incorrect inference mode")


def learn_who2com_forward(sd, cfg, x, training=True, inference="argmax"):
    """LearnWho2Com.forward, agent.py:565-673: query from agent 0, keys/values from agents 1..4, decoder on
    cat(own, aux)."""
    m, enc, dec = _cfg(cfg)
    n = 5
    b = x.shape[0]
    imgs = _split_agents(x, n)
    val = _value_maps(x, sd, m, enc, n, b)
    qk = policy_net4(imgs, sd, "query_key_net", enc)
    keys = km_generator(qk, sd, "key_net").view(n, b, -1).transpose(0, 1)[:, 1:]
    if m["query"]:
        query = km_generator(qk[:b], sd, "query_net").unsqueeze(1)
    else:
        query = torch.ones(b, 1, m["query_size"])
    aux, prob = single_request_attention(query, keys, val[:, 1:], sd, "attention_net", m["attention"], m["sparse"])
    action = torch.argmax(prob, dim=2)
    if training or inference == "softmax":
        return img_decoder(torch.cat((val[:, 0], aux), 1), sd, "decoder", dec), prob, action
    if inference == "argmax_test":
        sel = val[torch.arange(b), action[:, 0] + 1]
        return img_decoder(torch.cat((val[:, 0], sel), 1), sd, "decoder", dec), prob, action
    raise ValueError("Incorrect inference mode")


def mimo_all_agents_forward(sd, cfg, x):
    """MIMO_All_agents.forward catall branch, agent.py:924-981: each agent decodes the rotation-concatenated maps."""
    m, enc, dec = _cfg(cfg)
    n = m["agent_num"]
    b = x.shape[0]
    feat = img_encoder(_split_agents(x, n), sd, "encoder", enc, m["feat_squeezer"])
    fm = feat.view(n, b, *feat.shape[1:])
    if m["shuffle_features"] == "selection":
        # agent.py:934-947: one random.randint(0, n-1) per agent, in agent order (Python's global `random`)
        picks = [random.randint(0, n - 1) for _ in range(n)]
        rows = [torch.cat((fm[i], fm[picks[i]]), 1) for i in range(n)]
        action = torch.tensor(picks, dtype=torch.long).view(1, n).expand(b, n).contiguous()
        return img_decoder(torch.cat(rows, 0), sd, "decoder", dec, m["feat_squeezer"]), action
    if m["shuffle_features"] == "ComNet":
        # agent.py:948-961: cat(own, mean of the other agents' maps)
        rows = []
        for i in range(n):
            other = torch.zeros_like(fm[0])
            for j in range(n):
                if j != i:
                    other = other + fm[j]
            rows.append(torch.cat((fm[i], other / (n - 1)), 1))
        return img_decoder(torch.cat(rows, 0), sd, "decoder", dec, m["feat_squeezer"])
    rows = [torch.cat([fm[(i + j) % n] for j in range(n)], 1) for i in range(n)]
    return img_decoder(torch.cat(rows, 0), sd, "decoder", dec, m["feat_squeezer"])


def all_agents_forward(sd, cfg, x):
    """All_agents.forward catall / fixed2 branches, agent.py:437-469 (five separate encoders)."""
    m, enc, dec = _cfg(cfg)
    feats = [img_encoder(x[:, 3 * i:3 * i + 3], sd, "encoder%d" % (i + 1), enc, m["feat_squeezer"]) for i in range(5)]
    if m["shuffle_features"] == "selection":
        # agent.py:447-452,466-467: one random.randint(0, 4); the requester decodes cat(own, drawn agent's map)
        aux_id = random.randint(0, 4)
        pred = img_decoder(torch.cat((feats[0], feats[aux_id]), 1), sd, "decoder", dec, m["feat_squeezer"])
        return pred, torch.ones(feats[0].shape[0], dtype=torch.long) * aux_id
    if m["shuffle_features"] == "fixed2":
        feats = feats[:2]
    return img_decoder(torch.cat(feats, 1), sd, "decoder", dec, m["feat_squeezer"])


def forward(sd, cfg, x, train_stats=None, **kw):
    """Dispatch on cfg['model']['arch'] like ptsemseg.models.get_model (models/__init__.py:8-101). train_stats: a dict
    -> the model.train() forward (batch-statistics BatchNorm); it receives the updated BatchNorm buffers."""
    global _TRAIN_STATS
    arch = cfg["model"]["arch"]
    sd = {k: v.detach().to(torch.float32).cpu() for k, v in sd.items() if torch.is_floating_point(v)}
    x = x.detach().to(torch.float32).cpu()
    _TRAIN_STATS = train_stats
    try:
        return _dispatch(arch, sd, cfg, x, kw)
    finally:
        _TRAIN_STATS = None


_GRAD_MODE = False   # forward_with_grads() switches autograd on for one forward


def _dispatch(arch, sd, cfg, x, kw):
    with torch.set_grad_enabled(_GRAD_MODE):
        if arch == "Single_agent":
            return single_agent_forward(sd, cfg, x)
        if arch == "MIMOcom":
            return mimocom_forward(sd, cfg, x, **kw)
        if arch == "MIMOcomWho":
            return mimocom_forward(sd, cfg, x, who=True, **kw)
        if arch == "LearnWhen2Com":
            return learn_when2com_forward(sd, cfg, x, **kw)
        if arch == "LearnWho2Com":
            return learn_who2com_forward(sd, cfg, x, **kw)
        if arch == "MIMO_All_agents":
            return mimo_all_agents_forward(sd, cfg, x)
        if arch == "All_agents":
            return all_agents_forward(sd, cfg, x)
    raise ValueError("Model {} not available".format(arch))


def cross_entropy2d(logits, target):
    """ptsemseg/loss/loss.py:5-18 for equal input / target sizes: mean cross-entropy over the pixels, ignore_index 250."""
    n, c, h, w = logits.shape
    flat = logits.permute(0, 2, 3, 1).reshape(-1, c)
    return F.cross_entropy(flat, target.reshape(-1), ignore_index=250)


def forward_with_grads(sd, cfg, x, labels, **kw):
    """One training step's forward + backward as Trainer_*.train() runs it (trainer.py:659-670): model.train();
    outputs = model(images, training=True, ...); loss = cross_entropy2d(outputs, labels); loss.backward().
    Returns (outputs, loss, {state_dict key: gradient}) - the gradients autograd leaves on the reference's parameters
    (keys the forward does not touch are absent). Test infrastructure like the rest of this file."""
    global _TRAIN_STATS, _GRAD_MODE
    arch = cfg["model"]["arch"]
    leaves = {}
    for k, v in sd.items():
        if not torch.is_floating_point(v):
            continue
        t = v.detach().to(torch.float32).cpu().clone()
        if not (k.endswith("running_mean") or k.endswith("running_var")):
            t.requires_grad_(True)
        leaves[k] = t
    x = x.detach().to(torch.float32).cpu()
    _TRAIN_STATS, _GRAD_MODE = {}, True
    try:
        out = _dispatch(arch, leaves, cfg, x, kw)
    finally:
        _TRAIN_STATS, _GRAD_MODE = None, False
    pred = out[0] if isinstance(out, tuple) else out
    loss = cross_entropy2d(pred, labels.cpu())
    loss.backward()
    grads = {k: t.grad for k, t in leaves.items() if t.requires_grad and t.grad is not None}
    detach = lambda o: o.detach() if torch.is_tensor(o) else o
    out = tuple(detach(o) for o in out) if isinstance(out, tuple) else out.detach()
    return out, float(loss.detach()), grads


# --------------------------------------------------------------------------------------------- loader / eval glue
LOADER_MEAN = (103.939, 116.779, 123.68)  # airsim_loader.py:191, applied per channel AFTER the BGR flip


def loader_transform(img_rgb_u8, mean=LOADER_MEAN, img_norm=True):
    """airsimLoader.transform, airsim_loader.py:515-535, image half: one (H, W, 3) uint8 RGB frame ->
    (3, H, W) float32 (BGR, mean-subtracted in float64, /255)."""
    img = np.asarray(img_rgb_u8)[:, :, ::-1]      # :521
    img = img.astype(np.float64)                  # :522
    img = img - np.asarray(mean, dtype=np.float64)  # :523
    if img_norm:
        img = img.astype(float) / 255.0           # :524-525
    img = img.transpose(2, 0, 1)                  # :527
    return torch.from_numpy(np.ascontiguousarray(img)).float()  # :535


def views_from_frames(frames_u8, mean=LOADER_MEAN, img_norm=True):
    """(B, N, H, W, 3) uint8 RGB frames -> the (B, 3N, H, W) float32 batch forward() takes: the loader transform per
    frame, then the trainer's channel concat `torch.cat(tuple(images_list), dim=1)` (trainer.py:651,785)."""
    frames = np.asarray(frames_u8)
    b, n = frames.shape[:2]
    return torch.stack([torch.cat([loader_transform(frames[i, a], mean, img_norm) for a in range(n)], 0)
                        for i in range(b)], 0)


def labels_from_logits(logits):
    """`outputs.data.max(1)[1]`, trainer.py:804 (first maximal index on ties)."""
    return torch.as_tensor(logits).max(1)[1]


# --------------------------------------------------------------------------------------------- metrics
def confusion_matrix(label_true, label_pred, n_class):
    """runningScore._fast_hist, metrics.py:99-105."""
    lt = np.asarray(label_true).reshape(-1)
    lp = np.asarray(label_pred).reshape(-1)
    mask = (lt >= 0) & (lt < n_class)
    return np.bincount(n_class * lt[mask].astype(int) + lp[mask], minlength=n_class ** 2).reshape(n_class, n_class)


def mean_iou(hist):
    """runningScore.get_scores 'Mean IoU', metrics.py:176-183."""
    hist = hist.astype(np.float64)
    with np.errstate(divide="ignore", invalid="ignore"):
        iu = np.diag(hist) / (hist.sum(axis=1) + hist.sum(axis=0) - np.diag(hist))
    return float(np.nanmean(iu))


def miou_between(ref_logits, got_logits, n_class=11):
    """mIoU of argmax(got) against argmax(ref) used as labels (there is no ground truth without the dataset)."""
    lt = torch.as_tensor(ref_logits).max(1)[1].cpu().numpy()
    lp = torch.as_tensor(got_logits).max(1)[1].cpu().numpy()
    return mean_iou(confusion_matrix(lt, lp, n_class))
