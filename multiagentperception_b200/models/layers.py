"""Parameter containers for the conv blocks and backbones of the When2com models.

These nn.Modules exist to own parameters/buffers under exactly the state_dict names the reference uses
(SURVEY.md appendix D), so reference checkpoints load and optimizers / DataParallel see the usual module tree.
They are never *called* on the accelerated path: the engine (engine.py) reads their tensors, folds / packs them
once, and drives the sm_100a kernels. Calling one directly raises, so an accidental eager fallback cannot go
unnoticed.

Naming mirrors the reference modules each class stands in for:
  conv2DBatchNormRelu / deconv2DBatchNormRelu   ptsemseg/models/utils.py:87-120,148-168
  n_segnet_encoder / n_segnet_decoder           ptsemseg/models/backbone.py:12-55,99-140
  resnet_encoder / simple_decoder               ptsemseg/models/backbone.py:58-96,143-164
"""
import torch.nn as nn


class _Container(nn.Module):
    def forward(self, *a, **k):  # pragma: no cover - guard
        raise RuntimeError(
            "%s is a parameter container of the B200 engine and is not executed eagerly; call the model's "
            "forward() in eval mode on a CUDA device" % type(self).__name__)


class conv2DBatchNormRelu(_Container):
    """Conv2d(k, bias=True) + BatchNorm2d + ReLU; parameters live at cbr_unit.{0,1}.*"""

    def __init__(self, in_channels, n_filters, k_size=3, stride=1, padding=1):
        super().__init__()
        self.stride = int(stride)
        self.cbr_unit = nn.Sequential(
            nn.Conv2d(int(in_channels), int(n_filters), kernel_size=k_size, stride=stride, padding=padding, bias=True),
            nn.BatchNorm2d(int(n_filters)),
            nn.ReLU(inplace=True))

    @property
    def conv(self):
        return self.cbr_unit[0]

    @property
    def bn(self):
        return self.cbr_unit[1]


class deconv2DBatchNormRelu(_Container):
    """ConvTranspose2d(k3 s2 p1 op1, bias=True) + BatchNorm2d + ReLU; parameters at dcbr_unit.{0,1}.*"""

    def __init__(self, in_channels, n_filters, k_size=3, stride=2, padding=1, output_padding=1):
        super().__init__()
        if (k_size, stride, padding, output_padding) != (3, 2, 1, 1):
            raise ValueError("only the k3 s2 p1 op1 transposed conv of the reference decoders is supported")
        self.dcbr_unit = nn.Sequential(
            nn.ConvTranspose2d(int(in_channels), int(n_filters), kernel_size=3, stride=2, padding=1, output_padding=1,
                               bias=True),
            nn.BatchNorm2d(int(n_filters)),
            nn.ReLU(inplace=True))

    @property
    def conv(self):
        return self.dcbr_unit[0]

    @property
    def bn(self):
        return self.dcbr_unit[1]


class n_segnet_encoder(_Container):
    # (cout, stride) of conv1..conv13
    SPEC = ((64, 1), (64, 2), (128, 1), (128, 2), (256, 1), (256, 1), (256, 2), (512, 1), (512, 1), (512, 2),
            (512, 1), (512, 1), (512, 2))

    def __init__(self, n_classes=21, in_channels=3):
        super().__init__()
        self.in_channels = in_channels
        cin = in_channels
        for i, (cout, stride) in enumerate(self.SPEC, 1):
            setattr(self, "conv%d" % i, conv2DBatchNormRelu(cin, cout, 3, stride, 1))
            cin = cout

    def units(self):
        return [getattr(self, "conv%d" % i) for i in range(1, len(self.SPEC) + 1)]


class n_segnet_decoder(_Container):
    # (kind, cout) of deconv1..deconv12; 'd' = transposed conv, 'c' = conv; last cout = n_classes
    SPEC = (("d", 512), ("c", 512), ("c", 512), ("d", 512), ("c", 512), ("c", 256), ("d", 256), ("c", 128),
            ("d", 128), ("c", 64), ("d", 64), ("c", None))

    def __init__(self, n_classes=21, in_channels=512):
        super().__init__()
        self.in_channels = in_channels
        cin = in_channels
        for i, (kind, cout) in enumerate(self.SPEC, 1):
            cout = n_classes if cout is None else cout
            blk = deconv2DBatchNormRelu(cin, cout) if kind == "d" else conv2DBatchNormRelu(cin, cout, 3, 1, 1)
            setattr(self, "deconv%d" % i, blk)
            cin = cout

    def units(self):
        return [getattr(self, "deconv%d" % i) for i in range(1, len(self.SPEC) + 1)]


class _BasicBlock(_Container):
    def __init__(self, cin, cout, stride):
        super().__init__()
        self.conv1 = nn.Conv2d(cin, cout, 3, stride, 1, bias=False)
        self.bn1 = nn.BatchNorm2d(cout)
        self.relu = nn.ReLU(inplace=True)
        self.conv2 = nn.Conv2d(cout, cout, 3, 1, 1, bias=False)
        self.bn2 = nn.BatchNorm2d(cout)
        self.downsample = None
        if stride != 1 or cin != cout:
            self.downsample = nn.Sequential(nn.Conv2d(cin, cout, 1, stride, bias=False), nn.BatchNorm2d(cout))
        self.stride = stride


class _ResNet18Trunk(_Container):
    """State-dict layout of pretrainedmodels.resnet18 (torchvision resnet18 with fc renamed last_linear)."""

    def __init__(self):
        super().__init__()
        self.conv1 = nn.Conv2d(3, 64, 7, 2, 3, bias=False)
        self.bn1 = nn.BatchNorm2d(64)
        self.relu = nn.ReLU(inplace=True)
        self.maxpool = nn.MaxPool2d(3, 2, 1)
        cin = 64
        for li, cout in enumerate((64, 128, 256, 512), 1):
            stride = 1 if li == 1 else 2
            setattr(self, "layer%d" % li, nn.Sequential(_BasicBlock(cin, cout, stride), _BasicBlock(cout, cout, 1)))
            cin = cout
        self.avgpool = nn.AdaptiveAvgPool2d(1)
        self.last_linear = nn.Linear(512, 1000)  # unused by the forward path; kept for checkpoint key parity


class resnet_encoder(_Container):
    def __init__(self, n_classes=21, in_channels=3):
        super().__init__()
        if in_channels != 3:
            raise ValueError("resnet_encoder takes 3-channel images")
        self.feature_backbone = _ResNet18Trunk()
        fb = self.feature_backbone
        # the reference registers the same layers a second time under these names (backbone.py:65-69)
        self.backbone_0 = fb.conv1
        self.backbone_1 = nn.Sequential(fb.bn1, fb.relu, fb.maxpool, fb.layer1)
        self.backbone_2 = fb.layer2
        self.backbone_3 = fb.layer3
        self.backbone_4 = fb.layer4


class simple_decoder(_Container):
    def __init__(self, n_classes=21, in_channels=512):
        super().__init__()
        self.in_channels = in_channels
        self.pred = nn.Sequential(nn.Conv2d(in_channels, 256, 3, padding=1), nn.ReLU(inplace=True),
                                  nn.Conv2d(256, n_classes, 3, padding=1))


def get_encoder(name):
    try:
        return {"n_segnet_encoder": n_segnet_encoder, "resnet_encoder": resnet_encoder}[name]
    except KeyError:
        raise ValueError("Encoder {} not available".format(name))


def get_decoder(name):
    try:
        return {"n_segnet_decoder": n_segnet_decoder, "simple_decoder": simple_decoder}[name]
    except KeyError:
        raise ValueError("Decoder {} not available".format(name))
