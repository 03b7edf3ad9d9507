"""3-stage bottleneck ResNet patch embedder (models/ofa/resnet.py:86-225, frozen_bn.py:7-80), same parameter names.

BatchNorm (+ReLU, +residual add, running-stat update, frozen / eval mode) runs on this library's fused NHWC kernels
(ops.batch_norm); every 1x1 convolution is a tcgen05 GEMM on the NHWC bytes (ops.conv1x1) and every stride-1 3x3
convolution the implicit-GEMM kernel of csrc/conv.cu (ops.conv3x3), forward / dgrad / wgrad, in bf16.  The two stride-2 3x3
convolutions and every k > 1 convolution of the fp32 parity mode are a patch matrix (csrc/im2col.cu) times the weight on the
same GEMM (fp32 operands through the three-way bf16 split), the max-pool is csrc/pool.cu / csrc/im2col.cu: no cuDNN / ATen
convolution or pooling kernel is left on the path in either mode.  Output is NHWC-flattened [B, h*w, 1024] so `image_proj` consumes it without a transpose copy."""
import torch
import torch.nn as nn
import torch.nn.functional as F

from . import ops

BLOCKS = {"resnet50": [3, 4, 6], "resnet101": [3, 4, 23], "resnet152": [3, 8, 36]}


class FrozenBatchNorm2d(nn.Module):
    def __init__(self, num_features, eps=1e-5):
        super().__init__()
        self.eps = eps
        self.register_buffer("weight", torch.ones(num_features))
        self.register_buffer("bias", torch.zeros(num_features))
        self.register_buffer("running_mean", torch.zeros(num_features))
        self.register_buffer("running_var", torch.ones(num_features) - eps)

    momentum = 0.0


_TRACK = []
_GROUPS = [1]       # batch groups of the stem pass in flight (ResNetStem.forward(x, groups))


def _bn(mod, x, relu=False, residual=None):
    """nn.BatchNorm2d (batch statistics when the module is in training mode) or FrozenBatchNorm2d, as holders."""
    frozen = isinstance(mod, FrozenBatchNorm2d)
    training = mod.training and not frozen
    if training and mod.num_batches_tracked is not None:
        _TRACK.append(mod.num_batches_tracked)        # bumped once per stem forward with one fused add (ResNetStem.forward)
    return ops.batch_norm(x, mod.weight, mod.bias, mod.running_mean, mod.running_var, residual, relu, training,
                          0.1 if not frozen else 0.0, mod.eps, _GROUPS[0])


def _conv(mod, x):
    if mod.kernel_size == (1, 1) and mod.stride[0] == mod.stride[1]:
        return ops.conv1x1(x, mod.weight, mod.stride[0])          # tcgen05 GEMM on the NHWC bytes
    if (mod.kernel_size == (3, 3) and mod.stride == (1, 1) and mod.padding == (1, 1) and x.dtype == torch.bfloat16
            and mod.in_channels % 64 == 0 and mod.out_channels % 64 == 0):
        w = mod.weight
        if not w.is_contiguous(memory_format=torch.channels_last):     # once: [Cout][3][3][Cin] bytes, same shape / state dict
            w.data = w.data.contiguous(memory_format=torch.channels_last)
            if w.grad is not None:
                w.grad = None
        return ops.conv3x3(x, w)                                   # implicit-GEMM tcgen05 kernel (csrc/conv.cu)
    if (mod.kernel_size == (7, 7) and mod.stride == (2, 2) and mod.padding == (3, 3) and mod.in_channels == 3
            and x.dtype == torch.bfloat16):
        return ops.stem_conv7x7(x, mod.weight)                     # patch matrix + tcgen05 GEMM (csrc/pool.cu)
    # the two stride-2 3x3 convolutions and every k > 1 convolution of the fp32 parity mode: patch matrix + tcgen05 GEMM
    # (fp32 operands through the three-way bf16 split) -- no library convolution is left on the path
    assert mod.stride[0] == mod.stride[1] and mod.padding[0] == mod.padding[1] and mod.groups == 1 and mod.dilation == (1, 1)
    return ops.conv_im2col(x, mod.weight, mod.stride[0], mod.padding[0])


class Bottleneck(nn.Module):
    expansion = 4

    def __init__(self, inplanes, planes, stride, downsample, norm, drop_path_rate=0.0):
        super().__init__()
        self.drop_path_rate = drop_path_rate
        self.conv1 = nn.Conv2d(inplanes, planes, 1, bias=False)
        self.bn1 = norm(planes)
        self.conv2 = nn.Conv2d(planes, planes, 3, stride=stride, padding=1, bias=False)
        self.bn2 = norm(planes)
        self.conv3 = nn.Conv2d(planes, planes * 4, 1, bias=False)
        self.bn3 = norm(planes * 4)
        self.downsample = downsample

    def forward(self, x):
        # conv1 hands x back for the identity / downsample branch: that branch's gradient is then added in the epilogue of
        # conv1's dgrad GEMM instead of by a separate pass over the activation
        out, idn = ops.conv1x1(x, self.conv1.weight, fork=True)
        out = _bn(self.bn1, out, relu=True)
        out = _bn(self.bn2, _conv(self.conv2, out), relu=True)
        if self.downsample is not None:
            idn = _bn(self.downsample[1], _conv(self.downsample[0], idn))
        if self.drop_path_rate > 0.0 and self.training:        # resnet.py:5-20,130 (per-sample Bernoulli keep)
            out = _bn(self.bn3, _conv(self.conv3, out))
            keep = 1.0 - self.drop_path_rate
            m = torch.floor(keep + torch.rand(x.shape[0], 1, 1, 1, dtype=out.dtype, device=out.device))
            return F.relu(idn + out.div(keep) * m)
        return _bn(self.bn3, _conv(self.conv3, out), relu=True, residual=idn)   # relu(identity + bn3(conv3)) in one pass


class ResNetStem(nn.Module):
    def __init__(self, resnet_type, frozen_bn=False, drop_path_rate=0.0):
        super().__init__()
        norm = FrozenBatchNorm2d if frozen_bn else nn.BatchNorm2d
        self.inplanes = 64
        self.conv1 = nn.Conv2d(3, 64, 7, stride=2, padding=3, bias=False)
        self.bn1 = norm(64)
        layers = BLOCKS[resnet_type]
        self.drop_path_rate = drop_path_rate
        self.layer1 = self._make_layer(64, layers[0], 1, norm)
        self.layer2 = self._make_layer(128, layers[1], 2, norm)
        self.layer3 = self._make_layer(256, layers[2], 2, norm)
        for m in self.modules():
            if isinstance(m, nn.Conv2d):
                nn.init.kaiming_normal_(m.weight, mode="fan_out", nonlinearity="relu")
        self.last_hw = None

    def _make_layer(self, planes, blocks, stride, norm):
        down = None
        if stride != 1 or self.inplanes != planes * 4:
            down = nn.Sequential(nn.Conv2d(self.inplanes, planes * 4, 1, stride=stride, bias=False), norm(planes * 4))
        seq = [Bottleneck(self.inplanes, planes, stride, down, norm)]
        self.inplanes = planes * 4
        dpr = [x.item() for x in torch.linspace(0, self.drop_path_rate, blocks)]      # resnet.py:203-207
        for i in range(1, blocks):
            seq.append(Bottleneck(self.inplanes, planes, 1, None, norm, dpr[i]))
        return nn.Sequential(*seq)

    def forward(self, x, groups=1):
        """groups > 1: x holds the images of `groups` tasks back to back (equal counts); every BatchNorm normalises each
        group with its own batch statistics, so the result equals `groups` separate forwards while every convolution,
        GEMM and normalisation kernel runs once on the whole batch."""
        if groups > 1 and self.drop_path_rate > 0.0 and self.training:
            raise ValueError("grouped stem passes are not combined with ResNet drop-path (per-call random streams)")
        dt = self.conv1.weight.dtype
        _GROUPS[0] = groups
        del _TRACK[:]
        try:
            x = x.to(dt).contiguous(memory_format=torch.channels_last)
            x = _bn(self.bn1, _conv(self.conv1, x), relu=True)
            x = ops.max_pool3x3s2(x)
            x = self.layer3(self.layer2(self.layer1(x)))
        finally:
            if _TRACK:          # nn.BatchNorm2d bookkeeping: 94 counters, one multi-tensor kernel instead of 94 launches
                torch._foreach_add_(list(_TRACK), groups)
                del _TRACK[:]
            _GROUPS[0] = 1
        B, Cc, h, w = x.shape
        self.last_hw = (h, w)
        return x.permute(0, 2, 3, 1).reshape(B, h * w, Cc)
