"""2d3d-ResNet backbone with the reference's module/parameter surface and a B200 CUDA forward.

Mirrors /root/reference/backbone/resnet_2d3d.py: same class names, constructor arguments,
parameter names/shapes (so `state_dict()` round-trips with the reference's checkpoints) and the
same initialisation *and RNG consumption order* (nn.Conv3d default init, then kaiming_normal_
fan_out; resnet_2d3d.py:224-230) -- the nn.Conv3d / nn.BatchNorm3d children are parameter holders
only; their torch forward is never called.  Compute goes through dpc_b200.engine -> libdpc_b200.so.
"""
import torch
import torch.nn as nn
from torch.autograd.function import once_differentiable

from . import engine
from .arch import backbone_spec, NETWORKS

__all__ = ['ResNet2d3d_full', 'BasicBlock2d', 'BasicBlock3d', 'Bottleneck2d', 'Bottleneck3d', 'resnet18_2d3d_full',
           'resnet34_2d3d_full',
           'resnet50_2d3d_full', 'resnet101_2d3d_full', 'resnet152_2d3d_full', 'resnet200_2d3d_full',
           'neq_load_customized']


def conv3x3x3(in_planes, out_planes, stride=1, bias=False):         # resnet_2d3d.py:13-21
    return nn.Conv3d(in_planes, out_planes, kernel_size=3, stride=stride, padding=1, bias=bias)


def conv1x3x3(in_planes, out_planes, stride=1, bias=False):         # resnet_2d3d.py:23-31
    return nn.Conv3d(in_planes, out_planes, kernel_size=(1, 3, 3), stride=(1, stride, stride),
                     padding=(0, 1, 1), bias=bias)


def get_tensor(module, dotted):
    """module.a.0.weight by attribute walk (works on nn.DataParallel replicas too)"""
    obj = module
    for part in dotted.split('.'):
        obj = getattr(obj, part)
    return obj


class _BasicBlock(nn.Module):
    expansion = 1
    _conv = None

    def __init__(self, inplanes, planes, stride=1, downsample=None, track_running_stats=True,
                 use_final_relu=True):
        super().__init__()
        self.use_final_relu = use_final_relu
        self.conv1 = type(self)._conv(inplanes, planes, stride, bias=False)
        self.bn1 = nn.BatchNorm3d(planes, track_running_stats=track_running_stats)
        self.relu = nn.ReLU(inplace=True)
        self.conv2 = type(self)._conv(planes, planes, bias=False)
        self.bn2 = nn.BatchNorm3d(planes, track_running_stats=track_running_stats)
        self.downsample = downsample
        self.stride = stride

    def forward(self, x):
        raise RuntimeError('blocks are parameter holders; call the ResNet2d3d_full module (CUDA path)')


class BasicBlock3d(_BasicBlock):                                   # resnet_2d3d.py:47-80
    _conv = staticmethod(conv3x3x3)


class BasicBlock2d(_BasicBlock):                                   # resnet_2d3d.py:83-116
    _conv = staticmethod(conv1x3x3)


class _Bottleneck(nn.Module):
    expansion = 4
    _conv = None

    def __init__(self, inplanes, planes, stride=1, downsample=None, track_running_stats=True,
                 use_final_relu=True):
        super().__init__()
        self.use_final_relu = use_final_relu
        self.conv1 = nn.Conv3d(inplanes, planes, kernel_size=1, bias=False)
        self.bn1 = nn.BatchNorm3d(planes, track_running_stats=track_running_stats)
        self.conv2 = type(self)._conv(planes, planes, stride, bias=False)
        self.bn2 = nn.BatchNorm3d(planes, track_running_stats=track_running_stats)
        self.conv3 = nn.Conv3d(planes, planes * 4, kernel_size=1, bias=False)
        self.bn3 = nn.BatchNorm3d(planes * 4, track_running_stats=track_running_stats)
        self.relu = nn.ReLU(inplace=True)
        self.downsample = downsample
        self.stride = stride

    def forward(self, x):
        raise RuntimeError('blocks are parameter holders; call the ResNet2d3d_full module (CUDA path)')


class Bottleneck3d(_Bottleneck):                                   # resnet_2d3d.py:119-158
    _conv = staticmethod(conv3x3x3)


class Bottleneck2d(_Bottleneck):                                   # resnet_2d3d.py:161-202
    _conv = staticmethod(conv1x3x3)


class _BackboneFn(torch.autograd.Function):
    """x [NB,3,T,H,W] -> channels-last feature rows [NB*To*Ho*Wo, feature_size].
    `tensors` = the parameters in `names` order, then (running_mean, running_var) per name in `bn_names` (empty for
    track_running_stats=False).  With buffers: training -> batch statistics + momentum update of the buffers in place;
    eval -> the buffers are the statistics (and the backward treats them as constants)."""

    @staticmethod
    def forward(ctx, x, network, names, need, training, bn_names, *tensors):
        # `need` is decided by the caller: grad mode is always off inside Function.forward
        n = len(names)
        P = dict(zip(names, tensors[:n]))
        bufs = tensors[n:]
        bn_state = {k: (bufs[2 * i], bufs[2 * i + 1]) for i, k in enumerate(bn_names)} if bn_names else None
        rows, dims, bctx = engine.backbone_forward(network, x, P, need_ctx=need, bn_state=bn_state, training=training)
        ctx.bctx, ctx.names, ctx.nbuf = bctx, names, len(bufs)
        ctx.save_for_backward(*tensors[:n])
        return rows

    @staticmethod
    @once_differentiable
    def backward(ctx, drows):
        if ctx.bctx is None:
            raise RuntimeError('backbone forward ran without saving activations')
        P = dict(zip(ctx.names, ctx.saved_tensors))
        G = engine.backbone_backward(ctx.bctx, drows.contiguous(), P)
        ctx.bctx = None
        return (None,) * 6 + tuple(G[n] for n in ctx.names) + (None,) * ctx.nbuf


class ResNet2d3d_full(nn.Module):                                  # resnet_2d3d.py:205-270
    def __init__(self, block, layers, track_running_stats=True):
        super().__init__()
        self.inplanes = 64
        self.track_running_stats = track_running_stats
        self.conv1 = nn.Conv3d(3, 64, kernel_size=(1, 7, 7), stride=(1, 2, 2), padding=(0, 3, 3), bias=False)
        self.bn1 = nn.BatchNorm3d(64, track_running_stats=track_running_stats)
        self.relu = nn.ReLU(inplace=True)
        self.maxpool = nn.MaxPool3d(kernel_size=(1, 3, 3), stride=(1, 2, 2), padding=(0, 1, 1))
        if not isinstance(block, list):
            block = [block] * 4
        if block == [BasicBlock2d, BasicBlock2d, BasicBlock3d, BasicBlock3d]:
            kind = 'basic'
        elif block == [Bottleneck2d, Bottleneck2d, Bottleneck3d, Bottleneck3d]:
            kind = 'bottleneck'
        else:
            raise NotImplementedError('only the [2d,2d,3d,3d] BasicBlock / Bottleneck layouts of the reference are built')
        self.layer1 = self._make_layer(block[0], 64, layers[0])
        self.layer2 = self._make_layer(block[1], 128, layers[1], stride=2)
        self.layer3 = self._make_layer(block[2], 256, layers[2], stride=2)
        self.layer4 = self._make_layer(block[3], 256, layers[3], stride=2, is_final=True)
        for m in self.modules():                                   # resnet_2d3d.py:224-230
            if isinstance(m, nn.Conv3d):
                m.weight = nn.init.kaiming_normal_(m.weight, mode='fan_out')
                if m.bias is not None:
                    m.bias.data.zero_()
            elif isinstance(m, nn.BatchNorm3d):
                m.weight.data.fill_(1)
                m.bias.data.zero_()
        layers = tuple(layers)
        self.network = {v: k for k, v in NETWORKS.items()}.get((layers, kind))
        if self.network is None:
            raise NotImplementedError('layer counts %s (%s blocks) are not one of the reference networks' % (layers, kind))
        self._names = engine.backbone_param_names(self.network)
        assert [b['downsample'] for b in backbone_spec(self.network)] == \
            [blk.downsample is not None for l in (self.layer1, self.layer2, self.layer3, self.layer4) for blk in l]

    def _make_layer(self, block, planes, blocks, stride=1, is_final=False):   # resnet_2d3d.py:232-257
        downsample = None
        if stride != 1 or self.inplanes != planes * block.expansion:
            customized_stride = (1, stride, stride) if block in (BasicBlock2d, Bottleneck2d) else stride
            downsample = nn.Sequential(
                nn.Conv3d(self.inplanes, planes * block.expansion, kernel_size=1, stride=customized_stride, bias=False),
                nn.BatchNorm3d(planes * block.expansion, track_running_stats=self.track_running_stats))
        layers = [block(self.inplanes, planes, stride, downsample, track_running_stats=self.track_running_stats)]
        self.inplanes = planes * block.expansion
        if is_final:
            for _ in range(1, blocks - 1):
                layers.append(block(self.inplanes, planes, track_running_stats=self.track_running_stats))
            layers.append(block(self.inplanes, planes, track_running_stats=self.track_running_stats,
                                use_final_relu=False))
        else:
            for _ in range(1, blocks):
                layers.append(block(self.inplanes, planes, track_running_stats=self.track_running_stats))
        return nn.Sequential(*layers)

    # ---- CUDA path ---------------------------------------------------------------------------
    def out_dims(self, T, H, W):
        """(To, Ho, Wo) of the feature map for an input clip of T frames of HxW"""
        e = engine._out_extent
        H, W = e(e(H, 7, 2, 3), 3, 2, 1), e(e(W, 7, 2, 3), 3, 2, 1)  # stem conv, max-pool
        H, W = e(H, 3, 2, 1), e(W, 3, 2, 1)                         # layer2: stride (1,2,2)
        for _ in range(2):                                          # layer3, layer4
            T, H, W = e(T, 3, 2, 1), e(H, 3, 2, 1), e(W, 3, 2, 1)
        return T, H, W

    def forward_rows(self, x):
        """x [NB,3,T,H,W] -> (rows [NB*To*Ho*Wo, feature_size] channels-last, (To,Ho,Wo))"""
        if x.dim() != 5 or x.shape[1] != 3:
            raise ValueError('expected [NB,3,T,H,W], got %s' % (tuple(x.shape),))
        if not x.is_cuda:
            raise RuntimeError('dpc_b200 has no CPU path: input must be a CUDA tensor')
        x = x.contiguous().float()
        # attribute walk, not named_parameters(): nn.DataParallel replicas hold their (non-leaf) parameter
        # copies as plain attributes, so named_parameters() is empty there
        params = [get_tensor(self, n).contiguous() for n in self._names]
        need = torch.is_grad_enabled() and any(p.requires_grad for p in params)
        bn_names, bufs = (), []
        if self.track_running_stats:                                # nn.BatchNorm3d default (resnet_2d3d.py:206): LC / eval
            bn_names = tuple(k for k, m in self.named_modules() if isinstance(m, nn.BatchNorm3d))
            for k in bn_names:
                m = get_tensor(self, k)
                bufs += [m.running_mean, m.running_var]
        rows = _BackboneFn.apply(x, self.network, self._names, need, self.training, bn_names, *params, *bufs)
        if self.track_running_stats and self.training:
            for k in bn_names:
                get_tensor(self, k).num_batches_tracked += 1
        return rows, self.out_dims(x.shape[2], x.shape[3], x.shape[4])

    def forward(self, x):                                           # resnet_2d3d.py:259-270
        rows, (To, Ho, Wo) = self.forward_rows(x)
        return rows.view(x.shape[0], To, Ho, Wo, -1).permute(0, 4, 1, 2, 3)     # NCDHW view


def resnet18_2d3d_full(**kwargs):                                  # resnet_2d3d.py:274-278
    return ResNet2d3d_full([BasicBlock2d, BasicBlock2d, BasicBlock3d, BasicBlock3d], [2, 2, 2, 2], **kwargs)


def resnet34_2d3d_full(**kwargs):                                  # resnet_2d3d.py:280-284
    return ResNet2d3d_full([BasicBlock2d, BasicBlock2d, BasicBlock3d, BasicBlock3d], [3, 4, 6, 3], **kwargs)


def _bottleneck_factory(layers):
    def f(**kwargs):
        return ResNet2d3d_full([Bottleneck2d, Bottleneck2d, Bottleneck3d, Bottleneck3d], list(layers), **kwargs)
    return f


resnet50_2d3d_full = _bottleneck_factory((3, 4, 6, 3))             # resnet_2d3d.py:286-290
resnet101_2d3d_full = _bottleneck_factory((3, 4, 23, 3))           # :292-296
resnet152_2d3d_full = _bottleneck_factory((3, 8, 36, 3))           # :298-302
resnet200_2d3d_full = _bottleneck_factory((3, 24, 36, 3))          # :304-308


def neq_load_customized(model, pretrained_dict):
    """Partial checkpoint load with the reference's semantics and report (resnet_2d3d.py:310-333):
    keys present in both are taken from the checkpoint, everything else keeps the model's value."""
    own = model.state_dict()
    unused = [k for k in pretrained_dict if k not in own]
    missing = [k for k in own if k not in pretrained_dict]
    print('\n=======Check Weights Loading======')
    print('Weights not used from pretrained file:')
    for k in unused:
        print(k)
    print('---------------------------')
    print('Weights not loaded into new model:')
    for k in missing:
        print(k)
    print('===================================\n')
    own.update({k: v for k, v in pretrained_dict.items() if k in own})
    model.load_state_dict(own)
    return model
