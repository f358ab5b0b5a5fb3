"""CPU oracle for the DPC-RNN training path  --  TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this file.  The product path (dpc_b200/) never imports it and has no CPU fallback.

What it is
----------
A functional, state-dict driven restatement (fp32, or fp64 on request) of the reference hot path

    DPC_RNN.forward                      /root/reference/dpc/model_3d.py:46-98
    ResNet2d3d_full.forward              /root/reference/backbone/resnet_2d3d.py:259-270
    BasicBlock2d / BasicBlock3d.forward  /root/reference/backbone/resnet_2d3d.py:100-116, 64-80
    Bottleneck2d / Bottleneck3d.forward  /root/reference/backbone/resnet_2d3d.py:181-200, 139-158  (oracle only so far)
    ConvGRUCell / ConvGRU.forward        /root/reference/backbone/convrnn.py:24-34, 62-88
    loss / target (driver side)          /root/reference/dpc/main.py:178-185, 213-217
    calc_topk_accuracy                   /root/reference/utils/utils.py:38-55
    Adam(lr, weight_decay) step          /root/reference/dpc/main.py:81,229-231

The arithmetic of the reference lives in a third-party dependency (PyTorch ATen: conv3d, batch_norm,
max_pool3d, matmul, log_softmax; reference pins only "pytorch >= 0.4", README.md:30; installed here:
torch 2.11.0+cu128).  The oracle therefore restates the *composition* with torch.nn.functional CPU
ops, without nn.Module plumbing, `.cuda()` calls or Python-loop mask building.

Pinning
-------
The reference holds no golden vectors.  The oracle is pinned against the reference
itself, imported in the build container by oracle/make_golden.py; its outputs are committed under
tests/golden/ and checked by tests/test_oracle_golden.py (CPU, no reference needed at test time).
"""
import math
from collections import OrderedDict

import torch
import torch.nn.functional as F

# --------------------------------------------------------------------------------------------
# Architecture description (restated from backbone/resnet_2d3d.py:205-284, select_backbone.py:3-21)
# --------------------------------------------------------------------------------------------
NETWORKS = {
    # name: (blocks per stage)   -- block types are [2d, 2d, 3d, 3d] for all of them (resnet_2d3d.py:274-308)
    'resnet18': (2, 2, 2, 2),
    'resnet34': (3, 4, 6, 3),
    # Bottleneck networks (resnet_2d3d.py:119-202,286-308): restated and pinned here ahead of the product path,
    # which still raises NotImplementedError for them
    'resnet50': (3, 4, 6, 3),
    'resnet101': (3, 4, 23, 3),
    'resnet152': (3, 8, 36, 3),
    'resnet200': (3, 24, 36, 3),
}
BOTTLENECK = ('resnet50', 'resnet101', 'resnet152', 'resnet200')
STAGE_PLANES = (64, 128, 256, 256)      # layer4 narrowed to 256 (resnet_2d3d.py:222)
STAGE_IS3D = (False, False, True, True)
FEATURE_SIZE = 256                       # select_backbone.py:7,10 (BasicBlock networks)


def feature_size(network):
    """select_backbone.py:4-10: 1024 for the Bottleneck networks (256 planes x expansion 4), 256 for r18 / r34"""
    return 1024 if network in BOTTLENECK else FEATURE_SIZE


def backbone_spec(network):
    """List of blocks: dict(name, block, inplanes, planes, outplanes, stride, is3d, downsample, final_relu)."""
    if network not in NETWORKS:
        raise IOError('model type is wrong')            # select_backbone.py:19
    expansion = 4 if network in BOTTLENECK else 1       # resnet_2d3d.py:48,84,120,162
    spec = []
    inplanes = 64
    for si, nblocks in enumerate(NETWORKS[network]):
        planes = STAGE_PLANES[si]
        stride = 1 if si == 0 else 2
        for bi in range(nblocks):
            s = stride if bi == 0 else 1
            ds = (bi == 0) and (s != 1 or inplanes != planes * expansion)      # resnet_2d3d.py:234
            is_last = (si == 3 and bi == nblocks - 1)
            spec.append(dict(name='layer%d.%d' % (si + 1, bi), block='bottleneck' if expansion == 4 else 'basic',
                             inplanes=inplanes, planes=planes, outplanes=planes * expansion,
                             stride=s, is3d=STAGE_IS3D[si], downsample=ds,
                             final_relu=not is_last))              # resnet_2d3d.py:249-252
            inplanes = planes * expansion
    return spec


def param_shapes(network):
    """OrderedDict key -> shape, in the reference's state_dict order (SURVEY.md §3.4 trap 6)."""
    sh = OrderedDict()
    sh['backbone.conv1.weight'] = (64, 3, 1, 7, 7)
    sh['backbone.bn1.weight'] = (64,)
    sh['backbone.bn1.bias'] = (64,)
    for b in backbone_spec(network):
        p = 'backbone.' + b['name']
        k = (3, 3, 3) if b['is3d'] else (1, 3, 3)
        if b['block'] == 'bottleneck':                              # resnet_2d3d.py:123-137,165-179
            sh[p + '.conv1.weight'] = (b['planes'], b['inplanes'], 1, 1, 1)
            sh[p + '.bn1.weight'] = (b['planes'],)
            sh[p + '.bn1.bias'] = (b['planes'],)
            sh[p + '.conv2.weight'] = (b['planes'], b['planes']) + k
            sh[p + '.bn2.weight'] = (b['planes'],)
            sh[p + '.bn2.bias'] = (b['planes'],)
            sh[p + '.conv3.weight'] = (b['outplanes'], b['planes'], 1, 1, 1)
            sh[p + '.bn3.weight'] = (b['outplanes'],)
            sh[p + '.bn3.bias'] = (b['outplanes'],)
        else:
            sh[p + '.conv1.weight'] = (b['planes'], b['inplanes']) + k
            sh[p + '.bn1.weight'] = (b['planes'],)
            sh[p + '.bn1.bias'] = (b['planes'],)
            sh[p + '.conv2.weight'] = (b['planes'], b['planes']) + k
            sh[p + '.bn2.weight'] = (b['planes'],)
            sh[p + '.bn2.bias'] = (b['planes'],)
        if b['downsample']:
            sh[p + '.downsample.0.weight'] = (b['outplanes'], b['inplanes'], 1, 1, 1)
            sh[p + '.downsample.1.weight'] = (b['outplanes'],)
            sh[p + '.downsample.1.bias'] = (b['outplanes'],)
    D = feature_size(network)
    for cell in ('agg.ConvGRUCell_00', 'agg.cell_list.0'):          # registered twice, convrnn.py:55-58
        for g in ('reset_gate', 'update_gate', 'out_gate'):
            sh['%s.%s.weight' % (cell, g)] = (D, 2 * D, 1, 1)
            sh['%s.%s.bias' % (cell, g)] = (D,)
    for i in (0, 2):                                                # model_3d.py:36-40
        sh['network_pred.%d.weight' % i] = (D, D, 1, 1)
        sh['network_pred.%d.bias' % i] = (D,)
    return sh


def synthetic_state_dict(network, seed, dtype=torch.float32):
    """Portable, non-degenerate parameters: every tensor drawn from a seeded CPU generator
    (no LAPACK, so identical on every host with the same torch).  BN affine parameters are
    randomised too, which tests more than the reference's 1/0 initialisation does."""
    g = torch.Generator().manual_seed(seed)
    sd = OrderedDict()
    for k, shape in param_shapes(network).items():
        if k.startswith('agg.cell_list.0'):
            sd[k] = sd[k.replace('agg.cell_list.0', 'agg.ConvGRUCell_00')]   # same storage in the reference
            continue
        leaf = k.rsplit('.', 1)[1]
        is_bn = len(shape) == 1 and ('bn' in k or 'downsample.1' in k)
        if is_bn and leaf == 'weight':
            t = 0.5 + torch.rand(shape, generator=g)
        elif leaf == 'bias':
            t = 0.1 * torch.randn(shape, generator=g)
        else:
            if k.startswith('backbone.'):
                fan = shape[0] * shape[2] * shape[3] * shape[4]     # fan_out, like kaiming_normal_(mode='fan_out')
                std = math.sqrt(2.0 / fan)
            else:
                std = 1.0 / math.sqrt(shape[1])
            t = std * torch.randn(shape, generator=g)
        sd[k] = t.to(dtype)
    return sd


def reference_init_state_dict(network, seed=0):
    """Restatement of the reference's own initialisation:
    nn.Conv3d default init then kaiming_normal_(fan_out) (resnet_2d3d.py:224-230), BN 1/0,
    nn.Conv2d default init then orthogonal_/zero (convrnn.py:17-22, model_3d.py:100-105),
    consuming the global CPU RNG in the reference's order after torch.manual_seed(seed)
    (main.py:50).  The orthogonal part goes through LAPACK QR and is only bit-reproducible on the
    same host; use synthetic_state_dict() for portable fixtures."""
    import torch.nn.init as init
    torch.manual_seed(seed)
    shapes = param_shapes(network)
    sd = OrderedDict()

    def default_conv(wshape, bias):
        w = torch.empty(wshape)
        init.kaiming_uniform_(w, a=math.sqrt(5))
        b = None
        if bias:
            fan_in = wshape[1] * int(torch.tensor(wshape[2:]).prod())
            b = torch.empty(wshape[0])
            init.uniform_(b, -1 / math.sqrt(fan_in), 1 / math.sqrt(fan_in))
        return w, b

    bb_convs = [k for k in shapes if k.startswith('backbone.') and len(shapes[k]) == 5]
    # 1) construction order == state_dict order for the backbone convs (default init draws)
    for k in bb_convs:
        sd[k], _ = default_conv(shapes[k], False)
    # 2) `for m in self.modules()` re-initialises them in module order (== same order)
    for k in bb_convs:
        init.kaiming_normal_(sd[k], mode='fan_out')
    for k in shapes:
        if k.startswith('backbone.') and len(shapes[k]) == 1:
            sd[k] = torch.ones(shapes[k]) if k.endswith('weight') else torch.zeros(shapes[k])
    # 3) ConvGRUCell: three default Conv2d inits (reset, update, out), then orthogonal_ x3, zeros
    cell = 'agg.ConvGRUCell_00'
    for g in ('reset_gate', 'update_gate', 'out_gate'):
        w, b = default_conv(shapes['%s.%s.weight' % (cell, g)], True)
        sd['%s.%s.weight' % (cell, g)], sd['%s.%s.bias' % (cell, g)] = w, b
    for g in ('reset_gate', 'update_gate', 'out_gate'):
        init.orthogonal_(sd['%s.%s.weight' % (cell, g)])
    for g in ('reset_gate', 'update_gate', 'out_gate'):
        sd['%s.%s.bias' % (cell, g)].zero_()
    # 4) network_pred default inits
    for i in (0, 2):
        w, b = default_conv(shapes['network_pred.%d.weight' % i], True)
        sd['network_pred.%d.weight' % i], sd['network_pred.%d.bias' % i] = w, b
    # 5) _initialize_weights(agg): named_parameters() order (weight, bias per gate); duplicates skipped
    for g in ('reset_gate', 'update_gate', 'out_gate'):
        init.orthogonal_(sd['%s.%s.weight' % (cell, g)], 1)
        sd['%s.%s.bias' % (cell, g)].zero_()
    # 6) _initialize_weights(network_pred)
    for i in (0, 2):
        init.orthogonal_(sd['network_pred.%d.weight' % i], 1)
        sd['network_pred.%d.bias' % i].zero_()
    out = OrderedDict()
    for k in shapes:
        out[k] = sd[k.replace('agg.cell_list.0', 'agg.ConvGRUCell_00')]
    return out


# --------------------------------------------------------------------------------------------
# Forward restatement
# --------------------------------------------------------------------------------------------
def _bn(x, sd, prefix, eps=1e-5):
    # nn.BatchNorm3d(track_running_stats=False): batch statistics in train AND eval (model_3d.py:28)
    return F.batch_norm(x, None, None, sd[prefix + '.weight'], sd[prefix + '.bias'], True, 0.0, eps)


def backbone_forward(x, sd, network, prefix='backbone.', taps=None):
    """x: [NB,3,T,H,W] -> [NB,feature_size,T',H/32,W/32].  `taps` (dict) collects intermediates."""
    def tap(name, t):
        if taps is not None:
            taps[name] = t
    w = sd[prefix + 'conv1.weight']
    x = F.conv3d(x, w, None, (1, 2, 2), (0, 3, 3))                  # resnet_2d3d.py:211,260
    tap('stem.conv', x)
    x = F.relu(_bn(x, sd, prefix + 'bn1'))                          # :261-262
    x = F.max_pool3d(x, (1, 3, 3), (1, 2, 2), (0, 1, 1))            # :214,263
    tap('stem.out', x)
    for b in backbone_spec(network):
        p = prefix + b['name']
        if b['is3d']:
            s1, pad = (b['stride'],) * 3, (1, 1, 1)                 # conv3x3x3, :13-21
            sds = (b['stride'],) * 3
        else:
            s1, pad = (1, b['stride'], b['stride']), (0, 1, 1)      # conv1x3x3, :23-31
            sds = (1, b['stride'], b['stride'])                     # customized_stride, :236-239
        if b['block'] == 'bottleneck':                              # Bottleneck2d / 3d.forward, :139-158,181-200
            out = F.conv3d(x, sd[p + '.conv1.weight'], None, 1, 0)
            out = F.relu(_bn(out, sd, p + '.bn1'))
            out = F.conv3d(out, sd[p + '.conv2.weight'], None, s1, pad)
            out = F.relu(_bn(out, sd, p + '.bn2'))
            out = F.conv3d(out, sd[p + '.conv3.weight'], None, 1, 0)
            out = _bn(out, sd, p + '.bn3')
        else:
            out = F.conv3d(x, sd[p + '.conv1.weight'], None, s1, pad)
            out = F.relu(_bn(out, sd, p + '.bn1'))
            out = F.conv3d(out, sd[p + '.conv2.weight'], None, 1, pad)
            out = _bn(out, sd, p + '.bn2')
        if b['downsample']:
            res = F.conv3d(x, sd[p + '.downsample.0.weight'], None, sds, 0)
            res = _bn(res, sd, p + '.downsample.1')
        else:
            res = x
        out = out + res
        if b['final_relu']:
            out = F.relu(out)
        x = out
        tap(b['name'], x)
    return x


def gru_cell(x, h, sd, prefix='agg.cell_list.0.'):
    """ConvGRUCell.forward with kernel_size=1 (convrnn.py:24-34). x,h: [B,D,L,L]."""
    comb = torch.cat([x, h], 1)
    upd = torch.sigmoid(F.conv2d(comb, sd[prefix + 'update_gate.weight'], sd[prefix + 'update_gate.bias']))
    rst = torch.sigmoid(F.conv2d(comb, sd[prefix + 'reset_gate.weight'], sd[prefix + 'reset_gate.bias']))
    out = torch.tanh(F.conv2d(torch.cat([x, h * rst], 1), sd[prefix + 'out_gate.weight'],
                              sd[prefix + 'out_gate.bias']))
    return h * (1 - upd) + out * upd


def closed_form_mask(B, P, L, device='cpu'):
    """int8 mask [B,P,SQ,B,P,SQ]; equals the reference's loop construction (model_3d.py:86-96):
    1 = positive, -1 = temporal negative, -3 = spatial negative, 0 = easy negative.  Contiguous."""
    SQ = L * L
    b = torch.arange(B, device=device)
    p = torch.arange(P, device=device)
    s = torch.arange(SQ, device=device)
    same_b = (b.view(B, 1, 1, 1, 1, 1) == b.view(1, 1, 1, B, 1, 1))
    same_s = (s.view(1, 1, SQ, 1, 1, 1) == s.view(1, 1, 1, 1, 1, SQ))
    same_p = (p.view(1, P, 1, 1, 1, 1) == p.view(1, 1, 1, 1, P, 1))
    m = torch.zeros((B, P, SQ, B, P, SQ), dtype=torch.int8, device=device)
    m[(same_b & ~same_s).expand_as(m)] = -3
    m[(same_b & same_s & ~same_p).expand_as(m)] = -1
    m[(same_b & same_s & same_p).expand_as(m)] = 1
    return m


def dpc_forward(block, sd, network='resnet18', pred_step=3, dropout_masks=None, taps=None):
    """DPC_RNN.forward (model_3d.py:46-98).  block: [B,N,3,SL,H,W].
    dropout_masks: None (eval mode: dropout off) or a list of N-pred_step + pred_step tensors
    [B,D,L,L] holding the *scaled* keep mask (0 or 1/(1-p)) applied to the GRU state each step
    (convrnn.py:78), in call order; the last prediction-loop GRU step is dead work.
    Returns (score [B,P,SQ,B,P,SQ], mask int8)."""
    def tap(name, t):
        if taps is not None:
            taps[name] = t
    B, N, C, SL, H, W = block.shape
    D = feature_size(network)
    last_duration = int(math.ceil(SL / 4))                          # model_3d.py:24
    L = int(math.ceil(H / 32))                                      # model_3d.py:25
    x = block.reshape(B * N, C, SL, H, W)
    feat = backbone_forward(x, sd, network, taps=taps)
    feat = F.avg_pool3d(feat, (last_duration, 1, 1), stride=(1, 1, 1))   # :53
    feat_inf_all = feat.reshape(B, N, D, L, L)                      # pre-ReLU targets, :55
    feat = F.relu(feat).reshape(B, N, D, L, L)                      # :56-57
    feat_inf = feat_inf_all[:, N - pred_step:]
    tap('feature', feat)
    tap('feature_inf', feat_inf)
    mi = 0
    h = torch.zeros(B, D, L, L, dtype=block.dtype, device=block.device)   # convrnn.py:25-27
    for t in range(N - pred_step):                                  # agg over the first N-P blocks, :62
        h = gru_cell(feat[:, t], h, sd)
        if dropout_masks is not None:
            h = h * dropout_masks[mi]
        mi += 1
    tap('hidden_agg', h)
    pred = []
    for i in range(pred_step):                                      # :66-71
        w0, b0 = sd['network_pred.0.weight'], sd['network_pred.0.bias']
        w2, b2 = sd['network_pred.2.weight'], sd['network_pred.2.bias']
        p_tmp = F.conv2d(F.relu(F.conv2d(h, w0, b0)), w2, b2)
        pred.append(p_tmp)
        if i < pred_step - 1:                                       # the last GRU step is dead (trap 3)
            h = gru_cell(F.relu(p_tmp), h, sd)
            if dropout_masks is not None:
                h = h * dropout_masks[mi]
            mi += 1
    pred = torch.stack(pred, 1)                                     # [B,P,D,L,L]
    tap('pred', pred)
    P = pred_step
    pm = pred.permute(0, 1, 3, 4, 2).reshape(B * P * L * L, D)      # :81
    fm = feat_inf.permute(0, 1, 3, 4, 2).reshape(B * P * L * L, D)  # :82
    score = torch.matmul(pm, fm.t()).view(B, P, L * L, B, P, L * L)  # :83
    mask = closed_form_mask(B, P, L, block.device)
    return score, mask


def nce_loss(score, mask):
    """Driver-side loss (main.py:178-185, 213-217): target = argmax(mask == 1) per row, CE mean."""
    B, P, SQ, B2, NS, _ = mask.shape
    s = score.reshape(B * P * SQ, B2 * NS * SQ)
    target = (mask == 1).reshape(B * P * SQ, B2 * NS * SQ).to(torch.int64).argmax(1)
    return F.cross_entropy(s, target), s, target


def topk_accuracy(score_flat, target, topk=(1, 3, 5)):
    """utils/utils.py:38-55."""
    maxk = max(topk)
    _, pred = score_flat.topk(maxk, 1, True, True)
    correct = pred.t().eq(target.view(1, -1))
    return [correct[:k].reshape(-1).float().sum() / target.numel() for k in topk]


def train_step_grads(block, sd, network='resnet18', pred_step=3, dropout_masks=None):
    """forward + CE + backward through autograd on the functional graph.
    Returns (loss, score, grads dict keyed like the state_dict; duplicated GRU keys share a grad)."""
    leaves = OrderedDict()
    for k, v in sd.items():
        if k.startswith('agg.ConvGRUCell_00'):
            continue
        leaves[k] = v.detach().clone().requires_grad_(True)
    full = OrderedDict(leaves)
    for k in sd:
        if k.startswith('agg.ConvGRUCell_00'):
            full[k] = leaves[k.replace('agg.ConvGRUCell_00', 'agg.cell_list.0')]
    score, mask = dpc_forward(block, full, network, pred_step, dropout_masks)
    loss, _, _ = nce_loss(score, mask)
    gl = torch.autograd.grad(loss, list(leaves.values()), allow_unused=True)
    grads = OrderedDict()
    for (k, _), g in zip(leaves.items(), gl):
        grads[k] = g
    for k in sd:
        if k.startswith('agg.ConvGRUCell_00'):
            grads[k] = grads[k.replace('agg.ConvGRUCell_00', 'agg.cell_list.0')]
    return loss.detach(), score.detach(), grads


def adam_step(params, grads, state, lr=1e-3, wd=1e-5, betas=(0.9, 0.999), eps=1e-8):
    """torch.optim.Adam with L2 weight decay (not AdamW), main.py:81.  In place on `params`."""
    state['step'] = state.get('step', 0) + 1
    t = state['step']
    b1, b2 = betas
    for k, p in params.items():
        g = grads[k]
        if g is None:
            continue
        g = g + wd * p
        m = state.setdefault('m.' + k, torch.zeros_like(p))
        v = state.setdefault('v.' + k, torch.zeros_like(p))
        m.mul_(b1).add_(g, alpha=1 - b1)
        v.mul_(b2).addcmul_(g, g, value=1 - b2)
        bc1, bc2 = 1 - b1 ** t, 1 - b2 ** t
        denom = (v.sqrt() / math.sqrt(bc2)).add_(eps)
        p.addcdiv_(m, denom, value=-lr / bc1)


# --------------------------------------------------------------------------------------------
# CPU baseline timing (BASELINE config 1): fwd + CE + bwd + Adam, all host threads
# --------------------------------------------------------------------------------------------
def cpu_train_step_time(network='resnet18', img=128, batch=4, steps=3, warmup=1, seed=0):
    import time
    sd = synthetic_state_dict(network, seed)
    uniq = OrderedDict((k, v) for k, v in sd.items() if not k.startswith('agg.ConvGRUCell_00'))
    g = torch.Generator().manual_seed(1234)
    block = torch.randn(batch, 8, 3, 5, img, img, generator=g)
    state = {}
    times = []
    for i in range(warmup + steps):
        t0 = time.perf_counter()
        _, _, grads = train_step_grads(block, sd, network)
        adam_step(uniq, grads, state)
        dt = time.perf_counter() - t0
        if i >= warmup:
            times.append(dt)
    times.sort()
    return times[len(times) // 2]


# --------------------------------------------------------------------------------------------
# LC classifier: restatement of /root/reference/eval/model_3d_lc.py:47-65
# --------------------------------------------------------------------------------------------
def lc_param_shapes(network, num_class=101):
    """parameters AND buffers of LC in state_dict order (track_running_stats=True everywhere)"""
    sh = OrderedDict()

    def bn(prefix, c):
        sh[prefix + '.weight'] = (c,); sh[prefix + '.bias'] = (c,)
        sh[prefix + '.running_mean'] = (c,); sh[prefix + '.running_var'] = (c,); sh[prefix + '.num_batches_tracked'] = ()
    sh['backbone.conv1.weight'] = (64, 3, 1, 7, 7)
    bn('backbone.bn1', 64)
    for b in backbone_spec(network):
        p = 'backbone.' + b['name']
        k = (3, 3, 3) if b['is3d'] else (1, 3, 3)
        if b['block'] == 'bottleneck':                              # resnet_2d3d.py:123-137,165-179
            sh[p + '.conv1.weight'] = (b['planes'], b['inplanes'], 1, 1, 1)
            bn(p + '.bn1', b['planes'])
            sh[p + '.conv2.weight'] = (b['planes'], b['planes']) + k
            bn(p + '.bn2', b['planes'])
            sh[p + '.conv3.weight'] = (b['outplanes'], b['planes'], 1, 1, 1)
            bn(p + '.bn3', b['outplanes'])
        else:
            sh[p + '.conv1.weight'] = (b['planes'], b['inplanes']) + k
            bn(p + '.bn1', b['planes'])
            sh[p + '.conv2.weight'] = (b['planes'], b['planes']) + k
            bn(p + '.bn2', b['planes'])
        if b['downsample']:
            sh[p + '.downsample.0.weight'] = (b['outplanes'], b['inplanes'], 1, 1, 1)
            bn(p + '.downsample.1', b['outplanes'])
    D = feature_size(network)
    for cell in ('agg.ConvGRUCell_00', 'agg.cell_list.0'):
        for g in ('reset_gate', 'update_gate', 'out_gate'):
            sh['%s.%s.weight' % (cell, g)] = (D, 2 * D, 1, 1)
            sh['%s.%s.bias' % (cell, g)] = (D,)
    bn('final_bn', D)
    sh['final_fc.1.weight'] = (num_class, D)
    sh['final_fc.1.bias'] = (num_class,)
    return sh


def lc_synthetic_state_dict(network, seed, num_class=101):
    g = torch.Generator().manual_seed(seed)
    sd = OrderedDict()
    for k, shape in lc_param_shapes(network, num_class).items():
        if k.startswith('agg.cell_list.0'):
            sd[k] = sd[k.replace('agg.cell_list.0', 'agg.ConvGRUCell_00')]
            continue
        leaf = k.rsplit('.', 1)[1]
        if leaf == 'num_batches_tracked':
            t = torch.tensor(3, dtype=torch.int64)
        elif leaf == 'running_mean':
            t = 0.2 * torch.randn(shape, generator=g)
        elif leaf == 'running_var':
            t = 0.5 + torch.rand(shape, generator=g)
        elif len(shape) == 1 and leaf == 'weight':
            t = 0.5 + torch.rand(shape, generator=g)
        elif leaf == 'bias':
            t = 0.1 * torch.randn(shape, generator=g)
        elif k.startswith('backbone.'):
            t = math.sqrt(2.0 / (shape[0] * shape[2] * shape[3] * shape[4])) * torch.randn(shape, generator=g)
        else:
            t = torch.randn(shape, generator=g) / math.sqrt(shape[1])
        sd[k] = t
    return sd


def _bn_rs(x, sd, prefix, training, momentum=0.1, eps=1e-5, new_stats=None):
    """nn.BatchNorm(track_running_stats=True): train -> batch statistics (+ records the updated buffers in
    new_stats), eval -> running statistics"""
    rm, rv = sd[prefix + '.running_mean'], sd[prefix + '.running_var']
    if training:
        rm2, rv2 = rm.clone(), rv.clone()
        y = F.batch_norm(x, rm2, rv2, sd[prefix + '.weight'], sd[prefix + '.bias'], True, momentum, eps)
        if new_stats is not None:
            new_stats[prefix + '.running_mean'], new_stats[prefix + '.running_var'] = rm2, rv2
        return y
    return F.batch_norm(x, rm, rv, sd[prefix + '.weight'], sd[prefix + '.bias'], False, momentum, eps)


def lc_forward(block, sd, network='resnet18', training=False, new_stats=None):
    """LC.forward with dropout off (eval, or train with p = 0).  Returns (output [B,1,num_class], context [B,1,D])."""
    B, N, C, SL, H, W = block.shape
    D = feature_size(network)
    last_duration = int(math.ceil(SL / 4))
    L = int(math.ceil(H / 32))
    bn = lambda x, p: _bn_rs(x, sd, p, training, new_stats=new_stats)
    x = block.reshape(B * N, C, SL, H, W)
    x = F.conv3d(x, sd['backbone.conv1.weight'], None, (1, 2, 2), (0, 3, 3))
    x = F.max_pool3d(F.relu(bn(x, 'backbone.bn1')), (1, 3, 3), (1, 2, 2), (0, 1, 1))
    for b in backbone_spec(network):
        p = 'backbone.' + b['name']
        if b['is3d']:
            s1, pad, sds = (b['stride'],) * 3, (1, 1, 1), (b['stride'],) * 3
        else:
            s1, pad, sds = (1, b['stride'], b['stride']), (0, 1, 1), (1, b['stride'], b['stride'])
        if b['block'] == 'bottleneck':                              # Bottleneck2d / 3d.forward, resnet_2d3d.py:139-158,181-200
            out = F.relu(bn(F.conv3d(x, sd[p + '.conv1.weight'], None, 1, 0), p + '.bn1'))
            out = F.relu(bn(F.conv3d(out, sd[p + '.conv2.weight'], None, s1, pad), p + '.bn2'))
            out = bn(F.conv3d(out, sd[p + '.conv3.weight'], None, 1, 0), p + '.bn3')
        else:
            out = F.relu(bn(F.conv3d(x, sd[p + '.conv1.weight'], None, s1, pad), p + '.bn1'))
            out = bn(F.conv3d(out, sd[p + '.conv2.weight'], None, 1, pad), p + '.bn2')
        res = bn(F.conv3d(x, sd[p + '.downsample.0.weight'], None, sds, 0), p + '.downsample.1') if b['downsample'] else x
        out = out + res
        x = F.relu(out) if b['final_relu'] else out
    feat = F.relu(x)                                                 # model_3d_lc.py:53
    feat = F.avg_pool3d(feat, (last_duration, 1, 1), stride=1).reshape(B, N, D, L, L)
    h = torch.zeros(B, D, L, L, dtype=block.dtype)
    for t in range(N):                                               # context, _ = self.agg(feature)
        h = gru_cell(feat[:, t], h, sd)
    context = h.mean((2, 3)).unsqueeze(1)                            # avg_pool3d over (1, L, L)   [B,1,D]
    context = _bn_rs(context.transpose(-1, -2), sd, 'final_bn', training, new_stats=new_stats).transpose(-1, -2)
    output = F.linear(context, sd['final_fc.1.weight'], sd['final_fc.1.bias']).view(B, -1, sd['final_fc.1.weight'].shape[0])
    return output, context
