"""LC: the linear-classification / fine-tuning model of the reference's downstream task
(/root/reference/eval/model_3d_lc.py:12-75), on the same B200 kernels as DPC_RNN.

Same constructor, `forward(block) -> (output [B,1,num_class], context [B,1,D])`, parameter / buffer names
(backbone.* with BatchNorm running statistics, agg.*, final_bn.*, final_fc.1.*), so checkpoints from
`dpc/main.py` load with `neq_load_customized` exactly as `eval/test.py` does.
Differences from DPC_RNN that the kernels honour: track_running_stats=True (train: batch statistics + buffer
update; eval: running statistics), ReLU BEFORE the temporal average (model_3d_lc.py:53-55), the ConvGRU runs
over all N blocks, spatial mean, BatchNorm1d, Dropout + Linear.
"""
import itertools
import math

import torch
import torch.nn as nn
from torch.autograd.function import once_differentiable

from . import engine
from .select_backbone import select_resnet
from .convrnn import ConvGRU
from .resnet_2d3d import get_tensor

_calls = itertools.count()


class _LcHeadFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, rows, dims, B, N, training, gru_p, fc_p, seed, need, rm, rv, *params):
        P = dict(zip(engine.LC_PARAM_NAMES, params))
        out, context, hctx = engine.lc_head_forward(rows, dims, B, N, P, (rm, rv), training, gru_p, fc_p, seed, need_ctx=need)
        ctx.hctx = hctx
        ctx.save_for_backward(*params)
        return out, context

    @staticmethod
    @once_differentiable
    def backward(ctx, dout, dcontext):
        if ctx.hctx is None:
            raise RuntimeError('LC head forward ran without saving activations')
        P = dict(zip(engine.LC_PARAM_NAMES, ctx.saved_tensors))
        drows, G = engine.lc_head_backward(ctx.hctx, dout.contiguous(), dcontext, P)
        ctx.hctx = None
        return (drows,) + (None,) * 10 + tuple(G[n] for n in engine.LC_PARAM_NAMES)


class LC(nn.Module):
    def __init__(self, sample_size, num_seq, seq_len, network='resnet18', dropout=0.5, num_class=101):
        super().__init__()
        torch.cuda.manual_seed(666)                                  # model_3d_lc.py:16
        self.sample_size, self.num_seq, self.seq_len, self.num_class = sample_size, num_seq, seq_len, num_class
        print('=> Using RNN + FC model ')
        print('=> Use 2D-3D %s!' % network)
        self.last_duration = int(math.ceil(seq_len / 4))
        self.last_size = int(math.ceil(sample_size / 32))
        self.backbone, self.param = select_resnet(network, track_running_stats=True)
        self.param['num_layers'] = 1
        self.param['hidden_size'] = self.param['feature_size']
        print('=> using ConvRNN, kernel_size = 1')
        self.agg = ConvGRU(input_size=self.param['feature_size'], hidden_size=self.param['hidden_size'],
                           kernel_size=1, num_layers=self.param['num_layers'])
        self._initialize_weights(self.agg)
        self.final_bn = nn.BatchNorm1d(self.param['feature_size'])
        self.final_bn.weight.data.fill_(1)
        self.final_bn.bias.data.zero_()
        self.final_fc = nn.Sequential(nn.Dropout(dropout), nn.Linear(self.param['feature_size'], self.num_class))
        self._initialize_weights(self.final_fc)

    def forward(self, block):
        if block.dim() != 6:
            raise ValueError('expected block [B,N,C,SL,H,W], got %s' % (tuple(block.shape),))
        if not block.is_cuda:
            raise RuntimeError('dpc_b200 has no CPU path: input must be a CUDA tensor')
        (B, N, C, SL, H, W) = block.shape
        x = block.reshape(B * N, C, SL, H, W).contiguous().float()
        bb = self.backbone
        grad_on = torch.is_grad_enabled() and any(p.requires_grad for p in self.parameters())
        rows, dims = bb.forward_rows(x)                              # running statistics handled by the backbone module
        if self.training:
            self.final_bn.num_batches_tracked += 1
        if dims[0] != self.last_duration or dims[1] != self.last_size:
            raise ValueError('feature map %s does not match last_duration=%d / last_size=%d'
                             % (dims, self.last_duration, self.last_size))
        hp = [get_tensor(self, n).contiguous() for n in engine.LC_PARAM_NAMES]
        seed = (torch.initial_seed() * 0x9E3779B1 + next(_calls) * 1000003 + block.device.index * 7919) & 0x7FFFFFFFFFFFFFFF
        out, context = _LcHeadFn.apply(rows, dims, B, N, self.training, self.agg.dropout_p, float(self.final_fc[0].p),
                                       seed, grad_on, self.final_bn.running_mean, self.final_bn.running_var, *hp)
        return out.view(B, -1, self.num_class), context.view(B, 1, -1)

    def _initialize_weights(self, module):
        for name, param in module.named_parameters():
            if 'bias' in name:
                nn.init.constant_(param, 0.0)
            elif 'weight' in name:
                nn.init.orthogonal_(param, 1)
