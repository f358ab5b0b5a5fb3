"""DPC_RNN: drop-in for the reference module (/root/reference/dpc/model_3d.py:14-109).

Same constructor, `forward(block) -> [score, mask]`, `reset_mask()`, parameter names and
initialisation order, so `dpc/main.py` can `from dpc_b200.model_3d import DPC_RNN` and run
unchanged (wrapping in nn.DataParallel, Adam over .parameters(), state_dict checkpoints).
The forward/backward run on hand-written sm_100a kernels through libdpc_b200.so.
"""
import itertools
import math

import torch
import torch.nn as nn
from torch.autograd.function import once_differentiable

from . import engine
from .select_backbone import select_resnet
from .convrnn import ConvGRU

_dropout_calls = itertools.count()          # process-wide: DataParallel replicas are throw-away copies


class _HeadFn(torch.autograd.Function):
    """feature rows -> score [M, M] (pool/split, ConvGRU aggregate, predictor loop, score matmul)"""

    @staticmethod
    def forward(ctx, rows, dims, B, N, pred_step, dropout_p, seed, need, *params):
        # `need` is decided by the caller: grad mode is always off inside Function.forward
        P = dict(zip(engine.HEAD_PARAM_NAMES, params))
        score, hctx = engine.head_forward(rows, dims, B, N, pred_step, P, dropout_p, seed, need_ctx=need)
        ctx.hctx = hctx
        ctx.save_for_backward(*params)
        return score

    @staticmethod
    @once_differentiable
    def backward(ctx, dscore):
        if ctx.hctx is None:
            raise RuntimeError('head forward ran without saving activations')
        P = dict(zip(engine.HEAD_PARAM_NAMES, ctx.saved_tensors))
        drows, G = engine.head_backward(ctx.hctx, dscore.contiguous(), P)
        ctx.hctx = None
        return (drows, None, None, None, None, None, None, None) + tuple(G[n] for n in engine.HEAD_PARAM_NAMES)


class DPC_RNN(nn.Module):
    '''DPC with RNN'''

    def __init__(self, sample_size, num_seq=8, seq_len=5, pred_step=3, network='resnet50'):
        super().__init__()
        torch.cuda.manual_seed(233)                                  # model_3d.py:18 (lazy without CUDA)
        print('Using DPC-RNN model')
        self.sample_size = sample_size
        self.num_seq = num_seq
        self.seq_len = seq_len
        self.pred_step = pred_step
        self.last_duration = int(math.ceil(seq_len / 4))
        self.last_size = int(math.ceil(sample_size / 32))
        print('final feature map has size %dx%d' % (self.last_size, self.last_size))

        self.backbone, self.param = select_resnet(network, track_running_stats=False)
        self.param['num_layers'] = 1
        self.param['hidden_size'] = self.param['feature_size']
        self.agg = ConvGRU(input_size=self.param['feature_size'], hidden_size=self.param['hidden_size'],
                           kernel_size=1, num_layers=self.param['num_layers'])
        fs = self.param['feature_size']
        self.network_pred = nn.Sequential(nn.Conv2d(fs, fs, kernel_size=1, padding=0),
                                          nn.ReLU(inplace=True),
                                          nn.Conv2d(fs, fs, kernel_size=1, padding=0))
        self.mask = None
        self.relu = nn.ReLU(inplace=False)
        self._initialize_weights(self.agg)
        self._initialize_weights(self.network_pred)

    def _head_params(self):
        from .resnet_2d3d import get_tensor          # attribute walk: valid on nn.DataParallel replicas
        return [get_tensor(self, n).contiguous() for n in engine.HEAD_PARAM_NAMES]

    def forward(self, block):
        # block: [B, N, C, SL, H, W]
        if block.dim() != 6:
            raise ValueError('expected block [B,N,C,SL,H,W], got %s' % (tuple(block.shape),))
        (B, N, C, SL, H, W) = block.shape
        if N <= self.pred_step:
            raise ValueError('num_seq (%d) must exceed pred_step (%d)' % (N, self.pred_step))
        x = block.reshape(B * N, C, SL, H, W)
        rows, dims = self.backbone.forward_rows(x)
        To, Lh, Lw = dims
        if To != self.last_duration or Lh != self.last_size or Lw != self.last_size:
            # the reference would fail at its .view() (model_3d.py:55) for such shapes
            raise ValueError('feature map %s does not match last_duration=%d / last_size=%d'
                             % (dims, self.last_duration, self.last_size))
        p = self.agg.dropout_p if self.training else 0.0
        seed = 0
        if p > 0:
            seed = (torch.initial_seed() * 0x9E3779B1 + next(_dropout_calls) * 1000003
                    + block.device.index * 7919) & 0x7FFFFFFFFFFFFFFF
        hp = self._head_params()
        need = torch.is_grad_enabled() and (rows.requires_grad or any(q.requires_grad for q in hp))
        score = _HeadFn.apply(rows, dims, B, N, self.pred_step, p, seed, need, *hp)
        SQ = self.last_size ** 2
        score = score.view(B, self.pred_step, SQ, B, self.pred_step, SQ)
        if self.mask is None or self.mask.shape[0] != B or self.mask.device != block.device:
            # closed form of the loops at model_3d.py:86-96; contiguous int8 (SURVEY.md §3.4 trap 4)
            self.mask = engine.nce_mask(B, self.pred_step, SQ, block.device)
        return [score, self.mask]

    def _initialize_weights(self, module):
        for name, param in module.named_parameters():
            if 'bias' in name:
                nn.init.constant_(param, 0.0)
            elif 'weight' in name:
                nn.init.orthogonal_(param, 1)

    def reset_mask(self):
        self.mask = None
