"""ConvGRU with the reference's parameter surface (/root/reference/backbone/convrnn.py:4-88).

The nn.Conv2d gates are parameter holders with the reference's initialisation and RNG order
(default init, then orthogonal_/zero, convrnn.py:13-22); the cell is registered under both
`ConvGRUCell_00` and `cell_list.0` like the reference (convrnn.py:55-58), so state_dict keys match.
The DPC hot path (kernel_size=1, num_layers=1) runs fused inside dpc_b200.engine.head_forward.
"""
import torch
import torch.nn as nn


class ConvGRUCell(nn.Module):
    def __init__(self, input_size, hidden_size, kernel_size):
        super().__init__()
        self.input_size, self.hidden_size, self.kernel_size = input_size, hidden_size, kernel_size
        padding = kernel_size // 2
        self.reset_gate = nn.Conv2d(input_size + hidden_size, hidden_size, kernel_size, padding=padding)
        self.update_gate = nn.Conv2d(input_size + hidden_size, hidden_size, kernel_size, padding=padding)
        self.out_gate = nn.Conv2d(input_size + hidden_size, hidden_size, kernel_size, padding=padding)
        for gate in (self.reset_gate, self.update_gate, self.out_gate):
            nn.init.orthogonal_(gate.weight)
        for gate in (self.reset_gate, self.update_gate, self.out_gate):
            nn.init.constant_(gate.bias, 0.)

    def forward(self, input_tensor, hidden_state):
        """one GRU step (convrnn.py:24-34): input [B,C,H,W], hidden [B,C,H,W] or None -> new state [B,C,H,W].
        (Inside DPC_RNN / LC the whole recurrence runs fused; this is the reference's stand-alone cell surface.)"""
        if self.kernel_size != 1 or self.input_size != self.hidden_size or self.hidden_size % 64 != 0:
            raise NotImplementedError('ConvGRUCell CUDA path: kernel_size=1, input=hidden (a multiple of 64) only')
        if not input_tensor.is_cuda:
            raise RuntimeError('dpc_b200 has no CPU path: input must be a CUDA tensor')
        B, C, H, W = input_tensor.shape
        rows = input_tensor.permute(0, 2, 3, 1).reshape(B * H * W, C).contiguous().float()
        h0 = None if hidden_state is None else hidden_state.permute(0, 2, 3, 1).reshape(B * H * W, C).contiguous().float()
        params = [self.reset_gate.weight, self.reset_gate.bias, self.update_gate.weight, self.update_gate.bias,
                  self.out_gate.weight, self.out_gate.bias]
        need = torch.is_grad_enabled() and (input_tensor.requires_grad or (h0 is not None and h0.requires_grad)
                                            or any(q.requires_grad for q in params))
        H_all = _GruSeqFn.apply(rows, h0, B * H * W, 1, 0.0, 0, need, *[q.contiguous() for q in params])
        return H_all.view(B, H, W, C).permute(0, 3, 1, 2)


class ConvGRU(nn.Module):
    def __init__(self, input_size, hidden_size, kernel_size, num_layers, dropout=0.1):
        super().__init__()
        self.input_size, self.hidden_size = input_size, hidden_size
        self.kernel_size, self.num_layers = kernel_size, num_layers
        cells = []
        for i in range(num_layers):
            cell = ConvGRUCell(input_size if i == 0 else hidden_size, hidden_size, kernel_size)
            name = 'ConvGRUCell_' + str(i).zfill(2)
            setattr(self, name, cell)
            cells.append(getattr(self, name))
        self.cell_list = nn.ModuleList(cells)
        self.dropout_layer = nn.Dropout(p=dropout)

    @property
    def dropout_p(self):
        return float(self.dropout_layer.p)

    def forward(self, x, hidden_state=None):
        """x [B,T,C,H,W] -> (layer_output [B,T,C,H,W], last_state [B,1,C,H,W])  (convrnn.py:62-88).
        Built for the configuration the reference uses: kernel_size 1, one layer, input == hidden (256 / 1024)."""
        from . import engine
        if self.kernel_size != 1 or self.num_layers != 1 or self.input_size != self.hidden_size \
                or self.hidden_size % 64 != 0:
            raise NotImplementedError('ConvGRU CUDA path: kernel_size=1, num_layers=1, input=hidden (a multiple of 64) only')
        if not x.is_cuda:
            raise RuntimeError('dpc_b200 has no CPU path: input must be a CUDA tensor')
        B, T, C, H, W = x.shape
        h0 = None
        if hidden_state is not None and hidden_state[0] is not None:
            h0 = hidden_state[0].permute(0, 2, 3, 1).reshape(B * H * W, C).contiguous()
        rows = x.permute(1, 0, 3, 4, 2).reshape(T * B * H * W, C).contiguous().float()
        cell = self.cell_list[0]
        params = [cell.reset_gate.weight, cell.reset_gate.bias, cell.update_gate.weight, cell.update_gate.bias,
                  cell.out_gate.weight, cell.out_gate.bias]
        p = float(self.dropout_layer.p) if self.training else 0.0
        seed = (torch.initial_seed() * 0x9E3779B1 + next(_seq_calls) * 1000003) & 0x7FFFFFFFFFFFFFFF
        need = torch.is_grad_enabled() and (x.requires_grad or any(q.requires_grad for q in params))
        H_all = _GruSeqFn.apply(rows, h0, B * H * W, T, p, seed, need, *[q.contiguous() for q in params])
        out = H_all.view(T, B, H, W, C).permute(1, 0, 4, 2, 3)
        return out, out[:, -1:].contiguous()


import itertools
from torch.autograd.function import once_differentiable

_seq_calls = itertools.count()


class _GruSeqFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, rows, h0, R, T, p, seed, need, *params):
        from . import engine
        P = dict(zip(engine.HEAD_PARAM_NAMES[:6], params))
        H_all, steps = engine.gru_sequence_forward(rows, h0, P, R, T, p, seed)
        ctx.steps, ctx.rows, ctx.R, ctx.T, ctx.has_h0 = (steps if need else None), rows, R, T, h0 is not None
        ctx.save_for_backward(*params)
        return H_all

    @staticmethod
    @once_differentiable
    def backward(ctx, dH_all):
        from . import engine
        P = dict(zip(engine.HEAD_PARAM_NAMES[:6], ctx.saved_tensors))
        G = engine._new_head_grads(P, dH_all.device)
        dX, dh0 = engine.gru_sequence_backward(ctx.steps, ctx.rows, dH_all.contiguous(), None, P, ctx.R, ctx.T, G)
        ctx.steps = None
        return (dX, dh0 if ctx.has_h0 else None, None, None, None, None, None) + tuple(G[n] for n in engine.HEAD_PARAM_NAMES[:6])
