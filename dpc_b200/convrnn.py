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
        raise RuntimeError('ConvGRUCell is a parameter holder; the CUDA path runs inside DPC_RNN.forward')


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
        raise NotImplementedError('stand-alone ConvGRU.forward (eval/LC) is a SURVEY.md §8(f) "next" row; '
                                  'DPC_RNN.forward runs the fused CUDA GRU')
