"""TEST / BENCH INFRASTRUCTURE (not product): one training step of the reference, for bench.py's reference arms.

`reference_step()` drives the UNMODIFIED reference modules staged in baseline/_ref (oracle/make_ref.py) with the
training-loop lines of /root/reference/dpc/main.py restated verbatim (the driver itself cannot be imported:
tensorboardX / matplotlib are not installed, main.py:8-9):

    main.py:25      torch.backends.cudnn.benchmark = True
    main.py:58-66   DPC_RNN(...); nn.DataParallel(model); model.to(cuda)
    main.py:67      criterion = nn.CrossEntropyLoss()
    main.py:80-81   optim.Adam(model.parameters(), lr, weight_decay)      (defaults lr 1e-3, wd 1e-5: main.py:35-36)
    main.py:178-185 process_output(mask)
    main.py:198     [score_, mask_] = model(input_seq)
    main.py:209-218 flatten, target = argmax(mask == 1), loss, calc_topk_accuracy (utils/utils.py:38-55)
    main.py:229-231 optimizer.zero_grad(); loss.backward(); optimizer.step()

Two documented deviations, both forced by the host: on a CPU-only host the two hard-coded `.cuda()` calls
(model_3d.py:88, convrnn.py:27) are made identity (SURVEY.md 3.4 trap 5), and `target_.view` is `.reshape`
(trap 4: on ONE device the reference's mask is a non-contiguous view and torch >= 2 refuses the `.view`).

When baseline/_ref is absent, `port_step()` gives the same step on the oracle's functional restatement.
"""
import contextlib
import io

import torch
import torch.nn as nn
import torch.optim as optim

from . import dpc_oracle as O
from .make_ref import import_reference


def calc_topk_accuracy(output, target, topk=(1,)):
    """utils/utils.py:38-55"""
    maxk = max(topk)
    batch_size = target.size(0)
    _, pred = output.topk(maxk, 1, True, True)
    pred = pred.t()
    correct = pred.eq(target.view(1, -1).expand_as(pred))
    return [correct[:k].contiguous().view(-1).float().sum(0).mul_(1 / batch_size) for k in topk]


def reference_step(network, img, pred_step, device, device_ids=None, seed=0):
    """-> step(input_seq) -> loss tensor, or None when the reference modules are not staged.
    device: torch.device; device_ids: GPUs for nn.DataParallel (None on CPU)."""
    mod = import_reference()
    if mod is None:
        return None
    if device.type == 'cpu':
        torch.Tensor.cuda = lambda self, *a, **k: self                # trap 5 (test infrastructure only)
    torch.manual_seed(seed)                                            # main.py:50
    with contextlib.redirect_stdout(io.StringIO()):
        model = mod.DPC_RNN(sample_size=img, num_seq=8, seq_len=5, network=network, pred_step=pred_step)
    if device.type == 'cuda':
        model = nn.DataParallel(model, device_ids=device_ids)
    model = model.to(device)
    criterion = nn.CrossEntropyLoss()
    optimizer = optim.Adam(model.parameters(), lr=1e-3, weight_decay=1e-5)
    model.train()
    state = {}

    def step(input_seq):
        B = input_seq.size(0)
        [score_, mask_] = model(input_seq)
        if 'target' not in state:
            (_, NP, SQ, B2, NS, _) = mask_.size()
            state['target'] = (mask_ == 1)
            state['dims'] = (B2, NS, NP, SQ)
        B2, NS, NP, SQ = state['dims']
        score_flattened = score_.view(B * NP * SQ, B2 * NS * SQ)
        target_flattened = state['target'].reshape(B * NP * SQ, B2 * NS * SQ).to(int).argmax(dim=1)
        loss = criterion(score_flattened, target_flattened)
        state['topk'] = calc_topk_accuracy(score_flattened, target_flattened, (1, 3, 5))
        optimizer.zero_grad()
        loss.backward()
        optimizer.step()
        return loss

    return step


def port_step(network, img, pred_step, device, seed=0):
    """the same step on the oracle port (functional restatement, oracle/dpc_oracle.py)"""
    sd = {k: v.to(device) for k, v in O.synthetic_state_dict(network, seed).items()}
    params = {k: v.clone().requires_grad_(True) for k, v in sd.items() if not k.startswith('agg.ConvGRUCell_00')}
    full = dict(params)
    for k in sd:
        if k.startswith('agg.ConvGRUCell_00'):
            full[k] = params[k.replace('agg.ConvGRUCell_00', 'agg.cell_list.0')]
    optimizer = optim.Adam(list(params.values()), lr=1e-3, weight_decay=1e-5)

    def step(input_seq):
        score, mask = O.dpc_forward(input_seq, full, network, pred_step)
        loss, s, target = O.nce_loss(score, mask)
        O.topk_accuracy(s, target)
        optimizer.zero_grad(set_to_none=True)
        loss.backward()
        optimizer.step()
        return loss

    return step
