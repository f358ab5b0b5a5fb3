"""DEV TOOL (test infrastructure, not product): end-to-end parameter-gradient error of the product path as the batch
grows, against the oracle on the same GPU in fp64 -- with the fp32 oracle's own distance from fp64 beside it.

    python tests/grad_table_tool.py [--out gpurun_out/r2_grad_table.json] [--max_b 128]

Rows: 2d3d-R18, 128^2, eval mode (dropout off, BatchNorm batch statistics), B in {2, 8, 32, 128}, for two parameter sets:
  'reference-init'  the reference's own initialisation (oracle.reference_init_state_dict: kaiming fan_out / orthogonal /
                    BN 1,0 -- what training and bench.py start from)
  'synthetic'       the portable randomised parameters of the golden fixtures (random BN affine, non-orthogonal head:
                    large logits, CE in its saturated regime)
Columns: rel-L2 over ALL parameters and worst single tensor for  ours vs fp64,  fp32-oracle vs fp64,  ours vs fp32-oracle;
forward score error (max-norm relative) of ours and of the fp32 oracle vs fp64; max |score| (the CE backward turns an
ABSOLUTE score error e into a RELATIVE softmax error e, so gradient error ~ max|score| x forward relative error)."""
import argparse
import contextlib
import io
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch

from oracle import dpc_oracle as O


def gnorm_err(a, b):
    """(rel-L2 over all tensors, worst tensor rel-L2, its name); b = reference"""
    num = den = 0.0
    worst, wk = 0.0, None
    for k in b:
        if k not in a:
            continue
        d = (a[k].double() - b[k].double())
        r = float(d.norm() / b[k].double().norm().clamp_min(1e-300))
        if r > worst:
            worst, wk = r, k
        num += float(d.pow(2).sum())
        den += float(b[k].double().pow(2).sum())
    return (num / den) ** 0.5, worst, wk


def run(network, img, B, sd, do64=True):
    import dpc_b200
    with contextlib.redirect_stdout(io.StringIO()):
        m = dpc_b200.DPC_RNN(sample_size=img, num_seq=8, seq_len=5, network=network, pred_step=3)
    m.load_state_dict(sd, strict=True)
    m = m.cuda().eval()
    g = torch.Generator().manual_seed(1000 + B)
    block = torch.randn(B, 8, 3, 5, img, img, generator=g).cuda()
    score, _ = m(block)
    loss = dpc_b200.NCECriterion()(score)
    loss.backward()
    torch.cuda.synchronize()
    ours = {k: p.grad.detach().cpu() for k, p in m.named_parameters()}
    s_ours = score.detach().cpu()
    l_ours = float(loss)
    del m, score, loss
    torch.cuda.empty_cache()
    sdc = {k: v.cuda() for k, v in sd.items()}
    l32, s32, g32 = O.train_step_grads(block, sdc, network, 3)
    g32 = {k: v.cpu() for k, v in g32.items()}
    s32 = s32.cpu()
    l32 = float(l32)
    torch.cuda.empty_cache()
    row = dict(B=B, loss_ours=l_ours, loss_o32=l32, max_abs_score=float(s32.abs().max()))
    a, w, wk = gnorm_err(ours, g32)
    row.update(ours_vs_o32_all=a, ours_vs_o32_worst=w, ours_vs_o32_worst_name=wk)
    row['score_ours_vs_o32'] = float((s_ours - s32).abs().max() / s32.abs().max())
    if do64:
        t0 = time.time()
        sd64 = {k: v.double() for k, v in sdc.items()}
        l64, s64, g64 = O.train_step_grads(block.double(), sd64, network, 3)
        torch.cuda.synchronize()
        row['fp64_seconds'] = time.time() - t0
        g64 = {k: v.cpu() for k, v in g64.items()}
        s64 = s64.cpu()
        del sd64
        torch.cuda.empty_cache()
        a, w, wk = gnorm_err(ours, g64)
        row.update(ours_vs_o64_all=a, ours_vs_o64_worst=w, ours_vs_o64_worst_name=wk)
        a, w, wk = gnorm_err(g32, g64)
        row.update(o32_vs_o64_all=a, o32_vs_o64_worst=w, o32_vs_o64_worst_name=wk)
        row['score_ours_vs_o64'] = float((s_ours.double() - s64).abs().max() / s64.abs().max())
        row['score_o32_vs_o64'] = float((s32.double() - s64).abs().max() / s64.abs().max())
        row['loss_o64'] = float(l64)
    return row


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--out', default=os.path.join(ROOT, 'gpurun_out', 'r2_grad_table.json'))
    ap.add_argument('--max_b', type=int, default=128)
    ap.add_argument('--max_b64', type=int, default=128)
    ap.add_argument('--net', default='resnet18')
    ap.add_argument('--img', type=int, default=128)
    a = ap.parse_args()
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    rows = []
    for name, sd in (('reference-init', O.reference_init_state_dict(a.net, 0)), ('synthetic', O.synthetic_state_dict(a.net, 51))):
        t64 = 0.0
        for B in (2, 8, 32, 128):
            if B > a.max_b:
                continue
            do64 = B <= a.max_b64 and t64 * 4 < 240                  # fp64 cuDNN convs: bound the time
            r = run(a.net, a.img, B, sd, do64)
            t64 = r.get('fp64_seconds', 1e9)
            r['params'] = name
            rows.append(r)
            print(json.dumps(r), flush=True)
    os.makedirs(os.path.dirname(a.out), exist_ok=True)
    json.dump(rows, open(a.out, 'w'), indent=1)
    print('| params | B | max|score| | score ours/fp64 | score fp32/fp64 | grad ours/fp64 all (worst) | grad fp32/fp64 all (worst) | ours/fp32 all |')
    print('|---|---|---|---|---|---|---|---|')
    for r in rows:
        f = lambda k: ('%.2e' % r[k]) if k in r else '-'
        print('| %s | %d | %.1f | %s | %s | %s (%s) | %s (%s) | %s |' % (
            r['params'], r['B'], r['max_abs_score'], f('score_ours_vs_o64'), f('score_o32_vs_o64'), f('ours_vs_o64_all'),
            f('ours_vs_o64_worst'), f('o32_vs_o64_all'), f('o32_vs_o64_worst'), f('ours_vs_o32_all')))


if __name__ == '__main__':
    main()
