"""gpurun_out/<tag>_<key>.raw.csv + .details.txt (scripts/ncu_capture.sh reduces every `ncu --set full` report on the GPU box
to its raw-metric CSV and details page) -> profiles/<tag>_ncu_<key>.txt, profiles/<tag>_ncu_summary.json and
profiles/<tag>_launches*.txt.  No GPU needed.   python scripts/make_ncu_summary.py r2"""
import csv
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ROWS1 = 1024 * 5 * 32 * 32          # layer1 positions at B = 128 (1024 blocks x 5 frames x 32 x 32)
ROWS0 = 1024 * 5 * 64 * 64          # conv1 output positions
ROWSP = ROWS0 // 4                  # pooled positions
M2, M5 = 6144, 6468
HBM_GBS, BF16_TF = 6650.0, 1720.0   # fallbacks; MEASURED_PEAKS.json wins


def peaks():
    p = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.exists(p):
        d = json.load(open(p))
        return d.get('hbm_gbs', HBM_GBS), d.get('bf16_tflops', BF16_TF)
    return HBM_GBS, BF16_TF


# key -> (site, bound, algorithmic flops per launch, algorithmic bytes per launch)
SITES = {
    'conv_l3': ('layer3.1.conv1 fwd 256->256 3x3x3 [1024,3,8,8]', 'tensor', 695784701952, 409731072),
    'conv_l1': ('layer1.0.conv1 fwd 64->64 1x3x3 [1024,5,32,32] (halo-patch kernel)', 'tensor', 2 * ROWS1 * 64 * 64 * 9, 2 * ROWS1 * 64 * 4),
    'conv_l2': ('layer2.0.conv2 fwd 128->128 1x3x3 [1024,5,16,16] (persistent kernel)', 'tensor', 2 * (ROWS1 // 4) * 128 * 128 * 9, 2 * (ROWS1 // 4) * 128 * 4),
    'wgrad_l3': ('layer3.1.conv2 wgrad 256x256x27', 'tensor', 695784701952, 402653184),
    'wgrad_l1': ('layer1.1.conv2 wgrad 64x64x9 (halo-patch kernel)', 'tensor', 2 * ROWS1 * 64 * 64 * 9, 2 * ROWS1 * 64 * 4),
    'r34_conv_l3': ('R34 @ 224^2, B = 44: layer3 conv fwd 256->256 3x3x3 [352,3,14,14]', 'tensor', 2 * 352 * 3 * 196 * 256 * 256 * 27,
                    2 * 352 * 3 * 196 * 256 * 4),
    'stem_fwd': ('pooled stem forward: conv1 over s2d planes [1024,5,64,64,16] + bn1 sums + max-pool selection -> [.,32,32,64] fp32 + idx',
                 'tensor', 2 * ROWS0 * 64 * 147, ROWS0 * 64 + ROWSP * 64 * 5),
    'stem_bwd': ('pooled stem backward: conv1 recomputed + pool/BN backward + conv1 wgrad in one kernel', 'tensor',
                 4 * ROWS0 * 64 * 147, ROWS0 * 64 + ROWSP * 64 * 5),
    'bn_bwd': ('layer1 BN backward apply pass [5242880,64]', 'hbm', 0, 6039797760),
    'head_fwd': ('recurrent head forward (7 GRU steps + 3 predictions, R = 2048 rows) in one kernel', 'latency', 2 * 2048 * 256 * (7 * 1536 + 3 * 512), 0),
    'head_bwd': ('recurrent head backward in one kernel', 'latency', 2 * 2048 * 256 * (7 * 1536 + 3 * 512), 0),
    'score_fwd': ('score matmul fwd [6144,256] x [256,6144] -> fp32 (persistent A-resident kernel, fp16 pairs)', 'hbm',
                  2 * M2 * M2 * 256, M2 * M2 * 4 + 2 * M2 * 256 * 4),
    'score_dpred': ('score bwd: d(pred) = dS . finf  [6144,6144] x [6144,256]', 'tensor', 2 * M2 * M2 * 256, M2 * M2 * 4 + 2 * M2 * 256 * 4),
    'score_dfinf': ('score bwd: d(finf) = dS^T . pred (wgrad form)', 'tensor', 2 * M2 * M2 * 256, M2 * M2 * 4 + 2 * M2 * 256 * 4),
    'ce_fwd': ('NCE cross-entropy + top-k forward over the [6144,6144] score', 'hbm', 0, M2 * M2 * 4),
    'ce_bwd': ('NCE cross-entropy backward (softmax - onehot) -> d(score)', 'hbm', 0, 2 * M2 * M2 * 4),
    'augment': ('input pipeline: 128 clips x 40 uint8 frames 200x150 -> block [128,8,3,5,128,128] fp32 (crop/resize/flip/grey/jitter/normalise)',
                'hbm', 0, 128 * 40 * (150 * 200 * 3 + 3 * 128 * 128 * 4)),
}
KEYS = ['gpu__time_duration.sum', 'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'dram__bytes_read.sum', 'dram__bytes_write.sum', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'lts__t_sector_hit_rate.pct', 'launch__registers_per_thread', 'launch__grid_size', 'launch__block_size',
        'launch__shared_mem_per_block_dynamic', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'smsp__cycles_active.avg', 'sm__inst_executed.sum']


def raw(path):
    rows = list(csv.reader(open(path, errors='replace')))
    rows = [r for r in rows if len(r) > 10]
    return dict(zip(rows[0], rows[2])), dict(zip(rows[0], rows[1]))


def num(d, u, k):
    v = float(d[k].replace(',', ''))
    unit = u.get(k, '')
    scale = {'Gbyte': 1e9, 'Mbyte': 1e6, 'Kbyte': 1e3, 'byte': 1.0, 'ms': 1.0, 'us': 1e-3, 'ns': 1e-6, 's': 1e3,
             'msecond': 1.0, 'usecond': 1e-3, 'nsecond': 1e-6}.get(unit, 1.0)
    return v * scale


def main(tag):
    hbm, bf16 = peaks()
    summary = {}
    for key, (site, bound, flops, nbytes) in SITES.items():
        rp = os.path.join(ROOT, 'gpurun_out', '%s_%s.raw.csv' % (tag, key))
        if not os.path.exists(rp):
            print('missing', rp)
            continue
        d, u = raw(rp)
        ms = num(d, u, 'gpu__time_duration.sum')
        dram = num(d, u, 'dram__bytes_read.sum') + num(d, u, 'dram__bytes_write.sum')
        e = dict(kernel=d['Kernel Name'].split('(')[0].replace('<unnamed>::', ''), site=site, bound=bound, flops=flops, bytes=nbytes,
                 ncu_duration_ms=round(ms, 4), dram_bytes=int(dram),
                 tensor_pipe_active_pct=float(d['sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active']),
                 issue_active_pct=float(d['smsp__issue_active.avg.pct_of_peak_sustained_active']),
                 dram_pct_of_peak=float(d['gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed']))
        if nbytes:
            e['traffic_over_algorithmic'] = round(dram / nbytes, 2)
        if flops:
            e['achieved_tflops_algorithmic'] = round(flops / ms / 1e9, 1)
            if bound == 'tensor':
                e['frac_of_bf16_burst'] = round(flops / ms / 1e9 / bf16, 3)
                e['executed_mma_passes'] = 3
        if bound == 'hbm':
            e['achieved_gbs_algorithmic'] = round(nbytes / ms / 1e6, 1)
            e['frac_of_hbm'] = round(nbytes / ms / 1e6 / hbm, 3)
        summary[key] = e
        lines = ['# ncu --set full --clock-control none: %s' % site, '# kernel: %s' % d['Kernel Name'][:160], '']
        lines += ['%-72s %s %s' % (k, d[k], u.get(k, '')) for k in KEYS if k in d]
        lines += ['', 'derived: ' + json.dumps({k: v for k, v in e.items() if k not in ('kernel', 'site')}), '']
        dt = os.path.join(ROOT, 'gpurun_out', '%s_%s.details.txt' % (tag, key))
        if os.path.exists(dt):
            txt = open(dt, errors='replace').read()
            keep = []
            for block in ('GPU Speed Of Light Throughput', 'Memory Workload Analysis', 'Warp State Statistics', 'Launch Statistics', 'Occupancy'):
                i = txt.find('Section: ' + block)
                if i >= 0:
                    j = txt.find('Section: ', i + 10)
                    keep.append(txt[i:j if j > 0 else None].rstrip())
            lines += keep
        open(os.path.join(ROOT, 'profiles', '%s_ncu_%s.txt' % (tag, key)), 'w').write('\n'.join(lines) + '\n')
        print(key, {k: e[k] for k in ('ncu_duration_ms', 'tensor_pipe_active_pct', 'issue_active_pct', 'dram_pct_of_peak') if k in e},
              e.get('achieved_tflops_algorithmic'), e.get('frac_of_hbm'))
    json.dump(summary, open(os.path.join(ROOT, 'profiles', '%s_ncu_summary.json' % tag), 'w'), indent=1)
    for suffix, title in (('launches', 'python scripts/profile_step.py 128 (2 train steps, B=128, R18 128^2)'),
                          ('score_launches', 'python scripts/profile_score.py 6144 (2 iterations of score matmul + NCE fwd/bwd)'),
                          ('r34_launches', 'python scripts/profile_step.py 44 resnet34 224 (2 train steps, BASELINE config 5 per-GPU shard)')):
        lc = os.path.join(ROOT, 'gpurun_out', '%s_%s.csv' % (tag, suffix))
        if os.path.exists(lc):
            txt = subprocess.run([sys.executable, os.path.join(ROOT, 'scripts', 'launch_summary.py'), lc, 'ncu launch list -- ' + title],
                                 capture_output=True, text=True).stdout
            open(os.path.join(ROOT, 'profiles', '%s_%s.txt' % (tag, suffix)), 'w').write(txt)


if __name__ == '__main__':
    main(sys.argv[1] if len(sys.argv) > 1 else 'r2')
