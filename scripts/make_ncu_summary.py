"""gpurun_out/<tag>_*.ncu-rep (scripts/ncu_capture.sh) -> profiles/<tag>_ncu_<key>.txt, profiles/<tag>_ncu_summary.json
and profiles/<tag>_launches.txt.  Runs where ncu is installed; no GPU needed.   python scripts/make_ncu_summary.py r2"""
import csv
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ROWS1 = 1024 * 5 * 32 * 32          # layer1 positions at B = 128 (1024 blocks x 5 frames x 32 x 32)
ROWS0 = 1024 * 5 * 64 * 64          # conv1 output positions
HBM_GBS, BF16_TF = 6650.0, 1720.0   # fallbacks; MEASURED_PEAKS.json wins


def peaks():
    p = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.exists(p):
        d = json.load(open(p))
        return d.get('hbm_gbs_burst', d.get('hbm_gbs', HBM_GBS)), d.get('bf16_tflops_burst', d.get('bf16_tflops', BF16_TF))
    return HBM_GBS, BF16_TF


# key -> (site, bound, algorithmic flops per launch, algorithmic bytes per launch)
SITES = {
    'conv_l3': ('layer3.1.conv1 fwd 256->256 3x3x3 [1024,3,8,8]', 'tensor', 695784701952, 409731072),
    'conv_l1': ('layer1.0.conv1 fwd 64->64 1x3x3 [1024,5,32,32] (halo-patch kernel)', 'tensor', 2 * ROWS1 * 64 * 64 * 9, 2 * ROWS1 * 64 * 4),
    'conv_l2': ('layer2.0.conv2 fwd 128->128 1x3x3 [1024,5,16,16]', 'tensor', 2 * (ROWS1 // 4) * 128 * 128 * 9, 2 * (ROWS1 // 4) * 128 * 4),
    'wgrad_l3': ('layer3.1.conv2 wgrad 256x256x27', 'tensor', 695784701952, 402653184),
    'wgrad_l1': ('layer1.1.conv2 wgrad 64x64x9 (halo-patch kernel)', 'tensor', 2 * ROWS1 * 64 * 64 * 9, 2 * ROWS1 * 64 * 4),
    'stem_fwd': ('conv1 as 4x4 conv over space-to-depth planes [1024,5,64,64,16] -> [.,64] + bn1 stats', 'hbm',
                 2 * ROWS0 * 64 * 147, 2 * ROWS0 * 32 + ROWS0 * 64 * 4),
    'stem_wgrad': ('conv1 wgrad from space-to-depth planes and dy planes', 'hbm', 2 * ROWS0 * 64 * 147, 2 * ROWS0 * 32 + ROWS0 * 64 * 4),
    'stem_tail': ('maxpool+relu+bn1 backward, apply pass', 'hbm', 0, 13421772800),
    'bn_bwd': ('layer1 BN backward apply pass [5242880,64]', 'hbm', 0, 6039797760),
}


def raw(path):
    out = subprocess.run(['ncu', '-i', path, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    return dict(zip(rows[0], rows[2])), dict(zip(rows[0], rows[1]))


def num(d, u, k):
    v = float(d[k].replace(',', ''))
    unit = u.get(k, '')
    scale = {'Gbyte': 1e9, 'Mbyte': 1e6, 'Kbyte': 1e3, 'byte': 1.0, 'ms': 1.0, 'us': 1e-3, 'ns': 1e-6, 's': 1e3}.get(unit, 1.0)
    return v * scale


def main(tag):
    hbm, bf16 = peaks()
    summary = {}
    for key, (site, bound, flops, nbytes) in SITES.items():
        rep = os.path.join(ROOT, 'gpurun_out', '%s_%s.ncu-rep' % (tag, key))
        if not os.path.exists(rep):
            print('missing', rep)
            continue
        txt = subprocess.run([sys.executable, os.path.join(ROOT, 'scripts', 'ncu_summary.py'), rep], capture_output=True, text=True).stdout
        open(os.path.join(ROOT, 'profiles', '%s_ncu_%s.txt' % (tag, key)), 'w').write(txt)
        d, u = raw(rep)
        ms = num(d, u, 'gpu__time_duration.sum')
        dram = num(d, u, 'dram__bytes_read.sum') + num(d, u, 'dram__bytes_write.sum')
        e = dict(kernel=d['Kernel Name'].split('(')[0].replace('<unnamed>::', ''), site=site, bound=bound, flops=flops, bytes=nbytes,
                 ncu_duration_ms=round(ms, 4), dram_bytes=int(dram),
                 tensor_pipe_active_pct=float(d['sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active']),
                 dram_pct_of_peak=float(d['gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed']),
                 traffic_over_algorithmic=round(dram / nbytes, 2))
        if bound == 'tensor':
            e['achieved_tflops_algorithmic'] = round(flops / ms / 1e9, 1)
            e['frac_of_bf16_burst'] = round(flops / ms / 1e9 / bf16, 3)
            e['executed_mma_passes'] = 3
        else:
            e['achieved_gbs_algorithmic'] = round(nbytes / ms / 1e6, 1)
            e['frac_of_hbm'] = round(nbytes / ms / 1e6 / hbm, 3)
        summary[key] = e
        print(key, e)
    json.dump(summary, open(os.path.join(ROOT, 'profiles', '%s_ncu_summary.json' % tag), 'w'), indent=1)
    lc = os.path.join(ROOT, 'gpurun_out', '%s_launches.csv' % tag)
    if os.path.exists(lc):
        txt = subprocess.run([sys.executable, os.path.join(ROOT, 'scripts', 'launch_summary.py'), lc,
                              'ncu launch list -- python scripts/profile_step.py 128 (2 train steps, B=128, R18 128^2)'],
                             capture_output=True, text=True).stdout
        open(os.path.join(ROOT, 'profiles', '%s_launches.txt' % tag), 'w').write(txt)


if __name__ == '__main__':
    main(sys.argv[1] if len(sys.argv) > 1 else 'r2')
