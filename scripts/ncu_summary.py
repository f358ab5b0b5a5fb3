"""Compact summary of an .ncu-rep (run where ncu is installed; no GPU needed):
    python scripts/ncu_summary.py gpurun_out/prof.ncu-rep > profiles/rNN_<kernel>.txt
"""
import csv
import subprocess
import sys

WANT = [
    'Kernel Name', 'Grid Size', 'Block Size',
    'gpu__time_duration.sum',
    'dram__bytes_read.sum', 'dram__bytes_write.sum',
    'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
    'lts__t_bytes.sum', 'lts__throughput.avg.pct_of_peak_sustained_elapsed',
    'sm__throughput.avg.pct_of_peak_sustained_elapsed',
    'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active',
    'sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active',
    'TPC.TriageCompute.sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed',
    'sm__mem_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed',
    'sm__inst_executed_pipe_uniform.sum', 'sm__warps_active.avg.pct_of_peak_sustained_active',
    'launch__registers_per_thread', 'launch__shared_mem_per_block_dynamic', 'launch__occupancy_limit_shared_mem',
    'sm__cycles_elapsed.max', 'smsp__cycles_active.avg',
]


def main(path):
    out = subprocess.run(['ncu', '-i', path, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    for r in rows[2:]:
        print('-' * 100)
        for i, h in enumerate(hdr):
            if h in WANT or ('pipe_tensor' in h and 'pct_of_peak_sustained_elapsed' in h and 'realtime' in h):
                print('%-95s %s %s' % (h, r[i], units[i]))
        rd = dict(zip(hdr, r))
        try:
            b = float(rd['dram__bytes_read.sum'].replace(',', '')), float(rd['dram__bytes_write.sum'].replace(',', ''))
            print('%-95s %s' % ('dram read+write (as reported units)', b))
        except Exception:
            pass


if __name__ == '__main__':
    main(sys.argv[1])
