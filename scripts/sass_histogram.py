"""SASS instruction histogram of libdpc_b200.so (cuobjdump -sass): per kernel, the counts of the tensor-core / TMA /
mbarrier instructions that prove the sm_100a path (UTCHMMA / UTCQMMA = tcgen05.mma, UTMALDG = TMA load, UTCBAR =
tcgen05.commit, SYNCS = mbarrier, LDTM / STTM = TMEM load/store), plus the FFMA / HMMA counts that would indicate a CUDA-core
or legacy mma.sync inner loop.
usage: python scripts/sass_histogram.py [lib] > profiles/r2_sass_histogram.txt"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, 'dpc_b200', 'libdpc_b200.so')
sass = subprocess.run(['cuobjdump', '-sass', lib], capture_output=True, text=True, check=True).stdout
KEYS = ['UTCHMMA', 'UTCQMMA', 'UTMALDG', 'UTMASTG', 'UTCBAR', 'SYNCS', 'LDTM', 'STTM', 'UTCATOM', 'HMMA', 'FFMA', 'DFMA',
        'ATOMG', 'RED', 'LDG', 'STG', 'LDS', 'STS', 'SHFL']
per = collections.OrderedDict()
cur = None
for ln in sass.splitlines():
    m = re.search(r'Function : (\S+)', ln)
    if m:
        cur = m.group(1)
        per[cur] = collections.Counter()
        continue
    m = re.match(r'\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)', ln)
    if m and cur:
        per[cur][m.group(1).split('.')[0]] += 1
        per[cur]['_total'] += 1


def demangle(n):
    try:
        full = subprocess.run(['c++filt', n], capture_output=True, text=True).stdout.strip()
        return re.sub(r'^void ', '', full.replace('(anonymous namespace)::', '')).split('(')[0]
    except Exception:
        return n


print('SASS instruction histogram of %s (sm_100a), %d kernels' % (os.path.basename(lib), len(per)))
print('%-58s %7s ' % ('kernel', 'instrs') + ' '.join('%7s' % k for k in KEYS))
tot = collections.Counter()
for k, c in per.items():
    tot.update(c)
    name = demangle(k)[:58]
    print('%-58s %7d ' % (name, c['_total']) + ' '.join('%7d' % c[x] for x in KEYS))
print('%-58s %7d ' % ('TOTAL', tot['_total']) + ' '.join('%7d' % tot[x] for x in KEYS))
tc = [k for k, c in per.items() if c['UTCHMMA'] or c['UTCQMMA']]
print('\nkernels issuing tcgen05.mma (UTCHMMA): %d; kernels with TMA loads (UTMALDG): %d; legacy HMMA (mma.sync) anywhere: %d'
      % (len(tc), sum(1 for c in per.values() if c['UTMALDG']), tot['HMMA']))
