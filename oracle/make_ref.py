"""TEST / BENCH INFRASTRUCTURE (not product): recipe that stages the UNMODIFIED reference model modules for the
bench's reference arms.

    python oracle/make_ref.py            # copies, when /root/reference is present

The reference (TengdaHan/DPC) has no setup.py / pyproject, so `pip install /root/reference` is impossible; its model
path is four pure-Python files that need only torch.  They are copied -- byte for byte, never edited -- into the
git-ignored `baseline/_ref/`, which travels to the GPU box with the gpurun snapshot exactly like the built `.so`
(no reference source ever enters the git history).  `/root/reference` does not exist on the GPU box: bench.py only
ever reads `baseline/_ref/`, and falls back to the oracle port (oracle/dpc_oracle.py) when it is absent.

    dpc/model_3d.py             DPC_RNN                      (the module under test)
    backbone/resnet_2d3d.py     ResNet2d3d_full, blocks
    backbone/convrnn.py         ConvGRU
    backbone/select_backbone.py select_resnet
"""
import filecmp
import os
import shutil
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.environ.get('DPC_REFERENCE', '/root/reference')
DST = os.path.join(ROOT, 'baseline', '_ref')
FILES = ['dpc/model_3d.py', 'backbone/resnet_2d3d.py', 'backbone/convrnn.py', 'backbone/select_backbone.py']


def make(verbose=False):
    """-> True if baseline/_ref holds the four reference modules afterwards"""
    if os.path.isdir(SRC):
        os.makedirs(DST, exist_ok=True)
        for f in FILES:
            s, d = os.path.join(SRC, f), os.path.join(DST, os.path.basename(f))
            if not (os.path.exists(d) and filecmp.cmp(s, d, shallow=False)):
                shutil.copyfile(s, d)
                if verbose:
                    print('staged', f)
    return all(os.path.exists(os.path.join(DST, os.path.basename(f))) for f in FILES)


def import_reference():
    """-> the reference's `model_3d` module imported from baseline/_ref, or None when it is not staged"""
    if not all(os.path.exists(os.path.join(DST, os.path.basename(f))) for f in FILES):
        return None
    if DST not in sys.path:
        sys.path.insert(0, DST)
    import importlib
    return importlib.import_module('model_3d')


if __name__ == '__main__':
    ok = make(verbose=True)
    print('baseline/_ref %s' % ('ready' if ok else 'NOT available (no %s here)' % SRC))
