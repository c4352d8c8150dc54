"""Recipe for `oracle/_ref/`: the UNMODIFIED reference files of the fit path, so that the GPU box (which has no
/root/reference) can time and check against the reference's own `NemoV2/V3.step` (bench.py `--impl reference`,
`cpu_baseline.kind = "reference"`).  Test infrastructure, like everything under oracle/.

    python -m oracle.build_ref          # build container only; a no-op where /root/reference is absent

The reference is pure Python (no compiled code on this path), so "building" it is: import the fit model through
`oracle/ref_harness.py` in a clean subprocess, list every module that import pulled in from /root/reference, and copy
exactly those files -- byte for byte, same relative paths -- into `oracle/_ref/` with a MANIFEST of their sha256.
`oracle/_ref/` is git-ignored (reference sources never enter the history) but travels to the GPU box with the snapshot,
like the built `.so` files.  `ref_harness.REF` falls back to it when /root/reference does not exist.
"""
from __future__ import annotations

import hashlib
import os
import shutil
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = '/root/reference'
DST = os.path.join(ROOT, 'oracle', '_ref')

_PROBE = r'''
import sys
sys.path.insert(0, %r)
from nemo_cvpr2023_b200 import assets
from oracle import ref_harness as rh
assert rh.REF == %r, rh.REF
rh.setup(assets.make_smpl_assets(0))
for m in list(sys.modules.values()):
    f = getattr(m, '__file__', None)
    if f and f.startswith(%r + '/'):
        print('REF_FILE ' + f)
'''


def build(verbose=False):
    if not os.path.isdir(os.path.join(SRC, 'nemo')):
        return os.path.isdir(os.path.join(DST, 'nemo'))
    out = subprocess.run([sys.executable, '-c', _PROBE % (ROOT, SRC, SRC)], capture_output=True, text=True, cwd=ROOT)
    if out.returncode:
        raise RuntimeError('reference import failed:\n' + out.stderr[-2000:])
    files = sorted({l[9:] for l in out.stdout.splitlines() if l.startswith("REF_FILE ")})
    if os.path.isdir(DST):
        shutil.rmtree(DST)
    lines = []
    for f in files:
        rel = os.path.relpath(f, SRC)
        dst = os.path.join(DST, rel)
        os.makedirs(os.path.dirname(dst), exist_ok=True)
        shutil.copyfile(f, dst)
        lines.append('%s  %s' % (hashlib.sha256(open(f, 'rb').read()).hexdigest(), rel))
    open(os.path.join(DST, 'MANIFEST.sha256'), 'w').write('\n'.join(lines) + '\n')
    if verbose:
        print('oracle/_ref: %d reference files, %d KB' % (len(files), sum(os.path.getsize(f) for f in files) // 1024))
    return True


if __name__ == '__main__':
    build(verbose=True)
