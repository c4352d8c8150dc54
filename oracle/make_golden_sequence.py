"""Generates tests/golden/demo_sequence.npz: the `sequences` of the UNMODIFIED reference loader
`nemo.multi_view_sequence.DemoMultiViewSequence` (/root/reference, stubbed heavy imports) on the seeded directory tree
of tests/util.make_demo_tree.  Build container only:  python -m oracle.make_golden_sequence"""
import os
import sys
import tempfile
from unittest.mock import MagicMock

import numpy as np

from . import ref_harness as H


def main():
    H._install_stubs()
    for name in ['hmr.penn_action', 'hmr.video', 'hmr.hmr_model', 'hmr.img_utils', 'scipy.io']:
        if name not in sys.modules:
            try:
                __import__(name)
            except Exception:
                sys.modules[name] = MagicMock(name=name)
    sys.path.insert(0, H.REF)
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    from nemo import multi_view_sequence as R
    from tests.util import make_demo_tree
    out = {}
    with tempfile.TemporaryDirectory() as root:
        cfg = make_demo_tree(root)
        seq = R.DemoMultiViewSequence(cfg, 0, 9)
        out['num_frames'], out['num_views'] = seq.num_frames, seq.num_views
        out['IMG'] = np.array([seq.IMG_D0, seq.IMG_D1])
        out['framerate_multiplier'] = np.array(seq.framerate_multiplier)
        for v, s in enumerate(seq.sequences):
            for k in s:
                out['v%d_%s' % (v, k)] = np.array(s[k])
        out['image_0_2'] = seq.get_image(1, 2)
    dst = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'tests', 'golden', 'demo_sequence.npz')
    np.savez_compressed(dst, **out)
    print('wrote', dst, {k: np.shape(v) for k, v in out.items()})


if __name__ == '__main__':
    main()
