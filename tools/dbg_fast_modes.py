import os, sys, subprocess
import numpy as np
sys.path.insert(0, '/root/repo')
if len(sys.argv) > 1:
    from orb_slam3_ros_b200 import synth
    from orb_slam3_ros_b200.extractor import ORBextractor
    img = synth.frame(2160, 3840, 5)
    ge = ORBextractor(8000, 1.2, 12)
    ge(img, None, (0, 0))
    out = {}
    for l in range(12):
        out[f"k{l}"] = ge.debug_raw_keys(0, l)
    np.savez(sys.argv[1], **out)
else:
    env = dict(os.environ)
    subprocess.check_call([sys.executable, __file__, '/tmp/band.npz'], env=env)
    env['ORBB_FAST_MODE'] = 'split'
    subprocess.check_call([sys.executable, __file__, '/tmp/split.npz'], env=env)
    a, b = np.load('/tmp/band.npz'), np.load('/tmp/split.npz')
    for l in range(12):
        ka, kb = a[f"k{l}"], b[f"k{l}"]
        print(l, ka.shape, kb.shape, np.array_equal(ka, kb))
        if ka.shape == kb.shape and not np.array_equal(ka, kb):
            d = np.flatnonzero((ka != kb).any(1))
            print(' first diffs', d[:5], ka[d[:5]], kb[d[:5]])
        elif ka.shape != kb.shape:
            sa = set(map(tuple, ka.tolist())); sb = set(map(tuple, kb.tolist()))
            print(' only band', sorted(sa - sb)[:8]); print(' only split', sorted(sb - sa)[:8])
