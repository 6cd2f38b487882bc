"""Generate tests/golden/superpoint_*.npz by running the REAL reference SuperPoint (models/models/superpoint.py),
unmodified, with seeded weights on seeded images.  Authoring container only (needs /root/reference):
    python -m oracle.make_superpoint_golden
The reference constructor loads its own `weights/superpoint_v1.pth`; the fixture then replaces the parameters by the
seeded state dict of synthetic.make_superpoint_state_dict (the pretrained file cannot travel to the GPU box, the seeded
weights can be regenerated there).  Stored per case: keypoints (x, y), scores, descriptors, and float64 checksums of the
dense maps."""
import contextlib
import io
import json
import os
import sys

import numpy as np
import torch

REF = '/root/reference'
OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'tests', 'golden')

CASES = [
    dict(name='120x160_all', seed=1, height=120, width=160, batch=1, max_keypoints=-1, wseed=0),
    dict(name='240x320_top200_b2', seed=2, height=240, width=320, batch=2, max_keypoints=200, wseed=0),
    dict(name='480x640_top1024', seed=3, height=480, width=640, batch=1, max_keypoints=1024, wseed=1),
]


def main():
    if REF not in sys.path:
        sys.path.insert(0, REF)
    from models.models.superpoint import SuperPoint          # the unmodified reference
    from e2e_multi_view_matching_b200.synthetic import make_superpoint_state_dict, make_image
    torch.set_num_threads(os.cpu_count())
    os.makedirs(OUT, exist_ok=True)
    for case in CASES:
        with contextlib.redirect_stdout(io.StringIO()):
            sp = SuperPoint({'max_keypoints': case['max_keypoints']}).eval()
        sd = make_superpoint_state_dict(case['wseed'])
        sp.load_state_dict({k: torch.from_numpy(v) for k, v in sd.items()}, strict=True)
        img = make_image(case['seed'], case['height'], case['width'], case['batch'])
        with torch.no_grad():
            out = sp({'image': [torch.from_numpy(img)]})
        store = {}
        for b in range(case['batch']):
            store['keypoints%d' % b] = out['keypoints'][b].numpy().astype(np.int16)
            store['scores%d' % b] = out['scores'][b].numpy()
            store['descriptors%d' % b] = out['descriptors'][b].numpy().astype(np.float32)
        n = [int(out['keypoints'][b].shape[0]) for b in range(case['batch'])]
        # margin statistics that tell how robust the discrete decisions are: threshold margin and top-k margin
        sc = np.concatenate([out['scores'][b].numpy() for b in range(case['batch'])])
        meta = dict(case)
        meta['n_keypoints'] = n
        meta['min_threshold_margin'] = float(np.abs(sc - 0.005).min())
        np.savez_compressed(os.path.join(OUT, 'superpoint_%s.npz' % case['name']), meta=json.dumps(meta), **store)
        print(case['name'], 'keypoints', n, 'score range', float(sc.min()), float(sc.max()))


if __name__ == '__main__':
    main()
