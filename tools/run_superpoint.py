"""Timing of the SuperPoint front-end (dense part + keypoint extraction + descriptor sampling) on 480 x 640 images."""
import sys
import numpy as np
import torch
sys.path.insert(0, '.')
from e2e_multi_view_matching_b200.models.superpoint import SuperPoint
from e2e_multi_view_matching_b200.synthetic import make_superpoint_state_dict, make_image

B = int(sys.argv[1]) if len(sys.argv) > 1 else 5
sp = SuperPoint({'max_keypoints': 1024}).eval()
sp.load_state_dict({k: torch.from_numpy(v) for k, v in make_superpoint_state_dict(1).items()})
sp = sp.cuda()
img = torch.from_numpy(make_image(3, 480, 640, B)).cuda()
for _ in range(2):
    sp.dense(img)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(5):
    sp.dense(img)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 5
flops = 52.0e9 * B
print('dense part, %d images 480x640: %.2f ms (%.2f ms / image, %.1f TFLOP/s fp32)' % (B, ms, ms / B, flops / ms / 1e9))
e0.record()
for _ in range(3):
    out = sp({'image': [img]})
e1.record(); torch.cuda.synchronize()
print('forward incl. keypoint extraction + sampling: %.2f ms / image; %d keypoints' % (e0.elapsed_time(e1) / 3 / B, out['keypoints'][0].shape[0]))
