"""`Matching` shim of upstream SuperGlue (models/matching.py upstream; absent from the fork, see
SURVEY.md §2): runs a keypoint front-end when keypoints are not supplied, then SuperGlue, and
merges both outputs.  The SuperPoint front-end is out of scope for the kernels (SURVEY.md §8
f-3): pass any nn.Module producing keypoints/scores/descriptors lists as `superpoint`."""
import torch

from .superglue import SuperGlue


class Matching(torch.nn.Module):
    def __init__(self, config={}, superpoint=None):
        super().__init__()
        self.superpoint = superpoint
        self.superglue = SuperGlue(config.get('superglue', {}))

    def forward(self, data):
        pred = {}
        for i in ('0', '1'):
            if 'keypoints' + i not in data:
                if self.superpoint is None:
                    raise ValueError('keypoints%s missing and no front-end given' % i)
                p = self.superpoint({'image': data['image' + i]})
                pred.update({k + i: v for k, v in p.items()})
        data = {**data, **pred}
        for k in data:
            if isinstance(data[k], (list, tuple)):
                data[k] = torch.stack(data[k])
        pred = {**pred, **self.superglue(data)}
        return pred
