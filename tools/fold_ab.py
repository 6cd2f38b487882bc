"""A/B of the offline merge->mlp.0 fold on the golden cases: max |cuda - ref32| on the log-couplings, fold on / off."""
import json, os, sys
import numpy as np, torch
sys.path.insert(0, '.')
from tests.util import MATCHER_CASES, GOLDEN, load_case, case_inputs
from e2e_multi_view_matching_b200.models.multi_view_matcher import MultiViewMatcher
rep = json.load(open(os.path.join(GOLDEN, 'matcher_report.json')))
for name in MATCHER_CASES:
    meta, ref = load_case(name)
    sd, data = case_inputs(meta)
    line = '%-20s noise %.2e ' % (name, rep[name]['max_abs_ref32_vs_ref64'])
    for fold in (True, False):
        model = MultiViewMatcher({'multi_frame_matching': meta['multi'], 'GNN_layers': meta['layers'], 'conf_mlp': True,
                                  'fold_merge': fold}).eval()
        model.load_state_dict({k: torch.from_numpy(np.asarray(v)) for k, v in sd.items()}, strict=True)
        model = model.cuda()
        out = model({k: (torch.from_numpy(v).cuda() if isinstance(v, np.ndarray) else v) for k, v in data.items()})
        err = max(float(np.abs(out[k].cpu().numpy() - v).max()) for k, v in ref.items() if k.startswith('scores_'))
        mism = sum(int((out[k].cpu().numpy() != v).sum()) for k, v in ref.items() if k.startswith('matches'))
        line += ' fold=%d err %.2e mism %d' % (fold, err, mism)
    print(line, flush=True)
