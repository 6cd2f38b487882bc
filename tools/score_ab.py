"""Where does the 3xTF32 path's distance to the fp32 reference come from?  Golden cases with the score GEMM on the
tensor cores (3xTF32) vs on the CUDA cores (fp32), and the whole matcher in fp32 (mode 0): max abs error, max
error in excess of 3e-5*|Z|, mean signed error of the log-couplings."""
import json, os, sys
import numpy as np, torch
sys.path.insert(0, '.')
import e2e_multi_view_matching_b200 as pkg
from e2e_multi_view_matching_b200 import _lib
from tests.util import MATCHER_CASES, GOLDEN, load_case, case_inputs
from e2e_multi_view_matching_b200.models.multi_view_matcher import MultiViewMatcher
lib = _lib.lib()
rep = json.load(open(os.path.join(GOLDEN, 'matcher_report.json')))
for name in MATCHER_CASES:
    meta, ref = load_case(name)
    sd, data = case_inputs(meta)
    model = MultiViewMatcher({'multi_frame_matching': meta['multi'], 'GNN_layers': meta['layers'], 'conf_mlp': True}).eval()
    model.load_state_dict({k: torch.from_numpy(np.asarray(v)) for k, v in sd.items()}, strict=True)
    model = model.cuda()
    tdata = {k: (torch.from_numpy(v).cuda() if isinstance(v, np.ndarray) else v) for k, v in data.items()}
    line = '%-20s noise %.2e |' % (name, rep[name]['max_abs_ref32_vs_ref64'])
    for label, mode, sk in (('tc/tc', 3, 1), ('tc/fp32score', 3, 0), ('fp32', 0, 0)):
        pkg.set_math_mode(mode)
        lib.mvm_debug_set_score_kernel(sk)
        out = model(tdata)
        errs, exc, bias = [], [], []
        for k, v in ref.items():
            if k.startswith('scores_'):
                g = out[k].cpu().numpy().astype(np.float64)
                d = g - v
                errs.append(np.abs(d).max()); exc.append((np.abs(d) - 3e-5 * np.abs(v)).max()); bias.append(d.mean())
        line += ' %s: max %.1e exc %.1e bias %+.1e |' % (label, max(errs), max(exc), np.mean(bias))
    pkg.set_math_mode(3); lib.mvm_debug_set_score_kernel(1)
    print(line, flush=True)
