"""Quick device timing of the matcher forward (not the bench contract; see bench.py)."""
import sys, time
import numpy as np
import torch
sys.path.insert(0, '.')
from e2e_multi_view_matching_b200.models.multi_view_matcher import MultiViewMatcher
from oracle.weights import make_state_dict, make_view_inputs

def main(T=5, N=1024, B=1, layers=None, reps=5):
    layers = layers or (['self'] + ['cross'] * 3) * 7
    sd = make_state_dict(len(layers), seed=0)
    model = MultiViewMatcher({'GNN_layers': layers}).eval()
    model.load_state_dict({k: torch.from_numpy(np.asarray(v)) for k, v in sd.items()})
    model = model.cuda()
    data = make_view_inputs(1000, [N] * T, batch=B)
    data = {k: (torch.from_numpy(v).cuda() if isinstance(v, np.ndarray) else v) for k, v in data.items()}
    for _ in range(2):
        out = model(data)
    torch.cuda.synchronize()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(reps + 1)]
    ev[0].record()
    for i in range(reps):
        out = model(data)
        ev[i + 1].record()
    torch.cuda.synchronize()
    ts = [ev[i].elapsed_time(ev[i + 1]) for i in range(reps)]
    print('T=%d N=%d B=%d layers=%d: ms per forward' % (T, N, B, len(layers)), ['%.2f' % t for t in ts])

if __name__ == '__main__':
    import e2e_multi_view_matching_b200 as pkg
    for mode in (0, 3, 1):
        pkg.set_math_mode(mode)
        print('math mode', mode)
        main(B=4)
    sys.exit(0)
    main()
    main(T=2, N=1024, B=8, layers=['self', 'cross'] * 9)
