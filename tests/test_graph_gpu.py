"""GPU: the whole matcher forward (about 180 kernel launches for the 9-layer ragged 4-view case) is captured into a
CUDA graph and replayed -- no allocation, host synchronisation or host-side table copy happens inside the C-ABI
call -- and the replay is bit-identical to the eager run."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def test_matcher_forward_is_graph_capturable():
    from tests.util import load_case, case_inputs
    from e2e_multi_view_matching_b200.models.multi_view_matcher import MultiViewMatcher
    meta, _ = load_case('mv4_ragged_sharp')
    sd, data = case_inputs(meta)
    model = MultiViewMatcher({'multi_frame_matching': True, 'GNN_layers': meta['layers'], 'conf_mlp': True}).eval()
    model.load_state_dict({k: torch.from_numpy(np.asarray(v)) for k, v in sd.items()}, strict=True)
    model = model.cuda()
    tdata = {k: (torch.from_numpy(v).cuda() if isinstance(v, np.ndarray) else v) for k, v in data.items()}
    side = torch.cuda.Stream()
    with torch.cuda.stream(side):                 # warm-up off the default stream: tensor maps, attributes, workspace
        for _ in range(3):
            eager = model(tdata)
    torch.cuda.current_stream().wait_stream(side)
    torch.cuda.synchronize()
    eager = {k: v.clone() for k, v in eager.items() if v is not None}
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph):
        out = model(tdata)
    for _ in range(2):
        for v in out.values():
            if v is not None:
                v.zero_()
        graph.replay()
    torch.cuda.synchronize()
    for k in eager:
        assert torch.equal(eager[k], out[k]), k
