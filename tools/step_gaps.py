"""Where does a bench step's time go that no kernel accounts for?  Kernel timeline of two steps from torch.profiler
(CUPTI activity records, no replay): GPU busy time vs. step span, the largest idle gaps with the kernels on either
side, and the host-side enqueue time of a step.   python tools/step_gaps.py [cfg3]"""
import json, os, sys, time, tempfile
import numpy as np
import torch
sys.path.insert(0, '.')
import bench
from e2e_multi_view_matching_b200 import _lib
from e2e_multi_view_matching_b200.pipeline import MultiViewPipeline, PairPipeline

cfg = bench.CONFIGS[sys.argv[1] if len(sys.argv) > 1 else 'cfg3']
dev = torch.device('cuda:0')
torch.cuda.set_device(dev)
from e2e_multi_view_matching_b200.models.multi_view_matcher import MultiViewMatcher
is_tuple = cfg['kind'] == 'tuple'
B = cfg['batch']
sd = bench.make_weights(cfg)
model = MultiViewMatcher({'GNN_layers': cfg['layers'], 'multi_frame_matching': is_tuple}).eval()
model.load_state_dict({k: torch.from_numpy(np.asarray(v)) for k, v in sd.items()})
model = model.to(dev)
pipe = MultiViewPipeline(model) if is_tuple else PairPipeline(model, eval_mode='w8pt_ba')
data_np = bench.make_inputs(cfg, cfg['seed_base'], B)
data_dev = {k: torch.from_numpy(v).to(dev) for k, v in data_np.items() if isinstance(v, np.ndarray) and not k.startswith(('image', 'landmark'))}
data_dev.update({k: torch.empty(v.shape, device='meta') for k, v in data_np.items() if k.startswith('image')})
data_dev['ids'] = data_np['ids']
for _ in range(3):
    pipe(data_dev)
torch.cuda.synchronize()
# host enqueue time
ts = []
for _ in range(3):
    torch.cuda.synchronize()
    t0 = time.time()
    pipe(data_dev)
    t1 = time.time()
    torch.cuda.synchronize()
    t2 = time.time()
    ts.append((t1 - t0, t2 - t0))
print('host enqueue per step: %.1f ms   (step incl. GPU drain %.1f ms)' % (1e3 * np.median([a for a, b in ts]), 1e3 * np.median([b for a, b in ts])))
from torch.profiler import profile, ProfilerActivity
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    for _ in range(2):
        pipe(data_dev)
    torch.cuda.synchronize()
path = os.path.join(tempfile.gettempdir(), 'step_trace.json')
prof.export_chrome_trace(path)
ev = [e for e in json.load(open(path))['traceEvents'] if e.get('cat') in ('kernel', 'gpu_memcpy', 'gpu_memset')]
ev.sort(key=lambda e: e['ts'])
span = ev[-1]['ts'] + ev[-1]['dur'] - ev[0]['ts']
busy = 0.0
end = ev[0]['ts']
gaps = []
for i, e in enumerate(ev):
    if e['ts'] > end:
        gaps.append((e['ts'] - end, ev[i - 1]['name'].replace('(anonymous namespace)::', '')[:50] if i else '-', e['name'].replace('(anonymous namespace)::', '')[:50]))
        busy += e['dur']
    else:
        busy += max(0.0, e['ts'] + e['dur'] - end)
    end = max(end, e['ts'] + e['dur'])
print('2 steps: span %.2f ms, GPU busy %.2f ms, idle %.2f ms over %d gaps (%d kernels)' % (span / 1e3, busy / 1e3, (span - busy) / 1e3, len(gaps), len(ev)))
gaps.sort(reverse=True)
print('largest gaps (us): before <- after')
for g, a, b in gaps[:25]:
    print('  %8.1f   %s  ->  %s' % (g, a, b))
h = np.array([g for g, _, _ in gaps])
for lo, hi in ((0, 2), (2, 5), (5, 10), (10, 50), (50, 1e9)):
    m = (h >= lo) & (h < hi)
    print('  gaps in [%g, %g) us: n=%d total %.2f ms' % (lo, hi, m.sum(), h[m].sum() / 1e3))
by = {}
for e in ev:
    k = e['name'].replace('(anonymous namespace)::', '').replace('void ', '').split('(')[0][:70]
    by.setdefault(k, [0, 0.0])
    by[k][0] += 1; by[k][1] += e['dur']
print('kernel time by name (2 steps):')
for k, (n, t) in sorted(by.items(), key=lambda x: -x[1][1])[:30]:
    print('  %-72s n=%4d %9.2f ms' % (k, n, t / 1e3))
