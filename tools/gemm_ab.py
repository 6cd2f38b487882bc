"""Layer-GEMM shapes of the bench step on the persistent fp16x3 kernel: kernel durations from CUPTI (torch.profiler),
optionally with the experiment flags of a debug build (mvm_debug_set_gemm_exp, not in release builds)."""
import sys, json, os, tempfile
import numpy as np
import torch
sys.path.insert(0, '.')
from e2e_multi_view_matching_b200 import ops, _lib
from torch.profiler import profile, ProfilerActivity
lib = _lib.lib()
M = 14 * 5 * 1024
g = torch.Generator().manual_seed(0)
shapes = [('qkv 256->768', 256, 0, 768), ('mlp.0 512->512', 256, 256, 512), ('mlp.2 512->256', 512, 0, 256)]
flush = torch.empty(256 << 20, dtype=torch.uint8, device='cuda')
exps = [0]
if hasattr(lib, 'mvm_debug_set_gemm_exp'):
    exps = [0, 1, 2, 3]
for e in exps:
    if len(exps) > 1:
        lib.mvm_debug_set_gemm_exp(e)
    for name, K1, K2, N in shapes:
        a = torch.randn(M, K1, generator=g).cuda()
        a2 = torch.randn(M, K2, generator=g).cuda() if K2 else None
        w = (torch.randn(N, K1 + K2, generator=g) / 16).cuda()
        b = torch.randn(N, generator=g).cuda()
        for _ in range(2):
            ops.linear(a, w, bias=b, a2=a2, relu=True, tc_passes='h16')
        torch.cuda.synchronize()
        with profile(activities=[ProfilerActivity.CUDA]) as prof:
            for _ in range(5):
                flush.zero_()
                ops.linear(a, w, bias=b, a2=a2, relu=True, tc_passes='h16')
            torch.cuda.synchronize()
        path = os.path.join(tempfile.gettempdir(), 'gemm_trace.json')
        prof.export_chrome_trace(path)
        d = [ev['dur'] for ev in json.load(open(path))['traceEvents'] if ev.get('cat') == 'kernel' and 'gemm_tc_persist' in ev['name']]
        fl = 2.0 * M * (K1 + K2) * N
        print('exp %d  %-16s median %7.1f us  (%5.0f TFLOP/s algorithmic, x3 tensor passes)  n=%d' % (e, name, np.median(d), fl / np.median(d) / 1e6, len(d)))
if len(exps) > 1:
    lib.mvm_debug_set_gemm_exp(0)
    print('exp bit 0: splitters do not read the A tile from shared memory; bit 1: the A tile is not TMA-loaded')
