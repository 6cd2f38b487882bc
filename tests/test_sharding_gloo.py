"""CPU, world_size 2 over gloo: the host logic of the multi-GPU path (tuple shards, the per-step scalar
all-reduce, max-over-ranks timing) and bench.py's reference arm under a process group (rank 0 alone works)."""
import os
import subprocess
import sys

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group('gloo', rank=rank, world_size=world)
    from e2e_multi_view_matching_b200 import sharding
    ids = sharding.tuple_shard(rank, world, 14)
    loss = sharding.all_reduce_step_loss(torch.tensor([float(rank + 1)]))
    mx = sharding.max_over_ranks(10.0 + rank)
    gathered = [None] * world
    dist.all_gather_object(gathered, ids)
    if rank == 0:
        flat = [i for g in gathered for i in g]
        assert flat == list(range(1000, 1000 + 14 * world)), flat          # disjoint, contiguous, complete
        assert loss.item() == sum(range(1, world + 1))
        assert mx == 10.0 + world - 1
        assert sharding.whole_job_throughput(14, 10, world, 1000.0) == 14 * 10 * world
        open(out, 'w').write('ok')
    dist.barrier()
    dist.destroy_process_group()


def test_shards_and_collectives_world2(tmp_path):
    out = str(tmp_path / 'ok.txt')
    mp.spawn(_worker, args=(2, 29531, out), nprocs=2, join=True)
    assert open(out).read() == 'ok'


def _grad_worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group('gloo', rank=rank, world_size=world)
    from e2e_multi_view_matching_b200 import sharding
    torch.manual_seed(0)
    params = [torch.nn.Parameter(torch.zeros(5, 3)), torch.nn.Parameter(torch.zeros(7)), torch.nn.Parameter(torch.zeros(2, 2))]
    params[0].grad = torch.full((5, 3), float(rank + 1))
    params[1].grad = torch.arange(7.0) * (rank + 1)
    if rank == 0:
        params[2].grad = torch.ones(2, 2)                    # rank 1 has no gradient for this parameter
    calls = sharding.all_reduce_gradients(params, bucket_bytes=64)      # small buckets: several collectives
    mean = sum(range(1, world + 1)) / world
    assert calls >= 2
    assert torch.allclose(params[0].grad, torch.full((5, 3), mean))
    assert torch.allclose(params[1].grad, torch.arange(7.0) * mean)
    assert torch.allclose(params[2].grad, torch.ones(2, 2) / world)
    if rank == 0:
        open(out, 'w').write('ok')
    dist.barrier()
    dist.destroy_process_group()


def test_gradient_all_reduce_world2(tmp_path):
    out = str(tmp_path / 'ok.txt')
    mp.spawn(_grad_worker, args=(2, 29533, out), nprocs=2, join=True)
    assert open(out).read() == 'ok'


def _train_worker(rank, world, port, out):
    """training.train_step under a process group, on the float64 stand-ins of the stage kernels: every rank trains on its
    own tuple, the gradients are averaged (DDP's all-reduce), the ranks end with identical parameters, equal to one SGD
    step with the mean of the two per-rank gradients."""
    import json
    import types
    import numpy as np
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group('gloo', rank=rank, world_size=world)
    torch.set_num_threads(2)
    from tests import emul_ops
    from tests.test_train_host_logic import PATCHED, match_loss
    from oracle.make_train_backward_golden import build, CASES
    from e2e_multi_view_matching_b200 import ops, _lib, training
    from e2e_multi_view_matching_b200.models.multi_view_matcher import MultiViewMatcher
    for f in PATCHED:
        setattr(ops, f, getattr(emul_ops, f))
    _lib.require_cuda = lambda device, what: None
    training.compute_match_loss = match_loss                       # (the product's is a CUDA kernel)
    case = CASES[0]
    data_np, sd = build(case)

    def model_and_data(r):
        model = MultiViewMatcher({'multi_frame_matching': True, 'GNN_layers': case['layers'], 'conf_mlp': False, 'full_output': False})
        model.load_state_dict({k: torch.from_numpy(np.asarray(v)) for k, v in sd.items() if not k.startswith('conf_mlp')})
        data = {k: (torch.from_numpy(v[r:r + 1]) if isinstance(v, np.ndarray) and v.ndim >= 1 and v.shape[0] == case['batch'] else
                    (torch.from_numpy(v) if isinstance(v, np.ndarray) else v)) for k, v in data_np.items()}
        return model.train(), data
    opt = types.SimpleNamespace(pose_loss=False)
    model, data = model_and_data(rank)
    optimizer = torch.optim.SGD(model.parameters(), lr=1e-9)
    loss, _ = training.train_step(opt, data, model, optimizer, n_pairs=3)
    mine = torch.cat([p.detach().reshape(-1) for p in model.parameters()])
    both = [torch.zeros_like(mine) for _ in range(world)]
    dist.all_gather(both, mine)
    assert torch.equal(both[0], both[1])                            # synchronised replicas
    if rank == 0:
        # reference: the two per-rank gradients computed locally, averaged by hand
        grads = []
        for r in range(world):
            m2, d2 = model_and_data(r)
            ls, _ = training.run_matcher(opt, d2, m2)
            (ls['match_loss'] / 3.0).sum().backward()
            grads.append(torch.cat([p.grad.reshape(-1) for p in m2.parameters()]))
        m0, _ = model_and_data(0)
        p0 = torch.cat([p.detach().reshape(-1) for p in m0.parameters()])
        expect = p0 - 1e-9 * (grads[0] + grads[1]) / 2
        assert torch.allclose(mine, expect, rtol=0, atol=1e-7 * float(p0.abs().max())), float((mine - expect).abs().max())
        assert not torch.equal(mine, p0) and np.isfinite(float(loss))
        open(out, 'w').write('ok')
    dist.barrier()
    dist.destroy_process_group()


def test_data_parallel_train_step_world2(tmp_path):
    out = str(tmp_path / 'ok.txt')
    mp.spawn(_train_worker, args=(2, 29534, out), nprocs=2, join=True)
    assert open(out).read() == 'ok'


def test_no_process_group_is_a_noop():
    from e2e_multi_view_matching_b200 import sharding
    assert sharding.tuple_shard(0, 1, 3) == [1000, 1001, 1002]
    assert sharding.all_reduce_step_loss(torch.tensor([2.0])).item() == 2.0
    assert sharding.max_over_ranks(3.5) == 3.5
    assert sharding.all_reduce_gradients([torch.nn.Parameter(torch.zeros(2))]) == 0


def test_reference_arm_only_rank0_prints():
    """bench.py --impl reference under torchrun-style env vars: ranks > 0 exit 0 without work or output."""
    env = dict(os.environ, RANK='1', WORLD_SIZE='2', LOCAL_RANK='1', MASTER_ADDR='127.0.0.1', MASTER_PORT='29532')
    r = subprocess.run([sys.executable, os.path.join(ROOT, 'bench.py'), '--impl', 'reference', '--gpus', '2', '--steps', '1',
                        '--warmup', '0'], capture_output=True, text=True, env=env, timeout=120)
    assert r.returncode == 0 and r.stdout.strip() == '', (r.returncode, r.stdout[-200:], r.stderr[-300:])
