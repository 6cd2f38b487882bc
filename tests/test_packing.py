"""CPU: weight repacking (packing.py) -- BatchNorm folding, head permutation, the tf32 hi/lo planes -- on a CPU
buffer, against direct evaluation of the reference's layer definitions."""
import numpy as np
import torch

from e2e_multi_view_matching_b200 import packing
from e2e_multi_view_matching_b200.synthetic import make_state_dict

LAYERS = ['self', 'cross', 'self']


def _packed(fold_merge=False):
    sd = {k: torch.from_numpy(np.asarray(v)) for k, v in make_state_dict(len(LAYERS), seed=3).items()}
    return sd, packing.PackedMatcher(sd, LAYERS, conf_mlp=True, device='cpu', fold_merge=fold_merge)


def _get(pm, name, shape):
    n = int(np.prod(shape))
    off = pm.offsets[name]
    total = pm.struct.hi_offset
    raw = pm.flat[off:off + n].reshape(shape)
    hi = pm.flat[total + off:total + off + n].reshape(shape)
    lo = pm.flat[2 * total + off:2 * total + off + n].reshape(shape)
    return raw, hi, lo


def test_bn_folding_matches_conv_then_batchnorm():
    sd, pm = _packed()
    x = torch.randn(5, 512, 7, dtype=torch.float64)
    w, b = sd['gnn.layers.0.mlp.0.weight'].double(), sd['gnn.layers.0.mlp.0.bias'].double()
    ref = torch.nn.functional.conv1d(x, w, b)
    ref = torch.nn.functional.batch_norm(ref, sd['gnn.layers.0.mlp.1.running_mean'].double(), sd['gnn.layers.0.mlp.1.running_var'].double(),
                                         sd['gnn.layers.0.mlp.1.weight'].double(), sd['gnn.layers.0.mlp.1.bias'].double(), False, 0.0, 1e-5)
    wf, _, _ = _get(pm, 'l0_w_mlp0', (512, 512))
    bf, _, _ = _get(pm, 'l0_b_mlp0', (512,))
    got = torch.einsum('oc,bcn->bon', wf.double(), x) + bf.double()[None, :, None]
    assert (got - ref).abs().max() < 1e-5


def test_head_permutation_reproduces_the_reference_view():
    """`.view(B, 64, 4, N)` makes channel c the pair (d = c // 4, h = c % 4); the packed q rows are head-major."""
    sd, pm = _packed()
    src = packing.head_permutation()
    assert sorted(src.tolist()) == list(range(256))
    x = torch.randn(2, 256, 9)
    q_ref = torch.nn.functional.conv1d(x, sd['gnn.layers.1.attn.proj.0.weight'], sd['gnn.layers.1.attn.proj.0.bias']).view(2, 64, 4, 9)
    wq, _, _ = _get(pm, 'l1_w_qkv', (768, 256))
    bq, _, _ = _get(pm, 'l1_b_qkv', (768,))
    q = torch.einsum('oc,bcn->bon', wq[:256], x) + bq[:256][None, :, None]          # [B, h*64 + d, N]
    torch.testing.assert_close(q.view(2, 4, 64, 9).permute(0, 2, 1, 3), q_ref, rtol=1e-5, atol=1e-5)
    # merge consumes the head-major channels: its columns carry the same permutation
    wm, _, _ = _get(pm, 'l1_w_merge', (256, 256))
    torch.testing.assert_close(wm, sd['gnn.layers.1.attn.merge.weight'][:, :, 0][:, src])


def test_merge_folded_into_mlp0():
    """Default packing: attn.merge (linear, single consumer) is folded into the message half of mlp.0 --
    mlp.0+BN(cat[x, merge(a)]) evaluated with the reference's layers == folded matrix applied to cat[x, a]."""
    sd, pm = _packed(fold_merge=True)
    assert pm.struct.layers[1].w_merge is None and 'l1_w_merge' not in pm.offsets
    src = packing.head_permutation()
    x = torch.randn(3, 256, 11, dtype=torch.float64)
    a_ref = torch.randn(3, 256, 11, dtype=torch.float64)          # attention output in the reference's channel order
    p = 'gnn.layers.1.'
    msg = torch.nn.functional.conv1d(a_ref, sd[p + 'attn.merge.weight'].double(), sd[p + 'attn.merge.bias'].double())
    ref = torch.nn.functional.conv1d(torch.cat([x, msg], 1), sd[p + 'mlp.0.weight'].double(), sd[p + 'mlp.0.bias'].double())
    ref = torch.nn.functional.batch_norm(ref, sd[p + 'mlp.1.running_mean'].double(), sd[p + 'mlp.1.running_var'].double(),
                                         sd[p + 'mlp.1.weight'].double(), sd[p + 'mlp.1.bias'].double(), False, 0.0, 1e-5)
    wf, _, _ = _get(pm, 'l1_w_mlp0', (512, 512))
    bf, _, _ = _get(pm, 'l1_b_mlp0', (512,))
    a_packed = a_ref[:, src]                                      # the kernels' head-contiguous channel order
    got = torch.einsum('oc,bcn->bon', wf.double(), torch.cat([x, a_packed], 1)) + bf.double()[None, :, None]
    assert (got - ref).abs().max() < 2e-5


def test_tf32_planes():
    _, pm = _packed()
    for name, shape in (('l2_w_qkv', (768, 256)), ('w_final', (256, 256)), ('conf_wf0', (512, 512))):
        raw, hi, lo = _get(pm, name, shape)
        assert ((hi.view(torch.int32) & 0x1FFF) == 0).all() and ((lo.view(torch.int32) & 0x1FFF) == 0).all()   # tf32-representable
        assert ((raw - hi).abs() <= raw.abs() * 2.0 ** -11 + 1e-45).all()                                    # round to nearest
        assert ((raw.double() - hi.double() - lo.double()).abs() <= raw.abs().double() * 2.0 ** -21 + 1e-45).all()
    assert pm.struct.lo_offset == 2 * pm.struct.hi_offset and pm.struct.n_layers == len(LAYERS)
