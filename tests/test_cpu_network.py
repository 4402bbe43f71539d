"""Host logic of the U-Net backbone (nksr_b200/unet.py) that needs no GPU: the torch reference convolution against
explicit loops, the octant table, the key-matched restriction, the parameter layout."""
from types import SimpleNamespace

import numpy as np
import torch

from nksr_b200.unet import SparseUNet, gather_gemm, octant_of_children, restrict_to


def test_torch_gather_gemm_is_the_definition():
    g = torch.Generator().manual_seed(0)
    n_in, n_out, K, ci, co = 11, 9, 5, 4, 3
    x = torch.randn((n_in, ci), generator=g)
    w = torch.randn((K, ci, co), generator=g)
    b = torch.randn(co, generator=g)
    res = torch.randn((n_out, co), generator=g)
    idx = torch.randint(-1, n_in, (n_out, K), generator=g, dtype=torch.int32)
    ref = np.zeros((n_out, co))
    for i in range(n_out):
        acc = b.numpy().astype(np.float64) + res[i].numpy()
        for k in range(K):
            j = int(idx[i, k])
            if j >= 0:
                acc = acc + x[j].numpy().astype(np.float64) @ w[k].numpy().astype(np.float64)
        ref[i] = np.maximum(acc, 0.0)
    out = gather_gemm(x, idx, w, b, res, relu=True, impl="torch")
    assert np.allclose(out.numpy(), ref, atol=1e-5)


def test_octants_and_restriction():
    child8 = torch.tensor([[0, -1, 2, -1, -1, -1, -1, 1], [-1, 3, -1, -1, 4, -1, -1, -1]], dtype=torch.int32)
    assert octant_of_children(child8, 6).tolist() == [0, 7, 2, 1, 4, -1]
    src = SimpleNamespace(keys=[torch.tensor([2, 5, 9, 11])])
    dst = SimpleNamespace(keys=[torch.tensor([5, 6, 11, 40])])
    f = {0: torch.arange(8.0).reshape(4, 2)}
    out = restrict_to(f, src, dst)[0]
    assert out.tolist() == [[2.0, 3.0], [0.0, 0.0], [6.0, 7.0], [0.0, 0.0]]
    assert restrict_to(f, src, src) is f


def test_unet_parameter_layout_follows_the_hparams():
    """configs/default/train.yaml:9-18: kernel_dim 4, tree_depth 4, unet.f_maps 32"""
    from nksr_b200.network import NKSRNetwork
    net = NKSRNetwork(dict(backbone="unet", tree_depth=4, kernel_dim=4, unet=dict(f_maps=32)))
    u = net.backbone_net
    assert isinstance(u, SparseUNet) and u.channels == [32, 64, 128, 256]
    assert tuple(u.enc_a[0].weight.shape) == (27, 32, 32) and tuple(u.down[0].weight.shape) == (8, 32, 64)
    assert tuple(u.dec[1].weight.shape) == (27, 128, 64) and u.heads[2].out_features == 6 + 2 * 4
    # seeded: two instances carry the same weights; the stand-in's parameters do not depend on the backbone
    net2 = NKSRNetwork(dict(backbone="unet", tree_depth=4, kernel_dim=4))
    assert all(torch.equal(a, b) for a, b in zip(net.state_dict().values(), net2.state_dict().values()))
    pool = NKSRNetwork(dict(tree_depth=4, kernel_dim=4))
    for k, v in pool.state_dict().items():
        assert torch.equal(v, net.state_dict()[k])
    net2.load_state_dict(net.state_dict())
