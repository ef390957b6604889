"""CPU: the torch.library layer (vaegan_b200/ops.py) registers the contractions as real custom ops with shape
functions; there is no CPU kernel behind them - a CPU tensor must fail loudly in the dispatcher."""
import pytest
import torch

import vaegan_b200  # noqa: F401


def test_ops_registered_with_schemas():
    ns = torch.ops.vaegan_b200
    for name in ("conv2d_nhwc", "conv_transpose2d_nhwc", "conv_dgrad_nhwc", "conv_wgrad_nhwc"):
        assert hasattr(ns, name), name
    assert "Tensor? bias" in str(ns.conv2d_nhwc.default._schema)


def test_fake_shapes_follow_the_reference_layers():
    """Meta tensors through the registered shape functions: Conv2d(64,128,4,2,1) of gan_code.py:64 and
    ConvTranspose2d(128,64,4,2,1) of gan_code.py:37 on NHWC activations."""
    ns = torch.ops.vaegan_b200
    x = torch.empty((8, 32, 32, 64), dtype=torch.bfloat16, device="meta")
    w = torch.empty((128, 64, 4, 4), device="meta")
    y = ns.conv2d_nhwc(x, w, None, 2, 1)
    assert tuple(y.shape) == (8, 16, 16, 128) and y.dtype == torch.bfloat16
    assert tuple(ns.conv_dgrad_nhwc(y, w, False, 2, 1, 32, 32).shape) == (8, 32, 32, 64)
    dw = ns.conv_wgrad_nhwc(y, x, False, 4, 2, 1)
    assert tuple(dw.shape) == (128, 64, 4, 4) and dw.dtype == torch.float32
    wt = torch.empty((128, 64, 4, 4), device="meta")            # ConvTranspose2d.weight[Cin, Cout, k, k]
    xt = torch.empty((8, 16, 16, 128), dtype=torch.bfloat16, device="meta")
    yt = ns.conv_transpose2d_nhwc(xt, wt, 2, 1)
    assert tuple(yt.shape) == (8, 32, 32, 64)
    assert tuple(ns.conv_dgrad_nhwc(yt, wt, True, 2, 1, 16, 16).shape) == (8, 16, 16, 128)
    assert tuple(ns.conv_wgrad_nhwc(yt, xt, True, 4, 2, 1).shape) == (128, 64, 4, 4)


def test_no_cpu_fallback():
    x, w = torch.zeros((1, 8, 8, 16)), torch.zeros((16, 16, 4, 4))
    with pytest.raises((NotImplementedError, RuntimeError)):
        torch.ops.vaegan_b200.conv2d_nhwc(x, w, None, 2, 1)
