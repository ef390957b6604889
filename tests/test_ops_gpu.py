"""GPU: torch.ops.vaegan_b200.* (the torch.library layer over the C-ABI) against F.conv2d / F.conv_transpose2d with
autograd, fp32 mode at 1e-4 and bf16 mode at bf16 tolerance; torch.library.opcheck on the schemas / fake kernels."""
import pytest
import torch
import torch.nn.functional as F

import vaegan_b200  # noqa: F401
from tests.util import rel_err

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("dtype,tol", [(torch.float32, 1e-4), (torch.bfloat16, 2e-2)])
@pytest.mark.parametrize("transposed", [False, True])
def test_custom_op_matches_torch_with_autograd(dtype, tol, transposed):
    gen = torch.Generator().manual_seed(3)
    B, cin, cout, hw = 4, 64, 128, 16
    x = torch.randn(B, cin, hw, hw, generator=gen)
    w = 0.05 * (torch.randn(cin, cout, 4, 4, generator=gen) if transposed else torch.randn(cout, cin, 4, 4, generator=gen))
    if dtype == torch.bfloat16:                     # identical operand rounding on both sides
        x, w = x.bfloat16().float(), w.bfloat16().float()
    xr, wr = x.clone().requires_grad_(True), w.clone().requires_grad_(True)
    y_ref = F.conv_transpose2d(xr, wr, None, 2, 1) if transposed else F.conv2d(xr, wr, None, 2, 1)
    gy = torch.randn(y_ref.shape, generator=gen)
    if dtype == torch.bfloat16:
        gy = gy.bfloat16().float()
    y_ref.backward(gy)

    xg = x.permute(0, 2, 3, 1).contiguous().to(dtype).cuda().requires_grad_(True)
    wg = w.cuda().requires_grad_(True)
    ns = torch.ops.vaegan_b200
    y = ns.conv_transpose2d_nhwc(xg, wg, 2, 1) if transposed else ns.conv2d_nhwc(xg, wg, None, 2, 1)
    y.backward(gy.permute(0, 2, 3, 1).contiguous().to(dtype).cuda())
    assert rel_err(y.float().permute(0, 3, 1, 2), y_ref) < tol
    assert rel_err(xg.grad.float().permute(0, 3, 1, 2), xr.grad) < tol
    assert rel_err(wg.grad, wr.grad) < tol


def test_opcheck():
    x = torch.randn(2, 8, 8, 32, device="cuda").bfloat16()
    w = torch.randn(64, 32, 4, 4, device="cuda") * 0.05
    torch.library.opcheck(torch.ops.vaegan_b200.conv2d_nhwc.default, (x, w, None, 2, 1),
                          test_utils=("test_schema", "test_faketensor"))
    wt = torch.randn(32, 64, 4, 4, device="cuda") * 0.05
    torch.library.opcheck(torch.ops.vaegan_b200.conv_transpose2d_nhwc.default, (x, wt, 2, 1),
                          test_utils=("test_schema", "test_faketensor"))
