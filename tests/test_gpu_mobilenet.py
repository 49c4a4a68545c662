"""GPU: depthwise-separable / squeeze-excite path (BASELINE configs[3], SURVEY 8a rows a5 + a8): kernels against plain
fp32 PyTorch on identical fp16-rounded operands, and yolov3-mobilenet whole-model parity against fixtures produced by
the reference itself (oracle/gen_golden_mobilenet.py: eval forward and training step at 128x128)."""
import json

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from helpers import attach_hyp, build_model, golden, orc

pytestmark = pytest.mark.gpu

NAME = "yolov3-mobilenet"


def _nhwc(t, dtype=torch.float16):
    return t.permute(0, 2, 3, 1).contiguous().to(dtype).cuda()


def _nchw(t):
    return t.float().permute(0, 3, 1, 2).cpu()


@pytest.mark.parametrize("C,k,s,act", [(16, 3, 1, "relu6"), (72, 5, 2, "relu6"), (240, 3, 2, "h_swish"),
                                       (184, 3, 1, "h_swish"), (960, 5, 1, "h_swish"), (24, 3, 1, "linear")])
def test_depthwise_forward_backward(C, k, s, act):
    from b200yolo import ops
    g = torch.Generator().manual_seed(C * 10 + k)
    B, H, W = 3, 19, 22
    pad = (k - 1) // 2
    x = torch.randn(B, C, H, W, generator=g).half().float()
    w = (torch.randn(C, 1, k, k, generator=g) / k).float()
    w16 = w.half().float()
    scale = torch.rand(C, generator=g) + 0.5
    bias = torch.randn(C, generator=g) * 0.1
    # forward with folded BN + activation
    ref = orc.activation(F.conv2d(x.double(), w16.double(), None, s, pad, 1, C) * scale.double().view(1, -1, 1, 1)
                         + bias.double().view(1, -1, 1, 1), act)
    y = ops.dwconv2d(_nhwc(x), w.cuda(), k, s, pad, scale=scale.cuda(), bias=bias.cuda(), act=act)
    torch.cuda.synchronize()
    assert (_nchw(y) - ref.float()).abs().max() < 4e-3 * max(1.0, float(ref.abs().max()))
    # training forward: raw output + channel sums
    s1 = torch.zeros(C, device="cuda")
    s2 = torch.zeros(C, device="cuda")
    z = ops.dwconv2d(_nhwc(x), w.cuda(), k, s, pad, stats=(s1, s2))
    zr = F.conv2d(x.double(), w16.double(), None, s, pad, 1, C)
    zf = _nchw(z)
    assert (zf - zr.float()).abs().max() < 4e-3 * float(zr.abs().max())
    np.testing.assert_allclose(s1.cpu().numpy(), zf.sum(dim=(0, 2, 3)).numpy(), rtol=1e-4, atol=1e-3)
    np.testing.assert_allclose(s2.cpu().numpy(), (zf * zf).sum(dim=(0, 2, 3)).numpy(), rtol=1e-4, atol=1e-3)
    # backward: data and weight gradients (dz fp16 with a device-side 1/s factor), bf16 accumulation target
    Ho, Wo = zr.shape[2:]
    dz = torch.randn(B, C, Ho, Wo, generator=g).half().float()
    inv_s = torch.tensor([0.25], device="cuda")
    xr = x.clone().double().requires_grad_(True)
    wr = w16.clone().double().requires_grad_(True)
    F.conv2d(xr, wr, None, s, pad, 1, C).backward(dz.double() * 0.25)
    prev = torch.randn(B, C, H, W, generator=g).bfloat16()
    dx = ops.dwconv2d_bwd_data(_nhwc(dz), w.cuda(), (B, H, W, C), k, s, pad, out=_nhwc(prev.float(), torch.bfloat16),
                               accumulate=True, inv_scale=inv_s)
    dw = ops.dwconv2d_bwd_weight(_nhwc(x), _nhwc(dz), k, s, pad, alpha=2.0, inv_scale=inv_s)
    torch.cuda.synchronize()
    want_dx = prev.float() + xr.grad.float()
    assert (_nchw(dx) - want_dx).abs().max() < 1.6e-2 * max(1.0, float(want_dx.abs().max()))       # bf16 store
    np.testing.assert_allclose(dw.cpu().numpy(), 2.0 * wr.grad.float().numpy(), rtol=2e-3, atol=2e-3)


@pytest.mark.parametrize("C", [72, 120, 480, 960])
def test_squeeze_excite_forward_backward(C):
    from b200yolo import ops
    g = torch.Generator().manual_seed(C)
    B, H, W = 4, 10, 13
    cr = C // 4
    x = (torch.randn(B, C, H, W, generator=g) + 0.5).half().float()
    w1 = torch.randn(cr, C, generator=g) / C ** 0.5 * 3
    w2 = torch.randn(C, cr, generator=g) / cr ** 0.5 * 3
    dy = torch.randn(B, C, H, W, generator=g).bfloat16().float()
    xr = x.clone().double().requires_grad_(True)
    w1r, w2r = w1.clone().double().requires_grad_(True), w2.clone().double().requires_grad_(True)
    pooled = xr.mean(dim=(2, 3))
    gate = F.relu6(F.relu(pooled @ w1r.t()) @ w2r.t() + 3.0) / 6.0
    ref = xr * gate.view(B, C, 1, 1)
    ref.backward(dy.double())
    xn = _nhwc(x)
    y, ws = ops.se_fwd(xn, w1.cuda(), w2.cuda())
    prev = torch.randn(B, C, H, W, generator=g).bfloat16()
    dx, dw1, dw2 = ops.se_bwd(xn, _nhwc(dy, torch.bfloat16), w1.cuda(), w2.cuda(), ws,
                              _nhwc(prev.float(), torch.bfloat16), accumulate=True, grad_scale=0.5)
    torch.cuda.synchronize()
    assert (_nchw(y) - ref.detach().float()).abs().max() < 2e-3 * max(1.0, float(ref.detach().abs().max()))
    want = prev.float() + xr.grad.float()
    assert (_nchw(dx) - want).abs().max() < 1.6e-2 * max(1.0, float(want.abs().max()))
    np.testing.assert_allclose(dw1.cpu().numpy() * 2, w1r.grad.float().numpy(), rtol=2e-3, atol=2e-4 * float(w1r.grad.abs().max()))
    np.testing.assert_allclose(dw2.cpu().numpy() * 2, w2r.grad.float().numpy(), rtol=2e-3, atol=2e-4 * float(w2r.grad.abs().max()))


def _errs(got, ref):
    box_rel = ((got[..., :4] - ref[..., :4]).abs() / ref[..., :4].abs().clamp(min=1.0)).max().item()
    prob = (got[..., 4:] - ref[..., 4:]).abs().max().item()
    return box_rel, prob


def test_mobilenet_eval_forward_matches_reference():
    g = golden("%s_128_eval" % NAME)
    model = build_model(NAME, device="cuda").eval()
    x = orc.synth_images(2, 128, 128, seed=0)
    with torch.no_grad():
        io, p, _ = model(x.cuda())
        io2, _, _ = model(x.cuda())          # CUDA-graph replay
    torch.cuda.synchronize()
    assert torch.equal(io, io2)
    assert tuple(io.shape) == g["io"].shape
    b, pr = _errs(io.cpu(), torch.from_numpy(g["io"]))
    print("\n[%s 128x128 eval] vs fp32 reference: box_rel=%.3g prob_abs=%.3g" % (NAME, b, pr))
    assert b < 1e-2 and pr < 5e-3


def test_mobilenet_training_step_matches_reference():
    from utils import utils as my_utils
    g = golden("%s_128_train" % NAME)
    model = attach_hyp(build_model(NAME, device="cuda")).train()
    x = orc.synth_images(4, 128, 128, seed=0).cuda()
    t = orc.synth_targets(4, 6, 80, seed=1).cuda()
    losses = []
    for it in range(3):                         # eager, graph capture, replay
        model.zero_grad(set_to_none=True)
        pred, _ = model(x)
        loss, items = my_utils.compute_loss(pred, t, model)
        loss.backward()
        losses.append(float(loss.detach()))
        if it == 0:
            first_items = items.detach().cpu().numpy()
            grads = {k: p.grad.detach().float().cpu().clone() for k, p in model.named_parameters()}
            preds = [pi.detach().cpu() for pi in pred]
    torch.cuda.synchronize()
    p_abs = max(float((pi - torch.from_numpy(g["p%d" % i])).abs().max()) for i, pi in enumerate(preds))
    items_rel = float((np.abs(first_items - g["items"]) / np.abs(g["items"])).max())
    ref = dict(zip([str(n) for n in g["grad_names"]], g["grad_norms"]))
    # The BatchNorm bias of a linear projection conv that feeds only conv -> batch-stat BatchNorm has a mathematically
    # ZERO gradient (a per-channel shift is removed by the next normalisation); the reference shows ~1e-6 there and the
    # fp16-stored dz leaves ~1e-2 -- compared with gradient norms of 1..100 elsewhere.  Norm errors are therefore taken
    # relative to max(reference norm, 1e-3 x the median reference norm).
    floor = 1e-3 * float(np.median(list(ref.values())))
    rel = np.array([abs(float(grads[k].norm()) - v) / max(v, floor) for k, v in ref.items()])
    print("\n[%s 4x128x128 train] vs fp32 reference: max|dp| %.3g | loss items rel %.3g | grad-norm rel median %.3g worst "
          "%.3g | losses %s" % (NAME, p_abs, items_rel, np.median(rel), rel.max(), losses))
    assert all(np.isfinite(losses)) and max(abs(a - losses[0]) / abs(losses[0]) for a in losses) < 5e-3
    assert all(torch.isfinite(v).all() for v in grads.values())
    assert items_rel < 1.2e-2            # measured 5.2e-3
    assert np.median(rel) < 4.1e-2       # measured 2.0e-2
    assert p_abs < 0.5                   # measured 0.24
