import sys, os, math
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import conftest  # noqa
import torch
import torch.nn.functional as F
from b200yolo import ops


def case(B, H, W, Cin, Cout, k, stride, sc, extra_pitch=0):
    g = torch.Generator().manual_seed(1)
    pad = (k - 1) // 2
    w = (torch.randn(Cout, Cin, k, k, generator=g) / math.sqrt(Cin * k * k)).half()
    Ho, Wo = ops.conv_out_hw(H, W, k, stride, pad)
    dz_true = torch.randn(B, Cout, Ho, Wo, generator=g) * 1e-3
    dz16 = (dz_true * sc).half()
    x = torch.zeros(B, Cin, H, W, dtype=torch.double, requires_grad=True)
    F.conv2d(x, w.double(), None, stride=stride, padding=pad).backward(dz16.double() / sc)
    ref = x.grad.float()
    wt = ops.pack_dgrad_weights(w.float().cuda(), stride, pad, (H, W))
    buf = torch.zeros(B, H, W, Cin + extra_pitch, dtype=torch.bfloat16, device="cuda")
    out = buf[..., :Cin]
    aux = torch.zeros(8, 4, device="cuda")
    aux[3, 2] = 1.0 / sc
    inv = aux[3][2:3]
    ops.conv2d_bwd_data(dz16.permute(0, 2, 3, 1).contiguous().cuda(), wt, (B, H, W, Cin), k, stride, pad, out=out,
                        accumulate=True, inv_scale=inv)
    torch.cuda.synchronize()
    got = out.float().permute(0, 3, 1, 2).cpu()
    ratio = (got.abs().mean() / ref.abs().mean()).item()
    print("B%d %dx%d Cin%d Cout%d k%d s%d sc=%g pitch+%d -> mean ratio %.4g  max err/max %.3g" % (
        B, H, W, Cin, Cout, k, stride, sc, extra_pitch, ratio, ((got - ref).abs().max() / ref.abs().max()).item()))


case(4, 8, 8, 384, 256, 3, 1, 8192.0)
case(4, 8, 8, 384, 256, 3, 1, 8192.0, extra_pitch=0)
case(4, 4, 4, 256, 512, 3, 1, 4096.0)
case(4, 8, 8, 128, 256, 3, 1, 4096.0, extra_pitch=256)
case(4, 8, 8, 64, 128, 3, 2, 64.0)
case(4, 4, 4, 1024, 256, 1, 1, 4096.0)
