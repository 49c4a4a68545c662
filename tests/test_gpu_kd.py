"""GPU: knowledge-distillation losses of the YOLO head (utils/utils.py:435-520 compute_lost_KD /
KD2 / KD3) -- csrc/kd.cu through the C ABI and the reference's entry points in utils.utils -- against the reference's
own values and autograd gradients (tests/golden/kd_case.npz, oracle/gen_golden_kd.py; duplicate labels make several
matches share a cell).  Tolerances: loss 1e-5 relative, gradients 1e-4 of the layer's largest gradient (fp32 exp /
softmax in a different association than ATen's; measured values are printed)."""
from types import SimpleNamespace

import numpy as np
import pytest
import torch

from helpers import golden, orc

pytestmark = pytest.mark.gpu


def _setup():
    g = golden("kd_case")
    layers = []
    for i in range(3):
        av = torch.from_numpy(g["anchor_vec%d" % i])
        layers.append(SimpleNamespace(anchor_vec=av, anchor_wh=av.view(1, 3, 1, 1, 2)))
    model = SimpleNamespace(module_list=layers, yolo_layers=[0, 1, 2], hyp=dict(orc.DEFAULT_HYP), nc=80, gr=1.0)
    stu = [torch.from_numpy(g["s%d" % i]).cuda().requires_grad_(True) for i in range(3)]
    tea = [torch.from_numpy(g["t%d" % i]).cuda() for i in range(3)]
    return g, model, stu, tea, torch.from_numpy(g["targets"]).cuda()


def _check(tag, g, loss, stu):
    ref = g[tag + "_loss"].reshape(-1)
    got = loss.detach().cpu().numpy().reshape(-1)
    lrel = float(np.abs(got - ref).max() / np.abs(ref).max())
    worst = 0.0
    for i in range(3):
        r = g["%s_ds%d" % (tag, i)]
        d = stu[i].grad.cpu().numpy()
        worst = max(worst, float(np.abs(d - r).max() / max(np.abs(r).max(), 1e-30)))
    print("\n[%s] loss %.8g (reference %.8g, rel %.2g) | worst gradient error / layer max %.2g"
          % (tag, got[0], ref[0], lrel, worst))
    assert lrel < 1e-5 and worst < 1e-4


def test_kd1_soft_targets():
    from utils import utils as U
    g, model, stu, tea, t = _setup()
    loss = U.compute_lost_KD(stu, tea, 80, 3)
    assert loss.dim() == 0                                  # the reference returns a 0-dim tensor here
    loss.backward()
    _check("kd1", g, loss, stu)


def test_kd2_soft_targets_and_bounded_regression():
    from utils import utils as U
    g, model, stu, tea, t = _setup()
    loss, ratio = U.compute_lost_KD2(model, t, stu, tea)
    assert loss.shape == (1,) and isinstance(ratio, float)
    assert ratio == float(g["kd2_ratio"]), (ratio, float(g["kd2_ratio"]))
    loss.sum().backward()
    _check("kd2", g, loss, stu)


def test_kd3_soft_targets_and_teacher_boxes():
    from utils import utils as U
    g, model, stu, tea, t = _setup()
    loss = U.compute_lost_KD3(model, t, stu, tea)
    assert loss.shape == (1,)
    (2.0 * loss.sum()).backward()                          # upstream gradient != 1
    for s in stu:
        s.grad.mul_(0.5)
    _check("kd3", g, loss, stu)


def test_kd_cpu_tensors_fail_loudly():
    from utils import utils as U
    x = [torch.zeros(1, 3, 2, 2, 85)]
    with pytest.raises(RuntimeError):
        U.compute_lost_KD(x, x, 80, 1)
