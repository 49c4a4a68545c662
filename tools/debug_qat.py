import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
from helpers import *
sys.path.insert(0, os.path.join(ROOT, "tests"))
from test_gpu_qat import _model
mode = int(sys.argv[1])
g = np.load(os.path.join(ROOT, "tests", "golden", "_dbg_qat%d.npz" % mode))
qm = _model(mode).train()
x = orc.synth_images(2, 64, 64, seed=0).cuda()
outs = {}
def hook(i):
    def f(m, inp, out):
        outs[i] = out.detach()
    return f
for i, m in enumerate(qm.module_list):
    m.register_forward_hook(hook(i))
pred, _ = qm(x)
for k in sorted([k for k in g.files if k[1:].isdigit()], key=lambda s: int(s[1:])):
    i = int(k[1:])
    if i not in outs: continue
    ref = torch.from_numpy(g[k])
    mine = outs[i].permute(0, 3, 1, 2).cpu()
    print(k, tuple(ref.shape), "max abs diff %.4g  ref absmax %.4g  frac differing %.4g" % (float((mine - ref).abs().max()), float(ref.abs().max()), float(((mine - ref).abs() > 1e-6).float().mean())))
for i in (0, 1, 2, 3):
    c = qm.module_list[i][0]
    print(i, "batch_mean diff %.3g (max %.3g) batch_var diff %.3g (max %.3g)" % (float((c.batch_mean.cpu() - torch.from_numpy(g["L%d.bm" % i])).abs().max()), float(np.abs(g["L%d.bm" % i]).max()), float((c.batch_var.cpu() - torch.from_numpy(g["L%d.bv" % i])).abs().max()), float(np.abs(g["L%d.bv" % i]).max())),
          "scales w/a/b", float(c.weight_quantizer.scale), float(g["L%d.ws" % i]), float(c.activation_quantizer.scale), float(g["L%d.as" % i]), float(c.bias_quantizer.scale), float(g["L%d.bs" % i]))
