"""Debug driver: layer-by-layer comparison of the training-mode forward against the oracle."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import torch
from helpers import attach_hyp, build_model, module_defs, orc, cfg_path

name = sys.argv[1] if len(sys.argv) > 1 else "yolov3-tiny"
import models
model = attach_hyp(build_model(name, device="cuda")).train()
x = orc.synth_images(2, 64, 64, seed=0)
pred, feats = model(x.cuda())
torch.cuda.synchronize()
plan = model.engine().plan_for(x.cuda())
sd = orc.synth_state_dict(models.Darknet(cfg_path(name)).state_dict(), 0)
outs = orc.darknet_forward(module_defs(name), sd, x, name, training=True, return_layers=True)
from b200yolo import ops
defs = module_defs(name)
bad = 0
for r in plan.convs:
    i = r.i
    ref_i = i + 1 if r.res is not None else i
    ref = outs[ref_i]
    if r.head:
        got = r.y.buf[..., :r.Cout].permute(0, 3, 1, 2).float().cpu()
    else:
        got = ops.nhwc_to_nchw(r.y.view()).cpu()
    err = (got - ref).abs().max().item()
    scale = ref.abs().max().item()
    flag = "" if err < 0.02 * max(scale, 1) else "  <<<<<<"
    if flag:
        bad += 1
    if flag or i < 4:
        extra = ""
        if r.bn is not None:
            z = ops.nhwc_to_nchw(r.z.view()).cpu()
            extra = " z_absmax=%.3g mean0=%.4g invstd0=%.4g" % (z.abs().max(), float(r.mean[0]), float(r.invstd[0]))
        print("L%03d %s k%d s%d C%d->%d HxW=%dx%d err=%.4g refmax=%.4g%s%s" % (i, r.act, r.k, r.s, r.conv.in_channels,
              r.Cout, r.y.H, r.y.W, err, scale, extra, flag))
    if bad >= 6:
        break
print("done, bad layers:", bad)
