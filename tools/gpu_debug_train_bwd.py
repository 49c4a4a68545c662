"""Debug driver: per-layer activation-gradient comparison (engine vs oracle autograd)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import torch
from helpers import attach_hyp, build_model, module_defs, orc, cfg_path, anchor_vecs

name = sys.argv[1] if len(sys.argv) > 1 else "yolov3-tiny"
import models
from utils import utils as my_utils
model = attach_hyp(build_model(name, device="cuda")).train()
x = orc.synth_images(4, 128, 128, seed=0)
t = orc.synth_targets(4, 6, 80, seed=1)
from b200yolo import ops as _ops
_orig = _ops.conv2d_bwd_data
_log = []
def _wrapped(dy, w_packed_t, in_shape, k, stride, pad, out=None, accumulate=False, inv_scale=None):
    torch.cuda.synchronize()
    before = float(out.float().abs().mean())
    inv = None if inv_scale is None else float(inv_scale)
    dzmax = float(dy.float().abs().max())
    r = _orig(dy, w_packed_t, in_shape, k, stride, pad, out=out, accumulate=accumulate, inv_scale=inv_scale)
    torch.cuda.synchronize()
    _log.append((tuple(dy.shape), in_shape, k, stride, inv, dzmax, before, float(out.float().abs().mean()),
                 str(dy.dtype), str(out.dtype), inv_scale.data_ptr() if inv_scale is not None else 0))
    return r
_ops.conv2d_bwd_data = _wrapped
import b200yolo.train_engine as _te
pred, _ = model(x.cuda())
loss, items = my_utils.compute_loss(pred, t.cuda(), model)
loss.backward()
torch.cuda.synchronize()
plan = model.engine().last_plan
sd = orc.synth_state_dict(models.Darknet(cfg_path(name)).state_dict(), 0)
for k, v in sd.items():
    if v.dtype.is_floating_point and not k.endswith(('running_mean', 'running_var')):
        v.requires_grad_(True)
yolo_out, outs = orc.darknet_forward(module_defs(name), sd, x, name, training=True, return_layers='both')
for o in outs:
    if o.requires_grad:
        o.retain_grad()
p = [y for y in yolo_out]
rl, _ = orc.compute_loss(p, t, anchor_vecs(name), dict(orc.DEFAULT_HYP), 80, 1.0)
rl.backward()
from b200yolo import ops
print("dgrad calls (dy shape, in_shape, k, s, inv, max|dz|, mean|out| before, after, dtypes, inv ptr):")
for e in _log[:4]:
    print("  ", e)
print("aux rows (max|du|, s, 1/s):")
for r in plan.convs[-6:]:
    print("  L%03d" % r.i, plan.dz_aux[r.aux_row].tolist())
print("head scale", plan.head_scale.tolist())
for r in reversed(plan.convs):
    i = r.i
    ref_i = i + 1 if r.res is not None else i
    ref = outs[ref_i].grad
    if ref is None:
        continue
    g = plan.grad_of[id(r.y)]
    if r.head:
        got = g.buf[..., :r.Cout].float().permute(0, 3, 1, 2).cpu() * float(plan.head_scale[1])
    else:
        got = g.view().float().permute(0, 3, 1, 2).cpu()
    ratio = (got.abs().mean() / ref.abs().mean().clamp(min=1e-30)).item()
    err = ((got - ref).abs().max() / ref.abs().max().clamp(min=1e-30)).item()
    print("L%03d %s%s C%d @%d  mean|g| ours/ref = %.4g  max rel err = %.3g" % (i, r.act, "+res" if r.res is not None else "",
          r.Cout, r.y.H, ratio, err))

# ---- re-run L021's (last BN conv) backward standalone from the plan's own tensors ------------------------------
r = [c for c in plan.convs if c.bn is not None][-1]
print("re-running backward of L%03d standalone" % r.i)
dy = plan.grad_of[id(r.y)].view()
dz, dgam, dbet, aux = ops.bn_act_bwd(r.z.view(), dy, r.scale, r.shift, r.bn.weight.detach(), r.mean, r.invstd, r.act, r.slope)
print("  aux", aux.tolist(), "max|dz16|", float(dz.float().abs().max()))
src = r.src
fresh = torch.zeros((plan.B, src.H, src.W, src.C), dtype=torch.bfloat16, device="cuda")
ops.conv2d_bwd_data(dz, r.wT, (plan.B, src.H, src.W, r.conv.in_channels), r.k, r.s, r.p, out=fresh, accumulate=True,
                    inv_scale=aux[2:3])
torch.cuda.synchronize()
plan_g = plan.grad_of[id(src)].view().float()
print("  mean|fresh| %.4g   mean|plan grad of src| %.4g" % (float(fresh.float().abs().mean()), float(plan_g.abs().mean())))
# oracle: gradient w.r.t. the input of that conv
ref = outs[r.i - 1].grad
if ref is not None:
    print("  mean|oracle| %.4g" % float(ref.abs().mean()))
print("  dz_aux row in plan:", plan.dz_aux[r.aux_row].tolist())
print("  wT dtype", r.wT.dtype, "dz_scratch dtype", plan.dz_scratch.dtype, "src grad dtype", plan.grad_of[id(src)].buf.dtype)
