"""Golden fixtures for the QAT graphs: run the REFERENCE's own quantized=1 (google, shortcut_way=1) and quantized=2
(TPSQ) models on CPU for ONE training step (forward, compute_loss, backward) and an eval forward afterwards.
Output: tests/golden/yolov3_64_qat{1,2}.npz (+ _layers), tests/golden/yolov3-tiny_64_qat{1,2}.npz (the max-pool /
zero-pad layers of the tiny backbone inside the QAT graphs).   Run here only."""
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
REF = os.environ.get("B2Y_REFERENCE_ROOT", "/root/reference")
sys.dont_write_bytecode = True
sys.path.insert(0, REF)
sys.path.insert(0, os.path.join(HERE, "mpl_stub"))
os.chdir(REF)

import numpy as np  # noqa: E402
import torch  # noqa: E402
import models as ref_models  # noqa: E402
from utils import utils as ru  # noqa: E402

sys.path.insert(0, HERE)
import darknet_oracle as orc  # noqa: E402

STEPS = 100


def load_synth(qm, cfg):
    fm = ref_models.Darknet(cfg)
    sd = orc.synth_state_dict(fm.state_dict(), 0)
    new = {}
    for k, v in qm.state_dict().items():
        i = k.split('.')[1]
        if k.endswith('Conv2d.weight'):
            new[k] = sd[k]
        elif k.endswith('Conv2d.gamma'):
            new[k] = sd.get('module_list.%s.BatchNorm2d.weight' % i, torch.ones_like(v))
        elif k.endswith('Conv2d.beta'):
            new[k] = sd.get('module_list.%s.BatchNorm2d.bias' % i, torch.zeros_like(v))
        elif k.endswith('Conv2d.bias') and ('module_list.%s.Conv2d.bias' % i) in sd:
            new[k] = sd['module_list.%s.Conv2d.bias' % i]
        else:
            new[k] = v                      # running stats start at zero (first-batch initialisation), quantiser state
    qm.load_state_dict(new)


CFGS = {"yolov3": "cfg/yolov3/yolov3.cfg", "yolov3-tiny": "cfg/yolov3tiny/yolov3-tiny.cfg"}


def run(mode, tag, name="yolov3"):
    torch.manual_seed(0)
    cfg = CFGS[name]
    qm = ref_models.Darknet(cfg, quantized=mode, a_bit=8, w_bit=8, shortcut_way=1, steps=STEPS)
    load_synth(qm, cfg)
    qm.nc, qm.gr, qm.hyp = 80, 1.0, dict(orc.DEFAULT_HYP)
    B, S = 2, 64
    qm.train()
    x = orc.synth_images(B, S, S, seed=0)
    t = orc.synth_targets(B, 6, 80, seed=1)
    # module-boundary tensors of a few QAT conv layers (teacher forcing in tests/test_gpu_qat.py): input, output,
    # gradient w.r.t. the output and w.r.t. the input
    lay = {}
    LAYERS = (5, 13, 39, 63, 81) if name == "yolov3" else ()

    def fwd_hook(i):
        def f(m, inp, out):
            lay["L%d.x" % i] = inp[0].detach().numpy().copy()
            lay["L%d.y" % i] = out.detach().numpy().copy()
        return f

    def bwd_hook(i):
        def f(m, gin, gout):
            if gin[0] is not None:
                lay["L%d.gx" % i] = gin[0].detach().numpy().copy()
            lay["L%d.gy" % i] = gout[0].detach().numpy().copy()
        return f

    for i in LAYERS:
        qm.module_list[i][0].register_forward_hook(fwd_hook(i))
        qm.module_list[i][0].register_full_backward_hook(bwd_hook(i))
    pred, _ = qm(x)
    loss, items = ru.compute_loss(pred, t, qm)
    loss.backward()
    for i in LAYERS:
        c = qm.module_list[i][0]
        for n, p_ in c.named_parameters():
            if p_.grad is not None:
                g_ = p_.grad
                lay["L%d.gnorm.%s" % (i, n)] = np.float64(float(g_.norm()))
                lay["L%d.grad.%s" % (i, n)] = (g_ if g_.numel() <= 300000 else g_[:8]).numpy().copy()
        for n, b_ in c.named_buffers():
            if n.endswith(("scale", "running_mean", "running_var", "min_val", "max_val")) and b_.numel() <= 4096:
                lay["L%d.buf.%s" % (i, n)] = b_.detach().numpy().copy()
        for n, p_ in c.named_parameters():
            if n.endswith("scale"):
                lay["L%d.par.%s" % (i, n)] = p_.detach().numpy().copy()
    if LAYERS:
        np.savez_compressed(os.path.join(ROOT, "tests", "golden", "%s_64_%s_layers.npz" % (name, tag)), **lay)
    out = {"items": items.numpy(), "loss": loss.detach().numpy()}
    for i, pi in enumerate(pred):
        out["p%d" % i] = pi.detach().numpy()
    named = {k: p for k, p in qm.named_parameters() if p.grad is not None}
    out["grad_names"] = np.array(list(named))
    out["grad_norms"] = np.array([float(p.grad.norm()) for p in named.values()], dtype=np.float64)
    keep = [k for k in named if named[k].numel() <= 1024 and k.split('.')[1] in ('0', '1', '2', '79', '80', '81')]
    if name != "yolov3":
        keep = [k for k in named if named[k].numel() <= 4096]
    for k in keep:
        out["grad::" + k] = named[k].grad.numpy()
    sd = qm.state_dict()
    for k, v in sd.items():
        if k.endswith(('.scale', 'running_mean', 'running_var', 'min_val', 'max_val', 'input_scale')) and v.numel() <= 1024:
            if k.split('.')[1] in ('0', '1', '2', '3', '4', '79', '80', '81', '86'):
                out["state::" + k] = v.detach().numpy()
    all_scales = {k: float(v.detach().reshape(-1)[0]) for k, v in sd.items() if k.endswith('.scale') and v.numel() == 1}
    out["scale_names"] = np.array(list(all_scales))
    out["scale_values"] = np.array(list(all_scales.values()), dtype=np.float64)
    qm.eval()
    with torch.no_grad():
        io, p, _ = qm(orc.synth_images(B, S, S, seed=5))
    out["eval_io"] = io.numpy()
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "%s_64_%s.npz" % (name, tag)), **out)
    print(name, tag, "loss", float(loss), items.tolist(), "n scales", len(all_scales), "eval io mean", float(io.abs().mean()))


if __name__ == "__main__":
    torch.set_num_threads(8)
    names = [a for a in sys.argv[1:] if a in CFGS] or list(CFGS)
    tags = [a for a in sys.argv[1:] if a.startswith("qat")] or ["qat1", "qat2"]
    for name in names:
        for mode, tag in ((1, "qat1"), (2, "qat2")):
            if tag in tags:
                run(mode, tag, name)
