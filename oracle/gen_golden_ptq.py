"""Golden fixture for the INT8 / power-of-two PTQ path (config C4): run the REFERENCE's own PTQ flow
(PTQ.py:25-28, 76-102: Darknet(cfg, quantized=3), calibration forwards in .train(), then .eval() forward) on CPU and
save the calibrated quantiser state + outputs.  Run here only; output tests/golden/yolov3_64_ptq.npz."""
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

sys.path.insert(0, HERE)
import darknet_oracle as orc  # noqa: E402


def main():
    torch.manual_seed(0)
    cfg = "cfg/yolov3/yolov3.cfg"
    fm = ref_models.Darknet(cfg)
    sd = orc.synth_state_dict(fm.state_dict(), 0)
    qm = ref_models.Darknet(cfg, quantized=3, a_bit=8, w_bit=8, shortcut_way=1)
    # same mapping as load_darknet_weights(quant=True) (models.py:610-628): BN params live on the conv module
    qsd = qm.state_dict()
    new = {}
    for k, v in qsd.items():
        parts = k.split('.')
        i = parts[1]
        if k.endswith('Conv2d.weight'):
            new[k] = sd[k]
        elif k.endswith('Conv2d.gamma'):
            new[k] = sd.get('module_list.%s.BatchNorm2d.weight' % i, torch.ones_like(v))
        elif k.endswith('Conv2d.beta'):
            new[k] = sd.get('module_list.%s.BatchNorm2d.bias' % i, torch.zeros_like(v))
        elif k.endswith('Conv2d.running_mean'):
            new[k] = sd.get('module_list.%s.BatchNorm2d.running_mean' % i, torch.zeros_like(v))
        elif k.endswith('Conv2d.running_var'):
            new[k] = sd.get('module_list.%s.BatchNorm2d.running_var' % i, torch.ones_like(v))
        elif k.endswith('Conv2d.bias') and ('module_list.%s.Conv2d.bias' % i) in sd:
            new[k] = sd['module_list.%s.Conv2d.bias' % i]
        else:
            new[k] = v
    qm.load_state_dict(new)
    B, S = 2, 64
    qm.train()
    with torch.no_grad():
        for seed in (10, 11):
            qm(orc.synth_images(B, S, S, seed=seed))
    qm.eval()
    x = orc.synth_images(B, S, S, seed=0)
    with torch.no_grad():
        io, p, _ = qm(x)
    out = {"io": io.numpy()}
    for i, pi in enumerate(p):
        out["p%d" % i] = pi.numpy()
    for i, m in enumerate(qm.module_list):
        name = m.__class__.__name__
        if name == 'Sequential' and len(m) and hasattr(m[0], 'activation_quantizer'):
            c = m[0]
            out["L%d.a_scale" % i] = c.activation_quantizer.scale.numpy()
            out["L%d.w_scale" % i] = c.weight_quantizer.scale.numpy()
            out["L%d.b_scale" % i] = c.bias_quantizer.scale.numpy()
            out["L%d.q_bias" % i] = c.q_bias.numpy()
            out["L%d.bias" % i] = c.bias.detach().numpy()
        elif name.startswith('COSPTQuantizedShortcut'):
            out["L%d.scale_x" % i] = m.scale_x.numpy()
            out["L%d.scale_a" % i] = m.scale_a.numpy()
            out["L%d.scale_sum" % i] = m.scale_sum.numpy()
        elif name == 'COSPTQuantizedFeatureConcat':
            for k, v in m.state_dict().items():
                out["L%d.%s" % (i, k)] = v.numpy()
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "yolov3_64_ptq.npz"), **out)
    print("saved", len(out), "entries; io mean", float(io.abs().mean()))
    for k in list(out)[:40]:
        if 'scale' in k:
            print(k, out[k].reshape(-1)[:3])


if __name__ == "__main__":
    main()
