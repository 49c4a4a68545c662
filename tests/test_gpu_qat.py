"""GPU: quantisation-aware-training graphs (quantized=1 google QAT with shortcut_way=1, quantized=2 TPSQ; SURVEY 8a rows a22 /
a23) against fixtures produced by the reference itself (oracle/gen_golden_qat.py: ONE training step -- forward,
compute_loss, backward -- of yolov3 at 64x64, batch 2, then an eval forward).  The QAT layers run as module forwards
over NHWC tensors composed of the sm_100a kernels (b200yolo/qat.py)."""
import numpy as np
import pytest
import torch

from helpers import attach_hyp, cfg_path, golden, orc

pytestmark = pytest.mark.gpu

STEPS = 100


def _model(mode):
    import models
    fm = models.Darknet(cfg_path("yolov3"))
    sd = orc.synth_state_dict(fm.state_dict(), 0)
    qm = models.Darknet(cfg_path("yolov3"), quantized=mode, a_bit=8, w_bit=8, shortcut_way=1, steps=STEPS)
    with torch.no_grad():
        for i, m in enumerate(qm.module_list):
            pre = 'module_list.%d.' % i
            if m.__class__.__name__ == 'Sequential' and len(m) and hasattr(m[0], 'activation_quantizer'):
                c = m[0]
                c.weight.copy_(sd[pre + 'Conv2d.weight'])
                if (pre + 'BatchNorm2d.weight') in sd:
                    c.gamma.copy_(sd[pre + 'BatchNorm2d.weight'])
                    c.beta.copy_(sd[pre + 'BatchNorm2d.bias'])
                else:
                    c.bias.copy_(sd[pre + 'Conv2d.bias'])
    return attach_hyp(qm.cuda())


@pytest.mark.parametrize("mode,tag", [(1, "qat1"), (2, "qat2")])
def test_qat_training_step_and_eval_match_reference(mode, tag):
    from utils import utils as my_utils
    g = golden("yolov3_64_%s" % tag)
    qm = _model(mode).train()
    x = orc.synth_images(2, 64, 64, seed=0).cuda()
    t = orc.synth_targets(2, 6, 80, seed=1).cuda()
    pred, _ = qm(x)
    loss, items = my_utils.compute_loss(pred, t, qm)
    loss.backward()
    torch.cuda.synchronize()
    items_rel = float((np.abs(items.detach().cpu().numpy() - g["items"]) / np.abs(g["items"])).max())
    p_abs = max(float((pi.detach().cpu() - torch.from_numpy(g["p%d" % i])).abs().max()) for i, pi in enumerate(pred))
    p_rms = max(float((pi.detach().cpu() - torch.from_numpy(g["p%d" % i])).pow(2).mean().sqrt()) for i, pi in enumerate(pred))
    params = dict(qm.named_parameters())
    names = [str(n) for n in g["grad_names"]]
    ref = dict(zip(names, g["grad_norms"]))
    missing = [k for k in names if params[k].grad is None]
    floor = 1e-3 * float(np.median(list(ref.values())))
    rel = np.array([abs(float(params[k].grad.norm()) - v) / max(v, floor) for k, v in ref.items() if params[k].grad is not None])
    # quantiser scales after the step: powers of two -> exact equality expected (a fold computed from fp16-rounded
    # batch statistics may move a range across a power-of-two boundary in rare layers)
    sd = qm.state_dict()
    sn, sv = [str(n) for n in g["scale_names"]], g["scale_values"]
    diff = [(k, float(sd[k].reshape(-1)[0]), float(v)) for k, v in zip(sn, sv)
            if abs(float(sd[k].reshape(-1)[0]) - float(v)) > 1e-6 * abs(float(v))]
    stat = 0.0
    for k in g.files:
        if k.startswith("state::") and k.endswith(("running_mean", "running_var")):
            r = torch.from_numpy(g[k])
            stat = max(stat, float((sd[k[7:]].cpu() - r).abs().max() / r.abs().max().clamp(min=1e-6)))
    print("\n[%s train step vs reference] loss items rel %.3g | p abs %.3g rms %.3g | grad-norm rel median %.3g worst %.3g "
          "(%d params without grad) | scales differing %d of %d %s | running stats rel %.3g"
          % (tag, items_rel, p_abs, p_rms, np.median(rel), rel.max(), len(missing), len(diff), len(sn), diff[:3], stat))
    qm.eval()
    with torch.no_grad():
        io, p, _ = qm(orc.synth_images(2, 64, 64, seed=5).cuda())
    ref_io = torch.from_numpy(g["eval_io"])
    close = float(torch.isclose(io.cpu(), ref_io, rtol=2e-2, atol=2e-2).float().mean())
    print("[%s eval after the step] fraction of io within 2e-2: %.4f, max abs diff %.3g" % (tag, close, float((io.cpu() - ref_io).abs().max())))
    assert not missing
    assert items_rel < 5e-2
    assert np.median(rel) < 0.1
    assert len(diff) <= 0.05 * len(sn)
    assert stat < 2e-2
    assert close > 0.95
