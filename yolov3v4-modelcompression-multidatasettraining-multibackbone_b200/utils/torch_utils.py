"""Host-side torch helpers under the reference's names (reference utils/torch_utils.py):
init_seeds :7, select_device :16, time_synchronized :43, fuse_conv_and_bn :65-89, model_info :92,
scale_img :130, ModelEMA :141-189.  None of this is on the device hot path; it is small host glue that the
reference's scripts import by name."""
import math
import os
import time
from copy import deepcopy

import torch
import torch.backends.cudnn as cudnn
import torch.nn as nn
import torch.nn.functional as F

_PARALLEL_WRAPPERS = (nn.parallel.DataParallel, nn.parallel.DistributedDataParallel)


def init_seeds(seed=0):
    torch.manual_seed(seed)
    if seed == 0:  # reproducible library kernels for whatever still runs through cuDNN outside the hot path
        cudnn.deterministic, cudnn.benchmark = True, False


def select_device(device='', batch_size=None):
    """'cpu' | '' | '0' | '0,1,2,3' -> torch.device (sets CUDA_VISIBLE_DEVICES like the reference does)."""
    want_cpu = device.lower() == 'cpu'
    if device and not want_cpu:
        os.environ['CUDA_VISIBLE_DEVICES'] = device
        assert torch.cuda.is_available(), 'CUDA unavailable, invalid device %s requested' % device
    if want_cpu or not torch.cuda.is_available():
        print('Using CPU\n')
        return torch.device('cpu')
    count = torch.cuda.device_count()
    if count > 1 and batch_size:
        assert batch_size % count == 0, 'batch-size %g not multiple of GPU count %g' % (batch_size, count)
    for idx in range(count):
        props = torch.cuda.get_device_properties(idx)
        lead = 'Using CUDA ' if idx == 0 else ' ' * 11
        print("%sdevice%g _CudaDeviceProperties(name='%s', total_memory=%dMB)" %
              (lead, idx, props.name, props.total_memory / 2 ** 20))
    print('')
    return torch.device('cuda:0')


def time_synchronized():
    if torch.cuda.is_available():
        torch.cuda.synchronize()
    return time.time()


def find_modules(model, mclass=nn.Conv2d):
    return [i for i, m in enumerate(model.module_list) if isinstance(m, mclass)]


def fuse_conv_and_bn(conv, bn):
    """Inference BN fold into a biased Conv2d:
         w_f = diag(gamma / sqrt(eps + var)) . w ,  b_f = W_bn . b_conv + beta - gamma * mean / sqrt(var + eps)
    (the device-side twin is b2y_pack_conv_weights in csrc/pointwise.cu)."""
    dev = conv.weight.device
    out_c = conv.out_channels
    with torch.no_grad():
        folded = nn.Conv2d(conv.in_channels, out_c, kernel_size=conv.kernel_size, stride=conv.stride,
                           padding=conv.padding, groups=conv.groups, bias=True).to(dev)
        w_bn = torch.diag(bn.weight.div(torch.sqrt(bn.eps + bn.running_var)))
        folded.weight.copy_(torch.mm(w_bn, conv.weight.clone().view(out_c, -1)).view(folded.weight.shape))
        conv_b = conv.bias if conv.bias is not None else torch.zeros(out_c, device=dev)
        shift = bn.bias - bn.weight.mul(bn.running_mean).div(torch.sqrt(bn.running_var + bn.eps))
        folded.bias.copy_(torch.mm(w_bn, conv_b.reshape(-1, 1)).reshape(-1) + shift)
    return folded


def model_info(model, verbose=False):
    params = list(model.parameters())
    total = sum(p.numel() for p in params)
    trainable = sum(p.numel() for p in params if p.requires_grad)
    if verbose:
        print('%5s %40s %9s %12s %20s %10s %10s' % ('layer', 'name', 'gradient', 'parameters', 'shape', 'mu', 'sigma'))
        for i, (name, p) in enumerate(model.named_parameters()):
            print('%5g %40s %9s %12g %20s %10.3g %10.3g' % (i, name.replace('module_list.', ''), p.requires_grad,
                                                           p.numel(), list(p.shape), p.mean(), p.std()))
    print('Model Summary: %g layers, %g parameters, %g gradients' % (len(params), total, trainable))


def scale_img(img, ratio=1.0, same_shape=True):
    """Bilinear rescale of a [B,3,H,W] batch by `ratio`, padded (value 0.447) back to the original size, or to the
    next multiple of 64 when same_shape is False (test-time augmentation, models.py:486-489)."""
    height, width = img.shape[2:]
    new_h, new_w = int(height * ratio), int(width * ratio)
    out = F.interpolate(img, size=(new_h, new_w), mode='bilinear', align_corners=False)
    if not same_shape:
        height, width = (math.ceil(v * ratio / 64) * 64 for v in (height, width))
    return F.pad(out, [0, width - new_w, 0, height - new_h], value=0.447)


class ModelEMA:
    """EMA over every floating-point entry of the state_dict, decay ramp d(t) = decay * (1 - exp(-t/2000))."""

    def __init__(self, model, decay=0.9999, device=''):
        self.ema = deepcopy(model).eval()
        self.updates = 0
        self.decay = lambda step: decay * (1 - math.exp(-step / 2000))
        self.device = device
        if device:
            self.ema.to(device=device)
        for p in self.ema.parameters():
            p.requires_grad_(False)

    @staticmethod
    def _state(m, unwrap):
        return (m.module if unwrap else m).state_dict()

    def update(self, model):
        self.updates += 1
        d = self.decay(self.updates)
        wrapped = type(model) in _PARALLEL_WRAPPERS
        with torch.no_grad():
            src = self._state(model, wrapped)
            for key, avg in self._state(self.ema, wrapped).items():
                if avg.dtype.is_floating_point:
                    avg.mul_(d).add_(src[key].detach(), alpha=1. - d)

    def update_attr(self, model):
        for key, val in model.__dict__.items():
            if not key.startswith('_'):
                setattr(self.ema, key, val)
