"""QAT layer zoo (quantized=1) under the reference's class names (reference utils/quantized/quantized_google.py):
RangeTracker family :16-77, Round (STE) :81, Quantizer family :95-219, BNFold_QuantizedConv2d_For_FPGA :235-836,
QuantizedShortcut_max/min :839/1066, QuantizedFeatureConcat :1305.

Round 1 of the B200 build ships the device primitives these layers are made of -- b2y_minmax_f32 (range trackers),
b2y_fakequant_f32 (round-half-away + clamp + power-of-two scale), the BN fold, and the INT8 tcgen05 conv -- and keeps
the module classes (constructors, parameters, buffers, state_dict keys) so checkpoints load; the QAT *training*
forward (fake-quant with straight-through gradients) is scheduled after the PTQ path ("next").
"""
import torch
import torch.nn as nn
from torch.autograd import Function
from torch.nn import Parameter


class RangeTracker(nn.Module):
    def __init__(self, q_level):
        super().__init__()
        self.q_level = q_level

    def update_range(self, min_val, max_val):
        raise NotImplementedError

    @torch.no_grad()
    def forward(self, input):
        from b200yolo import ops
        if self.q_level == 'L':
            mm = ops.minmax(input.reshape(1, -1), per_row=False)
            min_val, max_val = mm[0, 0].reshape(1), mm[0, 1].reshape(1)
        else:  # 'C': per output channel (dim 0)
            mm = ops.minmax(input.reshape(input.shape[0], -1), per_row=True)
            min_val, max_val = mm[:, 0].reshape(-1, 1, 1, 1), mm[:, 1].reshape(-1, 1, 1, 1)
        self.update_range(min_val, max_val)


class GlobalRangeTracker(RangeTracker):
    def __init__(self, q_level, out_channels):
        super().__init__(q_level)
        shape = (1,) if q_level == 'L' else (out_channels, 1, 1, 1)
        self.register_buffer('min_val', torch.zeros(shape))
        self.register_buffer('max_val', torch.zeros(shape))
        self.register_buffer('first_w', torch.zeros(1))

    def update_range(self, min_val, max_val):
        if self.first_w == 0:
            self.first_w.add_(1)
            self.min_val.add_(min_val)
            self.max_val.add_(max_val)
        else:
            self.min_val.copy_(torch.min(min_val, self.min_val))
            self.max_val.copy_(torch.max(max_val, self.max_val))


class AveragedRangeTracker(RangeTracker):
    def __init__(self, q_level, out_channels, momentum=0.1):
        super().__init__(q_level)
        self.momentum = momentum
        shape = (1,) if q_level == 'L' else (out_channels, 1, 1, 1)
        self.register_buffer('min_val', torch.zeros(shape))
        self.register_buffer('max_val', torch.zeros(shape))
        self.register_buffer('first_a', torch.zeros(1))

    def update_range(self, min_val, max_val):
        if self.first_a == 0:
            self.first_a.add_(1)
            self.min_val.add_(min_val)
            self.max_val.add_(max_val)
        else:
            self.min_val.mul_(1 - self.momentum).add_(min_val * self.momentum)
            self.max_val.mul_(1 - self.momentum).add_(max_val * self.momentum)


class Round(Function):
    @staticmethod
    def forward(ctx, input):
        return torch.sign(input) * torch.floor(torch.abs(input) + 0.5)

    @staticmethod
    def backward(ctx, grad_output):
        return grad_output.clone()


class Quantizer(nn.Module):
    def __init__(self, bits, range_tracker, out_channels, FPGA, sign=True):
        super().__init__()
        self.bits, self.range_tracker, self.FPGA, self.sign = bits, range_tracker, FPGA, sign
        shape = (1,) if out_channels == -1 else (out_channels, 1, 1, 1)
        self.register_buffer('scale', torch.zeros(shape))
        self.register_buffer('zero_point', torch.zeros(shape))

    def forward(self, input):
        raise NotImplementedError("google-QAT fake-quant forward is scheduled after the PTQ path (DESIGN.md)")


class SymmetricQuantizer(Quantizer):
    pass


class AsymmetricQuantizer(Quantizer):
    pass


def _unsupported(name):
    class _Stub(nn.Module):
        def __init__(self, *a, **k):
            super().__init__()
            raise NotImplementedError("%s: the google-QAT graph is scheduled after the PTQ path (DESIGN.md)" % name)
    _Stub.__name__ = name
    return _Stub


BNFold_QuantizedConv2d_For_FPGA = _unsupported('BNFold_QuantizedConv2d_For_FPGA')
QuantizedShortcut_max = _unsupported('QuantizedShortcut_max')
QuantizedShortcut_min = _unsupported('QuantizedShortcut_min')
QuantizedFeatureConcat = _unsupported('QuantizedFeatureConcat')
