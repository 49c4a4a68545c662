"""TPSQ layer zoo (quantized=2, reference utils/quantized/quantized_TPSQ.py): learned power-of-two scale.
Scheduled after the PTQ path (DESIGN.md "next"); the class name is kept so that models.create_modules imports."""
import torch.nn as nn


class TPSQ_BNFold_QuantizedConv2d_For_FPGA(nn.Module):
    def __init__(self, *a, **k):
        super().__init__()
        raise NotImplementedError("TPSQ (quantized=2) is scheduled after the PTQ path (DESIGN.md)")
