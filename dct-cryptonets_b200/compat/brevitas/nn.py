"""QuantConv2d / QuantReLU / QuantIdentity with Brevitas' per-tensor float-scale semantics, restated (no Brevitas here):

* weights (Int8WeightPerTensorFloat, `weight_bit_width=b`, `narrow_range=True`): scale = max|W| / (2^(b-1) - 1) from the
  tensor's own statistics, integers clamp(round(W / scale)) in [-(2^(b-1)-1), 2^(b-1)-1];
* activations (Int8ActPerTensorFloat / the QuantReLU default Uint8ActPerTensorFloat, `bit_width=b`): a learned threshold
  parameter (initial value 1.0, which is also what an untrained Brevitas module uses in eval mode);
  signed: scale = threshold / 2^(b-1), integers in [-2^(b-1), 2^(b-1)-1]; unsigned: scale = threshold / (2^b - 1),
  integers in [0, 2^b - 1]; rounding half to even.

The threshold lives at `<layer>.act_quant.fused_activation_quant_proxy.tensor_quant.scaling_impl.value`, the path Brevitas
uses in its state dict (from memory), so that a QAT checkpoint (reference train.py:83-89) has a chance to load.
`tfx_act_quant()` / `tfx_weight_quant()` expose the quantiser to the circuit front-end (tfx_b200/circuit.py).
"""
import torch
import torch.nn as nn
import torch.nn.functional as F

from .quant import Int8ActPerTensorFloat, Int8WeightPerTensorFloat, Uint8ActPerTensorFloat


class _Holder(nn.Module):
    pass


def _threshold_tree(init: float) -> nn.Module:
    scaling = _Holder()
    scaling.value = nn.Parameter(torch.tensor(float(init)))
    tensor_quant = _Holder()
    tensor_quant.scaling_impl = scaling
    fused = _Holder()
    fused.tensor_quant = tensor_quant
    proxy = _Holder()
    proxy.fused_activation_quant_proxy = fused
    return proxy


class _QuantAct(nn.Module):
    def __init__(self, bit_width, signed, narrow_range, return_quant_tensor=False, scaling_init=None):
        super().__init__()
        if return_quant_tensor:
            raise NotImplementedError("return_quant_tensor=True is not used by the reference and not provided")
        self.act_bit_width = int(bit_width)
        self.signed, self.narrow_range = bool(signed), bool(narrow_range)
        self.act_quant = _threshold_tree(1.0 if scaling_init is None else scaling_init)

    @property
    def threshold(self) -> torch.Tensor:
        return self.act_quant.fused_activation_quant_proxy.tensor_quant.scaling_impl.value.abs()

    def int_range(self):
        b = self.act_bit_width
        if self.signed:
            lo = -(1 << (b - 1)) + (1 if self.narrow_range else 0)
            return lo, (1 << (b - 1)) - 1
        return 0, (1 << b) - 1 - (1 if self.narrow_range else 0)

    def quant_act_scale(self) -> torch.Tensor:
        b = self.act_bit_width
        return self.threshold / float((1 << (b - 1)) if self.signed else ((1 << b) - 1))

    def tfx_act_quant(self):
        lo, hi = self.int_range()
        return float(self.quant_act_scale().detach()), lo, hi

    def _quant(self, x):
        s = self.quant_act_scale()
        lo, hi = self.int_range()
        q = torch.clamp(torch.round(x / s), lo, hi)
        return x + (q * s - x).detach() if x.requires_grad else q * s       # straight-through estimator


class QuantIdentity(_QuantAct):
    def __init__(self, act_quant=Int8ActPerTensorFloat, return_quant_tensor=False, bit_width=None, scaling_init=None, **kwargs):
        super().__init__(bit_width if bit_width is not None else act_quant.bit_width, getattr(act_quant, "signed", True),
                         kwargs.pop("narrow_range", getattr(act_quant, "narrow_range", False)), return_quant_tensor, scaling_init)

    def forward(self, x):
        return self._quant(x)


class QuantReLU(_QuantAct):
    def __init__(self, act_quant=Uint8ActPerTensorFloat, return_quant_tensor=False, bit_width=None, scaling_init=None, **kwargs):
        super().__init__(bit_width if bit_width is not None else act_quant.bit_width, False,
                         kwargs.pop("narrow_range", getattr(act_quant, "narrow_range", False)), return_quant_tensor, scaling_init)

    def forward(self, x):
        return self._quant(F.relu(x))


class QuantConv2d(nn.Conv2d):
    def __init__(self, in_channels, out_channels, kernel_size, stride=1, padding=0, dilation=1, groups=1, bias=True,
                 weight_quant=Int8WeightPerTensorFloat, bias_quant=None, input_quant=None, output_quant=None,
                 return_quant_tensor=False, weight_bit_width=None, narrow_range=None, **kwargs):
        super().__init__(in_channels, out_channels, kernel_size, stride=stride, padding=padding, dilation=dilation, groups=groups,
                         bias=bool(bias))
        if return_quant_tensor or input_quant is not None or output_quant is not None:
            raise NotImplementedError("only weight quantisation is used by the reference and provided")
        self.weight_bit_width = int(weight_bit_width if weight_bit_width is not None else weight_quant.bit_width)
        self.narrow_range = bool(getattr(weight_quant, "narrow_range", True) if narrow_range is None else narrow_range)

    def weight_int_range(self):
        b = self.weight_bit_width
        return -(1 << (b - 1)) + (1 if self.narrow_range else 0), (1 << (b - 1)) - 1

    def quant_weight_scale(self) -> torch.Tensor:
        lo, hi = self.weight_int_range()
        amax = self.weight.detach().abs().max()
        return torch.where(amax > 0, amax / float(-lo if self.narrow_range else hi + 1), torch.ones_like(amax))

    def tfx_weight_quant(self):
        """(integer weights int32 [Cout][Cin/groups][kh][kw], scale)"""
        s = self.quant_weight_scale()
        lo, hi = self.weight_int_range()
        return torch.clamp(torch.round(self.weight.detach() / s), lo, hi).to(torch.int32), float(s)

    def quant_weight(self) -> torch.Tensor:
        s = self.quant_weight_scale()
        lo, hi = self.weight_int_range()
        q = torch.clamp(torch.round(self.weight / s), lo, hi) * s
        return self.weight + (q - self.weight).detach() if self.weight.requires_grad else q

    def forward(self, x):
        return F.conv2d(x, self.quant_weight(), self.bias, self.stride, self.padding, self.dilation, self.groups)
