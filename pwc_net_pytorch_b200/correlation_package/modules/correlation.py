"""Mirror of correlation_package/modules/correlation.py:6-27 (the nn.Module)."""
from torch.nn.modules.module import Module

from ..functions.correlation import CorrelationFunction


class Correlation(Module):
    """Same constructor defaults and attribute names as the reference module
    (modules/correlation.py:8-21); parameter-free, so state_dicts are unaffected."""

    def __init__(self, pad_size=0, kernel_size=0, max_displacement=0, stride1=1, stride2=2,
                 corr_multiply=1):
        super(Correlation, self).__init__()
        self.pad_size = pad_size
        self.kernel_size = kernel_size
        self.max_displacement = max_displacement
        self.stride1 = stride1
        self.stride2 = stride2
        self.corr_multiply = corr_multiply

    def forward(self, input1, input2):
        return CorrelationFunction.apply(input1, input2, self.pad_size, self.kernel_size,
                                         self.max_displacement, self.stride1, self.stride2,
                                         self.corr_multiply)

    def extra_repr(self):
        return (f"pad_size={self.pad_size}, kernel_size={self.kernel_size}, "
                f"max_displacement={self.max_displacement}, stride1={self.stride1}, "
                f"stride2={self.stride2}")
