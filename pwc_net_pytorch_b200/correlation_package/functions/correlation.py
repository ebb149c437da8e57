"""Mirror of correlation_package/functions/correlation.py:5-56 (the autograd Function).

`CorrelationFunction.apply(input1, input2, pad_size=3, kernel_size=3, max_displacement=20,
stride1=1, stride2=2, corr_multiply=1)` returns one tensor; backward returns
`(grad_input1, grad_input2) + (None,) * 6`, exactly like the reference.
"""
from ...functional import CorrelationFunction

__all__ = ["CorrelationFunction"]
