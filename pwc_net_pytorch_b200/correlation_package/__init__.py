"""Drop-in for the reference's correlation_package (same sub-module layout and names):

    from pwc_net_pytorch_b200.correlation_package.modules.correlation import Correlation
    from pwc_net_pytorch_b200.correlation_package.functions.correlation import CorrelationFunction

mirror `correlation_package/modules/correlation.py` and `.../functions/correlation.py` of
daigo0927/PWC-Net_pytorch.  `pwc_net_pytorch_b200.install_as_reference_modules()` additionally
registers the package under the reference's own top-level name `correlation_package`, so
`model.py:8` (`from correlation_package.modules.correlation import Correlation`) resolves here.
"""
