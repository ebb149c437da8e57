"""Re-export of the model-level checkers (oracle/model_ops.py) for the tests."""
from oracle.model_ops import CostVolumeOps, TorchCorrelationOps, deterministic_init  # noqa: F401
