"""The float64 stand-ins of the training stage ops live in oracle/train_ops.py (CPU restatement, test infrastructure)."""
from oracle.train_ops import *  # noqa: F401,F403
from oracle.train_ops import _ot, _attend, _mask  # noqa: F401
