"""Solver status codes (same integer values as torchode/status_codes.py:9-33)."""
from enum import Enum

SUCCESS = 0
GENERAL_ERROR = 1
REACHED_DT_MIN = 2
REACHED_MAX_STEPS = 3
INFINITE_NORM = 4


class Status(Enum):
    """Per-sample outcome of a solve; anything above SUCCESS is abnormal."""

    SUCCESS = SUCCESS
    GENERAL_ERROR = GENERAL_ERROR
    REACHED_DT_MIN = REACHED_DT_MIN
    REACHED_MAX_STEPS = REACHED_MAX_STEPS
    # non-finite error ratio (NaN / inf in y or f)
    INFINITE_NORM = INFINITE_NORM
