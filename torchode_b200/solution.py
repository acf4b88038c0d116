"""Result container (mirrors torchode/solution.py:6-23)."""
from typing import Any, Dict

import torch


class Solution:
    def __init__(self, ts: torch.Tensor, ys: torch.Tensor, stats: Dict[str, Any], status: torch.Tensor):
        self.ts = ts
        self.ys = ys
        self.stats = stats
        self.status = status

    def __repr__(self):
        return f"Solution(ts={self.ts}, ys={self.ys}, stats={self.stats}, status={self.status})"
