"""SimilarityConfig mirror (/root/reference/crates/cortex-core/src/vector/config.rs:3-87).

The thresholds feed the scan epilogues; their comparison operators are the
reference's: linker `score >= auto_link_threshold` (linker/rules.rs:50), dedup
`score >= dedup_threshold` via search_threshold (vector/index.rs:384-387).
"""
from __future__ import annotations

from dataclasses import dataclass

from .index import CortexError


def _clamp01(x: float) -> float:
    return min(max(float(x), 0.0), 1.0)


@dataclass
class SimilarityConfig:
    auto_link_threshold: float = 0.75
    dedup_threshold: float = 0.92
    contradiction_threshold: float = 0.80
    auto_link_k: int = 20

    @classmethod
    def new(cls) -> "SimilarityConfig":
        return cls()

    def with_auto_link_threshold(self, t: float) -> "SimilarityConfig":
        self.auto_link_threshold = _clamp01(t)
        return self

    def with_dedup_threshold(self, t: float) -> "SimilarityConfig":
        self.dedup_threshold = _clamp01(t)
        return self

    def with_contradiction_threshold(self, t: float) -> "SimilarityConfig":
        self.contradiction_threshold = _clamp01(t)
        return self

    def with_auto_link_k(self, k: int) -> "SimilarityConfig":
        self.auto_link_k = int(k)
        return self

    def validate(self) -> None:
        """config.rs:66-86"""
        if self.auto_link_threshold >= self.dedup_threshold:
            raise CortexError("auto_link_threshold must be less than dedup_threshold")
        if self.contradiction_threshold >= self.dedup_threshold:
            raise CortexError("contradiction_threshold must be less than dedup_threshold")
        if self.auto_link_k == 0:
            raise CortexError("auto_link_k must be greater than 0")
