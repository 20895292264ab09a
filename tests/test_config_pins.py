"""SimilarityConfig pins (vector/config.rs:89-135, vector/tests.rs:122-135) and the
threshold comparison semantics the scan epilogues must honour (linker/rules.rs:50,
linker/rules.rs:403-421)."""
import pytest

from cortex_b200 import CortexError, SimilarityConfig


def test_default_config():
    c = SimilarityConfig()
    assert (c.auto_link_threshold, c.dedup_threshold, c.contradiction_threshold, c.auto_link_k) == (0.75, 0.92, 0.80, 20)
    c.validate()


def test_config_builder():
    c = SimilarityConfig.new().with_auto_link_threshold(0.70).with_dedup_threshold(0.95).with_auto_link_k(30)
    assert (c.auto_link_threshold, c.dedup_threshold, c.auto_link_k) == (0.70, 0.95, 30)


def test_invalid_config():
    with pytest.raises(CortexError):
        SimilarityConfig.new().with_auto_link_threshold(0.95).with_dedup_threshold(0.90).validate()
    with pytest.raises(CortexError):
        SimilarityConfig.new().with_contradiction_threshold(0.95).validate()
    with pytest.raises(CortexError):
        SimilarityConfig.new().with_auto_link_k(0).validate()


def test_clamping():
    c = SimilarityConfig.new().with_auto_link_threshold(1.5).with_dedup_threshold(-0.5)
    assert c.auto_link_threshold == 1.0 and c.dedup_threshold == 0.0


def test_link_rule_threshold_is_inclusive():
    c = SimilarityConfig()
    assert 0.8 >= c.auto_link_threshold and not (0.5 >= c.auto_link_threshold)
    assert 0.75 >= c.auto_link_threshold  # `score >= threshold`, rules.rs:50
