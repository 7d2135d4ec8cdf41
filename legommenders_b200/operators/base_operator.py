"""Operator contract: see legommenders_b200/contracts.py (kept importable under the reference's module name)."""
from ..contracts import BaseOperator, BaseOperatorConfig  # noqa: F401
