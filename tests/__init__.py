"""Test suite: CPU tests (oracle vs goldens, host logic, gloo) and GPU parity tests (pytest -m gpu)."""
