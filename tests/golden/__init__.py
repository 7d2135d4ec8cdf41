"""Golden vectors minted from the live reference (make_golden.py) and the case definitions shared by the tests."""
