"""Test-harness stand-in for hydra (not installed here, no network).

Only used by tests/golden/make_golden.py to import the *reference* package from
/root/reference in this container; never imported by the product.
"""
from . import utils  # noqa: F401
