"""Stub of torchmetrics (tests/stubs/__init__.py explains): the three image metrics dn_model.py constructs."""
