"""Stub of nerfstudio 1.1.3 (tests/stubs/__init__.py explains)."""
__version__ = "1.1.3-stub"
