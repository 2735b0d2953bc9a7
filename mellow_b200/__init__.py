"""mellow_b200 -- B200-native implementation of Mellow's two-audio-plus-prompt inference path.

Public surface mirrors the reference package (``from mellow import MellowWrapper``, reference mellow/__init__.py:1).
"""
from .wrapper import MellowWrapper  # noqa: F401
from .engine import Engine, MellowNativeError  # noqa: F401
