"""Exceptions mirroring tiatoolbox/utils/exceptions.py (names and messages)."""
from __future__ import annotations


class FileNotSupportedError(Exception):
    """Raise No supported file found error (tiatoolbox/utils/exceptions.py:8-19)."""

    def __init__(self, message: str = "File format is not supported") -> None:
        self.message = message
        super().__init__(self.message)


class MethodNotSupportedError(Exception):
    """Raise No supported file found error (tiatoolbox/utils/exceptions.py:22-33)."""

    def __init__(self, message: str = "Method is not supported") -> None:
        self.message = message
        super().__init__(self.message)
