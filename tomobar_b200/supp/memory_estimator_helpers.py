"""Dry-run device-memory accounting with the interface of the reference's
``tomobar/supp/memory_estimator_helpers.py`` (``DeviceMemStack``): while a stack is active, a method called
with a SHAPE TUPLE instead of an array records the allocations it would make and returns the shape of
its result; ``highwater`` is then the peak number of bytes (httomolibgpu sizes its slice chunks from it,
``methodsDIR_CuPy.py:253-258, 437-441``).

The numbers are those of THIS implementation's allocation sequence, not the reference's."""

from __future__ import annotations

from collections import Counter
from typing import Optional

ALLOCATION_UNIT_SIZE = 512  # granularity of the caching allocators (CuPy's and torch's small-block size)


def _rounded(nbytes: int) -> int:
    return -(-int(nbytes) // ALLOCATION_UNIT_SIZE) * ALLOCATION_UNIT_SIZE


class DeviceMemStack:
    """``with DeviceMemStack() as stack: method(shape_tuple, ...)`` then read ``stack.highwater``.
    Nested ``with`` blocks share the outermost stack, like the reference's."""

    _active: Optional["DeviceMemStack"] = None
    _depth = 0

    def __init__(self) -> None:
        self._live = Counter()  # requested size -> number of live blocks of that size
        self.current = 0
        self.highwater = 0

    def __enter__(self) -> "DeviceMemStack":
        cls = DeviceMemStack
        if cls._depth == 0:
            cls._active = self
        cls._depth += 1
        return self

    def __exit__(self, exc_type, exc_value, traceback) -> None:
        cls = DeviceMemStack
        cls._depth -= 1
        if cls._depth == 0:
            cls._active = None

    @classmethod
    def instance(cls) -> Optional["DeviceMemStack"]:
        return cls._active

    @property
    def allocations(self):
        """Live block sizes (the reference keeps them in a list of the same name)."""
        return sorted(self._live.elements())

    def malloc(self, byte_count) -> None:
        byte_count = int(byte_count)
        self._live[byte_count] += 1
        self.current += _rounded(byte_count)
        self.highwater = max(self.highwater, self.current)

    def free(self, byte_count) -> None:
        byte_count = int(byte_count)
        if self._live[byte_count] <= 0:
            raise AssertionError(f"DeviceMemStack.free({byte_count}): no live block of that size")
        self._live[byte_count] -= 1
        self.current -= _rounded(byte_count)
