"""``PatchCoords`` -- the coordinate record returned by the crop functions.

Mirrors ``biapy/data/dataset.py:484-527`` (same constructor, same attribute names; 2D patches carry no
``z_*`` attributes, exactly like the reference).
"""
from __future__ import annotations

from typing import Optional, Tuple


class PatchCoords:
    def __init__(self, y_start: int, y_end: int, x_start: int, x_end: int,
                 z_start: Optional[int] = None, z_end: Optional[int] = None):
        self.y_start = y_start
        self.y_end = y_end
        self.x_start = x_start
        self.x_end = x_end
        if z_start is not None:
            self.z_start = z_start
        if z_end is not None:
            self.z_end = z_end

    def extract_shape_from_coords(self) -> Tuple[int, ...]:
        if hasattr(self, "z_start"):
            return (self.z_end - self.z_start, self.y_end - self.y_start, self.x_end - self.x_start)
        return (self.y_end - self.y_start, self.x_end - self.x_start)

    def __repr__(self) -> str:
        if hasattr(self, "z_start"):
            return (f"PatchCoords(z={self.z_start}:{self.z_end}, y={self.y_start}:{self.y_end}, "
                    f"x={self.x_start}:{self.x_end})")
        return f"PatchCoords(y={self.y_start}:{self.y_end}, x={self.x_start}:{self.x_end})"

    def __eq__(self, other) -> bool:
        return isinstance(other, PatchCoords) and self.__dict__ == other.__dict__
