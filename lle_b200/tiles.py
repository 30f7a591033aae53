"""`lle.tiles` (python/lle/tiles/__init__.pyi): Direction and the tile handles `Gem`, `Laser`, `LaserSource`."""
from .types import Direction, Gem, Laser, LaserSource

__all__ = ["Direction", "Gem", "Laser", "LaserSource"]
