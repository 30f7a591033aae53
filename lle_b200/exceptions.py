"""`lle.exceptions` (python/lle/exceptions/__init__.pyi; src/bindings/pyexceptions.rs): the exception types of the package."""
from .types import InvalidActionError, InvalidLevelError, InvalidWorldStateError, ParsingError

__all__ = ["InvalidActionError", "InvalidLevelError", "InvalidWorldStateError", "ParsingError"]
