"""`lle.observations.ObservationType` (python/lle/observations.py:37-63): the observation presets.  The generators
themselves are store epilogues of the step kernel; "rgb-image" needs the renderer and is not on this path."""
from enum import Enum


class ObservationType(str, Enum):
    NORMALIZED_STATE = "normalized-state"
    STATE = "state"
    RGB_IMAGE = "rgb-image"
    LAYERED = "layered"
    FLATTENED = "flattened"
    PARTIAL_3x3 = "partial3x3"
    PARTIAL_5x5 = "partial5x5"
    PARTIAL_7x7 = "partial7x7"
    LAYERED_PADDED = "layered-padded"
    LAYERED_PADDED_1AGENT = "layered-padded-1"
    LAYERED_PADDED_2AGENTS = "layered-padded-2"
    LAYERED_PADDED_3AGENTS = "layered-padded-3"
    AGENT0_PERSPECTIVE_LAYERED = "perspective"

    @staticmethod
    def from_str(s: str) -> "ObservationType":
        return ObservationType(s)

    def __str__(self) -> str:  # so that the value can be handed to `obs_type=` arguments directly
        return self.value
