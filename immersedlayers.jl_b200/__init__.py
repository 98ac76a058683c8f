"""ilm-b200: B200-native immersed-layer operator hot path (host-side mirror of
the ImmersedLayers.jl operator API over the C-ABI library libilm_b200.so)."""
from . import bodies, lgf  # noqa: F401
