"""rampvo_b200 — B200-native (sm_100a) implementation of RAMP-VO's per-frame recurrent-update hot
path behind the reference's own Python call signatures (uzh-rpg/rampvo: ramp.altcorr, ramp.fastba,
ramp.projective_ops, ramp.lietorch.SE3).  All compute goes through librampvo_b200.so
(include/rampvo_b200.h); nothing here falls back to PyTorch or the CPU."""
from . import _lib  # noqa: F401

__all__ = ["altcorr", "fastba", "projective_ops", "lietorch"]
