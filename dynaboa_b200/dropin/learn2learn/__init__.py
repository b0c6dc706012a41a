"""Minimal `learn2learn` name-space: only `algorithms.MAML`, which is all the reference imports."""
from . import algorithms  # noqa: F401
