"""Drop-in for the reference module of the same name (tf_ops/*/tf_sampling.py)."""
from .ops import farthest_point_sample, gather_point  # noqa: F401
