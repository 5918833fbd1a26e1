"""Drop-in for the reference module of the same name (tf_ops/*/tf_grouping.py)."""
from .ops import query_ball_point, group_point  # noqa: F401
