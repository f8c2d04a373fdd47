"""rrtplanner_b200 -- B200-native tree-expansion hot path behind rrtplanner's planner API."""
from . import worlds  # noqa: F401
