"""Print the measured on-chip roofline denominators (rrtplanner_b200/peaks.py) as one JSON line."""
import json
import sys
sys.path.insert(0, ".")
from rrtplanner_b200 import peaks

print(json.dumps(peaks.measure()))
