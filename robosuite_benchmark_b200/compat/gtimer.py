"""gtimer, as far as util/rlkit_custom.py uses it (timed_for, stamp, get_times): robosuite_benchmark_b200.rlkit_api.GTimer."""
from robosuite_benchmark_b200.rlkit_api import gt as _gt

stamp = _gt.stamp
timed_for = _gt.timed_for
get_times = _gt.get_times
reset = _gt.reset
