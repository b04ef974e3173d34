"""Constants the hot path's callers share (names as in reference tlbo/utils/constants.py)."""
import numpy as np

# trial states returned by ``SMBO_*.iterate`` (the reference's names, including its spelling)
SUCCESS, FAILDED, TIMEOUT = range(3)
# "no objective value": the largest int32 (a failed evaluation reports it)
MAXINT = int(np.iinfo(np.int32).max)
# variance floor of the GP / facade predictions (constants.py:11)
VERY_SMALL_NUMBER = 1e-10
