"""Constants of the hot path (reference tlbo/utils/constants.py)."""
MAXINT = 2 ** 31 - 1
SUCCESS = 0
FAILDED = 1
TIMEOUT = 2
# avoids numerical instabilities / division by zero (reference constants.py:11)
VERY_SMALL_NUMBER = 1e-10
