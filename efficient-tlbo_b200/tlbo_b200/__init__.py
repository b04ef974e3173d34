"""tlbo_b200 -- B200-native surrogate-and-acquisition hot path of efficient-tlbo.

Host-side mirror of the reference's surrogate / facade / acquisition / optimizer
interfaces (reference: tlbo/model, tlbo/facade, tlbo/acquisition_function,
tlbo/optimizer, tlbo/framework/smbo_offline.py) over hand-written sm_100a CUDA kernels
reached through the C ABI in ``include/tlbo_b200.h``.  No CPU fallback: the compute
entry points raise if the CUDA library or a CUDA device is missing.

Submodules are imported lazily so that ``python -m tlbo_b200.build`` works before the
library exists.
"""
__version__ = '0.1.0'
