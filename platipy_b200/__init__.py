"""
platipy_b200 -- B200-native (sm_100a) Demons registration / resampling / label-fusion engine, a drop-in
for the hot path of pyplati/platipy.  All compute runs in hand-written CUDA kernels in
``platipy_b200/csrc`` behind the C ABI declared in ``include/b200reg.h``; there is no CPU fallback.
"""
__version__ = "0.1.0"
