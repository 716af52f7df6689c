"""lcrnet_b200 -- B200 (sm_100a) implementation of the LCR-Net inference hot path.

All compute lives in ``liblcr_b200.so`` (hand-written CUDA behind the C ABI declared in
``include/lcr_b200.h``); this package is the host-side mirror of the reference's operator and
model interfaces (``utils.ext``, ``experiments/lcrnet/modules/ops``, ``data.py``, the model
family).  There is no CPU fallback: every operator raises if the CUDA library is missing.
"""
__version__ = '0.1.0'
