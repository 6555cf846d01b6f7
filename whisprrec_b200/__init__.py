"""whisprrec_b200: the BPRMF / LightGCN hot path of WhisprRec as hand-written sm_100a CUDA behind a C-ABI.

Layout
  csrc/            CUDA kernels + the C-ABI (include/whisprrec_b200.h)  -> libwhisprrec_b200.so
  _lib.py          ctypes binding
  main.py, helpers/, models/, utils/   host-side mirror of the reference's main.py / reader / runner / model API
"""
__version__ = '0.1.0'
