"""asv_subtools_b200 -- B200 (sm_100a) x-vector extraction + back-end scoring.

Drop-in for the PyTorch extraction / scoring hot path of Snowdar/asv-subtools: same model
blueprints (`model/xvector.py`, `model/ecapa_tdnn_xvector.py` with the reference's
constructor signatures and state_dict keys), same `extract_embedding()` surface
(`libs.nnet.framework.TopVirtualNnet`), same Kaldi ark/scp formats -- but every arithmetic
step runs in hand-written CUDA (`csrc/`, exported through the C ABI of `include/xvb200.h`).
There is no CPU fallback: without `libxvb200.so` or without a B200 the ops raise.
"""
from . import _lib  # noqa: F401  (fails loudly if the native library is missing)

__all__ = ["_lib"]
