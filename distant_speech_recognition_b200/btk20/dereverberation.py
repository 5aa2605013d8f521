"""btk20.dereverberation (dereverberation/dereverberation.i:46-185): WPE dereverberation in the subband domain — the single-channel
feature and the multi-channel estimator + per-channel feature streams, used as in unit_test/test_subband_dereverberator.py:53-170.
The estimation (theta / weighted correlation / Cholesky solve per bin) and the output stage run as CUDA kernels (csrc/btkb_wpe.cu)."""
from .._btk20host import (SingleChannelWPEDereverberationFeaturePtr, MultiChannelWPEDereverberationPtr,  # noqa: F401
                          MultiChannelWPEDereverberationFeaturePtr)
