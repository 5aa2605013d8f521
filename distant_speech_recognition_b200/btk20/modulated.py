"""btk20.modulated (modulated/modulated.i:91-192): the oversampled DFT filter banks."""
import numpy as np
from .._btk20host import OverSampledDFTAnalysisBankPtr, OverSampledDFTSynthesisBankPtr  # noqa: F401


def get_window(winType, winLen):
    """modulated/modulated.cc:47-83 get_window: 0 rectangle, 1 Hamming, 2 Hanning."""
    i = np.arange(winLen, dtype=np.float64)
    if winType == 0:
        return np.ones(winLen)
    if winType == 2:
        return 0.5 * (1.0 - np.cos(2.0 * np.pi * i / (winLen - 1)))
    return 0.54 - 0.46 * np.cos(2.0 * np.pi * i / (winLen - 1))
