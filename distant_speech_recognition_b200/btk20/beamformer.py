"""btk20.beamformer (beamformer/beamformer.i:46-568): snapshot / spectral-matrix arrays and the subband beamformers of the hot path."""
from .._btk20host import (SnapShotArrayPtr, SpectralMatrixArrayPtr, SubbandOrthogonalizerPtr, SubbandBeamformerPtr, SubbandDSPtr, SubbandGSCPtr, SubbandMVDRPtr,  # noqa: F401
                          SubbandMVDRGSCPtr, SubbandGSCRLSPtr, SubbandGSCLMSPtr, LmsConfig, SubbandGSCRLSNativePtr, RlsConfig, SubbandSOSNativePtr, calc_all_delays)
