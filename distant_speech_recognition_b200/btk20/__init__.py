"""btk20 — the reference's Python module surface (btk20.stream / feature / modulated / beamformer / postfilter /
pybeamformer) re-implemented over the C++ host mirror (`_btk20host`, pybind11) and the sm_100a CUDA library (`libbtkb.so`).
Class names, keyword arguments and iteration semantics follow the SWIG interface files of btk2.0 (btk20_src/*/*.i);
see INTEGRATION.md.  There is no CPU fallback: the first `next()` on a graph needs a CUDA device."""
from . import common, stream, feature, modulated, beamformer, postfilter, dereverberation  # noqa: F401
