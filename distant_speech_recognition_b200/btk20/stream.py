"""btk20.stream (stream/stream.i:24-237): stream handle classes and the Python->C++ adapter."""
from .. import _btk20host as _h

VectorCharFeatureStreamPtr = _h.VectorCharFeatureStreamPtr
VectorShortFeatureStreamPtr = _h.VectorShortFeatureStreamPtr
VectorFloatFeatureStreamPtr = _h.VectorFloatFeatureStreamPtr
VectorFeatureStreamPtr = _h.VectorFeatureStreamPtr
VectorComplexFeatureStreamPtr = _h.VectorComplexFeatureStreamPtr
# stream/pyStream.h:136-231: the Python -> C++ adapters of the other element types
PyVectorShortFeatureStreamPtr = _h.PyVectorShortFeatureStreamPtr
PyVectorFloatFeatureStreamPtr = _h.PyVectorFloatFeatureStreamPtr
PyVectorFeatureStreamPtr = _h.PyVectorFeatureStreamPtr


def PyVectorComplexFeatureStreamPtr(obj, nm="PyVectorComplexFeatureStream"):
    """stream/pyStream.h:25-133: wrap a Python object exposing __iter__/next/size/reset as a C++ complex stream.
    Objects that carry a native stream (the btk20.pybeamformer classes) are unwrapped so that the whole graph runs as
    one GPU submission instead of crossing the language boundary once per frame."""
    native = getattr(obj, "native_stream", None)
    if callable(native):
        return native()
    return _h.PyVectorComplexFeatureStreamPtr(obj, nm)
