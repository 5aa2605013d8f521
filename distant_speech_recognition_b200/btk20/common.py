"""btk20.common — exception names of common/jexception.h:44-161 as Python sees them through include/jexception.i:20-86."""
from .._btk20host import j_error  # noqa: F401
