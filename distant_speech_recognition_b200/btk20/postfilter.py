"""btk20.postfilter (postfilter/postfilter.i:46-188)."""
from .._btk20host import ZelinskiPostFilterPtr, McCowanPostFilterPtr, LefkimmiatisPostFilterPtr  # noqa: F401
