"""btk20.postfilter (postfilter/postfilter.i:46-90)."""
from .._btk20host import ZelinskiPostFilterPtr  # noqa: F401
