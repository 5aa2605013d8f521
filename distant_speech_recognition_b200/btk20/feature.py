"""btk20.feature (feature/feature.i:205-250): only the input source of the hot path."""
from .._btk20host import SampleFeaturePtr  # noqa: F401
