"""Drop-in module: same dotted name and entry points as the reference's `data_loading/sample_loading.py`
(the loader of the hot path's outputs; the classifiers that consume it are out of scope)."""
from decode_tonal_langauge_b200.samples import ClassificationSampleHandler  # noqa: F401
