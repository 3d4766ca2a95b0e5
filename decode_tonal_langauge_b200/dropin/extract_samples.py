"""Drop-in module: same dotted name and entry points as the reference's `extract_samples.py`;
the implementation lives in decode_tonal_langauge_b200 and runs on the B200."""
from decode_tonal_langauge_b200.stages import extract_samples_run as run  # noqa: F401
