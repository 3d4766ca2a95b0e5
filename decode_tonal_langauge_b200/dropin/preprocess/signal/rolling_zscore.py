"""Drop-in module: same dotted name and entry points as the reference's `preprocess/signal/rolling_zscore.py`;
the implementation lives in decode_tonal_langauge_b200 and runs on the B200."""
from decode_tonal_langauge_b200.steps import rolling_zscore as run  # noqa: F401
