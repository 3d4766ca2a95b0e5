"""Drop-in module: same dotted name and entry points as the reference's `preprocess/preprocessor.py`;
the implementation lives in decode_tonal_langauge_b200 and runs on the B200."""
from decode_tonal_langauge_b200.preprocessor import preprocess_modalities, preprocess_signal  # noqa: F401
