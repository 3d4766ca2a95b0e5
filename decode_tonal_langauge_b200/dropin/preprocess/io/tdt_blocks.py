"""Drop-in module: same dotted name and entry points as the reference's `preprocess/io/tdt_blocks.py`;
the implementation lives in decode_tonal_langauge_b200 and runs on the B200."""
from decode_tonal_langauge_b200.stages import load_block, save_block  # noqa: F401
