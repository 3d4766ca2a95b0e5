"""Drop-in module: same dotted name and entry points as the reference's `data_loading/utils.py`;
the implementation lives in decode_tonal_langauge_b200 and runs on the B200."""
from decode_tonal_langauge_b200.epochs import extract_block_id, match_filename  # noqa: F401
