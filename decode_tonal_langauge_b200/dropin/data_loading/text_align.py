"""Drop-in module: same dotted name and entry points as the reference's `data_loading/text_align.py`;
the implementation lives in decode_tonal_langauge_b200 and runs on the B200."""
from decode_tonal_langauge_b200.epochs import extract_ecog_audio  # noqa: F401
from decode_tonal_langauge_b200.textgrid_io import TextGrid, handle_textgrids, read_textgrid  # noqa: F401
