"""Drop-in module: same dotted name and entry points as the reference's `preprocess/pipelines/subject_block.py`;
the implementation lives in decode_tonal_langauge_b200 and runs on the B200."""
from decode_tonal_langauge_b200.stages import (  # noqa: F401
    generate_setup_name, get_block_id, iter_blocks, subject_block_run as run)
