"""Drop-in module: same dotted name and entry points as the reference's `preprocess.downsample (the spelling used by CONFIG.md and example_config.yaml)`;
the implementation lives in decode_tonal_langauge_b200 and runs on the B200."""
from preprocess.signal.downsample import *  # noqa: F401,F403
from preprocess.signal.downsample import run  # noqa: F401
