"""Drop-in module: same dotted name and entry points as the reference's `preprocess.car_rereference (the spelling used by CONFIG.md and example_config.yaml)`;
the implementation lives in decode_tonal_langauge_b200 and runs on the B200."""
from preprocess.signal.car_rereference import *  # noqa: F401,F403
from preprocess.signal.car_rereference import run  # noqa: F401
