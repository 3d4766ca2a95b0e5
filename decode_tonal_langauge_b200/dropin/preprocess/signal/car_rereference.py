"""Drop-in module: same dotted name and entry points as the reference's `preprocess/signal/car_rereference.py`;
the implementation lives in decode_tonal_langauge_b200 and runs on the B200."""
from decode_tonal_langauge_b200.steps import car_rereference as run  # noqa: F401
