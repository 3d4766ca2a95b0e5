"""Drop-in module: same dotted name and entry points as the reference's `channel_selection/active.py`;
the implementation lives in decode_tonal_langauge_b200 and runs on the B200."""
from decode_tonal_langauge_b200.selection import active_run as run  # noqa: F401
