"""Drop-in module: same dotted name and entry points as the reference's `channel_selection_main.py`;
the implementation lives in decode_tonal_langauge_b200 and runs on the B200."""
import sys

from decode_tonal_langauge_b200.stages import channel_selection_run as run  # noqa: F401

if __name__ == "__main__":
    from decode_tonal_langauge_b200.config import load_config
    if len(sys.argv) != 2:
        raise SystemExit("Usage: python channel_selection_main.py <config.yaml>")
    run(load_config(sys.argv[1]))
