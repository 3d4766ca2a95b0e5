"""Drop-in module: same dotted name and entry points as the reference's `preprocess_main.py`;
the implementation lives in decode_tonal_langauge_b200 and runs on the B200."""
import sys

from decode_tonal_langauge_b200.stages import main, preprocess_run as run  # noqa: F401

if __name__ == "__main__":
    if len(sys.argv) != 2:
        raise SystemExit("Usage: python preprocess_main.py <config.yaml>")
    main(sys.argv[1])
