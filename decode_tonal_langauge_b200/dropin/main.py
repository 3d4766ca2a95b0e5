"""Drop-in module: same dotted name and entry points as the reference's `main.py`;
the implementation lives in decode_tonal_langauge_b200 and runs on the B200."""
import sys

from decode_tonal_langauge_b200.stages import STAGES, run_pipeline  # noqa: F401

if __name__ == "__main__":
    if len(sys.argv) != 2:
        raise SystemExit("Usage: python main.py <config.yaml>")
    run_pipeline(sys.argv[1])
