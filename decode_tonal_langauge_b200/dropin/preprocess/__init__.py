"""Drop-in module: same dotted name and entry points as the reference's `preprocess/__init__.py`;
the implementation lives in decode_tonal_langauge_b200 and runs on the B200."""
