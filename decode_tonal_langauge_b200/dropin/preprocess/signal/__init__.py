"""Drop-in module: same dotted name and entry points as the reference's `preprocess/signal/`;
the implementation lives in decode_tonal_langauge_b200 and runs on the B200."""
