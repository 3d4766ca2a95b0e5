"""Drop-in module: same dotted name and entry points as the reference's `utils/config.py`;
the implementation lives in decode_tonal_langauge_b200 and runs on the B200."""
from decode_tonal_langauge_b200.config import (  # noqa: F401
    dict_to_namespace, generate_hash_name_from_config, load_config, update_configuration)
