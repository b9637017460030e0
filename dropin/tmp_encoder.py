"""train_hash2.py:18 does `from tmp_encoder import *`; the file is missing from the reference repo. Empty on purpose."""
