"""Stand-in: the reference imports h5py (train_hash2.py:4, vol_renderer.py:4) and never calls it on the hot path."""


def __getattr__(name):
    raise RuntimeError(f"h5py.{name}: h5py is not installed (stand-in module of the B200 drop-in launcher)")
