def __getattr__(name):
    raise RuntimeError(f"matplotlib.pyplot.{name}: matplotlib is not installed (stand-in module of the B200 drop-in launcher)")
