class Line2D:
    def __init__(self, *a, **k):
        raise RuntimeError("matplotlib.lines.Line2D: matplotlib is not installed (stand-in module of the B200 drop-in launcher)")
