"""Stand-in: nerf2mesh.py:1,10 imports trimesh and never uses it."""


def __getattr__(name):
    raise RuntimeError(f"trimesh.{name}: trimesh is not installed (stand-in module of the B200 drop-in launcher)")
