"""Run an UNMODIFIED script of RishabhSri14/Human-Body-Reconstruction (train_hash2.py, nerf2mesh.py) on the B200 drop-ins.

    python /path/to/this/repo/launch.py /path/to/Human-Body-Reconstruction/train_hash2.py --hash_size 19 --num_batch 4096 ...
    python /path/to/this/repo/launch.py /path/to/Human-Body-Reconstruction/nerf2mesh.py --ckpt_name default --bound_pth bounds_model.npy

`python script.py` puts the script's own directory in front of PYTHONPATH, so `PYTHONPATH=dropin python train_hash2.py`
would import the reference's hash_encoding.py / encoder.py / ... sitting next to the script.  This launcher builds the
import path itself:

    sys.path = [ <repo>/dropin,            # hash_encoding, encoder, test_hash, helper, vol_renderer, tmp_encoder, torchmcubes
                 <script directory>,       # everything else stays the reference's own (dataset.py, dataset_new.py, ...)
                 ... the interpreter's path ...,
                 <repo>/dropin/_optional ] # stand-ins for packages the scripts import but the hot path never uses
                                           # (h5py, matplotlib, trimesh, open3d) -- LAST, so real installs win

and runs the script with runpy as __main__, arguments untouched.  Nothing in the reference tree is edited or copied.
"""
import os
import runpy
import sys

ROOT = os.path.dirname(os.path.abspath(__file__))
DROPIN = os.path.join(ROOT, "dropin")
HOT_PATH_MODULES = ("hash_encoding", "encoder", "test_hash", "helper", "vol_renderer", "tmp_encoder", "torchmcubes")


def prepare(script: str) -> str:
    """Set up sys.path for `script` (see the module docstring).  Returns the absolute script path."""
    script = os.path.abspath(script)
    if not os.path.isfile(script):
        raise FileNotFoundError(script)
    script_dir = os.path.dirname(script)
    for p in (DROPIN, script_dir, ROOT):
        while p in sys.path:
            sys.path.remove(p)
    sys.path[0:0] = [DROPIN, script_dir]
    sys.path.append(ROOT)
    sys.path.append(os.path.join(DROPIN, "_optional"))
    for name in HOT_PATH_MODULES:                 # a previously imported namesake must not win
        sys.modules.pop(name, None)
    return script


def served_by_dropin() -> dict:
    """module name -> file, for the hot-path modules that have been imported; raises if any came from elsewhere."""
    out = {}
    for name in HOT_PATH_MODULES:
        mod = sys.modules.get(name)
        if mod is None:
            continue
        f = os.path.abspath(getattr(mod, "__file__", "") or "")
        if not f.startswith(DROPIN + os.sep):
            raise RuntimeError(f"{name} was imported from {f}, not from {DROPIN}")
        out[name] = f
    return out


def run(script: str, argv=()):
    """Run `script` as __main__ with `argv`; returns the script's globals (runpy)."""
    script = prepare(script)
    old_argv = sys.argv
    sys.argv = [script, *argv]
    try:
        return runpy.run_path(script, run_name="__main__")
    finally:
        sys.argv = old_argv
        served_by_dropin()


def main():
    if len(sys.argv) < 2 or sys.argv[1] in ("-h", "--help"):
        print(__doc__)
        sys.exit(0 if len(sys.argv) >= 2 else 2)
    run(sys.argv[1], sys.argv[2:])


if __name__ == "__main__":
    main()
