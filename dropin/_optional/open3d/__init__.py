"""Stand-in for the part of open3d nerf2mesh.py:101-107 touches: a TriangleMesh container, the Vector3*Vector converters,
LineSet.create_from_triangle_mesh and visualization.draw_geometries.  The containers hold numpy arrays; opening the viewer
raises -- unless HBR_MESH_OUT names a file, in which case the first mesh is written there as an ASCII PLY (a headless box
has no window to open, and the mesh is what the run is for)."""
import os

import numpy as np


def _np(a, dtype):
    if hasattr(a, "detach"):
        a = a.detach().cpu().numpy()
    return np.ascontiguousarray(np.asarray(a), dtype=dtype)


class _Utility:
    @staticmethod
    def Vector3dVector(a):
        return _np(a, np.float64).reshape(-1, 3)

    @staticmethod
    def Vector3iVector(a):
        return _np(a, np.int32).reshape(-1, 3)


class TriangleMesh:
    def __init__(self):
        self.vertices = np.zeros((0, 3))
        self.triangles = np.zeros((0, 3), dtype=np.int32)
        self.vertex_colors = np.zeros((0, 3))


class LineSet:
    def __init__(self, mesh=None):
        self.mesh = mesh

    @staticmethod
    def create_from_triangle_mesh(mesh):
        return LineSet(mesh)


class _Geometry:
    TriangleMesh = TriangleMesh
    LineSet = LineSet


def write_ply(path, mesh):
    v, f, c = np.asarray(mesh.vertices), np.asarray(mesh.triangles), np.asarray(mesh.vertex_colors)
    has_c = c.shape[0] == v.shape[0] and v.shape[0] > 0
    with open(path, "w") as fh:
        fh.write("ply\nformat ascii 1.0\n")
        fh.write(f"element vertex {v.shape[0]}\nproperty float x\nproperty float y\nproperty float z\n")
        if has_c:
            fh.write("property uchar red\nproperty uchar green\nproperty uchar blue\n")
        fh.write(f"element face {f.shape[0]}\nproperty list uchar int vertex_indices\nend_header\n")
        cc = np.clip(np.nan_to_num(c) * 255.0, 0, 255).astype(np.uint8) if has_c else None
        for i in range(v.shape[0]):
            line = f"{v[i, 0]:.6f} {v[i, 1]:.6f} {v[i, 2]:.6f}"
            if has_c:
                line += f" {cc[i, 0]} {cc[i, 1]} {cc[i, 2]}"
            fh.write(line + "\n")
        for t in f:
            fh.write(f"3 {t[0]} {t[1]} {t[2]}\n")


class _Visualization:
    @staticmethod
    def draw_geometries(geoms, window_name="", **kw):
        out = os.environ.get("HBR_MESH_OUT")
        if out:
            for g in geoms:
                if isinstance(g, TriangleMesh):
                    write_ply(out, g)
                    print(f"[open3d stand-in] wrote {len(g.vertices)} vertices / {len(g.triangles)} triangles to {out}")
                    return
        raise RuntimeError("open3d is not installed: no viewer to open (set HBR_MESH_OUT=<file.ply> to write the mesh instead)")


geometry = _Geometry
utility = _Utility
visualization = _Visualization
