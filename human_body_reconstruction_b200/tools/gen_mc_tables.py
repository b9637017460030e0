"""Derives the 256-case marching-cubes tables from first principles and writes
   csrc/mc_tables.h   (device tables for grid_mc.cu)
   oracle/mc_tables.py (the same tables for the CPU checker)

No table is copied from anywhere: for every corner configuration the iso-contour is traced on the six cube
faces (one segment per pair of crossings; on an ambiguous face -- four crossings -- every INSIDE corner is cut
off on its own, a rule that depends only on the face's four corner states so neighbouring cells agree and the
surface is watertight), the directed segments are chained into closed loops and each loop is fan-triangulated.

Conventions (shared with grid_mc.cu):
  corner v in 0..7 sits at (v&1, (v>>1)&1, (v>>2)&1) along grid axes (0,1,2); bit v of the case index is set
  iff the corner is inside (density < iso);
  edge e = 4*a + 2*hi + lo runs along axis a from the corner whose other two coordinates are (lo, hi)
  (in increasing axis order).
"""
import os

CORNER = [((v & 1), (v >> 1) & 1, (v >> 2) & 1) for v in range(8)]


def corner_id(c):
    return c[0] | (c[1] << 1) | (c[2] << 2)


def edge_of(u, v):
    """edge id of the cube edge joining corners u and v (they differ in exactly one coordinate)."""
    cu, cv = CORNER[u], CORNER[v]
    a = [i for i in range(3) if cu[i] != cv[i]]
    assert len(a) == 1
    a = a[0]
    others = [i for i in range(3) if i != a]
    lo, hi = cu[others[0]], cu[others[1]]
    return 4 * a + 2 * hi + lo


def edge_corners(e):
    a, r = divmod(e, 4)
    lo, hi = r & 1, r >> 1
    others = [i for i in range(3) if i != a]
    c0 = [0, 0, 0]
    c0[others[0]], c0[others[1]] = lo, hi
    c1 = list(c0)
    c1[a] = 1
    return corner_id(c0), corner_id(c1)


def faces_ccw():
    """Each face as 4 corner ids, counter-clockwise when seen from outside the cube."""
    out = []
    for a in range(3):
        u, w = [i for i in range(3) if i != a]          # u x w = +a  iff (u,w,a) is a cyclic permutation of (0,1,2)
        cyclic = (u, w, a) in ((0, 1, 2), (1, 2, 0), (2, 0, 1))
        for side in (0, 1):
            quad = []
            for (pu, pw) in ((0, 0), (1, 0), (1, 1), (0, 1)):
                c = [0, 0, 0]
                c[a], c[u], c[w] = side, pu, pw
                quad.append(corner_id(c))
            # (0,0)->(1,0)->(1,1)->(0,1) is CCW about +(u x w); outward normal is +a on side 1, -a on side 0
            ccw_about_plus_a = cyclic
            want_plus = side == 1
            if ccw_about_plus_a != want_plus:
                quad.reverse()
            out.append(quad)
    return out


FACES = faces_ccw()


def case_triangles(case):
    inside = [(case >> v) & 1 for v in range(8)]
    nxt = {}                                             # directed segments: start edge -> end edge
    for quad in FACES:
        exits, enters = [], []
        for i in range(4):
            u, v = quad[i], quad[(i + 1) % 4]
            if inside[u] and not inside[v]:
                exits.append((i, edge_of(u, v)))
            elif not inside[u] and inside[v]:
                enters.append((i, edge_of(u, v)))
        for i, e_out in exits:
            # the inside corner quad[i] is cut off: its exit edge is E_i, its enter edge is E_{i-1} when the
            # previous corner is outside; otherwise walk back over inside corners to the enter edge.
            j = i
            while inside[quad[(j - 1) % 4]]:
                j = (j - 1) % 4
            e_in = edge_of(quad[(j - 1) % 4], quad[j])
            assert e_out not in nxt
            nxt[e_out] = e_in
    tris, seen = [], set()
    for start in sorted(nxt):
        if start in seen:
            continue
        loop, e = [], start
        while e not in seen:
            seen.add(e)
            loop.append(e)
            e = nxt[e]
        assert e == start and len(loop) >= 3
        tris.extend(triangulate(loop))
    return tris


def coplanar(e0, e1):
    """True when two cube edges lie in one face of the cube (a chord between them lies in that face)."""
    cs = set(edge_corners(e0)) | set(edge_corners(e1))
    return any(cs <= set(q) for q in FACES)


def all_triangulations(poly):
    """All triangulations of a convex polygon given as a vertex list; each as a list of index triples."""
    if len(poly) < 3:
        return [[]]
    if len(poly) == 3:
        return [[tuple(poly)]]
    out = []
    a, b = poly[0], poly[-1]
    for k in range(1, len(poly) - 1):
        for left in all_triangulations(poly[:k + 1]):
            for right in all_triangulations(poly[k:]):
                out.append(left + [(a, poly[k], b)] + right)
    return out


def triangulate(loop):
    """Triangulation of one contour loop that avoids chords lying inside a cube face whenever possible: such a
    chord could coincide with a chord of the neighbouring cell and make four triangles meet in one edge."""
    m = len(loop)
    best, best_bad = None, None
    for tri in all_triangulations(list(range(m))):
        bad = 0
        for t in tri:
            for q in range(3):
                i, j = t[q], t[(q + 1) % 3]
                if (i - j) % m in (1, m - 1):
                    continue                       # boundary segment of the loop
                bad += coplanar(loop[i], loop[j])
        if best is None or bad < best_bad:
            best, best_bad = tri, bad
    return [tuple(loop[i] for i in t) for t in best]


def build():
    tri = [case_triangles(c) for c in range(256)]
    for c in range(256):
        used = {e for t in tri[c] for e in t}
        crossing = {e for e in range(12) if ((c >> edge_corners(e)[0]) & 1) != ((c >> edge_corners(e)[1]) & 1)}
        assert used == crossing, (c, used, crossing)
    return tri


def main():
    tri = build()
    ntri = [len(t) for t in tri]
    mx = max(ntri)
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    with open(os.path.join(root, "csrc", "mc_tables.h"), "w") as f:
        f.write("// GENERATED by tools/gen_mc_tables.py -- do not edit.\n#pragma once\n")
        f.write(f"#define HBR_MC_MAX_TRIS {mx}\n")
        f.write("static __device__ const unsigned char kMcNumTris[256] = {" + ",".join(map(str, ntri)) + "};\n")
        f.write(f"static __device__ const signed char kMcTris[256][{3 * mx}] = {{\n")
        for t in tri:
            flat = [e for tr in t for e in tr] + [-1] * (3 * mx - 3 * len(t))
            f.write("  {" + ",".join(map(str, flat)) + "},\n")
        f.write("};\n")
    with open(os.path.join(os.path.dirname(root), "oracle", "mc_tables.py"), "w") as f:
        f.write('"""GENERATED by human_body_reconstruction_b200/tools/gen_mc_tables.py -- test infrastructure copy."""\n')
        f.write(f"NUM_TRIS = {ntri!r}\n")
        f.write(f"TRIS = {[[list(x) for x in t] for t in tri]!r}\n")
    print("max triangles per cell:", mx, "total table triangles:", sum(ntri))


if __name__ == "__main__":
    main()
