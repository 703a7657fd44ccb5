"""Hand grouping (connected components on the 1/8-resolution depth image): CPU restatement + the reference itself.
TEST INFRASTRUCTURE ONLY (tests/, smoke(), bench legs).

Pinned: `ref_make_groups` runs the REFERENCE's own C++ (src/cpp_grouping/grouping.cpp compiled unchanged by oracle/Makefile
into oracle/_ref/libref_grouping.so); tests/test_grouping_oracle.py checks the NumPy restatement below against it on CPU.
"""
import ctypes
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
REF_SO = os.path.join(_HERE, '_ref', 'libref_grouping.so')
_ref = None


def ref_available():
    return os.path.exists(REF_SO)


def ref_make_groups(img, pct_thresh):
    """The reference's CppGrouping.make_groups (src/cpp_grouping/grouping.cpp:80-191, binding cpp_grouping.pyx:15-26).
    img uint16[h,w].  Returns (coords int32[n,3] = (y, x, group) in the reference's BFS order, g_info float32[2,3] =
    (size, centroid x, centroid y) for the right (row 0) and left (row 1) group; centroid entries are unspecified when size is 0)."""
    global _ref
    if _ref is None:
        _ref = ctypes.CDLL(REF_SO)
    img = np.ascontiguousarray(img, dtype=np.uint16)
    h, w = img.shape
    coords = np.zeros((h * w, 3), dtype=np.int32)
    g_info = np.zeros((2, 3), dtype=np.float32)
    _ref.ref_make_groups(img.ctypes.data_as(ctypes.c_void_p), ctypes.c_int(w), ctypes.c_int(h), coords.ctypes.data_as(ctypes.c_void_p),
                         g_info.ctypes.data_as(ctypes.c_void_p), ctypes.c_float(pct_thresh))
    n = int(g_info[0, 0]) + int(g_info[1, 0])
    return coords[:n], g_info


def stencil_from_coords(coords, h, w):
    """write_pixel_groups_to_stencil_image (src/3d_bz.py:243-250): group id per pixel, 0 elsewhere."""
    out = np.zeros((h, w), dtype=np.uint16)
    out[coords[:, 0], coords[:, 1]] = coords[:, 2].astype(np.uint16)
    return out


def make_groups(img, pct_thresh):
    """NumPy restatement of grouping.cpp:80-191.  4-connected components of the non-zero pixels in raster order of their first
    pixel; components with size / (w*h) <= pct_thresh (fp32) are dropped (:137); centroid = int sums / size in fp32 (:139-148);
    centroid x < w/2 -> candidate right group else left (:150-163); per side the strictly largest wins, so among equal sizes the
    component met first in raster order (:151,157).  Returns (stencil uint16[h,w] with 1 = right, 2 = left, g_info float32[2,3])."""
    img = np.asarray(img)
    h, w = img.shape
    fg = img != 0
    labels = np.full((h, w), -1, dtype=np.int64)
    comps = []                                   # (first raster index, pixel list) in raster order of the first pixel
    for y in range(h):
        for x in range(w):
            if not fg[y, x] or labels[y, x] >= 0:
                continue
            cid = len(comps)
            stack = [(y, x)]
            labels[y, x] = cid
            pix = []
            while stack:
                cy, cx = stack.pop()
                pix.append((cy, cx))
                for dy, dx in ((-1, 0), (1, 0), (0, -1), (0, 1)):
                    ny, nx = cy + dy, cx + dx
                    if 0 <= ny < h and 0 <= nx < w and fg[ny, nx] and labels[ny, nx] < 0:
                        labels[ny, nx] = cid
                        stack.append((ny, nx))
            comps.append(pix)
    best = {0: (0, -1), 1: (0, -1)}              # side -> (size, component id)
    g_info = np.zeros((2, 3), dtype=np.float32)
    for cid, pix in enumerate(comps):
        n = len(pix)
        if np.float32(n) / np.float32(w * h) <= np.float32(pct_thresh):
            continue
        sy = sum(p[0] for p in pix)
        sx = sum(p[1] for p in pix)
        c_y = np.float32(sy) / np.float32(n)
        c_x = np.float32(sx) / np.float32(n)
        side = 0 if c_x < np.float32(w) / np.float32(2.0) else 1
        if n > best[side][0]:
            best[side] = (n, cid)
            g_info[side] = (n, c_x, c_y)
    stencil = np.zeros((h, w), dtype=np.uint16)
    for side in (0, 1):
        if best[side][1] >= 0:
            stencil[labels == best[side][1]] = side + 1
    return stencil, g_info
