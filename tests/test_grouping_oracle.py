"""CPU: the NumPy restatement of the hand grouping against the REFERENCE's own C++ (src/cpp_grouping/grouping.cpp compiled
unchanged into oracle/_ref/libref_grouping.so) on seeded images, including ties, the size threshold and the empty image."""
import numpy as np
import pytest

from oracle import grouping_oracle as go

pytestmark = pytest.mark.skipif(not go.ref_available(), reason='oracle/_ref/libref_grouping.so not built (needs /root/reference at build time)')


def blob_image(h, w, seed, n_blobs=6, density=0.0):
    rng = np.random.default_rng(seed)
    img = np.zeros((h, w), np.uint16)
    yy, xx = np.mgrid[0:h, 0:w]
    for _ in range(n_blobs):
        cy, cx = rng.integers(0, h), rng.integers(0, w)
        ry, rx = rng.integers(2, max(3, h // 4)), rng.integers(2, max(3, w // 4))
        img[((yy - cy) / ry) ** 2 + ((xx - cx) / rx) ** 2 <= 1.0] = rng.integers(500, 4000)
    if density > 0:
        img[rng.random((h, w)) < density] = 1234                 # salt noise: many tiny components
    img[rng.random((h, w)) < 0.03] = 0                            # holes
    return img


def check(img, thresh):
    h, w = img.shape
    coords, g_ref = go.ref_make_groups(img, thresh)
    st_ref = go.stencil_from_coords(coords, h, w)
    st, g = go.make_groups(img, thresh)
    assert np.array_equal(st, st_ref)
    assert g[0, 0] == g_ref[0, 0] and g[1, 0] == g_ref[1, 0]
    for side in (0, 1):
        if g_ref[side, 0] > 0:                                    # centroids are unspecified in the reference for empty groups
            assert np.array_equal(g[side], g_ref[side])
    return st_ref, g_ref


@pytest.mark.parametrize('seed', range(8))
def test_restatement_matches_reference_cpp(seed):
    img = blob_image(60, 106, seed, density=0.02 if seed % 2 else 0.0)           # 848x480 shrunk by 8 (src/3d_bz.py:49-60)
    st, g = check(img, 0.005)
    assert st.max() >= 1


def test_ties_threshold_and_empty():
    img = np.zeros((20, 40), np.uint16)
    img[2:5, 2:5] = 7          # 9 px, right half, met first
    img[10:13, 3:6] = 7        # 9 px, right half, met second: loses the tie
    img[5:9, 30:34] = 9        # 16 px, left half
    img[15:17, 25:27] = 9      # 4 px, left half
    st, g = check(img, 0.0)
    assert st[3, 3] == 1 and st[11, 4] == 0 and st[6, 31] == 2 and st[15, 25] == 0
    assert g[0, 0] == 9 and g[1, 0] == 16
    st, g = check(img, 10 / 800.0)                                 # 9 px groups fall below the threshold
    assert g[0, 0] == 0 and g[1, 0] == 16
    check(np.zeros((16, 16), np.uint16), 0.01)
    check(np.full((16, 16), 5, np.uint16), 0.01)                   # one component, centroid exactly at w/2 -> left (not <)
