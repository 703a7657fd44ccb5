"""GPU parity of the device hand grouping against the REFERENCE's own C++ (oracle/_ref/libref_grouping.so) and the NumPy
restatement: stencil image bit-exact, group sizes and centroids identical."""
import numpy as np
import pytest

from conftest import to_dev, to_np

pytestmark = pytest.mark.gpu


def _ours(img, thresh):
    import torch
    from rdf_b200.grouping import CppGrouping
    h, w = img.shape
    st = torch.full((h, w), 77, dtype=torch.int16, device='cuda').view(torch.uint16)
    gi = torch.full((2, 3), -1.0, dtype=torch.float32, device='cuda')
    CppGrouping().make_groups_cu(to_dev(img), st, gi, thresh)
    torch.cuda.synchronize()
    return to_np(st), gi.cpu().numpy()


@pytest.mark.parametrize('h,w', [(60, 106), (30, 53), (120, 128), (7, 9), (4, 300), (1, 500), (16, 1000), (200, 33), (33, 65), (64, 32)])
@pytest.mark.parametrize('seed', range(4))
def test_matches_reference_cpp_and_oracle(h, w, seed):
    from oracle import grouping_oracle as go
    from test_grouping_oracle import blob_image
    img = blob_image(h, w, 100 * h + seed, density=0.02 if seed % 2 else 0.0)
    thresh = 0.005
    st, g = _ours(img, thresh)
    st_o, g_o = go.make_groups(img, thresh)
    assert np.array_equal(st, st_o) and np.array_equal(g, g_o)
    if True:                       # the reference build is required under -m gpu (tests/conftest.py)
        coords, g_ref = go.ref_make_groups(img, thresh)
        assert np.array_equal(st, go.stencil_from_coords(coords, h, w))
        for side in (0, 1):
            assert g[side, 0] == g_ref[side, 0]
            if g_ref[side, 0] > 0:
                assert np.array_equal(g[side], g_ref[side])


def test_ties_threshold_empty_and_host_signature():
    from oracle import grouping_oracle as go
    from rdf_b200.grouping import CppGrouping
    img = np.zeros((20, 40), np.uint16)
    img[2:5, 2:5] = 7
    img[10:13, 3:6] = 7                                            # same size, met second: loses the tie
    img[5:9, 30:34] = 9
    st, g = _ours(img, 0.0)
    assert st[3, 3] == 1 and st[11, 4] == 0 and st[6, 31] == 2 and g[0, 0] == 9 and g[1, 0] == 16
    st, g = _ours(np.zeros((16, 16), np.uint16), 0.01)
    assert not st.any() and not g.any()
    st, g = _ours(np.full((16, 16), 5, np.uint16), 0.01)           # one component, centroid x = 7.5 < w/2 = 8 -> right group
    assert (st == 1).all() and g[0, 0] == 256 and g[0, 1] == 7.5 and g[1, 0] == 0
    half = np.zeros((16, 17), np.uint16)
    half[:, 8] = 3                                                 # centroid x = 8 < 8.5 -> right; x = 9 would be left
    half[:, 10] = 3
    st, g = _ours(half, 0.01)
    assert (st[:, 8] == 1).all() and (st[:, 10] == 2).all()
    # host-array signature of the reference binding (cpp_grouping.pyx:15)
    coords = np.zeros((img.size, 3), np.int32)
    g_info = np.zeros((2, 3), np.float32)
    n = CppGrouping().make_groups(img, coords, g_info, 0.0)
    st_o, g_o = go.make_groups(img, 0.0)
    assert n == 25 and np.array_equal(go.stencil_from_coords(coords[:n], 20, 40), st_o) and np.array_equal(g_info, g_o)


def test_too_large_image_is_refused():
    import torch
    from rdf_b200.grouping import CppGrouping
    img = torch.zeros((200, 200), dtype=torch.int16, device='cuda').view(torch.uint16)
    with pytest.raises(Exception):
        CppGrouping().make_groups_cu(img, torch.zeros_like(img), torch.zeros((2, 3), device='cuda'), 0.1)
