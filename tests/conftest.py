import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, '3d-beats_b200')
for p in (ROOT, PKG):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line('markers', 'gpu: needs a CUDA device (run on a B200 with `-m gpu`)')


def pytest_collection_modifyitems(config, items):
    import torch
    if torch.cuda.is_available():
        return
    skip = pytest.mark.skip(reason='no CUDA device')
    for item in items:
        if 'gpu' in item.keywords:
            item.add_marker(skip)


def _missing_reference_builds():
    from oracle import ref_kernels as rk, ref_points as rp, grouping_oracle as go
    return [name for name, ok in (('oracle/_ref/libref_kernels.so', rk.available()), ('oracle/_ref/libref_points.so', rp.available()),
                                  ('oracle/_ref/libref_grouping.so', go.ref_available())) if not ok]


@pytest.fixture(autouse=True)
def _require_reference_builds(request):
    """A `-m gpu` run must FAIL, not pass vacuously, when the strongest checker - the reference's own kernels and C++ compiled
    unchanged (oracle/_ref, built by `make -C oracle ref` where /root/reference exists and shipped with the snapshot) - is absent."""
    if 'gpu' in request.keywords:
        missing = _missing_reference_builds()
        if missing:
            pytest.fail('reference builds missing: %s - run `python -c "import __graft_entry__ as g; g.build()"` where '
                        '/root/reference exists; the GPU parity tests compare against them' % ', '.join(missing), pytrace=False)


@pytest.fixture(scope='session')
def dev():
    import torch
    return torch.device('cuda:0')


def to_dev(a):
    """NumPy -> CUDA tensor, keeping uint16 bit patterns."""
    import torch
    a = np.ascontiguousarray(a)
    if a.dtype == np.uint16:
        return torch.from_numpy(a.view(np.int16)).cuda().view(torch.uint16)
    if a.dtype == np.uint32:
        return torch.from_numpy(a.view(np.int32)).cuda().view(torch.uint32)
    if a.dtype == np.uint64:
        return torch.from_numpy(a.view(np.int64)).cuda().view(torch.uint64)
    return torch.from_numpy(a).cuda()


def to_np(t):
    import torch
    sv = {torch.uint16: (torch.int16, np.uint16), torch.uint32: (torch.int32, np.uint32), torch.uint64: (torch.int64, np.uint64)}
    if t.dtype in sv:
        s, u = sv[t.dtype]
        return t.view(s).cpu().numpy().view(u)
    return t.cpu().numpy()


def filled_u16(shape, value=65535):
    import torch
    return torch.full(shape, np.uint16(value).view(np.int16).item(), dtype=torch.int16, device='cuda').view(torch.uint16)
