"""Drop-in shim for the reference's scripts (run_live.py, run_live_layered.py, test_on_saved_model.py, train_model.py; for the
live scripts and the product loop src/3d_bz.py also `cuda.points_ops` -> rdf_b200.points_ops (PointsOps with the reference's
pycuda call forms, gaussian_kernel), `calibrated_plane` -> rdf_b200.calibrated_plane and `cpp_grouping` -> rdf_b200.grouping
(CppGrouping)).

`import rdf_dropin` BEFORE the scripts' own imports: it registers the B200 implementation under the top-level module
names the reference uses (`decision_tree`, `cuda.mean_shift`, `cuda.py_nvcc_utils`, `engine.buffer`), so that
`from decision_tree import *`, `from cuda.mean_shift import MeanShift`, `import cuda.py_nvcc_utils as py_nvcc_utils` and
`from engine.buffer import GpuBuffer` (src/run_live_layered.py:6-16, src/train_model.py:1-12) resolve to rdf_b200.
Modules are registered in sys.modules rather than laid out as an on-disk `cuda/` package, because a top-level `cuda`
directory would shadow cuda-python's `cuda.bindings` (hard part 5); other `cuda.*` / `engine.*` submodules keep
resolving through the original packages' __path__ when those exist.
"""
import importlib
import os
import sys
import types

_PKG = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if _PKG not in sys.path:
    sys.path.insert(0, _PKG)

import rdf_b200.buffers as _buffers  # noqa: E402
import rdf_b200.calibrated_plane as _calibrated_plane  # noqa: E402
import rdf_b200.decision_tree as _decision_tree  # noqa: E402
import rdf_b200.grouping as _grouping  # noqa: E402
import rdf_b200.mean_shift as _mean_shift  # noqa: E402
import rdf_b200.points_ops as _points_ops  # noqa: E402
import rdf_b200.py_nvcc_utils as _py_nvcc_utils  # noqa: E402


def _package(name):
    """Existing package of that name (keeps its __path__, e.g. cuda-python's `cuda` or the reference's `engine`), else a stub."""
    try:
        return importlib.import_module(name)
    except Exception:
        mod = types.ModuleType(name)
        mod.__path__ = []
        sys.modules[name] = mod
        return mod


def install():
    sys.modules['decision_tree'] = _decision_tree
    cuda_pkg = _package('cuda')
    cuda_pkg.mean_shift = _mean_shift
    cuda_pkg.py_nvcc_utils = _py_nvcc_utils
    sys.modules['cuda.mean_shift'] = _mean_shift
    sys.modules['cuda.py_nvcc_utils'] = _py_nvcc_utils
    cuda_pkg.points_ops = _points_ops                  # from cuda.points_ops import *  (src/3d_bz.py:7)
    sys.modules['cuda.points_ops'] = _points_ops
    sys.modules['cpp_grouping'] = _grouping            # from cpp_grouping import CppGrouping  (src/3d_bz.py:22)
    sys.modules['calibrated_plane'] = _calibrated_plane   # from calibrated_plane import *  (src/run_live.py:8, src/run_live_layered.py:8)
    engine_pkg = _package('engine')
    engine_pkg.buffer = _buffers                       # GpuBuffer(shape, dtype).cu()  (src/engine/buffer.py:10-39)
    sys.modules['engine.buffer'] = _buffers


install()
