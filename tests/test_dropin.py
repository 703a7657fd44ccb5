"""The reference scripts' import lines resolve to rdf_b200 through compat/rdf_dropin.py (no GPU needed to import)."""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

SCRIPT = r'''
import sys
sys.path.insert(0, r"%s")
import rdf_dropin
from decision_tree import *                       # src/train_model.py:1, src/test_on_saved_model.py:1
from cuda.mean_shift import MeanShift             # src/3d_bz.py
import cuda.py_nvcc_utils as py_nvcc_utils_mod    # src/run_live.py
from engine.buffer import GpuBuffer               # src/run_live_layered.py
import argparse
for name in ['DecisionTree', 'DecisionForest', 'LayeredDecisionForest', 'DecisionTreeEvaluator', 'DecisionTreeTrainer',
             'DecisionTreeDatasetConfig', 'cu_array', 'py_nvcc_utils', 'MAX_UINT16', 'np', 'json', 'Image']:
    assert name in globals(), name
assert DecisionTree.get_config(16, 4) == (65535, 65536, 15)
p = argparse.ArgumentParser(); py_nvcc_utils.add_args(p)
a = p.parse_args(['--fatbin_in', 'x']); py_nvcc_utils.config_compiler(a)       # accepted and ignored
assert MeanShift.__module__ == 'rdf_b200.mean_shift' and GpuBuffer.__module__ == 'rdf_b200.buffers'
from cuda.points_ops import *                     # src/3d_bz.py:7
from cpp_grouping import CppGrouping              # src/3d_bz.py:22
assert PointsOps.__module__ == 'rdf_b200.points_ops' and CppGrouping.__module__ == 'rdf_b200.grouping'
assert gaussian_kernel(5, 2.0).shape == (5, 5)
import cuda.bindings                              # cuda-python is still importable next to the shim
print('ok')
''' % os.path.join(ROOT, '3d-beats_b200', 'compat')


def test_reference_import_lines_resolve():
    out = subprocess.run([sys.executable, '-c', SCRIPT], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0 and out.stdout.strip().endswith('ok'), out.stderr[-3000:]
