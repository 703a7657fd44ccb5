"""ctypes front end of oracle/rdf_oracle.c (multi-threaded C restatement).  TEST INFRASTRUCTURE ONLY:
importable from tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs."""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, '_build', 'librdf_oracle.so')
_lib = None


def build(force=False):
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(os.path.join(_HERE, 'rdf_oracle.c')):
        subprocess.check_call(['make', '-s', '-C', _HERE, _SO])
    return _SO


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = ctypes.CDLL(_SO)
        _lib.oracle_gini_gain.restype = ctypes.c_float
    return _lib


def _p(a):
    return a.ctypes.data_as(ctypes.c_void_p) if a is not None else None


def num_threads():
    return int(lib().oracle_num_threads())


def eval_forest(forest, depth, labels_out, labels_reduce=1, filter_images=None, filter_class=None, scale=1.0,
                probs_out=None, nthreads=0):
    forest = np.ascontiguousarray(forest, dtype=np.float32)
    T, NN, E = forest.shape
    C = (E - 7) // 2
    D = int(np.log2(NN + 1))
    N, H, W = depth.shape
    assert depth.dtype == np.uint16 and labels_out.dtype == np.uint16 and depth.flags.c_contiguous and labels_out.flags.c_contiguous
    assert labels_out.shape == (N, H // labels_reduce, W // labels_reduce)
    fc = -1 if filter_images is None else int(filter_class)
    rc = lib().oracle_eval_forest(_p(forest), T, D, C, _p(depth), N, H, W, _p(filter_images), fc, _p(labels_out),
                                  _p(probs_out), int(labels_reduce), ctypes.c_float(scale), int(nthreads))
    assert rc == 0
    return labels_out


def eval_tree(tree, depth, labels_out, nthreads=0):
    tree = np.ascontiguousarray(tree, dtype=np.float32)
    NN, E = tree.shape
    C = (E - 7) // 2
    D = int(np.log2(NN + 1))
    N, H, W = depth.shape
    rc = lib().oracle_eval_tree(_p(tree), D, C, _p(depth), N, H, W, _p(labels_out), int(nthreads))
    assert rc == 0
    return labels_out


def composite(label_images, conditions, composite_out):
    conditions = np.ascontiguousarray(conditions, dtype=np.int32)
    L = len(label_images)
    h, w = composite_out.shape
    ptrs = (ctypes.c_void_p * L)(*[im.ctypes.data for im in label_images])
    rc = lib().oracle_composite(ptrs, L, w, h, _p(conditions), _p(composite_out))
    assert rc == 0
    return composite_out


def mean_shift(labels, num_labels, variances, num_rounds):
    labels = np.ascontiguousarray(labels.reshape(labels.shape[-2:]))
    h, w = labels.shape
    variances = np.ascontiguousarray(variances, dtype=np.float32)
    means = np.zeros((num_labels, 2), dtype=np.float64)
    rc = lib().oracle_mean_shift(_p(labels), w, h, int(num_labels), _p(variances), int(num_rounds), _p(means))
    assert rc == 0
    return means


def train_hist(depth, labels, nodes_by_pixel, node_slot, num_slots, offsets, thresholds, num_classes,
               f_begin=0, f_end=None, nthreads=0):
    offsets = np.ascontiguousarray(offsets, dtype=np.float32)
    thresholds = np.ascontiguousarray(thresholds, dtype=np.float32)
    node_slot = np.ascontiguousarray(node_slot, dtype=np.int32)
    nodes_by_pixel = np.ascontiguousarray(nodes_by_pixel, dtype=np.int32)
    F, NT = thresholds.shape
    N, H, W = depth.shape
    f_end = F if f_end is None else f_end
    hist = np.zeros((num_slots, F, NT + 1, num_classes), dtype=np.uint32)
    rc = lib().oracle_train_hist(_p(depth), _p(labels), _p(nodes_by_pixel), N, H, W, _p(node_slot), _p(offsets),
                                 _p(thresholds), F, NT, int(num_classes), int(f_begin), int(f_end), _p(hist), int(nthreads))
    assert rc == 0
    return hist


def gini_gain(parent, left, right):
    parent = np.ascontiguousarray(parent, dtype=np.uint64)
    left = np.ascontiguousarray(left, dtype=np.uint64)
    right = np.ascontiguousarray(right, dtype=np.uint64)
    return np.float32(lib().oracle_gini_gain(_p(parent), _p(left), _p(right), int(parent.shape[0])))
