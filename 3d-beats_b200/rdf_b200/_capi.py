"""ctypes binding of librdf_b200.so (C ABI declared in include/rdf_b200.h).

This replaces the reference's kernel loader (`py_nvcc_utils.get_module(n).get_function(kernel)`,
src/cuda/py_nvcc_utils.py:25-37): kernels are compiled ahead of time for sm_100a and fetched as C symbols.
There is NO fallback: if the library is missing the import fails, and every compute call fails without a GPU.
"""
import ctypes
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get('RDF_B200_LIB') or os.path.join(_HERE, 'librdf_b200.so')   # override: kernel-variant experiments

c_void_p = ctypes.c_void_p
c_int = ctypes.c_int
c_float = ctypes.c_float
c_size_t = ctypes.c_size_t
c_uint32 = ctypes.c_uint32
c_int64 = ctypes.c_int64

# name -> argtypes, in the order of include/rdf_b200.h
SIGNATURES = {
    'rdf_version': [],
    'rdf_last_error': [],
    'rdf_forest_create': [c_void_p, c_int, c_int, c_int, c_void_p, ctypes.POINTER(c_void_p)],
    'rdf_forest_update': [c_void_p, c_void_p, c_void_p],
    'rdf_forest_destroy': [c_void_p],
    'rdf_forest_info': [c_void_p, ctypes.POINTER(c_int), ctypes.POINTER(c_int), ctypes.POINTER(c_int), ctypes.POINTER(c_size_t)],
    'rdf_eval_forest': [c_void_p, c_void_p, c_int, c_int, c_int, c_void_p, c_int, c_void_p, c_void_p, c_int, c_float, c_void_p],
    'rdf_depth_tex_create': [c_int, c_int, c_int, ctypes.POINTER(c_void_p)],
    'rdf_depth_tex_destroy': [c_void_p],
    'rdf_depth_tex_upload': [c_void_p, c_void_p, c_int, c_void_p],
    'rdf_eval_forest_tex': [c_void_p, c_void_p, c_void_p, c_int, c_void_p, c_int, c_void_p, c_int, c_void_p],
    'rdf_eval_forest_canonical': [c_void_p, c_int, c_int, c_int, c_void_p, c_int, c_int, c_int, c_void_p, c_int, c_void_p,
                                  c_void_p, c_int, c_float, c_void_p],
    'rdf_eval_tree': [c_void_p, c_int, c_int, c_void_p, c_int, c_int, c_int, c_void_p, c_void_p],
    'rdf_eval_tree_packed': [c_void_p, c_void_p, c_int, c_int, c_int, c_void_p, c_void_p],
    'rdf_composite': [c_void_p, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p],
    'rdf_layered_run': [ctypes.POINTER(c_void_p), c_int, ctypes.POINTER(c_int), ctypes.POINTER(c_int), c_void_p, c_int, c_int,
                        ctypes.POINTER(c_void_p), c_void_p, c_int, c_void_p, c_int, c_float, c_void_p],
    'rdf_layered_run_batch': [ctypes.POINTER(c_void_p), c_int, ctypes.POINTER(c_int), ctypes.POINTER(c_int), c_void_p, c_int, c_int, c_int,
                              ctypes.POINTER(c_void_p), c_void_p, c_int, c_void_p, c_int, c_float, ctypes.c_uint, c_void_p],
    'rdf_upload_frame': [c_void_p, c_void_p, c_size_t, c_void_p],
    'rdf_mean_shift_workspace_bytes': [c_int, c_int, c_int, ctypes.POINTER(c_size_t)],
    'rdf_mean_shift': [c_void_p, c_int, c_int, c_int, c_void_p, c_int, c_void_p, c_void_p, c_size_t, c_void_p],
    'rdf_mean_shift_batch': [c_void_p, c_int, c_int, c_int, c_int, c_void_p, c_int, c_void_p, c_void_p, c_size_t, c_void_p],
    'rdf_group_hands': [c_void_p, c_int, c_int, c_float, c_void_p, c_void_p, c_void_p],
    'rdf_condition_depth': [c_void_p, c_int, c_int, c_float, c_float, c_float, c_void_p, c_float, c_void_p, c_int, c_int, c_void_p,
                            c_void_p, c_void_p],
    'rdf_deproject_points': [c_void_p, c_int, c_int, c_int, c_float, c_float, c_float, c_void_p, c_void_p],
    'rdf_transform_points': [c_int, c_void_p, c_void_p, c_void_p],
    'rdf_filter_points_by_plane': [c_int, c_float, c_void_p, c_void_p],
    'rdf_remove_missing_points': [c_int, c_void_p, c_void_p, c_void_p],
    'rdf_setup_depth_for_forest': [c_int, c_void_p, c_void_p, c_void_p],
    'rdf_zeros_to_no_pixel': [c_int, c_void_p, c_void_p],
    'rdf_shrink_image': [c_void_p, c_int, c_int, c_int, c_void_p, c_void_p],
    'rdf_stencil_by_group': [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_void_p, c_void_p],
    'rdf_scatter_groups': [c_void_p, c_int, c_void_p, c_int, c_int, c_void_p],
    'rdf_grow_groups': [c_void_p, c_int, c_int, c_void_p, c_void_p],
    'rdf_stencil_hands': [c_void_p, c_int, c_int, c_void_p, c_int, c_int, c_int, ctypes.POINTER(c_int), ctypes.POINTER(c_int),
                          c_void_p, c_void_p],
    'rdf_flip_x': [c_void_p, c_int, c_int, c_void_p, c_void_p],
    'rdf_labels_to_rgba': [c_void_p, c_int, c_int, c_void_p, c_int, c_void_p, c_void_p],
    'rdf_depth_to_rgba': [c_void_p, c_int, c_int, c_int, c_int, c_void_p, c_void_p],
    'rdf_fingertip_z': [c_void_p, c_int, c_int, ctypes.POINTER(c_int), c_int, c_int, c_void_p, c_int, c_int, c_float, c_float, c_float,
                        c_float, c_void_p, c_void_p, c_void_p, c_void_p],
    'rdf_mean_shift_fingertips': [c_void_p, c_int, c_int, c_int, c_int, c_void_p, c_int, c_void_p, c_void_p, c_size_t,
                                  ctypes.POINTER(c_int), c_int, c_int, c_void_p, c_int, c_int, c_float, c_float, c_float, c_float,
                                  c_void_p, c_void_p, c_void_p, c_void_p],
    'rdf_synth_depth': [c_void_p, c_int, c_int, c_int, c_int, c_uint32, c_int, c_void_p],
    'rdf_synth_forest': [c_void_p, c_int, c_int, c_int, c_uint32, c_void_p],
    'rdf_selftest_fastdiv': [ctypes.c_uint, c_uint32, ctypes.POINTER(ctypes.c_ulonglong), ctypes.POINTER(ctypes.c_ulonglong)],
    'rdf_train_init': [c_void_p, c_int64, c_int, c_void_p, c_void_p, c_void_p],
    'rdf_train_hist': [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_void_p, c_int, c_void_p, c_void_p, c_int, c_int,
                       c_int, c_void_p, c_void_p],
    'rdf_train_bucket_workspace_bytes': [c_int64, c_int, ctypes.POINTER(c_size_t)],
    'rdf_train_bucket': [c_void_p, c_int64, c_void_p, c_int, c_void_p, c_size_t, c_void_p],
    'rdf_train_hist_bucketed': [c_void_p, c_void_p, c_int, c_int, c_int, c_void_p, c_int, c_void_p, c_void_p, c_int, c_int, c_int,
                                c_void_p, c_void_p],
    'rdf_train_hist_bucketed_p2p': [c_void_p, c_void_p, c_int, c_int, c_int, c_void_p, c_int, c_void_p, c_void_p, c_int, c_int, c_int,
                                    c_void_p, c_int, c_void_p],
    'rdf_train_pick_candidates': [c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_int, c_void_p,
                                  c_void_p, c_void_p, c_void_p],
    'rdf_train_pick_finalize': [c_int, c_void_p, c_void_p, c_void_p, c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int,
                                c_int, c_int, c_void_p, c_void_p, c_void_p, c_void_p],
    'rdf_train_pick_best': [c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_void_p, c_void_p, c_int, c_int, c_int,
                            c_int, c_int, c_void_p, c_void_p, c_void_p, c_void_p],
    'rdf_train_next_active': [c_void_p, c_int, c_int, c_int, c_void_p, c_int, c_void_p, c_void_p, c_void_p],
    'rdf_train_advance_pixels': [c_void_p, c_void_p, c_int, c_int, c_int, c_void_p, c_int, c_int, c_int, c_void_p],
}

_lib = None


def load():
    """Load librdf_b200.so.  Raises ImportError (never falls back) when it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f'{LIB_PATH} not found: build it with 3d-beats_b200/csrc/build.sh (or __graft_entry__.build()). '
            'rdf_b200 has no CPU or JIT fallback.')
    lib = ctypes.CDLL(LIB_PATH)
    for name, argtypes in SIGNATURES.items():
        fn = getattr(lib, name)          # AttributeError if the library is stale - fail loudly
        fn.argtypes = argtypes
        fn.restype = ctypes.c_char_p if name == 'rdf_last_error' else c_int
    _lib = lib
    return lib


class RdfError(RuntimeError):
    pass


def check(rc):
    if rc == 0:
        return
    msg = load().rdf_last_error().decode('utf-8', 'replace')
    if rc == -1:
        raise ValueError(msg)
    raise RdfError(f'librdf_b200 error {rc}: {msg}')


def stream_ptr():
    """cudaStream_t of torch's current stream (kernels are launched where the caller's torch work is queued)."""
    return c_void_p(torch.cuda.current_stream().cuda_stream)


def dptr(t):
    """Device pointer of a GPUArray / torch tensor (None -> NULL)."""
    if t is None:
        return None
    tensor = getattr(t, 'tensor', t)
    if not tensor.is_cuda:
        raise ValueError('expected a CUDA tensor; rdf_b200 has no CPU path')
    if not tensor.is_contiguous():
        raise ValueError('expected a C-contiguous tensor')
    return c_void_p(tensor.data_ptr())
