"""Host-side mirror of the reference's RDF library (src/decision_tree.py) over librdf_b200.so.

Same class / method names, positional argument order and semantics as the reference so that run_live.py,
run_live_layered.py, test_on_saved_model.py and train_model.py keep working (`from decision_tree import *`,
SURVEY section 8b).  Differences, all behind the same call surface:
  * kernels are ahead-of-time sm_100a code reached through a C ABI (no pycuda JIT, no CPU fallback);
  * forests keep a packed device shadow (32-byte node headers) that is refreshed lazily when `forest_cu` changed;
  * LayeredDecisionForest.run is one fused launch; training data stays uncompressed in HBM (no nvcomp).
"""
import ctypes
import json
import os
import os.path
from pathlib import Path

import numpy as np
import torch

from . import _capi
from . import buffers as cu_array          # `cu_array.GPUArray(...)`, `cu_array.to_gpu(...)` as in the reference
from . import py_nvcc_utils
from .buffers import GpuBuffer, GPUArray, as_gpuarray

try:                                        # callers rely on `Image` re-exported by `from decision_tree import *`
    from PIL import Image
except Exception:                           # pragma: no cover - PIL is optional for the hot path itself
    Image = None

MAX_UINT16 = np.uint16(65535)               # src/util.py:43
MAX_THREADS_PER_BLOCK = 1024                # kept for callers that import it (src/decision_tree.py:16)


def sizeof_fmt(num, suffix='B'):
    for unit in ['', 'K', 'M', 'G', 'T', 'P', 'E', 'Z']:
        if abs(num) < 1024.0:
            return "%3.1f %s%s" % (num, unit, suffix)
        num /= 1024.0
    return "%.1f%s%s" % (num, 'Y', suffix)


def _version_of(arr):
    """torch bumps `_version` on every in-place write (fill_, copy_): a free dirty flag for the packed shadow."""
    return arr.tensor._version


class DecisionTree():
    """src/decision_tree.py:124-144."""

    def __init__(self, max_depth, num_classes):
        self.max_depth = max_depth
        self.num_classes = num_classes
        self.TOTAL_TREE_NODES, self.MAX_LEAF_NODES, self.TREE_NODE_ELS = DecisionTree.get_config(max_depth, num_classes)
        self.tree_out_cu = GPUArray((self.TOTAL_TREE_NODES, self.TREE_NODE_ELS), dtype=np.float32)
        self.tree_out_cu.fill(np.float32(0.))
        self._handle = None
        self._packed_version = None

    def handle(self):
        """Packed one-tree shadow of tree_out_cu (re-packed when tree_out_cu was written), for the fast evaluation path."""
        lib = _capi.load()
        ver = _version_of(self.tree_out_cu)
        if self._handle is None:
            h = ctypes.c_void_p()
            _capi.check(lib.rdf_forest_create(_capi.dptr(self.tree_out_cu), 1, self.max_depth, self.num_classes, _capi.stream_ptr(),
                                              ctypes.byref(h)))
            self._handle = h
        elif ver != self._packed_version:
            _capi.check(lib.rdf_forest_update(self._handle, _capi.dptr(self.tree_out_cu), _capi.stream_ptr()))
        self._packed_version = ver
        return self._handle

    def invalidate(self):
        """Force a re-pack on next use: needed when tree_out_cu was written by something torch does not see (the training
        kernels of librdf_b200 write it through raw pointers)."""
        self._packed_version = None

    def __del__(self):
        if getattr(self, '_handle', None) is not None:
            try:
                _capi.load().rdf_forest_destroy(self._handle)
            except Exception:
                pass
            self._handle = None

    @staticmethod
    def get_config(max_depth, num_classes):
        TOTAL_TREE_NODES = (2 ** max_depth) - 1       # nodes of the complete tree
        MAX_LEAF_NODES = 2 ** max_depth               # children of the deepest level
        TREE_NODE_ELS = 7 + (num_classes * 2)         # ux,uy,vx,vy,thresh,l_next,r_next,l_pdf[C],r_pdf[C]
        return (TOTAL_TREE_NODES, MAX_LEAF_NODES, TREE_NODE_ELS)


class DecisionForest():
    """src/decision_tree.py:146-168.  `forest_cu` stays the canonical float32[T,2^D-1,7+2C] array; `handle()` returns the
    packed shadow used by the kernels, re-packed automatically after `forest_cu` was written."""

    @staticmethod
    def load(model_filename):
        forest_cpu = np.load(model_filename)
        num_trees = forest_cpu.shape[0]
        tree_depth = int(np.log2(forest_cpu.shape[1] + 1))
        num_classes = (forest_cpu.shape[2] - 7) // 2
        f = DecisionForest(num_trees, tree_depth, num_classes)
        f.forest_cu.set(np.ascontiguousarray(forest_cpu, dtype=np.float32))
        return f

    def __init__(self, num_trees, max_depth, num_classes):
        self.num_trees = num_trees
        self.max_depth = max_depth
        self.num_classes = num_classes
        self.TOTAL_TREE_NODES, self.MAX_LEAF_NODES, self.TREE_NODE_ELS = DecisionTree.get_config(max_depth, num_classes)
        self.forest_cu = GPUArray((self.num_trees, self.TOTAL_TREE_NODES, self.TREE_NODE_ELS), dtype=np.float32)
        self.forest_cu.fill(np.float32(0.))
        self._handle = None
        self._packed_version = None
        self._packed_ptr = None

    def handle(self):
        lib = _capi.load()
        ver, ptr = _version_of(self.forest_cu), self.forest_cu.ptr
        if self._handle is not None and ptr != self._packed_ptr:
            self._destroy()
        if self._handle is None:
            h = ctypes.c_void_p()
            _capi.check(lib.rdf_forest_create(_capi.dptr(self.forest_cu), self.num_trees, self.max_depth, self.num_classes,
                                              _capi.stream_ptr(), ctypes.byref(h)))
            self._handle = h
        elif ver != self._packed_version:
            _capi.check(lib.rdf_forest_update(self._handle, _capi.dptr(self.forest_cu), _capi.stream_ptr()))
        self._packed_version, self._packed_ptr = ver, ptr
        return self._handle

    def invalidate(self):
        """Force a re-pack on next use (needed only if forest_cu was written outside torch, e.g. by a foreign kernel)."""
        self._packed_version = None

    def _destroy(self):
        if self._handle is not None:
            try:
                _capi.load().rdf_forest_destroy(self._handle)
            except Exception:
                pass
            self._handle = None

    def __del__(self):
        self._destroy()


class DecisionTreeEvaluator():
    """src/decision_tree.py:267-347."""

    def __init__(self):
        self._lib = _capi.load()

    def get_labels(self, tree, depth_images_in, labels_out):
        depth_images_in = as_gpuarray(depth_images_in)
        labels_out = as_gpuarray(labels_out)
        num_images, dim_y, dim_x = depth_images_in.shape
        assert labels_out.shape == (num_images, dim_y, dim_x)
        assert depth_images_in.dtype == np.uint16 and labels_out.dtype == np.uint16
        if hasattr(tree, 'handle'):
            tree.invalidate()           # the trainer writes tree_out_cu through raw pointers: always re-pack (60 -> 64 B per node)
            _capi.check(self._lib.rdf_eval_tree_packed(tree.handle(), _capi.dptr(depth_images_in), num_images, dim_x, dim_y,
                                                       _capi.dptr(labels_out), _capi.stream_ptr()))
        else:
            _capi.check(self._lib.rdf_eval_tree(_capi.dptr(tree.tree_out_cu), tree.max_depth, tree.num_classes,
                                                _capi.dptr(depth_images_in), num_images, dim_x, dim_y, _capi.dptr(labels_out),
                                                _capi.stream_ptr()))

    def get_labels_forest(self, forest, depth_images_in, labels_out, labels_reduce=1, filter_images=None,
                          filter_images_class=None, scale_factor=1., probs_out=None):
        depth_images_in = as_gpuarray(depth_images_in)
        labels_out = as_gpuarray(labels_out)
        num_images, dim_y, dim_x = depth_images_in.shape
        assert labels_out.shape == (num_images, dim_y // labels_reduce, dim_x // labels_reduce)
        assert depth_images_in.dtype == np.uint16 and labels_out.dtype == np.uint16
        if filter_images is not None:                      # the reference tests truthiness; `is not None` is the intent
            filter_images = as_gpuarray(filter_images)
            assert filter_images_class is not None
            assert filter_images.shape == labels_out.shape
            assert filter_images.dtype == np.uint16
        if probs_out is not None:
            probs_out = as_gpuarray(probs_out)
            assert probs_out.shape == labels_out.shape + (forest.num_classes,) and probs_out.dtype == np.float32
        fclass = int(filter_images_class) if filter_images is not None else -1
        if forest.num_trees <= 8:
            _capi.check(self._lib.rdf_eval_forest(forest.handle(), _capi.dptr(depth_images_in), num_images, dim_x, dim_y,
                                                  _capi.dptr(filter_images), fclass, _capi.dptr(labels_out),
                                                  _capi.dptr(probs_out), int(labels_reduce), float(scale_factor),
                                                  _capi.stream_ptr()))
        else:
            _capi.check(self._lib.rdf_eval_forest_canonical(_capi.dptr(forest.forest_cu), forest.num_trees, forest.max_depth,
                                                            forest.num_classes, _capi.dptr(depth_images_in), num_images,
                                                            dim_x, dim_y, _capi.dptr(filter_images), fclass,
                                                            _capi.dptr(labels_out), _capi.dptr(probs_out), int(labels_reduce),
                                                            float(scale_factor), _capi.stream_ptr()))

    def make_composite_labels_image(self, images, dim_x, dim_y, labels_decision_tree, composite_image):
        images = as_gpuarray(images)
        labels_decision_tree = as_gpuarray(labels_decision_tree)
        composite_image = as_gpuarray(composite_image)
        _capi.check(self._lib.rdf_composite(_capi.dptr(images), int(images.shape[0]), int(dim_x), int(dim_y),
                                            _capi.dptr(labels_decision_tree), _capi.dptr(composite_image), _capi.stream_ptr()))


class LayeredDecisionForest():
    """src/decision_tree.py:171-264.  `run` is a single fused launch (rdf_layered_run)."""

    @staticmethod
    def load(config_filename, depth_dims, labels_reduce=1):
        cfg = json.loads(open(config_filename).read())
        cfg['root'] = str(Path(config_filename).parent)                   # model paths are relative to the config file
        return LayeredDecisionForest(cfg, depth_dims, labels_reduce)

    def __init__(self, cfg, depth_dims, labels_reduce):
        self.eval = DecisionTreeEvaluator()
        self.depth_dims = tuple(depth_dims)                       # (y, x)
        self.labels_reduce = labels_reduce
        self.labels_dims = tuple(d // labels_reduce for d in self.depth_dims[:2])

        # one (forest, filter layer, filter class) triple per layer; a layer without 'filter_model' is ungated.  (The reference's
        # second membership test is a constant-true string literal, SURVEY note N2, so only 'filter_model' decides.)
        root = cfg.get('root', '')
        self.m = [(spec['model'] if isinstance(spec['model'], DecisionForest) else DecisionForest.load(os.path.join(root, spec['model'])),
                   spec.get('filter_model'), spec['filter_model_class'] if 'filter_model' in spec else None)
                  for spec in cfg['layers']]
        self.num_models = L = len(self.m)

        # per-layer label maps and the device table of their addresses (make_composite_labels_image's first argument)
        self.label_images = [GpuBuffer(self.labels_dims, dtype=np.uint16) for _ in range(L)]
        self.labels_images_ptrs_cu = GpuBuffer((L,), dtype=np.int64)
        self.labels_images_ptrs_cu.cu().set(np.fromiter((b.cu().ptr for b in self.label_images), dtype=np.int64, count=L))

        # condition rows are (0, composite id) or (1, row offset of the next layer's block)   (src/decision_tree.py:209-220)
        conditions = np.asarray(cfg['conditions'], dtype=np.int32).reshape(-1, 2)
        self.labels_conditions_cu = GpuBuffer(conditions.shape, dtype=np.int32)
        self.labels_conditions_cu.cu().set(conditions)
        self.num_layered_classes = int(conditions[conditions[:, 0] == 0, 1].max())

        colors = np.asarray(cfg['label_colors'], dtype=np.uint8)
        assert colors.shape == (self.num_layered_classes, 4)
        self.label_colors = GpuBuffer(colors.shape, dtype=np.uint8)
        self.label_colors.cu().set(colors)

        # ctypes views of the same tables for the fused launch
        self._c_filter_model = (ctypes.c_int * L)(*[(-1 if fm is None else int(fm)) for _, fm, _ in self.m])
        self._c_filter_class = (ctypes.c_int * L)(*[(-1 if fc is None else int(fc)) for _, _, fc in self.m])
        self._c_label_ptrs = (ctypes.c_void_p * L)(*[b.cu().ptr for b in self.label_images])
        self._n_cond = int(conditions.shape[0])

    def run(self, depth_image, labels_image, scale_factor=1., composite_flip_x=None, label_images=None):
        """src/decision_tree.py:233-264 in one launch.  Two keyword extensions for the live product's per-hand loop
        (src/3d_bz.py:387-446), which runs this once per hand: with composite_flip_x = a sequence of N flags, depth_image holds N
        images (uint16[N,H,W]), labels_image N composite images and label_images = per-layer buffers uint16[N,h,w] (other than
        self.label_images, which hold one image); image n's composite is written mirrored in x when its flag is set (the
        reference's labels_image_2.set + flip_x after the left hand's run)."""
        depth = as_gpuarray(depth_image)
        labels = as_gpuarray(labels_image)
        assert depth.dtype == np.uint16 and labels.dtype == np.uint16
        N = 1 if composite_flip_x is None else len(composite_flip_x)
        assert depth.size == N * self.depth_dims[0] * self.depth_dims[1], 'depth image dims'
        assert labels.size == N * self.labels_dims[0] * self.labels_dims[1], 'labels image dims'
        L = self.num_models
        handles = (ctypes.c_void_p * L)(*[m.handle().value for m, _, _ in self.m])
        label_ptrs = self._c_label_ptrs
        if label_images is not None:
            assert len(label_images) == L and all(as_gpuarray(i).size == N * self.labels_dims[0] * self.labels_dims[1] for i in label_images)
            label_ptrs = (ctypes.c_void_p * L)(*[as_gpuarray(i).ptr for i in label_images])
        else:
            assert N == 1, 'a batch needs its own per-layer label buffers (label_images=...)'
        mask = 0 if composite_flip_x is None else sum(1 << n for n, f in enumerate(composite_flip_x) if f)
        _capi.check(self.eval._lib.rdf_layered_run_batch(
            handles, L, self._c_filter_model, self._c_filter_class, _capi.dptr(depth), N, self.depth_dims[1], self.depth_dims[0],
            label_ptrs, _capi.dptr(self.labels_conditions_cu.cu()), self._n_cond, _capi.dptr(labels),
            int(self.labels_reduce), float(scale_factor), mask, _capi.stream_ptr()))

    def run_unfused(self, depth_image, labels_image, scale_factor=1.):
        """The reference's launch sequence verbatim (fills, one forest eval per layer, composite): kept for parity tests
        of get_labels_forest(filter_images=...) and make_composite_labels_image."""
        labels_image = as_gpuarray(labels_image)
        depth_image = as_gpuarray(depth_image)
        labels_image.fill(MAX_UINT16)
        for i in self.label_images:
            i.cu().fill(MAX_UINT16)
        depth_img_dims = (1,) + self.depth_dims
        label_img_dims = (1,) + self.labels_dims
        for i in range(self.num_models):
            m, filter_model, filter_model_class = self.m[i]
            self.eval.get_labels_forest(
                m, depth_image.reshape(depth_img_dims), self.label_images[i].cu().reshape(label_img_dims),
                labels_reduce=self.labels_reduce,
                filter_images=self.label_images[filter_model].cu().reshape(label_img_dims) if (filter_model is not None) else None,
                filter_images_class=filter_model_class, scale_factor=scale_factor)
        self.eval.make_composite_labels_image(self.labels_images_ptrs_cu.cu(), self.labels_dims[1], self.labels_dims[0],
                                              self.labels_conditions_cu.cu(), labels_image.reshape(label_img_dims))


# ----------------------------------------------------------------------------------------------------------------
# dataset (src/decision_tree.py:21-122) - blocks stay uncompressed in HBM (the reference keeps them nvcomp-compressed)
# ----------------------------------------------------------------------------------------------------------------
class ResidentBlocks:
    """Replacement for CompressedBlocksStatic/Dynamic (src/compressed_blocks.py:9-208): 180 GB of HBM holds the blocks
    uncompressed, so `get_block_cu` is a device-to-device copy and `write_block` its inverse."""

    def __init__(self, num_blocks, block_shape, dtype, fill_block=None):
        self.num_blocks = num_blocks
        self.block_shape = tuple(block_shape)
        self.store = GPUArray((num_blocks,) + self.block_shape, dtype=dtype)
        if fill_block is not None:
            host = np.zeros(self.block_shape, dtype=dtype)
            for i in range(num_blocks):
                fill_block(i, host)
                self.store[i].set(host)

    def block(self, i):
        assert i < self.num_blocks
        return self.store[i]

    def get_block_cu(self, block_num, arr_out):
        arr_out = as_gpuarray(arr_out)
        assert arr_out.shape == self.block_shape
        arr_out.set(self.block(block_num))

    get_block = get_block_cu

    def write_block(self, block_num, arr_in):
        self.block(block_num).set(as_gpuarray(arr_in))


def _dataset_config(dataset_dir):
    """config.json of a dataset directory (`dataset_dir` ends with a separator, as in the reference's call sites)"""
    with open(dataset_dir + 'config.json') as fh:
        return json.load(fh)


class DecisionTreeDatasetConfig():
    """src/decision_tree.py:21-122: a directory of `%08d_depth.png` / `%08d_labels.png` pairs described by config.json, served
    as blocks of `images_per_block` uint16 images that live uncompressed in HBM (ResidentBlocks)."""

    @staticmethod
    def multiple(dataset_dir, images):
        """One dataset per (num_images, images_per_block, imgs_name) request (src/decision_tree.py:24-44)."""
        available = _dataset_config(dataset_dir)['num_images']
        assert sum(req[0] for req in images) <= available
        # The reference shuffles the whole index range here and then does not use it (src/decision_tree.py:30-31): every dataset
        # draws its own images below, so train and test sets are independent and may overlap.  The draw is kept so that a
        # seeded np.random stream reaches the per-dataset shuffles and the proposal generators in the reference's state.
        np.random.shuffle(list(range(available)))
        return tuple(DecisionTreeDatasetConfig(dataset_dir, num_images=n, images_per_block=per_block or n, imgs_name=name)
                     for n, per_block, name in images)

    def __init__(self, dataset_dir, num_images=0, images_per_block=0, imgs_name='data0'):
        self.dataset_dir = dataset_dir
        self.cfg = cfg = _dataset_config(dataset_dir)
        self.imgs_name = imgs_name
        self.img_dims = tuple(cfg['img_dims'])                                    # (x, y)
        self.id_to_color = {0: np.zeros(4, dtype=np.uint8)}                       # id 0 = unlabelled, transparent black
        self.id_to_color.update((int(i), np.array(c, dtype=np.uint8)) for i, c in cfg['id_to_color'].items())
        self.total_available_images = cfg['num_images']
        self.num_images = num_images
        if not num_images:                                                        # descriptor only (colours, dims): nothing is loaded
            return
        self.images_per_block = images_per_block or num_images
        assert num_images % self.images_per_block == 0
        self.num_image_blocks = num_images // self.images_per_block

        order = list(range(self.total_available_images))
        np.random.shuffle(order)                                                  # the reference's draw (a list, not an array)
        self._image_ids = order[:num_images]
        block_shape = (self.images_per_block, self.img_dims[1], self.img_dims[0])
        self.depth_blocks = ResidentBlocks(self.num_image_blocks, block_shape, np.uint16, lambda b, out: self._read_block(b, out, 'depth'))
        self.labels_blocks = ResidentBlocks(self.num_image_blocks, block_shape, np.uint16, lambda b, out: self._read_block(b, out, 'labels'))

    def _read_block(self, block, out, kind):
        """decode the PNGs of one block into the host staging array `out` (uint16[images_per_block, H, W])"""
        first = block * self.images_per_block
        for j, image_id in enumerate(self._image_ids[first:first + self.images_per_block]):
            out[j] = np.asarray(Image.open(f'{self.dataset_dir}/{image_id:08d}_{kind}.png')).astype(np.uint16)

    @classmethod
    def from_arrays(cls, depth, labels, num_classes, images_per_block=0, imgs_name='mem'):
        """In-memory dataset (no PNG directory): depth/labels uint16[N,H,W] as NumPy arrays or device arrays."""
        self = cls.__new__(cls)
        N, H, W = depth.shape
        self.dataset_dir = None
        self.cfg = {'img_dims': [W, H], 'num_images': N}
        self.imgs_name = imgs_name
        self.img_dims = (W, H)
        self.id_to_color = {i: np.array([(53 * i) % 256, (97 * i) % 256, (29 * i + 60) % 256, 255 if i else 0], dtype=np.uint8)
                            for i in range(num_classes)}
        self.total_available_images = N
        self.num_images = N
        self.images_per_block = images_per_block or N
        assert N % self.images_per_block == 0
        self.num_image_blocks = N // self.images_per_block
        block_shape = (self.images_per_block, H, W)
        self.depth_blocks = ResidentBlocks(self.num_image_blocks, block_shape, np.uint16)
        self.labels_blocks = ResidentBlocks(self.num_image_blocks, block_shape, np.uint16)
        for store, src in ((self.depth_blocks, depth), (self.labels_blocks, labels)):
            if isinstance(src, np.ndarray):
                store.store.set(np.ascontiguousarray(src, dtype=np.uint16).reshape(store.store.shape))
            else:
                store.store.set(as_gpuarray(src).reshape(store.store.shape))
        return self

    def num_classes(self):
        return len(self.id_to_color)

    def _color_table(self):
        """RGBA colours as one uint32 per class id, in id_to_color order"""
        ids = np.fromiter(self.id_to_color.keys(), dtype=np.int64)
        rgba = np.stack([self.id_to_color[int(i)] for i in ids]).astype(np.uint8)
        return ids, rgba, np.ascontiguousarray(rgba).view(np.uint32).ravel()

    def convert_colors_to_ids(self, labels_color):
        """RGBA label picture uint8[H,W,4] -> class ids uint16[H,W]; every pixel must carry a known colour (src/decision_tree.py:88-99)"""
        W, H = self.img_dims
        ids, _, packed = self._color_table()
        px = np.ascontiguousarray(labels_color, dtype=np.uint8).reshape(H, W, 4).view(np.uint32)[..., 0]
        hit = px[None] == packed[:, None, None]                                   # [classes, H, W]
        assert int(hit.sum()) == W * H, 'a pixel carries a colour that is not in id_to_color'
        out = np.zeros((H, W), dtype=np.uint16)
        for k, class_id in enumerate(ids):                                        # later ids win, as in the reference's loop
            out[hit[k]] = class_id
        return out

    def convert_ids_to_colors(self, labels_ids):
        """class ids uint16[N,H,W] -> RGBA uint8[N,H,W,4]; ids without a colour stay transparent black (src/decision_tree.py:101-110)"""
        assert labels_ids.shape[1:] == (self.img_dims[1], self.img_dims[0])
        ids, rgba, _ = self._color_table()
        lut = np.zeros((max(int(ids.max()), int(np.max(labels_ids, initial=0))) + 1, 4), dtype=np.uint8)
        lut[ids] = rgba
        return lut[labels_ids]

    def get_depth_block_cu(self, block_num, arr_out):
        self.depth_blocks.get_block_cu(block_num, arr_out)

    def get_labels_block_cu(self, block_num, arr_out):
        self.labels_blocks.get_block_cu(block_num, arr_out)

    def num_pixels(self):
        return self.num_images * self.img_dims[0] * self.img_dims[1]

    def images_shape(self):
        return (self.num_images, self.img_dims[1], self.img_dims[0])


# ----------------------------------------------------------------------------------------------------------------
# proposals (src/decision_tree.py:350-371): same distributions AND the same np.random draw order, so a seeded
# np.random state yields the reference's proposal stream.
# ----------------------------------------------------------------------------------------------------------------
FEATURE_MAGNITUDE_MAX = 14.
FEATURE_THRESHOLD_MAX = 11.  # _MIN = -_MAX


def make_random_offset():
    f_theta = np.random.uniform(0, np.pi * 2)
    magnitude = np.power(np.e, np.random.uniform(0, FEATURE_MAGNITUDE_MAX))   # linear in log space
    return np.array([np.cos(f_theta), np.sin(f_theta)]) * magnitude


def make_random_feature():
    return make_random_offset(), make_random_offset()


def make_random_threshold():
    return np.random.choice([-1, 1]) * np.power(np.e, np.random.uniform(0, FEATURE_THRESHOLD_MAX))


def make_random_features(n, arr):
    for i in range(n):
        (u, v), t = make_random_feature(), make_random_threshold()
        arr[i] = (u[0], u[1], v[0], v[1], t)


class DecisionTreeTrainer():
    """src/decision_tree.py:373-601: level-synchronous training of one tree.

    Reference form: every proposal is one (feature, threshold) pair, `NUM_PROPOSALS_PER_PROPOSAL_BLOCK` of them are scored
    per pass.  `thresholds_per_feature` > 1 switches to the cfg-4 form: a proposal block is P features with
    that many sorted thresholds each, evaluated once per (pixel, feature).
    Multi-GPU: when torch.distributed is initialised with world_size > 1, every rank holds a shard of the images and the
    split histograms are sum-allreduced (NCCL) before pick-best; all ranks then build the identical tree.
    """

    def __init__(self, NUM_IMAGES_PER_IMAGE_BLOCK, NUM_PROPOSALS_PER_PROPOSAL_BLOCK, thresholds_per_feature=1,
                 proposal_fn=None, process_group=None, hist_budget_bytes=8 << 30, exchange=None):
        """exchange (multi-GPU only).  With the images SHARDED over the ranks (every rank passes its own images):
          'p2p' = the reduction is fused into the histogram kernel (every counter is flushed over NVLink into the peer-mapped
                  buffer of the rank owning its feature, then each rank scores its feature slice and the small per-node winners
                  are all-gathered);
          'allreduce' = NCCL sum-allreduce of the whole histogram, then every rank scores everything.
        Default: 'p2p' when torch symmetric memory can be set up, else 'allreduce' (env RDF_TRAIN_EXCHANGE).
        With the dataset REPLICATED (every rank passes ALL images - 180 GB of HBM holds far more than a training set):
          'features' = every rank buckets all pixels, builds the COMPLETE histograms of its own slice of the proposal block's
                  features, scores them, and only the per-node winners are all-gathered: no histogram ever crosses NVLink
                  (4096 nodes: 8.5 GB of counters stay local); node records, next-active lists and nodes_by_pixel are recomputed
                  identically on every rank."""
        self.exchange = exchange or os.environ.get('RDF_TRAIN_EXCHANGE')
        self._lib = _capi.load()
        self.NUM_IMAGES_PER_IMAGE_BLOCK = NUM_IMAGES_PER_IMAGE_BLOCK
        self.NUM_PROPOSALS_PER_PROPOSAL_BLOCK = NUM_PROPOSALS_PER_PROPOSAL_BLOCK
        self.thresholds_per_feature = int(thresholds_per_feature)
        self.proposal_fn = proposal_fn
        self.process_group = process_group
        self.hist_budget_bytes = int(hist_budget_bytes)

    def allocate(self, dataset, NUM_RANDOM_FEATURES, MAX_TREE_DEPTH):
        self.NUM_RANDOM_FEATURES = NUM_RANDOM_FEATURES
        self.MAX_TREE_DEPTH = MAX_TREE_DEPTH
        if self.NUM_IMAGES_PER_IMAGE_BLOCK is None:
            self.NUM_IMAGES_PER_IMAGE_BLOCK = dataset.num_images
        assert dataset.num_images % self.NUM_IMAGES_PER_IMAGE_BLOCK == 0
        assert self.NUM_RANDOM_FEATURES % self.NUM_PROPOSALS_PER_PROPOSAL_BLOCK == 0
        self.NUM_IMAGE_BLOCKS = dataset.num_images // self.NUM_IMAGES_PER_IMAGE_BLOCK
        self.NUM_PROPOSAL_BLOCKS = self.NUM_RANDOM_FEATURES // self.NUM_PROPOSALS_PER_PROPOSAL_BLOCK
        C = dataset.num_classes()
        _, self.MAX_LEAF_NODES, _ = DecisionTree.get_config(MAX_TREE_DEPTH, C)
        P, NT = self.NUM_PROPOSALS_PER_PROPOSAL_BLOCK, self.thresholds_per_feature

        self.node_counts_cu = cu_array.zeros((self.MAX_LEAF_NODES, C), dtype=np.uint64)
        self.next_node_counts_cu = cu_array.zeros((self.MAX_LEAF_NODES, C), dtype=np.uint64)
        self.active_nodes_cu = cu_array.zeros((self.MAX_LEAF_NODES,), dtype=np.int32)
        self.next_active_nodes_cu = cu_array.zeros((self.MAX_LEAF_NODES,), dtype=np.int32)
        self.next_num_active_nodes_cu = cu_array.zeros((1,), dtype=np.int32)
        self.get_next_num_active_nodes = lambda: int(self.next_num_active_nodes_cu.get()[0])
        self.best_gain_seen_per_node = GPUArray((self.MAX_LEAF_NODES,), dtype=np.float32)
        self.node_slot_cu = GPUArray((self.MAX_LEAF_NODES,), dtype=np.int32)
        self.current_offsets = GPUArray((P, 4), dtype=np.float32)
        self.current_thresholds = GPUArray((P, NT), dtype=np.float32)
        self.current_proposals_block_cpu = None

        # whole dataset resident: nodes_by_pixel int32[N,H,W] (the reference keeps it nvcomp-compressed per block)
        self.nodes_by_pixel = GPUArray(dataset.images_shape(), dtype=np.int32)
        p_hist = P                                                    # features whose histograms this rank holds per slot
        dist0 = self._dist()
        if dist0 is not None and self.exchange == 'features':
            p_hist = (P + dist0.get_world_size(self.process_group) - 1) // dist0.get_world_size(self.process_group)
        per_slot = p_hist * (NT + 1) * C * 4
        self.MAX_SLOTS_PER_BLOCK = int(max(1, min(self.MAX_LEAF_NODES // 2 or 1, self.hist_budget_bytes // per_slot)))
        self.hist_cu = GPUArray((self.MAX_SLOTS_PER_BLOCK, p_hist, NT + 1, C), dtype=np.uint32)
        need = ctypes.c_size_t()
        _capi.check(self._lib.rdf_train_bucket_workspace_bytes(int(np.prod(dataset.images_shape())), self.MAX_SLOTS_PER_BLOCK,
                                                               ctypes.byref(need)))
        self.bucket_ws = GPUArray(((need.value + 3) // 4,), dtype=np.int32)
        self._p2p = None
        self._feat = None
        dist = self._dist()
        if dist is not None and self.exchange == 'features':
            group = self.process_group if self.process_group is not None else dist.group.WORLD
            world, rank = dist.get_world_size(group), dist.get_rank(group)
            L = self.MAX_LEAF_NODES
            dev = self.hist_cu.tensor.device
            self._feat = {'group': group, 'world': world, 'rank': rank, 'Fo': (P + world - 1) // world,
                          'cand_gain': torch.zeros((L,), dtype=torch.float32, device=dev),
                          'cand_idx': torch.zeros((L,), dtype=torch.int32, device=dev),
                          'cand_cnt': torch.zeros((L, 2, C), dtype=torch.int64, device=dev)}
            return
        ntp = 1
        while ntp < NT:
            ntp *= 2
        bucketed_fits = 8 * (16 + 4 * ntp + 4 * (NT + 1) * C + 1) + 64 <= 220 * 1024      # rdf_train_hist_bucketed's shared-memory need (TB_U features, TB_SMEM_KB)
        if dist is not None and self.exchange != 'allreduce' and bucketed_fits:
            try:
                self._setup_p2p(dist, P, NT, C)
            except Exception as e:                               # no symmetric memory on this system: NCCL allreduce instead
                if self.exchange == 'p2p':
                    raise
                self._p2p = None
                self._p2p_error = repr(e)

    def _setup_p2p(self, dist, P, NT, C):
        """Peer-mapped histogram buffers through torch symmetric memory (PyTorch is only the plumbing: allocation, handle
        exchange and the device-side barrier; the reduction itself is rdf_train_hist_bucketed_p2p)."""
        import torch.distributed._symmetric_memory as symm_mem
        group = self.process_group if self.process_group is not None else dist.group.WORLD
        world, rank = dist.get_world_size(group), dist.get_rank(group)
        Fo = (P + world - 1) // world
        n = self.MAX_SLOTS_PER_BLOCK * Fo * (NT + 1) * C
        buf = symm_mem.empty(n, dtype=torch.int32, device=torch.device('cuda', torch.cuda.current_device()))
        hdl = symm_mem.rendezvous(buf, group)
        ptrs = torch.tensor([int(x) for x in hdl.buffer_ptrs], dtype=torch.int64, device=buf.device)
        L = self.MAX_LEAF_NODES
        self._p2p = {
            'group': group, 'world': world, 'rank': rank, 'Fo': Fo, 'buf': buf, 'hdl': hdl, 'ptrs': ptrs,
            'cand_gain': torch.zeros((L,), dtype=torch.float32, device=buf.device),
            'cand_idx': torch.zeros((L,), dtype=torch.int32, device=buf.device),
            'cand_cnt': torch.zeros((L, 2, C), dtype=torch.int64, device=buf.device),
        }

    # -- proposal stream -------------------------------------------------------------------------------------
    def _next_proposals(self, level, block):
        """Next proposal block as (offsets float32[P,4], thresholds float32[P,NT]).  Multi-GPU: rank 0's block is broadcast, so
        every rank scores the same candidates whatever its own np.random state (or proposal_fn) would have produced.
        Thresholds must be finite (ValueError otherwise); with NT > 1 each feature's thresholds are sorted ascending here - the
        bucketed histogram searches them by bisection, and bin k then means "thresholds 0..k-1 are <= the feature"."""
        P, NT = self.NUM_PROPOSALS_PER_PROPOSAL_BLOCK, self.thresholds_per_feature
        if self.proposal_fn is not None:
            offsets, thresholds = self.proposal_fn(level, block)
            offsets = np.ascontiguousarray(offsets, dtype=np.float32).reshape(P, 4)
            thresholds = np.ascontiguousarray(thresholds, dtype=np.float32).reshape(P, NT)
        elif NT == 1:
            if getattr(self, 'current_proposals_block_cpu', None) is None:
                self.current_proposals_block_cpu = np.zeros((P, 5), dtype=np.float32)
            make_random_features(P, self.current_proposals_block_cpu)          # src/decision_tree.py:487
            offsets = np.ascontiguousarray(self.current_proposals_block_cpu[:, 0:4])
            thresholds = np.ascontiguousarray(self.current_proposals_block_cpu[:, 4:5])
        else:
            offsets = np.zeros((P, 4), dtype=np.float32)
            thresholds = np.zeros((P, NT), dtype=np.float32)
            for i in range(P):
                u, v = make_random_feature()
                offsets[i] = (u[0], u[1], v[0], v[1])
                thresholds[i] = np.array([make_random_threshold() for _ in range(NT)], dtype=np.float32)
        dist = self._dist()
        if dist is not None:
            from . import dist as rdist
            offsets, thresholds = rdist.broadcast_proposals(offsets, thresholds, self.process_group)
        if not np.isfinite(thresholds).all():
            raise ValueError('split thresholds must be finite (a NaN threshold routes every pixel right in evaluation but cannot be '
                             'placed in a sorted threshold table)')
        if NT > 1:
            thresholds = np.sort(thresholds, axis=1)
        return offsets, thresholds

    def _dist(self):
        if self.process_group is False:                     # explicitly local, even inside an initialised process group
            return None
        import torch.distributed as dist
        if dist.is_available() and dist.is_initialized() and dist.get_world_size(self.process_group) > 1:
            return dist
        return None

    def _level_block_p2p(self, lib, st, depth, labels, N, W, H, S, P, NT, C, num_active, level, D, tree):
        """One (proposal block, node block) with the feature-sharded exchange (include/rdf_b200.h)."""
        import torch.distributed as dist
        q = self._p2p
        world, rank, Fo = q['world'], q['rank'], q['Fo']
        local = q['buf'][:S * Fo * (NT + 1) * C]
        local.zero_()
        q['hdl'].barrier(channel=0)                                   # every owner buffer is zero before anyone pushes
        _capi.check(lib.rdf_train_hist_bucketed_p2p(_capi.dptr(depth), _capi.dptr(labels), N, W, H, _capi.dptr(self.bucket_ws), S,
                                                    _capi.dptr(self.current_offsets), _capi.dptr(self.current_thresholds), P, NT, C,
                                                    _capi.dptr(q['ptrs']), world, st()))
        q['hdl'].barrier(channel=0)                                   # all ranks' reductions have landed in my slice
        nloc = max(0, min(P, (rank + 1) * Fo) - rank * Fo)
        cg, ci, cc = q['cand_gain'][:num_active], q['cand_idx'][:num_active], q['cand_cnt'][:num_active]
        _capi.check(lib.rdf_train_pick_candidates(num_active, _capi.dptr(self.active_nodes_cu), _capi.dptr(self.node_slot_cu),
                                                  _capi.dptr(self.node_counts_cu), _capi.dptr(local), S, nloc, Fo, rank * Fo, NT, C,
                                                  _capi.dptr(cg), _capi.dptr(ci), _capi.dptr(cc), st()))
        ag = torch.empty((world, num_active), dtype=torch.float32, device=cg.device)
        ai = torch.empty((world, num_active), dtype=torch.int32, device=cg.device)
        ac = torch.empty((world, num_active, 2, C), dtype=torch.int64, device=cg.device)
        dist.all_gather_into_tensor(ag, cg, group=q['group'])
        dist.all_gather_into_tensor(ai, ci, group=q['group'])
        dist.all_gather_into_tensor(ac, cc, group=q['group'])
        _capi.check(lib.rdf_train_pick_finalize(num_active, _capi.dptr(self.active_nodes_cu), _capi.dptr(self.node_slot_cu),
                                                _capi.dptr(self.node_counts_cu), world, _capi.dptr(ag), _capi.dptr(ai), _capi.dptr(ac),
                                                _capi.dptr(self.current_offsets), _capi.dptr(self.current_thresholds), NT, C, level, D,
                                                _capi.dptr(tree.tree_out_cu), _capi.dptr(self.next_node_counts_cu),
                                                _capi.dptr(self.best_gain_seen_per_node), st()))

    def _level_block_features(self, lib, st, depth, labels, N, W, H, S, P, NT, C, num_active, level, D, tree):
        """One (proposal block, node block) with the feature-sharded search over a replicated dataset: complete local histograms
        of this rank's feature slice, local scoring, all-gather of the per-node winners, identical finalisation everywhere."""
        import torch.distributed as dist
        q = self._feat
        world, rank, Fo = q['world'], q['rank'], q['Fo']
        f0 = min(P, rank * Fo)
        nloc = max(0, min(P, (rank + 1) * Fo) - f0)
        cg, ci, cc = q['cand_gain'][:num_active], q['cand_idx'][:num_active], q['cand_cnt'][:num_active]
        if nloc > 0:
            hist = self.hist_cu.tensor.view(-1)[:S * nloc * (NT + 1) * C]
            hist.view(torch.int32).zero_()
            _capi.check(lib.rdf_train_hist_bucketed(_capi.dptr(depth), _capi.dptr(labels), N, W, H, _capi.dptr(self.bucket_ws), S,
                                                    _capi.dptr(self.current_offsets[f0:f0 + nloc]),
                                                    _capi.dptr(self.current_thresholds[f0:f0 + nloc]), nloc, NT, C, _capi.dptr(hist), st()))
            _capi.check(lib.rdf_train_pick_candidates(num_active, _capi.dptr(self.active_nodes_cu), _capi.dptr(self.node_slot_cu),
                                                      _capi.dptr(self.node_counts_cu), _capi.dptr(hist), S, nloc, nloc, f0, NT, C,
                                                      _capi.dptr(cg), _capi.dptr(ci), _capi.dptr(cc), st()))
        else:                                                        # more ranks than features: this rank has no candidate
            cg.fill_(-1.0); ci.fill_(0x7fffffff); cc.zero_()
        ag = torch.empty((world, num_active), dtype=torch.float32, device=cg.device)
        ai = torch.empty((world, num_active), dtype=torch.int32, device=cg.device)
        ac = torch.empty((world, num_active, 2, C), dtype=torch.int64, device=cg.device)
        dist.all_gather_into_tensor(ag, cg, group=q['group'])
        dist.all_gather_into_tensor(ai, ci, group=q['group'])
        dist.all_gather_into_tensor(ac, cc, group=q['group'])
        _capi.check(lib.rdf_train_pick_finalize(num_active, _capi.dptr(self.active_nodes_cu), _capi.dptr(self.node_slot_cu),
                                                _capi.dptr(self.node_counts_cu), world, _capi.dptr(ag), _capi.dptr(ai), _capi.dptr(ac),
                                                _capi.dptr(self.current_offsets), _capi.dptr(self.current_thresholds), NT, C, level, D,
                                                _capi.dptr(tree.tree_out_cu), _capi.dptr(self.next_node_counts_cu),
                                                _capi.dptr(self.best_gain_seen_per_node), st()))

    def train(self, dataset, tree):
        lib, st = self._lib, _capi.stream_ptr
        C = dataset.num_classes()
        D = self.MAX_TREE_DEPTH
        P, NT = self.NUM_PROPOSALS_PER_PROPOSAL_BLOCK, self.thresholds_per_feature
        W, H = dataset.img_dims
        N = dataset.num_images
        dist = self._dist()
        depth = dataset.depth_blocks.store.reshape((N, H, W))
        labels = dataset.labels_blocks.store.reshape((N, H, W))

        tree.tree_out_cu.fill(np.float32(0.))
        # root statistics + nodes_by_pixel = 0 where labelled else -1 (src/decision_tree.py:450-467)
        self.node_counts_cu.fill(0)
        _capi.check(lib.rdf_train_init(_capi.dptr(labels), N * H * W, C, _capi.dptr(self.nodes_by_pixel),
                                       _capi.dptr(self.node_counts_cu), st()))
        if dist is not None and self._feat is None:                  # sharded images: the root statistics are a sum over the ranks
            root = self.node_counts_cu.tensor[0].view(torch.int64)
            dist.all_reduce(root, group=self.process_group)
        self.active_nodes_cu.fill(np.int32(0))
        self.next_num_active_nodes_cu.fill(np.int32(1))

        for current_level in range(D):
            num_active_nodes = self.get_next_num_active_nodes()                 # one D2H sync per level (:479)
            if num_active_nodes == 0:
                break
            self.best_gain_seen_per_node.fill(np.float32(-1.))
            active_host = self.active_nodes_cu.tensor[:num_active_nodes].cpu().numpy()
            slot_blocks = [active_host[i:i + self.MAX_SLOTS_PER_BLOCK] for i in range(0, num_active_nodes, self.MAX_SLOTS_PER_BLOCK)]
            num_nodes_level = 1 << current_level

            single_block = len(slot_blocks) == 1
            for proposal_block_idx in range(self.NUM_PROPOSAL_BLOCKS):
                offsets, thresholds = self._next_proposals(current_level, proposal_block_idx)
                self.current_offsets.set(offsets)
                self.current_thresholds.set(thresholds)
                for nodes_in_block in slot_blocks:
                    S = len(nodes_in_block)
                    if not (single_block and proposal_block_idx > 0):
                        # group the active pixels by histogram slot: once per (level, node block), reused by every proposal
                        # block when the level fits one node block (the common case with 180 GB of HBM)
                        slot_host = np.full((num_nodes_level,), -1, dtype=np.int32)
                        slot_host[nodes_in_block] = np.arange(S, dtype=np.int32)
                        self.node_slot_cu[:num_nodes_level].set(slot_host)
                        _capi.check(lib.rdf_train_bucket(_capi.dptr(self.nodes_by_pixel), N * H * W, _capi.dptr(self.node_slot_cu), S,
                                                         _capi.dptr(self.bucket_ws), self.bucket_ws.nbytes, st()))
                    if self._feat is not None:
                        self._level_block_features(lib, st, depth, labels, N, W, H, S, P, NT, C, num_active_nodes, current_level, D, tree)
                        continue
                    if self._p2p is not None:
                        self._level_block_p2p(lib, st, depth, labels, N, W, H, S, P, NT, C, num_active_nodes, current_level, D, tree)
                        continue
                    hist = self.hist_cu[:S]
                    hist.fill(0)
                    rc = lib.rdf_train_hist_bucketed(_capi.dptr(depth), _capi.dptr(labels), N, W, H, _capi.dptr(self.bucket_ws), S,
                                                     _capi.dptr(self.current_offsets), _capi.dptr(self.current_thresholds),
                                                     P, NT, C, _capi.dptr(hist), st())
                    if rc == -3:        # RDF_ERR_UNSUPPORTED: a feature's [NT+1][C] histogram exceeds shared memory -> raster kernel
                        rc = lib.rdf_train_hist(_capi.dptr(depth), _capi.dptr(labels), _capi.dptr(self.nodes_by_pixel), N, W, H,
                                                _capi.dptr(self.node_slot_cu), S, _capi.dptr(self.current_offsets),
                                                _capi.dptr(self.current_thresholds), P, NT, C, _capi.dptr(hist), st())
                    _capi.check(rc)
                    if dist is not None:                                       # the path's only exchange step
                        dist.all_reduce(hist.tensor.view(torch.int32), group=self.process_group)
                    _capi.check(lib.rdf_train_pick_best(num_active_nodes, _capi.dptr(self.active_nodes_cu),
                                                        _capi.dptr(self.node_slot_cu), _capi.dptr(self.node_counts_cu),
                                                        _capi.dptr(hist), S, _capi.dptr(self.current_offsets),
                                                        _capi.dptr(self.current_thresholds), P, NT, C, current_level, D,
                                                        _capi.dptr(tree.tree_out_cu), _capi.dptr(self.next_node_counts_cu),
                                                        _capi.dptr(self.best_gain_seen_per_node), st()))

            _capi.check(lib.rdf_train_next_active(_capi.dptr(tree.tree_out_cu), current_level, D, C,
                                                  _capi.dptr(self.active_nodes_cu), num_active_nodes,
                                                  _capi.dptr(self.next_active_nodes_cu),
                                                  _capi.dptr(self.next_num_active_nodes_cu), st()))
            if current_level == D - 1:
                break
            self.node_counts_cu.set(self.next_node_counts_cu)                   # :574
            _capi.check(lib.rdf_train_advance_pixels(_capi.dptr(depth), _capi.dptr(self.nodes_by_pixel), N, W, H,
                                                     _capi.dptr(tree.tree_out_cu), current_level, D, C, st()))
            self.active_nodes_cu.set(self.next_active_nodes_cu)                 # :598
