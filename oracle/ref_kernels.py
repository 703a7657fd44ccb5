"""Launch the REFERENCE's own CUDA kernels (compiled unchanged for sm_100a into oracle/_ref/libref_kernels.so) with the
reference's launch geometry and host choreography.  TEST INFRASTRUCTURE ONLY (tests/, smoke(), bench.py reference legs).

This is parity oracle #2 and the same-box performance bar.  The host sequences below restate
the reference's Python drivers: get_labels / get_labels_forest / make_composite_labels_image
(src/decision_tree.py:277-347), LayeredDecisionForest.run (:233-264), MeanShift.run (src/cuda/mean_shift.py:19-59,
including its per-round D2H/H2D round trips) and DecisionTreeTrainer.train (:444-601).
"""
import ctypes
import os

import numpy as np
import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
SO_PATH = os.path.join(_HERE, '_ref', 'libref_kernels.so')
_lib = None


def available():
    return os.path.exists(SO_PATH)


def lib():
    global _lib
    if _lib is None:
        if not available():
            raise FileNotFoundError(f'{SO_PATH} missing: run `make -C oracle ref` where /root/reference exists')
        _lib = ctypes.CDLL(SO_PATH)
    return _lib


def _p(t):
    return ctypes.c_void_p(t.data_ptr()) if t is not None else None


def _st():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def _u16(t):
    assert t.dtype in (torch.uint16, torch.int16) and t.is_cuda and t.is_contiguous()
    return t


def _ok(rc):
    if rc != 0:
        raise RuntimeError(f'reference kernel launch failed ({rc})')


def eval_forest(forest, depth, labels, labels_reduce=1, filter_images=None, filter_class=None, scale=1.0):
    T, NN, E = forest.shape
    C = (E - 7) // 2
    D = int(np.log2(NN + 1))
    N, H, W = depth.shape
    assert tuple(labels.shape) == (N, H // labels_reduce, W // labels_reduce)
    _ok(lib().ref_eval_forest(T, N, W, H, C, D, _p(_u16(depth)), -1 if filter_images is None else int(filter_class),
                              _p(filter_images), _p(forest), _p(_u16(labels)), int(labels_reduce), ctypes.c_float(scale), _st()))


def eval_tree(tree, depth, labels):
    NN, E = tree.shape
    C = (E - 7) // 2
    D = int(np.log2(NN + 1))
    N, H, W = depth.shape
    _ok(lib().ref_eval_tree(N, W, H, C, D, _p(_u16(depth)), _p(tree), _p(_u16(labels)), _st()))


def composite(label_ptrs, num_images, dim_x, dim_y, conditions, composite_out):
    _ok(lib().ref_composite(_p(label_ptrs), int(num_images), int(dim_x), int(dim_y), _p(conditions), _p(composite_out), _st()))


def layered_run(forests, filters, conditions, depth, labels_reduce=1, scale=1.0):
    """LayeredDecisionForest.run: fills, one forest launch per layer, composite.  forests: list of CUDA float32 tensors;
    depth uint16[H,W] CUDA; returns (composite uint16[h,w], [layer uint16[h,w]])."""
    H, W = depth.shape[-2:]
    h, w = H // labels_reduce, W // labels_reduce
    dev = depth.device
    comp = torch.full((h, w), -1, dtype=torch.int16, device=dev).view(torch.uint16)
    imgs = [torch.full((1, h, w), -1, dtype=torch.int16, device=dev).view(torch.uint16) for _ in forests]
    d3 = depth.reshape(1, H, W)
    for i, f in enumerate(forests):
        fm, fc = filters[i]
        eval_forest(f, d3, imgs[i], labels_reduce, imgs[fm] if fm is not None else None, fc, scale)
    ptrs = torch.tensor([im.data_ptr() for im in imgs], dtype=torch.int64, device=dev)
    cond = torch.as_tensor(np.asarray(conditions, dtype=np.int32).reshape(-1, 2), device=dev)
    composite(ptrs, len(imgs), w, h, cond, comp)
    return comp, [im[0] for im in imgs]


def mean_shift(labels, num_labels, variances, num_rounds):
    """MeanShift.run with the reference's per-round host round trips (2 x D2H, host divide, H2D)."""
    h, w = labels.shape[-2:]
    dev = labels.device
    means = torch.zeros((num_labels, 2), dtype=torch.float64, device=dev)
    temp = torch.zeros((num_labels, 3), dtype=torch.float64, device=dev)
    var = torch.as_tensor(np.asarray(variances, dtype=np.float32), device=dev) if not isinstance(variances, torch.Tensor) else variances
    for i in range(num_rounds):
        temp.zero_()
        _ok(lib().ref_mean_shift_round(_p(labels), int(num_labels), w, h, _p(var), _p(means), i, _p(temp), _st()))
        temp_cpu = temp.cpu().numpy()
        means_cpu = means.cpu().numpy()
        with np.errstate(invalid='ignore', divide='ignore'):
            means_cpu += temp_cpu[:, 0:2] / temp_cpu[:, 2].reshape((num_labels, 1))
        means.copy_(torch.from_numpy(means_cpu))
    return means.cpu().numpy()


def train_tree(depth, labels, num_classes, max_depth, proposal_blocks_fn, max_next_nodes=None):
    """DecisionTreeTrainer.train (src/decision_tree.py:444-601) with one image block and one node block per level.
    depth/labels: CUDA uint16[N,H,W]; proposal_blocks_fn(level) -> iterable of np.float32[P,5].
    Returns the canonical tree as np.float32[2^D-1, 7+2C]."""
    L = lib()
    dev = depth.device
    N, H, W = depth.shape
    C, D = num_classes, max_depth
    max_leaf = 1 << D
    if max_next_nodes is None:
        max_next_nodes = max_leaf
    tree = torch.zeros(((1 << D) - 1, 7 + 2 * C), dtype=torch.float32, device=dev)
    labels_cpu = labels.cpu().view(torch.int16).numpy().view(np.uint16)
    node_counts = np.zeros((max_leaf, C), dtype=np.uint64)
    un, cnt = np.unique(labels_cpu, return_counts=True)
    for l, c in zip(un, cnt):
        if l > 0:
            node_counts[0, int(l)] += np.uint64(c)
    nodes_by_pixel = torch.from_numpy(np.where(labels_cpu > 0, 0, -1).astype(np.int32)).to(dev)
    node_counts_cu = torch.from_numpy(node_counts.view(np.int64)).to(dev)
    next_node_counts_cu = node_counts_cu.clone()
    active = torch.zeros((max_leaf,), dtype=torch.int32, device=dev)
    next_active = torch.zeros((max_leaf,), dtype=torch.int32, device=dev)
    num_next = torch.ones((1,), dtype=torch.int32, device=dev)
    best_gain = torch.zeros((max_leaf,), dtype=torch.float32, device=dev)
    for level in range(D):
        num_active = int(num_next.cpu()[0])
        if num_active == 0:
            break
        best_gain.fill_(-1.0)
        n_children = 1 << (level + 1)
        assert n_children <= max_next_nodes
        for proposals in proposal_blocks_fn(level):
            P = proposals.shape[0]
            prop = torch.from_numpy(np.ascontiguousarray(proposals, dtype=np.float32)).to(dev)
            counts = torch.zeros((P, max_next_nodes, C), dtype=torch.int64, device=dev)
            _ok(L.ref_evaluate_random_features(N, W, H, P, C, D, max_next_nodes, 0, n_children, _p(labels), _p(depth), _p(prop),
                                               _p(nodes_by_pixel), _p(counts), _st()))
            _ok(L.ref_pick_best_features(num_active, P, D, max_next_nodes, 0, n_children, C, level, _p(active),
                                         _p(node_counts_cu), _p(counts), _p(prop), _p(tree), _p(next_node_counts_cu),
                                         _p(best_gain), _st()))
        num_next.zero_()
        next_active.zero_()
        _ok(L.ref_get_active_nodes_next_level(level, D, C, _p(tree), _p(active), num_active, _p(next_active), _p(num_next), _st()))
        if level == D - 1:
            break
        node_counts_cu.copy_(next_node_counts_cu)
        _ok(L.ref_copy_pixel_groups(N, W, H, level, D, C, _p(depth), _p(nodes_by_pixel), _p(tree), _st()))
        active.copy_(next_active)
    torch.cuda.synchronize()
    return tree.cpu().numpy()
