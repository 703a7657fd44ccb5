"""NumPy CPU restatement of the 3d-beats RDF hot path.  TEST INFRASTRUCTURE ONLY.

This file is the *oracle*: only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
legs may import it.  The product (3d-beats_b200/) never does; it fails loudly when the CUDA library is missing.

Parity pinning: the reference ships no golden vectors, known-answer tests or saved models for this path
(SURVEY.md section 4 / 8c).  The oracle is therefore pinned against the reference's OWN kernels compiled unchanged
for sm_100a (oracle/ref_kernels -> oracle/_ref/libref_kernels.so) on a B200: (i) live, in tests/test_gpu_reference.py,
and (ii) through tests/golden/*.npz, which were produced by those reference kernels on a B200 with
tests/golden/make_golden.py and are re-checked against this oracle on CPU in tests/test_oracle_golden.py.

Each function cites the reference file:line (relative to the reference repository root) it restates.
All arithmetic is fp32 unless stated; integer work is exact.
"""
import numpy as np

MAX_UINT16 = 65535          # src/cuda/cu_utils.hpp:8, src/util.py:43
_F32 = np.float32
_I32_MIN = -(1 << 31)
_I32_MAX = (1 << 31) - 1


def float2int_rd(x):
    """CUDA __float2int_rd (cvt.rmi.s32.f32): floor, saturating, NaN -> 0.  Returns int64 holding int32 values."""
    x = np.asarray(x, dtype=np.float32)
    with np.errstate(invalid='ignore'):
        f = np.floor(x).astype(np.float64)
    out = np.zeros(x.shape, dtype=np.int64)
    nan = np.isnan(f)
    hi = f >= 2147483648.0
    lo = f <= -2147483648.0
    ok = ~(nan | hi | lo)
    out[ok] = f[ok].astype(np.int64)
    out[hi] = _I32_MAX
    out[lo] = _I32_MIN
    return out


def _wrap_i32(v):
    """int32 two's-complement wrap of an int64 array (coord + offset is an int add on the device)."""
    return ((v + (1 << 31)) & 0xFFFFFFFF) - (1 << 31)


def probe(depth, n, yy, xx):
    """Array3d<uint16>::get with default 65535 for out-of-bounds (src/cuda/cu_utils.hpp:58-62,79-86).
    Bounds are per image: a probe never reads a neighbouring frame."""
    N, H, W = depth.shape
    inb = (yy >= 0) & (yy < H) & (xx >= 0) & (xx < W)
    out = np.full(n.shape, MAX_UINT16, dtype=np.uint16)
    out[inb] = depth[n[inb], yy[inb], xx[inb]]
    return out


def compute_feature(depth, n, Y, X, ux, uy, vx, vy, scale=1.0):
    """Shotton depth-difference feature d(x+u/d(x)) - d(x+v/d(x)) (src/cuda/decision_tree_common.hpp:8-28).

    fp32 multiply (scale*u), IEEE fp32 divide by float(d), floor to int32, two probes, fp32 subtract.
    Returns float32[P].  d == 0 -> 0.0 (line 12)."""
    d = depth[n, Y, X]
    df = d.astype(np.float32)
    s = _F32(scale)
    with np.errstate(divide='ignore', invalid='ignore', over='ignore'):
        ox_u = float2int_rd((s * ux.astype(np.float32)) / df)
        oy_u = float2int_rd((s * uy.astype(np.float32)) / df)
        ox_v = float2int_rd((s * vx.astype(np.float32)) / df)
        oy_v = float2int_rd((s * vy.astype(np.float32)) / df)
    X64 = X.astype(np.int64)
    Y64 = Y.astype(np.int64)
    pu = probe(depth, n, _wrap_i32(Y64 + oy_u), _wrap_i32(X64 + ox_u)).astype(np.float32)
    pv = probe(depth, n, _wrap_i32(Y64 + oy_v), _wrap_i32(X64 + ox_v)).astype(np.float32)
    f = pu - pv
    f[d == 0] = _F32(0.0)
    return f


def _traverse(tree, depth, n, Y, X, scale, max_depth, num_classes):
    """One tree over P pixels.  Returns (leaf_row int64[P] or -1, leaf_side int64[P]).
    Loop structure of src/cuda/tree_eval.cu:95-128 / :174-210; node addressing src/cuda/cu_utils.hpp:32-39."""
    P = n.shape[0]
    g = np.zeros(P, dtype=np.int64)
    leaf_row = np.full(P, -1, dtype=np.int64)
    leaf_side = np.zeros(P, dtype=np.int64)
    active = np.arange(P)
    for j in range(max_depth):
        if active.size == 0:
            break
        rows = (1 << j) - 1 + g[active]
        nd = tree[rows]                                          # [A, 7+2C]
        f = compute_feature(depth, n[active], Y[active], X[active], nd[:, 0], nd[:, 1], nd[:, 2], nd[:, 3], scale)
        with np.errstate(invalid='ignore'):
            side = np.where(f < nd[:, 4], 0, 1).astype(np.int64)   # NaN threshold -> right (tree_eval.cu:106)
        flag = float2int_rd(nd[np.arange(rows.shape[0]), 5 + side])
        cont = flag == -1
        done = ~cont
        leaf_row[active[done]] = rows[done]
        leaf_side[active[done]] = side[done]
        g[active[cont]] = 2 * g[active[cont]] + side[cont]
        active = active[cont]
    return leaf_row, leaf_side


def best_pdf_chance(acc):
    """get_best_pdf_chance (src/cuda/tree_eval.cu:7-21): first class with strictly greatest value > 0, else 0."""
    P, C = acc.shape
    best = np.zeros(P, dtype=np.float32)
    lab = np.zeros(P, dtype=np.int64)
    for c in range(C):
        with np.errstate(invalid='ignore'):
            better = acc[:, c] > best
        best = np.where(better, acc[:, c], best)
        lab = np.where(better, c, lab)
    return lab


def eval_forest(forest, depth, labels_out, labels_reduce=1, filter_images=None, filter_class=None,
                scale=1.0, probs_out=None):
    """evaluate_image_using_forest (src/cuda/tree_eval.cu:24-137) with the host geometry/filter convention of
    DecisionTreeEvaluator.get_labels_forest (src/decision_tree.py:298-330).

    forest float32[T,2^D-1,7+2C]; depth uint16[N,H,W]; labels_out uint16[N,H//r,W//r] is modified IN PLACE and
    skipped pixels (filter mismatch, centre depth 0 or 65535) keep their previous value.  Tree pdfs are summed
    in tree order 0..T-1 (SURVEY note N1).  probs_out: optional float32[N,h,w,C] receiving sum/T for evaluated pixels.
    """
    T, NN, E = forest.shape
    C = (E - 7) // 2
    D = int(np.log2(NN + 1))
    N, H, W = depth.shape
    r = labels_reduce
    h, w = H // r, W // r
    assert labels_out.shape == (N, h, w)
    nn, yy, xx = np.meshgrid(np.arange(N), np.arange(h), np.arange(w), indexing='ij')
    nn = nn.ravel(); yy = yy.ravel(); xx = xx.ravel()
    keep = np.ones(nn.shape, dtype=bool)
    if filter_images is not None:
        assert filter_class is not None and filter_images.shape == labels_out.shape
        if int(filter_class) != -1:                             # tree_eval.cu:81-85
            keep &= filter_images[nn, yy, xx].astype(np.int64) == int(filter_class)
    Y = yy * r
    X = xx * r
    d = depth[nn, Y, X]
    keep &= (d != 0) & (d != MAX_UINT16)                         # tree_eval.cu:88-89
    nn = nn[keep]; yy = yy[keep]; xx = xx[keep]; Y = Y[keep]; X = X[keep]
    P = nn.shape[0]
    acc = np.zeros((P, C), dtype=np.float32)
    for t in range(T):
        leaf_row, leaf_side = _traverse(forest[t], depth, nn, Y, X, scale, D, C)
        got = leaf_row >= 0
        rows = leaf_row[got]
        sides = leaf_side[got]
        cols = 7 + sides[:, None] * C + np.arange(C)[None, :]
        acc[got] = acc[got] + forest[t][rows[:, None], cols]     # fp32 add, tree order
    lab = best_pdf_chance(acc)
    labels_out[nn, yy, xx] = lab.astype(np.uint16)
    if probs_out is not None:
        probs_out[nn, yy, xx, :] = acc / _F32(T)
    return labels_out


def eval_tree(tree, depth, labels_out):
    """evaluate_image_using_tree (src/cuda/tree_eval.cu:140-212): single tree, scale 1, labels_reduce 1.
    A pixel whose traversal falls off the last level with a -1 flag is NOT written."""
    NN, E = tree.shape
    C = (E - 7) // 2
    D = int(np.log2(NN + 1))
    N, H, W = depth.shape
    nn, yy, xx = np.meshgrid(np.arange(N), np.arange(H), np.arange(W), indexing='ij')
    nn = nn.ravel(); yy = yy.ravel(); xx = xx.ravel()
    d = depth[nn, yy, xx]
    keep = (d != 0) & (d != MAX_UINT16)
    nn = nn[keep]; yy = yy[keep]; xx = xx[keep]
    leaf_row, leaf_side = _traverse(tree, depth, nn, yy, xx, 1.0, D, C)
    got = leaf_row >= 0
    rows = leaf_row[got]; sides = leaf_side[got]
    cols = 7 + sides[:, None] * C + np.arange(C)[None, :]
    pdf = tree[rows[:, None], cols]
    lab = best_pdf_chance(pdf)
    labels_out[nn[got], yy[got], xx[got]] = lab.astype(np.uint16)
    return labels_out


def composite(label_images, conditions, composite_out):
    """make_composite_labels_image (src/cuda/tree_eval.cu:214-248).  label_images: list of uint16[h,w];
    conditions int32[n,2]; composite_out uint16[h,w] modified in place (untouched where the walk stops early)."""
    conditions = np.asarray(conditions, dtype=np.int32).reshape(-1, 2)
    h, w = composite_out.shape
    off = np.zeros((h, w), dtype=np.int64)
    alive = np.ones((h, w), dtype=bool)
    for img in label_images:
        l = img.astype(np.int64)
        alive &= (l != 0) & (l != MAX_UINT16)
        idx = np.where(alive, off + l - 1, 0)
        k = conditions[idx, 0]
        v = conditions[idx, 1]
        final = alive & (k == 0)
        composite_out[final] = v[final].astype(np.uint16)
        alive &= ~final
        off = np.where(alive, v, off)
    return composite_out


def layered_run(forests, filters, conditions, depth, labels_reduce=1, scale=1.0):
    """LayeredDecisionForest.run (src/decision_tree.py:233-264).  forests: list of canonical arrays;
    filters: list of (filter_model, filter_model_class) or (None, None); depth uint16[H,W] or [1,H,W].
    Returns (composite uint16[h,w], [per-layer uint16[h,w]])."""
    depth = depth.reshape((1,) + depth.shape[-2:])
    _, H, W = depth.shape
    h, w = H // labels_reduce, W // labels_reduce
    comp = np.full((h, w), MAX_UINT16, dtype=np.uint16)                      # decision_tree.py:237
    imgs = [np.full((1, h, w), MAX_UINT16, dtype=np.uint16) for _ in forests]  # :239-240
    for i, forest in enumerate(forests):
        fm, fc = filters[i]
        eval_forest(forest, depth, imgs[i], labels_reduce,
                    filter_images=imgs[fm] if fm is not None else None,
                    filter_class=fc, scale=scale)
    composite([im[0] for im in imgs], conditions, comp)
    return comp, [im[0] for im in imgs]


def mean_shift(labels, num_labels, variances, num_rounds):
    """MeanShift.run (src/cuda/mean_shift.py:19-59) + kernel `run` (src/cuda/mean_shift.cu:3-48).
    labels uint16[h,w] (or [1,h,w]); returns float64[num_labels,2] = (x,y) per class; NaN for empty classes."""
    labels = labels.reshape(labels.shape[-2:])
    variances = np.asarray(variances, dtype=np.float32)
    ys, xs = np.nonzero((labels != 0) & (labels != MAX_UINT16))
    k = labels[ys, xs].astype(np.int64) - 1
    # the device indexes temp_sum[l-1] unchecked; labels above num_labels are a caller error there
    sel = k < num_labels
    ys, xs, k = ys[sel], xs[sel], k[sel]
    cx = xs.astype(np.float64)
    cy = ys.astype(np.float64)
    means = np.zeros((num_labels, 2), dtype=np.float64)
    for it in range(num_rounds):
        S = np.zeros((num_labels, 3), dtype=np.float64)
        if it == 0:
            np.add.at(S[:, 0], k, cx)
            np.add.at(S[:, 1], k, cy)
            np.add.at(S[:, 2], k, 1.0)
        else:
            dx = cx - means[k, 0]
            dy = cy - means[k, 1]
            v2 = (variances[k] * variances[k]).astype(np.float64)   # fp32 product, then widened (mean_shift.cu:41)
            p = np.exp(-(dx * dx + dy * dy) / (2.0 * v2))
            np.add.at(S[:, 0], k, dx * p)
            np.add.at(S[:, 1], k, dy * p)
            np.add.at(S[:, 2], k, p)
        with np.errstate(invalid='ignore', divide='ignore'):
            means = means + S[:, 0:2] / S[:, 2:3]                   # mean_shift.py:53-55; 0/0 -> NaN
    return means


# ----------------------------------------------------------------------------------------------------------
# training (src/cuda/tree_train.cu, src/decision_tree.py:444-601)
# ----------------------------------------------------------------------------------------------------------

def train_hist_reference_form(depth, labels, nodes_by_pixel, proposals, num_classes,
                              max_nodes_block, elig_min, elig_max):
    """evaluate_random_features (src/cuda/tree_train.cu:4-64): counts[proposal][child - elig_min][label] += 1.
    proposals float32[P,5]; returns uint64[P, max_nodes_block, C]."""
    P = proposals.shape[0]
    counts = np.zeros((P, max_nodes_block, num_classes), dtype=np.uint64)
    nn, yy, xx = np.nonzero(nodes_by_pixel != -1)
    g = nodes_by_pixel[nn, yy, xx].astype(np.int64)
    ok = (2 * g >= elig_min) & (2 * g + 1 < elig_max)              # :42
    nn, yy, xx, g = nn[ok], yy[ok], xx[ok], g[ok]
    lab = labels[nn, yy, xx].astype(np.int64)
    for j in range(P):
        pj = proposals[j]
        ones = np.ones(nn.shape, dtype=np.float32)
        f = compute_feature(depth, nn, yy, xx, ones * pj[0], ones * pj[1], ones * pj[2], ones * pj[3], 1.0)
        child = 2 * g + np.where(f < pj[4], 0, 1)
        np.add.at(counts[j], (child - elig_min, lab), np.uint64(1))
    return counts


def train_hist(depth, labels, nodes_by_pixel, node_slot, num_slots, offsets, thresholds, num_classes):
    """Generalised split histogram (SURVEY 8d cfg 4): one feature evaluation per (pixel, feature), binned against
    NT sorted thresholds.  bin = #{k : t_k <= f} in 0..NT, so the reference's left set `f < t_k`
    (tree_train.cu:59-60) is bins 0..k.  node_slot: int32[num_nodes_at_level] -> slot or -1.
    Returns uint32[num_slots, F, NT+1, C]."""
    F, NT = thresholds.shape
    hist = np.zeros((num_slots, F, NT + 1, num_classes), dtype=np.uint32)
    nn, yy, xx = np.nonzero(nodes_by_pixel >= 0)
    g = nodes_by_pixel[nn, yy, xx].astype(np.int64)
    slot = np.asarray(node_slot)[g].astype(np.int64)
    ok = slot >= 0
    nn, yy, xx, slot = nn[ok], yy[ok], xx[ok], slot[ok]
    lab = labels[nn, yy, xx].astype(np.int64)
    ones = np.ones(nn.shape, dtype=np.float32)
    for j in range(F):
        o = offsets[j]
        f = compute_feature(depth, nn, yy, xx, ones * o[0], ones * o[1], ones * o[2], ones * o[3], 1.0)
        b = np.searchsorted(thresholds[j], f, side='right')          # #{t_k <= f}
        np.add.at(hist[:, j], (slot, b, lab), np.uint32(1))
    return hist


def _fma32(a, b, c):
    """fp32 fused multiply-add, emulated through float64 (a*b is exact in float64; the final double rounding can
    differ from a true fma only on exact float64->float32 ties; the C oracle uses fmaf and is authoritative)."""
    return (np.asarray(a, np.float64) * np.asarray(b, np.float64) + np.asarray(c, np.float64)).astype(np.float32)


def gini_impurity(counts):
    """gini_impurity (src/cuda/tree_train.cu:72-80) as compiled by nvcc 12.9 for sm_100a:
    s = cvt.rn.f32.u64(sum); p = fma(c_i/s, c_i/s, p) in class order; return 1 - p.   counts uint64[..., C]."""
    counts = np.asarray(counts, dtype=np.uint64)
    s = counts.sum(axis=-1).astype(np.float32)
    p = np.zeros(counts.shape[:-1], dtype=np.float32)
    with np.errstate(divide='ignore', invalid='ignore'):
        for i in range(counts.shape[-1]):
            pi = counts[..., i].astype(np.float32) / s
            p = _fma32(pi, pi, p)
    return _F32(1.0) - p


def gini_gain(parent, left, right):
    """gini_gain (src/cuda/tree_train.cu:82-89) in the compiled operation order (verified in PTX, nvcc 12.9):
    left_term = (l/p) * gini(left)  [mul];  remainder = fma(r/p, gini(right), left_term);  gain = gini(parent) - remainder.
    Returns float32; callers apply the `0 if a side is empty` rule (:158-160)."""
    p_sum = np.asarray(parent, np.uint64).sum(axis=-1).astype(np.float32)
    l_sum = np.asarray(left, np.uint64).sum(axis=-1).astype(np.float32)
    r_sum = np.asarray(right, np.uint64).sum(axis=-1).astype(np.float32)
    with np.errstate(divide='ignore', invalid='ignore'):
        left_term = (l_sum / p_sum) * gini_impurity(left)
        rem = _fma32(r_sum / p_sum, gini_impurity(right), left_term)
    return gini_impurity(parent) - rem


def count_above_cutoff(counts, cutoff=np.float32(0.999)):
    """count_above_cutoff (src/cuda/tree_train.cu:92-97): first class with float(c)/float(sum) >= cutoff, else -1."""
    counts = np.asarray(counts, dtype=np.uint64)
    s = _F32(counts.sum())
    with np.errstate(divide='ignore', invalid='ignore'):
        for i in range(counts.shape[0]):
            if _F32(counts[i]) / s >= cutoff:
                return i
    return -1


def pick_best(tree, level, max_depth, active_nodes, parent_counts, child_left, child_right, proposals,
              next_counts, best_gain_seen):
    """pick_best_features (src/cuda/tree_train.cu:99-236) for one proposal block.

    tree float32[2^D-1,7+2C] (in place); active_nodes int[A]; parent_counts uint64[2^D,C] indexed by node id;
    child_left/right uint64[A,P,C] = per active node, per proposal child histograms; proposals float32[P,5];
    next_counts uint64[2^D,C] (in place); best_gain_seen float32[A] (in place)."""
    C = parent_counts.shape[1]
    for i, node in enumerate(np.asarray(active_nodes)):
        node = int(node)
        par = parent_counts[node]
        L = child_left[i]
        R = child_right[i]
        l_sum = L.sum(axis=-1)
        r_sum = R.sum(axis=-1)
        g = gini_gain(np.broadcast_to(par, L.shape), L, R)
        g = np.where((l_sum == 0) | (r_sum == 0), _F32(0.0), g).astype(np.float32)   # :158-160
        best_g = _F32(-1.0)
        best_j = 0
        for j in range(g.shape[0]):                                                   # strict >, first max (:162)
            if g[j] > best_g:
                best_g = g[j]
                best_j = j
        if not (best_g > best_gain_seen[i]):                                          # :172
            continue
        best_gain_seen[i] = best_g
        row = (1 << level) - 1 + node
        tree[row, 0:5] = proposals[best_j]
        p_sum = _F32(par.sum())
        if best_g <= 0.0:                                                             # :190-198
            tree[row, 5] = 0.0
            tree[row, 6] = 0.0
            pdf = par.astype(np.float32) / p_sum
            tree[row, 7:7 + C] = pdf
            tree[row, 7 + C:7 + 2 * C] = pdf
            continue
        for side, cnt, child in ((0, L[best_j], 2 * node), (1, R[best_j], 2 * node + 1)):
            base = 7 + side * C
            cut = count_above_cutoff(cnt)
            if cut > -1:                                                              # :203-206 (stale pdf entries kept)
                tree[row, 5 + side] = 0.0
                tree[row, base + cut] = 1.0
            elif level == max_depth - 1:                                              # :208-212
                tree[row, 5 + side] = 0.0
                tree[row, base:base + C] = cnt.astype(np.float32) / _F32(cnt.sum())
            else:                                                                     # :213-217
                tree[row, 5 + side] = -1.0
                next_counts[child] = cnt


def next_active(tree, level, active_nodes):
    """get_active_nodes_next_level (src/cuda/tree_train.cu:238-273).  The device order is nondeterministic (atomic
    counter); this returns the children in ascending order - compare as sets / sort the device result."""
    out = []
    for node in np.asarray(active_nodes):
        row = (1 << level) - 1 + int(node)
        if tree[row, 5] == -1.0:
            out.append(2 * int(node))
        if tree[row, 6] == -1.0:
            out.append(2 * int(node) + 1)
    return np.array(sorted(out), dtype=np.int32)


def advance_pixels(tree, level, depth, nodes_by_pixel):
    """copy_pixel_groups (src/cuda/tree_train.cu:275-324): apply the chosen split to every active pixel, in place."""
    nn, yy, xx = np.nonzero(nodes_by_pixel != -1)
    g = nodes_by_pixel[nn, yy, xx].astype(np.int64)
    nd = tree[(1 << level) - 1 + g]
    f = compute_feature(depth, nn, yy, xx, nd[:, 0], nd[:, 1], nd[:, 2], nd[:, 3], 1.0)
    with np.errstate(invalid='ignore'):
        left = f < nd[:, 4]
    status = float2int_rd(nd[np.arange(g.shape[0]), np.where(left, 5, 6)])
    nxt = np.where(status != -1, -1, 2 * g + np.where(left, 0, 1))
    nodes_by_pixel[nn, yy, xx] = nxt.astype(np.int32)
    return nodes_by_pixel


def train_tree(depth, labels, num_classes, max_depth, proposal_blocks_fn):
    """DecisionTreeTrainer.train (src/decision_tree.py:444-601) in reference form (one threshold per proposal,
    single node block, all images in one block).  proposal_blocks_fn(level) -> iterable of float32[P,5] blocks.
    Returns the canonical tree float32[2^D-1, 7+2C]."""
    C = num_classes
    D = max_depth
    tree = np.zeros(((1 << D) - 1, 7 + 2 * C), dtype=np.float32)
    node_counts = np.zeros((1 << D, C), dtype=np.uint64)
    un, cnt = np.unique(labels, return_counts=True)                                   # :457-460
    for l, c in zip(un, cnt):
        if l > 0:
            node_counts[0, int(l)] += np.uint64(c)
    nodes_by_pixel = np.where(labels > 0, 0, -1).astype(np.int32)                      # :462-463
    active = np.array([0], dtype=np.int32)
    next_counts = node_counts.copy()
    for level in range(D):
        if active.size == 0:
            break
        best_seen = np.full(active.shape, -1.0, dtype=np.float32)                      # :483
        n_children = 1 << (level + 1)
        for proposals in proposal_blocks_fn(level):
            counts = train_hist_reference_form(depth, labels, nodes_by_pixel, proposals, C, n_children, 0, n_children)
            L = np.stack([counts[:, 2 * int(a), :] for a in active])                   # [A,P,C]
            R = np.stack([counts[:, 2 * int(a) + 1, :] for a in active])
            pick_best(tree, level, D, active, node_counts, L, R, proposals, next_counts, best_seen)
        nxt = next_active(tree, level, active)
        if level == D - 1:
            break
        node_counts = next_counts.copy()                                               # :574
        advance_pixels(tree, level, depth, nodes_by_pixel)
        active = nxt
    return tree
