"""Synthetic inputs for the RDF hot path.

Everything here is pure integer / NumPy host code so that the CPU oracle, the tests and the
device-side generators (csrc/rdf_synth.cu) agree bit for bit.  Nothing in this file touches a GPU.

Frame generators use 32-bit integer hashing only (murmur3 finaliser), forests follow the
proposal distribution of the reference (`src/decision_tree.py:353-367`: direction U(0,2pi),
magnitude e^U(0,14), threshold +-e^U(0,11)).
"""
import json
import os

import numpy as np

MAX_UINT16 = 65535

FEATURE_MAGNITUDE_MAX = 14.0   # src/decision_tree.py:353
FEATURE_THRESHOLD_MAX = 11.0   # src/decision_tree.py:354

_M32 = np.uint32(0xFFFFFFFF)


def mix32(h):
    """murmur3 32-bit finaliser on a uint32 ndarray (wraps mod 2^32)."""
    h = np.asarray(h, dtype=np.uint32).copy()
    h ^= h >> np.uint32(16)
    h *= np.uint32(0x85EBCA6B)
    h ^= h >> np.uint32(13)
    h *= np.uint32(0xC2B2AE35)
    h ^= h >> np.uint32(16)
    return h


def _hash_nyx(n0, num, H, W, seed):
    n = (np.arange(n0, n0 + num, dtype=np.uint32) * np.uint32(0x9E3779B1))[:, None, None]
    y = (np.arange(H, dtype=np.uint32) * np.uint32(0x85EBCA77))[None, :, None]
    x = (np.arange(W, dtype=np.uint32) * np.uint32(0xC2B2AE3D))[None, None, :]
    with np.errstate(over='ignore'):
        return mix32(np.uint32(seed & 0xFFFFFFFF) ^ n ^ y ^ x)


def depth_frames(kind, num, H, W, seed=1234, first_frame=0):
    """uint16[num,H,W] depth frames.  kind in {'dense-smooth','dense-noise','live-mask'}.

    dense-smooth: d = 3000 + ((3x + 2y + 37n) mod 1024) + (h & 31)           (every pixel valid)
    dense-noise : d = 1 + (h mod 65534)                                       (adversarial)
    live-mask   : dense-smooth inside a centred ellipse, 65535 elsewhere      (cfg 2; mirrors
                  convert_0s_to_maxuint in the product, src/3d_bz.py:416-420)
    """
    h = _hash_nyx(first_frame, num, H, W, seed)
    if kind == 'dense-noise':
        return (np.uint32(1) + (h % np.uint32(65534))).astype(np.uint16)
    n = np.arange(first_frame, first_frame + num, dtype=np.uint32)[:, None, None]
    y = np.arange(H, dtype=np.uint32)[None, :, None]
    x = np.arange(W, dtype=np.uint32)[None, None, :]
    d = np.uint32(3000) + ((np.uint32(3) * x + np.uint32(2) * y + np.uint32(37) * n) % np.uint32(1024)) + (h & np.uint32(31))
    d = d.astype(np.uint16)
    if kind == 'dense-smooth':
        return d
    if kind == 'live-mask':
        inside = ellipse_mask(H, W)
        return np.where(inside[None], d, np.uint16(MAX_UINT16)).astype(np.uint16)
    raise ValueError(f'unknown frame kind {kind!r}')


def ellipse_geometry(H, W):
    """centre and radii of the cfg-2 hand blob, scaled from 848x480 (cx=424, cy=240, rx=150, ry=110)."""
    return W // 2, H // 2, max(1, (150 * W) // 848), max(1, (110 * H) // 480)


def ellipse_mask(H, W):
    cx, cy, rx, ry = ellipse_geometry(H, W)
    y = np.arange(H, dtype=np.int64)[:, None]
    x = np.arange(W, dtype=np.int64)[None, :]
    return (x - cx) ** 2 * (ry * ry) + (y - cy) ** 2 * (rx * rx) <= (rx * rx) * (ry * ry)


def train_labels(num, H, W, first_frame=0):
    """cfg 4 label maps: label = 1 + ((x//53 + y//60 + n) mod 3), uint16[num,H,W]; every pixel labelled."""
    n = np.arange(first_frame, first_frame + num, dtype=np.int64)[:, None, None]
    y = np.arange(H, dtype=np.int64)[None, :, None]
    x = np.arange(W, dtype=np.int64)[None, None, :]
    return (1 + ((x // 53 + y // 60 + n) % 3)).astype(np.uint16)


def tree_node_els(num_classes):
    return 7 + 2 * num_classes


def _random_offsets(rng, n):
    theta = rng.uniform(0.0, 2.0 * np.pi, size=n)
    mag = np.exp(rng.uniform(0.0, FEATURE_MAGNITUDE_MAX, size=n))
    return mag * np.cos(theta), mag * np.sin(theta)


def _random_thresholds(rng, n):
    sign = rng.integers(0, 2, size=n) * 2 - 1
    return sign * np.exp(rng.uniform(0.0, FEATURE_THRESHOLD_MAX, size=n))


def random_forest(num_trees, max_depth, num_classes, seed=1234, ragged=False, leaf_bias=None):
    """Random-init forest in the canonical layout float32[T, 2^D-1, 7+2C] (src/decision_tree.py:160-168).

    full-depth (default): child flags -1 at levels 0..D-2 and 0 (leaf) at level D-1, so every evaluated
    pixel performs exactly T*D node-steps.  ragged: from level 2 down each child becomes a leaf with
    p=0.15 (correctness-only variant).  Leaf pdfs are dyadic k/1024 so fp32 sums are exact in any order
    (SURVEY note N1).  leaf_bias: optional float[C] added (as multiples of 1/1024, clipped) to every leaf
    pdf to skew the label distribution (cfg 2 layer 1).
    """
    rng = np.random.default_rng(seed)
    T, D, C = num_trees, max_depth, num_classes
    NN = (1 << D) - 1
    forest = np.zeros((T, NN, tree_node_els(C)), dtype=np.float32)
    n = T * NN
    ux, uy = _random_offsets(rng, n)
    vx, vy = _random_offsets(rng, n)
    th = _random_thresholds(rng, n)
    forest[:, :, 0] = ux.reshape(T, NN)
    forest[:, :, 1] = uy.reshape(T, NN)
    forest[:, :, 2] = vx.reshape(T, NN)
    forest[:, :, 3] = vy.reshape(T, NN)
    forest[:, :, 4] = th.reshape(T, NN)

    level_of = np.floor(np.log2(np.arange(1, NN + 1))).astype(np.int64)   # row r (0-based) is at level floor(log2(r+1))
    flags = np.full((T, NN, 2), -1.0, dtype=np.float32)
    flags[:, level_of == D - 1, :] = 0.0
    if ragged:
        early = rng.random((T, NN, 2)) < 0.15
        early[:, level_of < 2, :] = False
        flags[early] = 0.0
    forest[:, :, 5:7] = flags

    k = rng.integers(0, 1024, size=(T, NN, 2, C)).astype(np.float64)
    if leaf_bias is not None:
        k = np.clip(k + np.asarray(leaf_bias, dtype=np.float64)[None, None, None, :] * 1024.0, 0, 4095)
        k = np.floor(k)
    pdf = (k / 1024.0).astype(np.float32)
    is_leaf = (flags != -1.0)
    pdf = pdf * is_leaf[..., None]
    forest[:, :, 7:7 + C] = pdf[:, :, 0, :]
    forest[:, :, 7 + C:7 + 2 * C] = pdf[:, :, 1, :]
    return forest


def hash_forest(num_trees, max_depth, num_classes, seed=1234, trees=None):
    """Hash-defined full-depth forest (bit-exact twin of csrc/rdf_synth.cu:synth_forest_kernel).

    Used for forests too large to build with a NumPy RNG and ship over PCIe (cfg 5: 7.5 GiB).  Each float is
    assembled from hashed bits: log-uniform magnitude 2^[0,20) (u,v) / 2^[0,16) (thresh), random mantissa and
    sign.  Leaf pdfs are k/1024.  `trees` optionally selects a sub-range of trees (for sampling in tests).
    """
    T, D, C = num_trees, max_depth, num_classes
    NN = (1 << D) - 1
    E = tree_node_els(C)
    tsel = np.arange(T, dtype=np.uint32) if trees is None else np.asarray(trees, dtype=np.uint32)
    out = np.zeros((len(tsel), NN, E), dtype=np.float32)
    rows = np.arange(NN, dtype=np.uint32)
    level_of = np.floor(np.log2(np.arange(1, NN + 1))).astype(np.int64)
    last = level_of == D - 1
    with np.errstate(over='ignore'):
        for ti, t in enumerate(tsel):
            base = np.uint32(seed & 0xFFFFFFFF) ^ (np.uint32(t) * np.uint32(0x9E3779B1)) ^ (rows * np.uint32(0x85EBCA77))
            for e in range(E):
                h = mix32(base ^ np.uint32((e * 0xC2B2AE3D) & 0xFFFFFFFF))
                if e < 5:
                    span = np.uint32(20 if e < 4 else 16)
                    expo = np.uint32(127) + ((h >> np.uint32(24)) % span)
                    bits = ((h & np.uint32(1)) << np.uint32(31)) | (expo << np.uint32(23)) | ((h >> np.uint32(1)) & np.uint32(0x7FFFFF))
                    out[ti, :, e] = bits.view(np.float32)
                elif e < 7:
                    out[ti, :, e] = np.where(last, np.float32(0.0), np.float32(-1.0))
                else:
                    out[ti, :, e] = np.where(last, (h & np.uint32(1023)).astype(np.float32) / np.float32(1024.0), np.float32(0.0))
    return out


def layered_cfg2(seed=1234, max_depth=16, num_trees=3):
    """cfg 2: L1 none/hand/other (C=3) -> L2 10 finger parts (C=11) gated by L1 label 1.

    Returns (forests, cfg_dict, variances).  conditions=[[1,2],[0,11],[0,1],...,[0,10]]:
    L1 label 1 -> consult L2 (table offset 2); L1 label 2 -> id 11; L2 label k -> id k.
    """
    l1 = random_forest(num_trees, max_depth, 3, seed=seed + 0, leaf_bias=[-1.0, 0.2, 0.0])
    l2 = random_forest(num_trees, max_depth, 11, seed=seed + 1)
    conditions = [[1, 2], [0, 11]] + [[0, k] for k in range(1, 11)]
    colors = [[(37 * i) % 256, (91 * i) % 256, (53 * i + 80) % 256, 255] for i in range(1, 12)]
    cfg = {
        'layers': [
            {'model': 'layer1.npy'},
            {'model': 'layer2.npy', 'filter_model': 0, 'filter_model_class': 1},
        ],
        'conditions': conditions,
        'label_colors': colors,
    }
    variances = np.array([8.0] * 10 + [50.0], dtype=np.float32)
    return [l1, l2], cfg, variances


def write_layered_model(directory, forests, cfg, name='layered.json'):
    """Write the layered-model JSON + per-layer .npy exactly as LayeredDecisionForest.load expects
    (src/decision_tree.py:173-199)."""
    os.makedirs(directory, exist_ok=True)
    for layer, forest in zip(cfg['layers'], forests):
        np.save(os.path.join(directory, layer['model']), forest)
    path = os.path.join(directory, name)
    with open(path, 'w') as f:
        json.dump(cfg, f)
    return path


def random_proposals(num_features, num_thresholds, seed=1234):
    """cfg 4 proposals: offsets float32[F,4] = (ux,uy,vx,vy), thresholds float32[F,NT] sorted ascending."""
    rng = np.random.default_rng(seed)
    ux, uy = _random_offsets(rng, num_features)
    vx, vy = _random_offsets(rng, num_features)
    offsets = np.stack([ux, uy, vx, vy], axis=1).astype(np.float32)
    th = _random_thresholds(rng, num_features * num_thresholds).reshape(num_features, num_thresholds)
    thresholds = np.sort(th.astype(np.float32), axis=1)
    return offsets, thresholds


def random_node_assignment(labels, level, seed=1234):
    """cfg 4: node ids as produced by a depth-`level` random tree: int32 like labels, -1 where label==0.
    Uses a hash of 32x32 pixel tiles so that node membership is spatially clustered like a real tree."""
    N, H, W = labels.shape
    if level == 0:
        nodes = np.zeros((N, H, W), dtype=np.int32)
    else:
        h = _hash_nyx(0, N, (H + 31) // 32, (W + 31) // 32, seed ^ 0x5bd1e995)
        tile = (h % np.uint32(1 << level)).astype(np.int32)
        nodes = np.repeat(np.repeat(tile, 32, axis=1), 32, axis=2)[:, :H, :W].copy()
    nodes[labels == 0] = -1
    return nodes


def live_scene(H=480, W=848, seed=1234, num_hands=2):
    """A raw camera frame of the live product (src/3d_bz.py): a tilted table seen from above, hand-sized blobs 0-5 cm above it
    (their rims sink into the plane-clip threshold), a small distractor blob below the grouping size threshold, missing samples
    (zeros) sprinkled everywhere and a band without data.  Units are the camera's 0.1 mm (src/rs_util.py:28).

    Returns a dict: depth_raw uint16[H,W], pp float32[2], focal / fx / fy, plane float32[4,4] (camera -> plane space, row-major as
    CalibratedPlane.get_mat() returns it; plane-space z is 0 on the table and negative above it), plane_z_threshold (product
    default, src/3d_bz.py:54)."""
    focal = 425.5 * W / 848.0
    pp = np.array([W / 2.0 - 3.25, H / 2.0 + 1.75], dtype=np.float32)
    n = np.array([0.08, -0.35, 0.93])
    n /= np.linalg.norm(n)
    p0 = np.array([0.0, 0.0, 4200.0])
    xa = np.cross([0.0, 1.0, 0.0], n)
    xa /= np.linalg.norm(xa)
    ya = np.cross(n, xa)
    R = np.stack([xa, ya, n])
    plane = np.eye(4)
    plane[:3, :3] = R
    plane[:3, 3] = -R @ p0
    plane = plane.astype(np.float32)

    yy, xx = np.mgrid[0:H, 0:W].astype(np.float64)
    dirx, diry = (xx - float(pp[0])) / focal, (yy - float(pp[1])) / focal
    t_table = float(n @ p0) / (n[0] * dirx + n[1] * diry + n[2])
    h = _hash_nyx(seed & 0xFFFF, 1, H, W, seed)[0]
    depth = t_table + ((h & np.uint32(15)).astype(np.float64) - 8.0)
    blobs = [(0.29 * W, 0.56 * H, 0.135 * W, 0.19 * H, 520.0), (0.72 * W, 0.48 * H, 0.125 * W, 0.20 * H, 460.0),
             (0.50 * W, 0.12 * H, 0.035 * W, 0.06 * H, 300.0)]
    if num_hands < 2:
        blobs = blobs[:num_hands] + blobs[2:]
    for cx, cy, rx, ry, top in blobs:
        r2 = ((xx - cx) / rx) ** 2 + ((yy - cy) / ry) ** 2
        height = top * np.sqrt(np.clip(1.0 - r2, 0.0, None)) + 3.0 * np.sin(xx / 7.0) * np.cos(yy / 5.0)
        inside = r2 < 1.0
        depth = np.where(inside, t_table - height + ((h >> np.uint32(4)) & np.uint32(7)).astype(np.float64) - 3.0, depth)
    depth = np.clip(np.rint(depth), 1, 65534).astype(np.uint16)
    depth[(h % np.uint32(53)) == 0] = 0                       # missing samples
    depth[H - max(1, H // 24):, :] = 0                        # band without data
    depth[:, : max(1, W // 100)] = 0
    return dict(depth_raw=depth, pp=pp, focal=np.float32(focal), fx=np.float32(focal), fy=np.float32(focal * 1.0015),
                plane=plane, plane_z_threshold=np.float32(40.0))
