"""Random tiny cases for the fuzz tests (CPU: NumPy oracle vs C oracle; GPU: CUDA path vs C oracle): forests with special
floats, ragged leaves and zero nodes; depth frames with 0 / 65535 / extreme values; labels_reduce, scale and filter variants."""
import numpy as np

SPECIAL_OFFSETS = np.array([0.0, -0.0, 1.0, -1.0, 0.5, 1e-30, -1e-30, 3e5, -3e5, 2097152.0, 4.2e6, -4.2e6, 1e12, -1e12, 3e38,
                            np.inf, -np.inf, np.nan, 65535.0, -65535.0], dtype=np.float32)
SPECIAL_THRESH = np.array([0.0, -0.0, 0.5, -0.5, 1.0, -1.0, 65535.0, -65535.0, 65536.0, 1e9, -1e9, 3e38, -3e38, np.inf, -np.inf,
                           np.nan, 1e-20], dtype=np.float32)
SCALES = [1.0, 0.5, 0.37, 2.0, 1e-3, 1e-12, 1e12, -1.0]


def make_case(seed):
    rng = np.random.default_rng(seed)
    T = int(rng.integers(1, 10))
    D = int(rng.integers(1, 7))
    C = int(rng.integers(1, 13))
    N = int(rng.integers(1, 4))
    H = int(rng.integers(1, 24))
    W = int(rng.integers(1, 40))
    r = int(rng.integers(1, 4))
    scale = float(SCALES[rng.integers(0, len(SCALES))]) if rng.random() < 0.6 else 1.0
    NN = (1 << D) - 1
    forest = np.zeros((T, NN, 7 + 2 * C), np.float32)
    mag = np.exp(rng.uniform(0, 8, size=(T, NN, 4)))
    forest[:, :, 0:4] = (mag * rng.choice([-1.0, 1.0], size=mag.shape)).astype(np.float32)
    forest[:, :, 4] = (np.exp(rng.uniform(0, 11, size=(T, NN))) * rng.choice([-1.0, 1.0], size=(T, NN))).astype(np.float32)
    forest[:, :, 5:7] = rng.choice([-1.0, 0.0, -1.5, 3.0, -0.5], p=[0.6, 0.25, 0.05, 0.05, 0.05], size=(T, NN, 2))   # floor(-0.5) == -1 too
    forest[:, :, 7:] = rng.integers(0, 1024, size=(T, NN, 2 * C)) / 1024.0
    k = rng.random((T, NN, 4)) < 0.08                                      # special offsets / thresholds / zeroed nodes
    forest[:, :, 0:4][k] = SPECIAL_OFFSETS[rng.integers(0, len(SPECIAL_OFFSETS), size=int(k.sum()))]
    k = rng.random((T, NN)) < 0.1
    forest[:, :, 4][k] = SPECIAL_THRESH[rng.integers(0, len(SPECIAL_THRESH), size=int(k.sum()))]
    forest[rng.random((T, NN)) < 0.05] = 0.0
    kind = rng.integers(0, 3)
    if kind == 0:
        depth = rng.integers(0, 65536, size=(N, H, W))
    elif kind == 1:
        depth = 1000 + rng.integers(0, 64, size=(N, H, W))
    else:
        depth = rng.choice(np.array([0, 1, 2, 65534, 65535, 3000]), size=(N, H, W))
    depth = depth.astype(np.uint16)
    h, w = H // r, W // r
    filt = fclass = None
    if rng.random() < 0.4 and h > 0 and w > 0:
        filt = rng.integers(0, 3, size=(N, h, w)).astype(np.uint16)
        fclass = int(rng.integers(0, 3))
    return dict(forest=forest, depth=depth, r=r, scale=scale, filt=filt, fclass=fclass, shape=(N, h, w), C=C)
