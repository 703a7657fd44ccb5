"""The reference's live scripts' per-frame call sequences executed VERBATIM through the drop-in shim (compat/rdf_dropin.py):
run_live_layered.py:80-135 and run_live.py:79-124 - pycuda-style kernel calls with grid= / block=, GpuBuffer objects, the
layered forest loaded from its JSON config - and, kernel by kernel, src/3d_bz.py:159-259,390-456.  Every buffer the sequence
touches is compared bit for bit with the reference's own kernels (points_ops.cu, calibrated_plane.cu, tree_eval.cu compiled
unchanged, oracle/_ref) run in the same order.  Only camera / window / OpenGL lines are left out."""
import os
import sys
import types

import numpy as np
import pytest

from conftest import to_dev, to_np

pytestmark = pytest.mark.gpu

COMPAT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), '3d-beats_b200', 'compat')


def _shim():
    if COMPAT not in sys.path:
        sys.path.insert(0, COMPAT)
    import rdf_dropin  # noqa: F401


def _bits(a):
    return np.ascontiguousarray(a).view(np.uint32)


def test_run_live_layered_tick_sequence_through_the_shim(tmp_path):
    import torch
    _shim()
    ns = {}
    exec('from decision_tree import *\nfrom cuda.points_ops import *\nfrom calibrated_plane import *\n'
         'import cuda.py_nvcc_utils as py_nvcc_utils\nfrom engine.buffer import GpuBuffer', ns)     # run_live_layered.py:6-13
    from rdf_b200 import synth
    from oracle import ref_points as rp, ref_kernels as rk
    LayeredDecisionForest, PointsOps, CalibratedPlane, GpuBuffer = (ns[k] for k in ('LayeredDecisionForest', 'PointsOps', 'CalibratedPlane', 'GpuBuffer'))

    scene = synth.live_scene(480, 848, seed=21)
    forests, cfg, _ = synth.layered_cfg2(seed=5, max_depth=12)
    cfg_path = synth.write_layered_model(str(tmp_path), forests, cfg)

    # ---- __init__ (run_live_layered.py:31-52), camera replaced by the synthetic scene, plane supplied instead of RANSAC ----
    self = types.SimpleNamespace()
    self.PLANE_Z_OUTLIER_THRESHOLD = 40.
    self.calibrated_plane = CalibratedPlane(25000, self.PLANE_Z_OUTLIER_THRESHOLD)
    self.calibrated_plane.set_mat(scene['plane'])
    self.TRAIN_DIM_X = 848
    self.DIM_X, self.DIM_Y, self.FOCAL, self.PP = 848, 480, scene['focal'], scene['pp']
    self.LABELS_REDUCE = 2
    self.layered_rdf = LayeredDecisionForest.load(cfg_path, (self.DIM_Y, self.DIM_X), self.LABELS_REDUCE)
    self.points_ops = PointsOps()
    self.pts = GpuBuffer((self.DIM_Y, self.DIM_X, 4), dtype=np.float32)
    self.pts.cu().fill(np.float32(0.))                     # the reference's GpuBuffer starts zeroed (GL buffer storage)
    self.depth_image = GpuBuffer((1, self.DIM_Y, self.DIM_X), np.uint16)
    self.labels_image = GpuBuffer((1, self.DIM_Y // self.LABELS_REDUCE, self.DIM_X // self.LABELS_REDUCE), dtype=np.uint16)
    self.labels_image_rgba = GpuBuffer((self.DIM_Y // self.LABELS_REDUCE, self.DIM_X // self.LABELS_REDUCE, 4), dtype=np.uint8)

    # ---- tick (run_live_layered.py:80-135), verbatim from here ----
    depth_image = np.asanyarray(scene['depth_raw']).reshape((1, self.DIM_Y, self.DIM_X))
    self.depth_image.cu().set(depth_image)

    grid_dim = (1, (self.DIM_X // 32) + 1, (self.DIM_Y // 32) + 1)
    block_dim = (1,32,32)

    # convert depth image to points
    self.points_ops.deproject_points(
        np.array([1, self.DIM_X, self.DIM_Y, -1], dtype=np.int32),
        self.PP,
        np.float32(self.FOCAL),
        self.depth_image.cu(),
        self.pts.cu(),
        grid=grid_dim,
        block=block_dim)

    if not self.calibrated_plane.is_set():
        self.calibrated_plane.make(self.pts, (self.DIM_X, self.DIM_Y))

    # every point..
    grid_dim2 = (((self.DIM_X * self.DIM_Y) // 1024) + 1, 1, 1)
    block_dim2 = (1024, 1, 1)

    self.points_ops.transform_points(
        np.int32(self.DIM_X * self.DIM_Y),
        self.pts.cu(),
        self.calibrated_plane.get_mat(),
        grid=grid_dim2,
        block=block_dim2)

    self.calibrated_plane.filter_points_by_plane(
        np.int32(self.DIM_X * self.DIM_Y),
        np.float32(self.PLANE_Z_OUTLIER_THRESHOLD),
        self.pts.cu(),
        grid=grid_dim2,
        block=block_dim2)

    self.points_ops.setup_depth_image_for_forest(
        np.int32(self.DIM_X * self.DIM_Y),
        self.pts.cu(),
        self.depth_image.cu(),
        grid=grid_dim2,
        block=block_dim2)

    # run RDF!
    self.layered_rdf.run(self.depth_image, self.labels_image, self.DIM_X / self.TRAIN_DIM_X)

    # make RGBA image
    self.labels_image_rgba.cu().fill(0)
    self.points_ops.make_rgba_from_labels(
        np.uint32(self.DIM_X // self.LABELS_REDUCE),
        np.uint32(self.DIM_Y // self.LABELS_REDUCE),
        np.uint32(self.layered_rdf.num_layered_classes),
        self.labels_image.cu(),
        self.layered_rdf.label_colors.cu(),
        self.labels_image_rgba.cu(),
        grid = ((self.DIM_X // 32) + 1, (self.DIM_Y // 32) + 1, 1),
        block = (32,32,1))
    # ---- end of the verbatim block ----
    torch.cuda.synchronize()

    ref_pts, ref_depth = rp.live_frame_for_forest(scene['depth_raw'], scene['pp'], scene['focal'], scene['plane'], self.PLANE_Z_OUTLIER_THRESHOLD)
    assert np.array_equal(_bits(self.pts.cu().get()), _bits(ref_pts)), 'point image differs from the reference kernels'
    assert np.array_equal(self.depth_image.cu().get()[0], ref_depth), 'forest input depth differs from the reference kernels'
    assert (ref_depth == 65535).any() and (ref_depth != 65535).sum() > 20000          # table clipped away, hands kept
    ref_comp, _ = rk.layered_run([to_dev(f) for f in forests], [(None, None), (0, 1)], cfg['conditions'], to_dev(ref_depth),
                                 self.LABELS_REDUCE, self.DIM_X / self.TRAIN_DIM_X)
    torch.cuda.synchronize()
    ref_comp = to_np(ref_comp)
    assert np.array_equal(self.labels_image.cu().get()[0], ref_comp), 'composite label map differs from the reference kernels'
    assert len(np.unique(ref_comp)) >= 4
    h, w = ref_comp.shape
    ref_rgba = rp.make_rgba_from_labels(ref_comp, np.array(cfg['label_colors'], np.uint8), np.zeros((h, w, 4), np.uint8))
    assert np.array_equal(self.labels_image_rgba.cu().get(), ref_rgba)

    # CalibratedPlane.make is the one call of the script outside the replaced path: it must say so, not guess a plane
    with pytest.raises(NotImplementedError):
        CalibratedPlane(10, 40.).make(self.pts, (self.DIM_X, self.DIM_Y))


def test_run_live_tick_sequence_through_the_shim():
    """run_live.py:79-124: same conditioning, then one forest through DecisionTreeEvaluator.get_labels_forest."""
    import torch
    _shim()
    ns = {}
    exec('from decision_tree import *\nfrom cuda.points_ops import *\nfrom calibrated_plane import *\nfrom engine.buffer import GpuBuffer', ns)
    from rdf_b200 import synth
    from oracle import ref_points as rp, ref_kernels as rk
    DecisionForest, DecisionTreeEvaluator, PointsOps, CalibratedPlane, GpuBuffer = (
        ns[k] for k in ('DecisionForest', 'DecisionTreeEvaluator', 'PointsOps', 'CalibratedPlane', 'GpuBuffer'))
    scene = synth.live_scene(480, 848, seed=4)
    f = synth.random_forest(3, 12, 5, seed=9, ragged=True)

    self = types.SimpleNamespace()
    self.PLANE_Z_OUTLIER_THRESHOLD = 40.
    self.calibrated_plane = CalibratedPlane(25000, self.PLANE_Z_OUTLIER_THRESHOLD)
    self.calibrated_plane.set_mat(scene['plane'])
    self.DIM_X, self.DIM_Y, self.FOCAL, self.PP = 848, 480, scene['focal'], scene['pp']
    self.forest = DecisionForest(3, 12, 5)
    self.forest.forest_cu.set(f)
    self.decision_tree_evaluator = DecisionTreeEvaluator()
    self.points_ops = PointsOps()
    self.pts = GpuBuffer((self.DIM_Y, self.DIM_X, 4), dtype=np.float32)
    self.pts.cu().fill(np.float32(0.))
    self.depth_image = GpuBuffer((1, self.DIM_Y, self.DIM_X), np.uint16)
    self.labels_image = GpuBuffer((1, self.DIM_Y, self.DIM_X), dtype=np.uint16)

    depth_image = np.asanyarray(scene['depth_raw']).reshape((1, self.DIM_Y, self.DIM_X))
    self.depth_image.cu().set(depth_image)
    grid_dim = (1, (self.DIM_X // 32) + 1, (self.DIM_Y // 32) + 1)
    block_dim = (1,32,32)
    self.points_ops.deproject_points(
        np.array([1, self.DIM_X, self.DIM_Y, -1], dtype=np.int32),
        self.PP,
        np.float32(self.FOCAL),
        self.depth_image.cu(),
        self.pts.cu(),
        grid=grid_dim,
        block=block_dim)
    grid_dim2 = (((self.DIM_X * self.DIM_Y) // 1024) + 1, 1, 1)
    block_dim2 = (1024, 1, 1)
    self.points_ops.transform_points(
        np.int32(self.DIM_X * self.DIM_Y),
        self.pts.cu(),
        self.calibrated_plane.get_mat(),
        grid=grid_dim2,
        block=block_dim2)
    self.calibrated_plane.filter_points_by_plane(
        np.int32(self.DIM_X * self.DIM_Y),
        np.float32(self.PLANE_Z_OUTLIER_THRESHOLD),
        self.pts.cu(),
        grid=grid_dim2,
        block=block_dim2)
    self.points_ops.setup_depth_image_for_forest(
        np.int32(self.DIM_X * self.DIM_Y),
        self.pts.cu(),
        self.depth_image.cu(),
        grid=grid_dim2,
        block=block_dim2)
    self.labels_image.cu().fill(np.uint16(65535))
    self.decision_tree_evaluator.get_labels_forest(self.forest, self.depth_image.cu(), self.labels_image.cu())
    labels_image_cpu = self.labels_image.cu().get()
    torch.cuda.synchronize()

    _, ref_depth = rp.live_frame_for_forest(scene['depth_raw'], scene['pp'], scene['focal'], scene['plane'], self.PLANE_Z_OUTLIER_THRESHOLD)
    ref_labels = torch.full((1, 480, 848), -1, dtype=torch.int16, device='cuda').view(torch.uint16)
    rk.eval_forest(to_dev(f), to_dev(ref_depth).reshape(1, 480, 848), ref_labels)
    torch.cuda.synchronize()
    assert np.array_equal(self.depth_image.cu().get()[0], ref_depth)
    assert np.array_equal(labels_image_cpu, to_np(ref_labels))


def test_product_loop_kernels_in_pycuda_form_match_the_reference():
    """src/3d_bz.py:159-259 and run_per_hand_pipeline (:390-456) kernel by kernel in the reference's call form: remove_missing,
    shrink_image, write_pixel_groups_to_stencil_image, grow_groups, stencil_depth_image_by_group, flip_x, convert_0s_to_maxuint,
    make_depth_rgba - each against the reference's kernel."""
    import torch
    _shim()
    ns = {}
    exec('from cuda.points_ops import *\nfrom calibrated_plane import *\nfrom engine.buffer import GpuBuffer\nfrom cpp_grouping import CppGrouping', ns)
    from rdf_b200 import synth
    from oracle import ref_points as rp, grouping_oracle as go
    PointsOps, CalibratedPlane, GpuBuffer, CppGrouping = (ns[k] for k in ('PointsOps', 'CalibratedPlane', 'GpuBuffer', 'CppGrouping'))
    scene = synth.live_scene(480, 848, seed=33)
    DIM_X, DIM_Y, level = 848, 480, 3
    mm_dims = (DIM_Y >> level, DIM_X >> level)
    ops, plane = PointsOps(), CalibratedPlane(1, 40.)
    plane.set_mat(scene['plane'])
    pts = GpuBuffer((DIM_Y, DIM_X, 4), dtype=np.float32)
    pts.cu().fill(np.float32(0.))
    depth = GpuBuffer((DIM_Y, DIM_X), np.uint16)
    depth.cu().set(scene['depth_raw'])
    block_dim2, grid_dim2 = (1024, 1, 1), ((DIM_X * DIM_Y) // 1024 + 1, 1, 1)
    ops.deproject_points(np.array([1, DIM_X, DIM_Y, -1], dtype=np.int32), scene['pp'], np.float32(scene['focal']), depth.cu(), pts.cu(),
                         grid=(1, 27, 15), block=(1, 32, 32))
    ops.transform_points(np.int32(DIM_X * DIM_Y), pts.cu(), plane.get_mat(), grid=grid_dim2, block=block_dim2)
    plane.filter_points_by_plane(np.int32(DIM_X * DIM_Y), np.float32(40.), pts.cu(), grid=grid_dim2, block=block_dim2)
    ops.remove_missing_3d_points_from_depth_image(np.int32(DIM_X * DIM_Y), pts.cu(), depth.cu(), grid=grid_dim2, block=block_dim2)
    depth_mm = GpuBuffer(mm_dims, np.uint16)
    ops.shrink_image(np.array((DIM_X, DIM_Y), dtype=np.int32), np.int32(level), depth.cu(), depth_mm.cu(), grid=(4, 2, 1), block=(32, 32, 1))
    torch.cuda.synchronize()
    ref_depth, ref_mm = rp.condition_frame(scene['depth_raw'], scene['pp'], scene['focal'], scene['plane'], 40., None, level)
    assert np.array_equal(depth.cu().get(), ref_depth) and np.array_equal(depth_mm.cu().get(), ref_mm)

    # grouping on the host-array signature, coordinate scatter, grow (src/3d_bz.py:222-259)
    coords = np.zeros((mm_dims[0] * mm_dims[1], 3), dtype=np.int32)
    g_info = np.zeros((2, 3), dtype=np.float32)
    CppGrouping().make_groups(depth_mm.cu().get(), coords, g_info, 0.06)
    n = int(g_info[0, 0] + g_info[1, 0])
    assert n > 0
    ref_coords, ref_info = go.ref_make_groups(ref_mm, 0.06)
    ref_stencil = go.stencil_from_coords(ref_coords, mm_dims[0], mm_dims[1])
    assert np.array_equal(g_info[:, 0], ref_info[:, 0])
    coords_gpu = GpuBuffer(coords.shape, np.int32)
    coords_gpu.cu()[0:n, :].set(coords[0:n])
    groups_2, groups = GpuBuffer(mm_dims, np.uint16), GpuBuffer(mm_dims, np.uint16)
    groups_2.cu().fill(0)
    ops.write_pixel_groups_to_stencil_image(coords_gpu.cu(), np.int32(n), groups_2.cu(), np.array(mm_dims, dtype=np.int32),
                                            grid=(n // 32 + 1, 1, 1), block=(32, 1, 1))
    ops.grow_groups(np.array([mm_dims[1], mm_dims[0]], dtype=np.int32), groups_2.cu(), groups.cu(), grid=(4, 2, 1), block=(32, 32, 1))
    torch.cuda.synchronize()
    assert np.array_equal(groups_2.cu().get(), ref_stencil)
    ref_grown = rp.grow_groups(ref_stencil)
    assert np.array_equal(groups.cu().get(), ref_grown)

    # per hand (src/3d_bz.py:390-420)
    group_img, image_2 = GpuBuffer((DIM_Y, DIM_X), np.uint16), GpuBuffer((DIM_Y, DIM_X), np.uint16)
    for g_id, flip in ((1, False), (2, True)):
        group_img.cu().fill(0)
        ops.stencil_depth_image_by_group(np.array([DIM_X, DIM_Y], dtype=np.int32), np.int32(level), np.int32(g_id), groups.cu(), depth.cu(),
                                         group_img.cu(), grid=(27, 15, 1), block=(32, 32, 1))
        if flip:
            ops.flip_x(np.array([DIM_X, DIM_Y], dtype=np.int32), group_img.cu(), image_2.cu(), grid=(27, 15, 1), block=(32, 32, 1))
        else:
            image_2.cu().set(group_img.cu())
        ops.convert_0s_to_maxuint(np.int32(DIM_X * DIM_Y), image_2.cu(), grid=grid_dim2, block=block_dim2)
        torch.cuda.synchronize()
        assert np.array_equal(image_2.cu().get(), rp.hand_depth_image(ref_depth, ref_grown, level, g_id, flip))
    rgba = GpuBuffer(mm_dims + (4,), np.uint8)
    ops.make_depth_rgba(np.array([mm_dims[1], mm_dims[0]], dtype=np.int32), np.uint16(0), np.uint16(2), groups.cu(), rgba.cu(),
                        grid=(4, 2, 1), block=(32, 32, 1))
    torch.cuda.synchronize()
    assert np.array_equal(rgba.cu().get(), rp.make_depth_rgba(ref_grown, 0, 2))
