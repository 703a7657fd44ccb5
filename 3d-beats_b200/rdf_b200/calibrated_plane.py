"""Mirror of the part of the reference's `CalibratedPlane` (src/calibrated_plane.py:12-37) that sits on the per-frame path:
the plane matrix holder and `filter_points_by_plane` (src/cuda/calibrated_plane.cu:30-45), called pycuda-style by run_live.py:109-114,
run_live_layered.py:110-115 and src/3d_bz.py:191-196.

The one-off RANSAC calibration (`make`: make_plane_candidates + find_plane_ransac, src/calibrated_plane.py:39-93) is OUTSIDE the
hot path this library replaces: supply the plane with `set_mat` (e.g. a matrix saved from the reference's
own calibration); `make` says so instead of guessing one."""
import numpy as np

from .points_ops import PointsOps


class CalibratedPlane():
    def __init__(self, num_random_guesses=0, plane_z_outlier_threshold=40.):
        self.num_random_guesses = num_random_guesses
        self.plane_z_outlier_threshold = plane_z_outlier_threshold
        self.plane = None
        self._ops = PointsOps()

    def is_set(self):
        return self.plane is not None

    def get_mat(self):
        assert self.is_set()
        return self.plane

    def set_mat(self, mat):
        """camera -> plane space, float32[4,4] row-major (what the reference's make() leaves in `.plane`)"""
        mat = np.ascontiguousarray(mat, dtype=np.float32)
        assert mat.shape == (4, 4)
        self.plane = mat

    def make(self, pts_gpu, img_dims, start_mat=None):
        raise NotImplementedError(
            'plane calibration by RANSAC (src/calibrated_plane.py:39-93) is outside the decision-forest hot path rdf_b200 replaces: '
            'call CalibratedPlane.set_mat(float32[4,4]) with a calibrated camera->plane matrix')

    def filter_points_by_plane(self, num_pts, plane_z_threshold, pts, grid=None, block=None):
        self._ops.filter_points_by_plane(num_pts, plane_z_threshold, pts)
