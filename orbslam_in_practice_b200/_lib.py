"""ctypes binding of liborbx.so (include/orbx.h).  Used by tests/, bench.py and smoke().

This is a test/bench harness binding, not the product host layer: the reference is C++, so the
host-side mirror of its classes lives in orbslam_in_practice_b200/cpp/ (ORBSlam::ORBextractor,
ORBSlam::ORBmatcher over the same C ABI).  There is no CPU fallback: a missing library or a
missing sm_100 device raises.
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
SO_PATH = os.path.join(_HERE, "liborbx.so")

KEYPOINT_DTYPE = np.dtype([("x", "<f4"), ("y", "<f4"), ("size", "<f4"), ("angle", "<f4"),
                           ("response", "<f4"), ("octave", "<i4"), ("class_id", "<i4")])
CAND_DTYPE = np.dtype([("x", "<i2"), ("y", "<i2"), ("score", "<i4")])

# every symbol include/orbx.h declares (tests check the library exports all of them)
ABI_SYMBOLS = [
    "orbx_strerror", "orbx_last_cuda_error", "orbx_version", "orbx_build_id", "orbx_device_count",
    "orbx_create", "orbx_destroy", "orbx_nlevels", "orbx_capacity", "orbx_tables",
    "orbx_extract_host", "orbx_extract_host_begin", "orbx_extract_host_end", "orbx_extract_device", "orbx_set_pyramid_border", "orbx_set_device_split", "orbx_set_low_latency", "orbx_debug_guard_check", "orbx_set_input_format", "orbx_undistort_keypoints_device",
    "orbx_level_dims", "orbx_download_level", "orbx_level_device_ptr",
    "orbx_download_candidates", "orbx_download_kept", "orbx_max_candidates", "orbx_launch_count",
    "orbx_set_profiling", "orbx_stage_times", "orbm_set_profiling", "orbm_knn2_times",
    "orbm_knn2_pairs_workspace_bytes", "orbm_knn2_pairs_device", "orbm_create", "orbm_destroy", "orbm_launch_count", "orbm_hamming_pairs_host",
    "orbm_knn2_device", "orbm_knn2_host", "orbm_ratio_select_device", "orbm_merge_shards_device",
    "orbm_popc_peak", "orbm_search_init_workspace_bytes", "orbm_search_init_device", "orbm_search_init_host",
    "orbm_search_window_device", "orbm_search_window_host", "orbm_search_groups_device", "orbm_search_groups_host",
    "orbm_exchange_create", "orbm_exchange_open", "orbm_knn2_sharded_device", "orbm_exchange_status",
]


class OrbxError(RuntimeError):
    pass


class Params(C.Structure):
    _fields_ = [("nfeatures", C.c_int32), ("scale_factor", C.c_float), ("nlevels", C.c_int32),
                ("ini_th_fast", C.c_int32), ("min_th_fast", C.c_int32)]


_lib = None


class WindowParams(C.Structure):
    """orbm_window_params (include/orbx.h): per-query windowed search (SearchByProjection-style)."""
    _fields_ = [("radius", C.c_float), ("level_scale", C.c_float * 16),
                ("query_level_min", C.c_int32), ("query_level_max", C.c_int32),
                ("level_below", C.c_int32), ("level_above", C.c_int32), ("gate", C.c_int32), ("th_dist", C.c_int32),
                ("nnratio", C.c_float), ("check_orientation", C.c_int32), ("update_centers", C.c_int32),
                ("width", C.c_int32), ("height", C.c_int32), ("literal_gridid_bug", C.c_int32),
                ("use_bounds", C.c_int32), ("min_x", C.c_float), ("max_x", C.c_float), ("min_y", C.c_float), ("max_y", C.c_float)]

    @classmethod
    def projection(cls, th, scale_factors, width, height, th_dist=100, check_orientation=True, level_below=1, level_above=1):
        """Upstream ORB-SLAM2 frame-to-frame SearchByProjection: r = th * scaleFactor[octave], octaves +-1, best <= TH_HIGH."""
        p = cls()
        p.radius = th
        for i in range(16):
            p.level_scale[i] = scale_factors[i] if i < len(scale_factors) else 1.0
        p.query_level_min, p.query_level_max = 0, 15
        p.level_below, p.level_above, p.gate, p.th_dist, p.nnratio = level_below, level_above, 1, th_dist, 0.0
        p.check_orientation, p.update_centers = int(check_orientation), 0
        p.width, p.height, p.literal_gridid_bug = width, height, 0
        return p


def load():
    """Load liborbx.so; raises if it has not been built (no fallback path exists)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(SO_PATH):
        raise OrbxError("liborbx.so is missing: run `python -m orbslam_in_practice_b200.build` "
                        "(or __graft_entry__.build()); there is no CPU fallback")
    L = C.CDLL(SO_PATH)
    vp, i32, f32, sz, ll = C.c_void_p, C.c_int, C.c_float, C.c_size_t, C.c_longlong
    L.orbx_strerror.restype = C.c_char_p
    L.orbx_strerror.argtypes = [i32]
    L.orbx_last_cuda_error.restype = C.c_char_p
    L.orbx_build_id.restype = C.c_char_p
    L.orbx_create.argtypes = [C.POINTER(Params), i32, i32, i32, i32, C.POINTER(vp)]
    L.orbx_destroy.argtypes = [vp]
    L.orbx_nlevels.argtypes = [vp]
    L.orbx_capacity.argtypes = [vp]
    L.orbx_tables.argtypes = [vp] * 7
    L.orbx_extract_host.argtypes = [vp, vp, sz, sz, i32, i32, i32, vp, vp, vp]
    L.orbx_extract_host_begin.argtypes = [vp, vp, sz, sz, i32, i32, i32, vp, vp, vp]
    L.orbx_extract_host_end.argtypes = [vp]
    L.orbx_extract_device.argtypes = [vp, vp, sz, sz, i32, i32, i32, vp, vp, vp, vp]
    L.orbx_set_pyramid_border.argtypes = [vp, i32]
    L.orbx_set_device_split.argtypes = [vp, i32]
    L.orbx_set_low_latency.argtypes = [vp, i32]
    L.orbx_debug_guard_check.argtypes = [vp, C.POINTER(i32)]
    L.orbx_set_input_format.argtypes = [vp, i32, i32]
    L.orbx_undistort_keypoints_device.argtypes = [vp, vp, vp, i32, vp, vp, i32, vp]
    L.orbx_level_dims.argtypes = [vp, i32, C.POINTER(i32), C.POINTER(i32)]
    L.orbx_download_level.argtypes = [vp, i32, i32, i32, i32, vp, sz]
    L.orbx_level_device_ptr.argtypes = [vp, i32, i32, C.POINTER(vp), C.POINTER(sz)]
    L.orbx_download_candidates.argtypes = [vp, i32, i32, vp, i32, C.POINTER(i32)]
    L.orbx_download_kept.argtypes = [vp, i32, i32, vp, i32, C.POINTER(i32)]
    L.orbx_max_candidates.argtypes = [vp, i32]
    L.orbx_launch_count.restype = ll
    L.orbx_launch_count.argtypes = [vp]
    L.orbm_create.argtypes = [i32, i32, i32, C.POINTER(vp)]
    L.orbm_destroy.argtypes = [vp]
    L.orbm_launch_count.restype = ll
    L.orbm_launch_count.argtypes = [vp]
    L.orbm_hamming_pairs_host.argtypes = [vp, vp, vp, i32, vp]
    L.orbm_knn2_device.argtypes = [vp, vp, i32, vp, i32, i32, vp, vp, vp, vp]
    L.orbm_knn2_host.argtypes = [vp, vp, i32, vp, i32, i32, vp, vp, vp]
    L.orbm_ratio_select_device.argtypes = [vp, vp, vp, vp, i32, i32, f32, vp, vp]
    L.orbm_merge_shards_device.argtypes = [vp, vp, vp, vp, i32, i32, sz, vp, vp, vp, vp]
    L.orbx_set_profiling.argtypes = [vp, i32]
    L.orbx_stage_times.argtypes = [vp, vp]
    L.orbm_set_profiling.argtypes = [vp, i32]
    L.orbm_knn2_times.argtypes = [vp, C.POINTER(f32), C.POINTER(f32)]
    L.orbm_popc_peak.argtypes = [i32, C.POINTER(C.c_double), C.POINTER(C.c_double)]
    L.orbm_search_window_device.argtypes = [vp, vp, vp, vp, i32, vp, vp, i32, vp, vp, vp, vp, vp, sz, vp]
    L.orbm_search_window_host.argtypes = [vp, vp, vp, i32, vp, vp, i32, vp, vp, C.POINTER(i32), vp]
    L.orbm_search_groups_device.argtypes = [vp, vp, vp, vp, vp, i32, vp, vp, i32, vp, vp, i32, f32, i32, vp, sz, vp]
    L.orbm_search_groups_host.argtypes = [vp, vp, vp, vp, i32, vp, vp, vp, i32, vp, C.POINTER(i32), i32, f32, i32]
    L.orbm_knn2_pairs_workspace_bytes.restype = sz
    L.orbm_knn2_pairs_workspace_bytes.argtypes = [i32, i32]
    L.orbm_knn2_pairs_device.argtypes = [vp, vp, vp, i32, vp, vp, i32, vp, vp, vp, i32, f32, vp, vp, sz, vp]
    L.orbm_search_init_workspace_bytes.restype = sz
    L.orbm_search_init_workspace_bytes.argtypes = [i32, i32]
    L.orbm_exchange_create.argtypes = [vp, i32, i32, i32, vp]
    L.orbm_exchange_open.argtypes = [vp, vp]
    L.orbm_knn2_sharded_device.argtypes = [vp, vp, i32, vp, i32, i32, vp, vp, vp, i32, f32, vp, vp]
    L.orbm_exchange_status.argtypes = [vp]
    L.orbm_search_init_host.argtypes = [vp, vp, vp, i32, vp, vp, i32, vp, vp, C.POINTER(i32), i32, f32, i32, i32, i32, i32]
    L.orbm_search_init_device.argtypes = [vp, vp, vp, vp, i32, vp, vp, i32, vp, vp, vp, i32, f32, i32, i32, i32, i32, vp, sz, vp]
    _lib = L
    return L


def build_id():
    return load().orbx_build_id().decode()


def check(rc):
    if rc != 0:
        L = load()
        msg = L.orbx_strerror(rc).decode()
        cu = L.orbx_last_cuda_error().decode()
        raise OrbxError("liborbx: %s (%d)%s" % (msg, rc, (" [" + cu + "]") if rc in (-2, -6) and cu else ""))


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


class Extractor:
    """Batched ORBextractor over the C ABI (ORBextractor.h:29-97)."""

    def __init__(self, nfeatures=1000, scale_factor=1.2, nlevels=8, ini_th=20, min_th=7,
                 max_width=640, max_height=480, max_batch=1, device=0):
        L = load()
        self.params = Params(nfeatures, scale_factor, nlevels, ini_th, min_th)
        self.h = C.c_void_p()
        self.nlevels, self.max_batch, self.device = nlevels, max_batch, device
        check(L.orbx_create(C.byref(self.params), max_width, max_height, max_batch, device, C.byref(self.h)))
        self.capacity = L.orbx_capacity(self.h)
        t = [np.zeros(nlevels, np.float32) for _ in range(4)] + [np.zeros(nlevels, np.int32), np.zeros(16, np.int32)]
        check(L.orbx_tables(self.h, *[_p(a) for a in t]))
        (self.scale_factors, self.inv_scale_factors, self.level_sigma2, self.inv_level_sigma2,
         self.features_per_level, self.umax) = t

    def close(self):
        if getattr(self, "h", None) and self.h.value and _lib is not None:
            _lib.orbx_destroy(self.h)
            self.h = C.c_void_p()

    __del__ = close

    def set_input_format(self, channels, rgb_order=False):
        check(load().orbx_set_input_format(self.h, channels, int(rgb_order)))
        self.in_channels = channels

    def undistort_keypoints_device(self, in_ptr, out_ptr, n, cam, dist, literal_bug=False, stream=0):
        cam = np.ascontiguousarray(cam, np.float32); dist = np.ascontiguousarray(dist, np.float32)
        check(load().orbx_undistort_keypoints_device(self.h, in_ptr, out_ptr, n, _p(cam), _p(dist), int(literal_bug), stream))

    def set_device_split(self, nsplit):
        check(load().orbx_set_device_split(self.h, int(nsplit)))

    def guard_check(self):
        """ORBX_GUARD=1 handles: raises if a kernel wrote into a canary zone around one of the handle's device buffers."""
        bad = C.c_int(-1)
        check(load().orbx_debug_guard_check(self.h, C.byref(bad)))

    def set_low_latency(self, on):
        check(load().orbx_set_low_latency(self.h, int(on)))

    def set_pyramid_border(self, on):
        check(load().orbx_set_pyramid_border(self.h, int(on)))

    def extract_host(self, imgs, out=None):
        """imgs: (F,H,W) or (H,W) u8 host array -> (kps [F,cap], desc [F,cap,32], counts [F])."""
        imgs = np.asarray(imgs)
        C_ = getattr(self, "in_channels", 1)
        if imgs.ndim == (2 if C_ == 1 else 3):
            imgs = imgs[None]
        if C_ != 1:
            assert imgs.dtype == np.uint8 and imgs.ndim == 4 and imgs.shape[3] == C_ and imgs.strides[3] == 1 and imgs.strides[2] == C_
            F, H, W = imgs.shape[:3]
        else:
            assert imgs.dtype == np.uint8 and imgs.ndim == 3 and imgs.strides[2] == 1
            F, H, W = imgs.shape
        if out is None:
            out = (np.zeros((F, self.capacity), KEYPOINT_DTYPE), np.zeros((F, self.capacity, 32), np.uint8),
                   np.zeros(F, np.int32))
        kps, desc, counts = out
        check(load().orbx_extract_host(self.h, _p(imgs), imgs.strides[1], imgs.strides[0] if F > 1 else imgs.strides[1] * H,
                                       W, H, F, _p(kps), _p(desc), _p(counts)))
        return kps, desc, counts

    def extract_host_ptr(self, img_ptr, row_pitch, frame_stride, W, H, F, kps_ptr, desc_ptr, counts_ptr):
        check(load().orbx_extract_host(self.h, img_ptr, row_pitch, frame_stride, W, H, F, kps_ptr, desc_ptr, counts_ptr))

    def extract_host_begin(self, img_ptr, row_pitch, frame_stride, W, H, F, kps_ptr, desc_ptr, counts_ptr):
        check(load().orbx_extract_host_begin(self.h, img_ptr, row_pitch, frame_stride, W, H, F, kps_ptr, desc_ptr, counts_ptr))

    def extract_host_end(self):
        check(load().orbx_extract_host_end(self.h))

    def extract_device(self, img_ptr, row_pitch, frame_stride, W, H, F, kps_ptr, desc_ptr, counts_ptr, stream=0):
        check(load().orbx_extract_device(self.h, img_ptr, row_pitch, frame_stride, W, H, F, kps_ptr, desc_ptr,
                                         counts_ptr, stream))

    def level_dims(self, level):
        w, h = C.c_int(), C.c_int()
        check(load().orbx_level_dims(self.h, level, C.byref(w), C.byref(h)))
        return w.value, h.value

    def level(self, frame, level, blurred=False, border=0):
        w, h = self.level_dims(level)
        out = np.zeros((h + 2 * border, w + 2 * border), np.uint8)
        check(load().orbx_download_level(self.h, frame, level, int(blurred), border, _p(out), out.strides[0]))
        return out

    def _cands(self, fn, frame, level, cap):
        out = np.zeros(max(cap, 1), CAND_DTYPE)
        n = C.c_int()
        check(fn(self.h, frame, level, _p(out), cap, C.byref(n)))
        return out[:n.value].copy()

    def candidates(self, frame, level):
        return self._cands(load().orbx_download_candidates, frame, level, load().orbx_max_candidates(self.h, level))

    def kept(self, frame, level):
        return self._cands(load().orbx_download_kept, frame, level, self.capacity)

    def set_profiling(self, on):
        check(load().orbx_set_profiling(self.h, int(on)))

    def stage_times(self):
        """ms per stage of the last call: level0, resize chain, FAST, octree, blur, describe"""
        ms = np.zeros(6, np.float32)
        check(load().orbx_stage_times(self.h, _p(ms)))
        return ms

    @property
    def launches(self):
        return load().orbx_launch_count(self.h)


class Matcher:
    """ORBmatcher distance / best-2 search over the C ABI (ORBmatcher.cpp:37-67,128-144)."""

    def __init__(self, max_queries, max_db, device=0):
        self.h = C.c_void_p()
        self.max_q, self.max_db = max_queries, max_db
        check(load().orbm_create(max_queries, max_db, device, C.byref(self.h)))

    def close(self):
        if getattr(self, "h", None) and self.h.value and _lib is not None:
            _lib.orbm_destroy(self.h)
            self.h = C.c_void_p()

    __del__ = close

    def hamming_pairs(self, a, b):
        a = np.ascontiguousarray(a, np.uint8).reshape(-1, 32); b = np.ascontiguousarray(b, np.uint8).reshape(-1, 32)
        assert len(a) == len(b)
        out = np.zeros(len(a), np.int32)
        check(load().orbm_hamming_pairs_host(self.h, _p(a), _p(b), len(a), _p(out)))
        return out

    def knn2_host(self, q, db, index_base=0):
        q = np.ascontiguousarray(q, np.uint8).reshape(-1, 32); db = np.ascontiguousarray(db, np.uint8).reshape(-1, 32)
        d1, idx1, d2 = (np.zeros(len(q), np.int32) for _ in range(3))
        check(load().orbm_knn2_host(self.h, _p(q), len(q), _p(db), len(db), index_base, _p(d1), _p(idx1), _p(d2)))
        return d1, idx1, d2

    def knn2_device(self, q_ptr, nq, db_ptr, ndb, index_base, d1_ptr, idx1_ptr, d2_ptr, stream=0):
        check(load().orbm_knn2_device(self.h, q_ptr, nq, db_ptr, ndb, index_base, d1_ptr, idx1_ptr, d2_ptr, stream))

    def knn2_pairs_device(self, desc_ptr, counts_ptr, capacity, pair_a_ptr, pair_b_ptr, npairs, d1_ptr, idx1_ptr, d2_ptr,
                          th_low, ratio, match_ptr, ws_ptr, ws_bytes, stream=0):
        check(load().orbm_knn2_pairs_device(self.h, desc_ptr, counts_ptr, capacity, pair_a_ptr, pair_b_ptr, npairs, d1_ptr, idx1_ptr,
                                            d2_ptr, th_low, ratio, match_ptr, ws_ptr, ws_bytes, stream))

    def ratio_select_device(self, d1_ptr, idx1_ptr, d2_ptr, nq, th_low, ratio, match_ptr, stream=0):
        check(load().orbm_ratio_select_device(self.h, d1_ptr, idx1_ptr, d2_ptr, nq, th_low, ratio, match_ptr, stream))

    def merge_shards_device(self, d1_ptr, idx1_ptr, d2_ptr, nshards, nq, od1_ptr, oidx1_ptr, od2_ptr, stream=0,
                            shard_stride=0):
        check(load().orbm_merge_shards_device(self.h, d1_ptr, idx1_ptr, d2_ptr, nshards, nq, shard_stride,
                                              od1_ptr, oidx1_ptr, od2_ptr, stream))

    def search_init_device(self, kps_ptr, desc_ptr, counts_ptr, capacity, pair_a_ptr, pair_b_ptr, npairs, prev_ptr,
                           matches_ptr, nmatches_ptr, window, nnratio, check_ori, width, height, ws_ptr, ws_bytes,
                           stream=0, literal_bug=False):
        check(load().orbm_search_init_device(self.h, kps_ptr, desc_ptr, counts_ptr, capacity, pair_a_ptr, pair_b_ptr, npairs,
                                             prev_ptr, matches_ptr, nmatches_ptr, window, nnratio, int(check_ori), width,
                                             height, int(literal_bug), ws_ptr, ws_bytes, stream))

    # ---- peer-memory exchange for the database-sharded search ----
    def exchange_create(self, max_queries, rank, world):
        h = (C.c_ubyte * 64)()
        check(load().orbm_exchange_create(self.h, max_queries, rank, world, h))
        return bytes(h)

    def exchange_open(self, handles):
        buf = (C.c_ubyte * (64 * len(handles))).from_buffer_copy(b"".join(handles))
        check(load().orbm_exchange_open(self.h, buf))

    def knn2_sharded_device(self, q_ptr, nq, db_ptr, ndb, index_base, d1_ptr, idx1_ptr, d2_ptr, th_low, ratio, match_ptr, stream=0):
        check(load().orbm_knn2_sharded_device(self.h, q_ptr, nq, db_ptr, ndb, index_base, d1_ptr, idx1_ptr, d2_ptr, th_low, ratio,
                                              match_ptr, stream))

    def exchange_status(self):
        check(load().orbm_exchange_status(self.h))

    def search_init_host(self, kp1, desc1, kp2, desc2, prev_matched, window=100, nnratio=0.9, check_ori=True,
                         width=640, height=480, literal_bug=False):
        kp1 = np.ascontiguousarray(kp1, KEYPOINT_DTYPE); kp2 = np.ascontiguousarray(kp2, KEYPOINT_DTYPE)
        desc1 = np.ascontiguousarray(desc1, np.uint8); desc2 = np.ascontiguousarray(desc2, np.uint8)
        prev = np.ascontiguousarray(prev_matched, np.float32).copy()
        m12 = np.zeros(len(kp1), np.int32); n = C.c_int()
        check(load().orbm_search_init_host(self.h, _p(kp1), _p(desc1), len(kp1), _p(kp2), _p(desc2), len(kp2), _p(prev), _p(m12),
                                           C.byref(n), window, nnratio, int(check_ori), width, height, int(literal_bug)))
        return n.value, m12, prev

    def search_window_host(self, kp1, desc1, kp2, desc2, centers, params):
        """params: a WindowParams (orbm_window_params).  Returns (nmatches, matches12, centers)."""
        kp1 = np.ascontiguousarray(kp1, KEYPOINT_DTYPE); kp2 = np.ascontiguousarray(kp2, KEYPOINT_DTYPE)
        desc1 = np.ascontiguousarray(desc1, np.uint8); desc2 = np.ascontiguousarray(desc2, np.uint8)
        cen = np.ascontiguousarray(centers, np.float32).copy()
        m12 = np.zeros(len(kp1), np.int32); n = C.c_int()
        check(load().orbm_search_window_host(self.h, _p(kp1), _p(desc1), len(kp1), _p(kp2), _p(desc2), len(kp2), _p(cen), _p(m12),
                                             C.byref(n), C.byref(params)))
        return n.value, m12, cen

    def search_groups_host(self, kp1, desc1, group1, kp2, desc2, group2, th_dist=50, nnratio=0.7, check_ori=True):
        """SearchByBoW-style matching restricted to equal group ids (uint16, 0xffff = none).  Returns (nmatches, matches12)."""
        kp1 = np.ascontiguousarray(kp1, KEYPOINT_DTYPE); kp2 = np.ascontiguousarray(kp2, KEYPOINT_DTYPE)
        desc1 = np.ascontiguousarray(desc1, np.uint8); desc2 = np.ascontiguousarray(desc2, np.uint8)
        g1 = np.ascontiguousarray(group1, np.uint16); g2 = np.ascontiguousarray(group2, np.uint16)
        m12 = np.zeros(len(kp1), np.int32); n = C.c_int()
        check(load().orbm_search_groups_host(self.h, _p(kp1), _p(desc1), _p(g1), len(kp1), _p(kp2), _p(desc2), _p(g2), len(kp2),
                                             _p(m12), C.byref(n), th_dist, nnratio, int(check_ori)))
        return n.value, m12

    def search_window_device(self, kps_ptr, desc_ptr, counts_ptr, capacity, pair_a_ptr, pair_b_ptr, npairs, centers_ptr,
                             matches_ptr, nmatches_ptr, params, ws_ptr, ws_bytes, stream=0):
        check(load().orbm_search_window_device(self.h, kps_ptr, desc_ptr, counts_ptr, capacity, pair_a_ptr, pair_b_ptr, npairs,
                                               centers_ptr, matches_ptr, nmatches_ptr, C.byref(params), ws_ptr, ws_bytes, stream))

    def set_profiling(self, on):
        check(load().orbm_set_profiling(self.h, int(on)))

    def knn2_times(self):
        a, b = C.c_float(), C.c_float()
        check(load().orbm_knn2_times(self.h, C.byref(a), C.byref(b)))
        return a.value, b.value

    @property
    def launches(self):
        return load().orbm_launch_count(self.h)


def popc_peak(device=0):
    a, b = C.c_double(), C.c_double()
    check(load().orbm_popc_peak(device, C.byref(a), C.byref(b)))
    return a.value, b.value
