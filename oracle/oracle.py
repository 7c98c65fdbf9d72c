"""ctypes front end of the CPU ORACLE (oracle/orb_oracle.c).  TEST INFRASTRUCTURE ONLY.

May be imported only by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
--impl reference legs.  The product package never imports this module.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "liborb_oracle.so")

KEYPOINT_DTYPE = np.dtype([("x", "<f4"), ("y", "<f4"), ("size", "<f4"), ("angle", "<f4"),
                           ("response", "<f4"), ("octave", "<i4"), ("class_id", "<i4")])
CAND_DTYPE = np.dtype([("x", "<i2"), ("y", "<i2"), ("score", "<i4")])
assert KEYPOINT_DTYPE.itemsize == 28 and CAND_DTYPE.itemsize == 8


def build(force=False):
    src = os.path.join(_HERE, "orb_oracle.c")
    deps = [src, os.path.join(_HERE, "orb_oracle.h"), os.path.join(_HERE, "orb_pattern_31.inc")]
    if (not force and os.path.exists(_SO)
            and all(os.path.getmtime(_SO) >= os.path.getmtime(d) for d in deps)):
        return _SO
    os.makedirs(os.path.dirname(_SO), exist_ok=True)
    subprocess.check_call(["gcc", "-O3", "-fno-math-errno", "-ffp-contract=off", "-fPIC", "-std=gnu11", "-shared",
                           "-o", _SO, src, "-lm", "-lpthread"], cwd=_HERE)
    return _SO


def build_variant(out_dir, tag, flags):
    """The same restatement compiled with other optimisation flags ON THE BOX THAT RUNS IT (bench.py's cpu_baseline leg:
    -O2 as the reference's default build would, and -O3 -march=native) -- a timing aid, never a checker."""
    so = os.path.join(out_dir, "liborb_oracle_%s.so" % tag)
    subprocess.check_call(["gcc"] + list(flags) + ["-fno-math-errno", "-ffp-contract=off", "-fPIC", "-std=gnu11", "-shared",
                                                   "-o", so, os.path.join(_HERE, "orb_oracle.c"), "-lm", "-lpthread"], cwd=_HERE)
    L = C.CDLL(so)
    L.orbo_knn2_mt.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int]
    return L


_lib = None


def lib():
    global _lib
    if _lib is not None:
        return _lib
    L = C.CDLL(build())
    vp, i32, f32, sz = C.c_void_p, C.c_int, C.c_float, C.c_size_t
    L.orbo_create.restype = vp
    L.orbo_create.argtypes = [i32, f32, i32, i32, i32]
    L.orbo_destroy.argtypes = [vp]
    L.orbo_tables.argtypes = [vp] + [vp] * 6
    L.orbo_set_tiebreak.argtypes = [vp, i32]
    L.orbo_extract.restype = i32
    L.orbo_extract.argtypes = [vp, vp, i32, i32, sz, vp, vp, i32]
    L.orbo_level_dims.argtypes = [vp, i32, C.POINTER(i32), C.POINTER(i32)]
    L.orbo_level_pixels.restype = vp
    L.orbo_level_pixels.argtypes = [vp, i32]
    L.orbo_level_blurred.restype = vp
    L.orbo_level_blurred.argtypes = [vp, i32]
    L.orbo_level_candidates.restype = i32
    L.orbo_level_candidates.argtypes = [vp, i32, C.POINTER(vp)]
    L.orbo_level_kept.restype = i32
    L.orbo_level_kept.argtypes = [vp, i32, C.POINTER(vp)]
    L.orbo_level_retries.restype = i32
    L.orbo_level_retries.argtypes = [vp, i32]
    L.orbo_resize_linear_u8.argtypes = [vp, i32, i32, sz, vp, i32, i32, sz]
    L.orbo_gaussian7_s2_u8.argtypes = [vp, i32, i32, sz, vp, sz]
    L.orbo_fast9_nms.restype = i32
    L.orbo_fast9_nms.argtypes = [vp, i32, i32, sz, i32, vp, i32]
    L.orbo_fast9_score0.argtypes = [vp, i32, i32, sz, vp, sz]
    L.orbo_fast_atan2.restype = f32
    L.orbo_fast_atan2.argtypes = [f32, f32]
    L.orbo_cv_round.restype = i32
    L.orbo_cv_round.argtypes = [C.c_double]
    L.orbo_reflect101_border.argtypes = [vp, i32, i32, sz, vp, i32, sz]
    L.orbo_distribute_octree.restype = i32
    L.orbo_distribute_octree.argtypes = [vp, i32, i32, i32, i32, i32, vp, i32]
    L.orbo_ic_angle.restype = f32
    L.orbo_ic_angle.argtypes = [vp, sz, i32, i32, vp, C.POINTER(i32), C.POINTER(i32)]
    L.orbo_orb_descriptor.argtypes = [vp, sz, i32, i32, f32, vp]
    L.orbo_undistort_keypoints.argtypes = [vp, vp, i32, vp, vp, i32]
    L.orbo_cvt_gray_u8.argtypes = [vp, i32, i32, sz, i32, i32, vp, sz]
    L.orbo_descriptor_distance.restype = i32
    L.orbo_descriptor_distance.argtypes = [vp, vp]
    L.orbo_knn2.argtypes = [vp, i32, vp, i32, i32, vp, vp, vp]
    L.orbo_knn2_mt.argtypes = [vp, i32, vp, i32, i32, vp, vp, vp, i32]
    L.orbo_ratio_select.argtypes = [vp, vp, vp, i32, i32, f32, vp]
    L.orbo_merge_shards.argtypes = [vp, vp, vp, i32, i32, vp, vp, vp]
    L.orbo_search_groups.restype = i32
    L.orbo_search_groups.argtypes = [vp, vp, vp, i32, vp, vp, vp, i32, vp, i32, f32, i32]
    L.orbo_search_window.restype = i32
    L.orbo_search_window.argtypes = [vp, vp, i32, vp, vp, i32, vp, vp, vp]
    L.orbo_search_for_initialization.restype = i32
    L.orbo_search_for_initialization.argtypes = [vp, vp, i32, vp, vp, i32, vp, vp, i32, f32, i32, i32, i32, i32]
    L.orbo_extract_many.restype = i32
    L.orbo_extract_many.argtypes = [i32, f32, i32, i32, i32, vp, i32, i32, sz, i32, i32, vp]
    _lib = L
    return L


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def _u8(a):
    a = np.ascontiguousarray(a, dtype=np.uint8)
    return a


class OracleExtractor:
    """Mirror of ORBSlam::ORBextractor (include/ORBextractor.h:29-97) over the C oracle."""

    def __init__(self, nfeatures=1000, scale_factor=1.2, nlevels=8, ini_th=20, min_th=7):
        self.nfeatures, self.nlevels = nfeatures, nlevels
        self.h = lib().orbo_create(nfeatures, scale_factor, nlevels, ini_th, min_th)
        if not self.h:
            raise ValueError("bad extractor parameters")
        t = [np.zeros(nlevels, np.float32) for _ in range(4)] + [np.zeros(nlevels, np.int32), np.zeros(16, np.int32)]
        lib().orbo_tables(self.h, *[_p(a) for a in t])
        (self.scale_factors, self.inv_scale_factors, self.level_sigma2, self.inv_level_sigma2,
         self.features_per_level, self.umax) = t

    def __del__(self):
        if getattr(self, "h", None):
            lib().orbo_destroy(self.h)
            self.h = None

    def set_tiebreak(self, rule):
        lib().orbo_set_tiebreak(self.h, rule)

    def __call__(self, image):
        """-> (keypoints structured array, descriptors n x 32 u8)"""
        image = _u8(image)
        assert image.ndim == 2
        cap = self.nfeatures * 2 + 64 + 4 * image.shape[1] // max(1, image.shape[0]) * self.nlevels
        while True:
            kps = np.zeros(cap, KEYPOINT_DTYPE)
            desc = np.zeros((cap, 32), np.uint8)
            n = lib().orbo_extract(self.h, _p(image), image.shape[1], image.shape[0], image.strides[0],
                                   _p(kps), _p(desc), cap)
            if n >= 0:
                break
            cap *= 2
        return kps[:n].copy(), desc[:n].copy()

    # stage intermediates of the last call
    def level(self, l):
        w, h = C.c_int(), C.c_int()
        lib().orbo_level_dims(self.h, l, C.byref(w), C.byref(h))
        ptr = lib().orbo_level_pixels(self.h, l)
        return np.ctypeslib.as_array(C.cast(ptr, C.POINTER(C.c_uint8)), (h.value, w.value)).copy()

    def blurred(self, l):
        w, h = C.c_int(), C.c_int()
        lib().orbo_level_dims(self.h, l, C.byref(w), C.byref(h))
        ptr = lib().orbo_level_blurred(self.h, l)
        if not ptr:
            return None
        return np.ctypeslib.as_array(C.cast(ptr, C.POINTER(C.c_uint8)), (h.value, w.value)).copy()

    def _cands(self, fn, l):
        ptr = C.c_void_p()
        n = fn(self.h, l, C.byref(ptr))
        if n == 0:
            return np.zeros(0, CAND_DTYPE)
        buf = (C.c_char * (n * 8)).from_address(ptr.value)
        return np.frombuffer(buf, CAND_DTYPE, n).copy()

    def candidates(self, l):
        return self._cands(lib().orbo_level_candidates, l)

    def kept(self, l):
        return self._cands(lib().orbo_level_kept, l)

    def retries(self, l):
        return lib().orbo_level_retries(self.h, l)


def resize_linear(src, dw, dh):
    src = _u8(src)
    dst = np.zeros((dh, dw), np.uint8)
    lib().orbo_resize_linear_u8(_p(src), src.shape[1], src.shape[0], src.strides[0], _p(dst), dw, dh, dw)
    return dst


def gaussian7(src):
    src = _u8(src)
    dst = np.zeros_like(src)
    lib().orbo_gaussian7_s2_u8(_p(src), src.shape[1], src.shape[0], src.strides[0], _p(dst), dst.strides[0])
    return dst


def fast9_nms(img, threshold):
    img = _u8(img)
    cap = img.size // 4 + 16
    out = np.zeros(cap, CAND_DTYPE)
    n = lib().orbo_fast9_nms(_p(img), img.shape[1], img.shape[0], img.strides[0], threshold, _p(out), cap)
    return out[:n].copy()


def fast9_score0(img):
    img = _u8(img)
    sc = np.zeros(img.shape, np.int16)
    lib().orbo_fast9_score0(_p(img), img.shape[1], img.shape[0], img.strides[0], _p(sc), img.shape[1])
    return sc


def fast_atan2(y, x):
    return lib().orbo_fast_atan2(float(y), float(x))


def reflect101_border(src, border):
    src = _u8(src)
    dst = np.zeros((src.shape[0] + 2 * border, src.shape[1] + 2 * border), np.uint8)
    lib().orbo_reflect101_border(_p(src), src.shape[1], src.shape[0], src.strides[0], _p(dst), border, dst.strides[0])
    return dst


def distribute_octree(cands, width, height, N, tiebreak=0):
    cands = np.ascontiguousarray(cands, dtype=CAND_DTYPE)
    cap = len(cands) + 8
    out = np.zeros(cap, CAND_DTYPE)
    n = lib().orbo_distribute_octree(_p(cands), len(cands), width, height, N, tiebreak, _p(out), cap)
    return out[:n].copy()


def ic_angle(img, x, y, umax):
    img = _u8(img)
    umax = np.ascontiguousarray(umax, np.int32)
    m10, m01 = C.c_int(), C.c_int()
    a = lib().orbo_ic_angle(_p(img), img.strides[0], x, y, _p(umax), C.byref(m10), C.byref(m01))
    return a, m10.value, m01.value


def orb_descriptor(blurred, x, y, angle_deg):
    blurred = _u8(blurred)
    d = np.zeros(32, np.uint8)
    lib().orbo_orb_descriptor(_p(blurred), blurred.strides[0], x, y, float(angle_deg), _p(d))
    return d


def undistort_keypoints(kps, cam, dist, literal_bug=False):
    kps = np.ascontiguousarray(kps, KEYPOINT_DTYPE)
    out = np.zeros_like(kps)
    cam = np.ascontiguousarray(cam, np.float32); dist = np.ascontiguousarray(dist, np.float32)
    lib().orbo_undistort_keypoints(_p(kps), _p(out), len(kps), _p(cam), _p(dist), int(literal_bug))
    return out


def cvt_gray(img, rgb_order=False):
    img = _u8(img)
    h, w, c = img.shape
    dst = np.zeros((h, w), np.uint8)
    lib().orbo_cvt_gray_u8(_p(img), w, h, img.strides[0], c, int(rgb_order), _p(dst), w)
    return dst


def descriptor_distance(a, b):
    a, b = _u8(a), _u8(b)
    return lib().orbo_descriptor_distance(_p(a), _p(b))


def knn2(q, db, index_base=0, nthreads=1):
    q, db = _u8(q).reshape(-1, 32), _u8(db).reshape(-1, 32)
    nq = len(q)
    d1, idx1, d2 = (np.zeros(nq, np.int32) for _ in range(3))
    lib().orbo_knn2_mt(_p(q), nq, _p(db), len(db), index_base, _p(d1), _p(idx1), _p(d2), nthreads)
    return d1, idx1, d2


def ratio_select(d1, idx1, d2, th_low=50, ratio=0.7):
    d1, idx1, d2 = (np.ascontiguousarray(a, np.int32) for a in (d1, idx1, d2))
    m = np.zeros(len(d1), np.int32)
    lib().orbo_ratio_select(_p(d1), _p(idx1), _p(d2), len(d1), th_low, ratio, _p(m))
    return m


def merge_shards(d1, idx1, d2):
    """d1/idx1/d2: (nshards, nq) arrays, shards in ascending index-range order."""
    d1, idx1, d2 = (np.ascontiguousarray(a, np.int32) for a in (d1, idx1, d2))
    ns, nq = d1.shape
    o = [np.zeros(nq, np.int32) for _ in range(3)]
    lib().orbo_merge_shards(_p(d1), _p(idx1), _p(d2), ns, nq, *[_p(a) for a in o])
    return tuple(o)


def search_for_initialization(kp1, desc1, kp2, desc2, prev_matched, window=100, nnratio=0.9,
                              check_orientation=True, width=640, height=480, literal_bug=False):
    kp1 = np.ascontiguousarray(kp1, KEYPOINT_DTYPE)
    kp2 = np.ascontiguousarray(kp2, KEYPOINT_DTYPE)
    desc1, desc2 = _u8(desc1), _u8(desc2)
    prev = np.ascontiguousarray(prev_matched, np.float32).copy()
    m12 = np.zeros(len(kp1), np.int32)
    n = lib().orbo_search_for_initialization(_p(kp1), _p(desc1), len(kp1), _p(kp2), _p(desc2), len(kp2),
                                             _p(prev), _p(m12), window, nnratio, int(check_orientation),
                                             width, height, int(literal_bug))
    return n, m12, prev


class WindowParams(C.Structure):
    """orbo_window_params (same layout as orbm_window_params in include/orbx.h)."""
    _fields_ = [("radius", C.c_float), ("level_scale", C.c_float * 16),
                ("query_level_min", C.c_int32), ("query_level_max", C.c_int32),
                ("level_below", C.c_int32), ("level_above", C.c_int32), ("gate", C.c_int32), ("th_dist", C.c_int32),
                ("nnratio", C.c_float), ("check_orientation", C.c_int32), ("update_centers", C.c_int32),
                ("width", C.c_int32), ("height", C.c_int32), ("literal_gridid_bug", C.c_int32),
                ("use_bounds", C.c_int32), ("min_x", C.c_float), ("max_x", C.c_float), ("min_y", C.c_float), ("max_y", C.c_float)]


def window_params(radius, level_scale=None, query_levels=(0, 15), level_below=1, level_above=1, gate=1, th_dist=100,
                  nnratio=0.0, check_orientation=True, update_centers=False, width=640, height=480, literal_bug=False,
                  bounds=None):
    p = WindowParams()
    p.radius = radius
    ls = list(level_scale) if level_scale is not None else []
    for i in range(16):
        p.level_scale[i] = ls[i] if i < len(ls) else 1.0
    p.query_level_min, p.query_level_max = query_levels
    p.level_below, p.level_above, p.gate, p.th_dist, p.nnratio = level_below, level_above, gate, th_dist, nnratio
    p.check_orientation, p.update_centers = int(check_orientation), int(update_centers)
    p.width, p.height, p.literal_gridid_bug = width, height, int(literal_bug)
    if bounds is not None:                      # (min_x, max_x, min_y, max_y): Frame::FindimageBound for a distorted lens
        p.use_bounds = 1
        p.min_x, p.max_x, p.min_y, p.max_y = [float(v) for v in bounds]
    return p


def search_window(kp1, desc1, kp2, desc2, centers, params):
    kp1 = np.ascontiguousarray(kp1, KEYPOINT_DTYPE)
    kp2 = np.ascontiguousarray(kp2, KEYPOINT_DTYPE)
    desc1, desc2 = _u8(desc1), _u8(desc2)
    cen = np.ascontiguousarray(centers, np.float32).copy()
    m12 = np.zeros(len(kp1), np.int32)
    n = lib().orbo_search_window(_p(kp1), _p(desc1), len(kp1), _p(kp2), _p(desc2), len(kp2), _p(cen), _p(m12), C.byref(params))
    return n, m12, cen


def search_groups(kp1, desc1, group1, kp2, desc2, group2, th_dist=50, nnratio=0.7, check_orientation=True):
    kp1 = np.ascontiguousarray(kp1, KEYPOINT_DTYPE)
    kp2 = np.ascontiguousarray(kp2, KEYPOINT_DTYPE)
    desc1, desc2 = _u8(desc1), _u8(desc2)
    g1 = np.ascontiguousarray(group1, np.uint16); g2 = np.ascontiguousarray(group2, np.uint16)
    m12 = np.zeros(len(kp1), np.int32)
    n = lib().orbo_search_groups(_p(kp1), _p(desc1), _p(g1), len(kp1), _p(kp2), _p(desc2), _p(g2), len(kp2), _p(m12),
                                 th_dist, nnratio, int(check_orientation))
    return n, m12


def extract_many(imgs, nthreads, nfeatures=1000, scale_factor=1.2, nlevels=8, ini_th=20, min_th=7):
    imgs = _u8(imgs)
    nf, h, w = imgs.shape
    counts = np.zeros(nf, np.int32)
    lib().orbo_extract_many(nfeatures, scale_factor, nlevels, ini_th, min_th, _p(imgs), w, h, w * h, nf,
                            nthreads, _p(counts))
    return counts
