"""nct_b200 -- Python host-side mirror of the reference's operator surface over libnct.so.

The product is `libnct.so` (hand-written sm_100a CUDA behind the C ABI of include/nct.h)
and the `neural_color_transfer` CLI built on it.  This module is the ctypes binding used
by tests/ and bench.py; method names follow the reference's functions
(NCT/GeneralizedPatchMatch.cuh:18-50, NCT/main.cu:47-454) so the parity tests read like
the reference's call sites.  torch is used only for device memory, streams and
torch.distributed plumbing.

There is no CPU fallback: importing works without a GPU (so the ABI can be checked), but
`Context()` raises if libnct.so is missing or no sm_100 device is present.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional, Sequence

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("NCT_LIB") or os.path.join(_HERE, "libnct.so")  # NCT_LIB: experiment builds only

c_ctx_p = C.c_void_p
_i, _f, _d, _p, _ll = C.c_int, C.c_float, C.c_double, C.c_void_p, C.c_longlong

class Config(C.Structure):
    """nct_config (CT/Config.h:55-98 + NCT/main.cu:64-83)"""
    _fields_ = [("bds_weight", _d), ("var_eps", _d), ("nonlocal_weight", _d), ("local_weight", _d), ("wls_lambda_init", _d),
                ("cluster_num", _i), ("k_num", _i), ("patch_size", _i), ("wls_alpha", _d), ("pm_iters", _i),
                ("kmeans_iters", _i), ("wls_rel_tol", _d), ("stop_after_level", _i), ("feature_store", _i)]


# name -> (restype, argtypes); every symbol include/nct.h declares must be listed here
# (tests/test_abi.py cross-checks this table against the header).
ABI = {
    "nct_create": (_i, [_i, C.POINTER(c_ctx_p)]),
    "nct_destroy": (_i, [c_ctx_p]),
    "nct_last_error": (C.c_char_p, [c_ctx_p]),
    "nct_set_stream": (_i, [c_ctx_p, _p]),
    "nct_get_stream": (_p, [c_ctx_p]),
    "nct_synchronize": (_i, [c_ctx_p]),
    "nct_version": (_i, []),
    "nct_launch_count": (_ll, [c_ctx_p]),
    "nct_reset_launch_count": (None, [c_ctx_p]),
    "nct_profile_enable": (_i, [c_ctx_p, _i]),
    "nct_profile_reset": (_i, [c_ctx_p]),
    "nct_profile_get": (_i, [c_ctx_p, _i, C.POINTER(_d), C.POINTER(_ll)]),
    "nct_profile_stage_name": (C.c_char_p, [_i]),
    "nct_debug_read_scratch": (_i, [c_ctx_p, C.c_char_p, _p, C.c_size_t]),
    "nct_chw_to_hwc": (_i, [c_ctx_p, _p, _p, _i, _i, _i]),
    "nct_hwc_to_chw": (_i, [c_ctx_p, _p, _p, _i, _i, _i]),
    "nct_l2norm": (_i, [c_ctx_p, _p, _p, _i, _i, _i]),
    "nct_l2norm_f16": (_i, [c_ctx_p, _p, _p, _i, _i, _i]),
    "nct_nnf_init": (_i, [c_ctx_p, _p, _i, _i, _i, _i]),
    "nct_nnf_upsample": (_i, [c_ctx_p, _p, _i, _i, _p, _i, _i, _i, _i]),
    "nct_patchmatch": (_i, [c_ctx_p, _p, _p, _p, _p, C.POINTER(_i)]),
    "nct_patchmatch_bidir": (_i, [c_ctx_p, _p, _p, _p, _p, _p, _p, C.POINTER(_i)]),
    "nct_patchmatch_f16": (_i, [c_ctx_p, _p, _p, _p, _p, C.POINTER(_i)]),
    "nct_patchmatch_bidir_f16": (_i, [c_ctx_p, _p, _p, _p, _p, _p, _p, C.POINTER(_i)]),
    "nct_xorwow_table": (_i, [c_ctx_p, _p, _i, _i]),
    "nct_patchmatch_stats": (_i, [c_ctx_p, C.POINTER(_ll)]),
    "nct_patchmatch_count_evals": (_i, [c_ctx_p, _i]),
    "nct_reconstruct_bds": (_i, [c_ctx_p, _p, _p, _p, _p, _i, _i, _i, _i, _d, _d, _p]),
    "nct_bds_feature_error": (_i, [c_ctx_p, _p, _p, _p, _p, _i, _i, _i, _i, _i, _f, _f, _p, _p]),
    "nct_bgr2lab_u8": (_i, [c_ctx_p, _p, _p, _i]),
    "nct_lab2bgr_u8": (_i, [c_ctx_p, _p, _p, _i]),
    "nct_resize_linear_u8c3": (_i, [c_ctx_p, _p, _i, _i, _p, _i, _i]),
    "nct_resize_linear_f64c3": (_i, [c_ctx_p, _p, _i, _i, _p, _i, _i]),
    "nct_local_fit": (_i, [c_ctx_p, _p, _p, _i, _i, _d, _p, _p]),
    "nct_confidence_weights": (_i, [c_ctx_p, _p, _i, _p]),
    "nct_solve_nonlocal": (_i, [c_ctx_p, _p, _p, _p, _p, _p, _p, _p, _i, _i, _i, _d, _d, _d, _i, _d, C.POINTER(_i)]),
    "nct_solve_ls_cg": (_i, [c_ctx_p, _i, _i, _p, _p, _p, _p, _p, _i, _d, _i, C.POINTER(_i)]),
    "nct_upsample_coefficients": (_i, [c_ctx_p, _p, _p, _i, _i, _p, _i, _i, _p, _p, _p]),
    "nct_solve_wls": (_i, [c_ctx_p, _p, _p, _p, _p, _i, _i, _d, _d, _d, _i, C.POINTER(_i), C.POINTER(_d)]),
    "nct_solve_wls_jacobi": (_i, [c_ctx_p, _p, _p, _p, _p, _i, _i, _d, _d, _d, _i, C.POINTER(_i), C.POINTER(_d)]),
    "nct_apply_coefficients": (_i, [c_ctx_p, _p, _p, _p, _i, _i, _p, _p]),
    "nct_vgg19_num_layers": (_i, []),
    "nct_vgg19_layer_name": (C.c_char_p, [_i]),
    "nct_vgg19_layer_shape": (_i, [_i, C.POINTER(_i), C.POINTER(_i)]),
    "nct_vgg19_set_weights": (_i, [c_ctx_p, _i, _p, _p]),
    "nct_vgg19_set_engine": (_i, [c_ctx_p, _i]),
    "nct_conv3x3_fixedpoint": (_i, [c_ctx_p, _p, _p, _p, _p, _p, _i, _i, _i, _i]),
    "nct_solve_direct": (_i, [c_ctx_p, _i, _i, _i, _p, _p, _p, C.POINTER(_p), C.POINTER(_p), _d, _i, C.POINTER(_i), C.POINTER(_d)]),
    "nct_probe_read_bandwidth": (_i, [c_ctx_p, C.c_size_t, _i, C.POINTER(C.c_double)]),
    "nct_vgg19_level_dims": (_i, [_i, _i, C.POINTER(_i * 3)]),
    "nct_vgg19_features": (_i, [c_ctx_p, _p, _i, _i, _i, C.POINTER(_p)]),
    "nct_config_default": (None, [C.POINTER(Config)]),
    "nct_transfer_pair_dev": (_i, [c_ctx_p, _p, _i, _i, _p, _i, _i, C.POINTER(Config), _p]),
    "nct_transfer_pair": (_i, [c_ctx_p, _p, _i, _i, _p, _i, _i, C.POINTER(Config), _p]),
    "nct_vgg19_load_caffemodel": (_i, [c_ctx_p, C.c_char_p]),
    "nct_png_read": (_i, [C.c_char_p, C.POINTER(C.POINTER(C.c_uint8)), C.POINTER(_i), C.POINTER(_i)]),
    "nct_png_free": (None, [C.POINTER(C.c_uint8)]),
    "nct_png_write": (_i, [C.c_char_p, _p, _i, _i]),
    "nct_run_pairs": (_i, [c_ctx_p, C.c_char_p, C.c_char_p, C.POINTER(Config), _i, _i, C.POINTER(_i)]),
    "nct_run_pairs_ex": (_i, [c_ctx_p, C.c_char_p, C.c_char_p, C.POINTER(Config), _i, _i, _i, C.POINTER(_i), C.POINTER(_i), C.POINTER(_i)]),
    "nct_set_vis": (_i, [c_ctx_p, C.c_char_p, C.c_char_p]),
    "nct_cluster_features": (_i, [c_ctx_p, _p, _i, _i, _i, _i, _i, _p]),
    "nct_find_knns": (_i, [c_ctx_p, _p, _i, _i, _i, _p, _i, _i, _i, _p, _p]),
    "nct_find_knns_brute": (_i, [c_ctx_p, _p, _i, _i, _i, _p, _i, _i, _i, _p, _p]),
}



_lib = None


class NctError(RuntimeError):
    pass


def load_library(path: Optional[str] = None):
    """dlopen libnct.so and bind every ABI symbol; raises NctError when the CUDA extension
    is missing (the product path never falls back to anything else)."""
    global _lib
    if _lib is not None and path is None:
        return _lib
    path = path or LIB_PATH
    if not os.path.exists(path):
        raise NctError(
            f"{path} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(nvcc, sm_100a). libnct has no CPU fallback."
        )
    lib = C.CDLL(path)
    for name, (res, args) in ABI.items():
        fn = getattr(lib, name)  # AttributeError => symbol missing: fail loudly
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def png_read(path):
    """imread stand-in of the CLI (no GPU needed): PNG file -> uint8 BGR numpy array"""
    import numpy as np
    lib = load_library()
    buf = C.POINTER(C.c_uint8)()
    h, w = _i(0), _i(0)
    rc = lib.nct_png_read(path.encode(), C.byref(buf), C.byref(h), C.byref(w))
    if rc != 0:
        raise NctError(f"cannot read PNG {path} ({rc})")
    out = np.ctypeslib.as_array(buf, shape=(h.value, w.value, 3)).copy()
    lib.nct_png_free(buf)
    return out


def png_write(path, bgr):
    import numpy as np
    lib = load_library()
    a = np.ascontiguousarray(bgr, np.uint8)
    rc = lib.nct_png_write(path.encode(), a.ctypes.data_as(C.c_void_p), a.shape[0], a.shape[1])
    if rc != 0:
        raise NctError(f"cannot write PNG {path} ({rc})")


def write_caffemodel(path, weights, v1=True):
    """Serialise {layer: (w OIHW, bias)} as a binary caffe NetParameter (V1 `layers` records like the original
    VGG_ILSVRC_19_layers.caffemodel, or V2 `layer` records).  Used to exercise the real loader with synthetic weights."""
    import numpy as np

    def varint(v):
        out = bytearray()
        while True:
            b = v & 0x7F
            v >>= 7
            if v:
                out.append(b | 0x80)
            else:
                out.append(b)
                return bytes(out)

    def field(num, wire, payload):
        return varint((num << 3) | wire) + (varint(len(payload)) + payload if wire == 2 else payload)

    def blob(arr):
        arr = np.ascontiguousarray(arr, np.float32)
        shape = list(arr.shape) if arr.ndim == 4 else [1, 1, 1, arr.size]
        if v1:
            head = b"".join(field(i + 1, 0, varint(s)) for i, s in enumerate(shape))
        else:
            dims = list(arr.shape)
            head = field(7, 2, field(1, 2, b"".join(varint(d) for d in dims)))
        return head + field(5, 2, arr.tobytes())

    out = bytearray(field(1, 2, b"VGG_ILSVRC_19_layers"))
    for name, (w, b) in weights.items():
        if v1:
            layer = field(4, 2, name.encode()) + field(5, 0, varint(4)) + field(6, 2, blob(w)) + field(6, 2, blob(b))
            out += field(2, 2, layer)
        else:
            layer = field(1, 2, name.encode()) + field(2, 2, b"Convolution") + field(7, 2, blob(w)) + field(7, 2, blob(b))
            out += field(100, 2, layer)
    with open(path, "wb") as f:
        f.write(bytes(out))


def make_params(C_, ah, aw, bh, bw, iters=10, rs_max=32, patch=3):
    """The reference's 11-int PatchMatch parameter block (NCT/main.cu:204-214)."""
    return (_i * 11)(C_, ah, aw, bh, bw, patch, iters, rs_max, 0, 10, 1)


def level_sizes(n: int, levels: int = 5):
    """Feature-map side lengths conv1_1..conv5_1 of an n-pixel side under Caffe's ceil-mode
    2x2/2 pooling (caffe/layers/pooling_layer.cpp:90-93): n -> ceil((n-2)/2)+1."""
    out = [n]
    for _ in range(levels - 1):
        n = -(-(n - 2) // 2) + 1
        out.append(n)
    return out


def _ptr(t):
    if t is None:
        return None
    return C.c_void_p(t.data_ptr())


class Context:
    """One libnct context = one GPU + one stream (reference: the process-wide CUDA device
    of NCT/main.cu:562-570)."""

    def __init__(self, gpu_id: int = 0, stream=None):
        self.lib = load_library()
        h = c_ctx_p()
        rc = self.lib.nct_create(gpu_id, C.byref(h))
        if rc != 0:
            raise NctError(f"nct_create({gpu_id}) failed with {rc} (no CUDA sm_100 device? libnct has no CPU fallback)")
        self.h = h
        self.gpu_id = gpu_id
        if stream is not None:
            self.set_stream(stream)

    # -- plumbing
    def _check(self, rc):
        if rc != 0:
            raise NctError(f"libnct error {rc}: {self.lib.nct_last_error(self.h).decode()}")

    def close(self):
        if getattr(self, "h", None):
            self.lib.nct_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_stream(self, stream):
        """stream: torch.cuda.Stream, raw cudaStream_t int, or None (ctx-owned stream)."""
        handle = getattr(stream, "cuda_stream", stream) or 0
        self._check(self.lib.nct_set_stream(self.h, C.c_void_p(handle)))

    def synchronize(self):
        self._check(self.lib.nct_synchronize(self.h))

    @property
    def launch_count(self):
        return int(self.lib.nct_launch_count(self.h))

    def reset_launch_count(self):
        self.lib.nct_reset_launch_count(self.h)

    def profile(self, enable=True):
        """stage events on the context's stream: True / 1 = every stage, 2 = the PatchMatch stage only, False = off"""
        self._check(self.lib.nct_profile_enable(self.h, int(enable)))
        self._check(self.lib.nct_profile_reset(self.h))

    def profile_report(self):
        """{stage name: (device ms, spans)} accumulated since the last profile()/reset"""
        out = {}
        for st in range(8):
            ms, n = _d(0), _ll(0)
            self._check(self.lib.nct_profile_get(self.h, st, C.byref(ms), C.byref(n)))
            out[self.lib.nct_profile_stage_name(st).decode()] = (ms.value, n.value)
        return out

    def read_scratch(self, name, dtype, count):
        """test hook: copy an internal scratch buffer to a numpy array"""
        import numpy as np
        out = np.empty(count, dtype)
        self._check(self.lib.nct_debug_read_scratch(self.h, name.encode(), out.ctypes.data_as(C.c_void_p), out.nbytes))
        return out

    # -- feature helpers
    def chw_to_hwc(self, src, dst=None):
        import torch
        Cn, H, W = src.shape
        dst = dst if dst is not None else torch.empty((H, W, Cn), dtype=torch.float32, device=src.device)
        self._check(self.lib.nct_chw_to_hwc(self.h, _ptr(src), _ptr(dst), Cn, H, W))
        return dst

    def hwc_to_chw(self, src, dst=None):
        import torch
        H, W, Cn = src.shape
        dst = dst if dst is not None else torch.empty((Cn, H, W), dtype=torch.float32, device=src.device)
        self._check(self.lib.nct_hwc_to_chw(self.h, _ptr(src), _ptr(dst), Cn, H, W))
        return dst

    def norm(self, src_hwc, dst=None):
        """norm(dst, src, NULL, dim) of NCT/GeneralizedPatchMatch.cu:237-283 (HWC layout)."""
        import torch
        H, W, Cn = src_hwc.shape
        dst = dst if dst is not None else torch.empty_like(src_hwc)
        self._check(self.lib.nct_l2norm(self.h, _ptr(src_hwc), _ptr(dst), Cn, H, W))
        return dst

    def norm_f16(self, src_hwc):
        """the same normalisation rounded to FP16 (torch.float16 tensor): the FP16 feature store of the PatchMatch volumes"""
        import torch
        H, W, Cn = src_hwc.shape
        dst = torch.empty((H, W, Cn), dtype=torch.float16, device=src_hwc.device)
        self._check(self.lib.nct_l2norm_f16(self.h, _ptr(src_hwc), _ptr(dst), Cn, H, W))
        return dst

    # -- NNF
    def init_ann(self, ann, ah, aw, bh, bw):
        """init_Ann_kernel (NCT/GeneralizedPatchMatch.cu:527-544); ann: int32/uint32 [ah*aw] cuda tensor."""
        self._check(self.lib.nct_nnf_init(self.h, _ptr(ann), ah, aw, bh, bw))
        return ann

    def upsample(self, ann_half, ah_half, aw_half, ann, ah, aw, bh, bw):
        """upSample_kernel + D2D copy (NCT/GeneralizedPatchMatch.cu:546-580, NCT/main.cu:238-250)."""
        self._check(self.lib.nct_nnf_upsample(self.h, _ptr(ann_half), ah_half, aw_half, _ptr(ann), ah, aw, bh, bw))
        return ann

    def patchmatch_single(self, a, b, ann, annd, params):
        """patchmatch_single<<<>>> (NCT/GeneralizedPatchMatch.cu:677-831); a, b L2-normalised HWC, float32 or (FP16
        feature store) float16."""
        import torch
        fn = self.lib.nct_patchmatch_f16 if a.dtype == torch.float16 else self.lib.nct_patchmatch
        if a.dtype != b.dtype:
            raise NctError("a and b must have the same dtype")
        self._check(fn(self.h, _ptr(a), _ptr(b), _ptr(ann), _ptr(annd), params))

    def patchmatch_bidir(self, a, b, ann, annd, bnn, bnnd, params_ab):
        """Both launches of NCT/main.cu:283-284 fused (float32 or float16 volumes)."""
        import torch
        fn = self.lib.nct_patchmatch_bidir_f16 if a.dtype == torch.float16 else self.lib.nct_patchmatch_bidir
        if a.dtype != b.dtype:
            raise NctError("a and b must have the same dtype")
        self._check(fn(self.h, _ptr(a), _ptr(b), _ptr(ann), _ptr(annd), _ptr(bnn), _ptr(bnnd), params_ab))

    def xorwow_table(self, ncols, ndraws):
        import torch
        out = torch.empty((ncols, ndraws), dtype=torch.float32, device=f"cuda:{self.gpu_id}")
        self._check(self.lib.nct_xorwow_table(self.h, _ptr(out), ncols, ndraws))
        return out

    def count_evals(self, enable=True):
        self._check(self.lib.nct_patchmatch_count_evals(self.h, 1 if enable else 0))

    def patchmatch_stats(self):
        st = (_ll * 2)()
        self._check(self.lib.nct_patchmatch_stats(self.h, st))
        return int(st[0]), int(st[1])

    # -- BDS votes
    def reconstruct_bds(self, a_img, b_img, ann, bnn, w_cohen=1.0, w_complete=2.0):
        """reconstruct_bds (NCT/GeneralizedPatchMatch.cu:122-235); uint8 BGR cuda tensors (H, W, 3)."""
        import torch
        ah, aw, _ = a_img.shape
        bh, bw, _ = b_img.shape
        out = torch.empty((ah, aw, 3), dtype=torch.uint8, device=a_img.device)
        self._check(self.lib.nct_reconstruct_bds(self.h, _ptr(a_img), _ptr(b_img), _ptr(ann), _ptr(bnn), ah, aw, bh, bw,
                                                 float(w_cohen), float(w_complete), _ptr(out)))
        return out

    def bds_feature_error(self, c_norm, s_raw, ann, bnn, w_cohen=1.0, w_complete=2.0, want_vote=False):
        """avg_vote_bds_a/_b/avg_vote_bds + norm + feature_distance (NCT/main.cu:297-318)."""
        import torch
        ah, aw, Cn = c_norm.shape
        bh, bw, _ = s_raw.shape
        err = torch.empty(ah * aw, dtype=torch.float32, device=c_norm.device)
        vote = torch.empty((ah, aw, Cn), dtype=torch.float32, device=c_norm.device) if want_vote else None
        self._check(self.lib.nct_bds_feature_error(self.h, _ptr(c_norm), _ptr(s_raw), _ptr(ann), _ptr(bnn), Cn, ah, aw, bh,
                                                   bw, float(w_cohen), float(w_complete), _ptr(err), _ptr(vote)))
        return (err, vote) if want_vote else err

    # -- colour space / resize (OpenCV stand-ins)
    def bgr2lab(self, img):
        import torch
        out = torch.empty_like(img)
        self._check(self.lib.nct_bgr2lab_u8(self.h, _ptr(img), _ptr(out), img.numel() // 3))
        return out

    def lab2bgr(self, img):
        import torch
        out = torch.empty_like(img)
        self._check(self.lib.nct_lab2bgr_u8(self.h, _ptr(img), _ptr(out), img.numel() // 3))
        return out

    def resize_linear(self, img, dh, dw):
        """cv::resize(..., INTER_LINEAR) for uint8 or float64 (H, W, 3) cuda tensors."""
        import torch
        sh, sw, _ = img.shape
        out = torch.empty((dh, dw, 3), dtype=img.dtype, device=img.device)
        fn = self.lib.nct_resize_linear_u8c3 if img.dtype == torch.uint8 else self.lib.nct_resize_linear_f64c3
        self._check(fn(self.h, _ptr(img), sh, sw, _ptr(out), dh, dw))
        return out

    # -- colour fit
    def local_fit(self, cnt_lab, stl_lab, eps=0.6):
        import torch
        h, w, _ = cnt_lab.shape
        a = torch.empty((h, w, 3), dtype=torch.float64, device=cnt_lab.device)
        b = torch.empty_like(a)
        self._check(self.lib.nct_local_fit(self.h, _ptr(cnt_lab), _ptr(stl_lab), h, w, float(eps), _ptr(a), _ptr(b)))
        return a, b

    def confidence_weights(self, err):
        import torch
        wgt = torch.empty(err.numel(), dtype=torch.float64, device=err.device)
        self._check(self.lib.nct_confidence_weights(self.h, _ptr(err), err.numel(), _ptr(wgt)))
        return wgt

    def solve_nonlocal(self, a, b, weight, cnt_lab, stl_lab, knn_id, knn_w, layer, local_weight=0.125, alpha=1.2,
                       nonlocal_weight=2.0, knum=8, d_weight=1.0, want_iters=False):
        """solve_nonlocal_downsample_gpu_gradient (CT/ColorTransfer.cpp:548-949); a, b updated in place."""
        h, w, _ = cnt_lab.shape
        its = (_i * 3)()
        self._check(self.lib.nct_solve_nonlocal(self.h, _ptr(a), _ptr(b), _ptr(weight), _ptr(cnt_lab), _ptr(stl_lab),
                                                _ptr(knn_id), _ptr(knn_w), h, w, layer, local_weight, alpha,
                                                nonlocal_weight, knum, d_weight, its if want_iters else None))
        return list(its) if want_iters else None

    def solve_ls_cg(self, size, constraints, A, columns, rowindex, x, b, tolerance=1e-6, maxitrs=100):
        """solve_ls_cg_gpu (CT/SparseSolver_GPU.cu:3-198) with its own argument list: HOST numpy arrays, one-based CSR
        (A float64[nnz], columns / rowindex int32), x float64[size] updated in place.  Returns the iteration count."""
        import numpy as np

        A = np.ascontiguousarray(A, np.float64)
        columns = np.ascontiguousarray(columns, np.int32)
        rowindex = np.ascontiguousarray(rowindex, np.int32)
        b = np.ascontiguousarray(b, np.float64)
        assert x.dtype == np.float64 and x.flags["C_CONTIGUOUS"] and x.shape == (size,)
        its = _i(0)
        self._check(self.lib.nct_solve_ls_cg(self.h, size, constraints, A.ctypes.data, columns.ctypes.data, rowindex.ctypes.data,
                                             x.ctypes.data, b.ctypes.data, int(A.shape[0]), tolerance, maxitrs, C.byref(its)))
        return int(its.value)

    def upsample_coefficients(self, a_lvl, b_lvl, cnt_lab_full):
        import torch
        h, w, _ = a_lvl.shape
        H, W, _ = cnt_lab_full.shape
        a = torch.empty((H, W, 3), dtype=torch.float64, device=a_lvl.device)
        b = torch.empty_like(a)
        rough = torch.empty((H, W), dtype=torch.float64, device=a_lvl.device)
        self._check(self.lib.nct_upsample_coefficients(self.h, _ptr(a_lvl), _ptr(b_lvl), h, w, _ptr(cnt_lab_full), H, W,
                                                       _ptr(a), _ptr(b), _ptr(rough)))
        return a, b, rough

    def solve_wls(self, a, b, rough, cnt_lab_full, lam, alpha=1.2, rel_tol=0.0, max_iters=0, jacobi=False):
        """solve_WLS_roughness_cpu (CT/ColorTransfer.cpp:951-1125); a, b updated in place. Returns (iters, rel_res)."""
        H, W, _ = cnt_lab_full.shape
        it = _i(0)
        res = _d(0.0)
        fn = self.lib.nct_solve_wls_jacobi if jacobi else self.lib.nct_solve_wls
        self._check(fn(self.h, _ptr(a), _ptr(b), _ptr(rough), _ptr(cnt_lab_full), H, W, lam, alpha,
                                           rel_tol, max_iters, C.byref(it), C.byref(res)))
        return it.value, res.value

    def solve_direct(self, A, row_index, columns, B, one_based=True, rel_tol=0.0, max_iters=0):
        """solve_direct_cpu (CT/SparseSolver_CPU.h:35-43): upper-triangular CSR of an SPD matrix (numpy host arrays), B = six
        right-hand sides (6, n).  Returns (X (6, n), iterations, relative residual)."""
        import numpy as np
        A = np.ascontiguousarray(A, np.float64)
        ri = np.ascontiguousarray(row_index, np.int32)
        co = np.ascontiguousarray(columns, np.int32)
        B = np.ascontiguousarray(B, np.float64)
        n = ri.size - 1
        X = np.empty((6, n), np.float64)
        bp = (_p * 6)(*[B[k].ctypes.data for k in range(6)])
        xp = (_p * 6)(*[X[k].ctypes.data for k in range(6)])
        it, res = _i(0), _d(0.0)
        self._check(self.lib.nct_solve_direct(self.h, int(A.size), n, 1 if one_based else 0, A.ctypes.data_as(C.c_void_p),
                                              ri.ctypes.data_as(C.c_void_p), co.ctypes.data_as(C.c_void_p), bp, xp, rel_tol, max_iters,
                                              C.byref(it), C.byref(res)))
        return X, it.value, res.value

    def apply_coefficients(self, cnt_lab_full, a, b, want_lab=False):
        import torch
        H, W, _ = cnt_lab_full.shape
        out = torch.empty((H, W, 3), dtype=torch.uint8, device=a.device)
        lab = torch.empty_like(out) if want_lab else None
        self._check(self.lib.nct_apply_coefficients(self.h, _ptr(cnt_lab_full), _ptr(a), _ptr(b), H, W, _ptr(out), _ptr(lab)))
        return (out, lab) if want_lab else out

    # -- clustering / non-local neighbours
    def cluster_features(self, feat_norm, k=10, iterations=11):
        """ColorTransfer::clusterFeastures (CT/ColorTransfer.cpp:355-395); feat_norm (h, w, C) float32 cuda."""
        import torch
        h, w, Cn = feat_norm.shape
        labels = torch.empty(h * w, dtype=torch.int32, device=feat_norm.device)
        self._check(self.lib.nct_cluster_features(self.h, _ptr(feat_norm), h, w, Cn, k, iterations, _ptr(labels)))
        return labels

    def find_knns(self, labels, lw, lh, lab, samples, nlabels=10, brute=False):
        """ColorTransfer::findKnns (CT/ColorTransfer.cpp:397-423); lab (h, w, 3) uint8 cuda."""
        import torch
        h, w, _ = lab.shape
        ids = torch.empty((h * w, 8), dtype=torch.int32, device=lab.device)
        wts = torch.empty((h * w, 8), dtype=torch.float64, device=lab.device)
        fn = self.lib.nct_find_knns_brute if brute else self.lib.nct_find_knns
        self._check(fn(self.h, _ptr(labels), lw, lh, nlabels, _ptr(lab), h, w, samples, _ptr(ids), _ptr(wts)))
        return ids, wts

    # -- VGG-19 (Classifier)
    def load_vgg19_weights(self, weights):
        """weights: dict layer name -> (w OIHW float32 numpy, bias float32 numpy), Caffe blob layout."""
        import numpy as np
        for i in range(self.lib.nct_vgg19_num_layers()):
            name = self.lib.nct_vgg19_layer_name(i).decode()
            w, b = weights[name]
            w = np.ascontiguousarray(w, np.float32)
            b = np.ascontiguousarray(b, np.float32)
            self._check(self.lib.nct_vgg19_set_weights(self.h, i, w.ctypes.data_as(C.c_void_p), b.ctypes.data_as(C.c_void_p)))

    def set_vgg_engine(self, engine):
        """0 = FP32 CUDA cores, 1 = tcgen05 TF32, 2 = tcgen05 3xTF32, 3 = tcgen05 INT8 exact fixed point"""
        self._check(self.lib.nct_vgg19_set_engine(self.h, engine))

    def conv3x3_fixedpoint(self, x, w_oihw, bias, debug_acc=False):
        """One 3x3 conv + bias + ReLU through the exact fixed-point tensor-core engine.  x: (H, W, Cin) float32 cuda
        tensor (>= 0); w_oihw, bias: numpy (Caffe blob layout).  Returns out (H, W, Cout) [, acc int32 (4, H*W, Cout)]."""
        import numpy as np
        import torch
        H, W, cin = x.shape
        w = np.ascontiguousarray(w_oihw, np.float32)
        b = np.ascontiguousarray(bias, np.float32)
        cout = w.shape[0]
        out = torch.empty((H, W, cout), dtype=torch.float32, device=x.device)
        acc = torch.zeros((4, H * W, cout), dtype=torch.int32, device=x.device) if debug_acc else None
        if acc is not None:
            torch.cuda.synchronize()
        self._check(self.lib.nct_conv3x3_fixedpoint(self.h, _ptr(x), w.ctypes.data_as(C.c_void_p), b.ctypes.data_as(C.c_void_p), _ptr(out),
                                                    _ptr(acc) if acc is not None else None, H, W, cin, cout))
        return (out, acc) if debug_acc else out

    def probe_read_bandwidth(self, nbytes, passes):
        """GB/s of coalesced 16-byte reads over an nbytes buffer (fits L2: the L2 roofline; >> L2: the HBM one)."""
        g = C.c_double(0.0)
        self._check(self.lib.nct_probe_read_bandwidth(self.h, nbytes, passes, C.byref(g)))
        return g.value

    def level_dims(self, h, w):
        d = ((_i * 3) * 5)()
        self.lib.nct_vgg19_level_dims(h, w, d)
        return [tuple(d[l]) for l in range(5)]

    def predict(self, img_bgr, deepest_level=0, out=None):
        """Classifier::Predict (NCT/Classifier.cpp:59-143): uint8 BGR (h, w, 3) cuda tensor -> list of 5 HWC float32
        feature tensors (levels deeper than `deepest_level` are None)."""
        import torch
        h, w, _ = img_bgr.shape
        dims = self.level_dims(h, w)
        feats = list(out) if out is not None else [None] * 5
        ptrs = (_p * 5)()
        for l in range(deepest_level, 5):
            Cn, fh, fw = dims[l]
            if feats[l] is None:
                feats[l] = torch.empty((fh, fw, Cn), dtype=torch.float32, device=img_bgr.device)
            ptrs[l] = feats[l].data_ptr()
        self._check(self.lib.nct_vgg19_features(self.h, _ptr(img_bgr), h, w, deepest_level, ptrs))
        return feats

    # -- per-pair pipeline
    def default_config(self, **overrides):
        cfg = Config()
        self.lib.nct_config_default(C.byref(cfg))
        for k, v in overrides.items():
            setattr(cfg, k, v)
        return cfg

    def transfer_pair_dev(self, cnt, stl, cfg=None, out=None):
        """transfer_color_single_bds (NCT/main.cu:47-454) on device-resident uint8 BGR tensors."""
        import torch
        ch, cw, _ = cnt.shape
        sh, sw, _ = stl.shape
        out = out if out is not None else torch.empty_like(cnt)
        self._check(self.lib.nct_transfer_pair_dev(self.h, _ptr(cnt), ch, cw, _ptr(stl), sh, sw,
                                                   C.byref(cfg) if cfg is not None else None, _ptr(out)))
        return out

    def transfer_pair(self, cnt_host, stl_host, cfg=None, out_host=None):
        """Host-buffer entry point (what the CLI calls): uint8 BGR host tensors / numpy arrays (pinned preferred)."""
        import numpy as np
        import torch
        def hp(x):
            return C.c_void_p(x.data_ptr()) if isinstance(x, torch.Tensor) else x.ctypes.data_as(C.c_void_p)
        ch, cw, _ = cnt_host.shape
        sh, sw, _ = stl_host.shape
        if out_host is None:
            out_host = np.empty((ch, cw, 3), np.uint8)
        self._check(self.lib.nct_transfer_pair(self.h, hp(cnt_host), ch, cw, hp(stl_host), sh, sw,
                                               C.byref(cfg) if cfg is not None else None, hp(out_host)))
        return out_host

    # -- files
    def load_caffemodel(self, path):
        """Net::CopyTrainedLayersFrom (caffe/net.cpp:798): load conv1_1..conv5_1 from a binary .caffemodel"""
        self._check(self.lib.nct_vgg19_load_caffemodel(self.h, path.encode()))

    def run_pairs(self, input_dir, output_dir, cfg=None, rank=0, world=1):
        """transfer_single (NCT/main.cu:456-543) for the lines i % world == rank of <input_dir>/pairs.txt"""
        done = _i(0)
        self._check(self.lib.nct_run_pairs(self.h, input_dir.encode(), output_dir.encode(),
                                           C.byref(cfg) if cfg is not None else None, rank, world, C.byref(done)))
        return done.value

    def run_pairs_ex(self, input_dir, output_dir, cfg=None, rank=0, world=1, resume=False, vis=False):
        """nct_run_pairs_ex: the same with resume-by-existing-output / ENABLE_VIS artefacts; returns (done, failed, skipped)"""
        done, failed, skipped = _i(0), _i(0), _i(0)
        self._check(self.lib.nct_run_pairs_ex(self.h, input_dir.encode(), output_dir.encode(), C.byref(cfg) if cfg is not None else None,
                                              rank, world, (1 if resume else 0) | (2 if vis else 0), C.byref(done), C.byref(failed),
                                              C.byref(skipped)))
        return done.value, failed.value, skipped.value

    def set_vis(self, directory, prefix="pair"):
        """per-level debug artefacts of the reference's ENABLE_VIS build; directory None switches them off"""
        self._check(self.lib.nct_set_vis(self.h, directory.encode() if directory else None, prefix.encode() if prefix else None))
