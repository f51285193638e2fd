"""oracle/vgg.py -- CPU restatement of the VGG-19 trunk the reference runs through Caffe.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

Follows
  preprocessing    NCT/Classifier.cpp:211-275 (8U BGR -> float, minus mean (103.939, 116.779, 123.68), planar)
  graph            demo/model/vgg19/VGG_ILSVRC_19_layers_deploy.prototxt (conv 3x3 pad 1 stride 1, in-place ReLU,
                   2x2/2 MAX pool), up to conv5_1
  conv             caffe/layers/base_conv_layer.cpp:257-282 (im2col + SGEMM) -- cross-correlation, weights OIHW
  pool             caffe/layers/pooling_layer.cpp:86-170 (ceil mode, window clipped at the border)
  relu             caffe/layers/relu_layer.cpp:9-19

FP32 throughout (torch CPU; its convolution's summation order is not Caffe's, so conv parity is a tolerance test
like Caffe's own, caffe/test/test_convolution_layer.cpp:231-265, 1e-4).

`features_canonical` runs the same graph through oracle/conv_oracle.c, whose summation order is DEFINED (tap-major,
channel-minor, one fmaf chain per output) -- the bit-exact target of the product's FP32 convolution engine."""
from __future__ import annotations

import numpy as np

from .synth import VGG19_TRUNK

MEAN_BGR = (103.939, 116.779, 123.68)
LEVEL_OF = {"conv5_1": 0, "conv4_1": 1, "conv3_1": 2, "conv2_1": 3, "conv1_1": 4}


def features(img_bgr_u8, weights, deepest_level=0, im2col=False):
    """Returns [f0..f4] HWC float32 numpy arrays (None for levels deeper than requested)."""
    import torch
    import torch.nn.functional as F

    x = torch.from_numpy(np.ascontiguousarray(img_bgr_u8)).to(torch.float32)
    x = x - torch.tensor(MEAN_BGR, dtype=torch.float32)
    x = x.permute(2, 0, 1).unsqueeze(0).contiguous()
    out = [None] * 5
    with torch.no_grad():
        for name, cin, cout, pool_before in VGG19_TRUNK:
            if pool_before:
                x = F.max_pool2d(x, 2, 2, ceil_mode=True)
            w, b = weights[name]
            wt, bt = torch.from_numpy(w), torch.from_numpy(b)
            if im2col:  # the "Caffe CPU path": im2col + SGEMM (base_conv_layer.cpp:257-282)
                n, c, h, ww = x.shape
                cols = F.unfold(x, 3, padding=1)[0]
                x = (wt.reshape(cout, -1) @ cols + bt[:, None]).reshape(1, cout, h, ww)
            else:
                x = F.conv2d(x, wt, bt, padding=1)
            x = torch.relu(x)
            if name in LEVEL_OF:
                lvl = LEVEL_OF[name]
                out[lvl] = np.ascontiguousarray(x[0].permute(1, 2, 0).numpy())
                if lvl == deepest_level:
                    break
    return out


def features_canonical(img_bgr_u8, weights, deepest_level=0):
    """Same graph, canonical summation order (oracle/conv_oracle.c).  Returns [f0..f4] HWC float32 (None below
    deepest_level)."""
    from .pm import lib

    L = lib()
    img = np.ascontiguousarray(img_bgr_u8, dtype=np.uint8)
    H, W, _ = img.shape
    x = np.empty((H, W, 3), np.float32)
    L.orc_preprocess_bgr(img, x, H * W)
    out = [None] * 5
    for name, cin, cout, pool_before in VGG19_TRUNK:
        if pool_before:
            Ho, Wo = (H - 2 + 1) // 2 + 1, (W - 2 + 1) // 2 + 1
            y = np.empty((Ho, Wo, cin), np.float32)
            L.orc_maxpool2x2_ceil(x, y, H, W, cin)
            x, H, W = y, Ho, Wo
        w, b = weights[name]
        y = np.empty((H, W, cout), np.float32)
        L.orc_conv3x3_relu_canon(x, np.ascontiguousarray(w, dtype=np.float32), np.ascontiguousarray(b, dtype=np.float32), y, H, W, cin, cout)
        x = y
        if name in LEVEL_OF:
            lvl = LEVEL_OF[name]
            out[lvl] = x
            if lvl == deepest_level:
                break
    return out


# ------------------------------------------------------------------------------------------------------------------
# Fixed-point ("Q") convolution: the DEFINED arithmetic of the product's exact tensor-core engine (engine 3).
#
# The reference's conv numerics are unpinned (cuDNN picks an algorithm at run time; Caffe's own tolerance is 1e-4,
# caffe/test/test_convolution_layer.cpp:231-265).  A tensor core's internal accumulation order is not specified either,
# so a floating-point tensor-core conv can never be bit-exact against anything.  Integer MMAs with INT32 accumulation
# ARE exact, hence order-independent.  The Q convolution therefore defines, per layer (conv1_2 .. conv5_1; conv1_1 with
# Cin = 3 keeps the canonical FP32 order of conv_oracle.c):
#
#   E      = exponent with max(X) < 2^E (from the FP32 exponent field of the tensor's maximum; X >= 0: post-ReLU)
#   xq     = floor(x * 2^(31 - E))                   in [0, 2^31)
#   Ew[o]  = the same exponent for max |W[o, :, :, :]|, per output channel
#   wq     = rint(w * 2^(22 - Ew[o]))  (ties to even) in [-2^22, 2^22]
#   xq = sum_i d_i 256^(3-i), wq = sum_j e_j 256^(2-j): balanced base-256 digits (d_i, e_j in [-128, 127] except the
#        leading ones, d_0 in [0, 128], e_0 in [-64, 64]) -- zero-mean digits make the dropped cross terms unbiased
#   S      = sum over taps and channels of  sum_{i+j <= 3} d_i e_j 256^(3-i-j)        (exact integer, |S| < 2^53)
#   out    = max(fl32(S * 2^(E + Ew[o] - 37)) + bias[o], 0)     (one rounding to FP32, then one FP32 add)
#
# Measured against an FP64 convolution of the same inputs (tests/test_oracle_vgg_q.py): max error 3-6e-7 of the layer's
# range, below the canonical FP32 order's own 1-2.5e-6.
QX_DIGITS, QW_DIGITS, QW_BITS, Q_DMAX = 4, 3, 22, 3


def q_exponent(m):
    """E with m < 2^E, read off the float32 exponent field (zero / denormal maximum: E = -126)."""
    bits = int(np.float32(m).view(np.uint32))
    return ((bits >> 23) & 0xFF) - 126


def _balanced_digits(q, n):
    """q: int64 array -> n balanced base-256 digits, most significant first (the leading one takes the carry)."""
    ds, r = [], q
    for _ in range(n - 1):
        d = ((r + 128) & 0xFF) - 128
        ds.append(d)
        r = (r - d) >> 8
    ds.append(r)
    return ds[::-1]


def q_act_digits(x):
    """x: float32 >= 0 -> (E, [d0..d3] int64 arrays)."""
    E = q_exponent(x.max()) if x.size else -126
    xq = np.floor(x.astype(np.float64) * 2.0 ** (8 * QX_DIGITS - 1 - E)).astype(np.int64)
    return E, _balanced_digits(xq, QX_DIGITS)


def q_weight_digits(w_oihw):
    """-> (Ew int array [cout], [e0..e2] int64 arrays shaped like w)."""
    cout = w_oihw.shape[0]
    Ew = np.array([q_exponent(m) for m in np.abs(w_oihw.reshape(cout, -1)).max(1)], dtype=np.int64)
    wq = np.rint(w_oihw.astype(np.float64) * (2.0 ** (QW_BITS - Ew))[:, None, None, None]).astype(np.int64)
    return Ew, _balanced_digits(wq, QW_DIGITS)


def q_conv3x3_relu(x_hwc, w_oihw, bias, return_acc=False):
    """The Q convolution of one layer (see above).  x_hwc float32 >= 0 (H, W, Cin) -> float32 (H, W, Cout)."""
    import torch
    import torch.nn.functional as F

    E, xd = q_act_digits(x_hwc)
    Ew, wd = q_weight_digits(w_oihw)
    S = None
    accs = []
    for i in range(QX_DIGITS):
        # digit d_i meets  sum_{j: i+j <= DMAX} e_j 256^(DMAX-i-j)  -- every partial sum stays an exact float64 integer
        wsum = np.zeros(w_oihw.shape, np.float64)
        for j in range(QW_DIGITS):
            if i + j <= Q_DMAX:
                wsum += wd[j].astype(np.float64) * 256.0 ** (Q_DMAX - i - j)
        xt = torch.from_numpy(xd[i].astype(np.float64)).permute(2, 0, 1).unsqueeze(0)
        c = F.conv2d(xt, torch.from_numpy(wsum), None, padding=1)[0].permute(1, 2, 0).numpy()
        S = c if S is None else S + c
    assert np.abs(S).max() < 2.0 ** 53
    if return_acc:  # the four INT32 accumulators the tensor core holds (debugging aid): acc[d] = sum_{i+j=d} d_i * e_j
        for d in range(Q_DMAX + 1):
            a = None
            for i in range(QX_DIGITS):
                j = d - i
                if 0 <= j < QW_DIGITS:
                    xt = torch.from_numpy(xd[i].astype(np.float64)).permute(2, 0, 1).unsqueeze(0)
                    c = F.conv2d(xt, torch.from_numpy(wd[j].astype(np.float64)), None, padding=1)[0].permute(1, 2, 0).numpy()
                    a = c if a is None else a + c
            accs.append(a.astype(np.int64))
    sh = E + Ew - (8 * QX_DIGITS - 1) - QW_BITS + 8 * (QX_DIGITS - 1 + QW_DIGITS - 1 - Q_DMAX)
    y = (S * (2.0 ** sh)[None, None, :]).astype(np.float32) + np.asarray(bias, np.float32)[None, None, :]
    y = np.ascontiguousarray(np.maximum(y, np.float32(0)))
    return (y, accs) if return_acc else y


def features_fixedpoint(img_bgr_u8, weights, deepest_level=0):
    """The trunk with conv1_1 in the canonical FP32 order and conv1_2 .. conv5_1 as Q convolutions: the bit-exact target
    of the product's engine 3 (tcgen05 kind::i8).  Returns [f0..f4] HWC float32 (None below deepest_level)."""
    from .pm import lib

    L = lib()
    img = np.ascontiguousarray(img_bgr_u8, dtype=np.uint8)
    H, W, _ = img.shape
    x = np.empty((H, W, 3), np.float32)
    L.orc_preprocess_bgr(img, x, H * W)
    out = [None] * 5
    for name, cin, cout, pool_before in VGG19_TRUNK:
        if pool_before:
            Ho, Wo = (H - 2 + 1) // 2 + 1, (W - 2 + 1) // 2 + 1
            y = np.empty((Ho, Wo, cin), np.float32)
            L.orc_maxpool2x2_ceil(x, y, H, W, cin)
            x, H, W = y, Ho, Wo
        w, b = weights[name]
        if cin < 64:
            y = np.empty((H, W, cout), np.float32)
            L.orc_conv3x3_relu_canon(x, np.ascontiguousarray(w, dtype=np.float32), np.ascontiguousarray(b, dtype=np.float32), y, H, W, cin, cout)
        else:
            y = q_conv3x3_relu(x, w, b)
        x = y
        if name in LEVEL_OF:
            lvl = LEVEL_OF[name]
            out[lvl] = x
            if lvl == deepest_level:
                break
    return out
