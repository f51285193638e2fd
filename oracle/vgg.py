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
