"""TEST INFRASTRUCTURE ONLY (see oracle/__init__.py): numpy restatement of the reference's ResNet18 patch encoder
(models.py:13-77 with torchvision's BasicBlock: conv3x3-bn-relu-conv3x3-bn (+ downsample(x)) -relu), eval-mode
BatchNorm, in the reference's op order (convolution, THEN batch norm -- no folding).  Pinned against the reference itself:
tests/golden/make_golden_resnet.py imports /root/reference/models.py (timm stubbed) and stores its outputs.
"""
import numpy as np


def conv2d(x, w, stride, pad):
    """x [B, C, H, W], w [O, C, kh, kw] -> [B, O, Ho, Wo] (nn.Conv2d, bias=False)."""
    B, C, H, W = x.shape
    O, _, kh, kw = w.shape
    Ho, Wo = (H + 2 * pad - kh) // stride + 1, (W + 2 * pad - kw) // stride + 1
    xp = np.pad(x, ((0, 0), (0, 0), (pad, pad), (pad, pad)))
    cols = np.empty((B, C, kh, kw, Ho, Wo), dtype=x.dtype)
    for ky in range(kh):
        for kx in range(kw):
            cols[:, :, ky, kx] = xp[:, :, ky:ky + stride * Ho:stride, kx:kx + stride * Wo:stride]
    out = np.tensordot(w.reshape(O, -1), cols.reshape(B, C * kh * kw, Ho * Wo), axes=([1], [1]))      # [O, B, Ho*Wo]
    return out.transpose(1, 0, 2).reshape(B, O, Ho, Wo)


def batchnorm(x, p, prefix, eps=1e-5):
    """nn.BatchNorm2d in eval mode."""
    mean, var = p[prefix + ".running_mean"], p[prefix + ".running_var"]
    g, b = p[prefix + ".weight"], p[prefix + ".bias"]
    dt = x.dtype.type
    return (x - mean[None, :, None, None]) / np.sqrt(var[None, :, None, None] + dt(eps)) * g[None, :, None, None] + b[None, :, None, None]


def maxpool(x, k=3, stride=2, pad=1):
    B, C, H, W = x.shape
    Ho, Wo = (H + 2 * pad - k) // stride + 1, (W + 2 * pad - k) // stride + 1
    xp = np.pad(x, ((0, 0), (0, 0), (pad, pad), (pad, pad)), constant_values=-np.inf)
    out = np.full((B, C, Ho, Wo), -np.inf, dtype=x.dtype)
    for ky in range(k):
        for kx in range(k):
            out = np.maximum(out, xp[:, :, ky:ky + stride * Ho:stride, kx:kx + stride * Wo:stride])
    return out


def resnet18_features(p, x, dtype=np.float32, layers=(2, 2, 2, 2)):
    """models.py:54-73 up to the flatten: [B, 3, H, W] -> [B, 512]."""
    p = {k: np.asarray(v, dtype=dtype) for k, v in p.items() if not k.endswith("num_batches_tracked")}
    x = np.asarray(x, dtype=dtype)
    x = np.maximum(batchnorm(conv2d(x, p["conv1.weight"], 2, 3), p, "bn1"), 0)      # :55-57
    x = maxpool(x)                                                                   # :58
    for li, nb in enumerate(layers, start=1):                                        # :60-63
        for bi in range(nb):
            pre = f"layer{li}.{bi}"
            stride = 2 if (li > 1 and bi == 0) else 1
            identity = x
            out = np.maximum(batchnorm(conv2d(x, p[pre + ".conv1.weight"], stride, 1), p, pre + ".bn1"), 0)
            out = batchnorm(conv2d(out, p[pre + ".conv2.weight"], 1, 1), p, pre + ".bn2")
            if pre + ".downsample.0.weight" in p:
                identity = batchnorm(conv2d(x, p[pre + ".downsample.0.weight"], stride, 0), p, pre + ".downsample.1")
            x = np.maximum(out + identity, 0)
    return x.mean(axis=(2, 3))                                                       # :64-65


def resnet18_forward(p, x, dtype=np.float32):
    """... plus class_classifier (models.py:66) when the state dict has one."""
    f = resnet18_features(p, x, dtype)
    if "class_classifier.weight" in p:
        return f, f @ np.asarray(p["class_classifier.weight"], dtype).T + np.asarray(p["class_classifier.bias"], dtype)
    return f, None
