"""Oracle (test infrastructure): Keras inference semantics in torch-CPU fp32.

The reference's CNN forward is ``keras.Model.predict`` on a model loaded from
``keras_*_cnn.hdf5`` (segmenter.py:131-133,163).  Neither TensorFlow nor the
``.hdf5`` files exist in the build container and the architectures are not in
the reference source tree, so this interpreter restates the *published Keras
layer semantics* (channels_last) for the layer types a Sequential CNN of that
family uses, driven by the same ``model_config`` JSON a Keras HDF5 holds:

  InputLayer, Conv2D (valid/same, strides, bias), BatchNormalization
  (inference: gamma*(x-mean)/sqrt(var+eps)+beta), Activation / ReLU / Softmax,
  MaxPooling2D (valid/same), Dropout (identity), Flatten (H,W,C order),
  Dense (+ fused activation).

PARITY UNPINNED against TensorFlow (see oracle/__init__.py).  What it pins is
the arithmetic the CUDA kernels must reproduce for any weights: an fp32
torch-CPU evaluation, the "plain fp32 reference" for a floating-point kernel.
"""
import json

import numpy as np
import torch
import torch.nn.functional as F


def _act(x, name):
    if name in (None, 'linear'):
        return x
    if name == 'relu':
        return F.relu(x)
    if name == 'softmax':
        return F.softmax(x, dim=-1)
    if name == 'sigmoid':
        return torch.sigmoid(x)
    raise NotImplementedError('activation %r' % name)


def _same_pad(size, k, s):
    """TF/Keras 'same' padding (may be asymmetric: extra goes to the end)."""
    out = -(-size // s)
    total = max((out - 1) * s + k - size, 0)
    return total // 2, total - total // 2


class KerasLikeModel:
    """Evaluates a Keras-style Sequential config.  ``config`` is the decoded
    ``model_config`` dict (or JSON string); ``weights`` maps
    ``'<layer_name>/<var>'`` (kernel, bias, gamma, beta, moving_mean,
    moving_variance) to float32 arrays in Keras layouts (Conv2D kernel
    [kh,kw,cin,cout]; Dense kernel [in,out])."""

    def __init__(self, config, weights, threads=None):
        if isinstance(config, (str, bytes)):
            config = json.loads(config)
        cfg = config['config'] if 'config' in config else config
        self.layers = cfg['layers'] if isinstance(cfg, dict) else cfg
        self.w = {k: torch.from_numpy(np.ascontiguousarray(v, dtype=np.float32)) for k, v in weights.items()}
        self.threads = threads

    def _get(self, lname, var):
        for key in ('%s/%s' % (lname, var), '%s/%s:0' % (lname, var)):
            if key in self.w:
                return self.w[key]
        raise KeyError('%s/%s' % (lname, var))

    @torch.no_grad()
    def forward(self, x, capture=None):
        """x: float32 [n, H, W, C] (channels_last).  Returns float32 [n, K].
        ``capture``: optional dict filled with per-layer outputs (NHWC numpy)."""
        if self.threads:
            torch.set_num_threads(self.threads)
        t = torch.from_numpy(np.ascontiguousarray(x, dtype=np.float32))
        if t.dim() == 4:
            t = t.permute(0, 3, 1, 2).contiguous()      # NCHW for torch
        for layer in self.layers:
            cls, c = layer['class_name'], layer['config']
            name = c.get('name')
            if cls in ('InputLayer', 'Dropout', 'SpatialDropout2D', 'GaussianNoise'):
                pass
            elif cls == 'Conv2D':
                w = self._get(name, 'kernel').permute(3, 2, 0, 1).contiguous()
                b = self._get(name, 'bias') if c.get('use_bias', True) else None
                sh, sw = c.get('strides', (1, 1))
                if c.get('padding', 'valid') == 'same':
                    pt, pb = _same_pad(t.shape[2], w.shape[2], sh)
                    pl, pr = _same_pad(t.shape[3], w.shape[3], sw)
                    t = F.pad(t, (pl, pr, pt, pb))
                t = F.conv2d(t, w, b, stride=(sh, sw))
                t = _act(t.permute(0, 2, 3, 1), c.get('activation')).permute(0, 3, 1, 2)
            elif cls == 'BatchNormalization':
                eps = c.get('epsilon', 1e-3)
                g = self._get(name, 'gamma') if c.get('scale', True) else None
                be = self._get(name, 'beta') if c.get('center', True) else None
                mu, var = self._get(name, 'moving_mean'), self._get(name, 'moving_variance')
                inv = torch.rsqrt(var + eps)
                scale = inv * g if g is not None else inv
                shift = (be if be is not None else 0) - mu * scale
                shape = (1, -1, 1, 1) if t.dim() == 4 else (1, -1)
                t = t * scale.view(shape) + shift.view(shape)
            elif cls == 'Activation':
                t = _act(t, c['activation']) if t.dim() == 2 else \
                    _act(t.permute(0, 2, 3, 1), c['activation']).permute(0, 3, 1, 2)
            elif cls == 'ReLU':
                t = F.relu(t)
            elif cls == 'Softmax':
                t = F.softmax(t, dim=-1)
            elif cls == 'MaxPooling2D':
                ph, pw = c.get('pool_size', (2, 2))
                st = c.get('strides') or (ph, pw)
                if c.get('padding', 'valid') == 'same':
                    pt, pb = _same_pad(t.shape[2], ph, st[0])
                    pl, pr = _same_pad(t.shape[3], pw, st[1])
                    t = F.pad(t, (pl, pr, pt, pb), value=float('-inf'))
                t = F.max_pool2d(t, (ph, pw), stride=tuple(st))
            elif cls == 'Flatten':
                t = t.permute(0, 2, 3, 1).reshape(t.shape[0], -1)   # (H, W, C) order
            elif cls == 'Dense':
                t = t @ self._get(name, 'kernel')
                if c.get('use_bias', True):
                    t = t + self._get(name, 'bias')
                t = _act(t, c.get('activation'))
            else:
                raise NotImplementedError('Keras layer %s' % cls)
            if capture is not None:
                capture[name] = (t.permute(0, 2, 3, 1) if t.dim() == 4 else t).numpy().copy()
        return t.numpy()

    def predict(self, x, batch_size=1024, **_):
        """``keras.Model.predict`` stand-in (results independent of batch_size)."""
        outs = [self.forward(x[i:i + batch_size]) for i in range(0, len(x), batch_size)]
        return np.concatenate(outs) if outs else np.zeros((0, 0), np.float32)

    __call__ = predict
