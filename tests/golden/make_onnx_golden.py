#!/usr/bin/env python
"""Generates tests/golden/resnet_small.onnx + resnet_small.npz by running the REAL reference
``resnet.py`` (loaded by file path from /root/reference) through torch's own ONNX exporter.

    python tests/golden/make_onnx_golden.py        # build container only

``final.onnx`` (the asset the reference's production backend loads, vbx_segmenter.py:249-266) is a
torch export of resnet.py's ResNet101; it is not available off-line, so the reader is pinned on a
small network of the same class (Bottleneck blocks [2, 1, 1, 1], m_channels 4, feat_dim 16, embed_dim 8,
randomised BatchNorm statistics) exported by the same exporter: same node pattern (BatchNorm folded
into Conv, biases de-duplicated through Identity nodes, Gemm head).  The .npz holds the state_dict and
the module's output on a seeded input, so the test can check the blob functionally.

The `onnx` Python package is absent here; the TorchScript exporter only needs it for an onnxscript
post-processing step that does nothing for this model, so that step is bypassed.
"""
import importlib.util
import os
import warnings

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))


def main():
    warnings.filterwarnings('ignore')
    from torch.onnx._internal.torchscript_exporter import onnx_proto_utils
    onnx_proto_utils._add_onnxscript_fn = lambda proto, custom_opsets: proto
    spec = importlib.util.spec_from_file_location('ref_resnet', '/root/reference/inaSpeechSegmenter/resnet.py')
    ref = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(ref)
    torch.manual_seed(20260923)
    net = ref.ResNet(ref.Bottleneck, [2, 1, 1, 1], m_channels=4, feat_dim=16, embed_dim=8).eval()
    g = torch.Generator().manual_seed(7)
    for mod in net.modules():
        if isinstance(mod, torch.nn.BatchNorm2d):
            mod.weight.data = 0.8 + 0.4 * torch.rand(mod.weight.shape, generator=g)
            mod.bias.data = 0.1 * torch.randn(mod.bias.shape, generator=g)
            mod.running_mean.data = 0.2 * torch.randn(mod.running_mean.shape, generator=g)
            mod.running_var.data = 0.5 + torch.rand(mod.running_var.shape, generator=g)
    x = torch.randn(1, 16, 40, generator=g)
    with torch.no_grad():
        y = net(x.clone()).numpy()
    path = os.path.join(HERE, 'resnet_small.onnx')
    torch.onnx.export(net, (x.clone(),), path, dynamo=False, input_names=['input'], output_names=['output'],
                      opset_version=11, dynamic_axes={'input': {2: 'T'}})
    sd = {k: v.numpy() for k, v in net.state_dict().items()}
    np.savez_compressed(os.path.join(HERE, 'resnet_small.npz'), x=x.numpy(), y=y, **{'sd:' + k: v for k, v in sd.items()})
    print('wrote %s (%d bytes), output %s' % (path, os.path.getsize(path), y.shape))


if __name__ == '__main__':
    main()
