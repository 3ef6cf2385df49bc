"""Oracle (test infrastructure): VoiceFemininityScoring glue (reference vbx_segmenter.py:28-61,
129-202) restated independently of the product: intervals handled with exhaustive, brute-force
definitions (pyannote.core is absent, so this restates its documented semantics: Timeline.crop =
intersection, duration = total length, Annotation keyed by (segment, track)); the MLP is evaluated
with numpy from the Keras config.  PARITY UNPINNED (interspeech2023_*.hdf5 absent)."""
import numpy as np


def mlp_numpy(config, weights, x):
    layers = config['config']['layers']
    y = np.asarray(x, dtype=np.float32)
    for l in layers:
        c, n = l['config'], l['config'].get('name')
        if l['class_name'] == 'Dense':
            y = y @ weights[n + '/kernel'] + (weights[n + '/bias'] if c.get('use_bias', True) else 0)
            act = c.get('activation')
        elif l['class_name'] == 'Activation':
            act = c['activation']
        elif l['class_name'] == 'BatchNormalization':
            inv = 1.0 / np.sqrt(weights[n + '/moving_variance'] + np.float32(c.get('epsilon', 1e-3)))
            y = (y - weights[n + '/moving_mean']) * inv * weights[n + '/gamma'] + weights[n + '/beta']
            act = None
        else:
            act = None
        if act == 'relu':
            y = np.maximum(y, 0)
        elif act == 'sigmoid':
            y = 1.0 / (1.0 + np.exp(-y))
        y = y.astype(np.float32)
    return y


def femininity(vad_seg, xvectors, mlp, vad_thresh):
    """vad_seg: [(label, start, stop)], xvectors: [(key, (start, stop), x)], mlp: callable."""
    speech = [(b, e) for lab, b, e in vad_seg if lab == 'speech']
    speech_duration = float(sum(e - b for b, e in speech))
    if not speech_duration:
        return None, speech_duration, 0
    kept, mid = [], []
    for key, (a, b), x in xvectors:
        m = (a + b) / 2
        if not any(s < m < e for s, e in speech):
            continue
        inter = 0.0
        for s, e in speech:
            lo, hi = max(a, s), min(b, e)
            if hi > lo:
                inter += hi - lo
        ratio = inter / (b - a)
        if ratio >= vad_thresh:
            kept.append((key, (a, b), x))
        mid.append((ratio, key, (a, b), x))
    need = round(0.5 * len(mid))
    if len(kept) < need:
        ranked = sorted(range(len(mid)), key=lambda i: mid[i][0])[::-1]      # decreasing overlap
        arr = np.array([mid[i][0] for i in range(len(mid))])
        ranked = list(np.argsort(arr)[::-1])                                 # numpy order, like the reference
        for i in ranked[len(kept):len(kept) + (need - len(kept))]:
            kept.append((mid[i][1], mid[i][2], mid[i][3]))
    p = mlp(np.asarray([x for _, _, x in kept]))
    if len(p) > 1:
        p = np.squeeze(p)
    votes = {}
    for (_, seg, _), pi in zip(kept, p):
        votes[seg] = bool(np.all(np.asarray(pi) >= 0.5))
    return sum(votes.values()) / len(votes), speech_duration, len(kept)
