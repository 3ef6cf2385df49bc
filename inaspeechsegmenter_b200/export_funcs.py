"""Result exporters, byte-compatible with the reference's (export_funcs.py:29-39).

The reference delegates to pandas (``DataFrame.to_csv(sep='\\t', index=False)``)
and to pytextgrid; its tests compare output files with ``filecmp``
(run_test.py:112-127).  These writers reproduce those bytes directly: CSV
floats use Python's shortest round-trip repr (what pandas emits, e.g.
``29.080000000000002``), TextGrid numbers use ``%f`` and the exact layout of
``media/musanmix-smn-gender.TextGrid`` (including the trailing blank after
``tiers? <exists>``).
"""


def _csv_text(lseg):
    rows = ['labels\tstart\tstop']
    for label, start, stop in lseg:
        rows.append('%s\t%s\t%s' % (label, repr(float(start)), repr(float(stop))))
    return '\n'.join(rows) + '\n'


def seg2csv(lseg, fout=None):
    text = _csv_text(lseg)
    if fout is None:
        return text
    if hasattr(fout, 'write'):
        fout.write(text)
    else:
        with open(fout, 'w', newline='') as f:
            f.write(text)


def _textgrid_text(lseg):
    xmin, xmax = lseg[0][1], lseg[-1][2]
    out = ['File type = "ooTextFile"', 'Object class = "TextGrid"', '',
           'xmin = %f' % xmin, 'xmax = %f' % xmax, 'tiers? <exists> ', 'size = 1', 'item []:',
           '\titem [1]:', '\t\tclass = "IntervalTier"', '\t\tname = "inaSpeechSegmenter"',
           '\t\txmin = %f' % xmin, '\t\txmax = %f' % xmax, '\t\tintervals: size = %d' % len(lseg)]
    for i, (label, start, stop) in enumerate(lseg):
        out += ['\t\tintervals[%d]:' % (i + 1), '\t\t\t xmin = %f' % start, '\t\t\t xmax = %f' % stop,
                '\t\t\t text = "%s"' % label]
    return '\n'.join(out) + '\n'


def seg2textgrid(lseg, fout=None):
    text = _textgrid_text(lseg)
    if fout is None:
        return text
    if hasattr(fout, 'write'):
        fout.write(text)
    else:
        with open(fout, 'w', newline='') as f:
            f.write(text)
