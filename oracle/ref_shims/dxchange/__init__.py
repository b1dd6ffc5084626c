"""TEST INFRASTRUCTURE ONLY -- stand-in for `dxchange` (TIFF I/O only, no arithmetic; the real
package is an un-vendored git dependency of the reference, pyproject.toml:29)."""
import os
import numpy as np


def read_tiff(path, *a, **kw):
    if path.endswith('.npy') or os.path.exists(path + '.npy'):
        return np.load(path if path.endswith('.npy') else path + '.npy')
    from PIL import Image, ImageSequence
    im = Image.open(path)
    frames = [np.array(fr) for fr in ImageSequence.Iterator(im)]
    return np.squeeze(np.stack(frames)) if len(frames) > 1 else frames[0]


def write_tiff(data, fname='tmp/data', dtype=None, overwrite=True, **kw):
    d = os.path.dirname(fname)
    if d and not os.path.exists(d):
        os.makedirs(d, exist_ok=True)
    a = np.asarray(data)
    np.save(fname + '.npy', a.astype(dtype) if dtype is not None else a)
