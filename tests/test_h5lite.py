"""adorym_b200.h5lite: the reference's shipped HDF5 fixtures open without h5py and agree with the byte
offsets probed in SURVEY section 4.  The fixtures live in the reference tree, so these tests run in the build container."""
import os
import numpy as np
import pytest

from adorym_b200 import h5lite

REF = '/root/reference/demos'
FIX = {
    'adhesin/data_adhesin_360_soft_4d.h5': ('complex64', (500, 1, 64, 64), 2456),
    'cameraman_multidist/data_shift.h5': ('float32', (1, 4, 512, 512), 2432),
}


@pytest.mark.skipif(not os.path.isdir(REF), reason='reference tree not mounted (only present in the build container)')
@pytest.mark.parametrize('rel', sorted(FIX))
def test_reference_fixtures_open_without_h5py(rel):
    dt, shp, off = FIX[rel]
    path = os.path.join(REF, rel)
    f = h5lite.File(path, 'r')
    d = f['exchange/data']
    assert d.shape == shp and d.dtype == np.dtype(dt)
    ref = np.memmap(path, dtype=dt, mode='r', offset=off, shape=shp)
    ind = np.array([0, 2, 3]) if shp[1] > 3 else np.array([0])
    assert np.array_equal(d[0, ind], ref[0, ind])                   # the driver's per-minibatch fancy read
    assert np.array_equal(np.asarray(d[...]), np.asarray(ref))
    with pytest.raises(KeyError):
        f['metadata/theta']                                         # the driver relies on KeyError (ptychography.py:251-255)
    if 'metadata' in f:
        m = {k: f['metadata'][k][...] for k in f['metadata'].keys()}
        assert float(m['energy_ev']) == 17500. and float(m['psize_cm']) == 1e-4
        assert np.allclose(m['free_prop_cm'], [40., 60., 80., 100.])


@pytest.mark.skipif(not os.path.isdir(REF), reason='reference tree not mounted (only present in the build container)')
def test_driver_dataset_handle_reads_hdf5():
    from adorym_b200.ptychography import _Dataset
    ds = _Dataset('data_shift.h5', os.path.join(REF, 'cameraman_multidist'))
    assert ds.shape == (1, 4, 512, 512)
    assert float(ds.meta['energy_ev']) == 17500.


def test_not_hdf5_raises(tmp_path):
    p = tmp_path / 'x.h5'
    p.write_bytes(b'not an hdf5 file at all' * 10)
    with pytest.raises(IOError):
        h5lite.File(str(p), 'r')
