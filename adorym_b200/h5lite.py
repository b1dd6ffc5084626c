"""Minimal read-only HDF5 reader for the measured-data files of the reference (`exchange/data` [n_theta, n_spots, py,
px] + optional `metadata/*`), so that reconstruct_ptychography opens them without h5py.

Supports what h5py writes by default and what the reference's data files use: superblock version 0/1, old-style
groups (B-tree v1 + local heap + symbol nodes), version-1 object headers with continuation blocks, simple dataspaces,
fixed-point / floating-point / compound (complex = {r, i}) datatypes, CONTIGUOUS or compact layout.  Chunked,
compressed or new-style (version-2 object header / link-message) files raise NotImplementedError: open those with
h5py.  A contiguous dataset is exposed as an np.memmap, so the per-minibatch fancy reads of
adorym/forward_model.py:113-119 (`prj[i_theta, ind]`) touch only the frames they need."""
import struct
import numpy as np

UNDEF = 0xFFFFFFFFFFFFFFFF


class H5LiteError(NotImplementedError):
    pass


class _Buf(object):
    def __init__(self, path):
        self.path = path
        self.mm = np.memmap(path, dtype='uint8', mode='r')

    def rd(self, off, n):
        return bytes(self.mm[off:off + n])

    def u(self, off, n):
        return int.from_bytes(self.rd(off, n), 'little')


def _parse_datatype(b, off=0):
    """-> (numpy dtype, bytes consumed)."""
    cls_ver = b[off]
    cls, ver = cls_ver & 0x0F, cls_ver >> 4
    bits0, bits1, bits2 = b[off + 1], b[off + 2], b[off + 3]
    size = struct.unpack_from('<I', b, off + 4)[0]
    endian = '>' if (bits0 & 1) else '<'
    if cls == 0:        # fixed point: properties = bit offset (2), precision (2)
        signed = bool(bits0 & 0x08)
        return np.dtype('%s%s%d' % (endian, 'i' if signed else 'u', size)), 8 + 4
    if cls == 1:        # floating point: properties 12 bytes
        return np.dtype('%sf%d' % (endian, size)), 8 + 12
    if cls == 6:        # compound
        nmemb = bits0 | (bits1 << 8)
        p = off + 8
        names, fmts, offs = [], [], []
        for _ in range(nmemb):
            e = b.index(b'\x00', p)
            name = b[p:e].decode()
            if ver < 3:
                p += ((e - p) // 8 + 1) * 8          # name null-padded to a multiple of 8 bytes
                moff = struct.unpack_from('<I', b, p)[0]
                p += 4
                if ver == 1:
                    p += 1 + 3 + 4 + 4 + 16             # dimensionality, reserved, permutation, reserved, 4 dim sizes
            else:
                p = e + 1
                nb = max(1, (size.bit_length() + 7) // 8)
                moff = int.from_bytes(b[p:p + nb], 'little')
                p += nb
            mdt, used = _parse_datatype(b, p)
            p += used
            names.append(name)
            fmts.append(mdt)
            offs.append(moff)
        dt = np.dtype({'names': names, 'formats': fmts, 'offsets': offs, 'itemsize': size})
        if [n.lower() for n in names] in (['r', 'i'], ['real', 'imag']) and fmts[0] == fmts[1] and fmts[0].kind == 'f' \
                and offs == [0, fmts[0].itemsize] and size == 2 * fmts[0].itemsize:
            dt = np.dtype('%sc%d' % (fmts[0].byteorder if fmts[0].byteorder in '<>' else '<', size))   # h5py's complex
        return dt, p - off
    raise H5LiteError('h5lite: datatype class %d is not supported' % cls)


class Dataset(object):
    def __init__(self, arr):
        self._a = arr
        self.shape = tuple(arr.shape)
        self.dtype = arr.dtype

    def __getitem__(self, key):
        return np.asarray(self._a[key])

    def __len__(self):
        return self.shape[0]

    def __array__(self, dtype=None, copy=None):
        a = np.asarray(self._a)
        return a.astype(dtype) if dtype is not None else a


class Group(object):
    def __init__(self, f, btree, heap):
        self._f, self._btree, self._heap = f, btree, heap
        self._links = None

    def _load(self):
        if self._links is not None:
            return self._links
        buf = self._f._buf
        if buf.rd(self._heap, 4) != b'HEAP':
            raise H5LiteError('h5lite: bad local heap')
        heap_data = buf.u(self._heap + 24, 8)
        links = {}

        def walk(addr):
            sig = buf.rd(addr, 4)
            if sig == b'TREE':
                ntype, level, n = buf.u(addr + 4, 1), buf.u(addr + 5, 1), buf.u(addr + 6, 2)
                if ntype != 0:
                    raise H5LiteError('h5lite: unexpected B-tree node type')
                p = addr + 8 + 16           # left / right sibling
                for i in range(n):
                    child = buf.u(p + 8 + i * 16, 8)      # key (8) child (8) key (8) ...
                    walk(child)
            elif sig == b'SNOD':
                n = buf.u(addr + 6, 2)
                p = addr + 8
                for i in range(n):
                    name_off = buf.u(p, 8)
                    ohdr = buf.u(p + 8, 8)
                    e = heap_data + name_off
                    name = b''
                    while True:
                        c = buf.rd(e, 1)
                        if c == b'\x00':
                            break
                        name += c
                        e += 1
                    links[name.decode()] = ohdr
                    p += 40
            else:
                raise H5LiteError('h5lite: unexpected group node %r' % sig)
        walk(self._btree)
        self._links = links
        return links

    def keys(self):
        return list(self._load().keys())

    def __contains__(self, name):
        try:
            self[name]
            return True
        except KeyError:
            return False

    def __getitem__(self, path):
        node = self
        for part in [p for p in path.split('/') if p]:
            if not isinstance(node, Group):
                raise KeyError(path)
            links = node._load()
            if part not in links:
                raise KeyError(path)
            node = node._f._object(links[part])
        return node


class File(Group):
    """h5lite.File(path)['exchange/data'] -> Dataset (shape, dtype, numpy indexing); ['metadata/x'][...] -> ndarray."""

    def __init__(self, path, mode='r', **kw):
        if mode != 'r' or kw.get('driver') is not None:
            raise H5LiteError('h5lite is read-only and serial')
        self._buf = _Buf(path)
        b = self._buf
        if b.rd(0, 8) != b'\x89HDF\r\n\x1a\n':
            raise IOError('%s is not an HDF5 file' % path)
        ver = b.u(8, 1)
        if ver > 1 or b.u(13, 1) != 8 or b.u(14, 1) != 8:
            raise H5LiteError('h5lite: superblock version %d / offset size %d (open this file with h5py)' % (ver, b.u(13, 1)))
        p = 24 + (4 if ver == 1 else 0)
        p += 32                               # base address, free-space address, end of file, driver info
        # root group symbol table entry: link name offset (8), object header (8), cache type (4), reserved (4), scratch (16)
        ohdr = b.u(p + 8, 8)
        cache = b.u(p + 16, 4)
        if cache == 1:
            Group.__init__(self, self, b.u(p + 24, 8), b.u(p + 32, 8))
        else:
            g = self._object(ohdr)
            Group.__init__(self, self, g._btree, g._heap)

    def close(self):
        pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        return False

    def _messages(self, addr):
        b = self._buf
        ver = b.u(addr, 1)
        if ver != 1:
            raise H5LiteError('h5lite: version-%d object header (open this file with h5py)' % ver)
        nmsg = b.u(addr + 2, 2)
        size = b.u(addr + 8, 4)
        blocks = [(addr + 16, size)]
        out = []
        while blocks and len(out) < nmsg:
            p, left = blocks.pop(0)
            while left >= 8 and len(out) < nmsg:
                mtype, msize, mflags = b.u(p, 2), b.u(p + 2, 2), b.u(p + 4, 1)
                body = b.rd(p + 8, msize)
                if mtype == 0x10:
                    blocks.append((int.from_bytes(body[0:8], 'little'), int.from_bytes(body[8:16], 'little')))
                out.append((mtype, body))
                p += 8 + msize
                left -= 8 + msize
        return out

    def _object(self, addr):
        msgs = self._messages(addr)
        d = dict()
        for t, body in msgs:
            d.setdefault(t, body)
        if 0x11 in d:                          # symbol table message: a group
            return Group(self, int.from_bytes(d[0x11][0:8], 'little'), int.from_bytes(d[0x11][8:16], 'little'))
        if 0x01 not in d or 0x03 not in d or 0x08 not in d:
            raise H5LiteError('h5lite: object at %d is neither an old-style group nor a simple dataset' % addr)
        sp = d[0x01]
        sver, rank = sp[0], sp[1]
        off = 8 if sver == 1 else 4
        shape = tuple(int.from_bytes(sp[off + 8 * i: off + 8 * i + 8], 'little') for i in range(rank))
        dt, _ = _parse_datatype(d[0x03])
        if 0x0B in d:
            raise H5LiteError('h5lite: filtered (compressed) datasets are not supported')
        lay = d[0x08]
        lver = lay[0]
        if lver == 3:
            lclass = lay[1]
            if lclass == 1:
                daddr = int.from_bytes(lay[2:10], 'little')
                if daddr == UNDEF:
                    return Dataset(np.zeros(shape, dtype=dt))
                arr = np.memmap(self._buf.path, dtype=dt, mode='r', offset=daddr, shape=shape) if len(shape) else \
                    np.frombuffer(self._buf.rd(daddr, dt.itemsize), dtype=dt)[0]
                return Dataset(np.asarray(arr)) if not len(shape) else Dataset(arr)
            if lclass == 0:
                n = int.from_bytes(lay[2:4], 'little')
                return Dataset(np.frombuffer(lay[4:4 + n], dtype=dt).reshape(shape).copy())
            raise H5LiteError('h5lite: chunked datasets are not supported (open this file with h5py)')
        if lver in (1, 2):
            rank_l, lclass = lay[1], lay[2]
            if lclass == 1:
                daddr = int.from_bytes(lay[8:16], 'little')
                return Dataset(np.memmap(self._buf.path, dtype=dt, mode='r', offset=daddr, shape=shape))
        raise H5LiteError('h5lite: data layout version %d' % lver)
