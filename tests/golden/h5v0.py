"""Minimal pure-Python reader for the HDF5 files the reference ships as golden outputs of its own regression tests
(`tests/models_basic/*/*_ref.out`: superblock version 0, version-1 object headers, group B-trees + local heaps, contiguous
little-endian datasets, scalar / 1-D attributes).  There is no libhdf5 / h5py in this image; this covers exactly what those
files use and raises on anything else.  Test infrastructure only.

    f = H5File(path)
    f.attrs('/')                  -> {'Iterations': ..., 'dt': ..., ...}
    f.keys('/rxs/rx1')            -> ['Ex', 'Ey', ...]
    f.dataset('/rxs/rx1/Ez')      -> numpy array
"""
import struct

import numpy as np

SIGNATURE = b'\x89HDF\r\n\x1a\n'


class H5Error(ValueError):
    pass


class H5File(object):
    def __init__(self, path):
        with open(path, 'rb') as f:
            self.b = f.read()
        b = self.b
        if b[:8] != SIGNATURE:
            raise H5Error('not an HDF5 file')
        if b[8] != 0:
            raise H5Error('superblock version {} not supported'.format(b[8]))
        self.O, self.L = b[13], b[14]          # size of offsets / lengths
        if self.O != 8 or self.L != 8:
            raise H5Error('only 8-byte offsets and lengths are supported')
        # 8 signature + 8 version bytes + 2+2 (group K) + 4 (flags) = 24, then base, free-space, eof, driver addresses
        base, = struct.unpack_from('<Q', b, 24)
        if base != 0:
            raise H5Error('non-zero base address')
        root_entry = 24 + 4 * self.O
        self.root = self._symtab_entry(root_entry)['header']

    # ------------------------------------------------------------------ low level
    def _symtab_entry(self, off):
        name_off, header, cache = struct.unpack_from('<QQI', self.b, off)
        e = {'name_off': name_off, 'header': header, 'cache': cache}
        if cache == 1:
            e['btree'], e['heap'] = struct.unpack_from('<QQ', self.b, off + 24)
        return e

    def _messages(self, addr):
        """(type, flags, payload bytes) of every message of a version-1 object header, continuation blocks included."""
        b = self.b
        version, _, nmsg, _refs, size = struct.unpack_from('<BBHII', b, addr)
        if version != 1:
            raise H5Error('object header version {} not supported'.format(version))
        blocks = [(addr + 16, size)]
        out = []
        while blocks and len(out) < nmsg:
            pos, length = blocks.pop(0)
            end = pos + length
            while pos + 8 <= end and len(out) < nmsg:
                mtype, msize, mflags = struct.unpack_from('<HHB', b, pos)
                payload = b[pos + 8:pos + 8 + msize]
                pos += 8 + msize
                if mtype == 0x0010:      # continuation
                    coff, clen = struct.unpack_from('<QQ', payload, 0)
                    blocks.append((coff, clen))
                out.append((mtype, mflags, payload))
        return out

    def _heap_name(self, heap, off):
        b = self.b
        if b[heap:heap + 4] != b'HEAP':
            raise H5Error('bad local heap')
        data, = struct.unpack_from('<Q', b, heap + 8 + 2 * self.L)
        end = b.index(b'\0', data + off)
        return b[data + off:end].decode('utf-8')

    def _group_entries(self, btree, heap):
        """{name: object header address} of a group (B-tree version 1, node type 0)."""
        b = self.b
        out = {}
        if b[btree:btree + 4] != b'TREE':
            raise H5Error('bad B-tree node')
        ntype, level, used = struct.unpack_from('<BBH', b, btree + 4)
        if ntype != 0:
            raise H5Error('not a group B-tree')
        pos = btree + 8 + 2 * self.O
        children = []
        for n in range(used):
            pos += self.L               # key n
            child, = struct.unpack_from('<Q', b, pos)
            pos += self.O
            children.append(child)
        for child in children:
            if level > 0:
                out.update(self._group_entries(child, heap))
                continue
            if b[child:child + 4] != b'SNOD':
                raise H5Error('bad symbol table node')
            nsym, = struct.unpack_from('<H', b, child + 6)
            for s in range(nsym):
                e = self._symtab_entry(child + 8 + s * (2 * self.O + 24))
                out[self._heap_name(heap, e['name_off'])] = e['header']
        return out

    def _global_heap_object(self, ref):
        """bytes of a variable-length value: reference = length (4), collection address (8), object index (4)"""
        b = self.b
        length, addr, index = struct.unpack('<IQI', ref)
        if b[addr:addr + 4] != b'GCOL':
            raise H5Error('bad global heap collection')
        size, = struct.unpack_from('<Q', b, addr + 8)
        pos, end = addr + 16, addr + size
        while pos + 16 <= end:
            idx, _refs, _, osize = struct.unpack_from('<HHIQ', b, pos)
            if idx == index:
                return b[pos + 16:pos + 16 + length]
            if idx == 0:
                break
            pos += 16 + (osize + 7) // 8 * 8
        raise H5Error('global heap object {} not found'.format(index))

    def _children(self, header):
        for mtype, _, payload in self._messages(header):
            if mtype == 0x0011:          # symbol table message
                btree, heap = struct.unpack_from('<QQ', payload, 0)
                return self._group_entries(btree, heap)
        return None

    def _resolve(self, path):
        header = self.root
        for part in [p for p in path.split('/') if p]:
            kids = self._children(header)
            if kids is None or part not in kids:
                raise KeyError(path)
            header = kids[part]
        return header

    @staticmethod
    def _dtype(payload):
        """(numpy dtype or 'S<n>' string dtype, bytes consumed are not needed)"""
        cls = payload[0] & 0x0f
        bits0 = payload[1]
        size, = struct.unpack_from('<I', payload, 4)
        order = '>' if bits0 & 1 else '<'
        if cls == 1:
            return np.dtype(order + 'f' + str(size))
        if cls == 0:
            return np.dtype(order + ('i' if bits0 & 8 else 'u') + str(size))
        if cls == 3:
            return np.dtype('S' + str(size))
        if cls == 9:
            return 'vlen'       # variable-length (h5py stores str attributes this way): 16-byte global-heap references
        raise H5Error('datatype class {} not supported'.format(cls))

    def _dims(self, payload):
        version, rank, flags = payload[0], payload[1], payload[2]
        if version == 1:
            off = 8
        elif version == 2:
            off = 4
        else:
            raise H5Error('dataspace version {} not supported'.format(version))
        return tuple(struct.unpack_from('<Q', payload, off + self.L * n)[0] for n in range(rank))

    # ------------------------------------------------------------------ public
    def keys(self, path='/'):
        kids = self._children(self._resolve(path))
        if kids is None:
            raise KeyError('{} is not a group'.format(path))
        return sorted(kids)

    def attrs(self, path='/'):
        out = {}
        for mtype, _, p in self._messages(self._resolve(path)):
            if mtype != 0x000C:
                continue
            version = p[0]
            if version != 1:
                raise H5Error('attribute message version {} not supported'.format(version))
            nname, ntype, nspace = struct.unpack_from('<HHH', p, 2)
            pad = lambda n: (n + 7) // 8 * 8
            pos = 8
            name = p[pos:pos + nname].split(b'\0')[0].decode('utf-8')
            pos += pad(nname)
            dt = self._dtype(p[pos:pos + ntype])
            pos += pad(ntype)
            dims = self._dims(p[pos:pos + nspace])
            pos += pad(nspace)
            count = int(np.prod(dims)) if dims else 1
            if isinstance(dt, str):
                vals = [self._global_heap_object(p[pos + 16 * n:pos + 16 * n + 16]).decode('utf-8', 'replace') for n in range(count)]
                out[name] = vals[0] if not dims else vals
                continue
            val = np.frombuffer(p, dtype=dt, count=count, offset=pos)
            if dt.kind == 'S':
                val = [v.split(b'\0')[0].decode('utf-8', 'replace') for v in val]
            out[name] = (val[0] if not dims else (list(val) if dt.kind == 'S' else val.reshape(dims).copy()))
        return out

    def dataset(self, path):
        dt = dims = addr = nbytes = None
        compact = None
        for mtype, _, p in self._messages(self._resolve(path)):
            if mtype == 0x0001:
                dims = self._dims(p)
            elif mtype == 0x0003:
                dt = self._dtype(p)
            elif mtype == 0x0008:
                version = p[0]
                if version == 3:
                    cls = p[1]
                    if cls == 1:
                        addr, nbytes = struct.unpack_from('<QQ', p, 2)
                    elif cls == 0:
                        n, = struct.unpack_from('<H', p, 2)
                        compact = p[4:4 + n]
                    else:
                        raise H5Error('chunked datasets are not supported')
                elif version in (1, 2):
                    rank, cls = p[1], p[2]
                    if cls != 1:
                        raise H5Error('only contiguous version-1/2 layouts are supported')
                    addr, = struct.unpack_from('<Q', p, 8)
                else:
                    raise H5Error('layout version {} not supported'.format(version))
        if dt is None or dims is None or (addr is None and compact is None):
            raise H5Error('{} is not a readable dataset'.format(path))
        count = int(np.prod(dims)) if dims else 1
        if compact is not None:
            return np.frombuffer(compact, dtype=dt, count=count).reshape(dims).copy()
        if addr == 0xffffffffffffffff:
            return np.zeros(dims, dtype=dt)
        return np.frombuffer(self.b, dtype=dt, count=count, offset=addr).reshape(dims).copy()
