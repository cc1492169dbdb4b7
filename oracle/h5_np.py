"""Pure-Python/numpy reader for the HDF5 subset that SkyH5 files use -- TEST INFRASTRUCTURE ONLY.

The reference reads SkyH5 through h5py (src/pyradiosky/skymodel.py:3491-3766, helper
``_get_value_hdf5_group`` :270-302); h5py / libhdf5 are absent from this image, so this module restates
the published HDF5 file-format specification (version 0/1 superblock, version 1 object headers,
symbol-table groups, B-tree v1, local and global heaps, contiguous / compact / chunked layouts with the
deflate and shuffle filters) in the simplest possible way: one whole-file ``bytes`` object and
``struct``.  It exists to check the product reader (pyradiosky_b200/csrc/rsk_h5.cpp, written
independently in C++) value for value on the reference's own fixture files, and is pinned itself by the
known answers the reference's tutorial prints for ``gsm_icrs.skyh5`` (docs/tutorial.rst:1279-1297).
Nothing under pyradiosky_b200/ may import it.
"""

import struct
import zlib

import numpy as np

UNDEF = 0xFFFFFFFFFFFFFFFF


class H5Error(ValueError):
    pass


class Dataset:
    def __init__(self, f, addr, msgs):
        self.f = f
        self.addr = addr
        self.msgs = msgs
        self.shape = None
        self.dtype = None  # dict
        self.layout = None
        self.filters = []
        self.attrs_raw = []
        for t, body in msgs:
            if t == 0x0001:
                self.shape = f._dataspace(body)
            elif t == 0x0003:
                self.dtype = f._datatype(body)[0]
            elif t == 0x0008:
                self.layout = f._layout(body)
            elif t == 0x000B:
                self.filters = f._filters(body)
            elif t == 0x000C:
                self.attrs_raw.append(body)

    @property
    def attrs(self):
        out = {}
        for body in self.attrs_raw:
            k, v = self.f._attribute(body)
            out[k] = v
        return out

    def read(self):
        return self.f._read_dataset(self)


class Group:
    def __init__(self, f, addr, msgs):
        self.f = f
        self.addr = addr
        self.msgs = msgs
        self.links = {}
        self.attrs_raw = [b for t, b in msgs if t == 0x000C]
        for t, body in msgs:
            if t == 0x0011:
                bt, hp = struct.unpack_from("<QQ", body, 0)
                self.links = f._group_links(bt, hp)

    @property
    def attrs(self):
        return dict(self.f._attribute(b) for b in self.attrs_raw)

    def keys(self):
        return sorted(self.links)

    def __contains__(self, k):
        return k in self.links

    def __getitem__(self, k):
        return self.f._object(self.links[k])


class File:
    def __init__(self, path):
        with open(path, "rb") as fh:
            self.b = fh.read()
        b = self.b
        if b[:8] != b"\x89HDF\r\n\x1a\n":
            raise H5Error("not an HDF5 file")
        ver = b[8]
        if ver not in (0, 1):
            raise H5Error(f"superblock version {ver} not supported")
        if b[13] != 8 or b[14] != 8:
            raise H5Error("only 8-byte offsets/lengths supported")
        pos = 24 + (4 if ver == 1 else 0)
        self.base, _fs, self.eof, _drv = struct.unpack_from("<QQQQ", b, pos)
        pos += 32
        _name_off, self.root_addr = struct.unpack_from("<QQ", b, pos)
        self.root = self._object(self.root_addr)

    # ---- object headers -------------------------------------------------------------------------
    def _messages(self, addr):
        b = self.b
        ver = b[addr]
        if ver != 1:
            raise H5Error(f"object header version {ver} at {addr} not supported")
        nmsg = struct.unpack_from("<H", b, addr + 2)[0]
        hsize = struct.unpack_from("<I", b, addr + 8)[0]
        blocks = [(addr + 16, hsize)]
        msgs = []
        while blocks and len(msgs) < nmsg:
            p, size = blocks.pop(0)
            end = p + size
            while p + 8 <= end and len(msgs) < nmsg:
                t, sz, _flags = struct.unpack_from("<HHB", b, p)
                body = b[p + 8:p + 8 + sz]
                p += 8 + sz
                if t == 0x0010:
                    off, ln = struct.unpack_from("<QQ", body, 0)
                    blocks.append((off, ln))
                msgs.append((t, body))
        return msgs

    def _object(self, addr):
        msgs = self._messages(addr)
        types = {t for t, _ in msgs}
        if 0x0011 in types:
            return Group(self, addr, msgs)
        if 0x0008 in types:
            return Dataset(self, addr, msgs)
        return Group(self, addr, msgs)

    def __getitem__(self, path):
        obj = self.root
        for part in [p for p in path.split("/") if p]:
            obj = obj[part]
        return obj

    def __contains__(self, path):
        obj = self.root
        for part in [p for p in path.split("/") if p]:
            if not isinstance(obj, Group) or part not in obj:
                return False
            obj = obj[part]
        return True

    # ---- groups ---------------------------------------------------------------------------------
    def _heap_data(self, addr):
        b = self.b
        if b[addr:addr + 4] != b"HEAP":
            raise H5Error("bad local heap")
        size, _free, daddr = struct.unpack_from("<QQQ", b, addr + 8)
        return b[daddr:daddr + size]

    def _group_links(self, btree, heap):
        heapd = self._heap_data(heap)
        links = {}

        def walk(addr):
            b = self.b
            if b[addr:addr + 4] == b"TREE":
                ntype, level, n = struct.unpack_from("<BBH", b, addr + 4)
                if ntype != 0:
                    raise H5Error("expected a group B-tree")
                p = addr + 8 + 16
                for i in range(n):
                    child = struct.unpack_from("<Q", b, p + 8)[0]
                    walk(child)
                    p += 16
            elif b[addr:addr + 4] == b"SNOD":
                n = struct.unpack_from("<H", b, addr + 6)[0]
                p = addr + 8
                for i in range(n):
                    noff, oaddr = struct.unpack_from("<QQ", b, p)
                    name = heapd[noff:heapd.index(b"\0", noff)].decode("utf8")
                    links[name] = oaddr
                    p += 40
            else:
                raise H5Error("bad group node")

        walk(btree)
        return links

    # ---- messages -------------------------------------------------------------------------------
    def _dataspace(self, body):
        ver, rank, flags = body[0], body[1], body[2]
        if ver == 1:
            p = 8
        elif ver == 2:
            p = 4
            if body[3] == 2:
                return None  # null dataspace
        else:
            raise H5Error("dataspace version")
        return tuple(struct.unpack_from("<Q", body, p + 8 * i)[0] for i in range(rank))

    def _datatype(self, body, pos=0):
        """Returns (dict, bytes consumed)."""
        cv = body[pos]
        cls, ver = cv & 0x0F, cv >> 4
        bits = body[pos + 1] | (body[pos + 2] << 8) | (body[pos + 3] << 16)
        size = struct.unpack_from("<I", body, pos + 4)[0]
        d = {"class": cls, "size": size, "bits": bits, "version": ver}
        used = 8
        if cls == 0:
            if bits & 1:
                raise H5Error("big-endian integers not supported")
            d["np"] = np.dtype(("<i" if bits & 8 else "<u") + str(size))
            used += 4
        elif cls == 1:
            if bits & 1:
                raise H5Error("big-endian floats not supported")
            d["np"] = np.dtype("<f" + str(size))
            used += 12
        elif cls == 3:
            d["np"] = np.dtype("S" + str(size))
        elif cls == 9:
            d["vlen_string"] = (bits & 0x0F) == 1
            base, n = self._datatype(body, pos + 8)
            d["base"] = base
            used += n
        elif cls == 8:
            base, n = self._datatype(body, pos + 8)
            d["base"] = base
            d["np"] = base["np"]
            nmemb = bits & 0xFFFF
            p = pos + 8 + n
            for _ in range(nmemb):
                e = body.index(b"\0", p)
                ln = e - p + 1
                if ver < 3:
                    ln = (ln + 7) & ~7
                p += ln
            p += nmemb * base["size"]
            used = p - pos
        else:
            raise H5Error(f"datatype class {cls} not supported")
        return d, used

    def _layout(self, body):
        ver = body[0]
        if ver == 3:
            cls = body[1]
            if cls == 0:
                size = struct.unpack_from("<H", body, 2)[0]
                return {"class": 0, "data": body[4:4 + size]}
            if cls == 1:
                addr, size = struct.unpack_from("<QQ", body, 2)
                return {"class": 1, "addr": addr, "size": size}
            if cls == 2:
                nd = body[2]
                addr = struct.unpack_from("<Q", body, 3)[0]
                dims = struct.unpack_from("<" + "I" * nd, body, 11)
                return {"class": 2, "addr": addr, "chunk": dims}
            raise H5Error("layout class")
        if ver in (1, 2):
            nd, cls = body[1], body[2]
            p = 8
            addr = None
            if cls != 0:
                addr = struct.unpack_from("<Q", body, p)[0]
                p += 8
            dims = struct.unpack_from("<" + "I" * nd, body, p)
            p += 4 * nd
            if cls == 0:
                size = struct.unpack_from("<I", body, p)[0]
                return {"class": 0, "data": body[p + 4:p + 4 + size]}
            if cls == 1:
                return {"class": 1, "addr": addr, "size": None}
            return {"class": 2, "addr": addr, "chunk": dims}
        raise H5Error(f"layout version {ver} not supported")

    def _filters(self, body):
        ver, n = body[0], body[1]
        out = []
        p = 8 if ver == 1 else 2
        for _ in range(n):
            fid = struct.unpack_from("<H", body, p)[0]
            p += 2
            if ver == 1 or fid >= 256:
                nlen = struct.unpack_from("<H", body, p)[0]
                p += 2
            else:
                nlen = 0
            _flags, ncd = struct.unpack_from("<HH", body, p)
            p += 4
            if ver == 1:
                nlen = (nlen + 7) & ~7
            p += nlen
            cd = struct.unpack_from("<" + "I" * ncd, body, p)
            p += 4 * ncd
            if ver == 1 and ncd % 2:
                p += 4
            out.append((fid, cd))
        return out

    def _attribute(self, body):
        ver = body[0]
        nsz, tsz, ssz = struct.unpack_from("<HHH", body, 2)
        p = 8
        if ver == 3:
            p = 9
        pad = (lambda x: (x + 7) & ~7) if ver == 1 else (lambda x: x)
        name = body[p:p + nsz].split(b"\0")[0].decode("utf8")
        p += pad(nsz)
        dt, _ = self._datatype(body, p)
        p += pad(tsz)
        shape = self._dataspace(body[p:p + ssz])
        p += pad(ssz)
        n = int(np.prod(shape)) if shape else 1
        val = self._decode(dt, body[p:], n, shape or ())
        return name, val

    # ---- data -----------------------------------------------------------------------------------
    def _gheap_object(self, addr, index):
        b = self.b
        if b[addr:addr + 4] != b"GCOL":
            raise H5Error("bad global heap")
        csize = struct.unpack_from("<Q", b, addr + 8)[0]
        p = addr + 16
        end = addr + csize
        while p + 16 <= end:
            idx, _rc, _r, osz = struct.unpack_from("<HHIQ", b, p)
            if idx == 0:
                break
            if idx == index:
                return b[p + 16:p + 16 + osz]
            p += 16 + ((osz + 7) & ~7)
        raise H5Error("global heap object not found")

    def _decode(self, dt, raw, n, shape):
        if dt["class"] == 9:
            if not dt.get("vlen_string"):
                raise H5Error("only variable-length strings supported")
            out = []
            for i in range(n):
                ln, addr, idx = struct.unpack_from("<IQI", raw, 16 * i)
                out.append(self._gheap_object(addr, idx)[:ln] if addr not in (0, UNDEF) else b"")
            arr = np.array(out, dtype=object).reshape(shape)
            return arr if shape else arr[()]
        npdt = dt["np"]
        arr = np.frombuffer(raw[:n * npdt.itemsize], dtype=npdt).reshape(shape)
        return arr.copy() if shape else arr.reshape(())[()]

    def _read_dataset(self, ds):
        shape = ds.shape
        if shape is None:
            return None
        n = int(np.prod(shape)) if shape else 1
        dt = ds.dtype
        esize = dt["size"]
        lay = ds.layout
        if lay["class"] == 0:
            raw = lay["data"]
        elif lay["class"] == 1:
            if lay["addr"] == UNDEF:
                raw = b"\0" * (n * esize)
            else:
                raw = self.b[lay["addr"]:lay["addr"] + n * esize]
        else:
            raw = self._read_chunked(ds, n, esize)
        return self._decode(dt, raw, n, shape)

    def _read_chunked(self, ds, n, esize):
        shape = ds.shape
        chunk = ds.layout["chunk"][:-1]
        rank = len(shape)
        out = np.zeros(shape, dtype=np.dtype("V" + str(esize)))
        csize_bytes = int(np.prod(chunk)) * esize

        def walk(addr):
            b = self.b
            if addr == UNDEF:
                return
            if b[addr:addr + 4] != b"TREE":
                raise H5Error("bad chunk B-tree")
            ntype, level, nent = struct.unpack_from("<BBH", b, addr + 4)
            if ntype != 1:
                raise H5Error("expected a chunk B-tree")
            keysz = 8 + 8 * (rank + 1)
            p = addr + 24
            for _ in range(nent):
                csz, fmask = struct.unpack_from("<II", b, p)
                offs = struct.unpack_from("<" + "Q" * (rank + 1), b, p + 8)
                child = struct.unpack_from("<Q", b, p + keysz)[0]
                if level > 0:
                    walk(child)
                else:
                    raw = b[child:child + csz]
                    for k, (fid, cd) in reversed(list(enumerate(ds.filters))):
                        if fmask & (1 << k):
                            continue
                        if fid == 1:
                            raw = zlib.decompress(raw)
                        elif fid == 2:
                            es = cd[0] if cd else esize
                            ne = len(raw) // es
                            a = np.frombuffer(raw[:ne * es], dtype=np.uint8).reshape(es, ne)
                            raw = a.T.tobytes() + raw[ne * es:]
                        else:
                            raise H5Error(f"filter {fid} not supported")
                    if len(raw) < csize_bytes:
                        raise H5Error("short chunk")
                    carr = np.frombuffer(raw[:csize_bytes], dtype=out.dtype).reshape(chunk)
                    sl_out = tuple(slice(o, min(o + c, s)) for o, c, s in zip(offs[:-1], chunk, shape))
                    sl_in = tuple(slice(0, s.stop - s.start) for s in sl_out)
                    out[sl_out] = carr[sl_in]
                p += keysz + 8

        walk(ds.layout["addr"])
        return out.tobytes()


def tree(path):
    """{hdf5 path: (kind, shape, dtype description, attrs)} for every object in the file."""
    f = File(path)
    out = {}

    def rec(g, prefix):
        for k in g.keys():
            o = g[k]
            p = prefix + "/" + k
            if isinstance(o, Group):
                out[p] = ("group", None, None, o.attrs)
                rec(o, p)
            else:
                out[p] = ("dataset", o.shape, o.dtype, o.attrs)

    rec(f.root, "")
    return out
