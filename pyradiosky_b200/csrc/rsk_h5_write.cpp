// rsk_h5_write.cpp -- SkyH5 (HDF5) writer: replaces the h5py calls of SkyModel.write_skyh5
// (src/pyradiosky/skymodel.py:4871-5029, helper _add_value_hdf5_group :229-267).
//
// The file produced is the same flavour h5py writes with its default libver='earliest', so the reference
// (and h5py/libhdf5 in general) can read it: version 0 superblock, version 1 object headers, symbol-table
// groups (one B-tree node over up to 32 symbol nodes of 8 entries, libhdf5's default K values), datasets
// stored contiguous or chunked (B-tree v1 index, optional shuffle + deflate), string attributes as
// variable-length UTF-8 strings in a global heap collection (h5py hands those back as ``str``, which
// _get_value_hdf5_group compares with "latitude"/"longitude", :279-283).
//
// Raw data is written as soon as rsk_h5w_dataset() is called (nothing big is kept in memory); all
// metadata is laid out and written by rsk_h5w_close().
#include <cuda_runtime.h>
#include <fcntl.h>
#include <stdint.h>
#include <string.h>
#include <time.h>
#include <unistd.h>
#include <zlib.h>

#include <algorithm>
#include <atomic>
#include <map>
#include <memory>
#include <string>
#include <thread>
#include <vector>

#include "../../include/rsk_io.h"

namespace {

constexpr uint64_t kUndef = ~0ull;
constexpr int kLeafK = 4;        // group leaf node K: a symbol node holds 2K entries
constexpr int kInternalK = 16;   // group B-tree node K: a node holds 2K children
constexpr int kChunkK = 32;      // indexed-storage B-tree K (libhdf5 default; not stored in a v0 superblock)

struct Bytes {
  std::vector<uint8_t> v;
  void u8(uint8_t x) { v.push_back(x); }
  void u16(uint16_t x) { put(&x, 2); }
  void u32(uint32_t x) { put(&x, 4); }
  void u64(uint64_t x) { put(&x, 8); }
  void put(const void* p, size_t n) {
    const uint8_t* b = static_cast<const uint8_t*>(p);
    v.insert(v.end(), b, b + n);
  }
  void zeros(size_t n) { v.insert(v.end(), n, 0); }
  void pad8() { zeros((8 - v.size() % 8) % 8); }
  size_t size() const { return v.size(); }
};

struct ChunkOut {
  uint64_t addr;
  uint32_t size;
  uint64_t off[8];
};

struct Node {
  bool is_group = true;
  std::map<std::string, std::unique_ptr<Node>> kids;
  std::vector<std::pair<std::string, std::string>> attrs;
  // dataset
  int type_class = 0, elem_size = 0, rank = 0;
  int64_t dims[8] = {0};
  bool chunked = false;
  int64_t chunk[8] = {0};
  int gzip = 0;
  bool utf8 = false;
  uint64_t data_addr = kUndef, data_size = 0;
  std::vector<ChunkOut> chunks;
  // layout results
  uint64_t header_addr = kUndef;
  uint64_t btree_addr = kUndef, heap_addr = kUndef;
  std::vector<uint8_t> attr_refs;  // 16-byte global-heap references, one per attribute
};

}  // namespace

struct RskH5Writer {
  int fd = -1;
  uint64_t eof = 96;  // bump allocator; the superblock occupies [0, 96)
  Node root;
  bool failed = false;

  ~RskH5Writer() {
    if (fd >= 0) close(fd);
  }

  uint64_t alloc(uint64_t n) {
    eof = (eof + 7) & ~7ull;
    const uint64_t a = eof;
    eof += n;
    return a;
  }

  bool write_at(uint64_t addr, const void* p, size_t n) {
    const uint8_t* b = static_cast<const uint8_t*>(p);
    while (n > 0) {
      const ssize_t w = pwrite(fd, b, n, (off_t)addr);
      if (w <= 0) return false;
      b += w;
      addr += (uint64_t)w;
      n -= (size_t)w;
    }
    return true;
  }

  // parent node of path and the leaf name; nullptr if a parent is missing or not a group
  Node* parent_of(const std::string& path, std::string* leaf) {
    Node* cur = &root;
    std::vector<std::string> parts;
    size_t i = 0;
    while (i < path.size()) {
      while (i < path.size() && path[i] == '/') ++i;
      size_t j = i;
      while (j < path.size() && path[j] != '/') ++j;
      if (j > i) parts.push_back(path.substr(i, j - i));
      i = j;
    }
    if (parts.empty()) return nullptr;
    for (size_t k = 0; k + 1 < parts.size(); ++k) {
      auto it = cur->kids.find(parts[k]);
      if (it == cur->kids.end() || !it->second->is_group) return nullptr;
      cur = it->second.get();
    }
    *leaf = parts.back();
    return cur;
  }

  Node* find(const std::string& path) {
    std::string leaf;
    bool only_slashes = true;
    for (char c : path) only_slashes = only_slashes && c == '/';
    if (only_slashes) return &root;
    Node* p = parent_of(path, &leaf);
    if (!p) return nullptr;
    auto it = p->kids.find(leaf);
    return it == p->kids.end() ? nullptr : it->second.get();
  }
};

namespace {

void datatype_msg(Bytes& m, int type_class, int size, bool utf8) {
  if (type_class == RSK_H5_INT || type_class == RSK_H5_UINT) {
    m.u8(0x10);
    m.u8(type_class == RSK_H5_INT ? 0x08 : 0x00);
    m.u8(0);
    m.u8(0);
    m.u32((uint32_t)size);
    m.u16(0);
    m.u16((uint16_t)(8 * size));
  } else if (type_class == RSK_H5_FLOAT) {
    m.u8(0x11);
    m.u8(0x20);
    m.u8(size == 8 ? 0x3f : 0x1f);
    m.u8(0);
    m.u32((uint32_t)size);
    m.u16(0);
    m.u16((uint16_t)(8 * size));
    if (size == 8) {
      m.u8(52); m.u8(11); m.u8(0); m.u8(52); m.u32(1023);
    } else {
      m.u8(23); m.u8(8); m.u8(0); m.u8(23); m.u32(127);
    }
  } else {  // fixed-length string, null padded
    m.u8(0x13);
    m.u8(utf8 ? 0x11 : 0x01);
    m.u8(0);
    m.u8(0);
    m.u32((uint32_t)size);
  }
}

void add_msg(Bytes& hdr, uint16_t type, uint8_t flags, const Bytes& body, int* nmsg) {
  Bytes b = body;
  b.pad8();
  hdr.u16(type);
  hdr.u16((uint16_t)b.size());
  hdr.u8(flags);
  hdr.zeros(3);
  hdr.put(b.v.data(), b.size());
  ++*nmsg;
}

Bytes finish_header(const Bytes& msgs, int nmsg) {
  Bytes h;
  h.u8(1);
  h.u8(0);
  h.u16((uint16_t)nmsg);
  h.u32(1);
  h.u32((uint32_t)msgs.size());
  h.u32(0);
  h.put(msgs.v.data(), msgs.size());
  return h;
}

void attr_msgs(const Node& n, Bytes& msgs, int* nmsg) {
  for (size_t k = 0; k < n.attrs.size(); ++k) {
    Bytes a;
    const std::string& name = n.attrs[k].first;
    a.u8(1);
    a.u8(0);
    a.u16((uint16_t)(name.size() + 1));
    a.u16(20);  // datatype size
    a.u16(8);   // dataspace size
    a.put(name.c_str(), name.size() + 1);
    a.pad8();
    // variable-length UTF-8 string, null terminated
    a.u8(0x19); a.u8(0x01); a.u8(0x01); a.u8(0x00); a.u32(16);
    a.u8(0x10); a.u8(0); a.u8(0); a.u8(0); a.u32(1); a.u16(0); a.u16(8);
    a.pad8();
    a.u8(1); a.zeros(7);  // scalar dataspace, version 1
    a.put(n.attr_refs.data() + 16 * k, 16);
    add_msg(msgs, 0x000c, 0, a, nmsg);
  }
}

}  // namespace

extern "C" {

int rsk_h5w_create(const char* path, RskH5Writer** out) {
  if (!path || !out) return RSK_H5_ERR_NULL;
  *out = nullptr;
  const int fd = open(path, O_CREAT | O_TRUNC | O_RDWR, 0644);
  if (fd < 0) return RSK_H5_ERR_OPEN;
  RskH5Writer* w = new RskH5Writer();
  w->fd = fd;
  *out = w;
  return RSK_H5_OK;
}

int rsk_h5w_group(RskH5Writer* w, const char* path) {
  if (!w || !path) return RSK_H5_ERR_NULL;
  std::string leaf;
  Node* p = w->parent_of(path, &leaf);
  if (!p || p->kids.count(leaf)) return RSK_H5_ERR_STATE;
  if (p->kids.size() >= (size_t)(2 * kLeafK) * (2 * kInternalK)) return RSK_H5_ERR_UNSUPPORTED;
  p->kids[leaf] = std::make_unique<Node>();
  return RSK_H5_OK;
}

}  // extern "C"

namespace {

// A piece of the dataset that is in host memory: the sub-array [lo, lo + ext) in C order at ``base``.
struct Slab {
  const uint8_t* base;
  int64_t lo[8];
  int64_t ext[8];
  int64_t c_first, c_count;  // the chunks (C order of the chunk grid) that lie inside it
};

// Where the data comes from.  Host arrays are one slab; device arrays are staged through two pinned buffers:
// a slab is one chunk-row along the first axis times a run of chunk-rows along the second (all of the last
// axis), fetched with one pitched device->host copy while the previous slab is being chunked and written.
struct Source {
  const uint8_t* host = nullptr;
  const uint8_t* dev = nullptr;
  cudaStream_t stream = nullptr;
  uint8_t* stage[2] = {nullptr, nullptr};
  cudaEvent_t ev[2] = {nullptr, nullptr};
  size_t stage_bytes = 0;
  ~Source() {
    for (int i = 0; i < 2; ++i) {
      if (ev[i]) cudaEventDestroy(ev[i]);
      if (stage[i]) cudaFreeHost(stage[i]);
    }
  }
  int ensure(size_t bytes) {
    if (bytes <= stage_bytes) return RSK_H5_OK;
    for (int i = 0; i < 2; ++i) {
      if (stage[i]) cudaFreeHost(stage[i]);
      stage[i] = nullptr;
      cudaError_t e = cudaHostAlloc(reinterpret_cast<void**>(&stage[i]), bytes, cudaHostAllocDefault);
      if (e == cudaSuccess && !ev[i]) e = cudaEventCreateWithFlags(&ev[i], cudaEventDisableTiming);
      if (e != cudaSuccess) return (int)e;
    }
    stage_bytes = bytes;
    return RSK_H5_OK;
  }
};

int write_dataset_impl(RskH5Writer* w, const char* path, int type_class, int elem_size, int rank, const int64_t* dims,
                       Source& source, const int64_t* chunk, int gzip_level) {
  if (!w || !path) return RSK_H5_ERR_NULL;
  if (rank < 0 || rank > 7 || (rank > 0 && !dims)) return RSK_H5_ERR_SIZE;
  if (type_class < RSK_H5_INT || type_class > RSK_H5_STRING) return RSK_H5_ERR_TYPE;
  if (type_class == RSK_H5_FLOAT ? (elem_size != 4 && elem_size != 8)
                                 : type_class == RSK_H5_STRING ? elem_size < 1
                                                               : (elem_size != 1 && elem_size != 2 && elem_size != 4 && elem_size != 8))
    return RSK_H5_ERR_TYPE;
  std::string leaf;
  Node* p = w->parent_of(path, &leaf);
  if (!p || p->kids.count(leaf)) return RSK_H5_ERR_STATE;
  if (p->kids.size() >= (size_t)(2 * kLeafK) * (2 * kInternalK)) return RSK_H5_ERR_UNSUPPORTED;
  auto node = std::make_unique<Node>();
  Node& d = *node;
  d.is_group = false;
  d.type_class = type_class;
  d.elem_size = elem_size;
  d.rank = rank;
  int64_t count = 1;
  for (int a = 0; a < rank; ++a) {
    if (dims[a] < 0) return RSK_H5_ERR_SIZE;
    d.dims[a] = dims[a];
    count *= dims[a];
  }
  if (count > 0 && !source.host && !source.dev) return RSK_H5_ERR_NULL;
  if (type_class == RSK_H5_STRING) {
    if (!source.host && count > 0) return RSK_H5_ERR_TYPE;  // strings live on the host
    for (int64_t i = 0; i < count * elem_size; ++i) d.utf8 = d.utf8 || source.host[i] >= 0x80;
  }
  d.chunked = chunk != nullptr && rank > 0 && count > 0;
  const int64_t slab_target = 64ll << 20;
  if (!d.chunked) {
    d.data_size = (uint64_t)count * elem_size;
    if (d.data_size > 0) {
      d.data_addr = w->alloc(d.data_size);
      if (source.host) {
        if (!w->write_at(d.data_addr, source.host, d.data_size)) return RSK_H5_ERR_OPEN;
      } else {  // device array, contiguous layout: flat slabs through the two staging buffers
        const uint64_t per = (uint64_t)std::min<int64_t>(slab_target, (int64_t)d.data_size);
        int rc = source.ensure(per);
        if (rc != RSK_H5_OK) return rc;
        const uint64_t nslabs = (d.data_size + per - 1) / per;
        auto fetch = [&](uint64_t sidx) {
          const uint64_t o = sidx * per, nbytes = std::min(per, d.data_size - o);
          cudaMemcpyAsync(source.stage[sidx & 1], source.dev + o, nbytes, cudaMemcpyDeviceToHost, source.stream);
          cudaEventRecord(source.ev[sidx & 1], source.stream);
        };
        fetch(0);
        for (uint64_t sidx = 0; sidx < nslabs; ++sidx) {
          if (cudaEventSynchronize(source.ev[sidx & 1]) != cudaSuccess) return (int)cudaGetLastError();
          if (sidx + 1 < nslabs) fetch(sidx + 1);
          const uint64_t o = sidx * per, nbytes = std::min(per, d.data_size - o);
          if (!w->write_at(d.data_addr + o, source.stage[sidx & 1], nbytes)) return RSK_H5_ERR_OPEN;
        }
      }
    }
    p->kids[leaf] = std::move(node);
    return RSK_H5_OK;
  }

  d.gzip = gzip_level < 0 ? 0 : (gzip_level > 9 ? 9 : gzip_level);
  int64_t grid[8], nchunks = 1, cbytes = elem_size;
  for (int a = 0; a < rank; ++a) {
    if (chunk[a] < 1 || chunk[a] > 0xffffffffll) return RSK_H5_ERR_SIZE;
    d.chunk[a] = std::min(chunk[a], std::max<int64_t>(d.dims[a], 1));
    grid[a] = (d.dims[a] + d.chunk[a] - 1) / d.chunk[a];
    nchunks *= grid[a];
    cbytes *= d.chunk[a];
  }
  if (cbytes > 0xffffffffll) return RSK_H5_ERR_SIZE;

  // ---- slabs ----------------------------------------------------------------------------------------
  std::vector<Slab> slabs;
  if (source.host) {
    Slab sl = {};
    sl.base = source.host;
    for (int a = 0; a < rank; ++a) {
      sl.lo[a] = 0;
      sl.ext[a] = d.dims[a];
    }
    sl.c_first = 0;
    sl.c_count = nchunks;
    slabs.push_back(sl);
  } else {
    if (rank > 3) return RSK_H5_ERR_UNSUPPORTED;  // device arrays: up to three dimensions
    // (A, B, C) view with leading ones
    int64_t D[3] = {1, 1, 1}, Cc[3] = {1, 1, 1}, G[3] = {1, 1, 1};
    for (int a = 0; a < rank; ++a) {
      D[3 - rank + a] = d.dims[a];
      Cc[3 - rank + a] = d.chunk[a];
      G[3 - rank + a] = grid[a];
    }
    const int64_t row_bytes = Cc[0] * Cc[1] * D[2] * elem_size;
    const int64_t rows_per_slab = std::max<int64_t>(1, std::min(G[1], slab_target / std::max<int64_t>(row_bytes, 1)));
    int rc = source.ensure((size_t)(rows_per_slab * row_bytes));
    if (rc != RSK_H5_OK) return rc;
    for (int64_t ia = 0; ia < G[0]; ++ia)
      for (int64_t ib0 = 0; ib0 < G[1]; ib0 += rows_per_slab) {
        const int64_t ib1 = std::min(G[1], ib0 + rows_per_slab);
        Slab sl = {};
        int64_t lo3[3] = {ia * Cc[0], ib0 * Cc[1], 0};
        int64_t ext3[3] = {std::min(Cc[0], D[0] - lo3[0]), std::min(ib1 * Cc[1], D[1]) - lo3[1], D[2]};
        for (int a = 0; a < rank; ++a) {
          sl.lo[a] = lo3[3 - rank + a];
          sl.ext[a] = ext3[3 - rank + a];
        }
        sl.c_first = (ia * G[1] + ib0) * G[2];
        sl.c_count = (ib1 - ib0) * G[2];
        slabs.push_back(sl);
      }
  }
  auto fetch = [&](size_t si) {  // device -> staging buffer si & 1 (asynchronous)
    Slab& sl = slabs[si];
    int64_t lo3[3] = {0, 0, 0}, ext3[3] = {1, 1, 1}, D[3] = {1, 1, 1};
    for (int a = 0; a < rank; ++a) {
      lo3[3 - rank + a] = sl.lo[a];
      ext3[3 - rank + a] = sl.ext[a];
      D[3 - rank + a] = d.dims[a];
    }
    const size_t width = (size_t)(ext3[1] * D[2] * elem_size);
    cudaMemcpy2DAsync(source.stage[si & 1], width, source.dev + ((lo3[0] * D[1] + lo3[1]) * D[2]) * elem_size,
                      (size_t)(D[1] * D[2] * elem_size), width, (size_t)ext3[0], cudaMemcpyDeviceToHost, source.stream);
    cudaEventRecord(source.ev[si & 1], source.stream);
    sl.base = source.stage[si & 1];
  };

  // ---- chunks ---------------------------------------------------------------------------------------
  // Chunks are produced by a few threads: gather (+ shuffle + deflate).  Uncompressed chunks have a known
  // size, so their addresses are assigned up front and every thread writes its own chunks; deflated ones
  // are produced in batches and written in order.
  unsigned hc = std::thread::hardware_concurrency();
  const int nt = (int)std::max(1u, std::min(hc ? hc : 4u, 16u));
  auto chunk_offsets = [&](int64_t ci, uint64_t* off) {
    int64_t rem = ci;
    for (int a = rank - 1; a >= 0; --a) {
      off[a] = (uint64_t)((rem % grid[a]) * d.chunk[a]);
      rem /= grid[a];
    }
  };
  auto gather = [&](int64_t ci, const Slab& sl, std::vector<uint8_t>& raw) {
    uint64_t off[8];
    chunk_offsets(ci, off);
    raw.assign((size_t)cbytes, 0);
    int64_t lo[8], ext[8];
    for (int a = 0; a < rank; ++a) {
      lo[a] = (int64_t)off[a];
      ext[a] = std::min(d.chunk[a], d.dims[a] - lo[a]);
    }
    int64_t idx[8] = {0};
    const int last = rank - 1;
    for (;;) {  // copy the rows of this chunk that lie inside the dataset
      int64_t so = 0, dof = 0;
      for (int a = 0; a < rank; ++a) {
        so = so * sl.ext[a] + (lo[a] + idx[a] - sl.lo[a]);
        dof = dof * d.chunk[a] + idx[a];
      }
      memcpy(raw.data() + dof * elem_size, sl.base + so * elem_size, (size_t)ext[last] * elem_size);
      int a = last - 1;
      for (; a >= 0; --a) {
        if (++idx[a] < ext[a]) break;
        idx[a] = 0;
      }
      if (a < 0) break;
    }
  };
  auto run_parallel = [&](int64_t n, auto&& body) {
    std::atomic<int64_t> next{0};
    auto worker = [&]() {
      std::vector<uint8_t> raw, tmp;
      for (;;) {
        const int64_t i = next.fetch_add(1);
        if (i >= n) break;
        body(i, raw, tmp);
      }
    };
    std::vector<std::thread> th;
    for (int t = 1; t < nt && t < n; ++t) th.emplace_back(worker);
    worker();
    for (auto& t : th) t.join();
  };
  if (!d.gzip) {
    d.chunks.resize((size_t)nchunks);
    for (int64_t ci = 0; ci < nchunks; ++ci) {
      d.chunks[(size_t)ci].size = (uint32_t)cbytes;
      d.chunks[(size_t)ci].addr = w->alloc((uint64_t)cbytes);
      chunk_offsets(ci, d.chunks[(size_t)ci].off);
    }
  }
  if (source.dev) fetch(0);
  for (size_t si = 0; si < slabs.size(); ++si) {
    if (source.dev) {
      if (cudaEventSynchronize(source.ev[si & 1]) != cudaSuccess) return (int)cudaGetLastError();
      if (si + 1 < slabs.size()) fetch(si + 1);  // the other buffer: its chunks were written one iteration ago
    }
    const Slab& sl = slabs[si];
    if (!d.gzip) {
      std::atomic<bool> io_ok{true};
      run_parallel(sl.c_count, [&](int64_t i, std::vector<uint8_t>& raw, std::vector<uint8_t>&) {
        const int64_t ci = sl.c_first + i;
        gather(ci, sl, raw);
        if (!w->write_at(d.chunks[(size_t)ci].addr, raw.data(), (size_t)cbytes)) io_ok = false;
      });
      if (!io_ok) return RSK_H5_ERR_OPEN;
    } else {
      const int64_t batch = (int64_t)nt * 4;
      std::vector<std::vector<uint8_t>> outb((size_t)batch);
      std::vector<int> status((size_t)batch);
      for (int64_t c0 = 0; c0 < sl.c_count; c0 += batch) {
        const int64_t nb = std::min(batch, sl.c_count - c0);
        run_parallel(nb, [&](int64_t i, std::vector<uint8_t>& raw, std::vector<uint8_t>& sh) {
          gather(sl.c_first + c0 + i, sl, raw);
          sh.resize((size_t)cbytes);
          const size_t ne = (size_t)(cbytes / elem_size);
          for (size_t e = 0; e < ne; ++e)
            for (int byte = 0; byte < elem_size; ++byte) sh[(size_t)byte * ne + e] = raw[e * elem_size + byte];
          std::vector<uint8_t>& ob = outb[(size_t)i];
          uLongf dl = compressBound((uLong)cbytes);
          ob.resize(dl);
          status[(size_t)i] = compress2(ob.data(), &dl, sh.data(), (uLong)cbytes, d.gzip) == Z_OK ? RSK_H5_OK : RSK_H5_ERR_FILTER;
          ob.resize(dl);
        });
        for (int64_t i = 0; i < nb; ++i) {
          if (status[(size_t)i] != RSK_H5_OK) return status[(size_t)i];
          ChunkOut co;
          co.size = (uint32_t)outb[(size_t)i].size();
          co.addr = w->alloc(co.size);
          chunk_offsets(sl.c_first + c0 + i, co.off);
          if (!w->write_at(co.addr, outb[(size_t)i].data(), co.size)) return RSK_H5_ERR_OPEN;
          d.chunks.push_back(co);
        }
      }
    }
  }
  p->kids[leaf] = std::move(node);
  return RSK_H5_OK;
}

}  // namespace

extern "C" {

int rsk_h5w_dataset(RskH5Writer* w, const char* path, int type_class, int elem_size, int rank, const int64_t* dims,
                    const void* data, const int64_t* chunk, int gzip_level) {
  Source src;
  src.host = static_cast<const uint8_t*>(data);
  return write_dataset_impl(w, path, type_class, elem_size, rank, dims, src, chunk, gzip_level);
}

int rsk_h5w_dataset_device(RskH5Writer* w, const char* path, int type_class, int elem_size, int rank,
                           const int64_t* dims, const void* device_data, const int64_t* chunk, int gzip_level,
                           void* stream) {
  Source src;
  src.dev = static_cast<const uint8_t*>(device_data);
  src.stream = (cudaStream_t)stream;
  return write_dataset_impl(w, path, type_class, elem_size, rank, dims, src, chunk, gzip_level);
}

int rsk_h5w_attr_string(RskH5Writer* w, const char* path, const char* name, const char* value) {
  if (!w || !path || !name || !value) return RSK_H5_ERR_NULL;
  Node* n = w->find(path);
  if (!n) return RSK_H5_ERR_NOTFOUND;
  for (const auto& kv : n->attrs)
    if (kv.first == name) return RSK_H5_ERR_STATE;
  if (strlen(name) > 1000 || strlen(value) > 60000) return RSK_H5_ERR_SIZE;
  n->attrs.emplace_back(name, value);
  return RSK_H5_OK;
}

}  // extern "C"

namespace {

// ---- close: lay out and write the metadata ------------------------------------------------------------

struct Closer {
  RskH5Writer* w;
  bool ok = true;

  void put(uint64_t addr, const Bytes& b) { ok = ok && w->write_at(addr, b.v.data(), b.size()); }

  void collect(Node& n, std::vector<Node*>& all) {
    all.push_back(&n);
    for (auto& kv : n.kids) collect(*kv.second, all);
  }

  // every attribute string goes into global heap collections (<= 64 KiB each, >= 4 KiB)
  void write_global_heaps(std::vector<Node*>& all) {
    struct Pending { Node* n; const std::string* s; };
    std::vector<Pending> items;
    for (Node* n : all)
      for (auto& kv : n->attrs) items.push_back({n, &kv.second});
    size_t i = 0;
    while (i < items.size()) {
      Bytes col;
      col.put("GCOL", 4);
      col.u8(1);
      col.zeros(3);
      col.u64(0);  // size patched below
      std::vector<std::pair<Node*, std::pair<uint32_t, uint32_t>>> placed;  // node, (len, index)
      uint16_t index = 1;
      while (i < items.size() && index < 0xfff0) {
        const std::string& s = *items[i].s;
        const size_t need = 16 + ((s.size() + 7) & ~(size_t)7);
        if (col.size() + need + 16 > 65536 && index > 1) break;
        col.u16(index);
        col.u16(0);
        col.u32(0);
        col.u64(s.size());
        col.put(s.data(), s.size());
        col.pad8();
        placed.push_back({items[i].n, {(uint32_t)s.size(), index}});
        ++index;
        ++i;
      }
      size_t total = std::max<size_t>(4096, (col.size() + 16 + 7) & ~(size_t)7);
      const size_t free_size = total - col.size();  // includes the 16-byte header of object 0
      col.u16(0);
      col.u16(0);
      col.u32(0);
      col.u64(free_size);
      col.zeros(total - col.size());
      const uint64_t total64 = total;
      memcpy(col.v.data() + 8, &total64, 8);
      const uint64_t addr = w->alloc(total);
      put(addr, col);
      for (auto& pl : placed) {
        Bytes ref;
        ref.u32(pl.second.first);
        ref.u64(addr);
        ref.u32(pl.second.second);
        pl.first->attr_refs.insert(pl.first->attr_refs.end(), ref.v.begin(), ref.v.end());
      }
    }
  }

  // B-tree v1 over the chunks of a dataset; returns the root node address
  uint64_t write_chunk_btree(Node& d) {
    const int nd = d.rank + 1;
    const size_t keysz = 8 + 8 * (size_t)nd;
    const size_t node_bytes = 24 + (size_t)(2 * kChunkK) * (keysz + 8) + keysz;
    struct Entry {
      Bytes key;      // key of the first chunk below
      uint64_t child;
    };
    auto chunk_key = [&](uint32_t size, const uint64_t* off) {
      Bytes k;
      k.u32(size);
      k.u32(0);
      for (int a = 0; a < d.rank; ++a) k.u64(off[a]);
      k.u64(0);
      return k;
    };
    std::vector<Entry> level;
    for (const ChunkOut& c : d.chunks) level.push_back({chunk_key(c.size, c.off), c.addr});
    // right-most key: one chunk past the last one in every dimension (zero-width chunk)
    Bytes end_key;
    {
      const ChunkOut& c = d.chunks.back();
      end_key.u32(0);
      end_key.u32(0);
      for (int a = 0; a < d.rank; ++a) end_key.u64(c.off[a] + (uint64_t)d.chunk[a]);
      end_key.u64((uint64_t)d.elem_size);
    }
    int lvl = 0;
    for (;;) {
      std::vector<Entry> up;
      const size_t per = 2 * kChunkK;
      const size_t nnodes = (level.size() + per - 1) / per;
      std::vector<uint64_t> addrs(nnodes);
      for (size_t k = 0; k < nnodes; ++k) addrs[k] = w->alloc(node_bytes);
      for (size_t k = 0; k < nnodes; ++k) {
        const size_t e0 = k * per, e1 = std::min(level.size(), e0 + per);
        Bytes nb;
        nb.put("TREE", 4);
        nb.u8(1);
        nb.u8((uint8_t)lvl);
        nb.u16((uint16_t)(e1 - e0));
        nb.u64(k > 0 ? addrs[k - 1] : kUndef);
        nb.u64(k + 1 < nnodes ? addrs[k + 1] : kUndef);
        for (size_t e = e0; e < e1; ++e) {
          nb.put(level[e].key.v.data(), keysz);
          nb.u64(level[e].child);
        }
        const Bytes& last = e1 < level.size() ? level[e1].key : end_key;
        nb.put(last.v.data(), keysz);
        nb.zeros(node_bytes - nb.size());
        put(addrs[k], nb);
        up.push_back({level[e0].key, addrs[k]});
      }
      if (nnodes == 1) return addrs[0];
      level.swap(up);
      ++lvl;
    }
  }

  void write_dataset(Node& d) {
    Bytes msgs;
    int nmsg = 0;
    {
      Bytes m;  // dataspace, version 1, with maximum dimensions (= current)
      m.u8(1);
      m.u8((uint8_t)d.rank);
      m.u8(d.rank > 0 ? 1 : 0);
      m.zeros(5);
      for (int rep = 0; rep < (d.rank > 0 ? 2 : 0); ++rep)
        for (int a = 0; a < d.rank; ++a) m.u64((uint64_t)d.dims[a]);
      add_msg(msgs, 0x0001, 0, m, &nmsg);
    }
    {
      Bytes m;
      datatype_msg(m, d.type_class, d.elem_size, d.utf8);
      add_msg(msgs, 0x0003, 1, m, &nmsg);
    }
    {
      Bytes m;  // fill value, version 2: allocation time, write time, defined, size 0
      m.u8(2);
      m.u8(d.chunked ? 3 : 2);
      m.u8(d.chunked ? 0 : 2);
      m.u8(1);
      m.u32(0);
      add_msg(msgs, 0x0005, 1, m, &nmsg);
    }
    if (d.chunked && d.gzip) {
      Bytes m;  // filter pipeline, version 1: shuffle then deflate
      m.u8(1);
      m.u8(2);
      m.zeros(6);
      m.u16(2); m.u16(8); m.u16(1); m.u16(1);
      m.put("shuffle\0", 8);
      m.u32((uint32_t)d.elem_size);
      m.u32(0);
      m.u16(1); m.u16(8); m.u16(1); m.u16(1);
      m.put("deflate\0", 8);
      m.u32((uint32_t)d.gzip);
      m.u32(0);
      add_msg(msgs, 0x000b, 0, m, &nmsg);
    }
    {
      Bytes m;  // layout, version 3
      m.u8(3);
      if (d.chunked) {
        m.u8(2);
        m.u8((uint8_t)(d.rank + 1));
        m.u64(d.chunks.empty() ? kUndef : write_chunk_btree(d));
        for (int a = 0; a < d.rank; ++a) m.u32((uint32_t)d.chunk[a]);
        m.u32((uint32_t)d.elem_size);
      } else {
        m.u8(1);
        m.u64(d.data_addr);
        m.u64(d.data_size);
      }
      add_msg(msgs, 0x0008, 0, m, &nmsg);
    }
    {
      Bytes m;  // modification time, version 1
      m.u8(1);
      m.zeros(3);
      m.u32((uint32_t)time(nullptr));
      add_msg(msgs, 0x0012, 0, m, &nmsg);
    }
    attr_msgs(d, msgs, &nmsg);
    const Bytes h = finish_header(msgs, nmsg);
    d.header_addr = w->alloc(h.size());
    put(d.header_addr, h);
  }

  void write_group(Node& g) {
    for (auto& kv : g.kids) {
      if (kv.second->is_group) write_group(*kv.second);
      else write_dataset(*kv.second);
    }
    // local heap: "" at offset 0, then the link names (std::map keeps them in strcmp order), then one free block
    Bytes heap;
    heap.zeros(8);
    std::vector<uint64_t> name_off;
    for (auto& kv : g.kids) {
      name_off.push_back(heap.size());
      heap.put(kv.first.c_str(), kv.first.size() + 1);
      heap.pad8();
    }
    const uint64_t free_off = heap.size();
    const uint64_t free_len = 32;
    heap.u64(1);  // no next free block
    heap.u64(free_len);
    heap.zeros(free_len - 16);
    const uint64_t heap_hdr = w->alloc(32);
    const uint64_t heap_data = w->alloc(heap.size());
    {
      Bytes h;
      h.put("HEAP", 4);
      h.u8(0);
      h.zeros(3);
      h.u64(heap.size());
      h.u64(free_off);
      h.u64(heap_data);
      put(heap_hdr, h);
      put(heap_data, heap);
    }
    // symbol nodes of up to 2 * kLeafK entries
    std::vector<Node*> kids;
    for (auto& kv : g.kids) kids.push_back(kv.second.get());
    const size_t per = 2 * kLeafK;
    const size_t nsn = (kids.size() + per - 1) / per;
    const size_t snod_bytes = 8 + per * 40;
    std::vector<uint64_t> snod_addr;
    std::vector<uint64_t> last_name;  // heap offset of the largest name in each symbol node
    for (size_t k = 0; k < nsn; ++k) {
      const size_t e0 = k * per, e1 = std::min(kids.size(), e0 + per);
      Bytes s;
      s.put("SNOD", 4);
      s.u8(1);
      s.u8(0);
      s.u16((uint16_t)(e1 - e0));
      for (size_t e = e0; e < e1; ++e) {
        s.u64(name_off[e]);
        s.u64(kids[e]->header_addr);
        if (kids[e]->is_group) {
          s.u32(1);
          s.u32(0);
          s.u64(kids[e]->btree_addr);
          s.u64(kids[e]->heap_addr);
        } else {
          s.u32(0);
          s.u32(0);
          s.zeros(16);
        }
      }
      s.zeros(snod_bytes - s.size());
      const uint64_t a = w->alloc(snod_bytes);
      put(a, s);
      snod_addr.push_back(a);
      last_name.push_back(name_off[e1 - 1]);
    }
    // one B-tree node over the symbol nodes
    const size_t bt_bytes = 24 + (size_t)(2 * kInternalK + 1) * 8 + (size_t)(2 * kInternalK) * 8;
    Bytes bt;
    bt.put("TREE", 4);
    bt.u8(0);
    bt.u8(0);
    bt.u16((uint16_t)nsn);
    bt.u64(kUndef);
    bt.u64(kUndef);
    bt.u64(0);
    for (size_t k = 0; k < nsn; ++k) {
      bt.u64(snod_addr[k]);
      bt.u64(last_name[k]);
    }
    bt.zeros(bt_bytes - bt.size());
    g.btree_addr = w->alloc(bt_bytes);
    put(g.btree_addr, bt);
    g.heap_addr = heap_hdr;
    // object header: symbol table message + attributes
    Bytes msgs;
    int nmsg = 0;
    {
      Bytes m;
      m.u64(g.btree_addr);
      m.u64(g.heap_addr);
      add_msg(msgs, 0x0011, 0, m, &nmsg);
    }
    attr_msgs(g, msgs, &nmsg);
    const Bytes h = finish_header(msgs, nmsg);
    g.header_addr = w->alloc(h.size());
    put(g.header_addr, h);
  }

  int run() {
    std::vector<Node*> all;
    collect(w->root, all);
    write_global_heaps(all);
    write_group(w->root);
    w->eof = (w->eof + 7) & ~7ull;
    Bytes sb;
    static const uint8_t sig[8] = {0x89, 'H', 'D', 'F', '\r', '\n', 0x1a, '\n'};
    sb.put(sig, 8);
    sb.u8(0); sb.u8(0); sb.u8(0); sb.u8(0); sb.u8(0);
    sb.u8(8); sb.u8(8); sb.u8(0);
    sb.u16(kLeafK);
    sb.u16(kInternalK);
    sb.u32(0);
    sb.u64(0);
    sb.u64(kUndef);
    sb.u64(w->eof);
    sb.u64(kUndef);
    sb.u64(0);
    sb.u64(w->root.header_addr);
    sb.u32(1);
    sb.u32(0);
    sb.u64(w->root.btree_addr);
    sb.u64(w->root.heap_addr);
    put(0, sb);
    if (ok && ftruncate(w->fd, (off_t)w->eof) != 0) ok = false;
    return ok ? RSK_H5_OK : RSK_H5_ERR_OPEN;
  }
};

}  // namespace

extern "C" int rsk_h5w_close(RskH5Writer* w) {
  if (!w) return RSK_H5_ERR_NULL;
  Closer c{w};
  int rc;
  try {
    rc = c.run();
  } catch (...) {
    rc = RSK_H5_ERR_STATE;
  }
  delete w;
  return rc;
}
