// rsk_h5.cpp -- SkyH5 (HDF5 subset) reader and the file -> HBM streaming path.  See include/rsk_io.h for
// the supported subset and the reference code this replaces (h5py calls in
// src/pyradiosky/skymodel.py:3491-3766).  Written from the published HDF5 file-format specification;
// libhdf5 is not used.  The writer lives in rsk_h5_write.cpp.
#include <cuda_runtime.h>
#include <fcntl.h>
#include <stdint.h>
#include <string.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>
#include <zlib.h>

#include <algorithm>
#include <atomic>
#include <condition_variable>
#include <functional>
#include <map>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "../../include/rsk_io.h"

namespace {

constexpr uint64_t kUndef = ~0ull;

struct H5Err {
  int code;
  std::string msg;
};

[[noreturn]] void fail(int code, const std::string& msg) { throw H5Err{code, msg}; }

struct Dtype {
  int cls = -1;       // RSK_H5_INT ...
  int size = 0;       // bytes per element in the file
  int consumed = 0;   // bytes of the datatype message used
};

struct Filter {
  int id;
  std::vector<uint32_t> cd;
};

struct Msg {
  uint16_t type;
  const uint8_t* p;
  size_t len;
};

struct ChunkRec {
  uint64_t addr;
  uint32_t size;
  uint32_t mask;
  uint64_t off[8];
};

struct Dataset {
  int rank = 0;
  bool null_space = false;
  int64_t dims[8] = {0};
  int64_t count = 1;
  Dtype dt;
  int layout = -1;
  const uint8_t* compact = nullptr;
  uint64_t compact_size = 0;
  uint64_t addr = kUndef;  // contiguous data or chunk B-tree
  int64_t chunk[9] = {0};  // rank + 1 entries (the last is the element size), rank <= 8
  int chunk_nd = 0;
  std::vector<Filter> filters;
  std::vector<ChunkRec> chunks;
  bool chunks_loaded = false;
  std::vector<Msg> attrs;
};

enum Conv { kRaw = 0, kF64 = 1, kI64 = 2 };

// n elements of file type dt at src -> dst (raw copy, double, or int64)
void convert(const Dtype& dt, const uint8_t* src, void* dst, int conv, int64_t n) {
  if (conv == kRaw) {
    memcpy(dst, src, (size_t)n * dt.size);
    return;
  }
  auto run = [&](auto load) {
    if (conv == kF64) {
      double* d = static_cast<double*>(dst);
      for (int64_t i = 0; i < n; ++i) d[i] = (double)load(i);
    } else {
      int64_t* d = static_cast<int64_t*>(dst);
      for (int64_t i = 0; i < n; ++i) d[i] = (int64_t)load(i);
    }
  };
#define RSK_LOAD(T) run([&](int64_t i) { T v; memcpy(&v, src + i * sizeof(T), sizeof(T)); return v; })
  if (dt.cls == RSK_H5_FLOAT && dt.size == 8) {
    if (conv == kF64) memcpy(dst, src, (size_t)n * 8); else RSK_LOAD(double);
  } else if (dt.cls == RSK_H5_FLOAT && dt.size == 4) RSK_LOAD(float);
  else if (dt.cls == RSK_H5_INT && dt.size == 8) {
    if (conv == kI64) memcpy(dst, src, (size_t)n * 8); else RSK_LOAD(int64_t);
  } else if (dt.cls == RSK_H5_INT && dt.size == 4) RSK_LOAD(int32_t);
  else if (dt.cls == RSK_H5_INT && dt.size == 2) RSK_LOAD(int16_t);
  else if (dt.cls == RSK_H5_INT && dt.size == 1) RSK_LOAD(int8_t);
  else if (dt.cls == RSK_H5_UINT && dt.size == 8) RSK_LOAD(uint64_t);
  else if (dt.cls == RSK_H5_UINT && dt.size == 4) RSK_LOAD(uint32_t);
  else if (dt.cls == RSK_H5_UINT && dt.size == 2) RSK_LOAD(uint16_t);
  else if (dt.cls == RSK_H5_UINT && dt.size == 1) RSK_LOAD(uint8_t);
  else fail(RSK_H5_ERR_TYPE, "dataset type cannot be converted to a number");
#undef RSK_LOAD
}

// A few worker threads that live for one file -> HBM read (creating threads per slab costs more than
// filling the slab).  run(count, body) calls body(i, thread) for i in [0, count) and returns when all are done;
// the first exception is re-thrown on the caller's thread.
class Pool {
 public:
  explicit Pool(int n) : n_(n < 1 ? 1 : n) {
    for (int t = 1; t < n_; ++t) th_.emplace_back([this, t] { loop(t); });
  }
  ~Pool() {
    {
      std::lock_guard<std::mutex> g(m_);
      stop_ = true;
      ++gen_;
    }
    cv_.notify_all();
    for (auto& t : th_) t.join();
  }
  int size() const { return n_; }
  void run(int64_t count, const std::function<void(int64_t, int)>& body) {
    {
      std::lock_guard<std::mutex> g(m_);
      job_ = &body;
      count_ = count;
      next_.store(0);
      pending_ = n_ - 1;
      err_ = H5Err{0, ""};
      ++gen_;
    }
    cv_.notify_all();
    work(0);
    std::unique_lock<std::mutex> g(m_);
    done_.wait(g, [this] { return pending_ == 0; });
    job_ = nullptr;
    if (err_.code != 0) throw err_;
  }

 private:
  void work(int t) {
    try {
      for (;;) {
        const int64_t i = next_.fetch_add(1);
        if (i >= count_) break;
        (*job_)(i, t);
      }
    } catch (const H5Err& e) {
      std::lock_guard<std::mutex> g(m_);
      if (err_.code == 0) err_ = e;
      next_.store(count_);
    } catch (const std::exception& e) {
      std::lock_guard<std::mutex> g(m_);
      if (err_.code == 0) err_ = H5Err{RSK_H5_ERR_FORMAT, e.what()};
      next_.store(count_);
    }
  }
  void loop(int t) {
    int seen = 0;
    for (;;) {
      {
        std::unique_lock<std::mutex> g(m_);
        cv_.wait(g, [&] { return gen_ != seen; });
        seen = gen_;
        if (stop_) return;
      }
      work(t);
      {
        std::lock_guard<std::mutex> g(m_);
        --pending_;
      }
      done_.notify_one();
    }
  }
  int n_;
  std::vector<std::thread> th_;
  std::mutex m_;
  std::condition_variable cv_, done_;
  const std::function<void(int64_t, int)>* job_ = nullptr;
  int64_t count_ = 0;
  std::atomic<int64_t> next_{0};
  int pending_ = 0;
  int gen_ = 0;
  bool stop_ = false;
  H5Err err_{0, ""};
};

}  // namespace

struct RskH5File {
  int fd = -1;
  const uint8_t* b = nullptr;
  uint64_t size = 0;
  uint64_t base = 0;
  uint64_t root_addr = 0;
  std::string err;
  std::map<uint64_t, std::map<std::string, uint64_t>> group_cache;
  std::map<std::string, Dataset> ds_cache;
  // pinned staging ring of rsk_h5_read_device, kept for the life of the handle (pinning memory is slow)
  static constexpr int kRing = 3;
  uint8_t* stage[kRing] = {nullptr, nullptr, nullptr};
  cudaEvent_t stage_ev[kRing] = {nullptr, nullptr, nullptr};
  size_t stage_bytes = 0;

  ~RskH5File() {
    for (int i = 0; i < kRing; ++i) {
      if (stage_ev[i]) cudaEventDestroy(stage_ev[i]);
      if (stage[i]) cudaFreeHost(stage[i]);
    }
    if (b) munmap(const_cast<uint8_t*>(b), size);
    if (fd >= 0) close(fd);
  }

  void ensure_stage(size_t bytes) {
    if (bytes <= stage_bytes && stage[0]) return;
    for (int i = 0; i < kRing; ++i) {
      if (stage[i]) cudaFreeHost(stage[i]);
      stage[i] = nullptr;
    }
    stage_bytes = 0;
    for (int i = 0; i < kRing; ++i) {
      cudaError_t e = cudaHostAlloc(reinterpret_cast<void**>(&stage[i]), bytes, cudaHostAllocDefault);
      if (e == cudaSuccess && !stage_ev[i]) e = cudaEventCreateWithFlags(&stage_ev[i], cudaEventDisableTiming);
      if (e != cudaSuccess) throw H5Err{(int)e, cudaGetErrorString(e)};
    }
    stage_bytes = bytes;
  }

  const uint8_t* at(uint64_t addr, uint64_t len) const {
    if (addr == kUndef || addr + base < addr || addr + base + len < addr + base || addr + base + len > size)
      fail(RSK_H5_ERR_FORMAT, "address outside the file");
    return b + base + addr;
  }
  static uint16_t u16(const uint8_t* p) { uint16_t v; memcpy(&v, p, 2); return v; }
  static uint32_t u32(const uint8_t* p) { uint32_t v; memcpy(&v, p, 4); return v; }
  static uint64_t u64(const uint8_t* p) { uint64_t v; memcpy(&v, p, 8); return v; }

  void open(const char* path) {
    fd = ::open(path, O_RDONLY);
    if (fd < 0) fail(RSK_H5_ERR_OPEN, std::string("cannot open ") + path);
    struct stat st;
    if (fstat(fd, &st) != 0 || st.st_size < 96) fail(RSK_H5_ERR_FORMAT, "file too small to be HDF5");
    size = (uint64_t)st.st_size;
    void* m = mmap(nullptr, size, PROT_READ, MAP_PRIVATE, fd, 0);
    if (m == MAP_FAILED) fail(RSK_H5_ERR_OPEN, "mmap failed");
    b = static_cast<const uint8_t*>(m);
    static const uint8_t sig[8] = {0x89, 'H', 'D', 'F', '\r', '\n', 0x1a, '\n'};
    if (memcmp(b, sig, 8) != 0) fail(RSK_H5_ERR_FORMAT, "not an HDF5 file (signature; a user block is not supported)");
    const int ver = b[8];
    if (ver > 1)
      fail(RSK_H5_ERR_UNSUPPORTED, "superblock version " + std::to_string(ver) +
                                       " (libver='latest' files are not supported; rewrite with libver='earliest')");
    if (b[13] != 8 || b[14] != 8) fail(RSK_H5_ERR_UNSUPPORTED, "only 8-byte offsets and lengths are supported");
    size_t pos = 24 + (ver == 1 ? 4 : 0);
    base = u64(b + pos);
    if (base != 0) fail(RSK_H5_ERR_UNSUPPORTED, "non-zero base address");
    pos += 32;
    root_addr = u64(b + pos + 8);
  }

  // ---- object headers ----------------------------------------------------------------------------
  std::vector<Msg> messages(uint64_t addr) {
    const uint8_t* h = at(addr, 16);
    if (h[0] != 1) {
      if (memcmp(h, "OHDR", 4) == 0) fail(RSK_H5_ERR_UNSUPPORTED, "version 2 object headers are not supported");
      fail(RSK_H5_ERR_FORMAT, "bad object header");
    }
    const int nmsg = u16(h + 2);
    const uint32_t hsize = u32(h + 8);
    std::vector<std::pair<uint64_t, uint64_t>> blocks{{addr + 16, hsize}};
    std::vector<Msg> out;
    for (size_t bi = 0; bi < blocks.size() && (int)out.size() < nmsg; ++bi) {
      uint64_t p = blocks[bi].first;
      const uint64_t end = p + blocks[bi].second;
      at(p, blocks[bi].second);
      while (p + 8 <= end && (int)out.size() < nmsg) {
        const uint8_t* m = at(p, 8);
        const uint16_t type = u16(m), sz = u16(m + 2);
        const uint8_t flags = m[4];
        if (p + 8 + sz > end) fail(RSK_H5_ERR_FORMAT, "object header message overruns its block");
        const uint8_t* body = at(p + 8, sz);
        if (flags & 0x02) fail(RSK_H5_ERR_UNSUPPORTED, "shared object header messages are not supported");
        if (type == 0x0010) {
          if (sz < 16) fail(RSK_H5_ERR_FORMAT, "short continuation message");
          blocks.push_back({u64(body), u64(body + 8)});
        }
        out.push_back({type, body, sz});
        p += 8 + sz;
      }
    }
    return out;
  }

  // ---- groups ------------------------------------------------------------------------------------
  void walk_group(uint64_t node, const uint8_t* heap, uint64_t heap_size, std::map<std::string, uint64_t>& links,
                  int depth) {
    if (depth > 32) fail(RSK_H5_ERR_FORMAT, "group B-tree too deep");
    const uint8_t* p = at(node, 8);
    if (memcmp(p, "TREE", 4) == 0) {
      if (p[4] != 0) fail(RSK_H5_ERR_FORMAT, "expected a group B-tree node");
      const int n = u16(p + 6);
      const uint8_t* e = at(node + 24, (uint64_t)n * 16 + 8);
      for (int i = 0; i < n; ++i) walk_group(u64(e + 16 * i + 8), heap, heap_size, links, depth + 1);
    } else if (memcmp(p, "SNOD", 4) == 0) {
      const int n = u16(p + 6);
      const uint8_t* e = at(node + 8, (uint64_t)n * 40);
      for (int i = 0; i < n; ++i) {
        const uint64_t noff = u64(e + 40 * i), oaddr = u64(e + 40 * i + 8);
        if (noff >= heap_size) fail(RSK_H5_ERR_FORMAT, "link name outside the local heap");
        const void* z = memchr(heap + noff, 0, heap_size - noff);
        if (!z) fail(RSK_H5_ERR_FORMAT, "unterminated link name");
        links[std::string(reinterpret_cast<const char*>(heap + noff))] = oaddr;
      }
    } else {
      fail(RSK_H5_ERR_FORMAT, "bad group node signature");
    }
  }

  // links of the group whose object header is at addr; nullptr if the object is not a group
  const std::map<std::string, uint64_t>* group_links(uint64_t addr) {
    auto it = group_cache.find(addr);
    if (it != group_cache.end()) return &it->second;
    const std::vector<Msg> msgs = messages(addr);
    for (const Msg& m : msgs) {
      if (m.type == 0x0011) {
        if (m.len < 16) fail(RSK_H5_ERR_FORMAT, "short symbol table message");
        const uint64_t bt = u64(m.p), hp = u64(m.p + 8);
        const uint8_t* h = at(hp, 32);
        if (memcmp(h, "HEAP", 4) != 0) fail(RSK_H5_ERR_FORMAT, "bad local heap");
        const uint64_t hsize = u64(h + 8), daddr = u64(h + 24);
        const uint8_t* heap = at(daddr, hsize);
        std::map<std::string, uint64_t> links;
        walk_group(bt, heap, hsize, links, 0);
        return &(group_cache[addr] = std::move(links));
      }
      if (m.type == 0x0002 || m.type == 0x0006)
        fail(RSK_H5_ERR_UNSUPPORTED, "new-style (link message) groups are not supported");
    }
    return nullptr;
  }

  // object header address of path, or kUndef
  uint64_t resolve(const char* path) {
    uint64_t cur = root_addr;
    std::string s(path ? path : "");
    size_t i = 0;
    while (i < s.size()) {
      while (i < s.size() && s[i] == '/') ++i;
      size_t j = i;
      while (j < s.size() && s[j] != '/') ++j;
      if (j == i) break;
      const auto* links = group_links(cur);
      if (!links) return kUndef;
      auto it = links->find(s.substr(i, j - i));
      if (it == links->end()) return kUndef;
      cur = it->second;
      i = j;
    }
    return cur;
  }

  // ---- message decoders --------------------------------------------------------------------------
  Dtype datatype(const uint8_t* p, size_t len) {
    if (len < 8) fail(RSK_H5_ERR_FORMAT, "short datatype message");
    const int cls = p[0] & 0x0f, ver = p[0] >> 4;
    const uint32_t bits = p[1] | (p[2] << 8) | (p[3] << 16);
    Dtype d;
    d.size = (int)u32(p + 4);
    d.consumed = 8;
    switch (cls) {
      case 0:
        if (bits & 1) fail(RSK_H5_ERR_UNSUPPORTED, "big-endian integers");
        d.cls = (bits & 8) ? RSK_H5_INT : RSK_H5_UINT;
        d.consumed += 4;
        break;
      case 1:
        if (bits & 1) fail(RSK_H5_ERR_UNSUPPORTED, "big-endian floats");
        if (d.size != 4 && d.size != 8) fail(RSK_H5_ERR_UNSUPPORTED, "float size");
        d.cls = RSK_H5_FLOAT;
        d.consumed += 12;
        break;
      case 3:
        d.cls = RSK_H5_STRING;
        break;
      case 9: {
        if ((bits & 0x0f) != 1) fail(RSK_H5_ERR_UNSUPPORTED, "variable-length sequences");
        const Dtype basetype = datatype(p + 8, len - 8);
        d.cls = RSK_H5_VLEN_STRING;
        d.consumed += basetype.consumed;
        break;
      }
      case 8: {  // enum (h5py bool): treated as its integer base type
        const Dtype basetype = datatype(p + 8, len - 8);
        if (basetype.cls != RSK_H5_INT && basetype.cls != RSK_H5_UINT) fail(RSK_H5_ERR_UNSUPPORTED, "enum base type");
        const int nmemb = bits & 0xffff;
        size_t q = 8 + basetype.consumed;
        for (int k = 0; k < nmemb; ++k) {
          const void* z = q < len ? memchr(p + q, 0, len - q) : nullptr;
          if (!z) fail(RSK_H5_ERR_FORMAT, "bad enum member name");
          size_t nl = (size_t)(static_cast<const uint8_t*>(z) - (p + q)) + 1;
          if (ver < 3) nl = (nl + 7) & ~(size_t)7;
          q += nl;
        }
        q += (size_t)nmemb * basetype.size;
        d.cls = basetype.cls;
        d.consumed = (int)q;
        break;
      }
      default:
        fail(RSK_H5_ERR_UNSUPPORTED, "datatype class " + std::to_string(cls));
    }
    return d;
  }

  static void dataspace(const uint8_t* p, size_t len, int* rank, int64_t* dims, bool* null_space) {
    if (len < 4) fail(RSK_H5_ERR_FORMAT, "short dataspace message");
    const int ver = p[0];
    *rank = p[1];
    *null_space = false;
    size_t q;
    if (ver == 1) q = 8;
    else if (ver == 2) {
      q = 4;
      if (p[3] == 2) *null_space = true;
    } else fail(RSK_H5_ERR_UNSUPPORTED, "dataspace version");
    if (*rank > 8) fail(RSK_H5_ERR_UNSUPPORTED, "rank > 8");
    if (len < q + 8 * (size_t)*rank) fail(RSK_H5_ERR_FORMAT, "short dataspace message");
    for (int i = 0; i < *rank; ++i) dims[i] = (int64_t)u64(p + q + 8 * i);
  }

  Dataset& dataset(const char* path) {
    const std::string key(path ? path : "");
    auto it = ds_cache.find(key);
    if (it != ds_cache.end()) return it->second;
    const uint64_t addr = resolve(path);
    if (addr == kUndef) fail(RSK_H5_ERR_NOTFOUND, "no object named " + key);
    const std::vector<Msg> msgs = messages(addr);
    Dataset d;
    bool have_space = false, have_type = false;
    for (const Msg& m : msgs) {
      if (m.type == 0x0001) {
        dataspace(m.p, m.len, &d.rank, d.dims, &d.null_space);
        have_space = true;
      } else if (m.type == 0x0003) {
        d.dt = datatype(m.p, m.len);
        have_type = true;
      } else if (m.type == 0x0008) {
        parse_layout(m, d);
      } else if (m.type == 0x000b) {
        parse_filters(m, d);
      } else if (m.type == 0x000c) {
        d.attrs.push_back(m);
      }
    }
    if (d.layout < 0 || !have_space || !have_type) fail(RSK_H5_ERR_TYPE, key + " is not a dataset");
    d.count = d.null_space ? 0 : 1;
    for (int i = 0; i < d.rank; ++i) {
      if (d.dims[i] < 0 || (d.dims[i] > 0 && d.count > (int64_t)(INT64_MAX / 16) / d.dims[i]))
        fail(RSK_H5_ERR_FORMAT, "dataset extent overflows");
      d.count *= d.dims[i];
    }
    if (d.layout == 2) {
      if (d.chunk_nd != d.rank + 1) fail(RSK_H5_ERR_FORMAT, "chunk rank differs from the dataspace rank");
      int64_t cb = d.dt.size > 0 ? d.dt.size : 1;
      for (int a = 0; a < d.rank; ++a) {
        if (d.chunk[a] <= 0 || cb > ((int64_t)1 << 40) / d.chunk[a]) fail(RSK_H5_ERR_FORMAT, "bad chunk shape");
        cb *= d.chunk[a];
      }
      if (d.chunk[d.rank] != d.dt.size) fail(RSK_H5_ERR_FORMAT, "chunk element size differs from the datatype size");
    }
    return ds_cache[key] = std::move(d);
  }

  void parse_layout(const Msg& m, Dataset& d) {
    const uint8_t* p = m.p;
    if (m.len < 2) fail(RSK_H5_ERR_FORMAT, "short layout message");
    const int ver = p[0];
    if (ver == 3) {
      d.layout = p[1];
      if (d.layout == 0) {
        d.compact_size = u16(p + 2);
        if (m.len < 4 + d.compact_size) fail(RSK_H5_ERR_FORMAT, "short compact layout");
        d.compact = p + 4;
      } else if (d.layout == 1) {
        if (m.len < 18) fail(RSK_H5_ERR_FORMAT, "short contiguous layout");
        d.addr = u64(p + 2);
      } else if (d.layout == 2) {
        const int nd = p[2];
        if (nd < 1 || nd > 9 || m.len < 11 + 4 * (size_t)nd) fail(RSK_H5_ERR_FORMAT, "bad chunked layout");
        d.addr = u64(p + 3);
        for (int i = 0; i < nd; ++i) d.chunk[i] = u32(p + 11 + 4 * i);
        d.chunk_nd = nd;
      } else fail(RSK_H5_ERR_UNSUPPORTED, "layout class");
    } else if (ver == 1 || ver == 2) {
      if (m.len < 8) fail(RSK_H5_ERR_FORMAT, "short layout message");
      const int nd = p[1];
      d.layout = p[2];
      size_t q = 8;
      if (d.layout != 0) {
        d.addr = u64(p + q);
        q += 8;
      }
      if (nd > 9 || m.len < q + 4 * (size_t)nd) fail(RSK_H5_ERR_FORMAT, "bad layout message");
      for (int i = 0; i < nd; ++i) d.chunk[i] = u32(p + q + 4 * i);
      d.chunk_nd = nd;
      q += 4 * (size_t)nd;
      if (d.layout == 0) {
        d.compact_size = u32(p + q);
        d.compact = p + q + 4;
        if (m.len < q + 4 + d.compact_size) fail(RSK_H5_ERR_FORMAT, "short compact layout");
      }
    } else {
      fail(RSK_H5_ERR_UNSUPPORTED, "layout message version " + std::to_string(ver) +
                                       " (libver='latest' chunk indexes are not supported)");
    }
  }

  void parse_filters(const Msg& m, Dataset& d) {
    const uint8_t* p = m.p;
    if (m.len < 2) fail(RSK_H5_ERR_FORMAT, "short filter message");
    const int ver = p[0], n = p[1];
    size_t q = ver == 1 ? 8 : 2;
    for (int k = 0; k < n; ++k) {
      if (q + 8 > m.len) fail(RSK_H5_ERR_FORMAT, "short filter message");
      Filter f;
      f.id = u16(p + q);
      q += 2;
      size_t nlen = 0;
      if (ver == 1 || f.id >= 256) {
        nlen = u16(p + q);
        q += 2;
      }
      q += 2;  // flags
      const int ncd = u16(p + q);
      q += 2;
      if (ver == 1) nlen = (nlen + 7) & ~(size_t)7;
      q += nlen;
      if (q + 4 * (size_t)ncd > m.len) fail(RSK_H5_ERR_FORMAT, "short filter message");
      for (int c = 0; c < ncd; ++c) f.cd.push_back(u32(p + q + 4 * c));
      q += 4 * (size_t)ncd;
      if (ver == 1 && (ncd & 1)) q += 4;
      d.filters.push_back(std::move(f));
    }
  }

  // ---- chunk index -------------------------------------------------------------------------------
  void walk_chunks(uint64_t node, Dataset& d, int depth) {
    if (node == kUndef) return;  // no chunk allocated yet
    if (depth > 32) fail(RSK_H5_ERR_FORMAT, "chunk B-tree too deep");
    const uint8_t* p = at(node, 24);
    if (memcmp(p, "TREE", 4) != 0 || p[4] != 1) fail(RSK_H5_ERR_FORMAT, "bad chunk B-tree node");
    const int level = p[5], n = u16(p + 6);
    const size_t keysz = 8 + 8 * (size_t)(d.rank + 1);
    const uint8_t* e = at(node + 24, (uint64_t)n * (keysz + 8) + keysz);
    for (int i = 0; i < n; ++i) {
      const uint8_t* k = e + i * (keysz + 8);
      const uint64_t child = u64(k + keysz);
      if (level > 0) {
        walk_chunks(child, d, depth + 1);
      } else {
        ChunkRec r;
        r.size = u32(k);
        r.mask = u32(k + 4);
        for (int a = 0; a < d.rank; ++a) {
          r.off[a] = u64(k + 8 + 8 * a);
          if ((int64_t)r.off[a] >= d.dims[a] || r.off[a] % (uint64_t)d.chunk[a] != 0)
            fail(RSK_H5_ERR_FORMAT, "chunk offset outside the dataset");
        }
        r.addr = child;
        at(child, r.size);
        d.chunks.push_back(r);
      }
    }
  }

  void load_chunks(Dataset& d) {
    if (d.chunks_loaded) return;
    for (int a = 0; a < d.rank; ++a)
      if (d.chunk[a] <= 0) fail(RSK_H5_ERR_FORMAT, "bad chunk shape");
    walk_chunks(d.addr, d, 0);
    d.chunks_loaded = true;
  }

  int64_t chunk_bytes(const Dataset& d) const {
    int64_t n = d.dt.size;
    for (int a = 0; a < d.rank; ++a) n *= d.chunk[a];
    return n;
  }

  // Decoded bytes of one chunk: a pointer into the mapping when no filter applies, else into scratch.
  const uint8_t* decode_chunk(const Dataset& d, const ChunkRec& r, std::vector<uint8_t>& s0, std::vector<uint8_t>& s1) {
    const uint8_t* cur = b + base + r.addr;
    size_t cur_len = r.size;
    const size_t want = (size_t)chunk_bytes(d);
    std::vector<uint8_t>* bufs[2] = {&s0, &s1};
    int which = 0;
    for (int k = (int)d.filters.size() - 1; k >= 0; --k) {
      if (r.mask & (1u << k)) continue;
      const Filter& f = d.filters[k];
      std::vector<uint8_t>& out = *bufs[which];
      if (f.id == 1) {  // deflate
        out.resize(want + 64);
        uLongf dl = (uLongf)out.size();
        const int zr = uncompress(out.data(), &dl, cur, (uLong)cur_len);
        if (zr != Z_OK) fail(RSK_H5_ERR_FILTER, "deflate stream is corrupt");
        cur_len = dl;
      } else if (f.id == 2) {  // shuffle
        const size_t es = f.cd.empty() ? (size_t)d.dt.size : f.cd[0];
        const size_t ne = es ? cur_len / es : 0;
        out.resize(cur_len);
        for (size_t byte = 0; byte < es; ++byte) {
          const uint8_t* src = cur + byte * ne;
          for (size_t i = 0; i < ne; ++i) out[i * es + byte] = src[i];
        }
        memcpy(out.data() + ne * es, cur + ne * es, cur_len - ne * es);
      } else {
        fail(RSK_H5_ERR_UNSUPPORTED, "filter id " + std::to_string(f.id));
      }
      cur = out.data();
      which ^= 1;
    }
    if (cur_len < want) fail(RSK_H5_ERR_FORMAT, "chunk shorter than its shape");
    return cur;
  }

  // Scatter one decoded chunk into a destination laid out as the (sub)array
  // [dst_lo[a], dst_lo[a] + dst_ext[a]) of the dataset, C order; elements outside it are skipped.
  // file_off >= 0: the chunk is stored unfiltered at that file offset and needs no conversion, so long
  // rows are pread() straight into the destination (no page faults on the mapping, one copy).
  void scatter_chunk(const Dataset& d, const ChunkRec& r, const uint8_t* data, uint8_t* dst, const int64_t* dst_lo,
                     const int64_t* dst_ext, int conv, int out_size, int64_t file_off = -1) {
    const int rank = d.rank;
    if (rank == 0) {
      convert(d.dt, data, dst, conv, 1);
      return;
    }
    int64_t lo[8], hi[8];  // intersection, in dataset coordinates
    for (int a = 0; a < rank; ++a) {
      lo[a] = std::max<int64_t>((int64_t)r.off[a], dst_lo[a]);
      hi[a] = std::min<int64_t>(std::min<int64_t>((int64_t)r.off[a] + d.chunk[a], d.dims[a]), dst_lo[a] + dst_ext[a]);
      if (hi[a] <= lo[a]) return;
    }
    const int last = rank - 1;
    const int64_t run = hi[last] - lo[last];
    int64_t idx[8];
    for (int a = 0; a < rank; ++a) idx[a] = lo[a];
    for (;;) {
      int64_t so = 0, dof = 0;
      for (int a = 0; a < rank; ++a) {
        so = so * d.chunk[a] + (idx[a] - (int64_t)r.off[a]);
        dof = dof * dst_ext[a] + (idx[a] - dst_lo[a]);
      }
      if (file_off >= 0 && run * d.dt.size >= 16384) {
        uint8_t* to = dst + dof * out_size;
        int64_t left = run * d.dt.size, fo = file_off + so * d.dt.size;
        while (left > 0) {
          const ssize_t got = pread(fd, to, (size_t)left, (off_t)fo);
          if (got <= 0) fail(RSK_H5_ERR_OPEN, "read failed");
          to += got;
          fo += got;
          left -= got;
        }
      } else {
        convert(d.dt, data + so * d.dt.size, dst + dof * out_size, conv, run);
      }
      int a = last - 1;
      for (; a >= 0; --a) {
        if (++idx[a] < hi[a]) break;
        idx[a] = lo[a];
      }
      if (a < 0) break;
    }
  }

  // decode (if filtered) and scatter one chunk
  void place_chunk(const Dataset& d, const ChunkRec& r, std::vector<uint8_t>& s0, std::vector<uint8_t>& s1, uint8_t* dst,
                   const int64_t* dst_lo, const int64_t* dst_ext, int conv, int out_size) {
    bool filtered = false;
    for (size_t k = 0; k < d.filters.size(); ++k) filtered = filtered || !(r.mask & (1u << k));
    const bool identity = conv == kRaw || (conv == kF64 && d.dt.cls == RSK_H5_FLOAT && d.dt.size == 8) ||
                          (conv == kI64 && d.dt.cls == RSK_H5_INT && d.dt.size == 8);
    if (!filtered && identity) {
      if ((int64_t)r.size < chunk_bytes(d)) fail(RSK_H5_ERR_FORMAT, "chunk shorter than its shape");
      scatter_chunk(d, r, b + base + r.addr, dst, dst_lo, dst_ext, conv, out_size, (int64_t)(base + r.addr));
    } else {
      scatter_chunk(d, r, decode_chunk(d, r, s0, s1), dst, dst_lo, dst_ext, conv, out_size);
    }
  }

  template <class F>
  static void parallel_for(int64_t n, int nthreads, F&& body) {
    if (nthreads < 1) nthreads = 1;
    if (nthreads > 64) nthreads = 64;
    if ((int64_t)nthreads > n) nthreads = (int)std::max<int64_t>(n, 1);
    std::atomic<int64_t> next{0};
    std::vector<H5Err> errs((size_t)nthreads, H5Err{0, ""});
    auto worker = [&](int t) {
      try {
        for (;;) {
          const int64_t i = next.fetch_add(1);
          if (i >= n) break;
          body(i, t);
        }
      } catch (const H5Err& e) {
        errs[(size_t)t] = e;
        next.store(n);
      } catch (const std::exception& e) {
        errs[(size_t)t] = H5Err{RSK_H5_ERR_FORMAT, e.what()};
        next.store(n);
      }
    };
    if (nthreads == 1) {
      worker(0);
    } else {
      std::vector<std::thread> th;
      for (int t = 1; t < nthreads; ++t) th.emplace_back(worker, t);
      worker(0);
      for (auto& t : th) t.join();
    }
    for (const H5Err& e : errs)
      if (e.code != 0) throw e;
  }

  // Whole dataset -> host destination (raw bytes or converted)
  void read_host(Dataset& d, void* dst, int conv, int nthreads) {
    const int out_size = conv == kRaw ? d.dt.size : 8;
    if (d.count == 0) return;
    if (d.dt.cls == RSK_H5_VLEN_STRING && conv != kRaw) fail(RSK_H5_ERR_TYPE, "string dataset");
    if (d.layout == 0) {
      if ((int64_t)d.compact_size < d.count * d.dt.size) fail(RSK_H5_ERR_FORMAT, "compact data too short");
      convert(d.dt, d.compact, dst, conv, d.count);
    } else if (d.layout == 1) {
      if (d.addr == kUndef) {
        memset(dst, 0, (size_t)d.count * out_size);
        return;
      }
      const uint8_t* src = at(d.addr, (uint64_t)d.count * d.dt.size);
      const int64_t piece = 1 << 20;
      const int64_t np = (d.count + piece - 1) / piece;
      parallel_for(np, nthreads, [&](int64_t i, int) {
        const int64_t e0 = i * piece, n = std::min(piece, d.count - e0);
        convert(d.dt, src + e0 * d.dt.size, static_cast<uint8_t*>(dst) + e0 * out_size, conv, n);
      });
    } else {
      load_chunks(d);
      int64_t expect = 1;
      for (int a = 0; a < d.rank; ++a) expect *= (d.dims[a] + d.chunk[a] - 1) / d.chunk[a];
      if ((int64_t)d.chunks.size() < expect) memset(dst, 0, (size_t)d.count * out_size);
      int64_t zero[8] = {0};
      std::vector<std::vector<uint8_t>> s0((size_t)std::max(nthreads, 1)), s1((size_t)std::max(nthreads, 1));
      parallel_for((int64_t)d.chunks.size(), std::min(nthreads, 64), [&](int64_t i, int t) {
        const ChunkRec& r = d.chunks[(size_t)i];
        place_chunk(d, r, s0[(size_t)t], s1[(size_t)t], static_cast<uint8_t*>(dst), zero, d.dims, conv, out_size);
      });
    }
  }

  // ---- strings -----------------------------------------------------------------------------------
  std::string gheap_string(const uint8_t* ref) {
    const uint32_t len = u32(ref);
    const uint64_t addr = u64(ref + 4);
    const uint32_t index = u32(ref + 12);
    if (addr == 0 || addr == kUndef || len == 0) return std::string();
    const uint8_t* h = at(addr, 16);
    if (memcmp(h, "GCOL", 4) != 0) fail(RSK_H5_ERR_FORMAT, "bad global heap collection");
    const uint64_t csize = u64(h + 8);
    at(addr, csize);
    uint64_t p = 16;
    while (p + 16 <= csize) {
      const uint8_t* o = h + p;
      const uint16_t idx = u16(o);
      const uint64_t osz = u64(o + 8);
      if (idx == 0) break;
      if (p + 16 + osz > csize) fail(RSK_H5_ERR_FORMAT, "global heap object overruns its collection");
      if (idx == index) return std::string(reinterpret_cast<const char*>(o + 16), std::min<uint64_t>(len, osz));
      p += 16 + ((osz + 7) & ~7ull);
    }
    fail(RSK_H5_ERR_FORMAT, "global heap object not found");
  }

  static std::string fixed_string(const uint8_t* p, size_t n) {
    size_t len = 0;
    while (len < n && p[len] != 0) ++len;
    return std::string(reinterpret_cast<const char*>(p), len);
  }

  std::vector<std::string> strings_of(const Dtype& dt, const uint8_t* raw, int64_t count) {
    std::vector<std::string> out;
    for (int64_t i = 0; i < count; ++i) {
      if (dt.cls == RSK_H5_STRING) out.push_back(fixed_string(raw + i * dt.size, (size_t)dt.size));
      else if (dt.cls == RSK_H5_VLEN_STRING) out.push_back(gheap_string(raw + i * 16));
      else fail(RSK_H5_ERR_TYPE, "not a string type");
    }
    return out;
  }

  // ---- attributes --------------------------------------------------------------------------------
  struct Attr {
    Dtype dt;
    int rank = 0;
    int64_t dims[8] = {0};
    int64_t count = 1;
    const uint8_t* data = nullptr;
    size_t data_len = 0;
  };

  std::vector<Msg> attr_messages(const char* path) {
    const uint64_t addr = resolve(path);
    if (addr == kUndef) fail(RSK_H5_ERR_NOTFOUND, std::string("no object named ") + (path ? path : ""));
    std::vector<Msg> out;
    for (const Msg& m : messages(addr))
      if (m.type == 0x000c) out.push_back(m);
    return out;
  }

  bool find_attr(const char* path, const char* name, Attr* a) {
    for (const Msg& m : attr_messages(path)) {
      const uint8_t* p = m.p;
      if (m.len < 8) fail(RSK_H5_ERR_FORMAT, "short attribute message");
      const int ver = p[0];
      if (ver < 1 || ver > 3) fail(RSK_H5_ERR_UNSUPPORTED, "attribute message version");
      if (ver >= 2 && (p[1] & 3)) fail(RSK_H5_ERR_UNSUPPORTED, "shared attribute datatype/dataspace");
      const size_t nsz = u16(p + 2), tsz = u16(p + 4), ssz = u16(p + 6);
      size_t q = ver == 3 ? 9 : 8;
      auto pad = [&](size_t x) { return ver == 1 ? ((x + 7) & ~(size_t)7) : x; };
      if (q + pad(nsz) + pad(tsz) + pad(ssz) > m.len) fail(RSK_H5_ERR_FORMAT, "attribute message overruns");
      const std::string nm = fixed_string(p + q, nsz);
      q += pad(nsz);
      if (nm != name) continue;
      a->dt = datatype(p + q, tsz);
      q += pad(tsz);
      bool null_space;
      dataspace(p + q, ssz, &a->rank, a->dims, &null_space);
      q += pad(ssz);
      a->count = null_space ? 0 : 1;
      for (int i = 0; i < a->rank; ++i) a->count *= a->dims[i];
      a->data = p + q;
      a->data_len = m.len - q;
      const int64_t es = a->dt.cls == RSK_H5_VLEN_STRING ? 16 : a->dt.size;
      if ((int64_t)a->data_len < a->count * es) fail(RSK_H5_ERR_FORMAT, "attribute data too short");
      return true;
    }
    return false;
  }
};

namespace {

int64_t copy_out(const std::string& s, char* buf, int64_t cap) {
  const int64_t need = (int64_t)s.size() + 1;
  if (buf && cap > 0) {
    const int64_t n = std::min<int64_t>(need - 1, cap - 1);
    memcpy(buf, s.data(), (size_t)n);
    buf[n] = 0;
  }
  return need;
}

template <class F>
int64_t guarded(RskH5File* f, F&& body) {
  if (!f) return RSK_H5_ERR_NULL;
  try {
    return body();
  } catch (const H5Err& e) {
    f->err = e.msg;
    return e.code;
  } catch (const std::exception& e) {
    f->err = e.what();
    return RSK_H5_ERR_FORMAT;
  }
}

int default_threads(int n) {
  if (n > 0) return n;
  unsigned hc = std::thread::hardware_concurrency();
  if (hc == 0) hc = 4;
  return (int)std::min(hc, 16u);
}

}  // namespace

extern "C" {

int rsk_h5_open(const char* path, RskH5File** out) {
  if (!path || !out) return RSK_H5_ERR_NULL;
  *out = nullptr;
  RskH5File* f = new RskH5File();
  try {
    f->open(path);
    f->messages(f->root_addr);
  } catch (const H5Err& e) {
    const int code = e.code;
    delete f;
    return code;
  }
  *out = f;
  return RSK_H5_OK;
}

void rsk_h5_close(RskH5File* f) { delete f; }

const char* rsk_h5_last_error(RskH5File* f) { return f ? f->err.c_str() : "null handle"; }

int rsk_h5_kind(RskH5File* f, const char* path) {
  return (int)guarded(f, [&]() -> int64_t {
    const uint64_t addr = f->resolve(path);
    if (addr == kUndef) return RSK_H5_ABSENT;
    if (f->group_links(addr)) return RSK_H5_GROUP;
    for (const Msg& m : f->messages(addr))
      if (m.type == 0x0008) return RSK_H5_DATASET;
    return RSK_H5_GROUP;
  });
}

int64_t rsk_h5_list(RskH5File* f, const char* group, char* buf, int64_t cap) {
  return guarded(f, [&]() -> int64_t {
    const uint64_t addr = f->resolve(group);
    if (addr == kUndef) fail(RSK_H5_ERR_NOTFOUND, std::string("no group named ") + (group ? group : ""));
    const auto* links = f->group_links(addr);
    if (!links) fail(RSK_H5_ERR_TYPE, "not a group");
    std::string s;
    for (const auto& kv : *links) {
      if (!s.empty()) s += '\n';
      s += kv.first;
    }
    return copy_out(s, buf, cap);
  });
}

int rsk_h5_info(RskH5File* f, const char* path, RskH5Info* info) {
  if (!info) return RSK_H5_ERR_NULL;
  return (int)guarded(f, [&]() -> int64_t {
    memset(info, 0, sizeof(*info));
    const int kind = rsk_h5_kind(f, path);
    if (kind < 0) return kind;
    if (kind == RSK_H5_ABSENT) fail(RSK_H5_ERR_NOTFOUND, std::string("no object named ") + (path ? path : ""));
    info->kind = kind;
    if (kind == RSK_H5_GROUP) return RSK_H5_OK;
    Dataset& d = f->dataset(path);
    info->type_class = d.dt.cls;
    info->elem_size = d.dt.cls == RSK_H5_VLEN_STRING ? 16 : d.dt.size;
    info->rank = d.rank;
    for (int a = 0; a < d.rank; ++a) info->dims[a] = d.dims[a];
    info->count = d.count;
    info->layout = d.layout;
    info->nfilters = (int)d.filters.size();
    if (d.layout == 2) {
      f->load_chunks(d);
      for (int a = 0; a < d.rank; ++a) info->chunk[a] = d.chunk[a];
      info->nchunks = (int64_t)d.chunks.size();
    }
    return RSK_H5_OK;
  });
}

int rsk_h5_read(RskH5File* f, const char* path, void* dst, int64_t dst_bytes, int nthreads) {
  return (int)guarded(f, [&]() -> int64_t {
    Dataset& d = f->dataset(path);
    const int64_t es = d.dt.cls == RSK_H5_VLEN_STRING ? 16 : d.dt.size;
    if (d.count > 0 && !dst) return RSK_H5_ERR_NULL;
    if (dst_bytes < d.count * es) fail(RSK_H5_ERR_SIZE, "destination too small");
    f->read_host(d, dst, kRaw, default_threads(nthreads));
    return RSK_H5_OK;
  });
}

int rsk_h5_read_as(RskH5File* f, const char* path, int as_type, void* dst, int64_t count, int nthreads) {
  return (int)guarded(f, [&]() -> int64_t {
    if (as_type != RSK_H5_AS_F64 && as_type != RSK_H5_AS_I64) fail(RSK_H5_ERR_TYPE, "bad destination type");
    Dataset& d = f->dataset(path);
    if (count != d.count) fail(RSK_H5_ERR_SIZE, "element count differs from the dataset's");
    if (d.count > 0 && !dst) return RSK_H5_ERR_NULL;
    f->read_host(d, dst, as_type == RSK_H5_AS_F64 ? kF64 : kI64, default_threads(nthreads));
    return RSK_H5_OK;
  });
}

int64_t rsk_h5_read_strings(RskH5File* f, const char* path, char* buf, int64_t cap) {
  return guarded(f, [&]() -> int64_t {
    Dataset& d = f->dataset(path);
    if (d.dt.cls != RSK_H5_STRING && d.dt.cls != RSK_H5_VLEN_STRING) fail(RSK_H5_ERR_TYPE, "not a string dataset");
    const int64_t es = d.dt.cls == RSK_H5_VLEN_STRING ? 16 : d.dt.size;
    std::vector<uint8_t> raw((size_t)(d.count * es));
    f->read_host(d, raw.data(), kRaw, 1);
    std::string s;
    for (const std::string& x : f->strings_of(d.dt, raw.data(), d.count)) {
      s += x;
      s += '\0';
    }
    const int64_t need = (int64_t)s.size();
    if (buf && cap > 0) memcpy(buf, s.data(), (size_t)std::min<int64_t>(need, cap));
    return need;
  });
}

int64_t rsk_h5_attr_string(RskH5File* f, const char* path, const char* attr, char* buf, int64_t cap) {
  if (!attr) return RSK_H5_ERR_NULL;
  return guarded(f, [&]() -> int64_t {
    RskH5File::Attr a;
    if (!f->find_attr(path, attr, &a)) return RSK_H5_ERR_NOTFOUND;
    if (a.count < 1) return copy_out("", buf, cap);
    return copy_out(f->strings_of(a.dt, a.data, 1)[0], buf, cap);
  });
}

int64_t rsk_h5_attr_f64(RskH5File* f, const char* path, const char* attr, double* dst, int64_t cap) {
  if (!attr) return RSK_H5_ERR_NULL;
  return guarded(f, [&]() -> int64_t {
    RskH5File::Attr a;
    if (!f->find_attr(path, attr, &a)) return RSK_H5_ERR_NOTFOUND;
    if (a.count > cap) fail(RSK_H5_ERR_SIZE, "destination too small");
    if (a.count > 0 && !dst) return RSK_H5_ERR_NULL;
    convert(a.dt, a.data, dst, kF64, a.count);
    return a.count;
  });
}

// File -> HBM.  The dataset is viewed as (A, B, C) = (dims[0], dims[1], dims[2]) (leading 1s for lower
// ranks).  A slab is one chunk-row along A times a run of whole chunk-rows along B, so every chunk
// belongs to exactly one slab (nothing is inflated twice) and a slab lands in the device array with a
// single pitched copy.  Three pinned slabs rotate: while one is in flight the next is being filled.
int rsk_h5_read_device(RskH5File* f, const char* path, int as_type, void* dst_device, int64_t count, int nthreads,
                       int64_t slab_bytes, void* stream) {
  constexpr int kRing = RskH5File::kRing;
  if (!f) return RSK_H5_ERR_NULL;
  bool used[kRing] = {false, false, false};
  cudaStream_t st = (cudaStream_t)stream;
  const int64_t rc = guarded(f, [&]() -> int64_t {
    if (as_type != RSK_H5_AS_F64 && as_type != RSK_H5_AS_I64) fail(RSK_H5_ERR_TYPE, "bad destination type");
    const int conv = as_type == RSK_H5_AS_F64 ? kF64 : kI64;
    Dataset& d = f->dataset(path);
    if (count != d.count) fail(RSK_H5_ERR_SIZE, "element count differs from the dataset's");
    if (d.count == 0) return RSK_H5_OK;
    if (!dst_device) return RSK_H5_ERR_NULL;
    if (d.dt.cls == RSK_H5_STRING || d.dt.cls == RSK_H5_VLEN_STRING) fail(RSK_H5_ERR_TYPE, "string dataset");
    const int nt = d.count < (1 << 17) ? 1 : default_threads(nthreads);  // small columns: no helpers
    Pool pool(nt);
    if (slab_bytes <= 0) slab_bytes = 64ll << 20;
    uint8_t** stage = f->stage;
    cudaEvent_t* ev = f->stage_ev;
    auto cu = [&](cudaError_t e) {
      if (e != cudaSuccess) throw H5Err{(int)e, cudaGetErrorString(e)};
    };
    uint8_t* dst = static_cast<uint8_t*>(dst_device);

    if (d.layout != 2 || d.rank > 3) {
      // compact / contiguous (or an unusual rank): flat slabs
      std::vector<uint8_t> whole;
      const uint8_t* src = nullptr;
      Dtype sdt = d.dt;
      if (d.layout == 1 && d.addr != kUndef) {
        src = f->at(d.addr, (uint64_t)d.count * d.dt.size);
      } else {
        whole.resize((size_t)d.count * 8);
        f->read_host(d, whole.data(), conv, nt);
        src = whole.data();
        sdt.cls = conv == kF64 ? RSK_H5_FLOAT : RSK_H5_INT;
        sdt.size = 8;
      }
      const int64_t per = std::max<int64_t>(slab_bytes / 8, 1);
      const int64_t cap_elems = std::min(per, d.count);
      f->ensure_stage((size_t)cap_elems * 8);
      int slot = 0;
      for (int64_t e0 = 0; e0 < d.count; e0 += cap_elems, slot = (slot + 1) % kRing) {
        const int64_t n = std::min(cap_elems, d.count - e0);
        if (used[slot]) cu(cudaEventSynchronize(ev[slot]));
        const int64_t piece = 1 << 18;
        const int64_t np = (n + piece - 1) / piece;
        uint8_t* sbuf = stage[slot];
        const bool direct = whole.empty() && sdt.size == 8 &&
                            ((conv == kF64 && sdt.cls == RSK_H5_FLOAT) || (conv == kI64 && sdt.cls == RSK_H5_INT));
        pool.run(np, [&](int64_t i, int) {
          const int64_t a = i * piece, m = std::min(piece, n - a);
          if (direct) {  // pread: no page faults on the mapping
            uint8_t* to = sbuf + a * 8;
            int64_t left = m * 8, fo = (int64_t)(f->base + d.addr) + (e0 + a) * 8;
            while (left > 0) {
              const ssize_t got = pread(f->fd, to, (size_t)left, (off_t)fo);
              if (got <= 0) fail(RSK_H5_ERR_OPEN, "read failed");
              to += got;
              fo += got;
              left -= got;
            }
          } else {
            convert(sdt, src + (e0 + a) * sdt.size, sbuf + a * 8, conv, m);
          }
        });
        cu(cudaMemcpyAsync(dst + e0 * 8, sbuf, (size_t)n * 8, cudaMemcpyHostToDevice, st));
        cu(cudaEventRecord(ev[slot], st));
        used[slot] = true;
      }
      return RSK_H5_OK;
    }

    f->load_chunks(d);
    // (A, B, C) view
    int64_t D[3] = {1, 1, 1}, Cc[3] = {1, 1, 1};
    for (int a = 0; a < d.rank; ++a) {
      D[3 - d.rank + a] = d.dims[a];
      Cc[3 - d.rank + a] = d.chunk[a];
    }
    const int64_t nA = (D[0] + Cc[0] - 1) / Cc[0], nB = (D[1] + Cc[1] - 1) / Cc[1], nC = (D[2] + Cc[2] - 1) / Cc[2];
    std::vector<std::vector<int32_t>> bucket((size_t)(nA * nB));
    for (size_t i = 0; i < d.chunks.size(); ++i) {
      uint64_t o[3] = {0, 0, 0};
      for (int a = 0; a < d.rank; ++a) o[3 - d.rank + a] = d.chunks[i].off[a];
      bucket[(size_t)((o[0] / Cc[0]) * nB + o[1] / Cc[1])].push_back((int32_t)i);
    }
    const int64_t row_bytes = Cc[0] * Cc[1] * D[2] * 8;  // one chunk-row along B, all of C, Cc[0] planes
    const int64_t rows_per_slab = std::max<int64_t>(1, std::min(nB, slab_bytes / std::max<int64_t>(row_bytes, 1)));
    const size_t stage_bytes = (size_t)(rows_per_slab * row_bytes);
    f->ensure_stage(stage_bytes);
    std::vector<std::vector<uint8_t>> s0((size_t)nt), s1((size_t)nt);
    int slot = 0;
    for (int64_t ia = 0; ia < nA; ++ia) {
      const int64_t a0 = ia * Cc[0], ha = std::min(Cc[0], D[0] - a0);
      for (int64_t ib0 = 0; ib0 < nB; ib0 += rows_per_slab, slot = (slot + 1) % kRing) {
        const int64_t ib1 = std::min(nB, ib0 + rows_per_slab);
        const int64_t b0 = ib0 * Cc[1], nb = std::min(ib1 * Cc[1], D[1]) - b0;
        if (used[slot]) cu(cudaEventSynchronize(ev[slot]));
        uint8_t* sbuf = stage[slot];
        std::vector<int32_t> ids;
        for (int64_t ib = ib0; ib < ib1; ++ib)
          for (int32_t id : bucket[(size_t)(ia * nB + ib)]) ids.push_back(id);
        if ((int64_t)ids.size() < (ib1 - ib0) * nC) memset(sbuf, 0, (size_t)(ha * nb * D[2] * 8));
        // staging layout [ha][nb][C]: expressed in the dataset's own rank for scatter_chunk
        int64_t lo3[3] = {a0, b0, 0}, ext3[3] = {ha, nb, D[2]};
        int64_t lo[8], ext[8];
        for (int a = 0; a < d.rank; ++a) {
          lo[a] = lo3[3 - d.rank + a];
          ext[a] = ext3[3 - d.rank + a];
        }
        pool.run((int64_t)ids.size(), [&](int64_t i, int t) {
          const ChunkRec& r = d.chunks[(size_t)ids[(size_t)i]];
          f->place_chunk(d, r, s0[(size_t)t], s1[(size_t)t], sbuf, lo, ext, conv, 8);
        });
        cu(cudaMemcpy2DAsync(dst + ((a0 * D[1] + b0) * D[2]) * 8, (size_t)(D[1] * D[2] * 8), sbuf,
                             (size_t)(nb * D[2] * 8), (size_t)(nb * D[2] * 8), (size_t)ha, cudaMemcpyHostToDevice, st));
        cu(cudaEventRecord(ev[slot], st));
        used[slot] = true;
      }
    }
    return RSK_H5_OK;
  });
  for (int i = 0; i < kRing; ++i)  // the staging memory is reused by the next call: wait for its copies
    if (used[i]) cudaEventSynchronize(f->stage_ev[i]);
  return (int)rc;
}

const char* rsk_h5_strerror(int code) {
  switch (code) {
    case RSK_H5_OK: return "success";
    case RSK_H5_ERR_NULL: return "a required pointer is NULL";
    case RSK_H5_ERR_OPEN: return "cannot open, map or create the file";
    case RSK_H5_ERR_FORMAT: return "not an HDF5 file, or the file is structurally corrupt";
    case RSK_H5_ERR_UNSUPPORTED: return "HDF5 feature outside the supported SkyH5 subset";
    case RSK_H5_ERR_NOTFOUND: return "no such object or attribute";
    case RSK_H5_ERR_TYPE: return "object or datatype is not of the requested kind";
    case RSK_H5_ERR_SIZE: return "destination too small or element count mismatch";
    case RSK_H5_ERR_FILTER: return "a chunk filter failed";
    case RSK_H5_ERR_STATE: return "writer state error";
    default: return code > 0 ? cudaGetErrorString((cudaError_t)code) : "unknown error";
  }
}

}  // extern "C"
