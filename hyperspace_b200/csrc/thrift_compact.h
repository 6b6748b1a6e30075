// thrift_compact.h -- minimal Thrift Compact Protocol reader (host + device) and writer (host).
// Parquet's footer (FileMetaData) and page headers are Thrift-compact structs; parquet-mr/Spark and pyarrow
// both emit this protocol.  The reader is header-only and __host__ __device__ so the GPU page walker and the host
// footer parser share it.
#pragma once
#include <cstdint>
#include <string>
#include <vector>

#ifdef __CUDACC__
#define HS_HD __host__ __device__ __forceinline__
#else
#define HS_HD inline
#endif

namespace hs {
namespace thrift {

enum CType : uint8_t {
  T_STOP = 0, T_TRUE = 1, T_FALSE = 2, T_BYTE = 3, T_I16 = 4, T_I32 = 5, T_I64 = 6, T_DOUBLE = 7,
  T_BINARY = 8, T_LIST = 9, T_SET = 10, T_MAP = 11, T_STRUCT = 12
};

struct Reader {
  const uint8_t* p;
  const uint8_t* end;
  bool bad;

  HS_HD Reader(const uint8_t* b, const uint8_t* e) : p(b), end(e), bad(false) {}

  HS_HD uint8_t byte() {
    if (p >= end) {
      bad = true;
      return 0;
    }
    return *p++;
  }
  HS_HD uint64_t varint() {
    uint64_t v = 0;
    int shift = 0;
    for (int i = 0; i < 10; i++) {
      uint8_t b = byte();
      v |= (uint64_t)(b & 0x7f) << shift;
      if (!(b & 0x80)) return v;
      shift += 7;
    }
    bad = true;
    return v;
  }
  HS_HD int64_t zigzag() {
    uint64_t v = varint();
    return (int64_t)(v >> 1) ^ -(int64_t)(v & 1);
  }
  // Reads a field header; returns the compact type (T_STOP at struct end) and updates field id.
  HS_HD uint8_t field(int16_t& fid) {
    uint8_t h = byte();
    if (h == 0) return T_STOP;
    uint8_t type = h & 0x0f;
    uint8_t delta = h >> 4;
    if (delta) fid = (int16_t)(fid + delta);
    else fid = (int16_t)zigzag();
    return type;
  }
  HS_HD void skip_bytes(uint64_t n) {
    if ((uint64_t)(end - p) < n) {
      bad = true;
      p = end;
    } else {
      p += n;
    }
  }
  // list header: element type + count
  HS_HD uint32_t list(uint8_t& etype) {
    uint8_t h = byte();
    etype = h & 0x0f;
    uint32_t n = h >> 4;
    if (n == 15) n = (uint32_t)varint();
    return n;
  }
  // Skip one value of compact type `type` (iterative, explicit stack; nesting in Parquet metadata is shallow).
  HS_HD void skip(uint8_t type) {
    // stack entries: kind 0 = struct (read fields until STOP), kind 1 = list with `remaining` elements of etype
    struct Frame {
      uint8_t kind, etype;
      uint32_t remaining;
    };
    Frame st[12];
    int sp = 0;
    bool first = true;
    uint8_t cur = type;
    while (!bad) {
      if (!first) {
        if (sp == 0) return;
        Frame& f = st[sp - 1];
        if (f.kind == 0) {
          int16_t fid = 0;
          cur = field(fid);
          if (cur == T_STOP) {
            sp--;
            continue;
          }
        } else {
          if (f.remaining == 0) {
            sp--;
            continue;
          }
          f.remaining--;
          cur = f.etype;
        }
      }
      first = false;
      switch (cur) {
        case T_TRUE:
        case T_FALSE:
          // in a list a bool occupies one byte; as a field the value is in the header
          if (sp > 0 && st[sp - 1].kind == 1) byte();
          break;
        case T_BYTE: byte(); break;
        case T_I16:
        case T_I32:
        case T_I64: varint(); break;
        case T_DOUBLE: skip_bytes(8); break;
        case T_BINARY: skip_bytes(varint()); break;
        case T_LIST:
        case T_SET: {
          uint8_t et;
          uint32_t n = list(et);
          if (sp >= 12) { bad = true; return; }
          st[sp++] = Frame{1, et, n};
          break;
        }
        case T_MAP:  // parquet.thrift declares no maps
          bad = true;
          return;
        case T_STRUCT:
          if (sp >= 12) { bad = true; return; }
          st[sp++] = Frame{0, 0, 0};
          break;
        default: bad = true; return;
      }
      if (sp == 0) return;
    }
  }
};

// ---- host-side writer -----------------------------------------------------------------------------------
class Writer {
 public:
  std::vector<uint8_t> buf;

  void byte(uint8_t b) { buf.push_back(b); }
  void varint(uint64_t v) {
    while (v >= 0x80) {
      buf.push_back((uint8_t)(v | 0x80));
      v >>= 7;
    }
    buf.push_back((uint8_t)v);
  }
  void zigzag(int64_t v) { varint(((uint64_t)v << 1) ^ (uint64_t)(v >> 63)); }

  void struct_begin() {
    last_.push_back(0);
  }
  void struct_end() {
    byte(0);
    last_.pop_back();
  }
  void field(int16_t fid, uint8_t type) {
    int16_t& last = last_.back();
    int d = fid - last;
    if (d > 0 && d <= 15) byte((uint8_t)((d << 4) | type));
    else {
      byte(type);
      zigzag(fid);
    }
    last = fid;
  }
  void f_i32(int16_t fid, int32_t v) {
    field(fid, T_I32);
    zigzag(v);
  }
  void f_i64(int16_t fid, int64_t v) {
    field(fid, T_I64);
    zigzag(v);
  }
  void f_bool(int16_t fid, bool v) { field(fid, v ? T_TRUE : T_FALSE); }
  void f_string(int16_t fid, const std::string& s) {
    field(fid, T_BINARY);
    varint(s.size());
    buf.insert(buf.end(), s.begin(), s.end());
  }
  // returns the offset of the payload inside buf (so fixed-width values can be patched in later)
  size_t f_binary(int16_t fid, const void* d, size_t n) {
    field(fid, T_BINARY);
    varint(n);
    const uint8_t* b = (const uint8_t*)d;
    const size_t off = buf.size();
    buf.insert(buf.end(), b, b + n);
    return off;
  }
  void f_struct_begin(int16_t fid) {
    field(fid, T_STRUCT);
    struct_begin();
  }
  void f_list_begin(int16_t fid, uint8_t etype, uint32_t n) {
    field(fid, T_LIST);
    list_header(etype, n);
  }
  void list_header(uint8_t etype, uint32_t n) {
    if (n < 15) byte((uint8_t)((n << 4) | etype));
    else {
      byte((uint8_t)(0xf0 | etype));
      varint(n);
    }
  }
  // list elements that are structs are written with struct_begin()/struct_end(); i32 elements with zigzag()
  void string_elem(const std::string& s) {
    varint(s.size());
    buf.insert(buf.end(), s.begin(), s.end());
  }

 private:
  std::vector<int16_t> last_;
};

}  // namespace thrift
}  // namespace hs
