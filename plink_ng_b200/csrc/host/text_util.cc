#include "text_util.h"

#include <algorithm>

#include <dlfcn.h>

#include <cfloat>
#include <cmath>
#include <cstring>
#include <fstream>
#include <sstream>

namespace pl2host {

namespace {

// Round-half-to-even with the reference's tolerance band: values within 5e-9 of a .5 tie are
// treated as exact ties (kBankerRound8, plink2_string.cc:2232-2237).
inline uint32_t RoundTol(double x) {
  uint32_t r = static_cast<uint32_t>(static_cast<int32_t>(x));
  const double bump = (r & 1) ? 0.500000005 : 0.499999995;
  r += static_cast<uint32_t>(static_cast<int32_t>((x - static_cast<double>(r)) + bump));
  return r;
}

// digits of `v` zero-padded to `width`, trailing zeros removed (at least one digit kept when
// keep_one); returns new end
inline char* PutFrac(uint32_t v, int width, char* p) {
  char tmp[12];
  for (int i = width - 1; i >= 0; --i) {
    tmp[i] = static_cast<char>('0' + v % 10);
    v /= 10;
  }
  int n = width;
  while (n > 0 && tmp[n - 1] == '0') --n;
  memcpy(p, tmp, n);
  return p + n;
}

inline char* PutExp(uint32_t e, char sign, char* p) {
  *p++ = 'e';
  *p++ = sign;
  if (e >= 100) {
    *p++ = static_cast<char>('0' + e / 100);
    e %= 100;
  }
  *p++ = static_cast<char>('0' + e / 10);
  *p++ = static_cast<char>('0' + e % 10);
  return p;
}

// one leading digit + up to 5 decimals from x in [0.99999949999999, 9.9999949999999)
inline char* PutMantissa(double x, char* p) {
  const uint32_t v = RoundTol(x * 100000);
  *p++ = static_cast<char>('0' + v / 100000);
  const uint32_t rem = v % 100000;
  if (rem) {
    *p++ = '.';
    p = PutFrac(rem, 5, p);
  }
  return p;
}

}  // namespace

char* u32toa(uint32_t x, char* buf) {
  char tmp[12];
  int n = 0;
  do {
    tmp[n++] = static_cast<char>('0' + x % 10);
    x /= 10;
  } while (x);
  while (n) *buf++ = tmp[--n];
  return buf;
}

char* i32toa(int32_t x, char* buf) {
  if (x < 0) {
    *buf++ = '-';
    return u32toa(static_cast<uint32_t>(-static_cast<int64_t>(x)), buf);
  }
  return u32toa(static_cast<uint32_t>(x), buf);
}

char* dtoa_g(double x, char* p) {
  if (x != x) {
    memcpy(p, "nan", 3);
    return p + 3;
  }
  if (x < 0) {
    *p++ = '-';
    x = -x;
  }
  // Exponential notation (|x| < 1e-4 or >= 1e6, after rounding to 6 digits): bring x into [1, 10) by the binary
  // decomposition of its decimal exponent, largest power first.  The comparison bounds and the multiply order
  // are part of the output contract (a different scaling order rounds differently in the last printed digit).
  struct Pow10Step {
    uint32_t e;
    double below, up;    // x < below  -> x *= up   (small side)
    double from, down;   // x >= from  -> x *= down (large side)
  };
  static const Pow10Step kSteps[9] = {
      {256, 9.9999949999999e-256, 1.0e256, 9.9999949999999e255, 1.0e-256}, {128, 9.9999949999999e-128, 1.0e128, 9.9999949999999e127, 1.0e-128},
      {64, 9.9999949999999e-64, 1.0e64, 9.9999949999999e63, 1.0e-64},      {32, 9.9999949999999e-32, 1.0e32, 9.9999949999999e31, 1.0e-32},
      {16, 9.9999949999999e-16, 1.0e16, 9.9999949999999e15, 1.0e-16},      {8, 9.9999949999999e-8, 100000000, 9.9999949999999e7, 1.0e-8},
      {4, 9.9999949999999e-4, 10000, 9.9999949999999e3, 1.0e-4},           {2, 9.9999949999999e-2, 100, 9.9999949999999e1, 1.0e-2},
      {1, 9.9999949999999e-1, 10, 9.9999949999999e0, 1.0e-1}};
  if (x < 9.9999949999999e-5) {
    if (x == 0.0) {
      *p++ = '0';
      return p;
    }
    uint32_t xp10 = 0;
    for (const Pow10Step& st : kSteps) {
      if (x < st.below) {
        x *= st.up;
        xp10 += st.e;
      }
    }
    return PutExp(xp10, '-', PutMantissa(x, p));
  }
  if (x >= 999999.49999999) {
    if (x > DBL_MAX) {
      memcpy(p, "inf", 3);
      return p + 3;
    }
    uint32_t xp10 = 0;
    for (const Pow10Step& st : kSteps) {
      if (x >= st.from) {
        x *= st.down;
        xp10 += st.e;
      }
    }
    return PutExp(xp10, '+', PutMantissa(x, p));
  }
  if (x >= 0.99999949999999) {
    // k integer digits, 6 - k decimals
    static const double kUpper[5] = {9.9999949999999, 99.999949999999, 999.99949999999, 9999.9949999999, 99999.949999999};
    static const double kScale[6] = {100000, 10000, 1000, 100, 10, 1};
    static const uint32_t kDiv[6] = {100000, 10000, 1000, 100, 10, 1};
    int k = 0;
    while (k < 5 && !(x < kUpper[k])) ++k;  // k + 1 integer digits
    const uint32_t v = (k == 5) ? RoundTol(x) : RoundTol(x * kScale[k]);
    p = u32toa(v / kDiv[k], p);
    const uint32_t rem = v % kDiv[k];
    if (rem) {
      *p++ = '.';
      p = PutFrac(rem, 5 - k, p);
    }
    return p;
  }
  // 0.0001 <= x < 1
  *p++ = '0';
  *p++ = '.';
  if (x < 9.9999949999999e-3) {
    x *= 100;
    *p++ = '0';
    *p++ = '0';
  }
  if (x < 9.9999949999999e-2) {
    x *= 10;
    *p++ = '0';
  }
  char* q = PutFrac(RoundTol(x * 1000000), 6, p);
  if (q == p) *q++ = '0';
  return q;
}

// dtoa_g_p8: the 8-significant-digit sibling (reference contract: 2.0/include/plink2_string.cc:2641-2790; used
// by --make-grm-sparse, :5073).  Same shape as dtoa_g with the bounds moved to 9.9999999499999e-k, the
// mantissa carried as round(x * 10^(8 - integer digits)) and ties decided inside a +-5e-7 band (kBankerRound6).
namespace {
inline uint32_t RoundTol6(double x) {
  uint32_t r = static_cast<uint32_t>(static_cast<int32_t>(x));
  const double bump = (r & 1) ? 0.5000005 : 0.4999995;
  r += static_cast<uint32_t>(static_cast<int32_t>((x - static_cast<double>(r)) + bump));
  return r;
}
inline char* PutMantissa8(double x, char* p) {  // one leading digit + up to 7 decimals
  const uint32_t v = RoundTol6(x * 10000000);
  *p++ = static_cast<char>('0' + v / 10000000);
  const uint32_t rem = v % 10000000;
  if (rem) {
    *p++ = '.';
    p = PutFrac(rem, 7, p);
  }
  return p;
}
}  // namespace

char* dtoa_g_p8(double x, char* p) {
  if (x != x) {
    memcpy(p, "nan", 3);
    return p + 3;
  }
  char* const start = p;
  if (x < 0) {
    *p++ = '-';
    x = -x;
  }
  // same exponent decomposition as dtoa_g, with the 8-digit rounding bounds
  struct Pow10Step {
    uint32_t e;
    double below, up, from, down;
  };
  static const Pow10Step kSteps[9] = {
      {256, 9.9999999499999e-256, 1.0e256, 9.9999999499999e255, 1.0e-256}, {128, 9.9999999499999e-128, 1.0e128, 9.9999999499999e127, 1.0e-128},
      {64, 9.9999999499999e-64, 1.0e64, 9.9999999499999e63, 1.0e-64},      {32, 9.9999999499999e-32, 1.0e32, 9.9999999499999e31, 1.0e-32},
      {16, 9.9999999499999e-16, 1.0e16, 9.9999999499999e15, 1.0e-16},      {8, 9.9999999499999e-8, 100000000, 9.9999999499999e7, 1.0e-8},
      {4, 9.9999999499999e-4, 10000, 9.9999999499999e3, 1.0e-4},           {2, 9.9999999499999e-2, 100, 9.9999999499999e1, 1.0e-2},
      {1, 9.9999999499999e-1, 10, 9.9999999499999e0, 1.0e-1}};
  if (x < 9.9999999499999e-5) {
    if (x == 0.0) {
      *start = '0';  // the sign of -0 is dropped, as in the reference
      return start + 1;
    }
    uint32_t xp10 = 0;
    for (const Pow10Step& st : kSteps) {
      if (x < st.below) {
        x *= st.up;
        xp10 += st.e;
      }
    }
    return PutExp(xp10, '-', PutMantissa8(x, p));
  }
  if (x >= 99999999.499999) {
    if (x > DBL_MAX) {  // the reference prints " inf" (with the blank) for +infinity
      memcpy(start, (p == start) ? " inf" : "-inf", 4);
      return start + 4;
    }
    uint32_t xp10 = 0;
    for (const Pow10Step& st : kSteps) {
      if (x >= st.from) {
        x *= st.down;
        xp10 += st.e;
      }
    }
    return PutExp(xp10, '+', PutMantissa8(x, p));
  }
  if (x >= 0.99999999499999) {
    // k + 1 integer digits, 7 - k decimals
    static const double kUpper[7] = {9.9999999499999, 99.999999499999, 999.99999499999, 9999.9999499999, 99999.999499999, 999999.99499999, 9999999.9499999};
    static const double kScale[8] = {10000000, 1000000, 100000, 10000, 1000, 100, 10, 1};
    static const uint32_t kDiv[8] = {10000000, 1000000, 100000, 10000, 1000, 100, 10, 1};
    int k = 0;
    while (k < 7 && !(x < kUpper[k])) ++k;
    const uint32_t v = (k == 7) ? RoundTol6(x) : RoundTol6(x * kScale[k]);
    p = u32toa(v / kDiv[k], p);
    const uint32_t rem = v % kDiv[k];
    if (rem) {
      *p++ = '.';
      p = PutFrac(rem, 7 - k, p);
    }
    return p;
  }
  // 0.0001 <= x < 1
  *p++ = '0';
  *p++ = '.';
  if (x < 9.9999999499999e-3) {
    x *= 100;
    *p++ = '0';
    *p++ = '0';
  }
  if (x < 9.9999999499999e-2) {
    x *= 10;
    *p++ = '0';
  }
  char* q = PutFrac(RoundTol6(x * 100000000), 8, p);
  if (q == p) *q++ = '0';
  return q;
}

namespace {
// the three libzstd entry points a one-shot frame needs (stable ABI since zstd 1.0)
struct ZstdApi {
  size_t (*compress_bound)(size_t) = nullptr;
  size_t (*compress)(void*, size_t, const void*, size_t, int) = nullptr;
  unsigned (*is_error)(size_t) = nullptr;
  bool ok = false;
};
const ZstdApi& Zstd() {
  static const ZstdApi api = [] {
    ZstdApi a;
    void* h = dlopen("libzstd.so.1", RTLD_NOW | RTLD_GLOBAL);
    if (!h) h = dlopen("libzstd.so", RTLD_NOW | RTLD_GLOBAL);
    if (h) {
      a.compress_bound = reinterpret_cast<size_t (*)(size_t)>(dlsym(h, "ZSTD_compressBound"));
      a.compress = reinterpret_cast<size_t (*)(void*, size_t, const void*, size_t, int)>(dlsym(h, "ZSTD_compress"));
      a.is_error = reinterpret_cast<unsigned (*)(size_t)>(dlsym(h, "ZSTD_isError"));
      a.ok = a.compress_bound && a.compress && a.is_error;
    }
    return a;
  }();
  return api;
}
}  // namespace

bool OutFile::Open(const std::string& path, bool zst) {
  zst_ = zst;
  pos_ = 0;
  if (zst_ && !Zstd().ok) {
    ok_ = false;  // libzstd is not on this system: refuse rather than write an uncompressed file under a .zst name
    return false;
  }
  f_ = fopen(path.c_str(), "wb");
  buf_.resize(1 << 20);
  ok_ = f_ != nullptr;
  return ok_;
}

void OutFile::Flush() {
  if (f_ && pos_) {
    if (zst_) {
      const ZstdApi& z = Zstd();
      zbuf_.resize(z.compress_bound(pos_));
      const size_t n = z.compress(zbuf_.data(), zbuf_.size(), buf_.data(), pos_, 3);
      if (z.is_error(n) || fwrite(zbuf_.data(), 1, n, f_) != n) ok_ = false;
    } else if (fwrite(buf_.data(), 1, pos_, f_) != pos_) {
      ok_ = false;
    }
  }
  pos_ = 0;
}

bool OutFile::Close() {
  if (f_) {
    Flush();
    if (fclose(f_)) ok_ = false;
    f_ = nullptr;
  }
  return ok_;
}

char* OutFile::Reserve(size_t n) {
  if (pos_ + n > buf_.size()) {
    Flush();
    if (n > buf_.size()) buf_.resize(n);
  }
  return buf_.data() + pos_;
}

void OutFile::Write(const void* p, size_t n) {
  if (n >= buf_.size() / 2 && !zst_) {
    Flush();
    if (f_ && fwrite(p, 1, n, f_) != n) ok_ = false;
    return;
  }
  if (zst_) {  // large writes go through the frame buffer in pieces
    const char* src = static_cast<const char*>(p);
    while (n) {
      const size_t room = buf_.size() - pos_;
      if (!room) {
        Flush();
        continue;
      }
      const size_t take = n < room ? n : room;
      memcpy(buf_.data() + pos_, src, take);
      pos_ += take;
      src += take;
      n -= take;
    }
    return;
  }
  char* d = Reserve(n);
  memcpy(d, p, n);
  pos_ += n;
}

void OutFile::Puts(const char* s) { Write(s, strlen(s)); }

std::vector<std::string> SplitWs(const std::string& line) {
  std::vector<std::string> out;
  size_t i = 0, n = line.size();
  while (i < n) {
    while (i < n && (line[i] == ' ' || line[i] == '\t' || line[i] == '\r' || line[i] == '\n')) ++i;
    if (i >= n) break;
    size_t j = i;
    while (j < n && !(line[j] == ' ' || line[j] == '\t' || line[j] == '\r' || line[j] == '\n')) ++j;
    out.emplace_back(line, i, j - i);
    i = j;
  }
  return out;
}

namespace {
// Streaming decompressors resolved at run time (no link-time dependency, like the 'zs' writer): zstd for .zst input,
// zlib for gzip / bgzf input.  The reference's TextStream reads all three transparently for every text input
// (2.0/include/plink2_text.cc: plain / gzip / bgzf / zstd by magic number), so --pfile ... vzs, a .kin0.zst written by a
// previous run, or a gzipped score file just work.
struct ZInBuf {
  const void* src;
  size_t size, pos;
};
struct ZOutBuf {
  void* dst;
  size_t size, pos;
};
bool ZstdDecompressAll(const std::string& in, std::string* out, std::string* err) {
  void* h = dlopen("libzstd.so.1", RTLD_NOW | RTLD_GLOBAL);
  if (!h) h = dlopen("libzstd.so", RTLD_NOW | RTLD_GLOBAL);
  auto create = h ? reinterpret_cast<void* (*)()>(dlsym(h, "ZSTD_createDStream")) : nullptr;
  auto release = h ? reinterpret_cast<size_t (*)(void*)>(dlsym(h, "ZSTD_freeDStream")) : nullptr;
  auto step = h ? reinterpret_cast<size_t (*)(void*, ZOutBuf*, ZInBuf*)>(dlsym(h, "ZSTD_decompressStream")) : nullptr;
  auto is_error = h ? reinterpret_cast<unsigned (*)(size_t)>(dlsym(h, "ZSTD_isError")) : nullptr;
  if (!create || !release || !step || !is_error) {
    *err = "libzstd is not available to read a Zstandard-compressed input file.";
    return false;
  }
  void* ds = create();
  std::vector<char> chunk(1 << 20);
  ZInBuf ib{in.data(), in.size(), 0};
  size_t last = 0;
  while (ib.pos < ib.size) {  // concatenated frames: a return value of 0 ends a frame, the next call starts the next
    ZOutBuf ob{chunk.data(), chunk.size(), 0};
    last = step(ds, &ob, &ib);
    if (is_error(last)) {
      release(ds);
      *err = "Malformed Zstandard stream.";
      return false;
    }
    out->append(chunk.data(), ob.pos);
  }
  for (;;) {  // drain what is still buffered inside the decoder
    ZOutBuf ob{chunk.data(), chunk.size(), 0};
    const size_t rc = step(ds, &ob, &ib);
    if (is_error(rc)) break;
    out->append(chunk.data(), ob.pos);
    if (!ob.pos) break;
  }
  release(ds);
  if (last != 0) {
    *err = "Truncated Zstandard stream.";
    return false;
  }
  return true;
}

// zlib's z_stream, as laid out by every zlib 1.2+ build on LP64
struct ZStream {
  const unsigned char* next_in;
  unsigned avail_in;
  unsigned long total_in;
  unsigned char* next_out;
  unsigned avail_out;
  unsigned long total_out;
  const char* msg;
  void* state;
  void* zalloc;
  void* zfree;
  void* opaque;
  int data_type;
  unsigned long adler, reserved;
};
bool GzipDecompressAll(const std::string& in, std::string* out, std::string* err) {
  void* h = dlopen("libz.so.1", RTLD_NOW | RTLD_GLOBAL);
  if (!h) h = dlopen("libz.so", RTLD_NOW | RTLD_GLOBAL);
  auto init2 = h ? reinterpret_cast<int (*)(ZStream*, int, const char*, int)>(dlsym(h, "inflateInit2_")) : nullptr;
  auto inflate = h ? reinterpret_cast<int (*)(ZStream*, int)>(dlsym(h, "inflate")) : nullptr;
  auto reset = h ? reinterpret_cast<int (*)(ZStream*)>(dlsym(h, "inflateReset")) : nullptr;
  auto end = h ? reinterpret_cast<int (*)(ZStream*)>(dlsym(h, "inflateEnd")) : nullptr;
  auto version = h ? reinterpret_cast<const char* (*)()>(dlsym(h, "zlibVersion")) : nullptr;
  if (!init2 || !inflate || !reset || !end || !version) {
    *err = "zlib is not available to read a gzip-compressed input file.";
    return false;
  }
  ZStream zs;
  memset(&zs, 0, sizeof(zs));
  if (init2(&zs, 15 + 16, version(), static_cast<int>(sizeof(ZStream))) != 0) {
    *err = "zlib initialisation failed.";
    return false;
  }
  std::vector<unsigned char> chunk(1 << 20);
  zs.next_in = reinterpret_cast<const unsigned char*>(in.data());
  size_t left = in.size();
  bool ok = true;
  while (left || zs.avail_in) {
    if (!zs.avail_in) {
      const unsigned take = static_cast<unsigned>(std::min<size_t>(left, 1u << 30));
      zs.avail_in = take;
      left -= take;
    }
    zs.next_out = chunk.data();
    zs.avail_out = static_cast<unsigned>(chunk.size());
    const int rc = inflate(&zs, 0);
    out->append(reinterpret_cast<const char*>(chunk.data()), chunk.size() - zs.avail_out);
    if (rc == 1) {  // Z_STREAM_END: bgzf and `cat a.gz b.gz` are sequences of members
      if (!zs.avail_in && !left) break;
      reset(&zs);
    } else if (rc != 0) {
      ok = false;
      break;
    }
  }
  end(&zs);
  if (!ok) *err = "Malformed gzip stream.";
  return ok;
}
}  // namespace

bool ReadLines(const std::string& path, std::vector<std::string>* lines, std::string* err) {
  FILE* f = fopen(path.c_str(), "rb");
  if (!f) {
    *err = "Failed to open " + path + ".";
    return false;
  }
  std::string raw;
  {
    std::vector<char> chunk(1 << 20);
    size_t got;
    while ((got = fread(chunk.data(), 1, chunk.size(), f)) > 0) raw.append(chunk.data(), got);
    fclose(f);
  }
  std::string text;
  const unsigned char* b = reinterpret_cast<const unsigned char*>(raw.data());
  if (raw.size() >= 4 && b[0] == 0x28 && b[1] == 0xB5 && b[2] == 0x2F && b[3] == 0xFD) {
    if (!ZstdDecompressAll(raw, &text, err)) {
      *err += " (" + path + ")";
      return false;
    }
  } else if (raw.size() >= 2 && b[0] == 0x1F && b[1] == 0x8B) {
    if (!GzipDecompressAll(raw, &text, err)) {
      *err += " (" + path + ")";
      return false;
    }
  } else {
    text.swap(raw);
  }
  size_t pos = 0;
  while (pos < text.size()) {
    size_t end = text.find('\n', pos);
    if (end == std::string::npos) end = text.size();
    size_t stop = end;
    if (stop > pos && text[stop - 1] == '\r') --stop;
    lines->emplace_back(text, pos, stop - pos);
    pos = end + 1;
  }
  return true;
}

}  // namespace pl2host
