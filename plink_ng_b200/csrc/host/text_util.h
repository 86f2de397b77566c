// text_util.h - number formatting and small text helpers of the host program.
//
// dtoa_g restates the OUTPUT CONTRACT of the reference's dtoa_g (2.0/include/plink2_string.cc:2507-2639):
// printf("%g")-like, 6 significant digits, trailing zeros stripped, "e-05"-style exponents,
// lowercase nan/inf, round-half-to-even decided with a +-5e-9 tolerance band (the reference's
// kBankerRound8 constants, :2232) after scaling by the same exact powers of ten, so that
// .king / .kin0 / .rel / .eigenvec text is byte-identical.
#pragma once
#include <cstdint>
#include <cstdio>
#include <string>
#include <vector>

namespace pl2host {

// Appends the decimal text of `x`; returns the new end pointer.  `buf` needs >= 16 free bytes.
char* dtoa_g(double x, char* buf);
// 8 significant digits (the reference's dtoa_g_p8, plink2_string.cc:2641); `buf` needs >= 24 free bytes.
char* dtoa_g_p8(double x, char* buf);
char* u32toa(uint32_t x, char* buf);
char* i32toa(int32_t x, char* buf);

// Buffered file writer (the reference streams through a ~1 MB buffer as well).
class OutFile {
 public:
  OutFile() = default;
  ~OutFile() { Close(); }
  // zst: Zstandard-compressed output (the reference's 'zs' modifiers; its writers go through zstd's streaming
  // API, 2.0/include/plink2_zstfile / plink2_compress_stream).  Here every flushed buffer becomes one complete
  // frame (concatenated frames are a valid .zst stream); libzstd.so.1 is resolved at run time.
  bool Open(const std::string& path, bool zst = false);
  bool Close();  // true iff every write succeeded
  char* Reserve(size_t n);  // pointer to >= n writable bytes
  void Advance(char* new_end) { pos_ = static_cast<size_t>(new_end - buf_.data()); }
  void Write(const void* p, size_t n);
  void Puts(const char* s);
  bool ok() const { return ok_; }

 private:
  void Flush();
  FILE* f_ = nullptr;
  std::vector<char> buf_;
  size_t pos_ = 0;
  bool ok_ = true;
  bool zst_ = false;
  std::vector<char> zbuf_;
};

std::vector<std::string> SplitWs(const std::string& line);
bool ReadLines(const std::string& path, std::vector<std::string>* lines, std::string* err);

}  // namespace pl2host
