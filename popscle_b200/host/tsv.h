// tsv.h — gzip/plain text line reader with whitespace splitting: the role tsv_reader plays in the
// reference (tsv_reader.cpp:12-52: hts_open/hts_getline/ksplit with delimiter 0), over zlib only.
#pragma once
#include <zlib.h>

#include <cstdint>
#include <cstdio>
#include <cstring>
#include <stdexcept>
#include <string>
#include <vector>

namespace pscl_host {

struct host_error : std::runtime_error {  // turned into the reference's "FATAL ERROR -" + abort by main()
  using std::runtime_error::runtime_error;
};

class LineReader {
 public:
  explicit LineReader(const std::string& path) : path_(path) {
    gz_ = gzopen(path.c_str(), "rb");  // transparent for uncompressed files
    if (!gz_) throw host_error("Cannot open file " + path + " for reading");
    gzbuffer(gz_, 1 << 20);
    buf_.resize(1 << 20);
  }
  ~LineReader() { if (gz_) gzclose(gz_); }
  LineReader(const LineReader&) = delete;
  LineReader& operator=(const LineReader&) = delete;

  // next line without its terminator; false at end of file
  bool next(std::string& line) {
    line.clear();
    bool got = false;
    for (;;) {
      if (pos_ == len_) {
        int n = gzread(gz_, buf_.data(), (unsigned)buf_.size());
        if (n <= 0) return got;
        len_ = (size_t)n; pos_ = 0;
      }
      got = true;
      const char* b = buf_.data() + pos_;
      const char* e = (const char*)memchr(b, '\n', len_ - pos_);
      if (e) {
        line.append(b, e - b);
        pos_ += (size_t)(e - b) + 1;
        if (!line.empty() && line.back() == '\r') line.pop_back();
        return true;
      }
      line.append(b, len_ - pos_);
      pos_ = len_;
    }
  }
  // the same, as a view into the reader's buffer (valid until the next call): no copy per line
  bool next_span(const char*& b, size_t& n) {
    for (;;) {
      const char* s = buf_.data() + pos_;
      const char* e = pos_ < len_ ? (const char*)memchr(s, '\n', len_ - pos_) : nullptr;
      if (e) {
        b = s; n = (size_t)(e - s); pos_ += n + 1;
        if (n && b[n - 1] == '\r') --n;
        return true;
      }
      if (eof_) {  // last line without a terminator
        if (pos_ == len_) return false;
        b = s; n = len_ - pos_; pos_ = len_;
        return true;
      }
      const size_t rem = len_ - pos_;  // a partial line: move it to the front and read on
      if (rem == buf_.size()) buf_.resize(buf_.size() * 2);
      if (rem && pos_) memmove(buf_.data(), buf_.data() + pos_, rem);
      const int got = gzread(gz_, buf_.data() + rem, (unsigned)(buf_.size() - rem));
      if (got <= 0) eof_ = true;
      len_ = rem + (got > 0 ? (size_t)got : 0); pos_ = 0;
    }
  }
  const std::string& path() const { return path_; }

 private:
  std::string path_;
  gzFile gz_ = nullptr;
  std::vector<char> buf_;
  size_t pos_ = 0, len_ = 0;
  bool eof_ = false;
};

// whitespace-separated fields of `line` (in place: the line is cut into NUL-terminated tokens)
inline int split_ws(std::string& line, std::vector<char*>& f) {
  f.clear();
  char* s = &line[0];
  const size_t l = line.size();
  size_t i = 0;
  while (i < l) {
    while (i < l && isspace((unsigned char)s[i])) ++i;
    if (i >= l) break;
    f.push_back(s + i);
    while (i < l && !isspace((unsigned char)s[i])) ++i;
    if (i < l) s[i++] = 0;
  }
  return (int)f.size();
}

inline void split_char(const std::string& s, char d, std::vector<std::string>& out) {
  // the strings of `out` are reused (assign keeps their capacity): no allocation per field once warmed up
  size_t n = 0, b = 0;
  for (;;) {
    const size_t e = s.find(d, b);
    const size_t len = e == std::string::npos ? std::string::npos : e - b;
    if (n < out.size()) out[n].assign(s, b, len);
    else out.emplace_back(s, b, len);
    ++n;
    if (e == std::string::npos) break;
    b = e + 1;
  }
  out.resize(n);
}

// tsv_reader semantics: a row, or false at EOF *or at the first empty line* (tsv_reader.cpp:37-41)
inline bool next_row(LineReader& r, std::string& line, std::vector<char*>& f) {
  if (!r.next(line) || line.empty()) return false;
  split_ws(line, f);
  return true;
}

}  // namespace pscl_host
