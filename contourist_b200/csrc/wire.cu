// Wire formats: the text the reference's emitters produce, written by host threads.
//
// html_demo.py:118-161 (grid_html_page, emit_three_json), morph_geometry.py:91-128 (MorphTriangles.to_json,
// flatten_json_list) and triangulated.py:16-50 build their output as  "[" + sep.join(str(x) ...) + "]"  in Python:
// at 10^7 triangles the joins take longer than the extraction by four orders of magnitude.  This file is that one
// loop in C++: a 2D array of numbers -> "[" row (row_sep row)* "]", row = row_prefix value (col_sep value)* row_suffix,
// every value formatted exactly like Python's str(): integers in decimal, floats as repr(float) (shortest digits
// that round-trip; fixed notation for 1e-4 <= |x| < 1e16, else d.ddde+XX; always a ".0" on integral values; float32
// inputs widened first, as float(np.float32(x)) does).  No device code: nvcc is only the compiler driver here.
#include <charconv>
#include <cmath>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <string>
#include <thread>
#include <vector>

#include "common.cuh"

namespace {

inline char* put_int(char* p, long long v) {
  unsigned long long u = v < 0 ? 0ull - (unsigned long long)v : (unsigned long long)v;
  if (v < 0) *p++ = '-';
  char tmp[24];
  int n = 0;
  do {
    tmp[n++] = (char)('0' + u % 10);
    u /= 10;
  } while (u);
  while (n) *p++ = tmp[--n];
  return p;
}

// repr(float): CPython's format_float_short(..., 'r'): shortest round-trip digits, exponent form iff decpt <= -4 or > 16
inline char* put_double(char* p, double v) {
  if (std::isnan(v)) {
    memcpy(p, "nan", 3);
    return p + 3;
  }
  if (std::signbit(v)) {
    *p++ = '-';
    v = -v;
  }
  if (std::isinf(v)) {
    memcpy(p, "inf", 3);
    return p + 3;
  }
  if (v == 0.0) {
    memcpy(p, "0.0", 3);
    return p + 3;
  }
  char tmp[40];
  const std::to_chars_result r = std::to_chars(tmp, tmp + sizeof(tmp), v, std::chars_format::scientific);
  // tmp = d[.ddd]e[+-]XX
  char digits[24];
  int nd = 0;
  const char* q = tmp;
  for (; q < r.ptr && *q != 'e'; ++q)
    if (*q != '.') digits[nd++] = *q;
  ++q;                                                // past 'e'
  const bool eneg = *q == '-';
  ++q;
  int e10 = 0;
  for (; q < r.ptr; ++q) e10 = e10 * 10 + (*q - '0');
  if (eneg) e10 = -e10;
  const int decpt = e10 + 1;                          // value = 0.d1d2... * 10^decpt
  if (decpt <= -4 || decpt > 16) {
    *p++ = digits[0];
    if (nd > 1) {
      *p++ = '.';
      memcpy(p, digits + 1, nd - 1);
      p += nd - 1;
    }
    *p++ = 'e';
    *p++ = e10 < 0 ? '-' : '+';
    const int ae = e10 < 0 ? -e10 : e10;
    if (ae < 10) *p++ = '0';
    return put_int(p, ae);
  }
  if (decpt <= 0) {
    *p++ = '0';
    *p++ = '.';
    for (int i = 0; i < -decpt; ++i) *p++ = '0';
    memcpy(p, digits, nd);
    return p + nd;
  }
  if (decpt >= nd) {
    memcpy(p, digits, nd);
    p += nd;
    for (int i = nd; i < decpt; ++i) *p++ = '0';
    *p++ = '.';
    *p++ = '0';
    return p;
  }
  memcpy(p, digits, decpt);
  p += decpt;
  *p++ = '.';
  memcpy(p, digits + decpt, nd - decpt);
  return p + (nd - decpt);
}

struct Spec {
  const void* data;
  int dtype;                                          // ctr_wire_dtype
  int64_t rows, cols;
  std::string row_prefix, col_sep, row_suffix, row_sep;
};

inline char* put_value(char* p, const Spec& s, int64_t idx) {
  switch (s.dtype) {
    case CTR_WIRE_I32: return put_int(p, ((const int32_t*)s.data)[idx]);
    case CTR_WIRE_I64: return put_int(p, ((const int64_t*)s.data)[idx]);
    case CTR_WIRE_U32: return put_int(p, ((const uint32_t*)s.data)[idx]);
    case CTR_WIRE_F32: return put_double(p, (double)((const float*)s.data)[idx]);
    default: return put_double(p, ((const double*)s.data)[idx]);
  }
}

void format_rows(const Spec& s, int64_t r0, int64_t r1, bool first_chunk, std::string& out) {
  const size_t per_value = 26;                        // longest repr(float): -d.dddddddddddddddde-XXX = 24 characters
  const size_t per_row = s.row_prefix.size() + s.row_suffix.size() + s.row_sep.size() + (size_t)s.cols * (per_value + s.col_sep.size());
  const int64_t block = 4096;
  std::vector<char> buf((size_t)block * per_row + 64);
  out.clear();
  for (int64_t b0 = r0; b0 < r1; b0 += block) {
    const int64_t b1 = b0 + block < r1 ? b0 + block : r1;
    char* p = buf.data();
    for (int64_t r = b0; r < b1; ++r) {
      if (!(first_chunk && r == r0)) {
        memcpy(p, s.row_sep.data(), s.row_sep.size());
        p += s.row_sep.size();
      }
      memcpy(p, s.row_prefix.data(), s.row_prefix.size());
      p += s.row_prefix.size();
      for (int64_t c = 0; c < s.cols; ++c) {
        if (c) {
          memcpy(p, s.col_sep.data(), s.col_sep.size());
          p += s.col_sep.size();
        }
        p = put_value(p, s, r * s.cols + c);
      }
      memcpy(p, s.row_suffix.data(), s.row_suffix.size());
      p += s.row_suffix.size();
    }
    out.append(buf.data(), (size_t)(p - buf.data()));
  }
}

}  // namespace

extern "C" int ctr_wire_format(const void* data, int dtype, int64_t rows, int64_t cols, const char* row_prefix,
                               const char* col_sep, const char* row_suffix, const char* row_sep, int n_threads, char** out,
                               int64_t* out_len) {
  if (!out || !out_len || rows < 0 || cols < 0 || (rows * cols > 0 && !data) || dtype < CTR_WIRE_I32 || dtype > CTR_WIRE_F64)
    return CTR_ERR_BAD_ARG;
  Spec s{data, dtype, rows, cols, row_prefix ? row_prefix : "", col_sep ? col_sep : ",", row_suffix ? row_suffix : "",
         row_sep ? row_sep : ",\n"};
  int nth = n_threads > 0 ? n_threads : (int)std::thread::hardware_concurrency();
  if (nth < 1) nth = 1;
  if (nth > 64) nth = 64;
  if (rows < 65536) nth = 1;
  if ((int64_t)nth > rows) nth = rows > 0 ? (int)rows : 1;
  std::vector<std::string> parts((size_t)nth);
  std::vector<std::thread> th;
  try {
    for (int t = 1; t < nth; ++t)
      th.emplace_back([&, t] { format_rows(s, rows * t / nth, rows * (t + 1) / nth, false, parts[(size_t)t]); });
    format_rows(s, 0, rows / nth, true, parts[0]);
    for (auto& x : th) x.join();
  } catch (...) {
    for (auto& x : th)
      if (x.joinable()) x.join();
    return CTR_ERR_OOM;
  }
  size_t total = 2;
  for (auto& x : parts) total += x.size();
  char* buf = (char*)malloc(total + 1);
  if (!buf) return CTR_ERR_OOM;
  std::vector<size_t> off((size_t)nth);
  size_t o = 1;
  for (int t = 0; t < nth; ++t) {
    off[(size_t)t] = o;
    o += parts[(size_t)t].size();
  }
  buf[0] = '[';
  th.clear();
  for (int t = 1; t < nth; ++t) th.emplace_back([&, t] { memcpy(buf + off[(size_t)t], parts[(size_t)t].data(), parts[(size_t)t].size()); });
  memcpy(buf + off[0], parts[0].data(), parts[0].size());
  for (auto& x : th) x.join();
  buf[o] = ']';
  buf[o + 1] = 0;
  *out = buf;
  *out_len = (int64_t)(o + 1);
  return 0;
}

extern "C" void ctr_wire_free(char* p) { free(p); }
