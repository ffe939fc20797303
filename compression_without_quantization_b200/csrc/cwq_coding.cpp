// cwq_coding.cpp -- host-side integer arithmetic coder for the side information (group sizes), the
// native counterpart of the reference's only compiled component, code/coding.pyx (Cython):
//   ArithmeticCoder.__init__  coding.pyx:29-55   cumulative counts C, D and total R
//   ArithmeticCoder.encode    coding.pyx:59-125
//   ArithmeticCoder.decode    coding.pyx:129-216  (linear symbol search)
//   ArithmeticCoder.decode_fast coding.pyx:220-310 (the AVL tree of data_structures.py:188-215 is a
//                              monotone search over C, done here by bisection)
// Sequential by nature and tiny (one symbol per latent group): it stays on the host, like the
// reference's.  The strict comparisons (`low > half`, `low > quarter`) are the reference's and are kept.
#include <cstdint>
#include <string>
#include <vector>
#include "../../include/cwq.h"

namespace cwq { void set_error(const std::string& msg); }

namespace {

struct Model {
  std::vector<uint64_t> C, D;
  uint64_t R = 0;
  bool build(const int64_t* counts, int32_t n, std::string& err) {
    if (n < 1) { err = "empty alphabet"; return false; }
    C.resize(n); D.resize(n);
    uint64_t c = 0;
    for (int32_t i = 0; i < n; ++i) {
      if (counts[i] < 1) { err = "every symbol needs a positive count (the reference's symbol tree cannot hold equal bounds)"; return false; }
      C[i] = c; c += (uint64_t)counts[i]; D[i] = c;
    }
    R = c;
    if (R >= (1ull << 31)) { err = "total count must stay below 2^31"; return false; }
    return true;
  }
};

inline void emit(uint8_t* out, int64_t cap, int64_t& n, char first, int64_t s) {
  if (n < cap) out[n] = (uint8_t)first;
  ++n;
  const char other = first == '0' ? '1' : '0';
  for (int64_t i = 0; i < s; ++i) { if (n < cap) out[n] = (uint8_t)other; ++n; }
}

}  // namespace

#define AC_REQUIRE(cond, msg) do { if (!(cond)) { cwq::set_error(std::string(__func__) + ": " + (msg)); return CWQ_ERR_ARG; } } while (0)

extern "C" int cwq_ac_encode_host(const int64_t* h_counts, int32_t n_symbols, const int64_t* h_message, int64_t n_message,
                                  int32_t precision, uint8_t* h_ascii_bits, int64_t capacity, int64_t* h_n_bits) {
  AC_REQUIRE(h_counts && h_n_bits && (h_message || n_message == 0) && (h_ascii_bits || capacity == 0), "null argument");
  AC_REQUIRE(precision >= 8 && precision <= 32, "precision must be in [8, 32]");
  Model m; std::string err;
  AC_REQUIRE(m.build(h_counts, n_symbols, err), err);
  const uint64_t whole = 1ull << precision, half = whole >> 1, quarter = whole >> 2;
  uint64_t low = 0, high = whole;
  int64_t s = 0, n = 0;
  for (int64_t k = 0; k < n_message; ++k) {
    const int64_t sym = h_message[k];
    AC_REQUIRE(sym >= 0 && sym < n_symbols, "symbol outside the alphabet");
    const uint64_t width = high - low;
    high = low + (width * m.D[sym]) / m.R;
    low = low + (width * m.C[sym]) / m.R;
    while (high < half || low > half) {
      if (high < half) { emit(h_ascii_bits, capacity, n, '0', s); s = 0; low *= 2; high *= 2; }
      else { emit(h_ascii_bits, capacity, n, '1', s); s = 0; low = (low - half) * 2; high = (high - half) * 2; }
    }
    while (low > quarter && high < 3 * quarter) { s += 1; low = (low - quarter) * 2; high = (high - quarter) * 2; }
  }
  s += 1;
  if (low <= quarter) emit(h_ascii_bits, capacity, n, '0', s); else emit(h_ascii_bits, capacity, n, '1', s);
  *h_n_bits = n;
  if (n > capacity) { cwq::set_error("cwq_ac_encode_host: output capacity too small"); return CWQ_ERR_WORKSPACE; }
  return CWQ_OK;
}

extern "C" int cwq_ac_decode_host(const int64_t* h_counts, int32_t n_symbols, const uint8_t* h_ascii_bits, int64_t n_bits,
                                  int32_t precision, int32_t fast, int64_t* h_message, int64_t capacity, int64_t* h_n_message,
                                  int64_t* h_bits_used) {
  AC_REQUIRE(h_counts && h_n_message && (h_ascii_bits || n_bits == 0) && (h_message || capacity == 0), "null argument");
  AC_REQUIRE(precision >= 8 && precision <= 32, "precision must be in [8, 32]");
  Model m; std::string err;
  AC_REQUIRE(m.build(h_counts, n_symbols, err), err);
  const uint64_t whole = 1ull << precision, half = whole >> 1, quarter = whole >> 2;
  uint64_t low = 0, high = whole, z = 0;
  int64_t i = 0, n = 0;
  auto bit = [&](int64_t p) -> uint64_t { return (p < n_bits && h_ascii_bits[p] == '1') ? 1ull : 0ull; };
  while (i < precision && i < n_bits) { if (h_ascii_bits[i] == '1') z += 1ull << (precision - i - 1); ++i; }
  for (;;) {
    const uint64_t width = high - low;
    int64_t j = -1;
    if (fast) {
      // largest j with (width * C[j]) / R <= z - low   (data_structures.py:188-215, monotone in j)
      AC_REQUIRE(z >= low, "corrupt stream (value below the interval)");
      const uint64_t target = z - low;
      int64_t lo = 0, hi = n_symbols - 1;
      while (lo < hi) { int64_t mid = (lo + hi + 1) >> 1; if ((width * m.C[mid]) / m.R <= target) lo = mid; else hi = mid - 1; }
      j = lo;
    } else {
      for (int64_t t = 0; t < n_symbols; ++t) {   // coding.pyx:162-170: first j with low_ <= z < high_
        const uint64_t lo_ = low + (width * m.C[t]) / m.R, hi_ = low + (width * m.D[t]) / m.R;
        if (lo_ <= z && z < hi_) { j = t; break; }
      }
      AC_REQUIRE(j >= 0, "corrupt stream (no symbol interval contains the value)");
    }
    if (n < capacity) h_message[n] = j;
    ++n;
    high = low + (width * m.D[j]) / m.R;
    low = low + (width * m.C[j]) / m.R;
    if (j == 0) break;   // symbol 0 ends the message (coding.pyx:180, 271)
    // A message may hold far more symbols than the stream has bits (a skewed histogram codes a symbol in a fraction of a
    // bit): symbols past `capacity` are counted, not stored, and the caller retries with the count (CWQ_ERR_WORKSPACE).
    // Only the stream position bounds the loop: every symbol narrows the interval, so a corrupt stream runs out of bits.
    AC_REQUIRE(i <= n_bits + 64 * (int64_t)precision, "corrupt stream (no end symbol within the stream)");
    while (high < half || low > half) {
      if (high < half) { low *= 2; high *= 2; z *= 2; }
      else { low = (low - half) * 2; high = (high - half) * 2; z = (z - half) * 2; }
      z += bit(i); ++i;
    }
    while (low > quarter && high < 3 * quarter) {
      low = (low - quarter) * 2; high = (high - quarter) * 2; z = (z - quarter) * 2;
      z += bit(i); ++i;
    }
  }
  *h_n_message = n;
  if (h_bits_used) *h_bits_used = i;
  if (n > capacity) { cwq::set_error("cwq_ac_decode_host: output capacity too small"); return CWQ_ERR_WORKSPACE; }
  return CWQ_OK;
}
