// Range-coder side of compress()/decompress(): the bit-serial rANS state
// machine and the CDF quantiser.  HOST code: one rANS stream is a strictly sequential recurrence
// (x' depends on x for every symbol), so it runs on the host next to the device kernels that
// produce its inputs (clc_gc_symbols_indexes: symbols and scale-table indexes stay on the device
// until one pinned D2H copy per stream) -- this replaces the reference's per-symbol Python lists
// (`.tolist()` of every symbol / index / the whole CDF table per call, CLC_run.py:650-652,
// :693-694, :712) and compressai.ans' pybind11 vector conversions.
//
// Wire format = compressai.ans [upstream: cpp_exts/rans/rans_interface.cpp, ryg_rans rans64.h;
// un-vendored and un-pinned in the reference, restated from the published algorithm]:
// 64-bit state, L = 2^31, 32-bit little-endian renormalisation words written back to front,
// 16-bit probabilities, 4-bit bypass groups for out-of-range values.
#include "common.cuh"

#include <algorithm>
#include <vector>

namespace {

constexpr int kPrecision = 16;
constexpr int kBypassPrecision = 4;
constexpr uint32_t kMaxBypassVal = (1u << kBypassPrecision) - 1;
constexpr uint64_t kRansL = 1ull << 31;

struct Tables {
  const int32_t* cdfs; int32_t n_cdfs; int32_t stride; const int32_t* sizes; const int32_t* offsets;
};

inline bool tables_ok(const Tables& t) {
  if (!t.cdfs || !t.sizes || !t.offsets || t.n_cdfs < 1 || t.stride < 2) return false;
  for (int i = 0; i < t.n_cdfs; ++i)
    if (t.sizes[i] < 2 || t.sizes[i] > t.stride) return false;
  return true;
}

// Writer that fills a word buffer from its END towards its start (rans64 convention).
struct BackWriter {
  uint32_t* base; uint32_t* ptr;
  bool put(uint32_t w) {
    if (ptr == base) return false;
    *--ptr = w;
    return true;
  }
};

inline bool enc_put(uint64_t& x, BackWriter& w, uint32_t start, uint32_t freq) {
  const uint64_t x_max = ((kRansL >> kPrecision) << 32) * freq;
  if (x >= x_max) {
    if (!w.put((uint32_t)x)) return false;
    x >>= 32;
  }
  x = ((x / freq) << kPrecision) + (x % freq) + start;
  return true;
}

inline bool enc_put_bits(uint64_t& x, BackWriter& w, uint32_t val) {
  const uint32_t freq = 1u << (16 - kBypassPrecision);
  const uint64_t x_max = ((kRansL >> 16) << 32) * freq;
  if (x >= x_max) {
    if (!w.put((uint32_t)x)) return false;
    x >>= 32;
  }
  x = (x << kBypassPrecision) | val;
  return true;
}

}  // namespace

// pmf [n] -> cdf [n + 1]; compressai `pmf_to_quantized_cdf` (cpp_exts/ops/ops.cpp).
extern "C" int clc_pmf_to_quantized_cdf(const float* pmf, int32_t n, int32_t precision, int32_t* cdf_out) {
  if (!pmf || !cdf_out || n < 1 || precision < 1 || precision > 16) return CLC_ERR_INVALID_ARGUMENT;
  for (int i = 0; i < n; ++i)
    if (!(pmf[i] >= 0.f) || pmf[i] > 3.0e38f) return CLC_ERR_INVALID_ARGUMENT;   // negative, NaN or inf
  std::vector<uint32_t> cdf((size_t)n + 1);
  cdf[0] = 0;
  const float scale = (float)(1 << precision);
  for (int i = 0; i < n; ++i) cdf[i + 1] = (uint32_t)roundf(pmf[i] * scale);   // half away from zero, fp32
  uint64_t total = 0;
  for (uint32_t c : cdf) total += c;
  if (total == 0) return CLC_ERR_INVALID_ARGUMENT;
  for (auto& c : cdf) c = (uint32_t)((((uint64_t)1 << precision) * c) / total);
  for (int i = 1; i <= n; ++i) cdf[i] += cdf[i - 1];
  cdf[n] = 1u << precision;
  for (int i = 0; i < n; ++i) {
    if (cdf[i] == cdf[i + 1]) {
      // a zero-frequency symbol: steal one count from the least frequent symbol with freq > 1
      uint32_t best_freq = ~0u;
      int best_steal = -1;
      for (int j = 0; j < n; ++j) {
        const uint32_t freq = cdf[j + 1] - cdf[j];
        if (freq > 1 && freq < best_freq) { best_freq = freq; best_steal = j; }
      }
      if (best_steal < 0) return CLC_ERR_INVALID_ARGUMENT;   // more symbols than counts
      if (best_steal < i) {
        for (int j = best_steal + 1; j <= i; ++j) cdf[j]--;
      } else {
        for (int j = i + 1; j <= best_steal; ++j) cdf[j]++;
      }
    }
  }
  for (int i = 0; i <= n; ++i) cdf_out[i] = (int32_t)cdf[i];
  return CLC_OK;
}

// Upper bound on the stream length of clc_rans_encode for n symbols (exact worst case: one
// renormalisation word per coded item, <= 12 items per symbol when it escapes to bypass coding).
extern "C" size_t clc_rans_encode_capacity(int64_t n) {
  return n < 0 ? 0 : (size_t)(4 * (13 * (uint64_t)n + 4));
}

// RansEncoder.encode_with_indexes / BufferedRansEncoder.{encode_with_indexes, flush}
// (compressai.ans; call sites CLC_run.py:654, :712-713): ONE stream over `n` symbols.
//   symbols, indexes : host int32 [n];  cdfs : host int32 [n_cdfs, cdf_stride];
//   cdf_sizes, offsets : host int32 [n_cdfs]
//   out : host buffer of out_capacity bytes; the stream is written to out[0 .. *out_bytes).
extern "C" int clc_rans_encode(const int32_t* symbols, const int32_t* indexes, int64_t n, const int32_t* cdfs,
                               int32_t n_cdfs, int32_t cdf_stride, const int32_t* cdf_sizes,
                               const int32_t* offsets, uint8_t* out, size_t out_capacity, size_t* out_bytes) {
  const Tables t = {cdfs, n_cdfs, cdf_stride, cdf_sizes, offsets};
  if (n < 0 || (n > 0 && (!symbols || !indexes)) || !out || !out_bytes || !tables_ok(t))
    return CLC_ERR_INVALID_ARGUMENT;
  if ((reinterpret_cast<uintptr_t>(out) & 3) != 0) return CLC_ERR_INVALID_ARGUMENT;
  const size_t cap_words = out_capacity / 4;
  if (cap_words < 2) return CLC_ERR_WORKSPACE;
  uint32_t* base = reinterpret_cast<uint32_t*>(out);
  BackWriter w = {base, base + cap_words};
  uint64_t x = kRansL;
  // symbols are consumed BACKWARDS; within one symbol the bypass items were pushed after the
  // sentinel, so in reverse they come first: raw groups (high to low), the group count, the sentinel
  for (int64_t i = n - 1; i >= 0; --i) {
    const int32_t ci = indexes[i];
    if (ci < 0 || ci >= n_cdfs) return CLC_ERR_INVALID_ARGUMENT;
    const int32_t* cdf = cdfs + (int64_t)ci * cdf_stride;
    const int32_t max_value = cdf_sizes[ci] - 2;
    int64_t value = (int64_t)symbols[i] - offsets[ci];
    uint64_t raw = 0;
    bool bypass = false;
    if (value < 0) {
      raw = (uint64_t)(-2 * value - 1);
      value = max_value;
      bypass = true;
    } else if (value >= max_value) {
      raw = (uint64_t)(2 * (value - max_value));
      value = max_value;
      bypass = true;
    }
    if (bypass) {
      int n_bypass = 0;
      while ((raw >> (n_bypass * kBypassPrecision)) != 0) ++n_bypass;
      for (int j = n_bypass - 1; j >= 0; --j)
        if (!enc_put_bits(x, w, (uint32_t)((raw >> (j * kBypassPrecision)) & kMaxBypassVal))) return CLC_ERR_WORKSPACE;
      // group count: pushed as [15]*q then remainder -> reversed: remainder first
      const uint32_t q = (uint32_t)n_bypass / kMaxBypassVal, rem = (uint32_t)n_bypass % kMaxBypassVal;
      if (!enc_put_bits(x, w, rem)) return CLC_ERR_WORKSPACE;
      for (uint32_t j = 0; j < q; ++j)
        if (!enc_put_bits(x, w, kMaxBypassVal)) return CLC_ERR_WORKSPACE;
    }
    const uint32_t start = (uint32_t)cdf[value], freq = (uint32_t)(cdf[value + 1] - cdf[value]);
    if (freq == 0 || cdf[value + 1] > (1 << kPrecision)) return CLC_ERR_INVALID_ARGUMENT;
    if (!enc_put(x, w, start, freq)) return CLC_ERR_WORKSPACE;
  }
  if (!w.put((uint32_t)(x >> 32)) || !w.put((uint32_t)x)) return CLC_ERR_WORKSPACE;
  const size_t nwords = (size_t)((base + cap_words) - w.ptr);
  if (w.ptr != base) memmove(base, w.ptr, nwords * 4);
  *out_bytes = nwords * 4;
  return CLC_OK;
}

// RansDecoder.{set_stream, decode_stream, decode_with_indexes} (compressai.ans; call sites
// CLC_run.py:758-760, :793).  `state` = {x, next word position}: pass {0, 0} for the first call on a
// stream (set_stream); consecutive calls continue the same stream (decode_stream per slice).
extern "C" int clc_rans_decode(const uint8_t* stream, size_t stream_bytes, uint64_t* state, const int32_t* indexes,
                               int64_t n, const int32_t* cdfs, int32_t n_cdfs, int32_t cdf_stride,
                               const int32_t* cdf_sizes, const int32_t* offsets, int32_t* out) {
  const Tables t = {cdfs, n_cdfs, cdf_stride, cdf_sizes, offsets};
  if (!stream || !state || n < 0 || (n > 0 && (!indexes || !out)) || !tables_ok(t)) return CLC_ERR_INVALID_ARGUMENT;
  const size_t nwords = stream_bytes / 4;
  auto word = [&](size_t i) -> uint32_t {
    uint32_t v;
    memcpy(&v, stream + 4 * i, 4);      // the bytes object need not be 4-byte aligned
    return v;
  };
  uint64_t x = state[0];
  size_t pos = (size_t)state[1];
  if (pos == 0) {
    if (nwords < 2) return CLC_ERR_INVALID_ARGUMENT;
    x = (uint64_t)word(0) | ((uint64_t)word(1) << 32);
    pos = 2;
  }
  const uint64_t mask = (1ull << kPrecision) - 1;
  bool truncated = false;
  // Long streams: per-table bucket index over the top 8 bits of the cumulative frequency, lut[b] = the
  // symbol whose interval contains b * 256, so the search for `cum` only scans lut[b] .. lut[b + 1]
  // (a handful of entries) instead of bisecting the whole row.  Built per call (64 x ~3k entries).
  constexpr int kBuckets = 1 << 8, kShift = kPrecision - 8;
  std::vector<int32_t> lut;
  const bool use_lut = n >= 4096;
  if (use_lut) {
    lut.resize((size_t)n_cdfs * (kBuckets + 1));
    for (int c = 0; c < n_cdfs; ++c) {
      const int32_t* cdf = cdfs + (int64_t)c * cdf_stride;
      const int32_t size = cdf_sizes[c];
      int32_t* l = lut.data() + (size_t)c * (kBuckets + 1);
      int32_t sidx = 0;
      for (int b = 0; b < kBuckets; ++b) {
        const int32_t v = b << kShift;
        while (sidx + 1 < size && cdf[sidx + 1] <= v) ++sidx;
        l[b] = sidx;
      }
      l[kBuckets] = size - 2 > 0 ? size - 2 : 0;
    }
  }
  auto renorm = [&]() {
    if (x < kRansL) {
      if (pos >= nwords) { truncated = true; return; }
      x = (x << 32) | word(pos++);
    }
  };
  auto get_bits = [&]() -> uint32_t {
    const uint32_t v = (uint32_t)(x & kMaxBypassVal);
    x >>= kBypassPrecision;
    renorm();
    return v;
  };
  for (int64_t i = 0; i < n; ++i) {
    const int32_t ci = indexes[i];
    if (ci < 0 || ci >= n_cdfs) return CLC_ERR_INVALID_ARGUMENT;
    const int32_t* cdf = cdfs + (int64_t)ci * cdf_stride;
    const int32_t size = cdf_sizes[ci], max_value = size - 2;
    const uint32_t cum = (uint32_t)(x & mask);
    // first entry > cum (the table is increasing): bucketed / binary search instead of the upstream linear scan
    int32_t s;
    if (use_lut) {
      const int32_t* l = lut.data() + (size_t)ci * (kBuckets + 1);
      s = l[cum >> kShift];
      const int32_t hi = l[(cum >> kShift) + 1];
      while (s < hi && cdf[s + 1] <= (int32_t)cum) ++s;
    } else {
      const int32_t* it = std::upper_bound(cdf, cdf + size, (int32_t)cum);
      s = (int32_t)(it - cdf) - 1;
    }
    if (s < 0 || s > max_value) return CLC_ERR_INVALID_ARGUMENT;
    const uint32_t start = (uint32_t)cdf[s], freq = (uint32_t)(cdf[s + 1] - cdf[s]);
    x = (uint64_t)freq * (x >> kPrecision) + (x & mask) - start;
    renorm();
    int64_t value = s;
    if (s == max_value) {
      uint32_t val = get_bits();
      uint32_t n_bypass = val;
      while (val == kMaxBypassVal && !truncated) {
        val = get_bits();
        n_bypass += val;
      }
      if (n_bypass > 16) return CLC_ERR_INVALID_ARGUMENT;   // raw values are < 2^33
      uint64_t raw = 0;
      for (uint32_t j = 0; j < n_bypass; ++j) raw |= (uint64_t)get_bits() << (j * kBypassPrecision);
      value = (int64_t)(raw >> 1);
      if (raw & 1) value = -value - 1;
      else value += max_value;
    }
    if (truncated) return CLC_ERR_INVALID_ARGUMENT;
    out[i] = (int32_t)(value + offsets[ci]);
  }
  state[0] = x;
  state[1] = (uint64_t)pos;
  return CLC_OK;
}
