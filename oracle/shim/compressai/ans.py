"""compressai.ans restatement (TEST INFRASTRUCTURE -- pure Python, small inputs only).

CompressAI is an un-vendored, un-pinned dependency of the reference (README.md:41,:70) and is not
installed here, so this restates its published range coder [upstream: compressai/cpp_exts/rans/
rans_interface.cpp + ryg_rans/rans64.h] -- PARITY UNPINNED against real CompressAI bitstreams;
parity is anchored on the reference's own call sites (CLC_run.py:654-656, :712-713, :758-760, :793):

  * rANS with a 64-bit state, 32-bit renormalisation words, lower bound L = 2^31 (rans64.h);
  * 16-bit probability precision, 4-bit bypass precision (max_bypass_val = 15);
  * per symbol: value = symbol - offset[idx]; values outside [0, max_value) are coded as the
    sentinel `max_value` (= cdf_size - 2) followed by a bypass-coded raw value
    (negative: -2v-1, overflow: 2(v-max_value)), the number of 4-bit groups first (unary in
    steps of 15), then the groups, least significant first;
  * the encoder consumes the symbol list BACKWARDS and writes words from the end of the buffer
    towards its start; the string is the used tail of the buffer, little-endian uint32 words.

`pmf_to_quantized_cdf` restates compressai/cpp_exts/ops/ops.cpp.
"""
import struct

PRECISION = 16
BYPASS_PRECISION = 4
MAX_BYPASS_VAL = (1 << BYPASS_PRECISION) - 1
RANS64_L = 1 << 31
_M64 = (1 << 64) - 1


def pmf_to_quantized_cdf(pmf, precision=16):
    """list of float probabilities -> list of len(pmf)+1 cumulative frequencies summing to 2^precision,
    every symbol with frequency >= 1 (frequency stolen from the least frequent symbol with freq > 1)."""
    import numpy as np
    pmf = [float(np.float32(p)) for p in pmf]
    if any(p < 0 or p != p or p == float("inf") for p in pmf):
        raise ValueError("Invalid `pmf`, non-finite or negative element found.")
    scale = 1 << precision
    # std::round on float (half away from zero), evaluated in fp32 like the C++ lambda
    import math
    cdf = [0] + [int(math.floor(float(np.float32(p) * np.float32(scale)) + 0.5)) for p in pmf]
    total = sum(cdf)
    if total == 0:
        raise ValueError("Invalid `pmf`: at least one element must have a non-zero probability.")
    cdf = [(scale * c) // total for c in cdf]
    for i in range(1, len(cdf)):
        cdf[i] += cdf[i - 1]
    cdf[-1] = scale
    n = len(cdf)
    for i in range(n - 1):
        if cdf[i] == cdf[i + 1]:
            best_freq, best_steal = 1 << 32, -1
            for j in range(n - 1):
                freq = cdf[j + 1] - cdf[j]
                if 1 < freq < best_freq:
                    best_freq, best_steal = freq, j
            assert best_steal != -1
            if best_steal < i:
                for j in range(best_steal + 1, i + 1):
                    cdf[j] -= 1
            else:
                for j in range(i + 1, best_steal + 1):
                    cdf[j] += 1
    return cdf


def _expand(symbols, indexes, cdfs, cdfs_sizes, offsets):
    """(start, range, bypass) triples in coding order (rans_interface.cpp encode_with_indexes)."""
    out = []
    for s, ci in zip(symbols, indexes):
        cdf = cdfs[ci]
        max_value = cdfs_sizes[ci] - 2
        value = s - offsets[ci]
        raw = 0
        if value < 0:
            raw = -2 * value - 1
            value = max_value
        elif value >= max_value:
            raw = 2 * (value - max_value)
            value = max_value
        out.append((cdf[value], cdf[value + 1] - cdf[value], False))
        if value == max_value:
            n_bypass = 0
            while (raw >> (n_bypass * BYPASS_PRECISION)) != 0:
                n_bypass += 1
            val = n_bypass
            while val >= MAX_BYPASS_VAL:
                out.append((MAX_BYPASS_VAL, MAX_BYPASS_VAL + 1, True))
                val -= MAX_BYPASS_VAL
            out.append((val, val + 1, True))
            for j in range(n_bypass):
                v = (raw >> (j * BYPASS_PRECISION)) & MAX_BYPASS_VAL
                out.append((v, v + 1, True))
    return out


def _flush(syms):
    x = RANS64_L
    words = []          # emitted in reverse buffer order
    for start, rng, bypass in reversed(syms):
        if not bypass:
            x_max = ((RANS64_L >> PRECISION) << 32) * rng
            if x >= x_max:
                words.append(x & 0xFFFFFFFF)
                x >>= 32
            x = ((x // rng) << PRECISION) + (x % rng) + start
        else:
            freq = 1 << (16 - BYPASS_PRECISION)
            x_max = ((RANS64_L >> 16) << 32) * freq
            if x >= x_max:
                words.append(x & 0xFFFFFFFF)
                x >>= 32
            x = ((x << BYPASS_PRECISION) | start) & _M64
    words.append((x >> 32) & 0xFFFFFFFF)
    words.append(x & 0xFFFFFFFF)
    words.reverse()
    return struct.pack("<%dI" % len(words), *words)


class BufferedRansEncoder:
    def __init__(self):
        self._syms = []

    def encode_with_indexes(self, symbols, indexes, cdfs, cdfs_sizes, offsets):
        assert len(symbols) == len(indexes)
        self._syms.extend(_expand(symbols, indexes, cdfs, cdfs_sizes, offsets))

    def flush(self):
        out = _flush(self._syms)
        self._syms = []
        return out


class RansEncoder:
    def encode_with_indexes(self, symbols, indexes, cdfs, cdfs_sizes, offsets):
        enc = BufferedRansEncoder()
        enc.encode_with_indexes(symbols, indexes, cdfs, cdfs_sizes, offsets)
        return enc.flush()


class RansDecoder:
    def __init__(self):
        self._words, self._pos, self._x = (), 0, 0

    def set_stream(self, encoded):
        n = len(encoded) // 4
        self._words = struct.unpack("<%dI" % n, encoded[:4 * n])
        self._x = self._words[0] | (self._words[1] << 32)
        self._pos = 2

    def _renorm(self):
        if self._x < RANS64_L:
            self._x = (self._x << 32) | self._words[self._pos]
            self._pos += 1

    def _get_bits(self, n_bits):
        val = self._x & ((1 << n_bits) - 1)
        self._x >>= n_bits
        self._renorm()
        return val

    def decode_stream(self, indexes, cdfs, cdfs_sizes, offsets):
        out = []
        mask = (1 << PRECISION) - 1
        for ci in indexes:
            cdf = cdfs[ci]
            max_value = cdfs_sizes[ci] - 2
            cum = self._x & mask
            s = 0
            n = cdfs_sizes[ci]
            while s < n and not cdf[s] > cum:       # std::find_if(first element > cum_freq)
                s += 1
            s -= 1
            start, rng = cdf[s], cdf[s + 1] - cdf[s]
            self._x = rng * (self._x >> PRECISION) + (self._x & mask) - start
            self._renorm()
            value = s
            if value == max_value:
                val = self._get_bits(BYPASS_PRECISION)
                n_bypass = val
                while val == MAX_BYPASS_VAL:
                    val = self._get_bits(BYPASS_PRECISION)
                    n_bypass += val
                raw = 0
                for j in range(n_bypass):
                    raw |= self._get_bits(BYPASS_PRECISION) << (j * BYPASS_PRECISION)
                value = raw >> 1
                if raw & 1:
                    value = -value - 1
                else:
                    value += max_value
            out.append(value + offsets[ci])
        return out

    def decode_with_indexes(self, encoded, indexes, cdfs, cdfs_sizes, offsets):
        self.set_stream(encoded)
        return self.decode_stream(indexes, cdfs, cdfs_sizes, offsets)
