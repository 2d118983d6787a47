"""Range coder behind compress()/decompress(): same class and method names as the module the
reference imports (`BufferedRansEncoder`, `RansEncoder`, `RansDecoder`, CLC_run.py:10-11;
`pmf_to_quantized_cdf` behind `EntropyModel._pmf_to_cdf`), bound to the host entry points of
libclc_b200.so (csrc/rans.cu: clc_rans_encode / clc_rans_decode / clc_pmf_to_quantized_cdf).

Differences in kind, not in results: symbols / indexes / tables may be passed as int32 tensors or
numpy arrays (the reference passes Python lists built with `.tolist()`, CLC_run.py:650-652, :693-694);
lists are still accepted.  Device tensors are brought to the host with one copy per call.
"""
import ctypes as C

import numpy as np
import torch

from ._lib import call, lib


def _host_i32(a):
    """-> contiguous host int32 numpy array (1-D or 2-D) from a tensor / array / list."""
    if isinstance(a, torch.Tensor):
        a = a.detach()
        if a.dtype != torch.int32:
            a = a.to(torch.int32)
        return np.ascontiguousarray(a.cpu().numpy())
    return np.ascontiguousarray(np.asarray(a, dtype=np.int32))


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


class _Tables:
    """Quantised CDF table [n_cdfs, stride] + sizes + offsets, converted once per coder call."""

    def __init__(self, cdfs, cdfs_sizes, offsets):
        self.cdfs = _host_i32(cdfs)
        if self.cdfs.ndim != 2:
            raise ValueError(f"Invalid CDF size {self.cdfs.shape}")
        self.sizes = _host_i32(cdfs_sizes).reshape(-1)
        self.offsets = _host_i32(offsets).reshape(-1)
        if not (len(self.sizes) == len(self.offsets) == self.cdfs.shape[0]):
            raise ValueError("cdfs, cdfs_sizes and offsets must describe the same number of distributions")

    def args(self):
        return _p(self.cdfs), self.cdfs.shape[0], self.cdfs.shape[1], _p(self.sizes), _p(self.offsets)


def pmf_to_quantized_cdf(pmf, precision=16):
    """pmf (sequence of floats) -> list of len(pmf)+1 cumulative 16-bit frequencies."""
    p = np.ascontiguousarray(np.asarray(pmf, dtype=np.float32).reshape(-1))
    out = np.empty(len(p) + 1, dtype=np.int32)
    call("clc_pmf_to_quantized_cdf", _p(p), len(p), int(precision), _p(out))
    return out.tolist()


def _encode(symbols, indexes, tables):
    n = len(symbols)
    cap = lib().clc_rans_encode_capacity(n)
    out = np.empty(cap // 4, dtype=np.uint32)
    nbytes = C.c_size_t(0)
    call("clc_rans_encode", _p(symbols), _p(indexes), n, *tables.args(), _p(out), cap, C.byref(nbytes))
    return out.view(np.uint8)[:nbytes.value].tobytes()


class BufferedRansEncoder:
    """encode_with_indexes() may be called several times; flush() codes everything as ONE stream.  Every call's
    symbols are coded against the tables passed IN THAT CALL (as compressai's encoder resolves them): calls with
    different tables -- e.g. EntropyBottleneck and GaussianConditional symbols in one stream -- are merged into
    one combined table with the later calls' indexes shifted accordingly."""

    def __init__(self):
        self._sym, self._idx, self._tabs = [], [], []

    def encode_with_indexes(self, symbols, indexes, cdfs, cdfs_sizes, offsets):
        s, i = _host_i32(symbols).reshape(-1), _host_i32(indexes).reshape(-1)
        if len(s) != len(i):
            raise ValueError("symbols and indexes must have the same length")
        self._sym.append(s)
        self._idx.append(i)
        self._tabs.append(cdfs if isinstance(cdfs, _Tables) else _Tables(cdfs, cdfs_sizes, offsets))

    @staticmethod
    def _same(a, b):
        return a is b or (a.cdfs.shape == b.cdfs.shape and np.array_equal(a.cdfs, b.cdfs) and
                          np.array_equal(a.sizes, b.sizes) and np.array_equal(a.offsets, b.offsets))

    def flush(self):
        if not self._tabs:
            s = i = np.empty(0, dtype=np.int32)
            tables = _Tables(np.array([[0, 1 << 16]], dtype=np.int32), [2], [0])
        else:
            distinct, row0, idx = [], [], []
            for i_call, t in zip(self._idx, self._tabs):
                for j, u in enumerate(distinct):
                    if self._same(t, u):
                        base = row0[j]
                        break
                else:
                    base = sum(u.cdfs.shape[0] for u in distinct)
                    distinct.append(t)
                    row0.append(base)
                idx.append(i_call + base if base else i_call)
            if len(distinct) == 1:
                tables = distinct[0]
            else:
                stride = max(u.cdfs.shape[1] for u in distinct)
                cdfs = np.zeros((sum(u.cdfs.shape[0] for u in distinct), stride), dtype=np.int32)
                for u, b in zip(distinct, row0):
                    cdfs[b:b + u.cdfs.shape[0], :u.cdfs.shape[1]] = u.cdfs
                tables = _Tables(cdfs, np.concatenate([u.sizes for u in distinct]),
                                 np.concatenate([u.offsets for u in distinct]))
            s, i = np.concatenate(self._sym), np.concatenate(idx).astype(np.int32)
        self._sym, self._idx, self._tabs = [], [], []
        return _encode(s, i, tables)


class RansEncoder:
    def encode_with_indexes(self, symbols, indexes, cdfs, cdfs_sizes, offsets):
        enc = BufferedRansEncoder()
        enc.encode_with_indexes(symbols, indexes, cdfs, cdfs_sizes, offsets)
        return enc.flush()


class RansDecoder:
    def __init__(self):
        self._stream = None
        self._state = (C.c_uint64 * 2)(0, 0)

    def set_stream(self, encoded):
        self._stream = bytes(encoded)
        self._state[0] = self._state[1] = 0

    def decode_stream(self, indexes, cdfs, cdfs_sizes, offsets, as_tensor=False):
        """Next len(indexes) symbols of the stream.  Returns a list like the reference's coder
        (as_tensor=True: an int32 CPU tensor, no per-symbol Python objects)."""
        if self._stream is None:
            raise ValueError("set_stream() first")
        i = _host_i32(indexes).reshape(-1)
        tables = cdfs if isinstance(cdfs, _Tables) else _Tables(cdfs, cdfs_sizes, offsets)
        out = np.empty(len(i), dtype=np.int32)
        call("clc_rans_decode", self._stream, len(self._stream), self._state, _p(i), len(i), *tables.args(), _p(out))
        return torch.from_numpy(out) if as_tensor else out.tolist()

    def decode_with_indexes(self, encoded, indexes, cdfs, cdfs_sizes, offsets, as_tensor=False):
        self.set_stream(encoded)
        return self.decode_stream(indexes, cdfs, cdfs_sizes, offsets, as_tensor=as_tensor)
