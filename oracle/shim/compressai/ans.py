"""compressai.ans stub: the rANS coder is out of scope."""


class _Unavailable:
    def __init__(self, *a, **k):
        raise NotImplementedError("rANS coder is outside the hot-path scope of this oracle shim")


BufferedRansEncoder = _Unavailable
RansEncoder = _Unavailable
RansDecoder = _Unavailable
