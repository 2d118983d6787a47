"""CPU: the bitstream path.  The oracle restatement of the range coder and the
product's C coder (host entry points of libclc_b200.so) against the committed fixtures produced by the
reference's own compress()/decompress() run through the shim (tests/golden/coder.npz), bit for bit."""
import random

import numpy as np
import pytest
import torch

from conftest import load_golden


def _oracle_ans():
    import oracle
    oracle.enable_shim()
    from compressai import ans
    return ans


def _raw(g):
    return (g["raw_sym"].tolist(), g["raw_idx"].tolist(), g["raw_cdfs"].tolist(), g["raw_sizes"].tolist(),
            g["raw_offsets"].tolist(), g["raw_string"].numpy().tobytes())


def test_oracle_coder_vs_golden_and_round_trip():
    O = _oracle_ans()
    g = load_golden("coder.npz")
    sym, idx, cdfs, sizes, offsets, want = _raw(g)
    assert O.RansEncoder().encode_with_indexes(sym, idx, cdfs, sizes, offsets) == want
    assert O.RansDecoder().decode_with_indexes(want, idx, cdfs, sizes, offsets) == sym
    # the y stream of the reference's CLC.compress decodes to the symbols the reference handed to the coder
    tab = (g["gc_cdf"].tolist(), g["gc_cdf_length"].tolist(), g["gc_offset"].tolist())
    n = 4096                                              # prefix: the pure-Python coder is slow
    dec = O.RansDecoder()
    dec.set_stream(g["y_string"].numpy().tobytes())
    assert dec.decode_stream(g["indexes"][:n].tolist(), *tab) == g["symbols"][:n].tolist()


def test_c_coder_bit_exact_vs_golden():
    from clc_b200 import ans as A
    g = load_golden("coder.npz")
    sym, idx, cdfs, sizes, offsets, want = _raw(g)
    assert A.RansEncoder().encode_with_indexes(sym, idx, cdfs, sizes, offsets) == want
    assert A.RansDecoder().decode_with_indexes(want, idx, cdfs, sizes, offsets) == sym
    # tensors in, tensor out; streaming decode in pieces; buffered encoder fed in pieces
    ts, ti = torch.tensor(sym, dtype=torch.int32), torch.tensor(idx, dtype=torch.int32)
    assert A.RansEncoder().encode_with_indexes(ts, ti, g["raw_cdfs"], g["raw_sizes"], g["raw_offsets"]) == want
    d = A.RansDecoder()
    d.set_stream(want)
    a = d.decode_stream(ti[:1000], cdfs, sizes, offsets, as_tensor=True)
    b = d.decode_stream(ti[1000:], cdfs, sizes, offsets, as_tensor=True)
    assert torch.equal(torch.cat([a, b]), ts)
    e = A.BufferedRansEncoder()
    e.encode_with_indexes(sym[:7], idx[:7], cdfs, sizes, offsets)
    e.encode_with_indexes(sym[7:], idx[7:], cdfs, sizes, offsets)
    assert e.flush() == want
    # the whole y stream of the reference's CLC.compress (81 920 symbols, 64 scale-table CDFs)
    tab = (g["gc_cdf"], g["gc_cdf_length"], g["gc_offset"])
    y_string = g["y_string"].numpy().tobytes()
    assert A.RansEncoder().encode_with_indexes(g["symbols"].int(), g["indexes"].int(), *tab) == y_string
    out = A.RansDecoder().decode_with_indexes(y_string, g["indexes"].int(), *tab, as_tensor=True)
    assert torch.equal(out, g["symbols"].int())


def test_c_coder_edge_cases():
    from clc_b200 import ans as A
    cdfs, sizes, offsets = [[0, 40000, 65535, 65536]], [4], [-1]
    assert A.RansDecoder().decode_with_indexes(A.RansEncoder().encode_with_indexes([], [], cdfs, sizes, offsets),
                                               [], cdfs, sizes, offsets) == []
    assert len(A.BufferedRansEncoder().flush()) == 8          # an empty stream is the 64-bit initial state
    for sym in ([-1], [0], [1], [2], [-2], [2 ** 31 - 1], [-2 ** 31 + 1], [5, -7, 0, 123456, -1, -1]):
        s = A.RansEncoder().encode_with_indexes(sym, [0] * len(sym), cdfs, sizes, offsets)
        assert A.RansDecoder().decode_with_indexes(s, [0] * len(sym), cdfs, sizes, offsets) == sym
    with pytest.raises(RuntimeError, match="invalid argument"):
        A.RansEncoder().encode_with_indexes([0], [3], cdfs, sizes, offsets)           # index out of range
    with pytest.raises(RuntimeError, match="invalid argument"):
        A.RansDecoder().decode_with_indexes(b"\x00" * 8, [0] * 64, cdfs, sizes, offsets)   # truncated stream
    with pytest.raises(ValueError):
        A.RansDecoder().decode_stream([0], cdfs, sizes, offsets)                       # no stream set


def test_pmf_to_quantized_cdf_vs_oracle():
    O = _oracle_ans()
    from clc_b200 import ans as A
    rnd = random.Random(3)
    for trial in range(200):
        n = rnd.randint(1, 80)
        pmf = [rnd.random() ** rnd.choice([1, 3, 8]) * rnd.choice([1, 1e-3, 1e-7]) for _ in range(n)]
        tot = sum(pmf)
        pmf = [p / tot for p in pmf]
        got = A.pmf_to_quantized_cdf(pmf)
        assert got == O.pmf_to_quantized_cdf(pmf), trial
        assert got[0] == 0 and got[-1] == 65536 and all(b > a for a, b in zip(got, got[1:]))
    with pytest.raises(RuntimeError):
        A.pmf_to_quantized_cdf([0.5, float("nan")])
    with pytest.raises(RuntimeError):
        A.pmf_to_quantized_cdf([0.0, 0.0])


def test_update_tables_equal_reference_tables():
    """CLC.update() (CLC_run.py:486-491): the tables the drop-in modules build are the tables the
    reference built (through the shim) for the same detfill weights."""
    from clc_b200.models import CLC
    from oracle import detfill
    g = load_golden("coder.npz")
    m = detfill.fill_(CLC(N=64), seed=0).eval()
    assert m.update() is True
    assert m.update() is False                      # already initialised, force=False
    gc, eb = m.gaussian_conditional, m.entropy_bottleneck
    for mod, pre in ((gc, "gc"), (eb, "eb")):
        assert torch.equal(mod.quantized_cdf.cpu(), g[f"{pre}_cdf"].int())
        assert torch.equal(mod.cdf_length.cpu(), g[f"{pre}_cdf_length"].int())
        assert torch.equal(mod.offset.cpu(), g[f"{pre}_offset"].int())
    # the z stream of the reference decodes with these tables; dequantised values are integers + medians
    z_hat = eb.decompress([g["z_string"].numpy().tobytes()], tuple(int(v) for v in g["shape"]))
    assert z_hat.shape == (1, 192, 4, 4)
    med = eb._get_medians().reshape(1, -1, 1, 1)
    assert torch.equal(torch.round(z_hat - med) + med, z_hat)
    assert eb.compress(z_hat) == [g["z_string"].numpy().tobytes()]


def test_uninitialised_tables_raise_like_the_reference():
    import clc_b200
    eb = clc_b200.EntropyBottleneck(4)
    with pytest.raises(ValueError, match="Uninitialized CDFs"):
        eb.compress(torch.zeros(1, 4, 2, 2))
    gc = clc_b200.GaussianConditional(None)
    with pytest.raises(ValueError, match="Uninitialized CDFs"):
        gc.coder_tables()


def test_reference_itself_reproduces_coder_golden():
    from oracle import ref_loader
    if not ref_loader.available():
        pytest.skip("/root/reference not present (GPU box)")
    from oracle import detfill
    CLC, _ = ref_loader.import_models()
    g = load_golden("coder.npz")
    torch.set_num_threads(8)
    m = detfill.fill_(CLC(N=64).eval(), seed=0)
    m.update()
    x = detfill.det_image((1, 3, 256, 256), 11)
    refs = [detfill.det_image((1, 3, 256, 256), 12 + i) for i in range(3)]
    with torch.no_grad():
        out = m.compress(x, refs)
    assert out["strings"][0][0] == g["y_string"].numpy().tobytes()
    assert out["strings"][1][0] == g["z_string"].numpy().tobytes()


def test_full_size_round_trip_clic_stream():
    """BASELINE cfg4 size: the y stream of one 2048x1280 image (3 276 800 symbols over the 64 scale-table
    CDFs, heavy tails so that bypass coding occurs) -- encode -> decode is the identity, and a single
    flipped payload word is detected as a different symbol sequence."""
    import clc_b200
    from clc_b200 import ans as A
    from clc_b200.models import get_scale_table
    gc = clc_b200.GaussianConditional(None)
    gc.update_scale_table(get_scale_table())
    tab = gc.coder_tables()
    rng = np.random.default_rng(7)
    n = 3_276_800
    idx = rng.integers(0, 64, n).astype(np.int32)
    scale = get_scale_table().numpy()[idx]
    sym = np.rint(rng.standard_t(3, n) * scale).clip(-2 ** 20, 2 ** 20).astype(np.int32)
    s = A.RansEncoder().encode_with_indexes(sym, idx, tab, None, None)
    out = A.RansDecoder().decode_with_indexes(s, idx, tab, None, None, as_tensor=True).numpy()
    assert (out == sym).all()
    assert 0.5 * n < len(s) < 4 * n                       # ~1 byte per symbol for this mix
    # decoding is resumable at any split point (the per-slice decode_stream of decompress())
    d = A.RansDecoder()
    d.set_stream(s)
    parts = [d.decode_stream(idx[a:b], tab, None, None, as_tensor=True).numpy()
             for a, b in ((0, 1), (1, 655_360), (655_360, n))]
    assert (np.concatenate(parts) == sym).all()
    bad = bytearray(s)
    bad[len(bad) // 2] ^= 0x40
    try:
        out2 = A.RansDecoder().decode_with_indexes(bytes(bad), idx, tab, None, None, as_tensor=True).numpy()
        assert (out2 != sym).any()
    except RuntimeError:
        pass                                              # a corrupted stream may also run out of words


def test_coder_property_random_tables_vs_oracle():
    """Property test (hypothesis): for random quantised tables, offsets and symbol sequences -- including
    values far outside the table support on both sides -- the C coder's stream equals the oracle's byte for
    byte and decodes back to the input, whole or in two resumed pieces."""
    from hypothesis import given, settings, strategies as st
    O = _oracle_ans()
    from clc_b200 import ans as A

    @st.composite
    def case(draw):
        n_cdfs = draw(st.integers(1, 5))
        cdfs, sizes, offsets = [], [], []
        for _ in range(n_cdfs):
            m = draw(st.integers(1, 24))
            w = draw(st.lists(st.floats(1e-6, 1.0), min_size=m, max_size=m))
            tot = sum(w)
            c = O.pmf_to_quantized_cdf([x / tot for x in w] + [1e-9], 16)
            cdfs.append(c + [0] * (26 - len(c)))
            sizes.append(len(c))
            offsets.append(draw(st.integers(-40, 10)))
        n = draw(st.integers(0, 200))
        idx = draw(st.lists(st.integers(0, n_cdfs - 1), min_size=n, max_size=n))
        sym = [draw(st.one_of(st.integers(offsets[i] - 3, offsets[i] + sizes[i] + 1),
                              st.integers(-2 ** 31 + 1, 2 ** 31 - 1))) for i in idx]
        cut = draw(st.integers(0, n))
        return cdfs, sizes, offsets, idx, sym, cut

    @settings(max_examples=150, deadline=None)
    @given(case())
    def run(c):
        cdfs, sizes, offsets, idx, sym, cut = c
        want = O.RansEncoder().encode_with_indexes(sym, idx, cdfs, sizes, offsets)
        got = A.RansEncoder().encode_with_indexes(sym, idx, cdfs, sizes, offsets)
        assert got == want
        assert A.RansDecoder().decode_with_indexes(got, idx, cdfs, sizes, offsets) == sym
        d = A.RansDecoder()
        d.set_stream(got)
        assert d.decode_stream(idx[:cut], cdfs, sizes, offsets) + d.decode_stream(idx[cut:], cdfs, sizes, offsets) == sym
        assert O.RansDecoder().decode_with_indexes(got, idx, cdfs, sizes, offsets) == sym

    run()


def test_quantized_cdf_invariants_on_adversarial_pmfs():
    """clc_pmf_to_quantized_cdf against the INVARIANTS any 16-bit rANS table must satisfy (an authority that does
    not depend on the in-tree shim): cdf[0] = 0, cdf[-1] = 2^16, strictly increasing (every symbol keeps a
    non-zero frequency, also when its mass rounds to 0), and -- where no stealing was needed -- frequencies equal
    round(p * 2^16).  Adversarial inputs: masses straddling the 1/65536 boundary, one dominant symbol, many
    zero-mass symbols, EntropyBottleneck tables from perturbed parameters."""
    from hypothesis import given, settings, strategies as st
    from clc_b200.ans import pmf_to_quantized_cdf
    O = _oracle_ans()
    unit = 1.0 / 65536

    @st.composite
    def pmf(draw):
        kind = draw(st.integers(0, 3))
        m = draw(st.integers(1, 60))
        if kind == 0:      # masses around the quantisation step
            w = [draw(st.sampled_from([0.0, 0.49 * unit, 0.5 * unit, 0.51 * unit, unit, 1.5 * unit])) for _ in range(m)]
            w.append(max(1.0 - sum(w), unit))
        elif kind == 1:    # one dominant symbol, the rest tiny
            w = [draw(st.floats(0.0, 3e-6)) for _ in range(m)] + [1.0]
        elif kind == 2:    # smooth bell (what update() produces)
            s = draw(st.floats(0.3, 12.0))
            c = draw(st.floats(0, m))
            w = [np.exp(-0.5 * ((i - c) / s) ** 2) + 1e-12 for i in range(m + 1)]
        else:
            w = [draw(st.floats(1e-9, 1.0)) for _ in range(m + 1)]
        tot = sum(w)
        return [x / tot for x in w]

    @settings(max_examples=300, deadline=None)
    @given(pmf())
    def run(p):
        cdf = pmf_to_quantized_cdf(p, 16)
        assert len(cdf) == len(p) + 1 and cdf[0] == 0 and cdf[-1] == 65536
        freq = np.diff(np.asarray(cdf, dtype=np.int64))
        assert (freq >= 1).all(), "every symbol must stay decodable"
        assert cdf == O.pmf_to_quantized_cdf(p, 16)
        want = np.floor(np.asarray(p, dtype=np.float32).astype(np.float64) * 65536 + 0.5)
        if int(want.sum()) == 65536 and (want >= 1).all():   # no renormalisation, nothing stolen
            assert (freq == want).all()

    run()


def test_eb_tables_from_perturbed_parameters_round_trip():
    """EntropyBottleneck.update() tables built from perturbed parameters: quantised CDF rows are valid rANS
    tables, equal the oracle's update(), and symbols over (and beyond) the support survive encode -> decode."""
    import clc_b200
    from clc_b200 import ans as A
    from oracle import clc_oracle as O
    g = torch.Generator().manual_seed(3)
    eb = clc_b200.EntropyBottleneck(12)
    with torch.no_grad():
        for n, p in eb.named_parameters():
            if n == "quantiles":
                p[:, 0, 0] -= 4 * torch.rand(12, generator=g)
                p[:, 0, 1] += torch.randn(12, generator=g)
                p[:, 0, 2] += 4 * torch.rand(12, generator=g)
            else:
                p.add_(0.2 * torch.randn(p.shape, generator=g))
    eb.update(force=True)
    ob = O.EntropyBottleneck(12)
    ob.load_state_dict({k: v for k, v in eb.state_dict().items() if not k.startswith("_") or "matrix" in k
                        or "bias" in k or "factor" in k}, strict=False)
    ob.update(force=True)
    assert torch.equal(eb._quantized_cdf.cpu(), ob._quantized_cdf.cpu())
    assert torch.equal(eb._offset.cpu(), ob._offset.cpu()) and torch.equal(eb._cdf_length.cpu(), ob._cdf_length.cpu())
    tab = eb.coder_tables()
    cdf, ln = eb._quantized_cdf.numpy(), eb._cdf_length.numpy()
    for c in range(12):
        row = cdf[c, :ln[c]]
        assert row[0] == 0 and row[-1] == 65536 and (np.diff(row) >= 1).all()
    rng = np.random.default_rng(5)
    n = 20000
    idx = rng.integers(0, 12, n).astype(np.int32)
    sym = (eb._offset.numpy()[idx] + rng.integers(-3, ln[idx] + 1)).astype(np.int32)   # incl. out-of-support (bypass)
    s = A.RansEncoder().encode_with_indexes(sym, idx, tab, None, None)
    out = A.RansDecoder().decode_with_indexes(s, idx, tab, None, None, as_tensor=True).numpy()
    assert (out == sym).all()


def test_buffered_encoder_resolves_each_call_against_its_own_tables():
    """compressai's BufferedRansEncoder codes every call's symbols with the cdfs passed in THAT call; buffering
    calls with different tables must give the stream of the oracle coder fed the same calls."""
    O = _oracle_ans()
    from clc_b200 import ans as A
    t1 = ([O.pmf_to_quantized_cdf([0.2, 0.5, 0.3, 1e-9], 16)], [5], [-1])
    c2 = [O.pmf_to_quantized_cdf([0.1, 0.1, 0.6, 0.2, 1e-9], 16), O.pmf_to_quantized_cdf([0.7, 0.3, 1e-9], 16) + [0, 0]]
    t2 = (c2, [6, 4], [0, -3])
    s1, i1 = [0, 1, -1, 1, 0, 5], [0, 0, 0, 0, 0, 0]
    s2, i2 = [2, -3, 1, -2, 3, 0, 9], [0, 1, 0, 1, 0, 0, 1]
    want_e = O.BufferedRansEncoder()
    got_e = A.BufferedRansEncoder()
    for sy, ix, t in ((s1, i1, t1), (s2, i2, t2), (s1, i1, t1)):
        want_e.encode_with_indexes(sy, ix, *t)
        got_e.encode_with_indexes(sy, ix, *t)
    want, got = want_e.flush(), got_e.flush()
    assert got == want
    d = A.RansDecoder()
    d.set_stream(got)
    assert d.decode_stream(i1, *t1) == s1 and d.decode_stream(i2, *t2) == s2 and d.decode_stream(i1, *t1) == s1
