"""CPU tests of the drop-in boundary: the C-ABI library loads and exports every symbol include/ppp_b200.h declares
(no compute calls without a GPU), and fails loudly -- never silently falls back -- when there is no device."""
import ctypes
import os
import re

import pytest

from prophylo_b200 import capi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    hdr = open(os.path.join(ROOT, "include", "ppp_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    return sorted(set(re.findall(r"\b(ppp_[a-z_0-9]+)\s*\(", hdr)))


def test_library_exports_every_declared_symbol():
    lib = capi.load()
    syms = declared_symbols()
    assert len(syms) >= 15
    for s in syms:
        assert hasattr(lib, s), s
    assert set(syms) == set(capi.EXPORTS)
    assert lib.ppp_abi_version() == 2
    assert lib.ppp_words_per_row(1024) == 32 and lib.ppp_words_per_row(1459) == 48 and lib.ppp_words_per_row(1) == 4


def test_hit_layout_is_32_bytes():
    assert capi.HIT_DTYPE.itemsize == 32
    assert capi.HIT_DTYPE.fields["score"][1] == 8 and capi.HIT_DTYPE.fields["margin"][1] == 28
    assert ctypes.sizeof(capi.Filter) == 16


def test_no_silent_cpu_fallback():
    import torch

    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(capi.PPPError) as ei:
        capi.Context(0)
    assert ei.value.code == -2 and "no CPU path" in str(ei.value)


def test_native_bla_reader_matches_python_reader(tmp_path):
    """ppp_read_bla (C++, multi-threaded, no GPU) against the Python reader + packer that reproduce bin/ppp.pl:606-673 and
    BlastResult.pm:192-237 (themselves pinned byte for byte against the reference's ppp.pl stdout in test_drivers.py):
    repeated subjects (last taxon wins), false taxa, short lines, foreign taxa, empty files, no trailing newline."""
    import numpy as np

    from prophylo_b200 import capi, drivers, packer

    rng = np.random.default_rng(5)
    universe = [str(1000 + i) for i in range(300)]
    pool = universe + ["0", "", "999999", "abc", "00", "0.0"]
    paths = []
    for i in range(40):
        rows = []
        for _ in range(int(rng.integers(0, 200))):
            s = str(700000 + int(rng.integers(0, 90)))
            kind = rng.random()
            if kind < 0.03:
                rows.append(f"q\t{s}")                      # short line: taxon missing -> false
            elif kind < 0.05:
                rows.append("")                             # blank line
            elif kind < 0.07:
                rows.append(f"q\t{s}\t1e-5\t{rng.choice(pool)}\textra\tcols")
            else:
                rows.append(f"q\t{s}\t1e-{int(rng.integers(1, 99))}\t{rng.choice(pool)}")
        text = "\n".join(rows) + ("\n" if (rows and i % 3) else "")
        p = tmp_path / f"{i}.bla"
        p.write_text(text)
        paths.append(str(p))
    index = packer.universe_index(universe)
    exp = packer.pack_ranked_targets([drivers.read_bla(p) for p in paths], index)
    for nthreads in (1, 4):
        got = capi.read_bla_files(paths, universe, nthreads=nthreads)
        for g, e, name in zip(got, exp, ("idx", "rawdepth", "row_off")):
            assert g.dtype == e.dtype and np.array_equal(g, e), name
    exp = packer.pack_ranked_targets([drivers.read_bla(p) for p in paths], index, dup=True)  # -dup packing (ppp.pm:189-200)
    got = capi.read_bla_files(paths, universe, nthreads=3, dup=True)
    for g, e, name in zip(got, exp, ("idx", "rawdepth", "row_off")):
        assert g.dtype == e.dtype and np.array_equal(g, e), name
    assert got[0].size != capi.read_bla_files(paths, universe)[0].size
    import pytest

    with pytest.raises(capi.PPPError):
        capi.read_bla_files([str(tmp_path / "missing.bla")], universe)


def test_native_profile_reader_matches_python_reader(tmp_path):
    """ppp_read_profiles (C++: parse + shared universe + word-wise bit-plane packing, no GPU) against the Python restatement of
    Profile.pm:433-514 + the numpy packer: comments, a header line, repeated taxa (a non-zero value is never overwritten),
    Perl-false taxa, trailing empty fields, non-numeric ids (ordered after the numeric ones), files with no trailing newline."""
    import numpy as np

    from prophylo_b200 import capi, drivers

    rng = np.random.default_rng(8)
    ids = [str(x) for x in rng.choice(np.arange(100, 90000), 400, replace=False)] + ["Xca", "gec5", "12abc"]
    paths = []
    for i in range(30):
        rows = ["# family %d" % i]
        for t in ids:
            if rng.random() < 0.8:
                rows.append(f"{t}\t{int(rng.random() < 0.2)}" + ("\t" if rng.random() < 0.05 else ""))
        for t in rng.choice(ids, 10):  # repeated taxa, either value
            rows.append(f"{t}\t{int(rng.random() < 0.5)}")
        rows += ["0\t1", "# end"]
        p = tmp_path / f"f{i}.profile"
        p.write_text("\n".join(rows) + ("\n" if i % 2 else ""))
        paths.append(str(p))
    for header in (False, True):
        a = drivers.read_profile_family(paths, header=header, native=True)
        b = drivers.read_profile_family(paths, header=header, native=False)
        assert a[0] == b[0] and a[0][-3:] == ["12abc", "Xca", "gec5"]
        for x, y, name in zip(a[1:], b[1:], ("P", "Y", "p", "T")):
            assert x.dtype == y.dtype and np.array_equal(x, y), (header, name)
    # families written over ONE genome list (every file the same taxa in the same order: the reader's shared-mapping path), with
    # one file that adds a taxon, one that drops one and one in another order mixed in
    same = []
    for i in range(12):
        order = list(ids)
        if i == 4:
            order.append("777777")
        if i == 7:
            order.pop(3)
        if i == 9:
            order.reverse()
        p = tmp_path / f"g{i}.profile"
        p.write_text("taxon\tvalue\n" + "".join(f"{t}\t{int(rng.random() < 0.3)}\n" for t in order))
        same.append(str(p))
    for sel in (same[:4], same):
        a = drivers.read_profile_family(sel, header=True, native=True)
        b = drivers.read_profile_family(sel, header=True, native=False)
        assert a[0] == b[0]
        for x, y, name in zip(a[1:], b[1:], ("P", "Y", "p", "T")):
            assert x.dtype == y.dtype and np.array_equal(x, y), name
    bad = tmp_path / "bad.profile"
    bad.write_text("1\t1\n\n2\t0\n")
    with pytest.raises(capi.PPPError, match="must contain only two fields"):
        capi.read_profile_files([str(bad)])
    with pytest.raises(capi.PPPError, match="Can' open"):
        capi.read_profile_files([str(tmp_path / "missing.profile")])


def test_bla_packed_cache_roundtrip_and_invalidation(tmp_path):
    """ppp_read_bla_cached: the first call parses and writes the cache, the second maps it back (identical arrays, no parsing),
    a changed file, universe or flag invalidates it, a corrupt cache is rebuilt."""
    import time

    import numpy as np

    from prophylo_b200 import capi

    rng = np.random.default_rng(11)
    universe = [str(1000 + i) for i in range(500)]
    paths = []
    for i in range(200):
        rows = [f"q\t{700000 + int(rng.integers(0, 400))}\t1e-{int(rng.integers(1, 99))}\t{rng.choice(universe + ['0', '99999'])}" for _ in range(int(rng.integers(0, 300)))]
        p = tmp_path / f"{i}.bla"
        p.write_text("\n".join(rows) + "\n")
        paths.append(str(p))
    cache = str(tmp_path / "genome.pppcache")
    info = {}
    t0 = time.perf_counter()
    a = capi.read_bla_files(paths, universe, cache=cache, info=info)
    t_parse = time.perf_counter() - t0
    assert info["from_cache"] is False and os.path.exists(cache)
    t0 = time.perf_counter()
    b = capi.read_bla_files(paths, universe, cache=cache, info=info)
    t_cache = time.perf_counter() - t0
    assert info["from_cache"] is True
    plain = capi.read_bla_files(paths, universe)
    for x, y, z in zip(a, b, plain):
        assert np.array_equal(x, y) and np.array_equal(x, z)
    assert t_cache < t_parse  # reading one packed file beats parsing 200 text files
    capi.read_bla_files(paths, universe, dup=True, cache=cache, info=info)  # other flags: other key
    assert info["from_cache"] is False
    capi.read_bla_files(paths, universe[:-1], cache=cache, info=info)       # other universe
    assert info["from_cache"] is False
    capi.read_bla_files(paths, universe, cache=cache, info=info)
    assert info["from_cache"] is False  # the cache now holds the previous call's key
    capi.read_bla_files(paths, universe, cache=cache, info=info)
    assert info["from_cache"] is True
    with open(paths[3], "a") as fh:  # a changed file (size / mtime) invalidates
        fh.write("q\t1\t1e-5\t1001\n")
    c = capi.read_bla_files(paths, universe, cache=cache, info=info)
    assert info["from_cache"] is False and c[0].size >= a[0].size
    open(cache, "wb").write(b"garbage")
    d = capi.read_bla_files(paths, universe, cache=cache, info=info)
    assert info["from_cache"] is False and np.array_equal(c[0], d[0])
