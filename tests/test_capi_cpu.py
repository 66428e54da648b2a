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
