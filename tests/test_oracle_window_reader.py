"""The reader of RepeatResolver.c (Einlesen 293-429) as rr_msa_read_window / rr.Einlesen(path, von, bis): the window's columns of the
reads with a symbol at both ends of the window, Ausgelassen per line; bis follows shorter lines down.
  * against the row rule as the other oracle tests restate it (window_codes) on the golden MSAs;
  * against the UNMODIFIED reader itself (oracle/_ref/ref_window_driver dumps Ausgelassen and every group and coverage bitset)
    on texts with lower case, '_', unknown symbols at the window ends, blank ends, and a shorter line in the middle."""
import os
import subprocess

import numpy as np
import pytest

import oracle_lib as O
import repeatresolver_b200 as rr
from conftest import ROOT, golden_msa
from test_oracle_cliquer import window_codes

DRV = os.path.join(ROOT, "oracle", "_ref", "ref_window_driver")
TABLE = np.full(256, 5, dtype=np.uint8)
for _ch, _k in ((b"aA", 0), (b"cC", 1), (b"gG", 2), (b"tT", 3), (b"-_", 4)):
    for _c in _ch:
        TABLE[_c] = _k


@pytest.mark.parametrize("name,frac", [("tree_small", (0.1, 0.9)), ("distributed_small", (0.2, 0.8)), ("saturated", (0.0, 1.0)),
                                       ("initial_aligner_style", (0.25, 0.6))])
def test_window_reader_on_golden_msas(name, frac, tmp_path):
    text = golden_msa(name)
    if not text.endswith(b"\n"):
        text += b"\n"
    width = len(text.split(b"\n")[0])
    von, bis = int(frac[0] * (width - 1)), int(frac[1] * (width - 1))
    p = tmp_path / "M"
    p.write_bytes(text)
    msa, aus = rr.Einlesen(str(p), von, bis)
    codes = window_codes(text, von, bis)
    assert np.array_equal(TABLE[msa.cells()], codes)
    lines = text.split(b"\n")[:-1]
    assert list(aus) == [1 if (l[von:von + 1] != b" " and l[bis:bis + 1] != b" ") else -1 for l in lines]
    msa.close()


def test_window_reader_errors(tmp_path):
    p = tmp_path / "M"
    p.write_bytes(b"ACGT\nACGT")                                             # 326: the reference exits on a last line without newline
    with pytest.raises(rr.RRError):
        rr.Einlesen(str(p), 0, 3)
    p.write_bytes(b"ACGT\nAC\n")
    with pytest.raises(rr.RRError):
        rr.Einlesen(str(p), 2, 3)                                            # the second line does not reach column 2
    with pytest.raises(rr.RRError):
        rr.Einlesen(str(tmp_path / "missing"), 0, 3)                         # "MA is missing."
    with pytest.raises(rr.RRError):
        rr.Einlesen(str(p), 3, 1)
    p.write_bytes(b"")
    msa, aus = rr.Einlesen(str(p), 0, 5)
    assert (msa.rows, msa.cols, len(aus)) == (0, 0, 0)


def _quirky_text():
    rng = np.random.default_rng(4)
    rows, width = 90, 140
    sym = np.frombuffer(b"ACGT-acgt_N.", dtype=np.uint8)
    m = sym[rng.integers(0, len(sym), (rows, width))]
    for r in range(rows):                                                    # blank ends of different lengths
        a, b = int(rng.integers(0, 40)), int(rng.integers(0, 40))
        m[r, :a] = ord(" ")
        m[r, width - b:] = ord(" ")
    lines = [bytes(l) for l in m]
    lines[55] = lines[55][:118]                                              # a shorter line: bis drops to 117 from here on (328)
    return b"\n".join(lines) + b"\n"


@pytest.mark.skipif(not os.path.exists(DRV), reason="oracle/_ref is built in the build container only")
@pytest.mark.parametrize("von,bis", [(30, 100), (0, 139), (20, 125), (39, 117), (60, 60)])
def test_window_reader_against_the_unmodified_reference(von, bis, tmp_path):
    text = _quirky_text()
    p = tmp_path / "M"
    p.write_bytes(text)
    out = subprocess.run([DRV, str(p), str(von), str(bis)], capture_output=True, text=True)
    assert out.returncode == 0, out.stderr
    lines = out.stdout.splitlines()
    k = next(i for i, l in enumerate(lines) if l.startswith("A ") or l == "A") - 1
    R, N, sc, nlines = (int(x) for x in lines[k].split())
    ref_aus = [int(x) for x in lines[k + 1].split()[1:]]
    groups = [np.array([int(w, 16) for w in l.split()[1:]], dtype=np.uint64) for l in lines[k + 2:k + 2 + 5 * N]]
    cover = [np.array([int(w, 16) for w in l.split()[1:]], dtype=np.uint64) for l in lines[k + 2 + 5 * N:k + 2 + 6 * N]]
    msa, aus = rr.Einlesen(str(p), von, bis)
    assert (msa.rows, msa.cols, len(aus)) == (R, N, nlines) and list(aus) == ref_aus
    codes = TABLE[msa.cells()]
    for i in range(N):
        for g in range(5):
            assert np.array_equal(O.bitset_words(codes[:, i] == g), groups[5 * i + g]), (i, g)
        assert np.array_equal(O.bitset_words(codes[:, i] < 5), cover[i]), i
    assert 0 < R <= nlines and (R < nlines or von >= 40)
    msa.close()
