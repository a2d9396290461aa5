"""CPU checks of the product boundary: the C-ABI library loads and exports every symbol the header
declares, the ctypes table binds them all, the product never touches oracle/, and without a GPU the
public API fails loudly instead of falling back."""
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _header_symbols():
    text = open(os.path.join(ROOT, "include", "pyotf_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(pyotf_b200_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    from pyotf_b200 import _lib
    from pyotf_b200 import build as _build

    _build.build()
    lib = _lib.load()
    syms = _header_symbols()
    assert len(syms) >= 29
    for name in syms:
        assert hasattr(lib, name), f"{name} declared in include/pyotf_b200.h but not exported"
    assert sorted(_lib.PROTOTYPES) == syms, "ctypes prototypes and header declarations differ"
    assert lib.pyotf_b200_abi_version() == 1
    assert lib.pyotf_b200_last_error() is not None


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "pyotf_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", text, flags=re.M), f
                assert "pyotf_oracle" not in text and "refshim" not in text, f
                assert "/root/reference" not in text, f


def test_no_cpu_fallback_without_gpu():
    import torch

    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from pyotf_b200 import _lib
    from pyotf_b200.otf import HanserPSF, SheppardPSF
    from pyotf_b200.phaseretrieval import retrieve_phase

    m = HanserPSF(520, 0.85, 1.0, 130, 64, 300, 4)
    for attr in ("PSFi", "PSFa", "OTFi", "OTFa"):
        with pytest.raises(_lib.Pyotf_b200Error):
            getattr(m, attr)
    with pytest.raises(_lib.Pyotf_b200Error):
        SheppardPSF(520, 1.27, 1.33, 100, 32, 190, 24).PSFi
    with pytest.raises(_lib.Pyotf_b200Error):
        retrieve_phase(np.ones((3, 64, 64)), dict(wl=520, na=0.85, ni=1.0, res=130, zres=300), 3)


def test_parameter_validation_matches_reference_semantics():
    """BasePSF / NumericProperty behaviour (pyotf/otf.py:53-212, utils.py:32-73) needs no device."""
    from pyotf_b200.otf import HanserPSF, SheppardPSF

    m = HanserPSF(520, 0.85, 1.0, 130, 64, 300, 5)
    np.testing.assert_array_equal(m.zrange, (np.arange(5) - 3) * 300)
    m.zsize = 8
    assert len(m.zrange) == 8
    m.zrange = 100.0
    assert m.zrange.shape == (1,)
    with pytest.raises(ValueError):
        m.res = 520 / 0.85 / 2
    with pytest.raises(ValueError):
        m.zres = 0
    with pytest.raises(TypeError):
        m.size = 64.0
    with pytest.raises(ValueError):
        m.na = -1.0
    with pytest.raises(ValueError):
        m.vec_corr = "bogus"
    with pytest.raises(ValueError):
        m.condition = "bogus"
    with pytest.raises(ValueError):
        SheppardPSF(520, 1.27, 1.33, 100, 64, zres=200)       # BASELINE config 3 as written: >= wl/2/ni = 195.49
    s = SheppardPSF(520, 1.27, 1.33, 100, 64, zres=190)
    with pytest.raises(TypeError):
        s.dual = "yes"
    assert repr(s).startswith("SheppardPSF(wl=520") and repr(s).endswith("dual=False)")
    assert repr(m).startswith("HanserPSF(wl=520")


def test_host_unwrap_and_shard_bounds():
    from pyotf_b200._engine import shard_bounds
    from pyotf_b200.phaseretrieval import unwrap_phase

    y, x = np.mgrid[-1:1:96j, -1:1:96j]
    true = 7 * (x**2 + y**2) - 4 * x * y + 2 * x
    out = unwrap_phase(np.angle(np.exp(1j * true)))
    d = out - true
    assert abs(d - 2 * np.pi * np.round(d.mean() / (2 * np.pi))).max() < 1e-12
    with pytest.raises(ValueError):
        unwrap_phase(np.zeros(5))
    for n, w in ((25, 2), (25, 8), (128, 8), (3, 4), (8192, 8)):
        b = [shard_bounds(n, r, w) for r in range(w)]
        assert b[0][0] == 0 and b[-1][1] == n and all(b[i][1] == b[i + 1][0] for i in range(w - 1))
        sizes = [hi - lo for lo, hi in b]
        assert max(sizes) - min(sizes) <= 1
        assert sizes == [len(c) for c in np.array_split(np.arange(n), w)]


def test_host_unwrap_matches_plain_restatement_unpinned_vs_skimage():
    """The optimised host unwrap (zero-reliability edges joined as they are generated, all-zero row bands skipped,
    radix sort) against the plain all-edges / stable-sort / union-find restatement in oracle/unwrap_oracle.py:
    identical arrays on pupil-like (zero outside an aperture), dense, noisy, flat and tiny inputs."""
    from oracle import unwrap_oracle as uo
    from pyotf_b200.phaseretrieval import unwrap_phase

    rng = np.random.default_rng(5)
    y, x = np.mgrid[-1:1:40j, -1:1:47j]
    smooth = 9 * (x**2 + y**2) - 5 * x * y
    inside = x**2 + y**2 < 0.55
    cases = {
        "pupil": np.where(inside, np.angle(np.exp(1j * smooth)), 0.0),
        "noisy pupil": np.where(inside, np.angle(np.exp(1j * (smooth + 0.9 * rng.standard_normal(x.shape)))), 0.0),
        "dense": np.angle(np.exp(1j * smooth)),
        "noise": rng.uniform(-np.pi, np.pi, (23, 31)),
        "flat blocks": np.where((x > 0.3) & (y < 0.5), 1.0, 0.0) + np.where(x < -0.6, 2.5, 0.0),
        "zeros": np.zeros((9, 12)),
        "tiny": rng.uniform(-3, 3, (3, 5)),
        "one row": rng.uniform(-3, 3, (1, 7)),
        "one pixel": np.ones((1, 1)),
    }
    for name, a in cases.items():
        np.testing.assert_array_equal(unwrap_phase(a), uo.unwrap_phase_2d(a), err_msg=name)
