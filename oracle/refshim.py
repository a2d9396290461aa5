"""TEST INFRASTRUCTURE ONLY -- never imported by the product package.

Imports the *unmodified* upstream pyotf from a directory on disk (``/root/reference`` in
the build container) so that ``oracle/make_golden.py`` and the CPU validation tests can
pin ``oracle/pyotf_oracle.py`` against the reference's own numbers.

pyotf's numeric core (otf.py / phaseretrieval.py / zernike.py / utils.py) imports four
packages that are not installed here and that none of the hot-path functions reach
(matplotlib, tifffile, dphtools, skimage).  They are stubbed in ``sys.modules`` *before*
the import; the numeric code then runs unmodified.

The GPU box has no ``/root/reference``; ``load_reference()`` returns ``None`` there and
everything that needs the live reference is skipped -- parity on the GPU box is against
``oracle/pyotf_oracle.py`` and the committed ``tests/golden/*.npz`` vectors.
"""

import importlib
import os
import sys
import types
from unittest import mock

REFERENCE_ROOT = os.environ.get("PYOTF_REFERENCE_ROOT", "/root/reference")


def _stub_modules():
    stubs = {}
    for name in (
        "matplotlib",
        "matplotlib.pyplot",
        "matplotlib.font_manager",
        "matplotlib.patches",
        "matplotlib.patheffects",
        "matplotlib.colors",
        "mpl_toolkits",
        "mpl_toolkits.axes_grid1",
        "mpl_toolkits.axes_grid1.anchored_artists",
        "tifffile",
        "skimage",
    ):
        stubs[name] = mock.MagicMock(name=name)
    # skimage.restoration.unwrap_phase: identity stub (only PhaseRetrievalResult.phase uses
    # it, phaseretrieval.py:146; parity for that field is "unpinned")
    restoration = types.ModuleType("skimage.restoration")
    restoration.unwrap_phase = lambda x, *a, **k: x
    stubs["skimage.restoration"] = restoration
    dph = types.ModuleType("dphtools")
    dph_utils = types.ModuleType("dphtools.utils")

    def _unavailable(*a, **k):
        raise NotImplementedError("dphtools is not installed; stubbed by oracle/refshim.py")

    for fn in ("fft_pad", "slice_maker", "radial_profile"):
        setattr(dph_utils, fn, _unavailable)

    def bin_ndarray(ndarray, new_shape=None, bin_size=None, operation="sum"):
        """Stand-in for dphtools.utils.bin_ndarray with its default operation (block sum); the only caller on a
        tested path, WidefieldMicroscope.PSF (pyotf/microscope.py:121), normalises the result, so a block mean
        would give the same answer."""
        import numpy as np

        assert new_shape is None and operation == "sum"
        shape = []
        for n, b in zip(ndarray.shape, bin_size):
            shape += [n // b, b]
        return np.asarray(ndarray).reshape(shape).sum(axis=tuple(range(1, 2 * ndarray.ndim, 2)))

    dph_utils.bin_ndarray = bin_ndarray
    dph.utils = dph_utils
    stubs["dphtools"] = dph
    stubs["dphtools.utils"] = dph_utils
    return stubs


def reference_available():
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "pyotf"))


def load_reference():
    """Return the upstream ``pyotf`` package (numeric modules imported) or ``None``."""
    if not reference_available():
        return None
    if "pyotf" in sys.modules and getattr(sys.modules["pyotf"], "__file__", "").startswith(
        REFERENCE_ROOT
    ):
        return sys.modules["pyotf"]
    for name, mod in _stub_modules().items():
        sys.modules.setdefault(name, mod)
    sys.path.insert(0, REFERENCE_ROOT)
    try:
        pkg = importlib.import_module("pyotf")
        for sub in ("utils", "zernike", "otf", "phaseretrieval", "microscope"):
            importlib.import_module(f"pyotf.{sub}")
    finally:
        sys.path.remove(REFERENCE_ROOT)
    return pkg


def read_fixture_stack(path=None):
    """Read the 25x256x256 uint16 measured PSF shipped with the reference (PIL, page by page)."""
    import numpy as np
    from PIL import Image

    if path is None:
        path = os.path.join(
            REFERENCE_ROOT, "fixtures", "psf_wl520nm_z300nm_x130nm_na0.85_n1.0.tif"
        )
    im = Image.open(path)
    pages = []
    for i in range(getattr(im, "n_frames", 1)):
        im.seek(i)
        pages.append(np.array(im))
    return np.stack(pages)
