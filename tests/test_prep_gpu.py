"""N3 (GPU): prep_data_for_PR on the device (histogram mode, centre / pad / crop, clip) must be bit-identical to the
ORACLE's restatement of pyotf/utils.py:76-201 (oracle/pyotf_oracle.py: remove_bg, center_data and the no-resize branch
are pinned to the live reference in tests/test_oracle_vs_reference.py; the pad / crop branches depend on the un-vendored
dphtools fft_pad / slice_maker and are unpinned), and retrieve_phase must accept its output directly.  The product's own
host functions (pyotf_b200.utils) are held to the same oracle."""
import os

import numpy as np
import pytest

from conftest import GOLDEN, rel_l2

pytestmark = pytest.mark.gpu


def test_fixture_stack_bit_identical_and_feeds_retrieve_phase():
    from oracle import pyotf_oracle as orc
    from pyotf_b200.phaseretrieval import retrieve_phase
    from pyotf_b200.utils import prep_data_for_PR, prep_data_for_PR_dev

    raw = np.load(os.path.join(GOLDEN, "fixture_psf_520.npz"))["raw"]
    host = orc.prep_data_for_pr(raw, 256, 1.1)
    np.testing.assert_array_equal(prep_data_for_PR(raw, 256, 1.1), host)
    dev = prep_data_for_PR_dev(raw, 256, 1.1, precision="fp64")
    assert dev.shape == (25, 256, 256)
    np.testing.assert_array_equal(dev.cpu().numpy(), host)
    assert abs(host.sum() - 22370068.4) < 0.1
    g = np.load(os.path.join(GOLDEN, "cfg2_phase_retrieval.npz"))
    r = retrieve_phase(dev, dict(wl=520, na=0.85, ni=1.0, res=130, zres=300), 100, 1e-4, 1e-4, precision="fp64")
    assert len(r.mse) == 51 and rel_l2(r.mse, g["es_mse"]) <= 1e-11
    d32 = prep_data_for_PR_dev(raw, 256, 1.1, precision="fp32")
    np.testing.assert_array_equal(d32.cpu().numpy(), host.astype(np.float32))


@pytest.mark.parametrize("shape,xysize", [((10, 5, 5), None), ((10, 5, 4), None), ((10, 5, 4), 10), ((10, 5, 4), 3),
                                          ((7, 64, 48), 32), ((7, 64, 48), 100), ((9, 33, 33), 33), ((3, 40, 40), 17),
                                          ((6, 31, 57), 20)])
def test_pad_crop_centre_bit_identical_unpinned_dphtools_branches(shape, xysize):
    from oracle import pyotf_oracle as orc
    from pyotf_b200.utils import prep_data_for_PR_dev
    from pyotf_b200.utils import prep_data_for_PR as product_host

    prep_data_for_PR = orc.prep_data_for_pr
    rng = np.random.default_rng(sum(shape))
    a = (rng.poisson(200, shape) + 100).astype(np.uint16)
    a[tuple(rng.integers(0, s) for s in shape)] = 60000                # a unique maximum to centre on
    host = prep_data_for_PR(a, xysize, 1.05)
    np.testing.assert_array_equal(product_host(a, xysize, 1.05), host)
    dev = prep_data_for_PR_dev(a, xysize, 1.05, precision="fp64").cpu().numpy()
    assert dev.shape == host.shape
    np.testing.assert_array_equal(dev, host)
    # ties of the maximum resolve to the first occurrence like numpy.argmax
    b = a.copy()
    b[b == 60000] = 500
    b.flat[[5, b.size - 3]] = 50000
    np.testing.assert_array_equal(prep_data_for_PR_dev(b, xysize, 1.05, precision="fp64").cpu().numpy(),
                                  prep_data_for_PR(b, xysize, 1.05))


def test_rejects_other_dtypes():
    from pyotf_b200.utils import prep_data_for_PR_dev

    with pytest.raises(ValueError):
        prep_data_for_PR_dev(np.ones((3, 8, 8)), 8)
