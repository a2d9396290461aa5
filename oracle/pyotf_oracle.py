"""TEST INFRASTRUCTURE ONLY -- CPU restatement (numpy, float64) of pyotf's hot path.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline legs may import
this module, and only as the *checker* / the timed CPU arm.  The product package
``pyotf_b200`` never imports it (tests/test_no_oracle_in_product.py enforces that).

What it restates (all citations relative to the upstream repo david-hoffman/pyotf):

* ``hanser_*``     -- HanserPSF._gen_kr/_gen_pupil/_calc_defocus/_gen_psf   pyotf/otf.py:335-428
* ``psfi / otfi``  -- BasePSF.PSFi / OTFi                                    pyotf/otf.py:204-212
* ``easy_fft``...  -- shifted FFT helpers, psqrt, cart2pol                    pyotf/utils.py:15-29,150-158
* ``retrieve_phase_loop`` -- the iteration of retrieve_phase                 pyotf/phaseretrieval.py:79-135
* ``zernike_modes / noll_to_degrees``                                        pyotf/zernike.py:173-220,251-365
* ``aberrated_pupil``  -- apply_aberration's pupil                           pyotf/otf.py:626-642
* ``sheppard_*``   -- SheppardPSF._gen_kr/_gen_otf/PSFa                      pyotf/otf.py:497-592

Third-party arithmetic the reference leans on (not under /root/reference, both unpinned in
requirements.txt): ``numpy.fft`` (pocketfft; numpy 2.3.5 here) and
``scipy.special.eval_jacobi`` (scipy 1.18.1 here).  The FFTs are used as-is (they are the
reference's own FFT); the Jacobi polynomial is restated with the published three-term
recurrence so that the oracle has no scipy dependency on the GPU box.

Parity pin: ``tests/test_oracle_vs_reference.py`` runs every function here against the live
upstream import (``oracle/refshim.py``) when ``/root/reference`` exists, and
``tests/test_oracle_golden.py`` checks it against ``tests/golden/*.npz`` which were generated
from the live upstream import by ``oracle/make_golden.py`` (committed).  Parity is PINNED
for everything except ``PhaseRetrievalResult.phase`` (needs skimage.unwrap_phase, absent).
"""

import numpy as np
from numpy.fft import fftfreq, fftn, fftshift, ifftn, ifftshift

# ----------------------------------------------------------------------------------------
# utils.py
# ----------------------------------------------------------------------------------------


def easy_fft(data, axes=None):
    """Centered forward FFT: shift origin to corner, transform, shift back (utils.py:15-17)."""
    return fftshift(fftn(ifftshift(data, axes=axes), axes=axes), axes=axes)


def easy_ifft(data, axes=None):
    """Centered inverse FFT (utils.py:20-22)."""
    return ifftshift(ifftn(fftshift(data, axes=axes), axes=axes), axes=axes)


def cart2pol(y, x):
    """(rho, theta) from cartesian (utils.py:25-29)."""
    return np.hypot(y, x), np.arctan2(y, x)


def psqrt(data):
    """sqrt(max(0, x)) (utils.py:150-158)."""
    return np.sqrt(np.fmax(0, data))


# ----------------------------------------------------------------------------------------
# HanserPSF (otf.py:265-443)
# ----------------------------------------------------------------------------------------


def default_zrange(zsize, zres):
    """z = (arange(nz) - (nz+1)//2) * zres (otf.py:296-298)."""
    return (np.arange(zsize) - (zsize + 1) // 2) * zres


def hanser_kspace(wl, ni, res, size):
    """Unshifted pupil-plane coordinates (otf.py:335-344).

    Returns (k, kr, phi, kmag, kz) with kr[i, j] = hypot(k[i], k[j]),
    phi[i, j] = arctan2(k[i], k[j]) -- row index is y.
    """
    k = fftfreq(size, res)
    kxx, kyy = np.meshgrid(k, k)
    kr, phi = cart2pol(kyy, kxx)
    kmag = ni / wl
    kz = psqrt(kmag**2 - kr**2)
    return k, kr, phi, kmag, kz


def hanser_pupil(kr, na, wl):
    """Ideal pupil: 1 inside the coherent cut-off na/wl (otf.py:346-355)."""
    return (kr <= na / wl).astype(complex)


def hanser_defocus(kz, zrange):
    """exp(2 pi i kz z) per plane (otf.py:357-360)."""
    zrange = np.atleast_1d(np.asarray(zrange))
    return np.exp(2 * np.pi * 1j * kz * zrange[:, None, None])


def hanser_vector_factors(theta, phi, vec_corr):
    """Polarisation projection factors, in the reference's order (otf.py:405-415)."""
    plist = []
    if vec_corr in ("z", "total"):
        plist.append(np.sin(theta) * np.cos(phi))
        plist.append(np.sin(theta) * np.sin(phi))
    if vec_corr in ("y", "total"):
        plist.append((np.cos(theta) - 1) * np.sin(phi) * np.cos(phi))
        plist.append(np.cos(theta) * np.sin(phi) ** 2 + np.cos(phi) ** 2)
    if vec_corr in ("x", "total"):
        plist.append(np.cos(theta) * np.cos(phi) ** 2 + np.sin(phi) ** 2)
        plist.append((np.cos(theta) - 1) * np.sin(phi) * np.cos(phi))
    return plist


def hanser_psfa(
    wl, na, ni, res, size, zrange, vec_corr="none", condition="sine", pupil_base=None
):
    """Amplitude PSF (nvec, nz, N, N), spatially shifted (otf.py:362-428)."""
    _, kr, phi, kmag, kz = hanser_kspace(wl, ni, res, size)
    if pupil_base is None:
        pupil_base = hanser_pupil(kr, na, wl)
    else:
        assert pupil_base.ndim == 2
    pupil = pupil_base * hanser_defocus(kz, zrange)
    theta = np.arcsin((kr < kmag) * kr / kmag)
    if condition == "sine":
        pupil = pupil * (1.0 / np.sqrt(np.cos(theta)))
    elif condition == "herschel":
        pupil = pupil * (1.0 / np.cos(theta))
    elif condition != "none":
        raise ValueError(condition)
    if vec_corr != "none":
        pupils = pupil * np.array(hanser_vector_factors(theta, phi, vec_corr))[:, None]
    else:
        pupils = pupil[None]
    return fftshift(ifftn(pupils, axes=(2, 3)), axes=(2, 3))


def psfi(psfa):
    """Intensity PSF: sum over vector components of |PSFa|^2 (otf.py:204-207)."""
    return (abs(psfa) ** 2).sum(axis=0)


def otfi(psf_i):
    """Intensity OTF: centered 3D FFT of PSFi (otf.py:209-212)."""
    return easy_fft(psf_i)


def hanser_otfa(psfa):
    """Amplitude OTF: centered FFT over (z, y, x) of each component (otf.py:435-438)."""
    return easy_fft(psfa, axes=(1, 2, 3))


# ----------------------------------------------------------------------------------------
# zernike.py
# ----------------------------------------------------------------------------------------


def noll_to_degrees(j):
    """Noll index (1-based) -> (n, m) (zernike.py:173-220)."""
    j = np.atleast_1d(np.asarray(j, dtype=np.int64))
    if not (j > 0).all():
        raise ValueError("Invalid Noll index")
    n = np.ceil((-3 + np.sqrt(1 + 8 * j)) / 2).astype(np.int64)
    jr = j - n * (n + 1) // 2
    lo = (n % 4) < 2
    m1 = np.where(lo, jr, jr - 1)
    m2 = np.where(lo, -(jr - 1), -jr)
    m = np.where((n - m1) % 2 == 0, m1, m2)
    return n, m


def jacobi_p(k, a, x):
    """Jacobi polynomial P_k^{(a,0)}(x) via the standard three-term recurrence.

    Restates scipy.special.eval_jacobi(k, a, 0, x) as used at zernike.py:332
    (Abramowitz & Stegun 22.7.1 with beta = 0).
    """
    x = np.asarray(x, dtype=float)
    p0 = np.ones_like(x)
    if k == 0:
        return p0
    p1 = 0.5 * (a + (a + 2) * x)
    for i in range(2, k + 1):
        # (2i(i+a)(2i+a-2)) P_i = (2i+a-1)[(2i+a)(2i+a-2) x + a^2] P_{i-1}
        #                           - 2 (i+a-1)(i-1)(2i+a) P_{i-2}
        c = 2 * i + a
        a1 = 2 * i * (i + a) * (c - 2)
        a2 = (c - 1) * (c * (c - 2) * x + a * a)
        a3 = 2 * (i + a - 1) * (i - 1) * c
        p0, p1 = p1, (a2 * p1 - a3 * p0) / a1
    return p1


def zernike_mode(r, theta, n, m, norm=True):
    """One Zernike polynomial on (r, theta) (zernike.py:315-365)."""
    n, m = int(n), int(m)
    mneg = m < 0
    m = abs(m)
    if (n - m) % 2:
        return np.zeros_like(r)
    rad = np.zeros_like(r)
    valid = r <= 1.0
    if n == 0 and m == 0:
        rad[valid] = 1
    else:
        rp = r[valid]
        k = (n - m) // 2
        rad[valid] = (-1) ** k * rp**m * jacobi_p(k, m, 1 - 2 * rp**2)
    rad = rad * (np.sin(m * theta) if mneg else np.cos(m * theta))
    if norm:
        rad = rad * (np.sqrt(n + 1) if m == 0 else np.sqrt(2 * (n + 1)))
    return rad


def zernike_modes(r, theta, n, m, norm=True):
    """Stack of Zernike polynomials, squeezed like the reference (zernike.py:251-312)."""
    n, m = np.atleast_1d(n), np.atleast_1d(m)
    r = np.asarray(r, dtype=float)
    theta = np.asarray(theta, dtype=float)
    return np.array([zernike_mode(r, theta, nn, mm, norm) for nn, mm in zip(n, m)]).squeeze()


def aberrated_pupil(wl, na, ni, res, size, mcoefs, pcoefs):
    """Pupil built by apply_aberration from Noll-ordered coefficients (otf.py:618-642)."""
    if mcoefs is None:
        mcoefs = np.zeros_like(pcoefs)
    if pcoefs is None:
        pcoefs = np.zeros_like(mcoefs)
    mcoefs, pcoefs = np.asarray(mcoefs, dtype=float), np.asarray(pcoefs, dtype=float)
    assert len(mcoefs) == len(pcoefs)
    _, kr, phi, _, _ = hanser_kspace(wl, ni, res, size)
    r = kr * wl / na
    zerns = zernike_modes(r, phi, *noll_to_degrees(np.arange(len(mcoefs)) + 1))
    zerns = zerns.reshape((len(mcoefs),) + r.shape)
    phase = (zerns * pcoefs[:, None, None]).sum(0)
    mag = (zerns * mcoefs[:, None, None]).sum(0)
    mag = mag + abs(hanser_pupil(kr, na, wl))
    return mag * np.exp(1j * phase)


# ----------------------------------------------------------------------------------------
# phaseretrieval.py
# ----------------------------------------------------------------------------------------


def calc_mse(a, b):
    """Mean squared difference (phaseretrieval.py:401-403)."""
    return ((a - b) ** 2).mean()


def retrieve_phase_loop(
    data,
    wl,
    na,
    ni,
    res,
    zres=None,
    zrange=None,
    max_iters=200,
    pupil_tol=1e-8,
    mse_tol=1e-8,
    phase_only=False,
    return_history=False,
):
    """The Hanser 2003 iteration exactly as retrieve_phase runs it (phaseretrieval.py:79-135).

    The model is forced to vec_corr="none", condition="none" (phaseretrieval.py:74-76).
    Returns dict(pupil, mse, mse_diff, pupil_diff, psfi, iters[, history]) where ``pupil`` is
    the unshifted complex pupil after the last update and ``psfi`` the PSFi of the last
    *applied* pupil (= result.model.PSFi).
    """
    assert max_iters > 0 and data.ndim == 3 and data.shape[1] == data.shape[2]
    nz, size = data.shape[0], data.shape[-1]
    if zrange is None:
        zrange = default_zrange(nz, res if zres is None else zres)
    mag = psqrt(data)
    _, kr, _, _, kz = hanser_kspace(wl, ni, res, size)
    new_pupil = hanser_pupil(kr, na, wl)
    mask = new_pupil.real
    mse = np.zeros(max_iters)
    mse_diff = np.zeros(max_iters)
    pupil_diff = np.zeros(max_iters)
    old_mse = old_pupil = np.nan
    history = []
    for i in range(max_iters):
        defocus = hanser_defocus(kz, zrange)
        psfa = fftshift(ifftn(new_pupil * defocus, axes=(1, 2)), axes=(1, 2))
        model_psfi = abs(psfa) ** 2
        new_mse = calc_mse(data, model_psfi)
        mse[i] = new_mse
        if i > 0:
            mse_diff[i] = abs(old_mse - new_mse) / old_mse
            pupil_diff[i] = (abs(old_pupil - new_pupil) ** 2).mean() / (abs(old_pupil) ** 2).mean()
        else:
            mse_diff[i] = pupil_diff[i] = np.nan
        if return_history:
            history.append(new_pupil.copy())
        if pupil_diff[i] < pupil_tol or mse_diff[i] < mse_tol or mse[i] < mse_tol:
            break
        old_mse, old_pupil = new_mse, new_pupil
        new_psf = mag * np.exp(1j * np.angle(psfa))
        new_pupils = fftn(ifftshift(new_psf, axes=(1, 2)), axes=(1, 2))
        new_pupils /= defocus
        new_pupil = new_pupils.mean(0) * mask
        if phase_only:
            new_pupil = np.exp(1j * np.angle(new_pupil)) * mask
    out = dict(
        pupil=new_pupil,
        mse=mse[: i + 1],
        mse_diff=mse_diff[: i + 1],
        pupil_diff=pupil_diff[: i + 1],
        psfi=model_psfi,
        iters=i + 1,
        mask=mask,
    )
    if return_history:
        out["history"] = history
    return out


# ----------------------------------------------------------------------------------------
# SheppardPSF (otf.py:446-592)
# ----------------------------------------------------------------------------------------


def sheppard_shell(wl, na, ni, res, size, zres, zsize, dual=False):
    """Voxels of the spherical cap and their k coordinates (otf.py:497-533).

    Returns (valid_points (nz,N,N) bool, kzz, kyy, kxx on the shell, dk, dkz).
    """
    k = fftfreq(size, res)
    kz = fftfreq(zsize, zres)
    k_tot = np.meshgrid(kz, k, k, indexing="ij")
    kr = np.linalg.norm(k_tot, axis=0)
    kmag = ni / wl
    dk, dkz = k[1] - k[0], kz[1] - kz[0]
    kz_min = np.sqrt(kmag**2 - (na / wl) ** 2)
    assert kz_min >= 0
    if dk != dkz:
        with np.errstate(invalid="ignore"):
            dd = np.array((dkz, dk, dk)).reshape(3, 1, 1, 1)
            dkr = np.linalg.norm(np.array(k_tot) * dd, axis=0) / kr
        dkr[0, 0, 0] = 0.0
    else:
        dkr = dk
    kzz = abs(k_tot[0]) if dual else k_tot[0]
    valid = np.logical_and(abs(kr - kmag) < dkr, kzz > kz_min + dkr)
    return valid, k_tot[0][valid], k_tot[1][valid], k_tot[2][valid], dk, dkz


def sheppard_otfa(wl, na, ni, res, size, zres, zsize, vec_corr="none", condition="sine", dual=False):
    """Amplitude OTF on the shell, shifted (otf.py:535-582).  Scalar case ignores `condition`."""
    valid, kzz, kyy, kxx, _, _ = sheppard_shell(wl, na, ni, res, size, zres, zsize, dual)
    kn = np.linalg.norm((kxx, kyy, kzz), axis=0)
    m, n, s = kxx / kn, kyy / kn, kzz / kn
    if condition == "sine":
        a = 1.0 / np.sqrt(s)
    elif condition == "herschel":
        a = 1.0 / s
    elif condition == "none":
        a = 1.0
    else:
        raise ValueError(condition)
    if vec_corr != "none":
        plist = []
        if vec_corr in ("z", "total"):
            plist += [-m, -n]
        if vec_corr in ("y", "total"):
            plist += [-n * m / (1 + s), 1 - n**2 / (1 + s)]
        if vec_corr in ("x", "total"):
            plist += [1 - m**2 / (1 + s), -m * n / (1 + s)]
        otf = np.zeros((len(plist), zsize, size, size), dtype=complex)
        for o, p in zip(otf, plist):
            o[valid] = p * a
    else:
        otf = np.zeros((1, zsize, size, size), dtype=complex)
        otf[0][valid] = 1.0
    return fftshift(otf, axes=(1, 2, 3))


def sheppard_psfa(otfa):
    """Centered 3D inverse FFT of the shell (otf.py:589-592)."""
    return easy_ifft(otfa, axes=(1, 2, 3))


# ----------------------------------------------------------------------------------------------
# N4: microscope models (pyotf/microscope.py) -- restated for the widefield and the confocal microscope
# ----------------------------------------------------------------------------------------------
def bin_ndarray(a, bin_size):
    """dphtools.utils.bin_ndarray (un-vendored dependency, pyotf/microscope.py:24) with its default operation: the
    sum over bin_size blocks along every axis."""
    a = np.asarray(a)
    shape = []
    for n, b in zip(a.shape, bin_size):
        assert n % b == 0
        shape += [n // b, b]
    return a.reshape(shape).sum(axis=tuple(range(1, 2 * a.ndim, 2)))


def widefield_psf(model_psf, oversample_factor, oversampled_size):
    """WidefieldMicroscope.PSF (pyotf/microscope.py:106-124)."""
    psf = model_psf
    if oversample_factor > 1:
        if oversampled_size % 2 == 0:
            shift = oversample_factor // 2
            psf = np.roll(psf, (shift, shift), axis=(1, 2))
        psf = bin_ndarray(psf, (1, oversample_factor, oversample_factor))
    return psf / psf.sum()


def disk_kernel(radius):
    """_disk_kernel (pyotf/microscope.py:180-187)."""
    full_size = int(np.ceil(radius * 2))
    if full_size % 2 == 0:
        full_size += 1
    coords = np.indices((full_size, full_size)) - (full_size - 1) // 2
    r = np.sqrt((coords ** 2).sum(0))
    kernel = r < radius
    return kernel / kernel.sum()


def confocal_model_psf(psf_em, psf_exc, wl, na, res, pinhole_size):
    """ConfocalMicroscope.model_psf (pyotf/microscope.py:213-227)."""
    from scipy.signal import fftconvolve

    airy_unit = 1.22 * wl / na / res
    pixel_pinhole_radius = pinhole_size * airy_unit / 2
    if pixel_pinhole_radius > 1.5:
        psf_det = fftconvolve(psf_em, disk_kernel(pixel_pinhole_radius)[None], "same", axes=(1, 2))
    else:
        psf_det = psf_em
    return psf_det * psf_exc


def sim_model_psf(psf_i, otf_i, res, zres, ni, wl_exc, na_exc, wiener, coherent, dc, dc_suppress, orientations):
    """BaseSIMMicroscope.model_psf (pyotf/microscope.py:316-401) on the model's PSFi / OTFi ([nz, N, N])."""
    nz, n, _ = psf_i.shape
    x = (np.arange(n) - (n + 1) // 2) * res
    z = (np.arange(nz) - (nz + 1) // 2) * zres
    zz, yy, xx = z[:, None, None], x[None, :, None], x[None, None, :]
    freq = 2 * np.pi * ni / wl_exc
    alpha = np.arcsin(na_exc / ni)
    if wiener is None:
        psf = psf_i
    else:
        assert wiener >= 0
        otf = abs(otf_i) ** 2
        if wiener == 0:
            wiener_otf = otf > 1e-16
        else:
            w = otf.max() * wiener ** 2
            wiener_otf = otf / (otf + w)
        psf = np.fft.ifftshift(np.fft.ifftn(np.fft.fftshift(wiener_otf))).real        # easy_ifft(...).real
    if dc:
        base_pattern = np.exp(1j * zz * freq) * np.ones(psf.shape[1:], dtype=complex)[None]
    else:
        base_pattern = np.zeros(psf.shape, dtype=complex)
    if not coherent:
        sim_psf = np.zeros_like(psf)
    for orientation in orientations:
        exc_pattern = base_pattern if coherent else base_pattern.copy()      # coherent: the beams of all orientations add up
        rr = xx * np.cos(orientation) + yy * np.sin(orientation)
        for theta in (-alpha, alpha):
            exc_pattern += np.exp(1j * ((rr * np.sin(theta) + zz * np.cos(theta)) * freq))
        if not coherent:
            sim_psf += psf * (abs(exc_pattern) ** 2)
    if coherent:
        sim_psf = psf * (abs(exc_pattern) ** 2)
    elif dc_suppress:
        sim_psf -= (2 * len(orientations) - 1) * psf
    return sim_psf


# ----------------------------------------------------------------------------------------------
# N3: data preparation (pyotf/utils.py:76-147, 161-201)
# ----------------------------------------------------------------------------------------------
def center_data(data):
    """pyotf/utils.py:76-124: roll the maximum to the central index of every axis."""
    centered = data.copy()
    max_loc = np.unravel_index(data.argmax(), data.shape)
    for i, (x0, nx) in enumerate(zip(max_loc, data.shape)):
        centered = np.roll(centered, nx // 2 - x0, i)
    return centered


def remove_bg(data, multiplier):
    """pyotf/utils.py:127-147: subtract multiplier x the mode of an unsigned integer array."""
    if not np.issubdtype(data.dtype, np.integer):
        raise ValueError("Floating point numbers unsupported")
    if np.any(data < 0):
        raise ValueError("Negative values unsupported")
    mode = np.bincount(data.ravel()).argmax()
    return data - multiplier * float(mode)


def _fft_pad_constant(data, newshape):
    """Stand-in for dphtools.utils.fft_pad(mode="constant") (un-vendored; PARITY UNPINNED): zero padding that keeps the
    element at index n // 2 of every axis at index new_n // 2."""
    pads = []
    for n, m in zip(data.shape, newshape):
        before = m // 2 - n // 2
        pads.append((before, m - n - before))
    return np.pad(data, pads, mode="constant")


def _slice_maker(centers, width):
    """Stand-in for dphtools.utils.slice_maker (un-vendored; PARITY UNPINNED): start = centre - width // 2,
    stop = start + width, only the start clipped at 0 (a window over the lower edge is shorter)."""
    out = []
    for c in centers:
        lo = int(c) - int(width) // 2
        hi = lo + int(width)
        if hi <= 0:
            lo, hi = 0, 0
        out.append(slice(max(lo, 0), hi))
    return tuple(out)


def prep_data_for_pr(data, xysize=None, multiplier=1.05):
    """pyotf/utils.py:161-201.  The no-resize branch (xysize == ny == nx: BASELINE config 2) is pinned to the live
    reference; the pad / crop branches go through the two dphtools stand-ins above and are unpinned."""
    nz, ny, nx = data.shape
    data_without_bg = remove_bg(data, multiplier)
    if xysize is None:
        xysize = max(ny, nx)
    if xysize == ny == nx:
        pad_data = data_without_bg
    elif xysize >= max(ny, nx):
        pad_data = _fft_pad_constant(data_without_bg, (nz, xysize, xysize))
    else:
        my_slice = _slice_maker(((ny + 1) // 2, (nx + 1) // 2), xysize)
        pad_data = center_data(data_without_bg)[(Ellipsis,) + my_slice]
    return np.fmax(0, pad_data)
