/* pyotf_b200 -- C ABI of the B200-native PSF/OTF + phase-retrieval engine.
 *
 * pyotf (david-hoffman/pyotf) has no FFI layer: its boundary is a Python object API.  This
 * header is the C boundary *under* our Python mirror of that API (pyotf_b200/otf.py,
 * phaseretrieval.py, zernike.py); every entry point names the reference code it replaces.
 *
 * Conventions
 *   - every function returns 0 on success, non-zero on error; the message is available from
 *     pyotf_b200_last_error() (thread-local).  No C++ exception crosses this boundary.
 *   - `stream` is a cudaStream_t passed as void*; all work is enqueued asynchronously on it
 *     unless the function is documented as synchronising.
 *   - `*_dev` pointers are device pointers owned by the caller (torch allocations in the
 *     Python host); handles own only their small O(N^2) parameter tables.
 *   - dtype: 0 = float32 / complex64 ("fp32 mode"), 1 = float64 / complex128 ("fp64 mode").
 *   - complex arrays are interleaved (re, im).  "unshifted" = numpy.fft.fftfreq order,
 *     "shifted" = numpy.fft.fftshift order, exactly as the reference returns them.
 */
#ifndef PYOTF_B200_H
#define PYOTF_B200_H

#ifdef __cplusplus
extern "C" {
#endif

#define PYOTF_B200_ABI_VERSION 1

enum { PYOTF_F32 = 0, PYOTF_F64 = 1 };
enum { PYOTF_VEC_NONE = 0, PYOTF_VEC_X = 1, PYOTF_VEC_Y = 2, PYOTF_VEC_Z = 3, PYOTF_VEC_TOTAL = 4 };
enum { PYOTF_COND_NONE = 0, PYOTF_COND_SINE = 1, PYOTF_COND_HERSCHEL = 2 };

int pyotf_b200_abi_version(void);
int pyotf_b200_version(void);                 /* = pyotf_b200_abi_version (the name of SURVEY.md section 8b) */
const char* pyotf_b200_last_error(void);
/* sm count, max dynamic shared memory per block, compute capability of `device` */
int pyotf_b200_device_query(int device, int* sm_count, int* smem_optin_bytes, int* cc_major, int* cc_minor);

/* ------------------------------------------------------------------------------------
 * HanserPSF  (pyotf/otf.py:265-443)
 * ---------------------------------------------------------------------------------- */
typedef struct pyotf_hanser_params {
    double wl, na, ni, res; /* BasePSF.__init__                      otf.py:53-104 */
    int size;               /* x/y size N (power of two, 32..256 on the fused path)   */
    int vec_corr;           /* PYOTF_VEC_*                           otf.py:405-417 */
    int condition;          /* PYOTF_COND_*                          otf.py:396-403 */
    int dtype;              /* PYOTF_F32 / PYOTF_F64                                  */
    int support;            /* -1: no user pupil, ideal mask kr <= na/wl is applied
                               (HanserPSF._gen_pupil, otf.py:346-355);
                               -2: a pupil is supplied that lives on the ideal pupil's box
                               (phase retrieval, apply_aberration);
                               h >= 0: a user pupil is supplied whose non-zero support is
                               |f| <= h frequency bins (see pyotf_b200_pupil_support)    */
} pyotf_hanser_params;

typedef struct pyotf_hanser pyotf_hanser_t;

/* Builds the fp64 pupil-plane geometry on the device (replaces HanserPSF._gen_kr,
 * otf.py:335-344, and the z-independent factors of _gen_psf, otf.py:392-417). */
int pyotf_b200_hanser_create(const pyotf_hanser_params* params, void* stream, pyotf_hanser_t** out);
int pyotf_b200_hanser_destroy(pyotf_hanser_t* h);
/* K = edge of the compact support box, rows = allocated rows of every compact table / pupil
 * (K + zero padding: compact arrays are [rows][K]), f0 = first frequency bin, nvec = 1|2|6 */
int pyotf_b200_hanser_info(const pyotf_hanser_t* h, int* K, int* rows, int* f0, int* nvec);
/* copies the geometry tables out (any pointer may be NULL): kz [rows*K] f64, amp [nvec*rows*K]
 * in the handle's dtype, mask [rows*K] u8 -- used by the parity tests for the edge predicates */
int pyotf_b200_hanser_tables(const pyotf_hanser_t* h, double* kz_dev, void* amp_dev, unsigned char* mask_dev,
                             void* stream);

/* Largest |frequency bin| at which an unshifted size x size complex128 pupil is non-zero.
 * SYNCHRONISES the stream (returns one int to the host). */
int pyotf_b200_pupil_support(const void* pupil_c128_dev, int size, int* hmax_host, void* stream);

/* Gathers `batch` unshifted size x size complex128 pupils into the handle's compact centred
 * [rows][K] layout and dtype (the `pupil_base` argument of _gen_psf, otf.py:362-389). */
int pyotf_b200_hanser_pack_pupil(const pyotf_hanser_t* h, const void* pupil_c128_dev, int batch,
                                 void* pupilc_out_dev, void* stream);

/* K1 -- HanserPSF._gen_psf + BasePSF.PSFi  (otf.py:362-428, 204-207).
 *   z_dev        [nz] float64 plane positions (HanserPSF.zrange, otf.py:296-333)
 *   pupilc_dev   NULL (ideal pupil) or [batch] compact pupils, `pupil_stride` elements apart
 *                (0 = one pupil shared by the whole batch; rows*K for a dense batch)
 *   psfa_out_dev NULL or [batch][nvec][nz][N][N] complex, spatially shifted
 *   psfi_out_dev NULL or [batch][nz][N][N] real, spatially shifted                         */
int pyotf_b200_hanser_psf(const pyotf_hanser_t* h, const double* z_dev, int nz, const void* pupilc_dev,
                          int batch, long long pupil_stride, void* psfa_out_dev, void* psfi_out_dev,
                          void* stream);

/* pyotf_b200_hanser_psf with the plane positions given on the HOST, as the reference holds them (HanserPSF.zrange is
 * a numpy array, otf.py:296-333).  Up to 256 planes of a fused-size handle (32 / 64 / 128 / 256) travel by value in
 * the kernel arguments: a single-stack ideal-pupil launch then reads nothing but the handle's immutable tables
 * before its first store, so back-to-back launches overlap (programmatic dependent launch; each still waits for its
 * predecessor before its first store) with no promise needed from the caller.  Longer stacks and other sizes are
 * uploaded into a buffer the handle owns (stream-ordered) and take the ordinary path.  z_host may be reused as soon
 * as the call returns.  Results are bit-identical to pyotf_b200_hanser_psf. */
int pyotf_b200_hanser_psf_zhost(pyotf_hanser_t* h, const double* z_host, int nz, const void* pupilc_dev,
                                int batch, long long pupil_stride, void* psfa_out_dev, void* psfi_out_dev,
                                void* stream);

/* Streaming mode for back-to-back single-stack launches with the ideal pupil.  enable = 1: the caller promises
 * that z_dev is not written by the kernel that precedes a pyotf_b200_hanser_psf call in its stream (uploads by
 * cudaMemcpy are fine).  Consecutive launches then overlap (programmatic dependent launch): the next stack's
 * CTAs fill the SMs the previous launch leaves idle in its last wave; each launch still waits for its
 * predecessor before its first store, so results and their order are unchanged.  Default 0. */
int pyotf_b200_hanser_set_overlap(pyotf_hanser_t* h, int enable);

/* Full-grid pupil coordinates for the API hooks (HanserPSF._gen_kr, otf.py:335-344):
 * kr, phi, kz as [size*size] float64, unshifted; any output may be NULL. */
int pyotf_b200_hanser_kspace(double wl, double ni, double res, int size, double* kr_dev, double* phi_dev,
                             double* kz_dev, void* stream);
/* exp(2 pi i kz z) -> [nz][npix] complex128 (HanserPSF._calc_defocus, otf.py:357-360). */
int pyotf_b200_hanser_defocus(const double* kz_dev, long long npix, const double* z_dev, int nz, void* out_c128_dev,
                              void* stream);

/* ------------------------------------------------------------------------------------
 * Zernike polynomials  (pyotf/zernike.py:251-365) and apply_aberration (otf.py:595-645)
 * ---------------------------------------------------------------------------------- */
/* out[mode][pt] = Z_{n[mode]}^{m[mode]}(r[pt], theta[pt]), float64; zero for r > 1. */
int pyotf_b200_zernike(const double* r_dev, const double* theta_dev, long long npts, const int* n_dev,
                       const int* m_dev, int nmodes, int norm, double* out_dev, void* stream);
/* Normal equations of the Zernike least-squares fit (phaseretrieval._fit_to_zerns, phaseretrieval.py:406-432):
 * G [nmodes][nmodes + nrhs] = Z [Z^T | b^T] with Z [nmodes][npts] (pyotf_b200_zernike on the in-pupil pixels) and
 * b [nrhs][npts] the data vectors; all float64 on the device.  The nmodes x nmodes solve stays on the host. */
int pyotf_b200_zernike_gram(const double* z_dev, const double* b_dev, long long npts, int nmodes, int nrhs,
                            double* g_out_dev, void* stream);
/* Batched aberrated pupils in the handle's compact layout / dtype:
 *   pupil_b = (sum_j mcoefs[b][j] Z_j + |ideal pupil|) * exp(i sum_j pcoefs[b][j] Z_j)
 * with Z_j evaluated at r = kr*wl/na, theta = phi (otf.py:626-642).  n_dev/m_dev are the
 * (n, m) degrees of the nmodes modes (Noll order for apply_aberration); mcoefs/pcoefs are
 * [batch][nmodes] float64, either may be NULL.  The handle must have support == -2. */
int pyotf_b200_hanser_aberration_pupil(const pyotf_hanser_t* h, const int* n_dev, const int* m_dev, int nmodes,
                                       const double* mcoefs_dev, const double* pcoefs_dev, int batch,
                                       void* pupilc_out_dev, void* stream);

/* ------------------------------------------------------------------------------------
 * SheppardPSF  (pyotf/otf.py:446-592)
 * ---------------------------------------------------------------------------------- */
typedef struct pyotf_sheppard_params {
    double wl, na, ni, res, zres;
    int size, zsize;        /* N (x/y) and number of kz planes; any positive sizes */
    int vec_corr;           /* PYOTF_VEC_* */
    int condition;          /* PYOTF_COND_* (ignored by the scalar model, otf.py:577-579) */
    int dual;               /* 0 | 1: mirrored cap (otf.py:525-529) */
    int dtype;              /* PYOTF_F32 / PYOTF_F64 */
} pyotf_sheppard_params;

/* K3 generator -- SheppardPSF._gen_kr + _gen_otf (otf.py:497-582).
 *   otfa_out_dev  NULL or [nvec][zsize][size][size] complex, fftshift-ed (== SheppardPSF.OTFa)
 *   valid_out_dev NULL or [zsize][size][size] uint8, UNSHIFTED (== SheppardPSF.valid_points)
 * PSFa = pyotf_b200_fftn(OTFa, axes (1,2,3), inverse, shift_in, shift_out)  (otf.py:589-592);
 * PSFi = pyotf_b200_abs2_sum0(PSFa)                                         (otf.py:204-207). */
int pyotf_b200_sheppard_otfa(const pyotf_sheppard_params* params, void* otfa_out_dev, unsigned char* valid_out_dev,
                             void* stream);
/* K3 fused -- OTFa, PSFa (otf.py:589-592) and PSFi (otf.py:204-207) of one model in one call; any output
 * may be NULL (scratch is stream-ordered).  The generator records which kz planes / (kz, ky) rows hold
 * shell voxels and the first two passes of the 3D inverse FFT skip the all-zero lines (at BASELINE
 * config 3 only 87 of 256 planes and 17 % of the rows are occupied).
 *   otfa_out_dev, psfa_out_dev [nvec][zsize][size][size] complex, shifted; psfi_out_dev [zsize][size][size] */
int pyotf_b200_sheppard_psf(const pyotf_sheppard_params* params, void* otfa_out_dev, void* psfa_out_dev,
                            void* psfi_out_dev, void* stream);

/* ------------------------------------------------------------------------------------
 * Data preparation for phase retrieval  (pyotf/utils.py:76-201)
 * ---------------------------------------------------------------------------------- */
/* prep_data_for_PR for an unsigned 16-bit camera stack on the device: mode-based background removal
 * (remove_bg, utils.py:127-147), pad (xysize >= max(ny, nx)) or centre-on-maximum + crop (center_data,
 * utils.py:76-124), clip at zero (utils.py:201).
 *   data_u16_dev [nz][ny][nx];  xysize <= 0 means max(ny, nx);  out_dev [nz][xysize][xysize] in `dtype`
 *   mode_host    receives the background mode (may be NULL).  SYNCHRONISES the stream (the crop branch
 *   needs the arg-max on the host to lay out the gather). */
int pyotf_b200_prep_data_u16(const unsigned short* data_u16_dev, int nz, int ny, int nx, int xysize, double multiplier,
                             int dtype, void* out_dev, int* mode_host, void* stream);

/* ------------------------------------------------------------------------------------
 * Phase retrieval  (pyotf/phaseretrieval.py:32-149, the loop :95-135, _calc_mse :401-403)
 * ---------------------------------------------------------------------------------- */
typedef struct pyotf_pr pyotf_pr_t;
typedef struct pyotf_comm pyotf_comm_t;

/* K2 state for one measured stack.  `params` as for HanserPSF; vec_corr / condition / support
 * are overridden (retrieve_phase forces "none"/"none", phaseretrieval.py:74-76).
 *   z_dev        [nz] float64 plane positions of THIS rank's planes
 *   nz_total     planes over all ranks (== nz on one GPU): the z-mean and the MSE divide by it
 *   data_dev     [nz][N][N] measured intensities, fftshift-ed like PSFi; data_dtype PYOTF_F32/F64
 *   rows_per_cta 0 = auto, else 16 | 32 | 64 (tuning knob: output rows of a plane per CTA)
 * The pupil estimate starts as the ideal pupil (HanserPSF._gen_pupil, phaseretrieval.py:89). */
int pyotf_b200_pr_create(const pyotf_hanser_params* params, const double* z_dev, int nz, long long nz_total,
                         const void* data_dev, int data_dtype, int rows_per_cta, void* stream, pyotf_pr_t** out);
int pyotf_b200_pr_destroy(pyotf_pr_t* s);
/* K, rows, f0 of the compact pupil layout; rows per CTA; partial pupils per iteration */
int pyotf_b200_pr_info(const pyotf_pr_t* s, int* K, int* rows, int* f0, int* rows_per_cta, int* nparts);
/* Kernel launches one iteration of pyotf_b200_pr_run takes on this state (loop body of phaseretrieval.py:95-130):
 * 1 = the plane kernel finalizes the iteration itself behind a grid-wide arrival barrier (single GPU, every CTA
 * resident) -- an upper bound: one launch then runs up to 8 iterations back to back, a second grid barrier between
 * them (PYOTF_B200_PR_ITERS_PER_LAUNCH=1: one launch per iteration); 2 = plane kernel + finalize kernel (long stacks; z-sharded runs with a peer window, where the finalize
 * kernel also sums over the ranks); 3 = z-sharded with the NCCL fallback; 5 = the four-pass path of other sizes. */
int pyotf_b200_pr_kernels_per_iteration(const pyotf_pr_t* s, int sharded, int* n);
/* Inspection hook for the parity tests: device pointers of the intermediates of the last iteration
 * (four-pass path: W [nz][K][N] and P [nz][N][N] complex; both paths: partial pupils [nparts][K][K]). */
int pyotf_b200_pr_debug_buffers(const pyotf_pr_t* s, void** W, void** P, void** partial);
/* Replace the current pupil estimate by an unshifted N x N complex128 pupil (NULL: the ideal
 * pupil again).  Lets a caller re-seed single iterations (parity tests) or warm-start. */
int pyotf_b200_pr_set_pupil(pyotf_pr_t* s, const void* pupil_c128_dev, void* stream);
/* which = 0: the final new_pupil (phaseretrieval.py:144-148); 1: the pupil last applied to the
 * model (what result.model.PSFi shows).  Output: unshifted N x N complex128. */
int pyotf_b200_pr_get_pupil(const pyotf_pr_t* s, int which, void* pupil_c128_dev_out, void* stream);
/* Runs up to max_iters iterations of the loop, the convergence test
 * (pupil_diff < pupil_tol || mse_diff < mse_tol || mse < mse_tol, :114) evaluated on the device.
 * comm != NULL: planes are sharded over the ranks of `comm`; one all-reduce per iteration.
 * Host outputs (any may be NULL) hold iters_run entries each, like the reference's trimmed
 * arrays (:133-135).  SYNCHRONISES the stream before returning. */
int pyotf_b200_pr_run(pyotf_pr_t* s, int max_iters, double pupil_tol, double mse_tol, int phase_only,
                      pyotf_comm_t* comm, double* mse_host, double* mse_diff_host, double* pupil_diff_host,
                      int* iters_run_host, void* stream);
/* = pyotf_b200_pr_run (the name of SURVEY.md section 8b) */
int pyotf_b200_pr_iterate(pyotf_pr_t* s, int max_iters, double pupil_tol, double mse_tol, int phase_only,
                          pyotf_comm_t* comm, double* mse_host, double* mse_diff_host, double* pupil_diff_host,
                          int* iters_run_host, void* stream);

/* z-sharded phase retrieval without a collective library call per iteration: every rank exports a small window
 * through CUDA IPC and maps the windows of all other ranks of the box.  Each iteration, CTA b of a rank's finalize
 * kernel sums the rank's partial pupils for its 32 pixels, stores them into entry [rank] of EVERY window (posted
 * writes over NVLink, every 16-byte word carrying the iteration's epoch next to its payload), spins on the words of
 * CTA b of all ranks in its own window until they carry the epoch, and adds them in rank order: no fence, no
 * round trip; all ranks compute bit-identical pupils and stop at the same iteration, with two
 * kernels per iteration exactly like the single-GPU loop.  This replaces the ncclAllReduce that the reference's
 * single-process loop has no counterpart for:
 *   1. pr_window_alloc on every rank -> 64-byte handle;  2. exchange the handles (any transport);
 *   3. pr_window_open(all handles in rank order).  pr_run(comm != NULL) then uses the window.
 * One process per GPU, all GPUs on one node with peer access (NVLink / NVSwitch). */
int pyotf_b200_pr_window_alloc(pyotf_pr_t* s, int world, unsigned char handle_out[64]);
int pyotf_b200_pr_window_open(pyotf_pr_t* s, const unsigned char* handles /* [world][64] */, int rank, int world);

/* HOST function (host pointers): 2D phase unwrapping of one ny x nx image of wrapped phases,
 * Herraez et al. 2002 -- stand-in for skimage.restoration.unwrap_phase (phaseretrieval.py:146). */
int pyotf_b200_unwrap_phase_2d(const double* in_host, double* out_host, int ny, int nx);

/* ------------------------------------------------------------------------------------
 * Multi-GPU plumbing (one process per GPU)
 * ---------------------------------------------------------------------------------- */
/* dlopen the NCCL library the host process uses (path of libnccl.so.2; NULL/"" = default search) */
int pyotf_b200_comm_load(const char* libnccl_path);
/* rank 0 creates the 128-byte id and distributes it (torch.distributed / any side channel) */
int pyotf_b200_comm_unique_id(unsigned char id_out[128]);
int pyotf_b200_comm_create(const unsigned char id[128], int rank, int world, pyotf_comm_t** out);
int pyotf_b200_comm_destroy(pyotf_comm_t* c);
/* in-place sum all-reduce of n float64 on `stream` */
int pyotf_b200_comm_allreduce_f64(pyotf_comm_t* c, double* buf_dev, long long n, void* stream);

/* ------------------------------------------------------------------------------------
 * Shifted N-d FFT  (pyotf/utils.py:15-22; OTFi otf.py:209-212; OTFa otf.py:435-438, 589-592)
 * ---------------------------------------------------------------------------------- */
/* out = [fftshift](fftn)([ifftshift] in) over `axes` (forward), or
 * out = [ifftshift](ifftn)([fftshift] in) (inverse; numpy "backward" 1/n normalisation).
 * `in` is a C-contiguous array of `ndim` <= 8 dims, real (real_in = 1) or complex; `out` is
 * complex with the same shape and may alias `in` only when `in` is complex.  shift_in /
 * shift_out = 1 give easy_fft / easy_ifft; 0/0 is the plain transform. */
int pyotf_b200_fftn(const void* in_dev, void* out_dev, int ndim, const long long* shape, const int* axes, int naxes,
                    int inverse, int real_in, int shift_in, int shift_out, int dtype, void* stream);
/* easy_fft (direction < 0) / easy_ifft (direction > 0) of a C-contiguous [nz][ny][nx] array over all three axes
 * (pyotf/utils.py:15-22): pyotf_b200_fftn with ndim = naxes = 3 and both shifts (the form of SURVEY.md section 8b). */
int pyotf_b200_fft3d_shifted(const void* in_dev, void* out_dev, int nz, int ny, int nx, int direction, int real_in,
                             int dtype, void* stream);
/* out[i] = sum_v |a[v][i]|^2  (BasePSF.PSFi from an existing PSFa, otf.py:204-207) */
int pyotf_b200_abs2_sum0(const void* a_dev, int nvec, long long n, void* out_dev, int dtype, void* stream);

/* ---- N4: microscope models on the device-resident PSFi (pyotf/microscope.py) ---------------------------------
 * WidefieldMicroscope.PSF (microscope.py:106-124): np.roll by `shift` along y and x, sum factor x factor lateral
 * bins (dphtools.utils.bin_ndarray, un-vendored: block sum), optionally divide by the total.
 *   in [nz][size][size] real -> out [nz][size/factor][size/factor] real (dtype = PYOTF_F32 / PYOTF_F64). */
int pyotf_b200_bin_lateral(const void* in_dev, int nz, int size, int factor, int shift, void* out_dev, int normalise,
                           int dtype, void* stream);
/* ConfocalMicroscope.model_psf (microscope.py:180-187, 213-227): out = fftconvolve(em, disk(radius)[None], "same",
 * axes=(1, 2)) * exc with disk = (sqrt(i^2 + j^2) < radius) / count on a ceil(2 radius)|1 window; all [nz][size][size]. */
int pyotf_b200_disk_conv_mul(const void* em_dev, const void* exc_dev, int nz, int size, double radius, void* out_dev,
                             int dtype, void* stream);
/* BaseSIMMicroscope.model_psf, the Wiener-filtered OTF (pyotf/microscope.py:341-355): out[i] (real) =
 * a / (a + max(a) * wiener^2) with a = |otf[i]|^2 of the complex input, or (a > 1e-16) when wiener == 0. */
int pyotf_b200_wiener_otf(const void* otf_c_dev, long long n, double wiener, void* out_dev, int dtype, void* stream);
/* BaseSIMMicroscope.model_psf (pyotf/microscope.py:316-401): out = psf * excitation intensity of the structured
 * illumination (two beams at +-alpha per orientation, alpha = arcsin(na_exc / ni), plus a DC beam when dc != 0), the
 * orientations (host array, <= 16) summed incoherently (coherent == 0; dc_suppress subtracts (2 n - 1) psf) or added
 * coherently before squaring (lattice SIM).  psf [nz][size][size]: real, or complex when psf_is_complex (its real
 * part is used: easy_ifft(wiener_otf).real).  Coordinates are (arange(n) - (n + 1) // 2) * res / zres. */
int pyotf_b200_sim_psf(const void* psf_dev, int psf_is_complex, int nz, int size, double res, double zres, double ni,
                       double wl_exc, double na_exc, const double* orientations_host, int n_orient, int coherent, int dc,
                       int dc_suppress, void* out_dev, int dtype, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* PYOTF_B200_H */
