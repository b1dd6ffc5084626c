/*
 * adorym_b200 -- C ABI of the B200-native multislice forward-model + gradient + update hot path.
 *
 * Drop-in boundary for mdw771/adorym's hot path.  The reference is pure Python; the
 * "FFI" it would bind is a ctypes/pybind stub called from a torch.autograd.Function that replaces
 *   - adorym/propagate.py:131-288   multislice_propagate_batch  (+ wrappers.py:699-779 fft2/ifft2/convolve,
 *                                    wrappers.py:600-608 exp_complex, forward_model.py:313-331 patch gather,
 *                                    util.py:1327-1351 pad_object)            -> adm_transmission_fwd / _adj +
 *                                                                                adm_multislice_fwd / _bwd
 *   - adorym/forward_model.py:88-103, 353, 371-375 (norm + mismatch loss)     -> adm_detector_loss
 *   - adorym/util.py:536-552 apply_rotation / wrappers.py:1105-1147 grid_sample and its autograd transpose
 *                                                                              -> adm_rotate_fwd / adm_rotate_adj
 *   - adorym/optimizers.py:275-337  AdamOptimizer.apply_gradient              -> adm_adam_step
 *   - adorym/regularizers.py:30-110 + util.py:1427-1440 (L1 / reweighted L1 / TV value and gradient)
 *                                                                              -> adm_regularizer_l1tv
 *   - adorym/util.py:380-397 realign_image_fourier on the probe (+ transpose)  -> adm_probe_shift_fwd / _bwd
 *
 * Conventions
 *   - every pointer is a DEVICE pointer to contiguous fp32 (or int32) memory owned by the caller
 *     (PyTorch allocations); kernels never allocate, workspaces are passed in;
 *   - complex arrays are interleaved (re, im) pairs; the object is [..., 2] = (delta, beta) or (re, im);
 *   - `stream` is a cudaStream_t passed as void* (0 = legacy default stream); all work is enqueued
 *     asynchronously on it; no host synchronisation inside;
 *   - return value: 0 = ok, 1 = bad argument, 2 = unsupported configuration, 3 = CUDA launch error
 *     (cudaGetLastError); there is NO CPU fallback.
 */
#ifndef ADORYM_B200_H
#define ADORYM_B200_H

#ifdef __cplusplus
extern "C" {
#endif

#define ADM_ABI_VERSION 8

/* unknown_type */
#define ADM_DELTA_BETA 0
#define ADM_REAL_IMAG 1
/* free_prop */
#define ADM_PROP_NONE 0  /* free_prop_cm in (0, None): detector sees the exit wave (propagate.py:263)        */
#define ADM_PROP_FAR 1   /* free_prop_cm == 'inf': fftshift(fft2) or fftshift(ifft2) (propagate.py:264-270)  */
#define ADM_PROP_NEAR 2  /* finite distance: one more Fresnel convolution (propagate.py:271-280, :537-553)   */
/* loss_type / raw_data_type (forward_model.py:88-103) */
#define ADM_LOSS_LSQ 0
#define ADM_LOSS_POISSON 1
#define ADM_RAW_MAGNITUDE 0
#define ADM_RAW_INTENSITY 1

typedef struct adm_multislice_cfg {
    int P;                 /* patch / probe edge (square); supported: 64, 72                                  */
    int n_slices;          /* Nz, number of object slices traversed (binning == 1)                            */
    int B;                 /* diffraction patterns (probe positions) in this minibatch                        */
    int M;                 /* incoherent probe modes                                                          */
    int Y, X;              /* in-plane extent of the (rotated) object volume, layout [Nz][Y][X][2]            */
    float k1;              /* 2*pi*dz/lambda if scale_ri_by_k else 1 (propagate.py:215)                       */
    int sign_convention;   /* +1 or -1 (propagate.py:236-241)                                                 */
    int unknown_type;      /* ADM_DELTA_BETA / ADM_REAL_IMAG                                                  */
    int free_prop;         /* ADM_PROP_*                                                                      */
    int normalize_fft;     /* far field with norm='ortho' (wrappers.py:725-756)                               */
    int probe_per_pattern; /* 0: probe is [M][P][P][2]; 1: probe is [B][M][P][P][2] (per-position shifted)    */
    int h_separable;       /* 1: Hprep / HprepC were made by adm_prepare_transfer_function_separable ([2][P][2]) */
    int h_per_slice;       /* 1 (needs h_separable): Hprep / HprepC are [n_slices-1][2][P][2], one separable transfer
                              function per gap between slices z and z+1 -- sparse multislice with unequal slice
                              spacings (propagate.py:479-534 sparse_multislice_propagate_batch)                 */
    int hfree_separable;   /* 1: the free-space transfer function behind Hfree / HfreeC (ADM_PROP_NEAR) is an outer product
                              hy(ky) hx(kx) -- the paraxial Fresnel kernel of propagate.py:72 always is.  The tables keep
                              their general [P][P] form; the flag lets the three-pass kernels take the two factors out
                              of their first row / column.  0: unknown (five-pass kernels).                          */
} adm_multislice_cfg;

/* ABI / build identification. */
int adm_abi_version(void);

/* Permute + transpose + scale a [P][P] complex transfer function H (natural fftfreq order, as built by
 * adorym/propagate.py:62-81 get_kernel) into the layout the in-shared-memory mixed-radix FFT passes read.
 * out[px][py] = scale * (conj ? conj(H) : H)[perm(py)][perm(px)].  Use scale = 1/P^2 for a propagation step. */
int adm_prepare_transfer_function(int P, const float* H, int conj, float scale, float* out, void* stream);

/* Separable transfer function H[ky][kx] = hy[ky] * hx[kx] (the paraxial Fresnel kernel of propagate.py:72 is
 * separable; the non-paraxial one of :76-80 is not): hy, hx are [P][2] in natural fftfreq order;
 * out [2][P][2] = (hy | scale * hx) permuted (and conjugated if conj).  The kernels then keep H in two P-entry
 * shared-memory tables (warp-broadcast reads) instead of reading P^2 values per slice.                        */
int adm_prepare_transfer_function_separable(int P, const float* hy, const float* hx, int conj, float scale, float* out,
                                            void* stream);

/* Row pitch (in complex voxels) of the transmission volume and of its gradient: X rounded up to even, because the
 * TMA tile loads / reduce-stores need a row pitch that is a multiple of 16 bytes.                             */
#define ADM_TVOL_PITCH(X) (((X) + 1) & ~1)

/* The gradient volume dL/dT carries a margin of P voxels on the LOW side of y and x: the TMA reduce-add that
 * scatters the patch tiles rejects negative start coordinates (hardware probe: tools/tma_probe.cu), while overhang on
 * the high side is clipped by the TMA unit.  Object voxel (z, y, x) lives at [z][y + P][x + P].                 */
#define ADM_GVOL_MARGIN(P) (P)
#define ADM_GVOL_ROWS(Y, P) ((Y) + (P))
#define ADM_GVOL_PITCH(X, P) ((((X) + (P)) + 1) & ~1)

/* Transmission volume (replaces the per-slice w.exp_complex of propagate.py:241 / the complex object of :246):
 *   rot  [rows][X][2]   slice-major object volume (delta, beta) or (re, im), rows = Nz * Y
 *   tvol [rows][Xp][2]  out: T' = T - 1, T = exp(-k1 beta) (cos(s k1 delta) - i sin(s k1 delta)) for delta_beta,
 *                       T = re + i im for real_imag; pad columns x >= X are written as 0.  Xp = ADM_TVOL_PITCH(X).
 * Vacuum outside the object is T' = 0 for both object types (util.py:1335-1350).                            */
int adm_transmission_fwd(int unknown_type, float k1, int sign_convention, long long rows, int X, const float* rot,
                         float* tvol, void* stream);
/* Chain rule back to the object: g_rot [Nz][Y][X][2] (WRITTEN) from the margined gradient volume
 * g_tvol [Nz][ADM_GVOL_ROWS(Y,P)][ADM_GVOL_PITCH(X,P)][2] = (dL/dRe T, dL/dIm T) and tvol:
 *   dL/ddelta = -s k1 Im(g_T conj(T)), dL/dbeta = -k1 Re(g_T conj(T)); identity for real_imag.                */
int adm_transmission_adj(int unknown_type, float k1, int sign_convention, int Nz, int Y, int X, int P,
                         const float* g_tvol, const float* tvol, float* g_rot, void* stream);

/* Forward multislice for B*M wavefields (replaces propagate.py:131-288 + the gather of forward_model.py:313-331).
 *   tvol     [Nz][Y][Xp][2]     transmission volume T' (adm_transmission_fwd), slice-major, 16-byte aligned
 *   probe    [M][P][P][2] or [B][M][P][P][2]
 *   pos      [B][2] int32 (y, x) top-left corner of each patch in object pixels; may be negative / overhang:
 *            out-of-object voxels are vacuum (TMA zero fill of T'; util.py:1335-1350)
 *   Hprep    prepared in-object transfer function (adm_prepare_transfer_function, conj=0, scale=1/P^2)
 *   Hfree    prepared free-space transfer function for ADM_PROP_NEAR, else NULL
 *   psi_ckpt [B][M][Nz][P][P+2][2] out: wavefield entering every slice (consumed by adm_multislice_bwd), 16-byte
 *            aligned; rows are stored with the tile pitch P+2 at column offset (x & 1) so that one bulk copy brings
 *            a slice back in the layout the TMA reduce of the adjoint needs (pad columns are never written)
 *   det      [B][M][P][P][2]     out: complex wavefield on the detector (far field is fftshift-ed)
 *   workspace  adm_multislice_workspace_bytes(P) bytes, zero-filled ONCE after allocation (the kernels leave its
 *            flags zero), private to the stream, or NULL.  With a workspace and more wavefields than co-resident CTAs
 *            the launch is persistent: the B*M*Nz slice steps are cut into equal intervals, one per CTA, and a wavefield
 *            cut by a boundary is handed from CTA to CTA through the workspace (no idle tail of a partial last wave).
 *            NULL (or few wavefields): one CTA per wavefield.
 * cfg->k1 / unknown_type are not used here (they parameterise adm_transmission_*).
 * Kernel selection (same buffers and formats either way): P = 72 / 64 with h_separable, n_slices > 1 and far-field or
 * exit-wave detection run on the three-pass kernels (csrc/ms3_kernels.cuh); single-slice objects, a general transfer
 * function, near-field detection and adm_multislice_bwd_gpsi on the five / six-pass kernels (csrc/multislice.cu).
 * Environment ADM_MS3=0 selects the latter for every shape (A/B runs, tests).                                   */
long long adm_multislice_workspace_bytes(int P);
int adm_multislice_fwd(const adm_multislice_cfg* cfg, const float* tvol, const float* probe, const int* pos,
                       const float* Hprep, const float* Hfree, float* psi_ckpt, float* det, void* workspace,
                       void* stream);

/* Analytic adjoint of adm_multislice_fwd (replaces torch.autograd.grad over the taped slice loop,
 * wrappers.py:300-331).
 *   g_det     [B][M][P][P][2]   dL/d(det) as (dL/dRe, dL/dIm)
 *   HprepC    prepared conj in-object transfer function (conj=1, scale=1/P^2); HfreeC likewise or NULL
 *   g_tvol    [Nz][ADM_GVOL_ROWS(Y,P)][ADM_GVOL_PITCH(X,P)][2]  ACCUMULATED (+=) gradient w.r.t. T (dL/dRe T,
 *             dL/dIm T), margined (see ADM_GVOL_MARGIN): per slice one TMA reduce-add of the patch tile (overlapping
 *             patches are summed by the L2, overhang is clipped); zero it before the first call of a step
 *   g_probe   [M][P][P][2] ACCUMULATED over patterns, or [B][M][P][P][2] WRITTEN if probe_per_pattern; may be
 *             NULL when the probe gradient is not needed
 *   workspace as for adm_multislice_fwd (the same buffer may serve both calls of a step)                    */
int adm_multislice_bwd(const adm_multislice_cfg* cfg, const float* tvol, const int* pos, const float* HprepC,
                       const float* HfreeC, const float* psi_ckpt, const float* g_det, float* g_tvol,
                       float* g_probe, void* workspace, void* stream);
/* The same, additionally storing dL/d psi_z -- the gradient with respect to the wavefield ENTERING every slice z -- in
 * g_psi_ckpt [B][M][Nz][P][P][2] (natural order): with psi_ckpt it gives the derivative with respect to the gap between
 * slices z-1 and z (refinable slice positions of the sparse multislice model, propagate.py:479-534,
 * fresnel_propagate_wrapped :556-568):  dL/d gap = Re sum conj(g_z) IFFT2(i phi' FFT2(psi_z)), phi' = d phase / d gap.   */
int adm_multislice_bwd_gpsi(const adm_multislice_cfg* cfg, const float* tvol, const int* pos, const float* HprepC,
                            const float* HfreeC, const float* psi_ckpt, const float* g_det, float* g_tvol,
                            float* g_probe, float* g_psi_ckpt, void* workspace, void* stream);

/* Detector magnitude, mismatch loss and residual (forward_model.py:88-103,353,371-375).
 *   det [B][M][P][P][2]; meas [B][P][P] = |prj| as stored (magnitude or intensity, see raw_data_type);
 *   mask [P][P] (1 = valid pixel) or NULL (beamstop, forward_model.py:128-136)
 *   pred [B][P][P] out: sqrt(sum_m |det_m|^2);  g_det [B][M][P][P][2] out (may be NULL): dL/d(det);
 *   loss_per_pattern [B] out: sum over valid pixels of the per-pixel loss term (NOT divided);
 *   inv_n: 1 / (B * n_valid_pixels) -- the mean's divisor, applied to g_det only.                           */
int adm_detector_loss(int B, int M, int P, int loss_type, int raw_data_type, float poisson_multiplier, float inv_n,
                      const float* det, const float* meas, const float* mask, float* pred, float* g_det,
                      float* loss_per_pattern, void* stream);

/* Tomographic rotation about y (util.py:536-552 -> grid_sample bilinear / border / align_corners=False).
 *   obj  [Y][X][Z][2] (reference layout);  grid [X*Z][2] fp32 NORMALISED coordinates (gz, gx) exactly as
 *   wrappers.py:1137 builds them from the fp16 lookup table;  rot [Z][Y][X][2] (slice-major) out.          */
int adm_rotate_fwd(int Y, int X, int Z, const float* obj, const float* grid, float* rot, void* stream);
/* Exact transpose: g_obj [Y][X][Z][2] += R^T g_rot  (g_rot [Z][Y][X][2]).                                   */
int adm_rotate_adj(int Y, int X, int Z, const float* g_rot, const float* grid, float* g_obj, void* stream);

/* Object -> transmission volume in one pass (adm_rotate_fwd + adm_transmission_fwd without the intermediate rotated
 * volume): tvol [Z][Y][Xp][2] = T'( R_grid obj [Y][X][Z][2] ).  grid == NULL: no rotation, a pure
 * [Y][X][Z] -> [Z][Y][X] transposition (the object is already in the beam frame: rotate_out_of_loop,
 * ptychography.py:1011, or a single angle).  binning > 1 (propagate.py:207-254; delta_beta only): tvol has
 * ceil(Z / binning) slices, each the transmission of the SUM of (delta, beta) over `binning` rotated object slices;
 * binning == Z is the pure-projection model (propagate.py:158-186).                                              */
int adm_object_to_transmission(int Y, int X, int Z, const float* obj, const float* grid, int unknown_type, float k1,
                               int sign_convention, int binning, float* tvol, void* stream);
/* Its transpose on the planes y in [y_begin, y_end) (adm_transmission_adj + adm_rotate_adj without the intermediate):
 * g_obj [Y][X][Z][2] += R^T chain(g_tvol, tvol).  g_obj is NOT cleared: initialise it with zeros or with the
 * regulariser gradient (adm_regularizer_l1tv_range, accumulate = 0).  tvol may be NULL for ADM_REAL_IMAG.        */
int adm_transmission_to_object_adj(int Y, int X, int Z, int P, const float* g_tvol, const float* tvol, const float* grid,
                                   int unknown_type, float k1, int sign_convention, int binning, float* g_obj, int y_begin,
                                   int y_end, void* stream);

/* Fused Adam (optimizers.py:309-318) with optional epilogue.  In place on x, m, v (n floats each).
 *   step = i_batch of the reference (bias correction uses step+1); hyper-parameters are doubles (python floats):
 *   1-b1, 1-b2 and the bias corrections are formed in double and then rounded to fp32, as torch does, and the
 *   element update follows the reference's operation order with round-to-nearest fp32 ops (bit-exact vs torch CPU)
 *   flags: bit0 clip x to >= 0 (ptychography.py:1138-1139); bit1 zero channel 0 (absorption_only);
 *          bit2 zero channel 1 (phase_only) (ptychography.py:1140-1142); n must be even if bit1|bit2
 *   mask: per-voxel [n/2] multiplier applied to both channels (array_ops.py:239-244) or NULL               */
int adm_adam_step(float* x, float* m, float* v, const float* g, long long n, double step_size, double b1, double b2,
                  double eps, int step, int flags, const float* mask, void* stream);

/* Data-parallel object update over NVLink peer memory (replaces comm.allreduce + AdamOptimizer.apply_gradient,
 * ptychography.py:1113-1125, for `world` <= 8 ranks of one NVSwitch node).  x_ptrs / g_ptrs: HOST arrays of `world` DEVICE
 * pointers -- every rank's object x [n] and gradient g [n] buffers, peer-mapped into this process (symmetric memory /
 * CUDA IPC); m, v: this rank's Adam moments [n] (only its shard is touched).  Rank r pulls its shard of every g (summed in
 * rank order), applies adm_adam_step's update to its shard of x and pushes the result into every rank's x.  The caller
 * must place a cross-GPU barrier before the call (all g complete) and after it (all pushes complete).  n % 4 == 0.
 * x_multicast / g_multicast: NVLS multicast mappings of the same two symmetric buffers (both or neither; NULL = peer
 * loads / stores): the pull then is ONE multimem.ld_reduce per 16 bytes (the switch adds the G copies, fp32) and the push
 * ONE multimem.st (the switch replicates it), which halves the bytes on every link.                                  */
int adm_p2p_reduce_adam(float* const* x_ptrs, const float* const* g_ptrs, float* x_multicast, const float* g_multicast,
                        int world, int rank, float* m, float* v, long long n, double step_size, double b1, double b2,
                        double eps, int step, int flags, const float* mask, void* stream);

/* Gradient descent (optimizers.py:432-464: x <- x - step g) and momentum (optimizers.py:366-411: v <- gamma v + step g,
 * x <- x - v; v == NULL selects plain GD), in place, with the flags / mask epilogue of adm_adam_step.  The dynamic
 * step-size schedule of GDOptimizer is host arithmetic on `step_size`.                                          */
int adm_sgd_step(float* x, float* v, const float* g, long long n, double step_size, double gamma, int flags,
                 const float* mask, void* stream);

/* L1 / reweighted-L1 / TV regulariser on obj [Y][X][Z][2], delta_beta type (regularizers.py:30-110).
 *   value[0] += alpha_d*mean(w0|d|) + alpha_b*mean(w1|b|) + gamma*(TV3d(d) + TV3d(b));  g_obj += gradient.
 *   weight_l1 [Y][X][Z][2] or NULL; value / g_obj may be NULL.                                             */
int adm_regularizer_l1tv(int Y, int X, int Z, const float* obj, const float* weight_l1, float alpha_d, float alpha_b,
                         float gamma, float* value, float* g_obj, void* stream);
/* Same, with the gradient multiplied by the DEVICE scalar *g_scale (the upstream gradient autograd hands to the
 * regulariser node, wrappers.py:322; NULL = 1) and written (accumulate = 0) or added (accumulate = 1) to g_obj.  */
int adm_regularizer_l1tv_scaled(int Y, int X, int Z, const float* obj, const float* weight_l1, float alpha_d,
                                float alpha_b, float gamma, float* value, float* g_obj, const float* g_scale,
                                int accumulate, void* stream);

/* Same on the planes y in [y_begin, y_end) only (value and gradient restricted to those planes; the TV neighbours are
 * read from the whole object): lets the driver produce the object gradient slab by slab under the all-reduce.     */
int adm_regularizer_l1tv_range(int Y, int X, int Z, const float* obj, const float* weight_l1, float alpha_d, float alpha_b,
                               float gamma, float* value, float* g_obj, const float* g_scale, int accumulate, int y_begin,
                               int y_end, void* stream);

/* Sub-pixel Fourier shift of the probe per pattern (util.py:380-397 as used at forward_model.py:296-303).
 *   probe [M][P][P][2]; shift [B][2] (y, x); out [B][M][P][P][2].                                          */
int adm_probe_shift_fwd(int B, int M, int P, const float* probe, const float* shift, float* out, void* stream);
/* Transpose: g_probe [M][P][P][2] += sum_b S_b^H g_out[b];  g_shift [B][2] += dL/d(shift) (accumulated over modes; zero it first).        */
int adm_probe_shift_bwd(int B, int M, int P, const float* probe, const float* shift, const float* g_out,
                        float* g_probe, float* g_shift, void* stream);

/* ---------------------------------------------------------------------------------------------------------------
 * Large-frame near-field path: MultiDistModel (adorym/forward_model.py:819-1092) on frames of N = 128 .. 2048 pixels
 * per side (power of two), one thin object slice, D propagation distances, M probe modes.  Replaces, per distance and
 * mode, multislice_propagate_batch with n_slices == 1 + fresnel_propagate (propagate.py:131-288, :537-553).
 * --------------------------------------------------------------------------------------------------------------- */
typedef struct adm_multidist_cfg {
    int N;               /* frame edge; adm_lf_supported(N) != 0                                              */
    int M;               /* probe modes                                                                        */
    int D;               /* propagation distances                                                              */
    float k1;            /* as adm_multislice_cfg.k1                                                           */
    int sign_convention; /* +1 / -1                                                                            */
    int unknown_type;    /* ADM_DELTA_BETA / ADM_REAL_IMAG                                                     */
} adm_multidist_cfg;

/* 1 if frames of edge N are handled by the large-frame kernels. */
int adm_lf_supported(int N);
/* HOST array freq_of_pos[N]: numpy-fftfreq index held at each position of the (digit-reversed) spectrum the large-frame
 * FFT passes leave in place.  A transfer function H[ky][kx] (propagate.py:62-81) is prepared for the kernels as
 * Hprep[py][px] = H[freq_of_pos[py]][freq_of_pos[px]] / N^2 (conjugated for the adjoint).                     */
int adm_lf_spectrum_order(int N, int* freq_of_pos);
/* det[d][m] = IFFT2(H_d . FFT2(probe_m . T(obj)))  for all d, m.
 *   obj [N][N][2]; probe [M][N][N][2]; Hprep [D][N][N][2]; work1 [M][N][N][2]; det [D][M][N][N][2] out.       */
int adm_multidist_fwd(const adm_multidist_cfg* cfg, const float* obj, const float* probe, const float* Hprep,
                      float* work1, float* det, void* stream);
/* Analytic adjoint: g_det [D][M][N][N][2] -> g_obj [N][N][2] WRITTEN (chain rule to (delta, beta) included; may be
 * NULL), g_probe [M][N][N][2] WRITTEN (may be NULL).  HprepC = conj of Hprep; work1 [M][N][N][2], work2 [D][M][N][N][2]. */
int adm_multidist_bwd(const adm_multidist_cfg* cfg, const float* obj, const float* probe, const float* HprepC,
                      const float* g_det, float* work1, float* work2, float* g_obj, float* g_probe, void* stream);
/* Sub-pixel Fourier shift of measured frames (util.py:380-397 as used on the data at forward_model.py:1079-1088):
 *   out[f] = Re IFFT2(FFT2(data[f]) exp(-2 pi i (fx sx + fy sy))), shift [frames][2] = (sy, sx).
 *   order [N] int32 DEVICE copy of adm_lf_spectrum_order; work [frames][N][N][2]; spec [frames][N][N][2] out (the shifted
 *   spectrum, consumed by adm_frame_shift_bwd); out [frames][N][N].                                           */
int adm_frame_shift_fwd(int N, int frames, const float* data, const float* shift, const int* order, float* work,
                        float* spec, float* out, void* stream);
/* g_shift [frames][2] += d/d(shift) of sum(g_out . out).                                                      */
int adm_frame_shift_bwd(int N, int frames, const float* g_out, const float* spec, const int* order, float* work,
                        float* g_shift, void* stream);
/* adm_detector_loss for large frames (many CTAs per frame, deterministic two-stage sum) that also returns the
 * derivative with respect to the measured frame (needed when the data are shifted by a refinable offset):
 *   det [B][M][npix][2]; meas [B][npix] (|.| is applied, forward_model.py:1062); pred [B][npix] out;
 *   g_det [B][M][npix][2] out or NULL; g_meas [B][npix] out or NULL; partial [B][64] workspace;
 *   loss_per_pattern [B] out (sum, NOT divided); inv_n = 1 / (B npix) applied to the gradients only.          */
int adm_detector_loss_large(int B, int M, long long npix, int loss_type, int raw_data_type, float poisson_multiplier,
                            float inv_n, const float* det, const float* meas, float* pred, float* g_det, float* g_meas,
                            float* partial, float* loss_per_pattern, void* stream);

/* ---- affine resampling of a stack of planes (tilt refinement, affine alignment of the measured frames) -------------
 * Replaces adorym/wrappers.py:1158-1174 `affine_transform` (forward_model.py:1066-1072, optimize_prj_affine) and
 * wrappers.py:1177-1217 `rotate` (forward_model.py:274-276, optimize_tilt): F.affine_grid(m, align_corners=False) +
 * F.grid_sample(bilinear, padding_mode='border', align_corners=False), and what autograd derives from them.
 *   The array is dense [D0][D1][D2][C] fp32 (C = 1 or 2); dim_n / dim_h / dim_w in {0,1,2} say which memory dimension is
 *   the plane index N, the rows H and the columns W.  mats: row-major 2x3 matrices, [N][6] (mat_per_plane = 1) or [6].
 *   _fwd : out = resampled array (same layout as in).
 *   _adj : g_in += transpose applied to g_out (caller zero-fills g_in).
 *   _dmat: g_mats [N][6] += d/d(mats) of sum(g_out . out) per plane (caller zero-fills; sum over N for a shared matrix). */
int adm_affine_resample_fwd(int D0, int D1, int D2, int C, int dim_n, int dim_h, int dim_w, const float* in,
                            const float* mats, int mat_per_plane, float* out, void* stream);
int adm_affine_resample_adj(int D0, int D1, int D2, int C, int dim_n, int dim_h, int dim_w, const float* g_out,
                            const float* mats, int mat_per_plane, float* g_in, void* stream);
int adm_affine_resample_dmat(int D0, int D1, int D2, int C, int dim_n, int dim_h, int dim_w, const float* in,
                             const float* g_out, const float* mats, int mat_per_plane, float* g_mats, void* stream);

/* ---- ptychography with probes of 128 .. 2048 pixels (power of two): batched slice steps on the large-frame kernels ----
 * One step of the slice loop (propagate.py:208-258) for B*M wavefields, pattern b reading its patch (origin pos[b], vacuum
 * outside the object: util.py:1327-1351, forward_model.py:313-331) of slice z of the transmission volume
 * tvol [Nz][Y][ADM_TVOL_PITCH(X)][2]:   psi_out = IFFT2(Hprep . FFT2(psi_in (1 + T'_patch))).
 *   Hprep == NULL: modulation only (the last slice is not propagated, propagate.py:253); work is then unused.
 *   psi_in / psi_out / work [B*M][N][N][2]; Hprep [N][N][2] prepared as for adm_multidist_fwd (permuted, / N^2).        */
int adm_lf_patch_step_fwd(int N, int B, int M, int Y, int X, int z, const float* tvol, const int* pos, const float* psi_in,
                          const float* Hprep, float* work, float* psi_out, void* stream);
/* Adjoint of the step: g_psi [B*M][N][N][2] WRITTEN (may be NULL); g_tvol [Nz][Y + g_margin][ADM_GVOL_PITCH(X, g_margin)][2]
 * ACCUMULATED at slice z (dL/dT = g conj(psi_in), red.add: patches overlap; may be NULL).  HprepC = conj of Hprep.       */
int adm_lf_patch_step_bwd(int N, int B, int M, int Y, int X, int z, const float* tvol, const int* pos, const float* psi_in,
                          const float* HprepC, const float* g_out, float* work1, float* work2, float* g_psi, float* g_tvol,
                          int g_margin, void* stream);
/* Far-field detection of large frames (propagate.py:263-270, sign_convention +1): out = scale . fftshift(FFT2(in)), natural
 * order; in / work / out [frames][N][N][2]; order = DEVICE copy of adm_lf_spectrum_order.  _bwd: its adjoint.           */
int adm_lf_farfield_fwd(int N, int frames, const float* in, const int* order, float scale, float* work, float* out,
                        void* stream);
int adm_lf_farfield_bwd(int N, int frames, const float* g_out, const int* order, float scale, float* work, float* g_in,
                        void* stream);
/* Sub-pixel Fourier shift of COMPLEX frames (the per-position probe shift of forward_model.py:296-303 for large probes):
 * out[f] = IFFT2(FFT2(in[f]) exp(-2 pi i (fx sx + fy sy))); shift [frames][2] = (sy, sx); spec = shifted spectrum kept
 * for adm_frame_shift_complex_bwd, which ACCUMULATES g_shift [frames][2] (the gradient with respect to `in` is the same
 * call with -shift applied to g_out).                                                                                   */
int adm_frame_shift_complex_fwd(int N, int frames, const float* in, const float* shift, const int* order, float* work,
                                float* spec, float* out, void* stream);
int adm_frame_shift_complex_bwd(int N, int frames, const float* g_out, const float* spec, const int* order, float* work,
                                float* g_shift, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* ADORYM_B200_H */
