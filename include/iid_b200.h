/*
 * iid_b200.h -- C ABI of the B200-native elastic-scattering (Debye sum) hot path.
 *
 * This is the drop-in boundary for the three callables that pyIID's
 * ElasticScatter.set_processor binds (reference:
 * pyiid/experiments/elasticscatter/__init__.py:206-292 -> self.fq, self.grad,
 * self.grad_pdf) and for the Rw / chi^2 energy+forces that pyiid.calc.Calc1D
 * derives from them (pyiid/calc/calc_1d.py:78-95, pyiid/calc/__init__.py:10-105).
 *
 * Conventions
 *  - extern "C", plain pointers and sizes, no C++/torch types.
 *  - every function returns int: 0 = ok, negative = IID_E_* below, positive =
 *    a cudaError_t.  iid_last_error() returns a thread-local message.
 *  - "dev" pointers are device pointers on the handle's device; "host"
 *    pointers are ordinary host memory.  `stream` is a cudaStream_t passed as
 *    void* (NULL = the handle's own stream).  Device-pointer calls only
 *    ENQUEUE work; host-pointer calls (suffix _host) copy in, run, copy out
 *    and synchronise before returning.
 *  - precision: IID_FP32 computes the pair sums in float32 arithmetic the way
 *    the reference does (float32-rounded positions, float32 accumulators,
 *    float64 for the F(Q) reduction); IID_FP64 computes everything in float64.
 *  - a handle is used by one thread at a time.
 *  - there is no CPU fallback anywhere behind this interface.
 */
#ifndef IID_B200_H
#define IID_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define IID_FP32 0
#define IID_FP64 1

#define IID_POT_RW 0      /* master_kernel.get_rw      (:206-236) */
#define IID_POT_CHI_SQ 1  /* master_kernel.get_chi_sq  (:239-266) */

#define IID_E_BADARG (-1)
#define IID_E_NOSTRUCT (-2)   /* iid_set_structure not called yet      */
#define IID_E_NOTRANSFORM (-3)/* iid_set_transform not called yet      */
#define IID_E_NOMEM (-4)
#define IID_E_NODEVICE (-5)   /* no CUDA device / wrong architecture   */
#define IID_E_NOCHAIN (-6)    /* iid_leapfrog_chain_next: that chain was dropped */

typedef struct iid_handle iid_handle;

/* library / device ------------------------------------------------------- */
int iid_version(void);
const char *iid_last_error(void);
int iid_device_count(int *count);
/* sm count, clock (kHz), compute capability major*10+minor */
int iid_device_info(int device, int *sm_count, int *clock_khz, int *cc);

/* handle ------------------------------------------------------------------ */
/* Owns one stream, the staged (element-sorted) atom arrays, the work-item
 * lists and all scratch for one GPU.  Replaces the per-call allocation done
 * by atomics/gpu_atomics.py:89-280 and the thread-per-GPU farm of
 * gpu_wrappers/gpu_wrap.py:287-314. */
int iid_create(int device, int precision, iid_handle **out);
int iid_destroy(iid_handle *h);
int iid_get_stream(iid_handle *h, void **stream);
int iid_synchronize(iid_handle *h);

/* One handle over several GPUs of the box, in ONE process: what
 * set_processor('Multi-GPU') gives in the reference (the thread-per-GPU farm
 * gpu_wrappers/gpu_wrap.py:119-156, 287-314 driven by gpu_multithreading).
 * n_devices <= 0 = every visible device.  The handle owns one sub-handle per
 * device; device d computes shard (d, active) of every pass, `active` =
 * min(n_devices, atoms / 1500) per structure.  The F(Q) pair sums and the
 * force array are summed on the host in device order; gradient rows are
 * disjoint per device and are written straight into the caller's array.
 * Every host-buffer entry point accepts such a handle (the small float64
 * stages, springs and the device-resident sampler run on device 0; the
 * sampler states need a structure small enough for one device); the
 * device-pointer (enqueue-only) pair-sum calls and iid_set_shard return
 * IID_E_BADARG. */
int iid_create_multi(int n_devices, int precision, iid_handle **out);
/* devices behind a handle (1 for iid_create) and how many the current
 * structure uses */
int iid_handle_devices(iid_handle *h, int *n_devices, int *active);

/* Which slice of the work this handle computes (one process per GPU; the
 * caller sums the partial F(Q) / force outputs): work items rank, rank+world,
 * ... of the pair-triangle list (F(Q), force), and the gradient ROWS of the
 * i-tiles rank, rank+world, ... (full gradient: every row is computed by one
 * rank only, so there is no gradient collective).  Default (0, 1).  Replaces
 * gpu_wrap.gpu_multithreading for the one-process-per-GPU launch. */
int iid_set_shard(iid_handle *h, int rank, int world);

/* Structure = what stays fixed while positions move: per-atom element index
 * type_index[n] in [0, n_types) and the per-ELEMENT form-factor table
 * ftable[n_types][nq] sampled at Q_m = m*qbin.  Replaces the per-atom
 * `atoms.arrays['F(Q) scatter' | 'PDF scatter']` [n, nq] float32 arrays that
 * wrap_fq / wrap_fq_grad pull from the Atoms object
 * (cpu_wrappers/flat_multi_cpu_wrap.py:11-19) and re-upload per chunk
 * (atomics/gpu_atomics.py:143,248).  The normaliser
 * na[m] = N * mean_pairs(f_i f_j) (flat_multi_cpu_wrap.py:52-55) is computed
 * here in float64 from its closed form.  Host pointers. */
int iid_set_structure(iid_handle *h, int64_t n, const int32_t *type_index,
                      int64_t n_types, const double *ftable, int64_t nq,
                      double qbin);

/* The same with a second table norm_table[n_types][nq] (NULL = ftable) for the
 * normaliser: the pair sums use ftable, na uses norm_table.  For per-atom
 * factors that multiply the pair term but not <f>^2 -- the Debye-Waller factor
 * of isotropic atomic displacements, tau_ij(Q) = exp(-(u_i^2 + u_j^2) Q^2 / 2)
 * = t_i t_j, in the reference's fq = norm * omega * tau
 * (kernels/cpu_nxn.py:114-121 get_adp_fq, kernels/cpu_flat.py:156-174
 * get_adp_grad_fq with grad_tau = 0): ftable rows = f t, norm_table rows = f,
 * one row per (element, displacement) class. */
int iid_set_structure_norm(iid_handle *h, int64_t n, const int32_t *type_index,
                           int64_t n_types, const double *ftable,
                           const double *norm_table, int64_t nq, double qbin);

/* F(Q) -> G(r) as a dense real matrix T[nr][nq] (float64, host) reproducing
 * master_kernel.get_pdf_at_qmin :39-104 (zero below qmin, zero-pad, odd
 * extension, inverse FFT, linear re-binning onto rgrid, factor 2); the host
 * layer builds it once per experiment.  Needed by iid_fq_to_gr, iid_potential,
 * iid_energy_forces*. */
int iid_set_transform(iid_handle *h, int64_t nr, int64_t nq, const double *T);

/* Host-only: the Lagrange weights w[n_points] of the radial stencil for a
 * point u in [0, 1) past grid node k (node k - left + i gets w[i]) and the grid
 * step in units of 1/Q_max.  The fused evaluation's force table interpolates
 * with this stencil, its F(Q) phase and the F(Q) pair histogram of structures
 * beyond the fine grid spread with it; w must hold at least 16 doubles. */
int iid_stencil_weights(double u, double *w, int *n_points, int *left, double *qmax_h);

/* Host-only: the same for the stencils with which the F(Q) pair histogram of a
 * large structure spreads -- the shortest stencil / finest grid the structure
 * fits: tier 0 = 6 points, Q_max h = 0.0658; tier 1 = 8 points, Q_max h = 0.157;
 * tier 2 = the 12 points above.  All three keep the 4e-10 bound; the pass costs
 * two shared-memory atomics per point and pair. */
int iid_hist_stencil_weights(int tier, double u, double *w, int *n_points, int *left,
                             double *qmax_h);

/* Host-only (no device): the sharding plan iid_set_structure + iid_set_shard
 * would produce -- total work items, this rank's items and the (i, j) slots
 * they cover, for the triangle (F(Q), force) or square (gradient) list. */
int iid_plan_shard(int64_t n, const int32_t *type_index, int64_t n_types,
                   int sm_count, int triangle, int rank, int world,
                   int64_t *n_items_total, int64_t *n_items_mine,
                   int64_t *pair_slots_mine, int64_t *padded_atoms);

/* Host-only: the row jobs of the full-gradient pass that rank `rank` of
 * `world` runs (iid_set_structure + iid_set_shard build the same): 4 int32 per
 * job (itile, seg_begin, seg_end, dest: -1 = stores its rows into G, >= 0 =
 * piece slot of a split row), per segment (jbegin, jend, info, 0) with info =
 * element type | 1<<16 diagonal tile | 1<<17 gradient only | 1<<18 flush, per
 * split row (itile, first slot, one past the last slot, 0).  Output arrays may
 * be NULL (counts only); piece_div <= 0 = default. */
int iid_plan_rows(int64_t n, const int32_t *type_index, int64_t n_types,
                  int sm_count, int rank, int world, int piece_div,
                  int64_t *n_jobs, int64_t *n_segs, int64_t *n_fixes,
                  int32_t *jobs, int64_t jobs_cap, int32_t *segs,
                  int64_t segs_cap, int32_t *fixes, int64_t fixes_cap);

/* sizes ------------------------------------------------------------------- */
int iid_get_sizes(iid_handle *h, int64_t *n, int64_t *nq, int64_t *nr,
                  int64_t *n_items_fq, int64_t *n_items_grad);

/* pair sums (device pointers; enqueue only) ------------------------------- */
/* pos_dev: [n,3] float64 in the caller's atom order (rounded to float32
 * inside when the handle is IID_FP32, as wrap_fq does :11-12).
 *
 * iid_fq_partial: S[m] = sum over this shard's unordered pairs of
 *   f_i f_j sin(Q_m r_ij)/r_ij  -> S_dev[nq] float64 (overwritten).
 *   Replaces atomic_fq (atomics/cpu_atomics.py:60-78, gpu_atomics.py:89-189).
 * iid_fq_finish: F[m] = 2 S[m] / na[m], 0 where na == 0
 *   (flat_multi_cpu_wrap.py:49-60) -> F_dev[nq] float64. */
int iid_fq_partial(iid_handle *h, const double *pos_dev, double *S_dev,
                   void *stream);
int iid_fq_finish(iid_handle *h, const double *S_dev, double *F_dev,
                  void *stream);

/* iid_grad_fq_partial: this shard's ROWS of the NORMALISED gradient in the
 * reference's convention (SURVEY.md section 8a note 1)
 *   G[i,w,m] = (1/na[m]) sum_j f_i f_j a_ij(m) (q_j - q_i)_w,
 *   a = (Q cos(Q r) - sin(Q r)/r) / r^2
 * into G_dev[n][3][nq] (float32 for IID_FP32, float64 for IID_FP64).  Row
 * ownership: a block walks one i-tile against ALL j and stores each row once
 * (plain coalesced stores: no zero-fill, no atomics, bit-reproducible); rows
 * of atoms owned by other shards are NOT touched.  G_dev may be device memory
 * or the device alias of pinned, mapped host memory (iid_host_alloc /
 * iid_host_register + iid_host_device_pointer).  S_dev (may be NULL) receives
 * this shard's F(Q) pair sums as iid_fq_partial does (summed over the row jobs
 * in a fixed order).  Replaces atomic_grad_fq (cpu_atomics.py:81-102,
 * gpu_atomics.py:192-280) + the /na of flat_multi_cpu_wrap.py:93-102. */
int iid_grad_fq_partial(iid_handle *h, const double *pos_dev, void *G_dev,
                        double *S_dev, void *stream);

/* iid_force_partial: force[i,w] = sum_m wq[m] * G[i,w,m] without forming G:
 * one scalar per pair, sum_m wq[m] f_i f_j a_ij(m) / na[m], walked over the
 * pair triangle.  wq_dev[nq] float64, force_dev[n][3] float64 (overwritten
 * with this shard's partial).  Replaces get_grad_pdf + get_grad_rw's
 * contraction (master_kernel.py:276-347) without the N x 3 x R array. */
int iid_force_partial(iid_handle *h, const double *pos_dev,
                      const double *wq_dev, double *force_dev, void *stream);

/* float64 small stages (device pointers; enqueue only) -------------------- */
/* G[nr] = T F  (master_kernel.get_pdf_at_qmin) */
int iid_fq_to_gr(iid_handle *h, const double *F_dev, double *G_dev,
                 void *stream);
/* Rw or chi^2 of gcalc against target (get_rw / get_chi_sq), the scale, and
 * the chain-rule weight vector wq[m] = conv * sum_r c_r T[r][m] with c from
 * get_grad_rw / get_grad_chi_sq (master_kernel.py:293-375).
 * out_dev[4] = {energy = value*conv, scale, raw value, 0}. */
int iid_potential(iid_handle *h, const double *G_dev, const double *target_dev,
                  int potential, double conv, double *out_dev, double *wq_dev,
                  void *stream);
/* grad_pdf[rows][nr] = grad_fq[rows][nq] . T^T  (master_kernel.grad_pdf
 * :276-290 / gpu_wrap.grad_pdf :197-284): one row per (atom, direction).
 * grad_fq_dev is float32 (IID_FP32) or float64; output float64. */
int iid_grad_pdf(iid_handle *h, const void *grad_fq_dev, int64_t rows,
                 double *grad_pdf_dev, void *stream);

/* host-buffer entry points (copy in, compute, copy out, synchronise) ------ */
/* These are what a ctypes/numpy caller binds; single shard or partial
 * results when a shard is set.  F_host[nq], G_host[n*3*nq] (float32 or
 * float64 by precision), pdf_host[nr], forces_host[n*3] are float64 unless
 * noted. */
int iid_fq_host(iid_handle *h, const double *pos_host, double *F_host);
/* G_host in pinned, mapped memory (iid_host_alloc, iid_host_register) is
 * written by the kernel directly (no device copy, no download); any other
 * host pointer is filled through pipelined pinned staging. */
int iid_grad_fq_host(iid_handle *h, const double *pos_host, void *G_host,
                     double *F_host);
/* Pinned, mapped, portable host memory for output arrays; registration of an
 * existing host mapping (e.g. POSIX shared memory that the ranks of a
 * one-process-per-GPU job all write their gradient rows into -- the one host
 * array gpu_wrap.py:159-194 assembles); the device alias to pass as G_dev. */
int iid_host_alloc(int64_t bytes, void **ptr);
int iid_host_free(void *ptr);
int iid_host_register(void *ptr, int64_t bytes);
int iid_host_unregister(void *ptr);
int iid_host_device_pointer(void *host, void **dev);
int iid_pdf_host(iid_handle *h, const double *pos_host, double *pdf_host,
                 double *F_host);
/* One call per HMC leapfrog: energy = potential(G(r), target)*conv and
 * forces = grad(...)*conv exactly as Calc1D.calculate_energy/forces
 * (calc/calc_1d.py:78-95) with exp_function = get_pdf and
 * exp_grad_function = get_grad_pdf.  out_host[4] as iid_potential. */
int iid_energy_forces_host(iid_handle *h, const double *pos_host,
                           const double *target_host, int potential,
                           double conv, double *out_host,
                           double *forces_host, double *pdf_host);

/* Rw / chi^2 of two host vectors (calc/__init__.py wrap_rw :10-31,
 * wrap_chi_sq :33-54) plus the chain-rule vector c[len] with
 * grad[i,w] = sum_k c[k] dcalc[i,w,k] (wrap_grad_rw :56-79, wrap_grad_chi_sq
 * :82-105; master_kernel.py:293-375).  out_host[4] as iid_potential. */
int iid_rw_host(iid_handle *h, const double *gcalc_host, const double *gobs_host,
                int64_t len, int potential, double conv, double *out_host,
                double *c_host);
/* out[row] = sum_k A[row][k] c[k], A float32 (a_is_f32) or float64, host. */
int iid_contract_host(iid_handle *h, const void *A_host, int a_is_f32,
                      int64_t rows, int64_t len, const double *c_host,
                      double *out_host);
/* G(r) = T F for a host F(Q) (get_pdf's noise branch,
 * elasticscatter/__init__.py:371-390). */
int iid_fq_to_gr_host(iid_handle *h, const double *F_host, double *pdf_host);

/* --- Spring restraints (reference calc/spring_calc.py) ---------------------
 * The restraint calculators a refinement sums with the Rw / chi^2 potential
 * (calc/multi_calc.py:58-83).  sp_type: rep = pairs closer than rt repel
 * (spring_nrg / spring_force :107-147), att = pairs further than rt attract
 * (:269-311), com = atoms further than rt from the centre of mass `com`
 * (host, 3 doubles, COM only; :188-236).  Energies are sums over ORDERED pairs
 * as in the reference.  FP32 handles reproduce the reference's float32
 * arithmetic per pair; FP64 handles use float64 throughout.  These calls do
 * not need iid_set_structure. */
#define IID_SPRING_REP 0
#define IID_SPRING_COM 1
#define IID_SPRING_ATT 2
#define IID_MAX_RESTRAINTS 4
/* This rank's share (rows of 128 atoms rank, rank+world, ...) of energy[1],
 * force[n,3] and atomwise[n] (atomwise_spring_nrg :171-185; any may be NULL);
 * the outputs are zeroed first, partial results are summed by the caller. */
int iid_spring_partial(iid_handle *h, const double *pos_dev, int64_t n,
                       int sp_type, double k, double rt, const double *com,
                       double *energy_dev, double *force_dev,
                       double *atomwise_dev, void *stream);
/* Host arrays in and out; world must be 1. */
int iid_spring_host(iid_handle *h, const double *pos_host, int64_t n,
                    int sp_type, double k, double rt, const double *com,
                    double *energy_host, double *forces_host,
                    double *atomwise_host);
/* Energy a probe atom adds at each voxel centre ((i+.5) resolution, C order
 * [nx,ny,nz]); voxel_spring_nrg :150-168, :240-256, :314-332. */
int iid_spring_voxel_host(iid_handle *h, const double *pos_host, int64_t n,
                          int sp_type, double k, double rt, const double *com,
                          double resolution, int64_t nx, int64_t ny,
                          int64_t nz, double *voxels_host);
/* rep / att restraints evaluated inside iid_energy_forces_host (same CUDA
 * graph): their forces are added to forces_host, their summed energy is read
 * with iid_get_restraint_energy after the call.  count = 0 clears. */
int iid_set_restraints(iid_handle *h, int count, const int *sp_type,
                       const double *k, const double *rt);
int iid_get_restraint_energy(iid_handle *h, double *energy);

/* Device-resident sampler states (pyiid/sim/__init__.py:10-38 leapfrog;
 * nuts_hmc.py:15-88 buildtree keeps one Atoms object per tree node).  A state
 * is a numbered slot (q[n,3], p[n,3], f[n,3] float64) on the device.
 * iid_sampler_setup allocates n_slots zeroed slots and stores the masses [n]
 * and the cell centre [3] that Atoms.center() moves the bounding box to. */
int iid_sampler_setup(iid_handle *h, int64_t n_slots, const double *masses_host,
                      const double *cell_centre);
int iid_state_upload(iid_handle *h, int slot, const double *q_host,
                     const double *p_host, const double *f_host);
/* any of the three outputs may be NULL */
int iid_state_download(iid_handle *h, int slot, double *q_host, double *p_host,
                       double *f_host);
/* One leapfrog step src -> dst entirely on the device: p += step/2 f,
 * q += step p/m, energy + forces at the new q (as iid_energy_forces_host,
 * fused restraints included), p += step/2 f, optional centring.  Returns
 * out_host[9] = potential energy, scale, -, -, restraint energy, kinetic
 * energy, centring shift x y z, and (optionally) the new q and p.
 * target_host NULL = the target already resident.  world must be 1. */
int iid_leapfrog_host(iid_handle *h, int src, int dst, double step, int centre,
                      const double *target_host, int potential, double conv,
                      double *out_host, double *q_host, double *p_host);
/* n_steps (<= IID_LF_CHAIN) consecutive steps src -> dst[0] -> dst[1] -> ...
 * of the same size: what buildtree (nuts_hmc.py:15-88) asks for when it grows
 * a subtree of depth j (2^j leapfrogs in a row).  Structures small enough for
 * the fused evaluation walk the whole chain inside ONE cooperative launch.
 * out_host[n_steps][9], q_host / p_host [n_steps][n*3] (may be NULL) as
 * iid_leapfrog_host, one row per step. */
#define IID_LF_CHAIN 64
int iid_leapfrog_chain_host(iid_handle *h, int src, const int *dst, int n_steps,
                            double step, int centre, const double *target_host,
                            int potential, double conv, double *out_host,
                            double *q_host, double *p_host);
/* The same chain in two halves: _begin enqueues it and returns at once, _next
 * hands out the steps in order as the device completes them (out_host[9],
 * q_host / p_host [n*3] or NULL), so the caller's work on step i -- the U-turn
 * test, the tree bookkeeping -- overlaps the computation of step i + 1.
 * Steps not collected are dropped by the next call on the handle (which first
 * waits for the chain); _next for a chain that is not in flight any more
 * returns IID_E_NOCHAIN. */
int iid_leapfrog_chain_begin(iid_handle *h, int src, const int *dst, int n_steps,
                             double step, int centre, const double *target_host,
                             int potential, double conv, int64_t *chain_id);
int iid_leapfrog_chain_next(iid_handle *h, int64_t chain_id, double *out_host,
                            double *q_host, double *p_host);

/* Device array -> pageable host memory through pipelined pinned staging,
 * ordered after the work already enqueued on the handle's stream; complete on
 * return.  (The multi-GPU host layer uses it after the NCCL all-reduce.) */
int iid_download_host(iid_handle *h, const void *dev, void *host, int64_t bytes);

/* Tunables: "force_table" (0/1: tabulated radial force pass in FP32 mode),
 * "force_table_min_n" (atoms from which it is used), "graph" (0/1: CUDA-graph
 * replay of iid_energy_forces_host / iid_leapfrog_host), "cheb" (0/1:
 * three-term recurrence), "qspace_wq" (0/1: chain-rule weights in Q space),
 * "nw_max" (warps per block), "grad_nw_max" (warps per full-gradient block),
 * "grad_split" (0/1: F(Q) summed below the diagonal only, gradient-only bin
 * loop above it), "prod_unroll" (0/1: two pair set-ups per producer
 * iteration), "piece_div" (smallest piece of a split gradient row = row /
 * piece_div), "zero_copy" (0/1: gradient rows straight into mapped host
 * arrays), "acc_j" (FP32 full gradient: j atoms per float32 partial sum before
 * it is parked and re-started, 0 = one accumulator over the whole row),
 * "fused" (0/1: small structures evaluate in ONE cooperative launch),
 * "fused_table" (0/1: that launch takes its force pass from a float64 radial
 * table it builds itself, up to two element types; 0 = direct pass over the
 * Q bins), "fused_hist" (0/1: that launch sums F(Q) through a fixed-point
 * radial pair histogram in shared memory when the structure's bounding box
 * fits it, bit-reproducible; 0 = its direct pass over the Q bins),
 * "chain_in_kernel" (0/1: a leapfrog chain is ONE launch; 0 = one
 * launch per step behind one synchronisation), "fused_det" (accepted, no
 * effect: the launch's fixed-point sums are always bit-reproducible),
 * "fq_hist" (0/1: FP32 mode, the F(Q)-only pass of structures of at least
 * "fq_hist_min_n" atoms goes through a fixed-point radial pair histogram,
 * O(N^2 + K Q), bit-reproducible; 0 = the direct O(N^2 Q) kernel),
 * "det_fq" (0/1: the stand-alone F(Q) pass stores per-item
 * partial sums and adds them in item order -- F(Q), G(r), Rw reproducible).  Defaults can also be set with IID_* environment variables
 * before iid_create. */
int iid_set_option(iid_handle *h, const char *key, int64_t value);

/* instrumentation ---------------------------------------------------------- */
/* number of kernels this handle has launched since creation */
int iid_launch_count(iid_handle *h, int64_t *count);
/* device time (ms, CUDA events on the launching stream) of the last pair-sum
 * kernel launched by this handle, and its algorithmic pair*Q count */
int iid_last_kernel_ms(iid_handle *h, float *ms, double *pairq);
int iid_set_timing(iid_handle *h, int enabled);
/* Measured issue rates (micro-benchmarks, ~30 ms) of the pipes that bound the
 * pair sums, in lane-FMAs per second: out[0] scalar FFMA, out[1] packed FFMA2,
 * out[2] DFMA -- the denominators of the per-mode roofline fractions. */
int iid_measure_peaks(iid_handle *h, double *out);

#ifdef __cplusplus
}
#endif
#endif /* IID_B200_H */
