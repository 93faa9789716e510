/* ahkab_b200 — C-ABI of the B200-native batched Newton/MNA engine.
 *
 * The reference (ahkab v0.18, pure Python) has no FFI: the boundary this
 * library sits behind is the analysis dispatch dict
 *     analysis = {'op': dc_analysis.op_analysis, 'dc': dc_analysis.dc_analysis,
 *                 'tran': transient.transient_analysis, 'ac': ac.ac_analysis}
 * (ahkab/ahkab.py:860-863, invoked at ahkab/ahkab.py:714).  Each entry point
 * below replaces one of those callees for a *batch* of B independent
 * parameter instances of one circuit; INTEGRATION.md shows the ctypes stub a
 * maintainer would put behind that dict.
 *
 * Conventions
 *  - plain C: pointers and sizes only, no torch / C++ types.
 *  - `const double* vary`, x0, and every output pointer are DEVICE pointers
 *    owned by the caller; small descriptor arrays (circuit, options, sweep
 *    values, frequencies, vary_slots) are HOST pointers copied during the call.
 *  - every call is asynchronous on `stream` (a cudaStream_t passed as void*;
 *    NULL = the legacy default stream) and never synchronises the device.
 *  - return value: 0 = launched, <0 = argument / launch error (text from
 *    ahk_last_error()).  Per-instance solver failures are NOT errors: they are
 *    reported in the status[] words (AHK_ST_* bits) — the reference returns
 *    None / raises per circuit (dc_analysis.py:678-680, transient.py:425-426,
 *    transient.py:467-469), a batch cannot.
 *  - unknown ordering of every x vector: node voltages 1..nv-1 (reference
 *    node numbering, ground removed) then the currents of the voltage-defined
 *    elements (V, E, H, L) in circuit order — ahkab/dc_analysis.py:949-1069 +
 *    ahkab/utilities.py:99-131.
 *  - all arithmetic is IEEE float64 (complex128 for AC).
 */
#ifndef AHKAB_B200_H
#define AHKAB_B200_H

#ifdef __CUDACC_RTC__   /* NVRTC has no host headers */
typedef int int32_t;
typedef long long int64_t;
#else
#include <stdint.h>
#endif

#ifdef __cplusplus
extern "C" {
#endif

#define AHK_ABI_VERSION 1

/* ---- element types ------------------------------------------------------ */
enum {
  AHK_R = 1,   /* resistor            params: [R]                          */
  AHK_C = 2,   /* capacitor           params: [C, ic]                      */
  AHK_L = 3,   /* inductor (vde)      params: [L, ic]                      */
  AHK_K = 4,   /* inductor coupling   params: [K]  ref/ref2 = vde of L1/L2,
                                      aux0/aux1 = pbase of L1/L2           */
  AHK_V = 5,   /* V source (vde)      params: [dc, |ac|, arg(ac)] + the waveform
                                      block iff tf != AHK_TF_NONE: 8 slots,
                                      AHK_TF_PWL: 4 + 2*npts slots          */
  AHK_I = 6,   /* I source            params: as AHK_V                      */
  AHK_E = 7,   /* VCVS (vde)          params: [alpha]  n3,n4 = sense nodes  */
  AHK_G = 8,   /* VCCS                params: [alpha]  n3,n4 = sense nodes  */
  AHK_H = 9,   /* CCVS (vde)          params: [alpha]  ref = sensed vde     */
  AHK_F = 10,  /* CCCS                params: [alpha]  ref = sensed vde     */
  AHK_D = 11,  /* diode               params: [AREA]; model: IS N NBV ISR NR
                                      RS BV IKF VT  (ahkab/diode.py:352-424) */
  AHK_EKV = 12,/* EKV MOS  n1..n4 = d,g,s,b  params: [W L M N]; model: VTO KP
                  GAMMA PHI COX LAMBDA XJ UCRIT; flags = NPMOS (+1/-1)
                  (ahkab/ekv.py:444-752)                                     */
  AHK_MOSQ = 13,/* square-law MOS n1..n4 = d,g,s,b params: [W L M N ZVT ZKP];
                  model: VTO KP GAMMA PHI LAMBDA AVT AKP; flags = NPMOS
                  (ahkab/mosq.py:552-803)                                    */
  AHK_SW = 14  /* V-controlled switch n3,n4 = control nodes params: [is_on0];
                  model: VT VH RON ROFF (ahkab/switch.py:310-399)           */
};

/* time functions of V/I sources (ahkab/time_functions.py:402-840) */
enum {
  AHK_TF_NONE = 0,
  AHK_TF_PULSE = 1, /* v1 v2 td tr pw tf per            */
  AHK_TF_SIN = 2,   /* vo va freq td theta phi          */
  AHK_TF_EXP = 3,   /* v1 v2 td1 tau1 td2 tau2          */
  AHK_TF_SFFM = 4,  /* vo va fc mdi fs td               */
  AHK_TF_AM = 5,    /* sa fc fm oc td                   */
  AHK_TF_PWL = 6    /* npts td repeat repeat_time x0 y0 x1 y1 ... (4 + 2*npts slots,
                       linear interpolation / extrapolation, time_functions.py:730-828) */
};

#define AHK_FLAG_HAS_AC 0x100   /* V/I: |ac|, arg(ac) are meaningful       */
#define AHK_FLAG_HAS_IC 0x200   /* C/L: ic slot is meaningful              */

typedef struct {
  int32_t type;      /* AHK_*                                              */
  int32_t n1, n2, n3, n4; /* node numbers, reference numbering, 0 = ground */
  int32_t vde;       /* index among voltage-defined elements, else -1     */
  int32_t ref, ref2; /* see the table above, else -1                      */
  int32_t aux0, aux1;
  int32_t pbase;     /* first slot of the element's own parameter block   */
  int32_t mbase;     /* first slot of its model block, else -1            */
  int32_t flags;     /* NPMOS for MOS, AHK_FLAG_* bits                    */
  int32_t tf;        /* AHK_TF_* for V/I                                  */
} ahk_elem;

/* One circuit topology + nominal parameter values.
 * Replaces the `Circuit` argument of op_analysis/dc_analysis/
 * transient_analysis/ac_analysis (ahkab/dc_analysis.py:598,461;
 * ahkab/transient.py:121; ahkab/ac.py:146). */
typedef struct {
  int32_t abi_version;   /* AHK_ABI_VERSION                                */
  int32_t n_nodes;       /* nodes incl. ground = Circuit.get_nodes_number() */
  int32_t n_vde;         /* number of voltage-defined elements             */
  int32_t n_elems;
  const ahk_elem* elems; /* in circuit order                               */
  int32_t n_params;
  const double* nominal; /* [n_params]                                     */
  /* OP initial guess operator x0 = G . T  (ahkab/dc_guess.py:40-227):
   * G is topology-only (pinv of the port-incidence matrix, computed once on
   * the host exactly like the reference), T[i] = guess_scale[i] *
   * param[guess_slot[i]] + guess_const[i]  (guess_slot < 0: constant only). */
  int32_t n_guess;       /* 0 = no guess available                         */
  const double* guess_G; /* [n][n_guess] row-major, n = n_nodes-1+n_vde     */
  const int32_t* guess_slot;
  const double* guess_scale;
  const double* guess_const;
} ahk_circuit;

/* ahkab/options.py:50-132,163 snapshot (module globals in the reference). */
typedef struct {
  double vea, ver, iea, ier;
  double gmin, cmin, hmin;
  double nl_voltages_lock_factor;
  double transient_aposteriori_step_threshold;
  int32_t nl_voltages_lock;
  int32_t nr_damp_first_iters;
  int32_t dc_max_nr_iter;
  int32_t transient_max_nr_iter;
  int32_t ac_max_nr_iter;
  int32_t use_standard_solve_method;
  int32_t use_gmin_stepping;
  int32_t use_source_stepping;
  int32_t transient_prediction_as_x0;
  int32_t transient_use_aposteriori_step_control;
  int32_t transient_max_time_iter;
  int32_t reserved;
} ahk_options;

/* The batch axis: instance b uses nominal[] with slot vary_slots[k] replaced
 * by vary[k*B + b]  (slot-major so that consecutive instances coalesce). */
typedef struct {
  int64_t B;
  int32_t n_vary;
  const int32_t* vary_slots; /* host [n_vary]                              */
  const double* vary;        /* device [n_vary][B]                         */
} ahk_batch;

/* ---- per-instance status word ------------------------------------------ */
#define AHK_ST_OK          0x01 /* analysis produced a solution             */
#define AHK_ST_METHOD_MASK 0x06 /* last dc_solve: 0 std, 1 gmin-step, 2 src */
#define AHK_ST_METHOD_SHIFT 1
#define AHK_ST_GMIN_CHECK  0x08 /* op_solution.gmin_check found bad vars    */
#define AHK_ST_PASS2_FAIL  0x10 /* no-Gmin re-solve failed; x is pass 1's   */
#define AHK_ST_HMIN        0x20 /* transient step fell below hmin           */
#define AHK_ST_ROWS_FULL   0x40 /* transient output buffer exhausted        */
#define AHK_ST_SINGULAR    0x80 /* a singular Jacobian was met              */
#define AHK_ST_NR_FAIL    0x100 /* fixed-step transient: NR did not converge */

/* transient methods (ahkab/transient.py:58-66) */
enum {
  AHK_IMPLICIT_EULER = 0, AHK_TRAP = 1,
  AHK_GEAR1 = 11, AHK_GEAR2 = 12, AHK_GEAR3 = 13,
  AHK_GEAR4 = 14, AHK_GEAR5 = 15, AHK_GEAR6 = 16
};

typedef struct ahk_engine ahk_engine;

#ifndef __CUDACC_RTC__   /* host entry points: not part of a device-only (NVRTC) build */

/* Build the device-side stamp program for one topology. Host call. */
int ahk_create(const ahk_circuit* circ, const ahk_options* opts,
               ahk_engine** out);
void ahk_destroy(ahk_engine* e);
/* reduced MNA size n = n_nodes - 1 + n_vde */
int ahk_size(const ahk_engine* e);
/* Replace the options snapshot (cheap). */
int ahk_set_options(ahk_engine* e, const ahk_options* opts);
/* Tuning: lanes cooperating on one instance (1,2,4,8,16,32; 0 = heuristic):
 * restricts the kernel-variant choice to variants with that many lanes. */
int ahk_set_group(ahk_engine* e, int lanes);
/* Kernel variants.  A variant (NC, R, G) fixes at compile time the matrix columns
 * NC (>= n+1; 0 = run-time n, shared-memory fallback), the rows per lane R and the
 * lanes per instance G.  kind 0 = OP/DC/transient kernels, 1 = AC kernels.
 * ahk_variant_info writes {NC, R, G}; ahk_set_variant forces one (-1 = heuristic),
 * failing if it cannot hold the engine's circuit. */
int ahk_variant_count(int kind);
int ahk_variant_info(int kind, int id, int32_t* nc_r_g);
int ahk_set_variant(ahk_engine* e, int real_id, int ac_id);
/* Circuit-specialised kernels (no reference counterpart).  The stock OP/DC/transient
 * kernels interpret the stamp program; for large batches the engine can instead compile,
 * once per (circuit, options, analysis), the same per-instance source with the program,
 * options and memory layout folded in as constants (NVRTC, sm_100a) — same results,
 * several times fewer instructions.  mode 0 = never, 1 = always (an NVRTC failure is an
 * error), 2 = for batches of at least min_batch instances, default (2, 4096);
 * min_batch <= 0 keeps the current threshold.  The first call with a new combination pays
 * the compilation (seconds); the kernel is cached in the engine and, as a cubin, on disk
 * ($AHK_JIT_CACHE, default ~/.cache/ahkab_b200; "off" disables), so that other processes and
 * later runs with the same circuit skip NVRTC. */
int ahk_set_jit(ahk_engine* e, int mode, int64_t min_batch);
/* Tuning: shape of the specialised OP/DC/transient kernels — rows of the matrix per lane and
 * lanes per instance (1,2,4,8,16,32; rows_per_lane * lanes >= n).  Being compiled on demand
 * they are not bound to the compiled variant table.  (0, 0) = the heuristic's variant. */
int ahk_set_jit_shape(ahk_engine* e, int rows_per_lane, int lanes);
/* Tuning of the specialised OP/DC/transient kernels: ilp = devices one lane evaluates
 * interleaved (1..4, 0 = heuristic); threads = block size (multiple of 32, 0 = heuristic);
 * min_blocks = resident blocks per SM the register allocation is capped for (0 = heuristic). */
int ahk_set_jit_tuning(ahk_engine* e, int ilp, int threads, int min_blocks);
/* Static-schedule LU of the specialised kernels whose lanes hold every row (one lane per
 * instance or replicated rows): numpy.linalg.solve's LU (dc_analysis.py:822) eliminated
 * symbolically along the pivot sequence a calibration launch observed, emitted as straight-line
 * code on the structural non-zeros; every pivot is verified against dgesv's rule at run time and
 * a solve that leaves the sequence goes through the pivoting LU (same results either way).
 * mode: -1 default (environment AHK_SLU, else transients), 0 off, 1 transients, 2 OP / DC too. */
int ahk_set_jit_slu(ahk_engine* e, int mode);
/* out4 = { calibration launches, schedules replaced by the observed sequence, solves of
 * the calibration sample (first 32 instances; AC: the chunks of instance 0) that left the schedule
 * in the last calibration, solves of the sample in it }. */
int ahk_jit_slu_stats(const ahk_engine* e, int32_t* out4);
/* Shape of the last small-circuit kernel launch of this engine (diagnostics, tests):
 * out8 = { variant id (>= 1000: free shape), matrix columns, rows per lane, lanes per
 * instance, threads per block, blocks, 1 if circuit-specialised, analysis 0 OP/DC 1 tran 2 AC }. */
int ahk_last_launch(const ahk_engine* e, int32_t* out8);
/* Large linear circuits (n > 64): which kernels served the last call and what the static
 * elimination schedule executes (diagnostics; bench.py's roofline): out8 = { path (0 dense
 * workspace, 1 sparse dynamic, 2 static schedule, 3 static schedule compiled by NVRTC; -1 none
 * yet), factorisation flop per instance (both matrices), flop per (instance, sweep point),
 * factor values stored per matrix, structural matrix entries, value slots incl. fill-in,
 * unknowns, 0 }. */
int ahk_big_stats(const ahk_engine* e, int64_t* out8);
/* Launches through specialised kernels / NVRTC compilations so far (all engines). */
int64_t ahk_jit_launch_count(void);
int64_t ahk_jit_build_count(void);
/* Host-only check (needs NVRTC, no GPU): compiles the specialised OP/DC (kind = 0),
 * transient (kind = 1) or AC (kind = 2; variant ids of the AC table) kernel of this circuit
 * for kernel variant variant_id and returns the size of the sm_100a cubin. */
int ahk_jit_compile(const ahk_circuit* circ, const ahk_options* opts, int n_vary,
                    const int32_t* vary_slots, int kind, int method,
                    int use_step_control, int variant_id, int64_t* cubin_bytes);
const char* ahk_last_error(void);
/* Number of kernel launches issued by this library so far (all engines). */
int64_t ahk_launch_count(void);

/* Plumbing for chunked (pipelined) end-to-end runs: copies columns [lo, hi) of the
 * slot-major host array src_host[n_rows][row_len] (pinned memory for the copy to be
 * asynchronous) into the contiguous device array dst[n_rows][hi - lo] — the `vary`
 * layout of a sub-batch — with one strided asynchronous copy on `stream`. */
int ahk_h2d_columns(double* dst, const double* src_host, int n_rows,
                    int64_t row_len, int64_t lo, int64_t hi, void* stream);

/* Plumbing for end-to-end transients: ahk_tran writes [B][max_rows] padded rows, of which
 * instance b used rows[b].  This packs the used rows of all instances contiguously,
 * instance after instance (offsets[b] = rows[0] + ... + rows[b-1], device, int64), so that
 * the device->host read moves sum(rows) rows instead of B * max_rows.
 *   t_packed device [sum rows], x_packed device [sum rows][n] */
int ahk_pack_rows(const double* t_out, const double* x_out, const int32_t* rows,
                  const int64_t* offsets, int64_t B, int64_t max_rows, int n,
                  double* t_packed, double* x_packed, void* stream);

/* Diagnostics (no reference counterpart): launches an FP64-FMA-only kernel,
 * blocks x threads (<= 256) threads each running 8 independent chains of
 * `iters` DFMAs = blocks*threads*iters*16 flop.  bench.py times it with CUDA
 * events to obtain the measured FP64 roofline denominator.  `out` is a device
 * pointer to at least one double.  Not counted by ahk_launch_count(). */
int ahk_probe_fp64(int blocks, int threads, int iters, double* out,
                   void* stream);

/* Operating point — replaces dc_analysis.op_analysis
 * (ahkab/dc_analysis.py:598-687): guess -> dc_solve with Gmin -> dc_solve
 * without Gmin from x1 -> gmin_check; iterations = n1 + n2.
 *   x0     device [B][n] or NULL (NULL + use_guess -> dc_guess, else zeros)
 *   x      device [B][n]     solution
 *   resid  device [B][n] or NULL  residual of the last NR iteration ("error")
 *   iters  device [B]   total NR iterations
 *   status device [B]   AHK_ST_* */
int ahk_op(ahk_engine* e, const ahk_batch* batch, const double* x0,
           int use_guess, double* x, double* resid, int32_t* iters,
           int32_t* status, void* stream);

/* Circuit size, all four analyses: n = nodes - 1 + voltage-defined elements up to 512 (the
 * reference solves every size densely, ahkab/dc_analysis.py:796-822).  n <= 64: matrix in
 * registers / shared memory.  65 <= n <= 512: the general large-n engine — one warp per
 * instance, the dense matrix in a global workspace the library allocates (B x n x (n + 8)
 * doubles, grow-only), non-zero bitmaps, dense panels on the FP64 tensor cores; interpreted
 * kernels only (ahk_set_jit mode 1 is an error there).  Sources with a waveform enter OP / DC
 * with their dc slot (VSource.V(time=None), ahkab/components/sources/VSource.py:92-97). */

/* Operating point with HOST buffers (what a caller holding numpy arrays binds): the same
 * analysis as ahk_op, but parameters and results live in host memory (pinned memory for
 * the copies to be asynchronous).  The batch is cut into `nchunks` chunks that flow
 * through three internal streams — copy-in, kernels, copy-out — so that the PCIe traffic
 * of both directions overlaps the kernels.  Asynchronous: `stream` is made to wait for
 * the last copy; synchronise it before reading the results.
 *   vary_slots host [n_vary], vary_host host [n_vary][B] (slot-major, as ahk_batch.vary)
 *   x_host [B][n], resid_host [B][n] or NULL, iters_host [B], status_host [B] */
int ahk_op_host(ahk_engine* e, int64_t B, int n_vary, const int32_t* vary_slots,
                const double* vary_host, int use_guess, double* x_host,
                double* resid_host, int32_t* iters_host, int32_t* status_host,
                int nchunks, void* stream);

/* DC sweep — replaces dc_analysis.dc_analysis (ahkab/dc_analysis.py:461-595):
 * for every sweep value (host array, the caller builds it with the
 * reference's cumulative lin/log iterators, ahkab/utilities.py:246-378) the
 * full OP procedure, warm-started from the previous point.  (Linear circuits with
 * n > 64 take the sparse large-circuit path: both matrices are factored once per
 * instance, the points are solved independently — each is one exact Newton step —
 * and x0 / use_guess do not influence the result beyond rounding.)
 *   src_elem  index in circ->elems of the swept V or I source (its dc value is swept; a
 *             waveform on it is ignored, as in the reference)
 *   x_out   device [B][npts][n]; iters/status device [B][npts] */
int ahk_dc_sweep(ahk_engine* e, const ahk_batch* batch, int src_elem,
                 const double* sweep_values, int npts, const double* x0,
                 int use_guess, double* x_out, int32_t* iters,
                 int32_t* status, void* stream);

/* Transient — replaces transient.transient_analysis
 * (ahkab/transient.py:121-470) incl. the DF modules implicit_euler/trap/gear.
 *   x0       device [B][n] or NULL (zeros)
 *   tstep    = HMAX with step control, the fixed step without
 *   t_out    device [B][max_rows]; x_out device [B][max_rows][n]
 *            (no row for tstart, transient.py:319,388)
 *   rows     device [B] accepted steps written
 *   counters device [B][4]: dc_solve calls, NR iterations, LTE-rejected
 *            steps, NR-failure step cuts */
int ahk_tran(ahk_engine* e, const ahk_batch* batch, const double* x0,
             double tstart, double tstep, double tstop, int method,
             int use_step_control, int64_t max_rows, double* t_out,
             double* x_out, int32_t* rows, int32_t* counters,
             int32_t* status, void* stream);

/* AC — replaces ac.ac_analysis (ahkab/ac.py:146-320): per omega one
 * complex128 solve of (MNA + j*omega*AC + J(x_op) + Gmin) x = -Nac.
 *   x_op    device [B][n] linearisation point (required iff the circuit is
 *           non-linear), else NULL
 *   omegas  host [npts], rad/s (caller uses the reference's iterators)
 *   x_out   device [B][npts][n] complex128 (re,im interleaved) */
int ahk_ac(ahk_engine* e, const ahk_batch* batch, const double* x_op,
           const double* omegas, int npts, double* x_out, int32_t* status,
           void* stream);
#endif /* !__CUDACC_RTC__ */

#ifdef __cplusplus
}
#endif
#endif /* AHKAB_B200_H */
