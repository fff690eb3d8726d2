/* petite_b200 C ABI - the drop-in boundary of the B200 shower engine.
 *
 * PETITE (the reference) has no FFI of its own: its boundary is the Python class API
 *   Shower.generate_shower          reference src/PETITE/shower.py:603-708
 *   DarkShower.generate_dark_shower reference src/PETITE/dark_shower.py:806-849
 * and the methods they call.  The entry points below are what a ctypes binding inside those two
 * classes would call (INTEGRATION.md shows the stub); each comment names the reference lines the
 * call replaces.  Plain C types only; no torch types cross this boundary.
 *
 * Conventions: every function returns 0 on success and a negative pb_status otherwise;
 * pb_last_error() returns a static, per-engine message.  An engine is bound to one CUDA device and is
 * not thread-safe; distinct engines are independent.  All kernels are enqueued on the stream passed
 * in (a cudaStream_t cast to void*, NULL = default stream).  Pointers marked [host] are host memory,
 * [dev] device memory owned by the caller (e.g. torch.Tensor.data_ptr()); the library never frees
 * caller memory.
 */
#ifndef PETITE_B200_H
#define PETITE_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct pb_engine_s* pb_engine;

enum pb_status {
  PB_OK = 0,
  PB_ERR_CUDA = -1,
  PB_ERR_ARG = -2,
  PB_ERR_CAPACITY = -3,   /* particle stack too small for the showers requested */
  PB_ERR_STATE = -4,      /* tables missing */
  PB_ERR_NO_SAMPLE = -5   /* a sampler exhausted max_n_integrators sweeps (shower.py:460-461) */
};

/* process codes (generation_process); 0-7 follow shower.py:36 process_code */
enum pb_process {
  PB_BREM = 0, PB_ANN = 1, PB_PAIRPROD = 2, PB_COMP = 3, PB_MOLLER = 4, PB_BHABHA = 5, PB_MUONE = 6,
  PB_MUONBREM = 7, PB_DARKBREM = 8, PB_DARKANN = 9, PB_DARKCOMP = 10, PB_DARKMUONBREM = 11,
  PB_SMDECAY = 12, PB_BSMDECAY = 13, PB_NONE = 14, PB_INPUT = 15, PB_NPROC_SAMPLED = 12
};

/* Shower / DarkShower constructor state that the kernels need (shower.py:100-142, 189-208;
 * dark_shower.py:69-129).  All energies GeV, lengths m unless noted. */
typedef struct pb_config {
  double Z_T, A_T, rho;          /* physical_constants.py:49-55 */
  double dEdx_GeV_per_m;         /* 0.1 * dEdx[MeV/cm]  (shower.py:649) */
  double mT_sampler;             /* event_info['mT'] at sampling time = A_T (SURVEY Q-19) */
  double min_energy;             /* Shower(min_energy) */
  double Eg_min, Ee_min;         /* shower.py:214-215 */
  double maxF_fudge;             /* maxF_fudge_global */
  double rescale_MCS;            /* rescale_MCS */
  double min_calc[5];            /* shower.py:242-246 for e-, e+, gamma, mu-, mu+ (in that order) */
  int64_t max_sweeps;            /* max_n_integrators */
  /* dark sector (ignored by pb_run_showers) */
  double mV, g_e, kinetic_mixing, Zeff;
  double E_res_ann, E_thr_comp;  /* dark_shower.py:234-235 */
  int32_t bound_electron;
  int32_t reserved;
} pb_config;

/* Structure-of-arrays particle stack in HBM (the Particle list of shower.py:621, particle.py:50-175).
 * One record = 4 x double4 + 32 bytes of ids + 8 bytes of counters = 168 B.  Records [0, n_primaries) are the primaries, daughters
 * are appended wave by wave; record order within a wave is not the reference's creation order (host
 * side restores it from parent/child links, see petite_b200/shower.py). */
typedef struct pb_stack {
  double* p0;      /* [dev] capacity x 4 : E, px, py, pz at creation            (particle._p0) */
  double* r0w;     /* [dev] capacity x 4 : x, y, z at creation, weight          (particle._r0, ids['weight']) */
  double* pf;      /* [dev] capacity x 4 : four-momentum after propagation      (particle._pf) */
  double* rf;      /* [dev] capacity x 4 : position after propagation, 4th = 0  (particle._rf) */
  int32_t* ids;    /* [dev] capacity x 8 : pid, parent slot (-1 primary), gen<<16 | child_bit<<15 | flags<<8 | process, shower id,
                                           Philox particle key (2 words), weight again (the two halves of the double): one
                                           32-byte sector per record carries everything the wave kernels need besides the
                                           four-vectors */
  int32_t* aux;    /* [dev] capacity x 2 : accept/reject trials used, dE/dx sub-steps taken */
  int64_t capacity;
} pb_stack;

#define PB_FLAG_SHORT_LIVED 1   /* ids['stability'] == 'short-lived' */
#define PB_FLAG_LONG_LIVED 4    /* ids['stability'] == 'long-lived': pi+-, K+- decay in flight (particle.py:410-422) */
#define PB_FLAG_NO_SAMPLE 2     /* sampler exhausted */

/* Primaries (host SoA): what the caller's loop over Particle objects feeds generate_shower with. */
typedef struct pb_primaries {
  const double* p;        /* [host] n x 4 */
  const double* r;        /* [host] n x 3 */
  const double* weight;   /* [host] n */
  const double* mass;     /* [host] n   ids['mass'] (particle.py:125-131 back-computed value if absent) */
  const int32_t* pid;     /* [host] n */
  const int32_t* flags;   /* [host] n   PB_FLAG_SHORT_LIVED for pi0 etc., PB_FLAG_LONG_LIVED for pi+-, K+- */
  int64_t n;
  int64_t on_device;      /* 0: the six arrays are host memory (copied inside the call); 1: they are [dev] already */
} pb_primaries;

typedef struct pb_counters {
  int64_t n_particles;     /* records written (primaries + daughters) */
  int64_t n_waves;
  int64_t n_steps;         /* propagate_particle calls that actually stepped */
  int64_t n_substeps;      /* dE/dx + MCS sub-steps */
  int64_t n_samples;       /* accepted draw_sample calls */
  int64_t n_trials;        /* integrand evaluations */
  int64_t n_no_sample;     /* samplers that gave up */
  int64_t n_launches;      /* kernels launched by the engine for this call */
  int64_t max_wave;        /* widest wave */
  int64_t n_charged;       /* records that went through the dE/dx sub-step loop (e+-, mu+-) */
} pb_counters;

/* Per-kernel device times (CUDA events on the launching stream) and per-process sampler counts of the LAST
 * pb_run_showers / pb_run_dark call; filled only while profiling is on (it adds two events per launch). */
enum pb_kernel { PB_K_INIT = 0, PB_K_PROPAGATE = 1 /* k_loop */, PB_K_SCAN = 2, PB_K_FILL = 3, PB_K_SAMPLE = 4, PB_K_EMIT = 5,
                 PB_K_FINALIZE = 6, PB_K_N = 7 };
typedef struct pb_profile {
  double ms[8];            /* summed device time per pb_kernel */
  int64_t launches[8];
  int64_t trials[16];      /* integrand evaluations per pb_process */
  int64_t samples[16];     /* accepted samples per pb_process */
} pb_profile;

/* Tally layout (doubles) written by pb_tally: see PB_TALLY_* ; counts are exact integers stored as fp64 so that one
 * NCCL all-reduce(sum) over the buffer combines ranks. Species order: e-, e+, gamma, mu-, mu+, V(4900022), other. */
#define PB_TALLY_NSPECIES 7
#define PB_TALLY_EBINS 64        /* log10(E/GeV) in [-3, 3) */
#define PB_TALLY_TBINS 32        /* log10(theta) in [-7, 1) */
#define PB_TALLY_COUNT 0                                   /* [7] records per species (creation)            */
#define PB_TALLY_WSUM 8                                    /* [7] sum of weights per species                */
#define PB_TALLY_WESUM 16                                  /* [7] sum of weight * E0 per species            */
#define PB_TALLY_EHIST 32                                  /* [7][64] weighted spectrum of creation energy  */
#define PB_TALLY_THIST (32 + 7 * 64)                       /* [7][32] weighted polar-angle spectrum         */
#define PB_TALLY_SIZE 1024

const char* pb_version(void);
const char* pb_last_error(pb_engine e);

/* Shower.__init__ (shower.py:100-142) minus the table reads, which stay in Python. */
int pb_create(pb_engine* out, int device, const pb_config* cfg);
void pb_destroy(pb_engine e);
int pb_set_config(pb_engine e, const pb_config* cfg);

/* One n*sigma(E) interpolant of set_NSigmas (shower.py:273-295) / set_dark_NSigmas; table_id = pb_process
 * for the 8 SM processes.  Linear in (E, n*sigma) with 0 outside, evaluated like scipy interp1d. */
int pb_upload_nsigma(pb_engine e, int table_id, const double* E /*[host]*/, const double* nsigma /*[host]*/, int n);

/* All trained maps of one process (set_samples, shower.py:210-215; set_dark_samples, dark_shower.py:196-201):
 * grid = n_energy rows of sum_d(ninc[d]+1) fp64 node positions (axes concatenated), E_inc and max_F per row. */
int pb_upload_maps(pb_engine e, int process, const double* grid /*[host]*/, int n_energy, int dim,
                   const int32_t* ninc /*[host]*/, const double* E_inc /*[host]*/,
                   const double* max_F /*[host]*/, int neval);

/* generate_shower (shower.py:603-708) for a batch of independent primaries.  Shower i of this call uses the
 * Philox root key (seed, first_shower_id + i).  global_ms mirrors the GlobalMS argument. */
int pb_run_showers(pb_engine e, const pb_primaries* prim, uint64_t seed, uint64_t first_shower_id,
                   int global_ms, pb_stack* stack, pb_counters* counters /*[host] out*/, void* stream);

/* DarkShower constructor tables needed on the device (dark_shower.py:219-593), all [host]:
 *   weights   _brem_elec_numerical_weight, _brem_positron_numerical_weight, _annihilation_numerical_weight,
 *             _muon_brem_numerical_weight (dark_shower.py:447-452): linear, 0 outside
 *   nsdark    _NSigmaDarkComp (dark_shower.py:306): log10 nodes, log-log interpolation, 1e-20 outside
 *   drate     _d_rate_dict_{elec_brem, positron_brem, positron_ann, muon_brem} (dark_shower.py:590-593):
 *             E[n] saved energies (ascending) and table[n][10][2] = (bin-centre energy, rate)
 *   min_E     _minimum_calculable_dark_energy for DarkBrem, DarkAnn, DarkComp, DarkMuonBrem (dark_shower.py:236-245) */
typedef struct pb_dark_tables {
  const double* w_E[4];
  const double* w_y[4];
  int32_t w_n[4];
  const double* nsdark_comp_lx;
  const double* nsdark_comp_ly;
  int32_t nsdark_comp_n;
  int32_t pad;
  const double* d_E[4];
  const double* d_table[4];
  int32_t d_n[4];
  double min_E[4];
} pb_dark_tables;
int pb_upload_dark(pb_engine e, const pb_dark_tables* t);

/* generate_dark_shower (dark_shower.py:806-849) over the first n_sm records of an SM stack (the output of
 * pb_run_showers): for every record and every active process with a positive weight, one dark-vector record is
 * appended to `dark` (pid 4900022, parent = SM slot, weight = wg * parent weight, r0 = parent's final position).
 * active_mask: bit (1 << pb_process) for PB_DARKBREM, PB_DARKANN, PB_DARKCOMP, PB_DARKMUONBREM, PB_BSMDECAY. */
int pb_run_dark(pb_engine e, const pb_stack* sm, int64_t n_sm, uint32_t active_mask, pb_stack* dark,
                pb_counters* counters /*[host] out*/, void* stream);

/* draw_sample / draw_dark_sample (shower.py:401-465, dark_shower.py:649-704) for n independent incoming energies:
 * sample i uses the Philox key (seed, first_id + i).  lu_key < 0 selects the map row as the reference does (Q-1).
 * x_out [host] n x 4 (map variables, unused slots 0), ntrials_out [host] n (the VB counter; -1 = "No Sample Found"). */
int pb_draw_samples(pb_engine e, int process, const double* E /*[host]*/, int64_t n, int lu_key, uint64_t seed,
                    uint64_t first_id, double* x_out, int32_t* ntrials_out, void* stream);

/* do_find_max_work (utilities/find_maxes.py:55-119) for every trained map row of one process with the engine's target
 * (Z_T, A_T): n_trials sweeps of `neval` points, max_F = max over sweeps of max(wgt*f) (a sweep containing a NaN never
 * raises it), sigma = sum(wgt*f)/n_trials.  mT > 0 overrides event_info['mT'] (the table builder uses the true
 * nuclear mass, the sampler A_T: SURVEY Q-19).  max_F_out, sigma_out: [host] n_energy each. */
int pb_find_max(pb_engine e, int process, int n_trials, uint64_t seed, double mT, double* max_F_out, double* sigma_out);

/* One VEGAS training sweep for n_energy maps of one process (utilities/generate_integrators.py:49-110 ->
 * all_processes.py:1160-1226 -> vegas' AdaptiveMap training data): n_points uniform y per map are pushed through the
 * CURRENT node grids `grid` ([host] n_energy rows, same layout as pb_upload_maps) and the integrand of `process` at
 * E_inc[k]; per axis and increment the sums of (jac*f)^2 and the hit counts are returned in d_out / n_out ([host]
 * n_energy x sum(ninc+1), slot i of an axis = increment i, last slot unused), the plain MC integral in integral_out.
 * The grid refinement itself (smoothing, alpha damping, rebinning) is done by the caller (petite_b200/train.py).
 * Z_T/A_T/mT/mV of the integrand come from the engine's configuration. */
int pb_train_accumulate(pb_engine e, int process, const double* grid, int n_energy, int dim, const int32_t* ninc,
                        const double* E_inc, int64_t n_points, uint64_t seed, double mT, double* d_out, double* n_out,
                        double* integral_out);
/* The same with the training weight |jac*f|^power (0 < power <= 64) in d_out.  power = 2 is pb_train_accumulate, i.e. VEGAS'
 * variance criterion; the maps are USED for accept/reject sampling (shower.py:401-465), whose cost is max / mean of jac*f, and
 * a larger power trains for exactly that: power = 8 gives maps with 2.3-2.8x the accept rate of the shipped ones for Brem,
 * PairProd and MuonBrem (profiles/r02_final/exp_train_pow2.log); the integral through the map is unbiased for any power. */
int pb_train_accumulate_p(pb_engine e, int process, const double* grid, int n_energy, int dim, const int32_t* ninc,
                          const double* E_inc, int64_t n_points, uint64_t seed, double mT, double power, double* d_out,
                          double* n_out, double* integral_out);

/* Histogram / yield tallies over records [first, first+n) of a stack, ACCUMULATED into tally[PB_TALLY_SIZE] [dev]. */
int pb_tally(pb_engine e, const pb_stack* stack, int64_t first, int64_t n, double* tally /*[dev]*/, void* stream);

/* detector_cut / transverse_position (shower.py:815-864): straight-line extrapolation of records [first, first+n) from
 * their creation point to the planes z = z_det[k] and test r_T in (inner_radius, radius); records are first filtered by
 * E_lo < E < E_hi (pass -inf/+inf for no cut).  weight_pass [host] n_det = summed weights inside the ring ("TotalWeight"),
 * weight_all [host] 1 = summed weight after the energy cut ("Efficiency" = ratio); mask [dev] n x n_det bytes or NULL
 * ("SampleW"). */
int pb_detector_cut(pb_engine e, const pb_stack* stack, int64_t first, int64_t n, const double* z_det /*[host]*/, int n_det,
                    double radius, double inner_radius, double E_lo, double E_hi, double* weight_pass, double* weight_all,
                    uint8_t* mask /*[dev]*/, void* stream);

int pb_set_profiling(pb_engine e, int level);   /* 0 off; 1 = time k_loop and k_sample only; 2 = time every kernel */
int pb_get_profile(pb_engine e, pb_profile* out /*[host]*/);

/* FP64 FMA-chain microbenchmark on the engine's device: the measured FP64 roofline denominator (TFLOP/s). */
int pb_measure_fp64_peak(pb_engine e, double* tflops /*[host] out*/);

/* Deterministic-piece probes used by the parity tests (all arrays [host]). what: */
enum pb_probe {
  PB_PROBE_DSIGMA = 0,   /* in: n x (1 + dim): E_inc, x[dim];            out: n     dsigma (integrands of all_processes.py) */
  PB_PROBE_NSIGMA = 1,   /* in: n x 1: E;  process = table id;           out: n     n*sigma(E) */
  PB_PROBE_MAP = 2,      /* in: n x (1 + dim): energy row index, y[dim]; out: n x (dim+1): x[dim], jac */
  PB_PROBE_MCS = 3,      /* in: n x 9: p4[4], dist_m, m_lepton, u_sign, z1, z2 ... see tests; out: n x 4 */
  PB_PROBE_KIN = 4,      /* in: n x 8: E, mass, x[4], u_az, pad;          out: n x 8 two four-vectors (parent along z) */
  PB_PROBE_PHILOX = 5,   /* in: n x 6 (as doubles): key0,key1,c0,stream,c2,c3; out: n x 2 doubles */
  PB_PROBE_HOTMATH = 6,  /* in: n x 4: x_log (> 0), x_exp in [1/20, 1/6], theta, u in [0,1);
                            out: n x 7: hot_log, hot_exp_neg_step, sin, cos (theta), sin, cos (2 pi u), fast_rcp(x_log)
                            (out stride >= 10 adds fast_sqrt0(x_log), fast_rsqrt(x_log), fast_sqrt0 at 0 and below;
                            out stride >= 12 adds hot_cospi(2 u - 1), hot_cospi(2 u): the azimuth factor of the 4-D sampler integrands) */
  PB_PROBE_MCS_FAST = 7, /* the folded multiple-scattering form of the sub-step loop; in / out as PB_PROBE_MCS (particle mass = m_lepton) */
  PB_PROBE_SUBSTEP = 8,  /* ONE iteration of the dE/dx + MCS loop exactly as k_loop runs it (shower.py:559-581).
                            in: n x 12: pid, E, px, py, pz, x, y, z, key0, key1, sub-step index, multiple scattering on/off;
                            out: n x 10: loop ended (1) / sub-step applied (0), E, px, py, pz, x, y, z, delta_z, next index */
  PB_PROBE_DARKKIN = 9,  /* dark-vector four-momentum in the parent frame (kinematics.py:43-68, 134-183, 267-299); process =
                            DarkBrem / DarkMuonBrem / DarkAnn / DarkComp.  in: n x 10: E, mV, x[4], u1, u2, Pe, cos(theta_e);
                            out: n x 4 */
  PB_PROBE_PROPAGATE = 10 /* propagate_particle (shower.py:509-601) of single particles with the engine's draw protocol.
                            in: n x 12: pid, E, px, py, pz, x, y, z, mass, key0, key1, multiple scattering on/off;
                            out: n x 9: pf[4], rf[3], sub-steps, 1 if propagated (0: below threshold, untouched) */
};
int pb_probe(pb_engine e, int what, int process, const double* in, int64_t n, int in_stride, double* out, int out_stride);

/* Replay (parity mode): n independent particle-steps - propagate_particle (shower.py:509-601), the process choice (:665-698),
 * draw_sample's accept/reject loop (:451-459), kinematics and rotation (:467-507) - run by the wave kernels' own device
 * functions, with every random number taken from a TAPE recorded from a reference run instead of the Philox protocol.
 * The tape of step i is tape[tape_off[i] .. tape_off[i+1]) and holds, in the reference's consumption order:
 *   charged: per loop iteration  random(), uniform(6, 20) [, sign, z1, z2, uniform(0, 2 pi) / 2 pi  if a sub-step was applied
 *            with multiple scattering on];  then the final step  random() [, sign, z1, z2, u_phi];
 *   photon:  random() (free path);
 *   then the process choice uniform, then per tested trial  y_0 .. y_{dim-1}, u_accept,  then the kinematics azimuth uniform
 *   (two uniforms for a short-lived decay instead of all of the above).
 * particles [host]: n x 10 doubles: pid, E, px, py, pz, x, y, z, mass, flags (1 = multiple scattering on, 2 = short-lived, 4 = long-lived).
 * out [host]: n x 32 doubles: 0 status (0 ok, 1 tape ran out, 2 tape not used up, 3 no sample), 1 sub-steps, 2 process code,
 *   3 trials, 4-7 pf, 8-10 rf, 11 kept daughters (bit 0 / bit 1), 12 pid_a, 13-16 p_a, 17 pid_b, 18-21 p_b, 22-25 sampled x,
 *   26 tape entries consumed, 27 weight factor, 28 propagated (1) / below threshold (0). */
int pb_replay(pb_engine e, int64_t n, const double* particles, const double* tape, const int64_t* tape_off, double* out);

/* Dark set-up quadratures: the integrals DarkShower.__init__ computes with scipy.integrate.quad - the cumulative
 * interaction integrals II(E) (src/PETITE/shower.py:298-320), the emission weights (dark_shower.py:337-399) and the dRate/dE bins
 * (dark_shower.py:454-493) - one GPU thread per integral, through a restatement of QUADPACK's QAGS (epsabs = epsrel = 1.49e-8,
 * limit = 50, the scipy defaults; petite_b200/csrc/quadpack.cuh) so that the subdivision and extrapolation decisions, and with them
 * the reference's numbers, are reproduced.  Tables are linear interpolants with a fill value outside their range (scipy interp1d).
 *   kind 0: f(E) = table[tab](E)                                                                  (shower.py:305-320)
 *   kind 1: f(E) = 10^table[tab](log10 E) / dEdx_cm * exp(-sum_k (table[surv_k](Ei) - table[surv_k](E)) / dEdx_m / cmtom), 0 if the
 *           sum is negative or E > Ei, and 0 if `cut` and the first factor is below 1e-18       (dark_shower.py:311-335)
 * ier [host, optional]: QUADPACK's ier * 1000 + number of sub-intervals used. */
typedef struct pb_quad_call { int32_t kind, tab, surv[3], cut; double Ei, a, b; } pb_quad_call;
int pb_quad_batch(pb_engine e, int n_tabs, const int32_t* tab_n, const double* const* tab_x, const double* const* tab_y,
                  const double* tab_fill, double dEdx_GeV_per_m, const pb_quad_call* calls, int64_t n, double* result,
                  double* abserr, int32_t* ier);

#ifdef __cplusplus
}
#endif
#endif
