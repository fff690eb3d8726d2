// petite_b200 shower engine: wavefront stepping of independent EM showers on a B200 (sm_100a).
//
// One "wave" = every particle created by the previous wave, across all showers of the batch.  Per wave:
//   k_propagate   one thread / particle: free path or dE/dx + multiple-scattering sub-steps, process choice,
//                 threshold + map look-up key -> bucket (process, energy row); bucket histogram
//   k_bucket_scan one CTA: exclusive scan of the histogram, tile table (<= TILE samples of one bucket per tile)
//   k_bucket_fill counting-sort scatter of the wave's particles into bucket order
//   k_sample      persistent CTAs pull tiles; the bucket's VEGAS node grid is staged in shared memory by a TMA
//                 bulk copy; lane groups run counter-indexed accept/reject trials (first accepted trial wins)
//   k_emit        one thread / particle in bucket order: kinematics, rotation to the lab, daughter records
//                 appended with one warp-aggregated atomic per warp
// The host loop only needs the new tail of the stack after each wave.
// Reference behaviour reproduced: src/PETITE/shower.py:401-708 (see include/petite_b200.h and DESIGN.md).
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>
#include <stdlib.h>
#include <math.h>
#include <string>
#include <vector>
#include <algorithm>

#include "../../include/petite_b200.h"
#include "physics.cuh"
#include "quadpack.cuh"

namespace pb {

constexpr int LU_MAX = 256;                 // energy rows per process (100 in data/, 150 in data_400GeV/)
constexpr int NBUCKET = 16 * LU_MAX;        // bucket = process * LU_MAX + row
#ifndef PB_TILE
#define PB_TILE 192
#endif
#ifndef PB_SAMPLE_THREADS
#define PB_SAMPLE_THREADS 128
#endif
#ifndef PB_SAMPLE_MINB
#define PB_SAMPLE_MINB 4
#endif
#ifndef PB_SAMPLE_MINB_SM
#define PB_SAMPLE_MINB_SM 5              // SM-pass kernel (no dark integrands): ~96 registers
#endif
#ifndef PB_SAMPLE_MINB_T
#define PB_SAMPLE_MINB_T 3               // kernels evaluating T > 1 trials per lane and round trade occupancy for ILP
#endif
#ifndef PB_SAMPLE_MINB_SM_T
#define PB_SAMPLE_MINB_SM_T 5            // T = 2: 100 registers; measured (4, 2): minb 3 / 4 / 5 -> 62.5 / 57.3 / 57.0 ms per step, (4, 1): 62.3
#endif
#ifndef PB_SAMPLE_MINB_DB
#define PB_SAMPLE_MINB_DB 4              // dark-brem family kernel (FAM = 3)
#endif
#ifndef PB_SAMPLE_G_DARK_DEFAULT
#define PB_SAMPLE_G_DARK_DEFAULT 8       // dark pass: ~70-300 trials per sample, wide groups waste little (measured 77 vs 85 ms, config 3)
#endif
#ifndef PB_SAMPLE_T_DB_DEFAULT
#define PB_SAMPLE_T_DB_DEFAULT 2
#endif
#ifndef PB_FIN_MINB
#define PB_FIN_MINB 8                    // 64 registers, 8 CTAs per SM: measured 17.8 -> 15.3 ms per step (10: 15.7, 12: 16.2, uncapped 94 registers: 17.8) although it spills 92 B
#endif
#ifdef PB_FIN_MINB
#define PB_FIN_BOUNDS __launch_bounds__(128, PB_FIN_MINB)
#else
#define PB_FIN_BOUNDS __launch_bounds__(128)
#endif
#ifdef PB_EMIT_MINB
#define PB_EMIT_BOUNDS __launch_bounds__(128, PB_EMIT_MINB)
#else
#define PB_EMIT_BOUNDS __launch_bounds__(128)
#endif
#ifndef PB_LOOP_MINB
#define PB_LOOP_MINB 6
#endif
#ifndef PB_LOOP_CAP_DEFAULT
#ifndef PB_LOOP_EARLY_MCS
#define PB_LOOP_EARLY_MCS 1              // substep(): multiple-scattering draw issued at the top of the iteration (ILP)
#endif
#ifndef PB_LOOP_CG
#define PB_LOOP_CG 1                     // k_loop's 16-byte record copies bypass L1 (cp.async.cg): the 23 KB of L1 stay with the tables (49.2 -> 48.6 ms per config-2 step)
#endif
#ifndef PB_DRAIN_LANES_DEFAULT
#define PB_DRAIN_LANES_DEFAULT 16        // k_loop: once the charged list is dry, a warp with at most this many live tracks pauses them (carry-over) instead of draining at falling occupancy; 0 = off
#endif
#define PB_LOOP_CAP_DEFAULT 32           // sub-steps a track may take per k_loop launch before it is carried into the next wave (0: no limit); measured 16 / 32 / 64 / 128 / off: 155.3 / 145.4 / 146.3 / 150.0 / 151.2 ms per config-2 step
#endif
#ifndef PB_SAMPLE_G_DEFAULT
#define PB_SAMPLE_G_DEFAULT 4
#endif
#ifndef PB_SAMPLE_T_DEFAULT
#define PB_SAMPLE_T_DEFAULT 2            // SM pass: two trials per lane and round (ILP)
#endif
#ifndef PB_SAMPLE_T_DARK_DEFAULT
#define PB_SAMPLE_T_DARK_DEFAULT 1       // generic kernel (dark pass, stand-alone sampling)
#endif
constexpr int TILE = PB_TILE;               // samples per tile
constexpr int SAMPLE_THREADS = PB_SAMPLE_THREADS;
constexpr int GRID_SMEM_DOUBLES = 4096;     // >= padded stride of a 4-D map row (3964 -> 3968)

// Coarse look-up ("which node can E belong to at most") keyed by the top bits of the IEEE representation: key(E) = hi32(E) >> 14 is
// monotone in E > 0 and splits every octave into 64 bins; coarse[key(E) - key0] = searchsorted_left(x, upper edge of the bin) is an
// UPPER bound of searchsorted_left(x, E), and walking it down to the first node below E gives the exact index after one or two
// L1-resident loads - no logarithm, no binary search, no rounding question (integer compares only).
constexpr int COARSE_SHIFT = 14;
struct Coarse { const int* ub; int key0, nb; };
struct NSigmaTable { const double4* node; double xmin, xmax; Coarse c; int n; int pad; };   // node = (x, y, slope to next, 0)
struct MapInfo {
  const double* grid;   // nE rows, `stride` doubles each (padded to a multiple of 16 doubles)
  const double* E;
  const double* maxF;
  int nE, dim, stride, B;
  double invB;               // 1 / B (points per sweep): wgt = jac / B
  Coarse c;                  // coarse look-up over the row energies
  int ninc[4];
  double dninc[4];           // ninc as doubles (the trial multiplies by it twice: no int->fp64 conversion in the loop)
  int off[4];
};
struct Tables {
  NSigmaTable ns[16];
  NSigmaTable sp[3];         // total n*sigma of a charged species (e-, e+, mu+-) on the union grid of its processes: the sub-step
                             // loop only needs the sum, one node load and one hint instead of two or three (build_species_tables)
  MapInfo map[N_SAMPLED];
};

struct DarkTables {
  NSigmaTable w[4];          // brem_elec, brem_positron, annihilation, muon_brem
  NSigmaTable nsdark_comp;   // log10 nodes
  const double* dE[4];       // dRate/dE: saved energies ...
  const double* dT[4];       // ... and [n][10][2] tables, same order as w
  int dn[4];
  double min_E[4];           // DarkBrem, DarkAnn, DarkComp, DarkMuonBrem
};
struct DarkCand {            // candidate dark emissions of one pb_run_dark call (capacity 2 x n_sm)
  int* slot;                 // SM record
  int* proc;
  double* wg;                // GetBSMWeights value
  double* pf;                // [4] parent four-momentum at the interaction point
  int* ntr;
  int* count;                // [1]
};

struct Stack {
  double* p0; double* r0w; double* pf; double* rf;
  int4* ids;                 // 2 x int4 per record: (pid, parent, info, shower), (key.x, key.y, weight lo, weight hi) - ONE 32-byte
                             // sector holds everything k_emit / k_finalize / k_loop need besides the four-vectors
  int2* aux;
  long long capacity;
};
__device__ __forceinline__ int4 ld_meta(const Stack& S, long long s) { return S.ids[2 * s]; }
__device__ __forceinline__ int4 ld_kw(const Stack& S, long long s) { return S.ids[2 * s + 1]; }       // key + weight
__device__ __forceinline__ uint2 kw_key(int4 kw) { return make_uint2((uint32_t)kw.x, (uint32_t)kw.y); }
__device__ __forceinline__ double kw_weight(int4 kw) { return __hiloint2double(kw.w, kw.z); }
__device__ __forceinline__ void st_ids(Stack& S, long long s, int4 meta, uint2 key, double w) {
  S.ids[2 * s] = meta;
  S.ids[2 * s + 1] = make_int4((int)key.x, (int)key.y, __double2loint(w), __double2hiint(w));
}

// Wave bookkeeping kept on the device so that waves can be enqueued back to back without a host round trip.
struct WaveState {
  long long begin, end;      // stack slots of the current wave
  long long capacity;        // of the particle stack
  long long tot_charged;
  int n, n_charged;          // entries of the current wave (new records + carried tracks), size of its charged list
  int n_new, n_carry;        // new records [begin, begin + n_new) have wave-local index = slot - begin; the carried tracks (sub-step loop
                             // paused in an earlier wave, see k_loop) follow with local indices n_new .. n - 1, slot = carry list entry
  int parity;                // which pair of index lists the current wave reads
  int status;                // 0 running, 1 finished (empty wave), 2 stack capacity exhausted, 3 scratch too small (host grows it)
  int waves, max_wave;
  int work_cap, order_cap;   // capacities of the per-wave scratch / of each index list
  int iters, pad;            // k_wave_begin calls of this run (= wave-kernel sequences launched, the last ones empty)
};

struct Work {            // per-wave scratch, sized to the widest wave seen so far
  int* bucket;           // [n] bucket of particle (begin + i)
  int2* sorted;          // [n] (wave-local index, bucket) in bucket order: k_emit reads ONE coalesced array instead of sorted -> bucket
  double* xs;            // [n x 4] accepted sample (map variables)
  double* sE;            // [n] incoming energy, gathered into bucket order by k_bucket_fill (contiguous per tile)
  uint2* skey;           // [n] particle key, likewise
  int* hist;             // [NBUCKET]
  int* offsets;          // [NBUCKET + 1]
  int* cursor;           // [NBUCKET]
  int* tile_base;        // [2][NBUCKET + 1] exclusive prefix of the per-bucket tile counts (k_bucket_scan), heavy buckets (class 0: their
                         // tiles come first) and the others (class 1); k_bucket_fill expands it into the tile table:
  int* tile_sz;          // [NBUCKET] samples per tile of the bucket in this wave (cost-normalised, see k_bucket_scan)
  unsigned long long* bstat;   // [NBUCKET][2] accepted samples and trials the sampler has spent in the bucket so far (all waves, all runs)
  int* tile_bucket;      // [max_tiles]
  int* tile_start;
  int* tile_count;
  int* ctrl;             // [0] n_tiles, [1] tile cursor, [2] k_loop chunk cursor
  int* list[4];          // wave-local index lists, two parities x {charged (dE/dx-stepping), everything else}: the wave
                         // reads list[2*parity + k], k_emit builds list[2*(parity^1) + k] for the next wave
  int* carry[2];         // slots of the tracks carried into the current wave (carry[parity]) / paused in it (carry[parity ^ 1])
  int loop_cap;          // sub-steps a track may take per launch (power of two; 0 = unlimited): see k_loop
  int drain_lanes;       // wide waves: live tracks per warp at or below which a warp whose list is dry pauses them (0 = off): see k_loop
  struct WaveState* ws;  // device-resident wave bookkeeping (lets the host enqueue several waves per synchronisation)
  unsigned long long* tail;      // [0] stack tail (next free record); [1] = (n_neutral_next << 32) | n_charged_next; [2] tracks paused by this wave
  unsigned long long* counters;  // [CNT_N]: steps, substeps, samples, trials, no_sample, overflow, per-process trials/samples
  unsigned long long* tlog;      // measurement aid (PB_TILE_LOG): per tile of ONE chosen wave {start ns, end ns, bucket << 32 | count, SM}; else nullptr
  int tlog_wave, tlog_cap;
  int tile_norm;                 // 1: cost-normalised tiles, heavy buckets first (k_bucket_scan); 0 (PB_TILE_NORM=0): TILE samples per tile, bucket order
};

enum { CNT_STEPS = 0, CNT_SUBSTEPS, CNT_SAMPLES, CNT_TRIALS, CNT_NOSAMPLE, CNT_OVERFLOW, CNT_PROC_TRIALS = 8,
       CNT_PROC_SAMPLES = 24, CNT_N = 40 };

__device__ __forceinline__ bool is_charged(int pid) { return pid == 11 || pid == -11 || pid == 13 || pid == -13; }
// stack slot of wave-local entry i (WaveState::n_new)
__device__ __forceinline__ long long wave_slot(const Work& W, long long begin, int n_new, int parity, int i) {
  return i < n_new ? begin + i : (long long)W.carry[parity][i - n_new];
}
constexpr int AUX_PAUSED = -2;     // aux.x of a track whose sub-step loop was paused in this wave (k_finalize skips it)
__device__ __forceinline__ int pack_info(int gen, int child_bit, int flags, int process) {
  return (gen << 16) | (child_bit << 15) | ((flags & 0x7f) << 8) | (process & 0xff);
}

// ------------------------------------------------------------------------------------------ n*sigma(E)
// scipy interp1d (linear, bounds_error=False, fill_value=0) as used by shower.py:280-295.  Nodes are packed as
// (x_i, y_i, slope_i = (y_{i+1}-y_i)/(x_{i+1}-x_i), 0): one 32-byte sector per evaluation, no division.
// `hi` follows scipy: hi = clip(searchsorted_left(x, E), 1, n-1), lo = hi - 1, y = slope_lo * (E - x_lo) + y_lo.
__device__ __forceinline__ double4 ld_node(const double4* p) {
  const double2* q = reinterpret_cast<const double2*>(p);
  double2 a = __ldg(q), b = __ldg(q + 1);
  return make_double4(a.x, a.y, b.x, b.y);
}
__device__ __forceinline__ int nsigma_locate(const NSigmaTable& T, double E) {
  int lo = 0, hi = T.n;
  while (lo < hi) {                       // searchsorted(side='left')
    int mid = (lo + hi) >> 1;
    if (__ldg(reinterpret_cast<const double*>(&T.node[mid])) < E) lo = mid + 1; else hi = mid;
  }
  return min(max(lo, 1), T.n - 1);
}
__device__ __forceinline__ int coarse_ub(const Coarse& c, double E) {
  int b = (__double2hiint(E) >> COARSE_SHIFT) - c.key0;
  return __ldg(c.ub + min(max(b, 0), c.nb - 1));
}
// same result as nsigma_locate through the coarse table (exact for any increasing grid of positive energies)
__device__ __forceinline__ int nsigma_locate_c(const NSigmaTable& T, double E) {
  int hi = coarse_ub(T.c, E);
  while (hi > 1 && !(__ldg(reinterpret_cast<const double*>(&T.node[hi - 1])) < E)) --hi;
  return hi;
}
__device__ __forceinline__ double nsigma_at(const NSigmaTable& T, int hi, double E) {
  if (T.n < 2) return 0.0;
  if (!(E >= T.xmin && E <= T.xmax)) return (E == E) ? 0.0 : E;
  double4 nd = ld_node(&T.node[hi - 1]);
  return __dadd_rn(__dmul_rn(nd.z, E - nd.x), nd.y);
}
__device__ __forceinline__ double nsigma_eval(const NSigmaTable& T, double E) {
  if (T.n < 2) return 0.0;
  return nsigma_at(T, nsigma_locate(T, E), E);
}
// energy only ever decreases along a track: walk the hint down instead of searching again
__device__ __forceinline__ double nsigma_hinted(const NSigmaTable& T, int& hi, double E) {
  if (T.n < 2) return 0.0;
  if (!(E >= T.xmin && E <= T.xmax)) return (E == E) ? 0.0 : E;
  double4 nd = ld_node(&T.node[hi - 1]);
  while (hi > 1 && !(nd.x < E)) { --hi; nd = ld_node(&T.node[hi - 1]); }
  return __dadd_rn(__dmul_rn(nd.z, E - nd.x), nd.y);
}

// tables entering the mean free path of a charged species, in the reference's summation order (shower.py:357-368)
__device__ __forceinline__ void species_tables(int pid, int* t) {
  if (pid == 11) { t[0] = P_BREM; t[1] = P_MOLLER; t[2] = -1; }
  else if (pid == -11) { t[0] = P_BREM; t[1] = P_BHABHA; t[2] = P_ANN; }
  else { t[0] = P_MUONBREM; t[1] = P_MUONE; t[2] = -1; }
}
__device__ __forceinline__ int species_index(int pid) { return pid == 11 ? 0 : (pid == -11 ? 1 : 2); }
__device__ __forceinline__ double mfp_from(double ns) { return (ns <= 0.0) ? 1.0e12 : kCmToM / ns; }   // shower.py:386-389

__device__ __forceinline__ double nsigma_c(const NSigmaTable& T, double E) {
  if (T.n < 2) return 0.0;
  if (!(E >= T.xmin && E <= T.xmax)) return (E == E) ? 0.0 : E;
  int hi = coarse_ub(T.c, E);
  double4 nd = ld_node(&T.node[hi - 1]);
  while (hi > 1 && !(nd.x < E)) { --hi; nd = ld_node(&T.node[hi - 1]); }
  return __dadd_rn(__dmul_rn(nd.z, E - nd.x), nd.y);
}

// Up to three tables at the SAME energy (the species' processes in the final step and in the process choice).  K separate nsigma_c
// calls are a chain of 2 K dependent look-ups (coarse table -> node, then the rare walk), and k_finalize spends half of its stall
// samples waiting on such chains (profiles/r02f): here the K coarse look-ups are issued together, then the K node loads, then the
// walks.  Same nodes, same two roundings per table as nsigma_c.  t[k] < 0: no table (result 0).
__device__ __forceinline__ void nsigma_c3(const Tables& T, const int* t, int nt, double E, double* out) {
  bool ok[3]; int h[3]; double4 nd[3];
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    ok[k] = k < nt && t[k] >= 0 && T.ns[t[k] < 0 ? 0 : t[k]].n >= 2 && E >= T.ns[t[k] < 0 ? 0 : t[k]].xmin && E <= T.ns[t[k] < 0 ? 0 : t[k]].xmax;
    h[k] = 1;
  }
#pragma unroll
  for (int k = 0; k < 3; ++k) if (ok[k]) h[k] = coarse_ub(T.ns[t[k]].c, E);
#pragma unroll
  for (int k = 0; k < 3; ++k) nd[k] = ok[k] ? ld_node(&T.ns[t[k]].node[h[k] - 1]) : make_double4(0.0, 0.0, 0.0, 0.0);
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    if (ok[k]) {
      const NSigmaTable& Tk = T.ns[t[k]];
      while (h[k] > 1 && !(nd[k].x < E)) { --h[k]; nd[k] = ld_node(&Tk.node[h[k] - 1]); }
      out[k] = __dadd_rn(__dmul_rn(nd[k].z, E - nd[k].x), nd[k].y);
    } else out[k] = (k < nt && t[k] >= 0 && T.ns[t[k]].n >= 2 && !(E == E)) ? E : 0.0;      // out of range: 0; NaN stays NaN (scipy fill_value)
  }
}

// SURVEY Q-1: argmin |E_i - E| + 1, clamped to the last row (shower.py:416-426).  lo = first row with E_row >= E comes from the
// coarse table (upper bound) walked down.
__device__ __forceinline__ int lookup_row(const MapInfo& m, double E) {
  int n = m.nE;
  int lo = coarse_ub(m.c, E);
  while (lo > 0 && !(__ldg(m.E + lo - 1) < E)) --lo;
  int best;
  if (lo <= 0) best = 0;
  else if (lo >= n) best = n - 1;
  else best = (fabs(__ldg(m.E + lo - 1) - E) <= fabs(__ldg(m.E + lo) - E)) ? lo - 1 : lo;   // argmin keeps the first minimum
  int lu = best + 1;
  return lu >= n ? n - 1 : lu;
}

// ------------------------------------------------------------------------------------------ wave kernels
// One dE/dx + multiple-scattering sub-step state of a charged track (shower.py:559-581)
struct Track {
  V4 p; double rx, ry, rz;
  double mass, iKp, pmin, delta_z, pn, ipn;     // iKp = mass / (1e3 m_species): 1 / (momentum scale of the MCS width)
  uint2 key;
  int sp, hint;    // species table (Tables::sp) and the look-up hint into it
  int it;          // loop iterations done == accepted sub-steps while the loop is alive
};

// Track set-up computed where the particle is created (k_emit / k_init_primaries, all lanes busy) instead of at refill
// time inside k_loop (3 of 32 lanes busy): |p| and the species-table hint, parked in the record's not-yet-used rf slot.
__device__ __forceinline__ void store_track_setup(const Material& M, const Tables& T, Stack& S, long long slot, int pid, double mass,
                                                  double E, double px, double py, double pz) {
  const NSigmaTable& Ts = T.sp[species_index(pid)];
  int h = Ts.n >= 2 ? coarse_ub(Ts.c, E) : 1;          // an upper bound of the node index: nsigma_hinted walks it down (exactly) at the first sub-step
  double pn = norm3_nofma(px, py, pz);
  double pmin = fmax(fmax(M.min_calc[pid_class(pid)], M.min_energy), mass);           // shower.py:532-533
  double2* rfp = reinterpret_cast<double2*>(S.rf + 4 * slot);
  rfp[0] = make_double2(pn, __hiloint2double(0, h));
  rfp[1] = make_double2(pmin, 1.0 / pn);
}

// ---- draw sources.  Every decision function below takes its random numbers from a "draw source": PhiloxDraws is the engine's
// protocol (each draw a pure function of (particle key, index, stream)); TapeDraws (replay mode, pb_replay) hands out
// the numbers a REFERENCE run consumed, in the reference's own order, so that the same device code can be
// driven by the reference's uniform stream.
struct PhiloxDraws {
  static constexpr bool kOrderFree = true;      // draws are pure functions of (key, index): they may be computed early, or for nothing
  uint2 key;
  __device__ __forceinline__ D2 substep(uint32_t it) { D2 u = draw2(key, it, ST_SUBSTEP); return D2{u.a, 6.0 + 14.0 * u.b}; }   // u_hard, U(6, 20)
  __device__ __forceinline__ McsDraw mcs(uint32_t it, uint32_t pc) { return mcs_draw(key, it, pc); }
  __device__ __forceinline__ double final_u() { return draw2(key, 0, ST_FINAL).a; }
  __device__ __forceinline__ double choice_u() { return draw2(key, 0, ST_CHOICE).a; }
  __device__ __forceinline__ double kin_u(int proc) { return draw2(key, 0, ST_KIN, 0, proc).a; }
  __device__ __forceinline__ D2 decay_u(int proc) { return draw2(key, 0, ST_DECAY, 0, proc); }
  __device__ __forceinline__ D2 decay_x(uint32_t loop, uint32_t i) { return draw2(key, i, ST_DECAY, loop, P_SMDECAY); }   // i-th (x, u) pair of accept/reject loop 1..3 of a decay in flight
};
struct TapeDraws {             // sequential reader; `over` is set if the tape runs out (a decision differed from the recording)
  static constexpr bool kOrderFree = false;
  const double* t; long long pos, end; bool over;
  __device__ __forceinline__ double next() { if (pos >= end) { over = true; return 0.5; } return t[pos++]; }
  __device__ __forceinline__ D2 substep(uint32_t) { double a = next(), b = next(); return D2{a, b}; }                 // random(), uniform(6, 20)
  __device__ __forceinline__ McsDraw mcs(uint32_t, uint32_t) {                                                       // choice, gauss, gauss, uniform(0, 2 pi) / 2 pi
    McsDraw d; d.sign = next(); double z1 = next(), z2 = next(); d.radial = sqrt(z1 * z1 + z2 * z2); d.uphi = next(); return d;
  }
  __device__ __forceinline__ double final_u() { return next(); }
  __device__ __forceinline__ double choice_u() { return next(); }
  __device__ __forceinline__ double kin_u(int) { return next(); }
  __device__ __forceinline__ D2 decay_u(int) { double a = next(), b = next(); return D2{a, b}; }
  __device__ __forceinline__ D2 decay_x(uint32_t, uint32_t) { double a = next(), b = next(); return D2{a, b}; }
};

// One iteration of the dE/dx + multiple-scattering loop (shower.py:559-581) on a track: true = the loop ends here (energy
// below threshold, or the hard scatter was drawn), false = one sub-step was applied.  Shared by k_loop and PB_PROBE_SUBSTEP.
template <class DS>
__device__ __forceinline__ bool substep(const Material& M, const Tables& T, Track& t, int ms_e, DS& ds) {
  if (!(t.p.E >= t.pmin)) return true;                                // loop condition (shower.py:559)
#if PB_LOOP_EARLY_MCS
  // the multiple-scattering draw of this sub-step, issued before the energy-loss chain instead of after it: its Philox rounds (integer
  // pipe) and the log / sqrt of the radial variable overlap the dependent FP64 chain below (the loop is latency-bound at 5-6 warps
  // per scheduler).  Wasted when the loop ends in this iteration (1 in ~9).  k_loop: 80 registers + 16 B spilled -> 74, no spill;
  // 47.3 -> 46.5 ms per config-2 step.  (Resolving the hard-scatter branch after the energy-loss chain as well: no further gain.)
  McsDraw d_early{0.0, 0.0, 0.0};
  if (DS::kOrderFree && ms_e) d_early = ds.mcs((uint32_t)t.it, 0);
#endif
  double ns = nsigma_hinted(T.sp[t.sp], t.hint, t.p.E);                // sum over the species' processes (shower.py:357-368)
  double mfp = (ns <= 0.0) ? 1.0e12 : kCmToM * fast_rcp(ns);          // shower.py:386-389
  D2 u = ds.substep((uint32_t)t.it);
  double iv = fast_rcp(u.b);                                          // delta_z = mfp / U(6, 20); delta_z / mfp = 1 / U
  t.delta_z = mfp * iv;
  if (u.a > hot_exp_neg_step(iv)) return true;                        // hard scatter (shower.py:564)
  // lose_energy (particle.py:143-153) with |p| carried along the track instead of recomputed
  double Eu = t.p.E - M.dEdx * t.delta_z;
  if (Eu <= t.mass) Eu = t.mass;
  double p3f = fast_sqrt0(__dsub_rn(__dmul_rn(Eu, Eu), __dmul_rn(t.mass, t.mass)));      // exactly 0 at Eu == mass, as the reference's sqrt
  if (p3f > 0.0) {
    double r = p3f * t.ipn;
    t.p = V4{Eu, t.p.x * r, t.p.y * r, t.p.z * r};
    t.pn = p3f;
    double inv = fast_rcp(p3f);
    t.ipn = inv;
    double s = t.delta_z * inv;
    t.rx += t.p.x * s; t.ry += t.p.y * s; t.rz += t.p.z * s;
    if (ms_e) {
#if PB_LOOP_EARLY_MCS
      McsDraw d = DS::kOrderFree ? d_early : ds.mcs((uint32_t)t.it, 0);
#else
      McsDraw d = ds.mcs((uint32_t)t.it, 0);
#endif
      t.p = mcs_fast(M, t.p, p3f, inv, M.rho * (t.delta_z * (1.0 / kCmToM)), t.iKp, d.sign, d.radial, d.uphi);
    }
  } else {
    t.p = V4{t.mass, 0.0, 0.0, 0.0};
    t.pn = 0.0; t.ipn = 0.0;
  }
  ++t.it;
  return false;
}

// First kernel of every wave: turn what the previous wave appended (stack tail, list sizes) into this wave's extent.
// Idempotent when it has to pause (status 3), so the host can grow the scratch and re-enqueue the same wave.
// Inside the wave-loop graph (pb_run_showers) it also drives the WHILE node: the loop goes on as long as this wave is a real one.
__device__ __forceinline__ void wave_begin(Work& W) {
  WaveState& ws = *W.ws;
  ws.iters += 1;
  if (ws.iters > (1 << 20)) ws.status = 4;                 // safety net for the device-side loop: no shower has a million waves
  if (ws.status != 0 && ws.status != 3) { ws.n = 0; ws.n_charged = 0; ws.n_new = 0; ws.n_carry = 0; return; }
  unsigned long long tail = W.tail[0], lists = W.tail[1], carried = W.tail[2];
  long long begin = ws.status == 3 ? ws.begin : ws.end;
  long long end = (long long)min(tail, (unsigned long long)ws.capacity);
  long long n_new = end - begin, n = n_new + (long long)carried;
  ws.status = 0;
  if (n <= 0) { ws.status = 1; ws.n = 0; ws.n_charged = 0; ws.n_new = 0; ws.n_carry = 0; return; }
  if (end + 2 * n > ws.capacity) { ws.status = 2; ws.n = 0; ws.n_charged = 0; ws.n_new = 0; ws.n_carry = 0; return; }
  if (n > ws.work_cap || 2 * n > ws.order_cap) { ws.status = 3; ws.begin = begin; ws.n = 0; ws.n_charged = 0; ws.n_new = 0; ws.n_carry = 0; return; }
  ws.begin = begin; ws.end = end;
  ws.n = (int)n; ws.n_new = (int)n_new; ws.n_carry = (int)carried;
  ws.n_charged = (int)(lists & 0xffffffffull);
  ws.parity ^= 1;
  ws.tot_charged += ws.n_charged;
  ws.waves += 1;
  ws.max_wave = max(ws.max_wave, (int)n);
  W.tail[1] = 0; W.tail[2] = 0;
  W.ctrl[1] = 0; W.ctrl[2] = 0; W.ctrl[4] = 0; W.ctrl[5] = 0; W.ctrl[6] = 0;
}
__global__ void k_wave_begin(Work W) { wave_begin(W); }
__global__ void k_wave_begin_cond(Work W, cudaGraphConditionalHandle h) {
  wave_begin(W);
  cudaGraphSetConditional(h, W.ws->status == 0 ? 1u : 0u);     // status 1 / 2 / 3: this iteration's kernels find n = 0, then the loop ends
}

// Sub-step loop of propagate_particle, charged species only.  Persistent warps pull 32-entry chunks of the wave's charged
// list; a lane that finishes its track (hard scatter drawn, or energy below threshold) stores it and immediately
// takes the next entry of the chunk, so the warp stays converged on the loop body whatever the per-track sub-step
// count (geometric, mean ~7, tail > 50).  The final partial step and the process choice are done by k_finalize.
//
// Refill runs with ~3 of 32 lanes, so it must not wait on HBM: each warp keeps a three-deep software pipeline --
// chunk cursor (atomic, consumed one chunk later) -> the chunk's list indices (register, consumed one chunk later) ->
// the chunk's records copied asynchronously (cp.async / LDGSTS, no registers) into the idle half of a per-warp
// double buffer in shared memory -> consumed with shared-memory reads.
#ifndef PB_LOOP_CHUNK
#define PB_LOOP_CHUNK 32
#endif
constexpr int LOOP_CHUNK = PB_LOOP_CHUNK;      // entries per pipeline stage (<= 32: one per lane)
struct LoopBuf {                 // one chunk of track records: p0, r0w, track set-up (rf), ids  (a carried track: pf, rf, -, ids, aux)
  double2 v[6][LOOP_CHUNK];
  int4 meta[LOOP_CHUNK];
  int4 kw[LOOP_CHUNK];
  int2 aux[LOOP_CHUNK];
  int idx[LOOP_CHUNK];
};
__device__ __forceinline__ void cp_async16(void* dst, const void* src) {
#if PB_LOOP_CG
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((uint32_t)__cvta_generic_to_shared(dst)), "l"(src) : "memory");
#else
  asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" ::"r"((uint32_t)__cvta_generic_to_shared(dst)), "l"(src) : "memory");
#endif
}
__device__ __forceinline__ void cp_async8(void* dst, const void* src) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"((uint32_t)__cvta_generic_to_shared(dst)), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }

// **Bounded sub-steps per launch (carry-over).**  The sub-step count of a track is geometric (mean ~9, 1 track in 40 above 32, the
// longest of a million above 150), a track cannot be split, and a wave used to end with its longest track: 50-130 us of idle SMs
// per launch, and in narrow waves the whole critical path.  With Work::loop_cap = K (a power of two) a track is PAUSED when its
// sub-step index reaches a multiple of K: its state is stored where the finished state would go (pf, rf with delta_z, aux.y = index,
// aux.x = AUX_PAUSED), its slot is appended to the carry list of the NEXT wave, and that wave's k_loop resumes it next to its own
// fresh tracks.  k_finalize skips paused tracks; a carried track that finishes joins the finalize -> bucket -> sample -> emit flow of
// the wave it finishes in (wave-local index n_new + k, see WaveState).  The draws are indexed by (particle key, sub-step index), so
// the track is the one an uninterrupted loop would have produced; daughters only appear a wave or two later.
__global__ void __launch_bounds__(128, PB_LOOP_MINB)
k_loop(const __grid_constant__ Material M, const __grid_constant__ Tables T, Stack S, Work W,
       const double* __restrict__ prim_mass, int ms_e) {
  __shared__ __align__(16) LoopBuf s_buf[4][2];
  const long long begin = W.ws->begin;
  const int n_charged = W.ws->n_charged, n_new = W.ws->n_new;
  const int n_carry = W.ws->n_carry;
  const int n_work = n_charged + n_carry;                  // the carried tracks first (they resume in full warps), then the fresh charged ones
  if (n_work <= 0) return;
  const int parity = W.ws->parity;
  const int* __restrict__ order_c = W.list[2 * parity];
  const int* __restrict__ carry_in = W.carry[parity];
  int* __restrict__ carry_out = W.carry[parity ^ 1];
  // **Drain pause.**  When the list is dry a warp finishes the <= 32 tracks it holds at falling occupancy: the slowest of 32 geometric
  // sub-step counts takes ~30 iterations where the work is worth ~8 (a mid-size wave ran 23.6 of 32 lanes, profiles/r02z).  In a wave wide
  // enough to have filled every warp, a warp that is down to drain_lanes live tracks pauses them like tracks at the cap: the next
  // wave resumes them FIRST, packed into full warps.  Scheduling only (see the carry-over note above).
  const int drain_lanes = (W.drain_lanes > 0 && n_work >= 2 * (int)(gridDim.x * blockDim.x)) ? W.drain_lanes : 0;
  // sub-step cap: in a wide wave the drain pause already removes the tail, and a cap of 2 K costs fewer extra waves (measured with
  // K = 32: 129.8 -> 128.7 ms per config-2 step, 123 -> 114 waves); narrow waves, where the cap IS the critical path, keep K
  const unsigned cap_mask = W.loop_cap > 0 ? (unsigned)(drain_lanes ? 2 * W.loop_cap : W.loop_cap) - 1u : 0xffffffffu;
  const int lane = threadIdx.x & 31;
  LoopBuf* buf = s_buf[threadIdx.x >> 5];
  const unsigned lt_mask = (1u << lane) - 1u;
  // small waves: smaller chunks, so that the three chunks every warp holds in its pipeline do not starve the other warps
  const int chunk = min(LOOP_CHUNK, max(1, n_work / (3 * (int)(gridDim.x * (blockDim.x / 32)))));
  // ---- pipeline state (warp-uniform unless noted)
  int which = 0;                 // half of the double buffer being consumed
  int pos = 0, cnt = 0;          // consumed / valid entries of that half
  int cnt_fly = 0;               // valid entries of the chunk being copied into the other half
  int pre_idx = -1, pre_cnt = 0; // (per lane) wave-local index of the chunk after that, and its size
  int c_raw = 0;                 // (lane 0) cursor of the chunk after that; read one chunk later
  bool dry = false;              // the list is exhausted
  auto fetch_cursor = [&]() { if (lane == 0) c_raw = atomicAdd(&W.ctrl[2], chunk); };
  auto fetch_index = [&]() {     // consumes c_raw, issues the index load
    int c = __shfl_sync(0xffffffffu, c_raw, 0);
    pre_cnt = min(max(n_work - c, 0), chunk);
    const int w = c + lane;
    pre_idx = (lane < pre_cnt) ? (w < n_carry ? n_new + w : order_c[w - n_carry]) : -1;
  };
  auto issue_copy = [&](int half) {   // consumes pre_idx, starts the record copies into buf[half]
    LoopBuf& B = buf[half];
    if (lane < pre_cnt) {
      const bool resumed = pre_idx >= n_new;
      long long s = resumed ? (long long)carry_in[pre_idx - n_new] : begin + pre_idx;
      const double2* pa = reinterpret_cast<const double2*>((resumed ? S.pf : S.p0) + 4 * s);
      const double2* pb = reinterpret_cast<const double2*>((resumed ? S.rf : S.r0w) + 4 * s);
      cp_async16(&B.v[0][lane], pa); cp_async16(&B.v[1][lane], pa + 1);
      cp_async16(&B.v[2][lane], pb); cp_async16(&B.v[3][lane], pb + 1);
      if (resumed) cp_async8(&B.aux[lane], S.aux + s);
      else {
        const double2* sup = reinterpret_cast<const double2*>(S.rf + 4 * s);     // store_track_setup
        cp_async16(&B.v[4][lane], sup); cp_async16(&B.v[5][lane], sup + 1);
      }
      cp_async16(&B.meta[lane], S.ids + 2 * s);
      cp_async16(&B.kw[lane], S.ids + 2 * s + 1);
      B.idx[lane] = pre_idx;
    }
    cnt_fly = pre_cnt;
  };
  auto advance = [&]() {          // the consumed half is empty: switch to the one in flight, refill the pipeline behind it
    cp_async_wait_all();
    __syncwarp();
    which ^= 1; pos = 0; cnt = cnt_fly;
    issue_copy(which ^ 1);
    fetch_index();
    fetch_cursor();
  };
  fetch_cursor(); fetch_index(); fetch_cursor();
  issue_copy(1); fetch_index(); fetch_cursor();
  int cur = -1;                       // wave-local index of the lane's track; -1: needs a track, -2: no more work
  Track t;
  unsigned long long c_sub = 0;
  for (;;) {
    // ---- refill
    unsigned need = __ballot_sync(0xffffffffu, cur == -1);
    while (need) {
      if (pos >= cnt) {
        if (!dry) { advance(); dry = (cnt == 0); }
        if (dry) { if (cur == -1) cur = -2; break; }
      }
      int avail = cnt - pos;
      int rank = __popc(need & lt_mask);
      bool take = (cur == -1) && rank < avail;
      if (take) {
        const LoopBuf& B = buf[which];
        const int e = pos + rank;
        cur = B.idx[e];
        double2 a0 = B.v[0][e], a1 = B.v[1][e], b0 = B.v[2][e], b1 = B.v[3][e];
        t.p = V4{a0.x, a0.y, a1.x, a1.y};
        t.rx = b0.x; t.ry = b0.y; t.rz = b1.x;
        int4 meta = B.meta[e];
        t.key = kw_key(B.kw[e]);
        int pid = meta.x;
        t.mass = pid_mass(pid); t.iKp = 1e-3;
        if (meta.y < 0) {          // a primary: its slot is its index in the call's primary arrays
          t.mass = prim_mass[cur < n_new ? begin + cur : (long long)carry_in[cur - n_new]];
          t.iKp = t.mass / (1e3 * pid_mass(pid));
        }
        t.sp = species_index(pid);
        if (cur < n_new) {         // fresh track: set-up stored at creation (store_track_setup)
          double2 s0 = B.v[4][e], s1 = B.v[5][e];
          t.pmin = s1.x;
          t.pn = s0.x; t.ipn = s1.y;
          t.hint = __double2loint(s0.y);
          t.delta_z = 0.0; t.it = 0;
        } else {                   // carried track: the state k_loop stored when it paused it
          t.delta_z = b1.y;
          t.it = B.aux[e].y;
          t.pmin = fmax(fmax(M.min_calc[pid_class(pid)], M.min_energy), t.mass);
          // |p| and its reciprocal exactly as substep() carried them (a track is only ever paused after a sub-step): bit-identical resume
          t.pn = fast_sqrt0(__dsub_rn(__dmul_rn(t.p.E, t.p.E), __dmul_rn(t.mass, t.mass)));
          t.ipn = t.pn > 0.0 ? fast_rcp(t.pn) : 0.0;
          t.hint = T.sp[t.sp].n >= 2 ? coarse_ub(T.sp[t.sp].c, t.p.E) : 1;
        }
      }
      pos += min(__popc(need), avail);
      need = __ballot_sync(0xffffffffu, cur == -1);
    }
    if (__all_sync(0xffffffffu, cur == -2)) break;
    // ---- one sub-step for every live lane
    bool done = false, pause = false;
    const bool drain = dry && __popc(__ballot_sync(0xffffffffu, cur >= 0)) <= drain_lanes;     // warp-uniform
    if (cur >= 0) {
      PhiloxDraws ds{t.key};
      done = substep(M, T, t, ms_e, ds);
      if (!done) { ++c_sub; pause = drain || ((unsigned)t.it & cap_mask) == 0u; }     // only ever after a sub-step: the launch makes progress on every track
    }
    // tracks created together reach the cap together: one atomic per warp for the carry-list slots
    const unsigned pmask = __ballot_sync(0xffffffffu, pause);
    int pbase = 0;
    if (pmask) {
      const int leader = __ffs(pmask) - 1;
      if (lane == leader) pbase = (int)atomicAdd(&W.tail[2], (unsigned long long)__popc(pmask));
      pbase = __shfl_sync(0xffffffffu, pbase, leader);
    }
    if (done || pause) {
      long long s = cur < n_new ? begin + cur : (long long)carry_in[cur - n_new];
      double2* pfp = reinterpret_cast<double2*>(S.pf + 4 * s);
      double2* rfp = reinterpret_cast<double2*>(S.rf + 4 * s);
      pfp[0] = make_double2(t.p.E, t.p.x); pfp[1] = make_double2(t.p.y, t.p.z);
      rfp[0] = make_double2(t.rx, t.ry);   rfp[1] = make_double2(t.rz, t.delta_z);
      S.aux[s] = make_int2(pause ? AUX_PAUSED : 0, t.it);
      if (pause) carry_out[pbase + __popc(pmask & lt_mask)] = (int)s;
      cur = -1;
    }
  }
  for (int o = 16; o > 0; o >>= 1) c_sub += __shfl_down_sync(0xffffffffu, c_sub, o);
  if (lane == 0 && c_sub) atomicAdd(&W.counters[CNT_SUBSTEPS], c_sub);
}

// Second half of propagate_particle + the process choice for ONE particle: the final partial step of a charged track
// (shower.py:583-598) or the photon's free path (:538-553), then np.random.choice over the species' processes (:665-698) and the
// sample_scattering threshold (:469).  In: the state the sub-step loop left (charged) or the creation state (neutral).
// Returns the bucket (process * LU_MAX + map row; P_NONE: no hard scatter; P_SMDECAY: short-lived).
template <class DS>
__device__ __forceinline__ int finalize_one(const Material& M, const Tables& T, DS& ds, bool charged, int pid, int flags, double mass,
                                            double E_start, double delta_z, int ms_e, V4& p, double& rx, double& ry, double& rz, bool& stepped) {
  int bucket = P_NONE * LU_MAX;
  const int cls = pid_class(pid);
  stepped = false;
  if (charged) {
    double pmin = fmax(fmax(M.min_calc[cls], M.min_energy), mass);
    if (!(E_start < pmin)) {                                            // shower.py:534-536: otherwise untouched
      stepped = true;
      int tb[3];
      species_tables(pid, tb);
      double distC = ds.final_u();                                      // shower.py:583-598
      double last;
      if (p.E < pmin) last = distC * delta_z;
      else {
        double v[3];
        nsigma_c3(T, tb, 3, p.E, v);
        double ns = v[0] + v[1];
        if (tb[2] >= 0) ns += v[2];
        double mfp = mfp_from(ns);
        last = mfp * log(1.0 / (1.0 + (exp(-delta_z / mfp) - 1) * distC));
      }
      p = lose_energy(p, mass, M.dEdx * last);
      double pn = norm3_nofma(p.x, p.y, p.z);
      if (pn > 0.0) {
        double sc = last / pn;
        rx += p.x * sc; ry += p.y * sc; rz += p.z * sc;
        if (ms_e) {                                                     // SURVEY Q-12: electron mass here
          McsDraw d = ds.mcs(MCS_FINAL_INDEX, 0);
          // the folded form the sub-step loop uses (golden-checked against the reference like mcs_apply, test_mcs_fast_vs_reference_golden)
          p = mcs_fast(M, p, pn, fast_rcp(pn), M.rho * (last * (1.0 / kCmToM)), mass * (1.0 / (1e3 * kMe)), d.sign, d.radial, d.uphi);
        }
      }
    }
  } else {
    if (flags & (PB_FLAG_SHORT_LIVED | PB_FLAG_LONG_LIVED)) {
      bucket = P_SMDECAY * LU_MAX;                                      // particle.py:391-424, decays in k_emit
    } else if (pid == 22) {
      double pmin = fmax(fmax(M.min_calc[cls], M.min_energy), mass);
      if (!(p.E < pmin)) {                                              // shower.py:538-553 (MS_g is always False)
        stepped = true;
        const int tg[3] = {P_PAIRPROD, P_COMP, -1};
        double v[3];
        nsigma_c3(T, tg, 2, p.E, v);
        double mfp = mfp_from(v[0] + v[1]);
        double distC = ds.final_u();
        double dist = mfp * log(1.0 / (1.0 - distC));
        double pn = norm3_nofma(p.x, p.y, p.z);
        rx += p.x / pn * dist; ry += p.y / pn * dist; rz += p.z / pn * dist;
      }
    }
  }
  if (cls >= 0 && !(flags & (PB_FLAG_SHORT_LIVED | PB_FLAG_LONG_LIVED))) {
    // process choice (shower.py:665-698) and the sample_scattering threshold (shower.py:469)
    double Ef = p.E;
    int cand[3]; double c[3]; int nc;
    if (pid == 11) { cand[0] = P_BREM; cand[1] = P_MOLLER; nc = 2; }
    else if (pid == -11) { cand[0] = P_BREM; cand[1] = P_ANN; cand[2] = P_BHABHA; nc = 3; }
    else if (pid == 22) { cand[0] = P_PAIRPROD; cand[1] = P_COMP; nc = 2; }
    else { cand[0] = P_MUONE; cand[1] = P_MUONBREM; nc = 2; }
    double SC = 0.0;
    if (nc == 2) cand[2] = -1;
    nsigma_c3(T, cand, nc, Ef, c);
    for (int k = 0; k < nc; ++k) SC += c[k];
    if (!(SC == 0.0 || SC != SC)) {
      double u = ds.choice_u();
      // np.random.choice: cdf = cumsum(p); cdf /= cdf[-1]; searchsorted(u, 'right')
      double cdf[3]; double acc = 0.0;
      for (int k = 0; k < nc; ++k) { acc += c[k] / SC; cdf[k] = acc; }
      int pick = nc - 1;
      for (int k = nc - 1; k >= 0; --k) if (u < cdf[k] / acc) pick = k;
      int proc = cand[pick];
      double thr = fmax(fmax(M.min_calc[cls], M.min_energy), mass);
      if (!(Ef <= thr)) bucket = proc * LU_MAX + lookup_row(T.map[proc], Ef);
    }
  }
  return bucket;
}

// Per particle of the wave (charged list first, then the rest): finalize_one -> bucket + histogram.
__global__ void PB_FIN_BOUNDS
k_finalize(const __grid_constant__ Material M, const __grid_constant__ Tables T, Stack S, Work W,
           const double* __restrict__ prim_mass, int ms_e) {
  const long long begin = W.ws->begin;
  const int n = W.ws->n, n_charged = W.ws->n_charged, n_new = W.ws->n_new;
  const int* __restrict__ order_c = W.list[2 * W.ws->parity];
  const int* __restrict__ order_n = W.list[2 * W.ws->parity + 1];
  const int* __restrict__ carry = W.carry[W.ws->parity];
  unsigned long long c_steps = 0;
  // entries: the wave's charged list, its other records, then the carried tracks (all charged, wave-local index = their position)
  for (int j = blockIdx.x * blockDim.x + threadIdx.x; j < n; j += gridDim.x * blockDim.x) {
    const bool charged = j < n_charged || j >= n_new;
    int i = j < n_charged ? order_c[j] : (j < n_new ? order_n[j - n_charged] : j);
    long long s = j < n_new ? begin + i : (long long)carry[j - n_new];
    int4 meta = ld_meta(S, s);
    PhiloxDraws ds{kw_key(ld_kw(S, s))};
    int pid = meta.x;
    int flags = (meta.z >> 8) & 0x7f;
    double mass = (meta.y < 0) ? prim_mass[s] : pid_mass(pid);
    V4 p; double rx, ry, rz;
    int nsub = 0;
    bool paused = false;
    double delta_z = 0.0, E_start = 0.0;
    if (charged) {
      const double2* pfp = reinterpret_cast<const double2*>(S.pf + 4 * s);
      const double2* rfp = reinterpret_cast<const double2*>(S.rf + 4 * s);
      double2 a0 = pfp[0], a1 = pfp[1], b0 = rfp[0], b1 = rfp[1];
      p = V4{a0.x, a0.y, a1.x, a1.y};
      rx = b0.x; ry = b0.y; rz = b1.x;
      delta_z = b1.y;
      const int2 ax = S.aux[s];
      nsub = ax.y;
      // "never propagated" (creation energy below the species threshold, shower.py:534-536) without gathering the p0 sector for it:
      // a track that took no sub-step still has its creation energy, and one that took any started at or above the threshold
      E_start = nsub == 0 ? a0.x : HUGE_VAL;
      paused = ax.x == AUX_PAUSED;     // sub-step loop paused in this wave (k_loop): nothing to finalize yet, the record stays as k_loop
                                       // left it.  (Tested AFTER finalize_one: the 2 % of wasted evaluations cost less than a branch that
                                       // every track would wait on with the latency of this load.)
    } else {
      const double2* p0p = reinterpret_cast<const double2*>(S.p0 + 4 * s);
      const double2* r0p = reinterpret_cast<const double2*>(S.r0w + 4 * s);
      double2 a0 = p0p[0], a1 = p0p[1], b0 = r0p[0], b1 = r0p[1];
      p = V4{a0.x, a0.y, a1.x, a1.y};
      rx = b0.x; ry = b0.y; rz = b1.x;
    }
    bool stepped;
    int bucket = finalize_one(M, T, ds, charged, pid, flags, mass, E_start, delta_z, ms_e, p, rx, ry, rz, stepped);
    if (paused) bucket = P_NONE * LU_MAX;
    else {
      if (stepped) c_steps += 1;
      double2* pfp = reinterpret_cast<double2*>(S.pf + 4 * s);
      double2* rfp = reinterpret_cast<double2*>(S.rf + 4 * s);
      pfp[0] = make_double2(p.E, p.x); pfp[1] = make_double2(p.y, p.z);
      rfp[0] = make_double2(rx, ry);   rfp[1] = make_double2(rz, mass);
      S.aux[s] = make_int2(0, nsub);
    }
    W.bucket[i] = bucket;
    // warp-aggregated histogram: one atomic per distinct bucket in the warp
    unsigned peers = __match_any_sync(__activemask(), bucket);
    if ((threadIdx.x & 31) == __ffs(peers) - 1) atomicAdd(&W.hist[bucket], __popc(peers));
  }
  for (int o = 16; o > 0; o >>= 1) c_steps += __shfl_down_sync(0xffffffffu, c_steps, o);
  if ((threadIdx.x & 31) == 0 && c_steps) atomicAdd(&W.counters[CNT_STEPS], c_steps);
}

// Exclusive scan over NBUCKET bins + tile bookkeeping; one CTA of 1024 threads, NBUCKET/1024 bins per thread.
// Tiles are COST-normalised: the accept/reject cost of a sample differs by an order of magnitude between buckets (10 GeV photons
// in lead: Brem / PairProd rows need 12-40 trials per sample, the annihilation rows 100-170), and a 192-sample tile of such a row
// used to run 300-500 us at the end of a launch whose other CTAs had finished (profiles/r02_summary.md, tile timeline).  The engine
// therefore keeps, per bucket, the samples and trials spent so far (bstat), gives the heavy buckets (more than twice the launch's
// average trials per sample) proportionally smaller tiles and puts those tiles at the front of the tile table.  Scheduling only: the samples themselves
// are a pure function of (particle key, trial index).
constexpr int TILE_MIN = 32;
#ifndef PB_TILE_HEAVY_COST
#define PB_TILE_HEAVY_COST 4.f
#endif
constexpr float TILE_HEAVY_COST = PB_TILE_HEAVY_COST;     // a heavy bucket's tile may cost this many average tiles
__global__ void __launch_bounds__(1024) k_bucket_scan(Work W) {
  constexpr int PER = NBUCKET / 1024;
  __shared__ int s_cnt[1024], s_t0[1024], s_t1[1024];
  __shared__ float s_w[32], s_c[32];
  int t = threadIdx.x;
  int cnt[PER], til[PER], heavy[PER];
  float avg[PER];
  // expected trials per sample of every bucket (16 until the bucket has a history) and of this launch as a whole
  float wsum = 0.f, csamp = 0.f;
  for (int k = 0; k < PER; ++k) {
    int b = t * PER + k;
    int c = W.hist[b];
    cnt[k] = c; avg[k] = 16.f;
    if (b < N_SAMPLED * LU_MAX) {
      const unsigned long long ns = W.bstat[2 * b], nt = W.bstat[2 * b + 1];
      if (ns >= 64) avg[k] = fmaxf((float)nt / (float)ns, 1.f);
      wsum += avg[k] * (float)c; csamp += (float)c;
    }
  }
  for (int o = 16; o > 0; o >>= 1) { wsum += __shfl_down_sync(0xffffffffu, wsum, o); csamp += __shfl_down_sync(0xffffffffu, csamp, o); }
  if ((t & 31) == 0) { s_w[t >> 5] = wsum; s_c[t >> 5] = csamp; }
  __syncthreads();
  float wall = 0.f, call = 0.f;
  for (int k = 0; k < 32; ++k) { wall += s_w[k]; call += s_c[k]; }
  const float avg_all = call > 0.f ? wall / call : 16.f;
  // share of the launch's expected trials that sits in heavy buckets: rare stragglers (the annihilation rows of an SM wave, < 1 %) get
  // the special treatment below; when the heavy buckets ARE the work (DarkAnn in a dark pass: 90 %) they are scheduled like any other
  float hw = 0.f;
  for (int k = 0; k < PER; ++k) if (avg[k] > 2.f * avg_all && t * PER + k < N_SAMPLED * LU_MAX) hw += avg[k] * (float)cnt[k];
  for (int o = 16; o > 0; o >>= 1) hw += __shfl_down_sync(0xffffffffu, hw, o);
  __syncthreads();
  if ((t & 31) == 0) s_w[t >> 5] = hw;
  __syncthreads();
  float hall = 0.f;
  for (int k = 0; k < 32; ++k) hall += s_w[k];
  const bool norm = W.tile_norm && hall < 0.25f * wall;
  // the buckets more than twice as expensive per sample as the launch average ("heavy") go first, with tiles shrunk so that one of
  // them costs at most TILE_HEAVY_COST average tiles.  Not smaller: a tile ends when its slowest WARP is done, the trial count of a
  // sample is geometric, and the fewer samples a warp gets the larger the spread between the four warps (normalising every tile to
  // the average cost was measured 15 % slower on the dark pass of config 3, where DarkAnn needs 300 trials per sample)
  int csum = 0, t0sum = 0, t1sum = 0;
  for (int k = 0; k < PER; ++k) {
    int b = t * PER + k;
    int c = cnt[k];
    til[k] = 0; heavy[k] = 0;
    if (b < N_SAMPLED * LU_MAX) {
      int tsz = TILE;
      heavy[k] = norm && avg[k] > 2.f * avg_all;
      if (heavy[k]) tsz = max(TILE_MIN, min(TILE, (int)((float)TILE * TILE_HEAVY_COST * avg_all / avg[k]) & ~31));
      W.tile_sz[b] = tsz;
      til[k] = (c + tsz - 1) / tsz;
    }
    csum += c;
    if (heavy[k]) t0sum += til[k]; else t1sum += til[k];
    W.hist[b] = 0;
    W.cursor[b] = 0;
  }
  s_cnt[t] = csum; s_t0[t] = t0sum; s_t1[t] = t1sum;
  __syncthreads();
  for (int o = 1; o < 1024; o <<= 1) {           // Hillis-Steele inclusive scan
    int a = 0, b = 0, c = 0;
    if (t >= o) { a = s_cnt[t - o]; b = s_t0[t - o]; c = s_t1[t - o]; }
    __syncthreads();
    s_cnt[t] += a; s_t0[t] += b; s_t1[t] += c;
    __syncthreads();
  }
  int cbase = s_cnt[t] - csum, t0base = s_t0[t] - t0sum, t1base = s_t1[t] - t1sum;
  int* tb0 = W.tile_base; int* tb1 = W.tile_base + NBUCKET + 1;
  for (int k = 0; k < PER; ++k) {
    int b = t * PER + k;
    W.offsets[b] = cbase;
    tb0[b] = t0base; tb1[b] = t1base;          // the tile table itself is written by k_bucket_fill (wide), one thread per tile
    cbase += cnt[k];
    if (heavy[k]) t0base += til[k]; else t1base += til[k];
  }
  if (t == 1023) {
    W.offsets[NBUCKET] = cbase; tb0[NBUCKET] = t0base; tb1[NBUCKET] = t1base;
    W.ctrl[0] = t0base + t1base; W.ctrl[1] = 0; W.ctrl[4] = 0; W.ctrl[5] = 0; W.ctrl[6] = 0;
  }
  if (t == 0) W.ctrl[2] = 0;
}

// Where a sample's incoming energy / Philox key come from and where its trial count goes: the SM pass indexes the
// wave's stack records directly, the dark pass goes through its candidate list.
struct SampleIO {
  const double* E4;        // incoming energy of entry i at E4[4*i]
  const uint2* key;        // particle keys: the key of entry / record k sits at key[k * key_stride + key_off]
  int key_stride, key_off; // (1, 0) for a plain key array; (4, 2) for the ids records of a stack
  const int* key_index;    // entry i uses record key_index[i] (nullptr: record off + i)
  int* ntr;                // trials used by entry i at ntr[i * ntr_stride]; -1 if the sampler gave up
  int ntr_stride;
  const WaveState* ws;     // SM pass: entry i is stack slot ws->begin + i (E4/key/ntr then point at slot 0)
};

// n_explicit < 0: the wave size comes from the device-side wave state.  Besides the counting-sort scatter of the indices, the
// sampler's two per-entry inputs (incoming energy, Philox key) are gathered into bucket order here, where the reads are
// coalesced, so that k_sample stages a tile with contiguous loads instead of two dependent scattered ones per sample.
__global__ void __launch_bounds__(256) k_bucket_fill(Work W, SampleIO io, int n_explicit) {
  const int n = n_explicit < 0 ? W.ws->n : n_explicit;
  // tile table: tile t belongs to the bucket b with tile_base[b] <= t < tile_base[b + 1] (binary search over the 4097-entry prefix)
  const int n_tiles = W.ctrl[0], n_heavy = W.tile_base[NBUCKET];
  for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < n_tiles; t += gridDim.x * blockDim.x) {
    const int* tb = t < n_heavy ? W.tile_base : W.tile_base + NBUCKET + 1;      // heavy buckets' tiles first
    const int u = t < n_heavy ? t : t - n_heavy;
    int lo = 0, hi = NBUCKET;                       // last b with tb[b] <= u
    while (hi - lo > 1) { int mid = (lo + hi) >> 1; if (tb[mid] <= u) lo = mid; else hi = mid; }
    const int tsz = W.tile_sz[lo];
    const int start = W.offsets[lo] + (u - tb[lo]) * tsz;
    W.tile_bucket[t] = lo;
    W.tile_start[t] = start;
    W.tile_count[t] = min(tsz, W.offsets[lo + 1] - start);
  }
  // Counting-sort scatter, aggregated per CTA: the kernel used to be bound by the latency of same-address atomics (every warp of a
  // wide wave bumps the cursors of the same few hot buckets: 65 % of its stall samples waited on that atomic, 7 % of the issue slots
  // were busy).  A CTA now ranks FILL_CHUNK entries per bucket in shared memory first and makes ONE global atomic per bucket it saw.
  __shared__ int s_cnt[NBUCKET], s_base[NBUCKET];
  for (int k = threadIdx.x; k < NBUCKET; k += blockDim.x) s_cnt[k] = 0;
  __syncthreads();
  constexpr int FILL_PER = 4;
  const int chunk = FILL_PER * (int)blockDim.x;
  for (int c0 = blockIdx.x * chunk; c0 < n; c0 += gridDim.x * chunk) {
    int b[FILL_PER], r[FILL_PER];
    double E[FILL_PER]; uint2 key[FILL_PER];
#pragma unroll
    for (int e = 0; e < FILL_PER; ++e) {
      const int i = c0 + e * (int)blockDim.x + (int)threadIdx.x;
      b[e] = -1; r[e] = 0; E[e] = 0.0; key[e] = make_uint2(0, 0);
      if (i < n) {
        b[e] = W.bucket[i];
        r[e] = atomicAdd(&s_cnt[b[e]], 1);                     // rank among the CTA's entries of this bucket
        if (b[e] < N_SAMPLED * LU_MAX) {
          const size_t rec = io.ws ? (size_t)wave_slot(W, io.ws->begin, io.ws->n_new, io.ws->parity, i) : (size_t)i;   // SM pass: stack slot
          E[e] = io.E4[4 * rec];
          key[e] = io.key[(io.key_index ? (size_t)io.key_index[i] : rec) * io.key_stride + io.key_off];
        }
      }
    }
    __syncthreads();
#pragma unroll
    for (int e = 0; e < FILL_PER; ++e)
      if (b[e] >= 0 && r[e] == 0) s_base[b[e]] = atomicAdd(&W.cursor[b[e]], s_cnt[b[e]]);
    __syncthreads();
#pragma unroll
    for (int e = 0; e < FILL_PER; ++e) {
      if (b[e] < 0) continue;
      const int i = c0 + e * (int)blockDim.x + (int)threadIdx.x;
      const int pos = W.offsets[b[e]] + s_base[b[e]] + r[e];
      W.sorted[pos] = make_int2(i, b[e]);
      if (b[e] < N_SAMPLED * LU_MAX) { W.sE[pos] = E[e]; W.skey[pos] = key[e]; }
    }
    __syncthreads();
#pragma unroll
    for (int e = 0; e < FILL_PER; ++e)
      if (b[e] >= 0 && r[e] == 0) s_cnt[b[e]] = 0;
    __syncthreads();
  }
}

// ---- TMA (bulk async copy) helpers: global -> shared, completion on an mbarrier
__device__ __forceinline__ uint32_t smem_addr(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_addr(bar)), "r"(count));
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_addr(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void tma_bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(smem_addr(dst)), "l"(src), "r"(bytes), "r"(smem_addr(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_LOOP:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra DONE;\n"
      "bra WAIT_LOOP;\n"
      "DONE:\n"
      "}\n" ::"r"(smem_addr(bar)), "r"(parity) : "memory");
}

// T accept/reject trials (indices t0 .. t0 + T - 1 of one sample) evaluated by ONE lane in straight-line code: the T
// dependency chains interleave (the register file holds too few warps to hide an FP64 chain by thread-level parallelism
// alone).  Per trial: Philox doubles D[0..dim] = y_0..y_{dim-1}, u_accept; map y -> x through the staged grid (vegas
// AdaptiveMap: x = g[i] + (g[i+1]-g[i]) * (y*ninc - i), jac = prod ninc*(g[i+1]-g[i])); accept iff
// max_F * u < (jac / B) * f(x)  (shower.py:453-459).  Returns the accept mask (bit k = trial t0 + k) and the point of the
// LOWEST accepted trial in x.  FAM: -1 = every sampled process, 2 = the SM pass (4-D folded forms + the five 1-D SM processes),
// 3 = the dark-bremsstrahlung family (DarkBrem, DarkMuonBrem: the 3-D folded form only).
template <int DIM, int FAM, int T>
__device__ __forceinline__ unsigned trial_block(const Material& M, const MapInfo& mi, const double* __restrict__ g, int proc,
                                                double E, const SampleConst& sc, double maxF, uint2 key, uint32_t t0, double* x) {
  double D[T][DIM + 2];      // the uniforms PLUS ONE (doubles in [1, 2)): consumed as fma(D, s, -s) == (D - 1) * s, bit for bit
#pragma unroll
  for (int k = 0; k < T; ++k) {
    const uint32_t t = t0 + k;
    uint32_t s0, s1;
    if (DIM == 4) {          // two calls: four map coordinates, and the accept uniform from the calls' 2 x 24 spare bits
      D2 d0 = draw2s_p1(key, t, ST_VEGAS, 0, proc, s0), d1 = draw2s_p1(key, t, ST_VEGAS, 1, proc, s1);
      D[k][0] = d0.a; D[k][1] = d0.b; D[k][2] = d1.a; D[k][3] = d1.b;
      D[k][4] = u48p1(s0, s1);
    } else {
#pragma unroll
      for (int j = 0; j < (DIM + 2) / 2; ++j) {
        D2 d = draw2s_p1(key, t, ST_VEGAS, j, proc, s0);
        D[k][2 * j] = d.a; D[k][2 * j + 1] = d.b;
      }
    }
  }
  double xx[T][4], jac[T];
#pragma unroll
  for (int k = 0; k < T; ++k) { jac[k] = 1.0; xx[k][0] = 0.0; xx[k][1] = 0.0; xx[k][2] = 0.0; xx[k][3] = 0.0; }
#pragma unroll
  for (int d = 0; d < DIM; ++d) {
    const double dn = mi.dninc[d];
    const int nm1 = mi.ninc[d] - 1;
    const double* gd = g + mi.off[d];
#pragma unroll
    for (int k = 0; k < T; ++k) {
      double yn = fma(D[k][d], dn, -dn);        // y * ninc
      int iy = min((int)yn, nm1);
      double g0 = gd[iy], g1 = gd[iy + 1];
      double inc = g1 - g0;
      xx[k][d] = __dadd_rn(g0, __dmul_rn(inc, yn - iy));
      jac[k] *= inc * dn;
    }
  }
  double f[T];
  if (DIM == 4) {
    if (proc == P_PAIRPROD) {
#pragma unroll
      for (int k = 0; k < T; ++k) f[k] = ds_pairprod_fast(sc, E, xx[k]);
    } else {
      const double ml = (proc == P_BREM) ? kMe : kMmu;
#pragma unroll
      for (int k = 0; k < T; ++k) f[k] = ds_brem_fast(M, sc, E, ml, xx[k]);
    }
  } else if (FAM == 2) {       // the five 1-D SM processes only
#pragma unroll
    for (int k = 0; k < T; ++k) {
      switch (proc) {
        case P_COMP: f[k] = ds_compton(M, E, 0.0, xx[k][0]); break;
        case P_ANN: f[k] = ds_annihilation(E, 0.0, M.Eg_min, xx[k][0]); break;
        case P_MOLLER: f[k] = ds_moller(E, M.Ee_min, xx[k][0]); break;
        case P_BHABHA: f[k] = ds_bhabha(E, M.Ee_min, xx[k][0]); break;
        default: f[k] = ds_muone(E, M.Ee_min, xx[k][0]); break;
      }
    }
  } else if (DIM == 3) {
    const double ml = (proc == P_DARKBREM) ? kMe : kMmu;
#pragma unroll
    for (int k = 0; k < T; ++k) f[k] = ds_darkbrem_fast(M, sc, E, ml, xx[k]);      // the two 3-D processes
  } else if (FAM == -1 && DIM == 1 && proc == P_DARKANN) {
#pragma unroll
    for (int k = 0; k < T; ++k) f[k] = ds_darkann_c(M, sc, E, xx[k][0]);
  } else {
#pragma unroll
    for (int k = 0; k < T; ++k) f[k] = dsigma(M, proc, E, xx[k]);
  }
  unsigned am = 0;
#pragma unroll
  for (int k = T - 1; k >= 0; --k) {
    if (fma(maxF, D[k][DIM], -maxF) < (jac[k] * mi.invB) * f[k]) {   // max_F * u
      am |= 1u << k;
      x[0] = xx[k][0]; x[1] = xx[k][1]; x[2] = xx[k][2]; x[3] = xx[k][3];
    }
  }
  return am;
}

// Persistent sampling kernel.  G lanes cooperate on one sample and every lane evaluates T consecutive trials per round:
// lane l of the group evaluates trials (r G + l) T .. (r G + l) T + T - 1 in round r and the lowest accepted trial wins -
// identical to the reference's sequential first-accept rule because every trial's uniforms are a pure function of
// (particle key, trial index).  Groups pull the next sample of the tile from a shared cursor as soon as they finish.
__host__ __device__ constexpr bool proc_is_sm(int p) { return p < P_DARKBREM; }
__host__ __device__ constexpr bool proc_is_4d(int p) { return p == P_BREM || p == P_PAIRPROD || p == P_MUONBREM; }

// family_mode (FAM == -1 only): 1 = leave the dark-brem tiles to the FAM = 3 launch of the same pass (own tile cursor, ctrl[4])
template <int G, int FAM, int T>
__global__ void __launch_bounds__(SAMPLE_THREADS, (FAM == 2) ? (T > 1 ? PB_SAMPLE_MINB_SM_T : PB_SAMPLE_MINB_SM) : (FAM == 3) ? PB_SAMPLE_MINB_DB : (T > 1 ? PB_SAMPLE_MINB_T : PB_SAMPLE_MINB))
k_sample(const __grid_constant__ Material M, const __grid_constant__ Tables T_, SampleIO io, Work W, int family_mode) {
  const Tables& Tb = T_;
  __shared__ __align__(128) double s_grid[GRID_SMEM_DOUBLES];
  __shared__ __align__(16) double s_E[TILE], s_cb[TILE], s_cc[TILE], s_cd[TILE];   // per-entry energy and sampler constants
  __shared__ __align__(16) uint2 s_key[TILE];
  __shared__ int s_idx[TILE];
  __shared__ __align__(8) uint64_t s_bar;
  __shared__ int s_tile, s_cursor;
  const int lane = threadIdx.x & 31;
  const int sub = lane % G;
  const int gbase = lane - sub;
  const unsigned gmask = (G == 32) ? 0xffffffffu : (((1u << G) - 1u) << gbase);
  uint32_t phase = 0;
  unsigned long long c_trials = 0, c_samples = 0, c_fail = 0;
  __shared__ unsigned long long s_ptrials, s_psamples;
  if (threadIdx.x == 0) { mbar_init(&s_bar, 1); s_ptrials = 0; s_psamples = 0; }
  const long long w_begin = io.ws ? io.ws->begin : 0;
  const int w_new = io.ws ? io.ws->n_new : 0x7fffffff, w_par = io.ws ? io.ws->parity : 0;
  auto rec_of = [&](int i) -> size_t { return io.ws ? (size_t)wave_slot(W, w_begin, w_new, w_par, i) : (size_t)i; };
  __syncthreads();
  for (;;) {
    if (threadIdx.x == 0) { s_tile = atomicAdd(&W.ctrl[(FAM == -1 && family_mode == 1) ? 4 : 1], 1); s_cursor = 0; }
    __syncthreads();
    int tile = s_tile;
    if (tile >= W.ctrl[0]) break;
    int bucket = W.tile_bucket[tile], tstart = W.tile_start[tile], tcount = W.tile_count[tile];
    int proc = bucket / LU_MAX, lu = bucket % LU_MAX;
    {
      const bool db = proc == P_DARKBREM || proc == P_DARKMUONBREM;
      if ((FAM == 3 && !db) || (FAM == -1 && family_mode == 1 && db)) { __syncthreads(); continue; }     // the other launch's tile
    }
    const MapInfo& mi = Tb.map[proc];
    const bool tlog = W.tlog != nullptr && threadIdx.x == 0 && io.ws != nullptr && io.ws->waves == W.tlog_wave && tile < W.tlog_cap;
    if (tlog) { unsigned long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); W.tlog[4 * tile] = t; }
    if (threadIdx.x == 0) {
      uint32_t bytes = (uint32_t)mi.stride * 8u;
      mbar_expect_tx(&s_bar, bytes);
      tma_bulk_g2s(s_grid, mi.grid + (size_t)lu * mi.stride, bytes, &s_bar);
    }
    // stage the tile's entries (contiguous in bucket order, see k_bucket_fill) while the node grid is in flight, and
    // compute the per-sample constants of the 4-D integrands here with every lane busy
    for (int k = threadIdx.x; k < tcount; k += SAMPLE_THREADS) {
      double Ek = W.sE[tstart + k];
      s_E[k] = Ek;
      s_key[k] = W.skey[tstart + k];
      s_idx[k] = W.sorted[tstart + k].x;
      if (FAM != 3 && proc_is_4d(proc)) {
        SampleConst c = (proc == P_PAIRPROD) ? pairprod_const(M, Ek) : brem_const(M, Ek, proc == P_BREM ? kMe : kMmu);
        s_cb[k] = c.b; s_cc[k] = c.c; s_cd[k] = c.d;
      } else if (FAM != 2 && (proc == P_DARKBREM || proc == P_DARKMUONBREM)) {
        SampleConst c = darkbrem_const(M, Ek, proc == P_DARKBREM ? kMe : kMmu);
        s_cb[k] = c.b; s_cc[k] = c.c; s_cd[k] = c.e;          // |p|, tconv, 1/|p|
      } else if (FAM == -1 && proc == P_DARKANN) {
        SampleConst c = darkann_const(M, Ek);
        s_cb[k] = c.b; s_cc[k] = c.c; s_cd[k] = c.d;          // beta, u_max, 2 x prefactor
      }
    }
    double maxF = __ldg(mi.maxF + lu) * M.fudge;
    mbar_wait(&s_bar, phase);
    phase ^= 1;
    __syncthreads();
    const uint32_t max_trials = (uint32_t)min(M.max_trials, 0xffffff00LL);
    // group state
    int cur = -1;          // entry index, -1 = need a new one, -2 = tile exhausted
    double E = 0.0; uint2 key = make_uint2(0, 0);
    SampleConst sc{0, 0, 0, 0, 0};
    uint32_t next_t = 0;   // first trial index this sample has not evaluated yet
    for (;;) {
      if (cur == -1) {
        int j = 0;
        if (sub == 0) j = atomicAdd(&s_cursor, 1);
        if (G > 1) j = __shfl_sync(gmask, j, gbase);
        if (j < tcount) {
          cur = s_idx[j];
          E = s_E[j];
          key = s_key[j];
          next_t = 0;
          if (FAM == 3) sc = SampleConst{0.0, s_cb[j], s_cc[j], M.i2mT, s_cd[j]};
          else if (proc == P_PAIRPROD) sc = SampleConst{E - 2 * kMe, s_cb[j], s_cc[j], s_cd[j], 0.0};
          else if (proc == P_BREM) sc = SampleConst{E - kMe - M.Eg_min, s_cb[j], s_cc[j], s_cd[j], kMe * kMe};
          else if (proc == P_MUONBREM) sc = SampleConst{E - kMmu - M.Eg_min, s_cb[j], s_cc[j], s_cd[j], kMmu * kMmu};
          else if (FAM != 2 && (proc == P_DARKBREM || proc == P_DARKMUONBREM)) sc = SampleConst{0.0, s_cb[j], s_cc[j], M.i2mT, s_cd[j]};
          else if (FAM == -1 && proc == P_DARKANN) sc = SampleConst{0.0, s_cb[j], s_cc[j], s_cd[j], (M.mV * M.mV) / (2.0 * kMe * (E + kMe))};
        } else cur = -2;
      }
      // ---- drain help.  Once the tile's cursor is exhausted a group that finishes has nothing left to fetch; instead of
      // idling until the slowest sample of the warp is accepted (the geometric tail: 1 sample in 200 needs > 100 trials) it
      // joins a sample that is still open in its warp.  The open samples ("owners", k of them) share the idle groups evenly:
      // idle group i helps owner i mod k as that sample's (1 + i / k)-th group, with the owner's state fetched by shuffles
      // (the helper's own registers are dead), and the m groups on a sample cover trials next_t .. next_t + m G T - 1 of the
      // round.  Trial indices, not lanes, define the draws, so the first accepted INDEX wins whoever evaluated it.
      const unsigned own = __ballot_sync(0xffffffffu, sub == 0 && cur >= 0);
      const unsigned idl = __ballot_sync(0xffffffffu, sub == 0 && cur == -2);
      if (own == 0) break;                                // every group is out of work: the tile is done
      const bool help = (G < 32) && idl != 0;             // warp-uniform
      bool helper = false;
      uint32_t r = 0, m = 1;                              // my group's rank among the groups on my sample, and their number
      unsigned peers = gmask;
      if (help) {
        const unsigned below = (1u << gbase) - 1u;
        const int k = __popc(own), n_idle = __popc(idl);
        int j;                                            // which owner (in lane order) my group works for
        if (cur == -2) { const int ir = __popc(idl & below); j = ir % k; r = 1 + ir / k; helper = true; }
        else j = __popc(own & below);
        m = 1 + (n_idle - j + k - 1) / k;
        const int src = __fns(own, 0, j + 1);             // leader lane of that owner
        const double Eh = __shfl_sync(0xffffffffu, E, src);
        const unsigned kx = __shfl_sync(0xffffffffu, key.x, src), ky = __shfl_sync(0xffffffffu, key.y, src);
        const double sa = __shfl_sync(0xffffffffu, sc.a, src), sb = __shfl_sync(0xffffffffu, sc.b, src), sc_ = __shfl_sync(0xffffffffu, sc.c, src),
                     sd = __shfl_sync(0xffffffffu, sc.d, src), se = __shfl_sync(0xffffffffu, sc.e, src);
        const int ch = __shfl_sync(0xffffffffu, cur, src);
        const uint32_t nh = __shfl_sync(0xffffffffu, next_t, src);
        if (helper) { E = Eh; key = make_uint2(kx, ky); sc = SampleConst{sa, sb, sc_, sd, se}; cur = ch; next_t = nh; }
        peers = __match_any_sync(0xffffffffu, cur);
      }
      double x[4] = {0.0, 0.0, 0.0, 0.0};
      unsigned am = 0;
      const uint32_t code0 = (r * G + sub) * T;           // my first trial of this round, relative to next_t
      const uint32_t t0 = next_t + code0;
      if (cur >= 0 && t0 < max_trials) {
        if (FAM == 2) {
          if (proc_is_4d(proc)) am = trial_block<4, 2, T>(M, mi, s_grid, proc, E, sc, maxF, key, t0, x);
          else am = trial_block<1, 2, T>(M, mi, s_grid, proc, E, sc, maxF, key, t0, x);
        } else if (FAM == 3) {
          am = trial_block<3, 3, T>(M, mi, s_grid, proc, E, sc, maxF, key, t0, x);
        } else switch (mi.dim) {
          case 4: am = trial_block<4, -1, T>(M, mi, s_grid, proc, E, sc, maxF, key, t0, x); break;
          case 3: am = trial_block<3, -1, T>(M, mi, s_grid, proc, E, sc, maxF, key, t0, x); break;
          default: am = trial_block<1, -1, T>(M, mi, s_grid, proc, E, sc, maxF, key, t0, x); break;
        }
        if (T > 1 && max_trials - t0 < (uint32_t)T) am &= (1u << (max_trials - t0)) - 1u;      // trials at or beyond max_trials do not exist
      }
      const uint32_t mine = am ? code0 + (uint32_t)(__ffs(am) - 1) : 0xffffffffu;
      uint32_t best = mine;                               // lowest accepted trial of my sample in this round
      if (help) best = __reduce_min_sync(peers, mine);
      else if (G > 1) {
#pragma unroll
        for (int o = G / 2; o > 0; o >>= 1) best = min(best, __shfl_xor_sync(0xffffffffu, best, o));
      }
      if (cur >= 0) {
        if (best != 0xffffffffu) {
          if (mine == best) {
            double2* xo = reinterpret_cast<double2*>(W.xs + 4 * (size_t)cur);
            xo[0] = make_double2(x[0], x[1]); xo[1] = make_double2(x[2], x[3]);
            int ntr = (int)(next_t + best + 1);
            io.ntr[rec_of(cur) * io.ntr_stride] = ntr;
            c_trials += ntr; c_samples += 1;
          }
          cur = -1;
        } else {
          next_t += m * (G * T);
          if (next_t >= max_trials) {                      // "No Sample Found" (shower.py:460-461)
            if (sub == 0 && !helper) {
              io.ntr[rec_of(cur) * io.ntr_stride] = -1;
              W.bucket[cur] = P_NONE * LU_MAX;
              W.xs[4 * (size_t)cur] = __longlong_as_double(0x7ff8000000000000LL);     // NaN sample: the emit kernels skip this entry
              c_trials += (unsigned long long)max_trials; c_fail += 1;
            }
            cur = -1;
          }
        }
      }
      if (helper) cur = -2;
    }
    // per-tile counter flush (a tile is one process): warp reduce -> shared -> one global atomic per CTA
    for (int o = 16; o > 0; o >>= 1) {
      c_trials += __shfl_down_sync(0xffffffffu, c_trials, o);
      c_samples += __shfl_down_sync(0xffffffffu, c_samples, o);
      c_fail += __shfl_down_sync(0xffffffffu, c_fail, o);
    }
    if (lane == 0) {
      if (c_trials) atomicAdd(&s_ptrials, c_trials);
      if (c_samples) atomicAdd(&s_psamples, c_samples);
      if (c_fail) atomicAdd(&W.counters[CNT_NOSAMPLE], c_fail);
    }
    c_trials = 0; c_samples = 0; c_fail = 0;
    __syncthreads();   // everyone is done with s_grid (and the counters) before the next tile overwrites it
    if (tlog) {
      unsigned long long t; unsigned sm;
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
      asm volatile("mov.u32 %0, %%smid;" : "=r"(sm));
      W.tlog[4 * tile + 1] = t; W.tlog[4 * tile + 2] = ((unsigned long long)bucket << 32) | (unsigned)tcount; W.tlog[4 * tile + 3] = ((unsigned long long)s_ptrials << 16) | sm;
    }
    if (threadIdx.x == 0) {
      if (s_ptrials) { atomicAdd(&W.counters[CNT_TRIALS], s_ptrials); atomicAdd(&W.counters[CNT_PROC_TRIALS + proc], s_ptrials); atomicAdd(&W.bstat[2 * bucket + 1], s_ptrials); }
      if (s_psamples) { atomicAdd(&W.counters[CNT_SAMPLES], s_psamples); atomicAdd(&W.counters[CNT_PROC_SAMPLES + proc], s_psamples); atomicAdd(&W.bstat[2 * bucket], s_psamples); }
      s_ptrials = 0; s_psamples = 0;
    }
  }
}

// Hard scatter or decay of ONE particle: sampled point -> two four-vectors in the parent frame (kinematics.py) -> lab frame
// (particle.py:176-185, shower.py:471-480), PDG ids of the two products (shower.py:77-96) and the weight factor (decays).
// Decay in flight of a long-lived meson, pi+- / K+- -> mu nu (particle.py:363-389 draw_x_sample / prob_decay_b_int, :410-422): the
// distance to the decay point is drawn by accept/reject from tot_rate exp(-tot_rate x) on [0, 4 / tot_rate], and EACH daughter's
// weight carries BR x P(decay before interaction) evaluated at its own, independently drawn x (the reference calls
// prob_decay_b_int once per daughter dictionary).  One accept/reject loop: `loop` = 1 (path), 2, 3 (the two weights).
template <class DS>
__device__ __forceinline__ double decay_draw_x(DS& ds, uint32_t loop, double tot_rate) {
  const double x_max = 4.0 / (tot_rate * 1.0);
  double x = 0.0;
  for (uint32_t i = 0; i < (1u << 20); ++i) {
    D2 u = ds.decay_x(loop, i);
    x = DS::kOrderFree ? x_max * u.a : u.a;                     // np.random.uniform(0, x_max); a tape holds the generator's output itself
    if (u.b < (tot_rate * exp(-tot_rate * x)) / tot_rate) break;
  }
  return x;
}
// The three accept/reject loops of one decay in flight.  Only primaries can be long-lived (daughters are created 'stable'), so the
// engine runs this once per primary in k_init_primaries and k_emit only reads the three numbers (PRE = true): the loops and their
// exponentials stay out of k_emit's register budget.  The tape-driven replay (k_replay) draws in line, in the reference's order.
template <class DS>
__device__ __forceinline__ void decay_in_flight_draws(DS& ds, int pid, double E, double mass, double* dz, double* wfac, double* wfac_b) {
  const bool kaon = pid == 321 || pid == -321;
  const double int_len = kaon ? 2.2875e-1 : 1.796e-1, ctau0 = kaon ? 3.711 : 7.8045, br = kaon ? 0.6356 : 0.9998;     // particle.py:15-20, 44-47
  const double gamma = E / mass;                                // the energy at creation (particle.py:358, 381)
  const double beta = sqrt(1 - 1 / (gamma * gamma));
  const double ctau = ctau0 * gamma * beta;
  const double tot_rate = 1.0 / ctau + 1.0 / int_len;
  *dz = decay_draw_x(ds, 1, tot_rate);
  const double den = 1 + ctau / int_len;
  *wfac = br * ((1 - exp(-tot_rate * decay_draw_x(ds, 2, tot_rate))) / den);
  *wfac_b = br * ((1 - exp(-tot_rate * decay_draw_x(ds, 3, tot_rate))) / den);
}
template <bool PRE, class DS>
__device__ __forceinline__ void scatter_products(int proc, int pid, int flags, V4 pf, double mass, const double* x, DS& ds,
                                                 V4& da, V4& db, int& pid_a, int& pid_b, double& wfac, double& wfac_b, double& dz) {
  wfac = 1.0; wfac_b = 1.0; dz = 0.0;
  if (proc == P_SMDECAY) {
    double m1 = 0.0;
    if (flags & PB_FLAG_LONG_LIVED) {                          // pi+- / K+- -> mu nu in flight (particle.py:410-422)
      pid_a = pid > 0 ? -13 : 13; pid_b = pid > 0 ? 14 : -14;
      m1 = kMmu;
      if (!PRE) decay_in_flight_draws(ds, pid, pf.E, mass, &dz, &wfac, &wfac_b);      // PRE: the caller applies the numbers k_init_primaries drew
    } else {                                                   // pi0 -> gamma gamma (particle.py:391-409)
      pid_a = 22; pid_b = 22;
      wfac = wfac_b = 0.98823;                                 // particle.py:40 meson_decay_dict[111]
    }
    D2 u = ds.decay_u(P_SMDECAY);
    if (!DS::kOrderFree) u.a = 0.5 * (u.a + 1.0);               // a tape holds cos(theta) = np.random.uniform(-1, 1) as drawn (particle.py:226)
    two_body_decay(pf, mass, m1, 0.0, u.a, u.b, &da, &db);
    return;
  }
  double u_az = ds.kin_u(proc);
  double E0 = pf.E;
  switch (proc) {                                              // shower.py:77-96
    case P_BREM: case P_MUONBREM: kin_brem(E0, mass, x, u_az, &da, &db); pid_a = pid; pid_b = 22; break;
    case P_PAIRPROD: kin_pairprod(E0, x, u_az, &da, &db); pid_a = -11; pid_b = 11; break;
    case P_COMP: kin_compton(E0, 0.0, x[0], u_az, &da, &db); pid_a = 11; pid_b = 22; break;
    case P_ANN: kin_annihilation(E0, 0.0, x[0], u_az, &da, &db); pid_a = 22; pid_b = 22; break;
    case P_MOLLER: case P_BHABHA: kin_ee(E0, x[0], u_az, &da, &db); pid_a = pid; pid_b = 11; break;
    case P_MUONE: kin_mue(E0, x[0], u_az, &da, &db); pid_a = pid; pid_b = 11; break;
  }
  Rot R = rotation_to(pf);                                     // shower.py:471,479-480
  da = rotate(R, da); db = rotate(R, db);
}

// Kinematics + rotation + daughter append, in bucket order (warps are process-coherent).
__global__ void PB_EMIT_BOUNDS
k_emit(const __grid_constant__ Material M, const __grid_constant__ Tables T, Stack S, Work W, int wave_order,
       const double* __restrict__ decay_aux) {
  const long long begin = W.ws->begin;
  const int n = W.ws->n, n_new = W.ws->n_new, parity = W.ws->parity;
  int* __restrict__ next_c = W.list[2 * (parity ^ 1)];
  int* __restrict__ next_n = W.list[2 * (parity ^ 1) + 1];
  const int lane = threadIdx.x & 31;
  // CTA-uniform grid-stride loop (the append below uses full-warp shuffles and two CTA barriers)
  __shared__ int s_tot[4], s_tch[4];
  __shared__ unsigned long long s_base[4], s_lbase[4];
  for (int cbase = blockIdx.x * blockDim.x; cbase < n; cbase += gridDim.x * blockDim.x) {
  const int j = cbase + (int)threadIdx.x;
  V4 da{0, 0, 0, 0}, db{0, 0, 0, 0};
  int pid_a = 0, pid_b = 0, proc = P_NONE;
  bool keep_a = false, keep_b = false;
  long long slot = 0;
  int4 meta = make_int4(0, 0, 0, 0);
  uint2 key = make_uint2(0, 0);
  double wgt = 0.0, rx = 0, ry = 0, rz = 0;
  if (j < n) {
    int i, bucket;
    if (wave_order) { i = j; bucket = W.bucket[i]; }
    else { int2 ib = W.sorted[j]; i = ib.x; bucket = ib.y; }
    proc = bucket / LU_MAX;
    double2 a0 = make_double2(0.0, 0.0), a1 = a0, b0 = a0, b1 = a0, x01 = a0, x23 = a0;
    if (proc != P_NONE) {
      // every load of this particle is issued here, behind the one coalesced (index, bucket) read: two levels of memory latency
      slot = wave_slot(W, begin, n_new, parity, i);
      const double2* pfp = reinterpret_cast<const double2*>(S.pf + 4 * slot);
      const double2* rfp = reinterpret_cast<const double2*>(S.rf + 4 * slot);
      const double2* xp = reinterpret_cast<const double2*>(W.xs + 4 * (size_t)i);
      a0 = pfp[0]; a1 = pfp[1]; b0 = rfp[0]; b1 = rfp[1];
      if (proc < N_SAMPLED) { x01 = xp[0]; x23 = xp[1]; }
      meta = ld_meta(S, slot);
      { int4 kw = ld_kw(S, slot); key = kw_key(kw); wgt = kw_weight(kw); }
      if (x01.x != x01.x) {                                          // the sampler gave up ("No Sample Found", shower.py:460-461)
        S.ids[2 * slot].z = meta.z | (PB_FLAG_NO_SAMPLE << 8);
        proc = P_NONE;
      }
    }
    if (proc != P_NONE) {
      V4 pf{a0.x, a0.y, a1.x, a1.y};
      rx = b0.x; ry = b0.y; rz = b1.x;
      double mass = b1.y;
      int pid = meta.x;
      double x[4] = {x01.x, x01.y, x23.x, x23.y};
      PhiloxDraws ds{key};
      double wfac, wfac_b, dz;
      scatter_products<true>(proc, pid, (meta.z >> 8) & 0x7f, pf, mass, x, ds, da, db, pid_a, pid_b, wfac, wfac_b, dz);
      wgt *= wfac;
      keep_a = da.E > M.min_energy;                                // shower.py:704-706
      keep_b = db.E > M.min_energy;
    }
  }
  // CTA-aggregated append: one atomic on the stack tail and one on the packed (neutral, charged) list counters per CTA (every warp of
  // the launch used to bump the same two words and wait for the result: 12 % of the kernel's stall samples)
  const bool ch_a = keep_a && is_charged(pid_a), ch_b = keep_b && is_charged(pid_b);
  int cnt = (keep_a ? 1 : 0) + (keep_b ? 1 : 0);
  int cch = (ch_a ? 1 : 0) + (ch_b ? 1 : 0);
  int incl = cnt, inch = cch;
  for (int o = 1; o < 32; o <<= 1) {
    int v = __shfl_up_sync(0xffffffffu, incl, o), w = __shfl_up_sync(0xffffffffu, inch, o);
    if (lane >= o) { incl += v; inch += w; }
  }
  int total = __shfl_sync(0xffffffffu, incl, 31), total_ch = __shfl_sync(0xffffffffu, inch, 31);
  const int wid = threadIdx.x >> 5;
  if (lane == 0) { s_tot[wid] = total; s_tch[wid] = total_ch; }
  __syncthreads();
  if (threadIdx.x == 0) {
    int t_all = 0, c_all = 0;
    for (int k = 0; k < 4; ++k) { t_all += s_tot[k]; c_all += s_tch[k]; }
    unsigned long long b0 = 0, l0 = 0;
    if (t_all > 0) {
      b0 = atomicAdd(&W.tail[0], (unsigned long long)t_all);
      l0 = atomicAdd(&W.tail[1], ((unsigned long long)(t_all - c_all) << 32) | (unsigned long long)c_all);
    }
    for (int k = 0; k < 4; ++k) {        // per-warp bases: records, and the (neutral << 32 | charged) list positions
      s_base[k] = b0; s_lbase[k] = l0;
      b0 += (unsigned long long)s_tot[k];
      l0 += ((unsigned long long)(s_tot[k] - s_tch[k]) << 32) | (unsigned long long)s_tch[k];
    }
  }
  __syncthreads();
  const unsigned long long base = s_base[wid], lbase = s_lbase[wid];
  // decay in flight of a long-lived primary (pi+-, K+-: particle.py:410-422): path and weight factors were drawn by k_init_primaries
  // (decay_aux: dz, factor of the first daughter, of the second).  The parent ends where it decays, the daughters start there.
  // Read here, late, so that the rare branch holds no registers across the kinematics above.
  const bool in_flight = proc == P_SMDECAY && ((meta.z >> 8) & PB_FLAG_LONG_LIVED);
  if (in_flight) {
    const double dz = decay_aux[3 * slot];
    const double2* pfp = reinterpret_cast<const double2*>(S.pf + 4 * slot);
    const double2 q0 = pfp[0], q1 = pfp[1];
    const double pn = sqrt(q0.y * q0.y + q1.x * q1.x + q1.y * q1.y);
    if (pn > 0.0) {
      rx += q0.y / pn * dz; ry += q1.x / pn * dz; rz += q1.y / pn * dz;
      double2* rfo = reinterpret_cast<double2*>(S.rf + 4 * slot);
      const double m = rfo[1].y;
      rfo[0] = make_double2(rx, ry); rfo[1] = make_double2(rz, m);
    }
  }
  if (cnt) {
    long long dst = (long long)base + (incl - cnt);
    int ci = (int)(lbase & 0xffffffffu) + (inch - cch);
    int ni = (int)(lbase >> 32) + ((incl - cnt) - (inch - cch));
    const long long next_begin = begin + n_new;
    int gen = ((meta.z >> 16) & 0xffff) + 1;
    for (int bit = 0; bit < 2; ++bit) {
      bool keep = bit ? keep_b : keep_a;
      if (!keep) continue;
      if (dst >= S.capacity) { atomicAdd(&W.counters[CNT_OVERFLOW], 1ull); break; }
      V4 d = bit ? db : da;
      double2* p0p = reinterpret_cast<double2*>(S.p0 + 4 * dst);
      double2* r0p = reinterpret_cast<double2*>(S.r0w + 4 * dst);
      p0p[0] = make_double2(d.E, d.x); p0p[1] = make_double2(d.y, d.z);
      const double wd = in_flight ? wgt * decay_aux[3 * slot + 1 + bit] : wgt;      // each daughter of a decay in flight has its own factor
      r0p[0] = make_double2(rx, ry);   r0p[1] = make_double2(rz, wd);
      st_ids(S, dst, make_int4(bit ? pid_b : pid_a, (int)slot, pack_info(gen, bit, 0, proc), meta.w), child_key(key, bit), wd);
      if (bit ? ch_b : ch_a) {
        next_c[ci++] = (int)(dst - next_begin);
        store_track_setup(M, T, S, dst, bit ? pid_b : pid_a, pid_mass(bit ? pid_b : pid_a), d.E, d.x, d.y, d.z);
      } else next_n[ni++] = (int)(dst - next_begin);
      ++dst;
    }
  }
  }   // grid-stride loop
}

__global__ void k_init_primaries(const __grid_constant__ Material M, const __grid_constant__ Tables T, Stack S, Work W,
                                 const double* __restrict__ p, const double* __restrict__ r, const double* __restrict__ w,
                                 const double* __restrict__ mass, const int* __restrict__ pid, const int* __restrict__ flags,
                                 long long n, unsigned long long seed, unsigned long long first_id, double* __restrict__ decay_aux) {
  long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (i >= n) return;
  if (flags[i] & PB_FLAG_LONG_LIVED) {                     // decay in flight: the draws k_emit will apply (scatter_products<true>)
    PhiloxDraws ds{root_key(seed, first_id + (unsigned long long)i)};
    decay_in_flight_draws(ds, pid[i], p[4 * i], mass[i], decay_aux + 3 * i, decay_aux + 3 * i + 1, decay_aux + 3 * i + 2);
  }
  for (int k = 0; k < 4; ++k) S.p0[4 * i + k] = p[4 * i + k];
  for (int k = 0; k < 3; ++k) S.r0w[4 * i + k] = r[3 * i + k];
  S.r0w[4 * i + 3] = w[i];
  st_ids(S, i, make_int4(pid[i], -1, pack_info(0, 0, flags[i], P_INPUT), (int)i), root_key(seed, first_id + (unsigned long long)i), w[i]);
  const bool ch = is_charged(pid[i]) && !(flags[i] & (PB_FLAG_SHORT_LIVED | PB_FLAG_LONG_LIVED));
  unsigned long long old = atomicAdd(&W.tail[1], ch ? 1ull : (1ull << 32));
  if (ch) {
    W.list[2][(int)(old & 0xffffffffu)] = (int)i;          // parity 1: the first k_wave_begin flips 0 -> 1
    store_track_setup(M, T, S, i, pid[i], mass[i], p[4 * i], p[4 * i + 1], p[4 * i + 2], p[4 * i + 3]);
  } else W.list[3][(int)(old >> 32)] = (int)i;
}


// ------------------------------------------------------------------------------------------ dark pass
// log-log table of dark_shower.py:31-46 (interpolate1d): 10 ** lin(log10 E), 1e-20 outside the grid
__device__ __forceinline__ double loglog_eval(const NSigmaTable& T, double E) {
  double lx = log10(E);
  if (T.n < 2 || !(lx >= T.xmin && lx <= T.xmax)) return pow(10.0, -20.0);
  return pow(10.0, nsigma_at(T, nsigma_locate(T, lx), lx));
}

// np.sum over 10 doubles (pairwise summation kernel: 8 accumulators, then the remainder)
__device__ __forceinline__ double np_sum10(const double* a) {
  double r = __dadd_rn(__dadd_rn(__dadd_rn(a[0], a[1]), __dadd_rn(a[2], a[3])), __dadd_rn(__dadd_rn(a[4], a[5]), __dadd_rn(a[6], a[7])));
  r = __dadd_rn(r, a[8]);
  return __dadd_rn(r, a[9]);
}

// GetBSMWeights (dark_shower.py:595-647) + the pre-sampling half of produce_bsm_particle (:721-768): interaction-energy
// bin, MCS + energy loss down to it, threshold short-cuts -> candidate (SM slot, process, weight, four-momentum, bucket).
__global__ void __launch_bounds__(128)
k_dark_prepare(const __grid_constant__ Material M, const __grid_constant__ Tables T, const __grid_constant__ DarkTables D,
               Stack S, Work W, DarkCand C, long long first, long long n, unsigned active) {
  long long s = first + blockIdx.x * (long long)blockDim.x + threadIdx.x;      // SM records [first, n) of this chunk
  const int lane = threadIdx.x & 31;
  int nc = 0;
  int c_proc[2]; double c_wg[2]; V4 c_pf[2]; int c_bucket[2];
  if (s < n) {
    int4 meta = ld_meta(S, s);
    int pid = meta.x;
    const double2* p0p = reinterpret_cast<const double2*>(S.p0 + 4 * s);
    double2 a0 = p0p[0], a1 = p0p[1];
    V4 p0{a0.x, a0.y, a1.x, a1.y};
    double E0 = p0.E;
    double mass = S.rf[4 * s + 3];
    uint2 key = kw_key(ld_kw(S, s));
    const double pre = M.g_e * M.g_e / (4 * kPi * kAlpha);
    int procs[2] = {-1, -1};
    if (pid == 11) procs[0] = P_DARKBREM;
    else if (pid == -11) { procs[0] = P_DARKBREM; procs[1] = P_DARKANN; }
    else if (pid == 22) procs[0] = P_DARKCOMP;
    else if (pid == 13 || pid == -13) procs[0] = P_DARKMUONBREM;
    else if (pid == 111 || pid == 221 || pid == 331) procs[0] = P_BSMDECAY;        // pi0, eta, eta' -> gamma V (dark_shower.py:633-638)
    // the reference visits active_processes in list order; host code restores that order, here DarkBrem < DarkAnn
    for (int k = 0; k < 2; ++k) {
      int proc = procs[k];
      if (proc < 0 || !((active >> proc) & 1u)) continue;
      double wg = 0.0;
      int wt = -1;                                                      // weight / drate table
      if (proc == P_BSMDECAY) {
        double r = M.mV / mass;
        const double br = pid == 111 ? 0.98823 : (pid == 221 ? 0.3936 : 0.02307);      // particle.py:40-48 meson_twobody_branchingratios
        if (!(r >= 1.0)) { double q = 1.0 - r * r; wg = 2 * M.eps * M.eps * (q * q * q) * br; }
      } else {
        double thr = D.min_E[proc - P_DARKBREM];
        if (E0 < thr) continue;
        if (proc == P_DARKCOMP) {
          if (E0 < M.min_calc[2]) continue;
          wg = pre * loglog_eval(D.nsdark_comp, E0) / (nsigma_c(T.ns[P_PAIRPROD], E0) + nsigma_c(T.ns[P_COMP], E0));
        } else {
          wt = (proc == P_DARKBREM) ? (pid == 11 ? 0 : 1) : (proc == P_DARKANN ? 2 : 3);
          wg = pre * nsigma_eval(D.w[wt], E0);
        }
      }
      if (!(wg > 0.0)) continue;
      const double2* pfp = reinterpret_cast<const double2*>(S.pf + 4 * s);
      double2 f0 = pfp[0], f1 = pfp[1];
      V4 pf{f0.x, f0.y, f1.x, f1.y};
      int bucket;
      if (proc == P_BSMDECAY) {
        bucket = P_BSMDECAY * LU_MAX;
      } else {
        if (wt >= 0) {                                                  // dark_shower.py:738-759
          const double* Es = D.dE[wt];
          int nE = D.dn[wt];
          int ie;
          if (E0 < __ldg(Es)) ie = 0;
          else if (E0 > __ldg(Es + nE - 1)) ie = nE - 1;
          else { int lo = 0, hi = nE; while (lo < hi) { int mid = (lo + hi) >> 1; if (__ldg(Es + mid) <= E0) lo = mid + 1; else hi = mid; } ie = lo - 1; }
          double Ei = __ldg(Es + ie);
          const double* tab = D.dT[wt] + (size_t)ie * 20;
          double rate[10];
          for (int b = 0; b < 10; ++b) rate[b] = __ldg(tab + 2 * b + 1);
          double tot = np_sum10(rate);
          if (tot == 0.0) continue;                                     // "return None"
          for (int b = 0; b < 10; ++b) rate[b] = rate[b] / tot;
          double norm = np_sum10(rate);
          double cdf[10], acc = 0.0;
          for (int b = 0; b < 10; ++b) { acc = __dadd_rn(acc, rate[b] / norm); cdf[b] = acc; }
          double u = draw2(key, 0, ST_DBIN, 0, proc).a;
          int pick = 9;
          for (int b = 9; b >= 0; --b) if (u < cdf[b] / acc) pick = b;
          double E_int = __ldg(tab + 2 * pick) + (E0 - Ei);
          double dist = (E0 - E_int) / M.dEdx;
          pf = p0;
          double pn = norm3_nofma(pf.x, pf.y, pf.z);
          if (pn > 0) {                                                 // SURVEY Q-12: default (electron) m_lepton
            McsDraw d = mcs_draw(key, MCS_FINAL_INDEX, proc);
            pf = mcs_scatter(M, pf, pn, M.rho * (dist / kCmToM), kMe, mass, d);
          }
          pf = lose_energy(pf, mass, E0 - E_int);
        }
        double E = pf.E;
        if ((proc == P_DARKANN && E <= M.E_res_ann) || (proc == P_DARKCOMP && E <= M.E_thr_comp)) bucket = P_BSMDECAY * LU_MAX + 1;
        else bucket = proc * LU_MAX + lookup_row(T.map[proc], E);
      }
      c_proc[nc] = proc; c_wg[nc] = wg; c_pf[nc] = pf; c_bucket[nc] = bucket;
      ++nc;
    }
  }
  int incl = nc;
  for (int o = 1; o < 32; o <<= 1) { int v = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += v; }
  int total = __shfl_sync(0xffffffffu, incl, 31);
  int base = 0;
  if (lane == 31 && total > 0) base = atomicAdd(C.count, total);
  base = __shfl_sync(0xffffffffu, base, 31);
  for (int k = 0; k < nc; ++k) {
    int i = base + (incl - nc) + k;
    C.slot[i] = (int)s; C.proc[i] = c_proc[k]; C.wg[i] = c_wg[k]; C.ntr[i] = 0;
    double2* o = reinterpret_cast<double2*>(C.pf + 4 * (size_t)i);
    o[0] = make_double2(c_pf[k].E, c_pf[k].x); o[1] = make_double2(c_pf[k].y, c_pf[k].z);
    W.bucket[i] = c_bucket[k];
    atomicAdd(&W.hist[c_bucket[k]], 1);
  }
}

// Second half of produce_bsm_particle (dark_shower.py:761-804) and the TwoBody_BSMDecay branch (:836-844): dark-vector
// four-momentum in the parent frame, rotation to the lab, record appended to the dark stack.
__global__ void __launch_bounds__(128)
k_dark_emit(const __grid_constant__ Material M, Stack S, Stack O, Work W, DarkCand C, int n_cand) {
  int j = blockIdx.x * blockDim.x + threadIdx.x;
  const int lane = threadIdx.x & 31;
  bool keep = false;
  V4 v{0, 0, 0, 0};
  double wgt = 0.0, rx = 0, ry = 0, rz = 0;
  int slot = 0, proc = P_NONE, ntr = 0;
  int4 meta = make_int4(0, 0, 0, 0);
  uint2 key = make_uint2(0, 0);
  if (j < n_cand) {
    const int2 ib = W.sorted[j];
    const int i = ib.x;
    int bucket = ib.y;
    if (bucket / LU_MAX < N_SAMPLED && C.ntr[i] < 0) bucket = P_NONE * LU_MAX;      // the sampler gave up on this candidate
    if (bucket / LU_MAX != P_NONE) {
      slot = C.slot[i]; proc = C.proc[i]; ntr = C.ntr[i];
      double wg = C.wg[i];
      const double2* pfp = reinterpret_cast<const double2*>(C.pf + 4 * (size_t)i);
      double2 f0 = pfp[0], f1 = pfp[1];
      V4 pf{f0.x, f0.y, f1.x, f1.y};
      const double2* rfp = reinterpret_cast<const double2*>(S.rf + 4 * (size_t)slot);
      double2 b0 = rfp[0], b1 = rfp[1];
      rx = b0.x; ry = b0.y; rz = b1.x;
      double mass = b1.y;
      meta = ld_meta(S, slot);
      int4 kw = ld_kw(S, slot);
      key = kw_key(kw);
      double w0 = kw_weight(kw);
      const double mV = M.mV, me = kMe;
      double E = pf.E;
      if (proc == P_BSMDECAY) {
        D2 u = draw2(key, 0, ST_DECAY, 0, P_BSMDECAY);
        V4 g;
        two_body_decay(pf, mass, 0.0, mV, u.a, u.b, &g, &v);
      } else {
        if (bucket == P_BSMDECAY * LU_MAX + 1) {
          if (proc == P_DARKANN) v = V4{sqrt(E * E - me * me + mV * mV), 0.0, 0.0, sqrt(E * E - me * me)};
          else v = V4{sqrt(E * E + mV * mV), 0.0, 0.0, E};
        } else {
          const double2* xp = reinterpret_cast<const double2*>(W.xs + 4 * (size_t)i);
          double2 x01 = xp[0], x23 = xp[1];
          double x[4] = {x01.x, x01.y, x23.x, x23.y};
          if (proc == P_DARKCOMP) {                                      // bound electron (dark_shower.py:774-784)
            double c = electron_wave_function(M.Zeff, kAlpha * M.Zeff * me / sqrt(3.0));
            double pe = 0.0;
            for (uint32_t it = 0;; ++it) {                               // draw_pe_sample (dark_shower.py:710-719)
              D2 u = draw2(key, it, ST_PE, 0, proc);
              pe = 1e-3 * u.a;
              if (u.b < electron_wave_function(M.Zeff, pe) / c) break;
            }
            double c0 = -1.0 + 2.0 * draw2(key, 0, ST_C0, 0, proc).a;
            double ss = me * me + 2 * E * (sqrt(me * me + pe * pe) - c0 * pe);
            double Ee = (ss - mV * mV + me * me) / (2 * sqrt(ss));
            if (Ee < me) { v = V4{mV, 0.0, 0.0, 0.0}; wg = 0.0; }
            else { D2 u = draw2(key, 0, ST_KIN, 0, proc); v = kin_compton_bound_V(E, mV, x[0], pe, c0, u.a, u.b); }
          } else if (proc == P_DARKANN) {
            v = kin_darkann_V(E, mV, x[0]);
          } else {
            v = kin_darkbrem_V(E, mV, x, draw2(key, 0, ST_KIN, 0, proc).a);
          }
        }
        Rot R = rotation_to(pf);
        v = rotate(R, v);
      }
      wgt = wg * w0;
      keep = true;
    }
  }
  unsigned ball = __ballot_sync(0xffffffffu, keep);
  int total = __popc(ball);
  unsigned long long base = 0;
  if (lane == 0 && total > 0) base = atomicAdd(&W.tail[0], (unsigned long long)total);
  base = __shfl_sync(0xffffffffu, base, 0);
  if (keep) {
    long long dst = (long long)base + __popc(ball & ((1u << lane) - 1u));
    if (dst >= O.capacity) { atomicAdd(&W.counters[CNT_OVERFLOW], 1ull); return; }
    double2* p0p = reinterpret_cast<double2*>(O.p0 + 4 * dst);
    double2* r0p = reinterpret_cast<double2*>(O.r0w + 4 * dst);
    double2* pfp = reinterpret_cast<double2*>(O.pf + 4 * dst);
    double2* rfp = reinterpret_cast<double2*>(O.rf + 4 * dst);
    p0p[0] = make_double2(v.E, v.x); p0p[1] = make_double2(v.y, v.z);
    pfp[0] = make_double2(v.E, v.x); pfp[1] = make_double2(v.y, v.z);
    r0p[0] = make_double2(rx, ry);   r0p[1] = make_double2(rz, wgt);
    rfp[0] = make_double2(rx, ry);   rfp[1] = make_double2(rz, M.mV);
    int gen = ((meta.z >> 16) & 0xffff) + 1;
    st_ids(O, dst, make_int4(4900022, slot, pack_info(gen, proc == P_BSMDECAY ? 1 : 0, 0, proc), meta.w), child_key(key, 16u + (uint32_t)proc), wgt);
    O.aux[dst] = make_int2(ntr, 0);
  }
}


// ------------------------------------------------------------------------------------------ stand-alone sampling
__global__ void __launch_bounds__(256)
k_prepare_draws(const __grid_constant__ Tables T, Work W, DarkCand C, uint2* keys, const double* __restrict__ E, int n,
                int process, int lu_key, unsigned long long seed, unsigned long long first_id) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  double Ei = E[i];
  C.pf[4 * (size_t)i] = Ei;
  C.ntr[i] = 0;
  keys[i] = root_key(seed, first_id + (unsigned long long)i);
  const MapInfo& mi = T.map[process];
  int lu = (lu_key < 0 || lu_key > mi.nE) ? lookup_row(mi, Ei) : min(lu_key, mi.nE - 1);
  int bucket = process * LU_MAX + lu;
  W.bucket[i] = bucket;
  atomicAdd(&W.hist[bucket], 1);
}

// find_maxes: one CTA per map row, thread t owns sweep t (B points) - utilities/find_maxes.py:105-110
__global__ void __launch_bounds__(128)
k_find_max(const __grid_constant__ Material M, const __grid_constant__ Tables T, int process, int n_trials,
           unsigned long long seed, double* __restrict__ out_max, double* __restrict__ out_sum) {
  __shared__ __align__(128) double s_grid[GRID_SMEM_DOUBLES];
  __shared__ double s_max[128], s_sum[128];
  const MapInfo& mi = T.map[process];
  int row = blockIdx.x;
  const double* g = mi.grid + (size_t)row * mi.stride;
  for (int k = threadIdx.x; k < mi.stride; k += blockDim.x) s_grid[k] = g[k];
  __syncthreads();
  double E = mi.E[row];
  uint2 key = root_key(seed, ((unsigned long long)process << 32) | (unsigned long long)row);
  double bmax_all = 0.0, sum_all = 0.0;
  for (int sweep = threadIdx.x; sweep < n_trials; sweep += blockDim.x) {
    double bmax = -1.0 / 0.0, bsum = 0.0;
    bool has_nan = false;
    for (int j = 0; j < mi.B; ++j) {
      uint32_t t = (uint32_t)(sweep * mi.B + j);
      double D[6];
      for (int c = 0; c < (mi.dim + 1) / 2; ++c) { D2 d = draw2(key, t, ST_VEGAS, c, process); D[2 * c] = d.a; D[2 * c + 1] = d.b; }
      double x[4] = {0, 0, 0, 0}, jac = 1.0;
      for (int d = 0; d < mi.dim; ++d) {
        int ninc = mi.ninc[d];
        double yn = D[d] * ninc;
        int iy = min((int)yn, ninc - 1);
        double g0 = s_grid[mi.off[d] + iy], g1 = s_grid[mi.off[d] + iy + 1];
        double inc = g1 - g0;
        x[d] = __dadd_rn(g0, __dmul_rn(inc, yn - iy));
        jac *= inc * ninc;
      }
      double mm = (jac / mi.B) * dsigma(M, process, E, x);
      if (mm != mm) has_nan = true;
      bmax = fmax(bmax, mm);
      bsum += mm;
    }
    if (!has_nan && bmax > bmax_all) bmax_all = bmax;      // np.max of a sweep with a NaN is NaN and never wins
    sum_all += bsum;
  }
  s_max[threadIdx.x] = bmax_all; s_sum[threadIdx.x] = sum_all;
  __syncthreads();
  if (threadIdx.x == 0) {
    double m = 0.0, s = 0.0;
    for (int k = 0; k < blockDim.x; ++k) { m = fmax(m, s_max[k]); s += s_sum[k]; }
    out_max[row] = m;
    out_sum[row] = s / n_trials;
  }
}

// VEGAS training data (sum of (jac f)^2 and counts per axis increment): blockIdx.y = map (energy), dynamic shared memory =
// node grid + fp64 accumulators + counters of that map
__global__ void __launch_bounds__(256)
k_train(const __grid_constant__ Material M, int process, int dim, int stride, int4 ninc4, const double* __restrict__ grids,
        const double* __restrict__ E_inc, long long n_points, unsigned long long seed, double* __restrict__ d_out,
        double* __restrict__ n_out, double* __restrict__ integral_out, double power) {
  extern __shared__ double s_mem[];
  double* s_grid = s_mem;
  double* s_d = s_mem + stride;
  int* s_n = reinterpret_cast<int*>(s_mem + 2 * stride);
  const int row = blockIdx.y;
  const int ninc[4] = {ninc4.x, ninc4.y, ninc4.z, ninc4.w};
  int off[4]; { int o = 0; for (int d = 0; d < 4; ++d) { off[d] = o; o += (d < dim) ? ninc[d] + 1 : 0; } }
  for (int k = threadIdx.x; k < stride; k += blockDim.x) { s_grid[k] = grids[(size_t)row * stride + k]; s_d[k] = 0.0; s_n[k] = 0; }
  __syncthreads();
  const double E = E_inc[row];
  uint2 key = root_key(seed, ((unsigned long long)process << 32) | (unsigned long long)row);
  double acc = 0.0;
  const long long per_block = (n_points + gridDim.x - 1) / gridDim.x;
  const long long p0 = blockIdx.x * per_block, p1 = min(n_points, p0 + per_block);
  for (long long pt = p0 + threadIdx.x; pt < p1; pt += blockDim.x) {
    double D[6];
    for (int c = 0; c < (dim + 1) / 2; ++c) { D2 d = draw2(key, (uint32_t)pt, ST_VEGAS, c, (uint32_t)(pt >> 32)); D[2 * c] = d.a; D[2 * c + 1] = d.b; }
    double x[4] = {0, 0, 0, 0}, jac = 1.0;
    int iy[4];
    for (int d = 0; d < dim; ++d) {
      double yn = D[d] * ninc[d];
      iy[d] = min((int)yn, ninc[d] - 1);
      double g0 = s_grid[off[d] + iy[d]], g1 = s_grid[off[d] + iy[d] + 1];
      double inc = g1 - g0;
      x[d] = g0 + inc * (yn - iy[d]);
      jac *= inc * ninc[d];
    }
    double f = jac * dsigma(M, process, E, x);
    if (f == f && fabs(f) < 1e300) {
      acc += f;
      // training weight |jac f|^p.  p = 2 is Lepage's variance criterion (what vegas trains on); the SAMPLER's figure of merit is
      // mean / max of jac f, and a larger p flattens the peaks that set max_F (p = 8: 2.3-2.8x the accept rate of the shipped maps,
      // profiles/r02_final/exp_train_pow2.log).  Only ratios within an axis matter to the refinement; the clamp keeps the sums finite.
      double f2 = (power == 2.0) ? f * f : fmin(pow(fabs(f), power), 1e250);
      for (int d = 0; d < dim; ++d) { atomicAdd(&s_d[off[d] + iy[d]], f2); atomicAdd(&s_n[off[d] + iy[d]], 1); }
    }
  }
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_down_sync(0xffffffffu, acc, o);
  if ((threadIdx.x & 31) == 0 && acc != 0.0) atomicAdd(&integral_out[row], acc / (double)n_points);
  __syncthreads();
  for (int k = threadIdx.x; k < stride; k += blockDim.x)
    if (s_n[k]) { atomicAdd(&d_out[(size_t)row * stride + k], s_d[k]); atomicAdd(&n_out[(size_t)row * stride + k], (double)s_n[k]); }
}

// ------------------------------------------------------------------------------------------ tallies
__device__ __forceinline__ int species_of(int pid) {
  switch (pid) { case 11: return 0; case -11: return 1; case 22: return 2; case 13: return 3; case -13: return 4; case 4900022: return 5; }
  return 6;
}
__global__ void __launch_bounds__(256) k_tally(Stack S, long long first, long long n, double* __restrict__ tally) {
  // per-warp private histograms (8 x 1024 doubles of dynamic shared memory) and register accumulators for the three
  // per-species sums: almost every record of a shower batch falls into the same handful of bins, and one shared copy
  // serialises on them
  extern __shared__ double s_t[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  double* mine = s_t + warp * PB_TALLY_SIZE;
  for (int k = threadIdx.x; k < 8 * PB_TALLY_SIZE; k += blockDim.x) s_t[k] = 0.0;
  __syncthreads();
  double cnt[PB_TALLY_NSPECIES], ws[PB_TALLY_NSPECIES], wes[PB_TALLY_NSPECIES];
#pragma unroll
  for (int k = 0; k < PB_TALLY_NSPECIES; ++k) { cnt[k] = 0.0; ws[k] = 0.0; wes[k] = 0.0; }
  // warp-uniform loop (every lane takes part in the match operations below; lanes past the end carry weight 0 into bin 0)
  for (long long base = (blockIdx.x * (long long)blockDim.x + warp * 32); base < n; base += (long long)gridDim.x * blockDim.x) {
    const long long i = base + lane;
    const bool valid = i < n;
    long long s = first + (valid ? i : 0);
    const double2* p0p = reinterpret_cast<const double2*>(S.p0 + 4 * s);
    double2 a0 = p0p[0], a1 = p0p[1];
    const int4 kw = ld_kw(S, s);                  // pid and weight sit in the record's one 32-byte ids sector (r0w is not touched)
    double w = valid ? kw_weight(kw) : 0.0;
    int sp = species_of(ld_meta(S, s).x);
    if (valid) {
#pragma unroll
      for (int k = 0; k < PB_TALLY_NSPECIES; ++k)
        if (sp == k) { cnt[k] += 1.0; ws[k] += w; wes[k] += w * a0.x; }
    }
    int eb = (int)floor((log10(a0.x) + 3.0) * (PB_TALLY_EBINS / 6.0));
    eb = min(max(eb, 0), PB_TALLY_EBINS - 1);
    double pt = sqrt(a0.y * a0.y + a1.x * a1.x);
    double th = atan2(pt, a1.y);
    int tb = (th > 0) ? (int)floor((log10(th) + 7.0) * (PB_TALLY_TBINS / 8.0)) : 0;
    tb = min(max(tb, 0), PB_TALLY_TBINS - 1);
    // Almost every record of a shower batch falls into the same handful of bins, and a shared-memory fp64 atomicAdd is a
    // compare-and-swap loop: 32 lanes on one bin retry 32 times.  Lanes with the same (bin, weight) - weights are 1 except
    // below a decay - elect a leader that adds weight x count once: a few atomics per warp instead of 32 serialised ones.
    const unsigned same_w = __match_any_sync(0xffffffffu, __double_as_longlong(w));
    const unsigned ge = __match_any_sync(0xffffffffu, sp * PB_TALLY_EBINS + eb) & same_w;
    const unsigned gt = __match_any_sync(0xffffffffu, sp * PB_TALLY_TBINS + tb) & same_w;
    if (w != 0.0) {
      if (lane == __ffs(ge) - 1) atomicAdd(&mine[PB_TALLY_EHIST + sp * PB_TALLY_EBINS + eb], w * (double)__popc(ge));
      if (lane == __ffs(gt) - 1) atomicAdd(&mine[PB_TALLY_THIST + sp * PB_TALLY_TBINS + tb], w * (double)__popc(gt));
    }
  }
#pragma unroll
  for (int k = 0; k < PB_TALLY_NSPECIES; ++k) {
    double c = cnt[k], a = ws[k], b = wes[k];
    for (int o = 16; o > 0; o >>= 1) {
      c += __shfl_down_sync(0xffffffffu, c, o); a += __shfl_down_sync(0xffffffffu, a, o); b += __shfl_down_sync(0xffffffffu, b, o);
    }
    if (lane == 0) { mine[PB_TALLY_COUNT + k] = c; mine[PB_TALLY_WSUM + k] = a; mine[PB_TALLY_WESUM + k] = b; }
  }
  __syncthreads();
  for (int k = threadIdx.x; k < PB_TALLY_SIZE; k += blockDim.x) {
    double v = 0.0;
    for (int wv = 0; wv < 8; ++wv) v += s_t[wv * PB_TALLY_SIZE + k];
    if (v != 0.0) atomicAdd(&tally[k], v);
  }
}

// detector_cut (shower.py:815-864): one thread per record, loop over detector planes
constexpr int MAX_DET = 16;
struct DetPlanes { double z[MAX_DET]; int n; };
__global__ void __launch_bounds__(256)
k_detector_cut(Stack S, long long first, long long n, const __grid_constant__ DetPlanes D, double radius, double inner,
               double E_lo, double E_hi, double* __restrict__ sums /*[MAX_DET + 1]*/, unsigned char* __restrict__ mask) {
  double acc[MAX_DET + 1];
#pragma unroll
  for (int k = 0; k <= MAX_DET; ++k) acc[k] = 0.0;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    long long s = first + i;
    const double2* p0p = reinterpret_cast<const double2*>(S.p0 + 4 * s);
    const double2* r0p = reinterpret_cast<const double2*>(S.r0w + 4 * s);
    double2 a0 = p0p[0], a1 = p0p[1], b0 = r0p[0], b1 = r0p[1];
    bool in_E = (a0.x < E_hi) && (a0.x > E_lo);
    if (in_E) acc[MAX_DET] += b1.y;
#pragma unroll
    for (int k = 0; k < MAX_DET; ++k) {
      if (k >= D.n) break;
      double T = (D.z[k] - b1.x) / a1.y;                 // (z - z0) / pz
      double xf = b0.x + T * a0.y, yf = b0.y + T * a1.x;
      double rT = sqrt(xf * xf + yf * yf);
      bool pass = in_E && (rT > inner) && (rT < radius);
      if (pass) acc[k] += b1.y;
      if (mask) mask[i * D.n + k] = pass ? 1 : 0;
    }
  }
#pragma unroll
  for (int k = 0; k <= MAX_DET; ++k) {
    double v = acc[k];
    for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
    if ((threadIdx.x & 31) == 0 && v != 0.0) atomicAdd(&sums[k], v);
  }
}

// FP64 roofline denominator: 8 independent DFMA chains per thread
__global__ void __launch_bounds__(256) k_fp64_peak(double* out, int iters, double a, double b) {
  double x0 = threadIdx.x, x1 = x0 + 1, x2 = x0 + 2, x3 = x0 + 3, x4 = x0 + 4, x5 = x0 + 5, x6 = x0 + 6, x7 = x0 + 7;
  for (int i = 0; i < iters; ++i) {
    x0 = fma(x0, a, b); x1 = fma(x1, a, b); x2 = fma(x2, a, b); x3 = fma(x3, a, b);
    x4 = fma(x4, a, b); x5 = fma(x5, a, b); x6 = fma(x6, a, b); x7 = fma(x7, a, b);
  }
  double r = x0 + x1 + x2 + x3 + x4 + x5 + x6 + x7;
  if (r == 12345.678) out[0] = r;
}

// ------------------------------------------------------------------------------------------ replay (parity mode)
// One complete particle-step per thread - sub-step loop, final step, process choice, accept/reject sampling, kinematics - through
// the SAME device functions the wave kernels call (substep, finalize_one, the sampler's integrand forms, scatter_products), with
// every random number taken from a tape recorded from a reference run (TapeDraws).  A step that makes the reference's decisions
// consumes exactly its tape segment.
constexpr int REPLAY_IN = 10, REPLAY_OUT = 32;
__global__ void k_replay(const __grid_constant__ Material M, const __grid_constant__ Tables T, long long n, const double* __restrict__ in,
                         const double* __restrict__ tape, const long long* __restrict__ tape_off, double* __restrict__ out) {
  long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (i >= n) return;
  const double* a = in + i * REPLAY_IN;
  double* o = out + i * REPLAY_OUT;
  TapeDraws ds{tape, tape_off[i], tape_off[i + 1], false};
  const int pid = (int)a[0];
  const double mass = a[8];
  const int fl = (int)a[9];
  const int ms_e = fl & 1, flags = ((fl & 2) ? PB_FLAG_SHORT_LIVED : 0) | ((fl & 4) ? PB_FLAG_LONG_LIVED : 0);
  const bool charged = is_charged(pid) && !flags;
  V4 p{a[1], a[2], a[3], a[4]};
  double rx = a[5], ry = a[6], rz = a[7], delta_z = 0.0;
  int nsub = 0;
  if (charged) {                                                  // k_loop: track set-up, then substep() until the loop ends
    Track t;
    t.p = p; t.rx = rx; t.ry = ry; t.rz = rz;
    t.key = make_uint2(0, 0); t.it = 0;
    t.mass = mass; t.iKp = mass / (1e3 * pid_mass(pid));
    t.pmin = fmax(fmax(M.min_calc[pid_class(pid)], M.min_energy), t.mass);
    t.sp = species_index(pid);
    t.pn = norm3_nofma(p.x, p.y, p.z); t.ipn = 1.0 / t.pn;
    t.hint = nsigma_locate(T.sp[t.sp], p.E);
    t.delta_z = 0.0;
    while (!substep(M, T, t, ms_e, ds) && !ds.over) {}
    p = t.p; rx = t.rx; ry = t.ry; rz = t.rz; delta_z = t.delta_z; nsub = t.it;
  }
  bool stepped;
  int bucket = finalize_one(M, T, ds, charged, pid, flags, mass, a[1], delta_z, ms_e, p, rx, ry, rz, stepped);
  int proc = bucket / LU_MAX, lu = bucket % LU_MAX;
  double x[4] = {0.0, 0.0, 0.0, 0.0};
  long long ntr = 0;
  int status = 0;
  if (proc < N_SAMPLED) {                                         // k_sample: sequential first-accept over the recorded trials
    const MapInfo& mi = T.map[proc];
    const double* g = mi.grid + (size_t)lu * mi.stride;
    const double maxF = mi.maxF[lu] * M.fudge, E = p.E;
    SampleConst sc{0, 0, 0, 0, 0};
    if (proc == P_PAIRPROD) sc = pairprod_const(M, E);
    else if (proc == P_BREM || proc == P_MUONBREM) sc = brem_const(M, E, proc == P_BREM ? kMe : kMmu);
    bool found = false;
    while (!found && !ds.over && ntr < M.max_trials) {
      ++ntr;
      double jac = 1.0, xx[4] = {0.0, 0.0, 0.0, 0.0};
      for (int d = 0; d < mi.dim; ++d) {
        double yn = ds.next() * mi.dninc[d];
        int iy = min((int)yn, mi.ninc[d] - 1);
        double g0 = g[mi.off[d] + iy], g1 = g[mi.off[d] + iy + 1];
        double inc = g1 - g0;
        xx[d] = __dadd_rn(g0, __dmul_rn(inc, yn - iy));
        jac *= inc * mi.dninc[d];
      }
      double u = ds.next();
      double f;
      if (proc == P_PAIRPROD) f = ds_pairprod_fast(sc, E, xx);
      else if (proc == P_BREM || proc == P_MUONBREM) f = ds_brem_fast(M, sc, E, proc == P_BREM ? kMe : kMmu, xx);
      else f = dsigma(M, proc, E, xx);
      if (!ds.over && maxF * u < (jac * mi.invB) * f) { found = true; x[0] = xx[0]; x[1] = xx[1]; x[2] = xx[2]; x[3] = xx[3]; }
    }
    if (!found) { status = 3; proc = P_NONE; }
  }
  V4 da{0, 0, 0, 0}, db{0, 0, 0, 0};
  int pid_a = 0, pid_b = 0;
  double wfac = 1.0, wfac_b = 1.0, dz = 0.0;
  if (proc != P_NONE && proc != P_INPUT) scatter_products<false>(proc, pid, flags, p, mass, x, ds, da, db, pid_a, pid_b, wfac, wfac_b, dz);
  if (dz != 0.0) {
    const double pn = sqrt(p.x * p.x + p.y * p.y + p.z * p.z);
    if (pn > 0.0) { rx += p.x / pn * dz; ry += p.y / pn * dz; rz += p.z / pn * dz; }
  }
  if (ds.over) status = 1;
  else if (ds.pos != ds.end && status == 0) status = 2;
  o[0] = status; o[1] = nsub; o[2] = proc; o[3] = (double)ntr;
  o[4] = p.E; o[5] = p.x; o[6] = p.y; o[7] = p.z; o[8] = rx; o[9] = ry; o[10] = rz;
  o[11] = (da.E > M.min_energy ? 1 : 0) + (db.E > M.min_energy ? 2 : 0);                   // daughters kept (shower.py:704-706)
  o[12] = pid_a; o[13] = da.E; o[14] = da.x; o[15] = da.y; o[16] = da.z;
  o[17] = pid_b; o[18] = db.E; o[19] = db.x; o[20] = db.y; o[21] = db.z;
  o[22] = x[0]; o[23] = x[1]; o[24] = x[2]; o[25] = x[3];
  o[26] = (double)(ds.pos - tape_off[i]); o[27] = wfac; o[28] = stepped ? 1.0 : 0.0;
  o[29] = wfac_b;
}

// ------------------------------------------------------------------------------------------ dark set-up quadratures (row f-3)
// The integrals DarkShower.__init__ computes with scipy.integrate.quad (dark_shower.py:311-399, 454-493; shower.py:298-354), one
// thread per integral, through the QAGS restatement in quadpack.cuh.  Tables are evaluated like scipy's interp1d (linear,
// fill value outside): the same +, -, *, / as the host twin petite_b200.shower.LinearTable, never contracted.
struct QuadTab { const double* x; const double* y; int n; int pad; double fill; };
struct QuadTabs { const QuadTab* t; double dEdx_m; };
__device__ __forceinline__ double quad_lin(const QuadTab& T, double v) {
  int lo = 0, hi = T.n;
  while (lo < hi) { int mid = (lo + hi) >> 1; if (T.x[mid] < v) lo = mid + 1; else hi = mid; }      // searchsorted(side='left')
  hi = min(max(lo, 1), T.n - 1);
  lo = hi - 1;
  double slope = __ddiv_rn(__dsub_rn(T.y[hi], T.y[lo]), __dsub_rn(T.x[hi], T.x[lo]));
  double out = __dadd_rn(__dmul_rn(slope, __dsub_rn(v, T.x[lo])), T.y[lo]);
  return (v < T.x[0] || v > T.x[T.n - 1]) ? T.fill : out;
}
struct QuadIntegrand {
  const QuadTab* t; pb_quad_call c; double dEdx_m;
  __device__ double operator()(double E) const {
    if (c.kind == 0) return quad_lin(t[c.tab], E);                          // shower.py:298-320: the n*sigma table itself
    // dark_shower.py:311-335: n*sigma_dark(E) / dEdx * survival(E, Ei)
    double v = pow(10.0, quad_lin(t[c.tab], log10(E)));                     // dark_shower.py:31-46 (log-log table)
    if (c.cut && v < 1.0e-18) return 0.0;
    double d = 0.0;                                                        // shower.py:322-354: sum over the species' processes
    for (int k = 0; k < 3; ++k)
      if (c.surv[k] >= 0) d = __dadd_rn(d, __dsub_rn(quad_lin(t[c.surv[k]], c.Ei), quad_lin(t[c.surv[k]], E)));
    if (d < 0.0 || E > c.Ei) return 0.0;
    const double dEdx_cm = __dmul_rn(dEdx_m, kCmToM);
    return __dmul_rn(__ddiv_rn(v, dEdx_cm), exp(__ddiv_rn(__ddiv_rn(-d, dEdx_m), kCmToM)));
  }
};
__global__ void __launch_bounds__(64) k_quad(QuadTabs T, const pb_quad_call* __restrict__ calls, long long n, double* __restrict__ result,
                                             double* __restrict__ abserr, int* __restrict__ ier) {
  long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (i >= n) return;
  QuadIntegrand f{T.t, calls[i], T.dEdx_m};
  pbq::QagsOut o = pbq::qags(f, f.c.a, f.c.b, 1.49e-8, 1.49e-8);           // scipy.integrate.quad defaults
  result[i] = o.result;
  if (abserr) abserr[i] = o.abserr;
  if (ier) ier[i] = o.ier * 1000 + o.last;
}

// ------------------------------------------------------------------------------------------ probes (tests)
__global__ void k_probe(const __grid_constant__ Material M, const __grid_constant__ Tables T, int what, int process,
                        const double* __restrict__ in, long long n, int is, double* __restrict__ out, int os) {
  long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (i >= n) return;
  const double* a = in + i * is;
  double* o = out + i * os;
  switch (what) {
    case PB_PROBE_DSIGMA:
      if (process >= 32) {      // the folded forms used inside k_sample
        int p = process - 32;
        if (p == P_PAIRPROD) { SampleConst sc = pairprod_const(M, a[0]); o[0] = ds_pairprod_fast(sc, a[0], a + 1); }
        else if (p == P_DARKANN) { SampleConst sc = darkann_const(M, a[0]); o[0] = ds_darkann_c(M, sc, a[0], a[1]); }
        else if (p == P_DARKBREM || p == P_DARKMUONBREM) {
          double ml = (p == P_DARKBREM) ? kMe : kMmu;
          SampleConst sc = darkbrem_const(M, a[0], ml);
          sc.d = M.i2mT;
          o[0] = ds_darkbrem_fast(M, sc, a[0], ml, a + 1);
        }
        else { double ml = (p == P_BREM) ? kMe : kMmu; SampleConst sc = brem_const(M, a[0], ml); o[0] = ds_brem_fast(M, sc, a[0], ml, a + 1); }
      } else o[0] = dsigma(M, process, a[0], a + 1);
      break;
    case PB_PROBE_NSIGMA: o[0] = nsigma_eval(T.ns[process], a[0]); break;
    case PB_PROBE_MAP: {
      const MapInfo& mi = T.map[process];
      const double* g = mi.grid + (size_t)((int)a[0]) * mi.stride;
      double jac = 1.0;
      for (int d = 0; d < mi.dim; ++d) {
        int ninc = mi.ninc[d];
        double yn = a[1 + d] * ninc;
        int iy = min((int)yn, ninc - 1);
        double g0 = g[mi.off[d] + iy], g1 = g[mi.off[d] + iy + 1];
        double inc = g1 - g0;
        o[d] = __dadd_rn(g0, __dmul_rn(inc, yn - iy));
        jac *= inc * ninc;
      }
      o[mi.dim] = jac;
    } break;
    case PB_PROBE_MCS: {   // in: p4[4], dist_m, m_lepton, sign, z1, z2, u_phi
      V4 p{a[0], a[1], a[2], a[3]};
      double pn = norm3_nofma(p.x, p.y, p.z);
      V4 q = (pn > 0) ? mcs_apply(M, p, pn, M.rho * (a[4] / kCmToM), a[5], a[5], a[6], sqrt(a[7] * a[7] + a[8] * a[8]), a[9]) : p;
      o[0] = q.E; o[1] = q.x; o[2] = q.y; o[3] = q.z;
    } break;
    case PB_PROBE_KIN: {   // in: E, mass, x[4], u_az, u2 ; out: two four-vectors
      V4 va{0, 0, 0, 0}, vb{0, 0, 0, 0};
      switch (process) {
        case P_BREM: case P_MUONBREM: kin_brem(a[0], a[1], a + 2, a[6], &va, &vb); break;
        case P_PAIRPROD: kin_pairprod(a[0], a + 2, a[6], &va, &vb); break;
        case P_COMP: kin_compton(a[0], 0.0, a[2], a[6], &va, &vb); break;
        case P_ANN: kin_annihilation(a[0], 0.0, a[2], a[6], &va, &vb); break;
        case P_MOLLER: case P_BHABHA: kin_ee(a[0], a[2], a[6], &va, &vb); break;
        case P_MUONE: kin_mue(a[0], a[2], a[6], &va, &vb); break;
        case P_SMDECAY: two_body_decay(V4{a[0], a[2], a[3], a[4]}, a[1], 0.0, 0.0, a[6], a[7], &va, &vb); break;
      }
      o[0] = va.E; o[1] = va.x; o[2] = va.y; o[3] = va.z; o[4] = vb.E; o[5] = vb.x; o[6] = vb.y; o[7] = vb.z;
    } break;
    case PB_PROBE_HOTMATH: {
      o[0] = hot_log(a[0]);
      o[1] = hot_exp_neg_step(a[1]);
      hot_sincos(a[2], &o[2], &o[3]);
      hot_sincos_2pi(a[3], &o[4], &o[5]);
      o[6] = fast_rcp(a[0]);
      if (os >= 10) { o[7] = fast_sqrt0(a[0]); o[8] = fast_rsqrt(a[0]); o[9] = fast_sqrt0(0.0) + fast_sqrt0(-a[0]); }
      if (os >= 12) { o[10] = hot_cospi(2.0 * a[3] - 1.0); o[11] = hot_cospi(2.0 * a[3]); }
    } break;
    case PB_PROBE_MCS_FAST: {   // in as PB_PROBE_MCS; the particle's mass is m_lepton
      V4 p{a[0], a[1], a[2], a[3]};
      double pn = norm3_nofma(p.x, p.y, p.z);
      V4 q = (pn > 0) ? mcs_fast(M, p, pn, fast_rcp(pn), M.rho * (a[4] * (1.0 / kCmToM)), 1e-3, a[6], sqrt(a[7] * a[7] + a[8] * a[8]), a[9]) : p;
      o[0] = q.E; o[1] = q.x; o[2] = q.y; o[3] = q.z;
    } break;
    case PB_PROBE_SUBSTEP: {    // the track set-up of k_loop's refill (and of store_track_setup), then one substep()
      Track t;
      int pid = (int)a[0];
      t.p = V4{a[1], a[2], a[3], a[4]};
      t.rx = a[5]; t.ry = a[6]; t.rz = a[7];
      t.key = make_uint2((uint32_t)a[8], (uint32_t)a[9]);
      t.it = (int)a[10];
      t.mass = pid_mass(pid); t.iKp = 1e-3;
      t.pmin = fmax(fmax(M.min_calc[pid_class(pid)], M.min_energy), t.mass);
      t.sp = species_index(pid);
      t.pn = norm3_nofma(t.p.x, t.p.y, t.p.z); t.ipn = 1.0 / t.pn;
      t.hint = nsigma_locate(T.sp[t.sp], t.p.E);
      t.delta_z = 0.0;
      PhiloxDraws ds{t.key};
      bool done = substep(M, T, t, (int)a[11], ds);
      o[0] = done ? 1.0 : 0.0; o[1] = t.p.E; o[2] = t.p.x; o[3] = t.p.y; o[4] = t.p.z; o[5] = t.rx; o[6] = t.ry; o[7] = t.rz;
      o[8] = t.delta_z; o[9] = (double)t.it;
    } break;
    case PB_PROBE_PROPAGATE: {  // propagate_particle (shower.py:509-601) of one particle with the engine's own draws
      // in: pid, E, px, py, pz, x, y, z, mass, key0, key1, multiple scattering on/off; out: pf[4], rf[3], sub-steps, propagated
      const int pid = (int)a[0];
      const double mass = a[8];
      const int ms = (int)a[11];
      PhiloxDraws ds{make_uint2((uint32_t)a[9], (uint32_t)a[10])};
      V4 p{a[1], a[2], a[3], a[4]};
      double rx = a[5], ry = a[6], rz = a[7], delta_z = 0.0;
      int nsub = 0;
      const bool charged = is_charged(pid);
      if (charged) {
        Track t;
        t.p = p; t.rx = rx; t.ry = ry; t.rz = rz; t.key = ds.key; t.it = 0;
        t.mass = mass; t.iKp = mass / (1e3 * pid_mass(pid));
        t.pmin = fmax(fmax(M.min_calc[pid_class(pid)], M.min_energy), t.mass);
        t.sp = species_index(pid);
        t.pn = norm3_nofma(p.x, p.y, p.z); t.ipn = 1.0 / t.pn;
        t.hint = nsigma_locate(T.sp[t.sp], p.E);
        t.delta_z = 0.0;
        while (!substep(M, T, t, ms, ds)) {}
        p = t.p; rx = t.rx; ry = t.ry; rz = t.rz; delta_z = t.delta_z; nsub = t.it;
      }
      bool stepped;
      finalize_one(M, T, ds, charged, pid, 0, mass, a[1], delta_z, ms, p, rx, ry, rz, stepped);
      o[0] = p.E; o[1] = p.x; o[2] = p.y; o[3] = p.z; o[4] = rx; o[5] = ry; o[6] = rz; o[7] = nsub; o[8] = stepped ? 1.0 : 0.0;
    } break;
    case PB_PROBE_DARKKIN: {    // in: E, mV, x[4], u1, u2, Pe, cte
      V4 v{0, 0, 0, 0};
      if (process == P_DARKBREM || process == P_DARKMUONBREM) v = kin_darkbrem_V(a[0], a[1], a + 2, a[6]);
      else if (process == P_DARKANN) v = kin_darkann_V(a[0], a[1], a[2]);
      else if (process == P_DARKCOMP) v = kin_compton_bound_V(a[0], a[1], a[2], a[8], a[9], a[6], a[7]);
      o[0] = v.E; o[1] = v.x; o[2] = v.y; o[3] = v.z;
    } break;
    case PB_PROBE_PHILOX: {
      D2 d = draw2(make_uint2((uint32_t)a[0], (uint32_t)a[1]), (uint32_t)a[2], (uint32_t)a[3], (uint32_t)a[4], (uint32_t)a[5]);
      o[0] = d.a; o[1] = d.b;
    } break;
  }
}

}  // namespace pb

// =============================================================================================== host side
using namespace pb;

struct pb_engine_s {
  int device = 0;
  pb_config cfg{};
  Material mat{};
  Tables tab{};
  std::vector<void*> owned;      // device allocations for tables
  std::vector<double> ns_x[16], ns_y[16];   // host copies of the n*sigma tables (for the per-species sums)
  bool species_dirty = true;
  Work work{};
  long long work_n = 0;          // capacity of per-wave scratch
  void* work_blob = nullptr;
  void* fixed_blob = nullptr;    // hist/offsets/cursor/ctrl/tail/counters
  int* order_blob = nullptr; long long order_cap = 0;
  DarkTables dark{}; bool dark_ready = false;
  void* cand_blob = nullptr; long long cand_cap = 0; DarkCand cand{};
  double* prim_mass = nullptr; long long prim_cap = 0;
  void* prim_stage = nullptr; size_t prim_stage_bytes = 0;
  int n_sm = 148;
  int profiling = 0;             // 0 off, 1 = the two dominant kernels only (k_loop, k_sample), 2 = every kernel
  int sample_group = PB_SAMPLE_G_DEFAULT;    // lanes cooperating on one accept/reject sample (tuning knob, PB_SAMPLE_G)
  int sample_trials = PB_SAMPLE_T_DEFAULT;   // trials per lane and round (PB_SAMPLE_T): ILP inside the lane
  int sample_group_dark = PB_SAMPLE_G_DARK_DEFAULT;   // lanes per sample in the dark pass / stand-alone sampling (PB_SAMPLE_G_DARK)
  int sample_trials_dark = PB_SAMPLE_T_DARK_DEFAULT;   // the same for the generic kernel (PB_SAMPLE_T_DARK)
  int sample_trials_db = PB_SAMPLE_T_DB_DEFAULT;       // the same for the dark-brem family kernel (PB_SAMPLE_T_DB)
  int sample_split_db = 1;                             // PB_SAMPLE_SPLIT_DB=0: dark pass through one generic launch
  int sample_generic = 0;        // PB_SAMPLE_GENERIC=1: SM pass through the generic kernel (with the dark integrands compiled in)
  int emit_wave_order = 0;       // PB_EMIT_ORDER=1: k_emit walks the wave in record order (coalesced) instead of bucket order
  static constexpr int LOOKAHEAD = 8;   // waves enqueued per host synchronisation while the shower tail shrinks
  cudaEvent_t ev[2 * 8] = {};
  cudaEvent_t evp[LOOKAHEAD][2][2] = {};   // lookahead slot x {k_loop, k_sample} x {start, stop}
  WaveState* h_ws = nullptr;             // pinned read-back of the device wave state (+ tail)
  // host-free wave loop: one CUDA graph = WHILE node around the seven wave kernels, condition set on the device
  int use_graph = 1;                     // PB_GRAPH=0: stream launches with a host synchronisation per (growing) wave
  cudaGraph_t wave_graph = nullptr;
  cudaGraphExec_t wave_exec = nullptr;
  cudaStream_t cap_stream = nullptr;     // capture stream for building the graph body
  cudaStream_t aux_stream = nullptr;     // set-up / probe entry points (find_max, training, probes, peak measurement) run here, not on the
                                         // default stream: they must not serialise against another handle's wave loop
  std::vector<unsigned char> graph_sig;  // kernel arguments the instantiated graph was built with
  pb_profile prof{};
  std::string err;
  std::string tlog_path;
};

#define PB_CUDA(e, call)                                                                         \
  do {                                                                                           \
    cudaError_t _c = (call);                                                                     \
    if (_c != cudaSuccess) {                                                                     \
      (e)->err = std::string(#call) + ": " + cudaGetErrorString(_c);                             \
      return PB_ERR_CUDA;                                                                        \
    }                                                                                            \
  } while (0)

static void derive_material(pb_engine e) {
  const pb_config& c = e->cfg;
  Material& m = e->mat;
  memset(&m, 0, sizeof(m));
  m.Z = c.Z_T; m.A = c.A_T; m.rho = c.rho; m.dEdx = c.dEdx_GeV_per_m; m.mT = c.mT_sampler;
  m.min_energy = c.min_energy; m.Eg_min = c.Eg_min; m.Ee_min = c.Ee_min;
  m.fudge = c.maxF_fudge; m.rescale_mcs = c.rescale_MCS;
  for (int i = 0; i < 5; ++i) m.min_calc[i] = c.min_calc[i];
  // all_processes.py:92-103 ; host libm pow, as CPython's float ** float
  double a0 = 184.15 * pow(2.718, -0.5) * pow(c.Z_T, -1.0 / 3.0) / kMe;
  m.ff_a0sq = a0 * a0;
  m.ff_Z2a04 = (c.Z_T * c.Z_T) * pow(a0, 4);
  // all_processes.py:123-133
  double c1 = pow(111 * pow(c.Z_T, -1.0 / 3) / kMe, 2);
  m.dff_c1 = c1;
  m.dff_c2 = 0.164 * pow(c.A_T, -2.0 / 3);
  m.dff_ic2 = 1.0 / m.dff_c2;
  m.dff_ap2 = pow(773.0 * pow(c.Z_T, -2.0 / 3) / kMe, 2);
  m.dff_inel_pref = c.Z_T / (c1 * c1 * (c.Z_T * c.Z_T));
  m.dff_pref = (c.Z_T * c.Z_T) * (c1 * c1);
  m.Z23 = pow(c.Z_T, 2.0 / 3.0);
  m.i2mT = 1.0 / (2.0 * c.mT_sampler);
  m.mcs_C4 = 0.157 * c.Z_T * (c.Z_T + 1) / c.A_T;
  m.mcs_Cw = m.mcs_C4 / (2.007e-5 * m.Z23);
  m.mcs_c3 = 3.34 * (c.Z_T * kAlpha) * (c.Z_T * kAlpha);
  m.me4 = pow(kMe, 4);
  m.mV4 = pow(c.mV, 4);
  m.mV = c.mV; m.g_e = c.g_e; m.eps = c.kinetic_mixing; m.Zeff = c.Zeff;
  m.E_res_ann = c.E_res_ann; m.E_thr_comp = c.E_thr_comp; m.bound_electron = c.bound_electron;
  m.max_trials = 0;
}

extern "C" const char* pb_version(void) { return "petite_b200 0.1 (sm_100a)"; }
extern "C" const char* pb_last_error(pb_engine e) { return e ? e->err.c_str() : "null engine"; }

extern "C" int pb_create(pb_engine* out, int device, const pb_config* cfg) {
  if (!out || !cfg) return PB_ERR_ARG;
  pb_engine e = new pb_engine_s();
  e->device = device;
  cudaError_t c = cudaSetDevice(device);
  if (c != cudaSuccess) { delete e; return PB_ERR_CUDA; }
  cudaDeviceProp prop;
  if (cudaGetDeviceProperties(&prop, device) == cudaSuccess) e->n_sm = prop.multiProcessorCount;
  e->cfg = *cfg;
  derive_material(e);
  if (const char* g = getenv("PB_SAMPLE_G")) e->sample_group = atoi(g);
  if (const char* g = getenv("PB_SAMPLE_T")) e->sample_trials = atoi(g);
  if (const char* g = getenv("PB_SAMPLE_G_DARK")) e->sample_group_dark = atoi(g);
  if (const char* g = getenv("PB_SAMPLE_T_DARK")) e->sample_trials_dark = atoi(g);
  if (const char* g = getenv("PB_SAMPLE_GENERIC")) e->sample_generic = atoi(g);
  if (const char* g = getenv("PB_SAMPLE_T_DB")) e->sample_trials_db = atoi(g);
  if (const char* g = getenv("PB_SAMPLE_SPLIT_DB")) e->sample_split_db = atoi(g);
  if (const char* g = getenv("PB_EMIT_ORDER")) e->emit_wave_order = atoi(g);
  if (const char* g = getenv("PB_GRAPH")) e->use_graph = atoi(g);
  e->work.tile_norm = 1;
  if (const char* g = getenv("PB_TILE_NORM")) e->work.tile_norm = atoi(g);
  e->work.drain_lanes = PB_DRAIN_LANES_DEFAULT;
  if (const char* g = getenv("PB_DRAIN_LANES")) e->work.drain_lanes = std::min(std::max(atoi(g), 0), 32);
  e->work.loop_cap = PB_LOOP_CAP_DEFAULT;
  if (const char* g = getenv("PB_LOOP_CAP")) e->work.loop_cap = atoi(g);
  if (e->work.loop_cap < 0 || (e->work.loop_cap & (e->work.loop_cap - 1))) e->work.loop_cap = 0;      // powers of two only
  if (const char* g = getenv("PB_TILE_LOG")) {            // "<file>:<wave>": per-tile timeline of k_sample in that wave of every run (appended)
    std::string v(g);
    size_t c = v.rfind(':');
    if (c != std::string::npos) {
      e->tlog_path = v.substr(0, c);
      e->work.tlog_wave = atoi(v.c_str() + c + 1);
      e->work.tlog_cap = 1 << 17;
      if (cudaMalloc(&e->work.tlog, sizeof(unsigned long long) * 4 * (size_t)e->work.tlog_cap) != cudaSuccess) e->work.tlog = nullptr;
      else cudaMemset(e->work.tlog, 0, sizeof(unsigned long long) * 4 * (size_t)e->work.tlog_cap);
    }
  }
  size_t fixed = sizeof(int) * (NBUCKET * 6 + 3 + 16) + sizeof(unsigned long long) * (8 + CNT_N + 2 * NBUCKET) + sizeof(WaveState) + 64;
  if (cudaMalloc(&e->fixed_blob, fixed) != cudaSuccess) { delete e; return PB_ERR_CUDA; }
  cudaMemset(e->fixed_blob, 0, fixed);
  bool ok = cudaStreamCreateWithFlags(&e->aux_stream, cudaStreamNonBlocking) == cudaSuccess;
  for (int i = 0; i < 2 * 8 && ok; ++i) ok = cudaEventCreate(&e->ev[i]) == cudaSuccess;
  for (int a = 0; a < pb_engine_s::LOOKAHEAD && ok; ++a) for (int b = 0; b < 2 && ok; ++b) for (int c = 0; c < 2 && ok; ++c)
    ok = cudaEventCreate(&e->evp[a][b][c]) == cudaSuccess;
  if (ok) ok = cudaMallocHost(&e->h_ws, sizeof(WaveState) + 32) == cudaSuccess;
  if (!ok) { pb_destroy(e); return PB_ERR_CUDA; }
  char* p = (char*)e->fixed_blob;
  e->work.tail = (unsigned long long*)p; p += 8 * sizeof(unsigned long long);
  e->work.counters = (unsigned long long*)p; p += CNT_N * sizeof(unsigned long long);
  e->work.bstat = (unsigned long long*)p; p += 2 * NBUCKET * sizeof(unsigned long long);
  e->work.hist = (int*)p; p += NBUCKET * sizeof(int);
  e->work.offsets = (int*)p; p += (NBUCKET + 1) * sizeof(int);
  e->work.cursor = (int*)p; p += NBUCKET * sizeof(int);
  e->work.tile_base = (int*)p; p += 2 * (NBUCKET + 1) * sizeof(int);
  e->work.tile_sz = (int*)p; p += NBUCKET * sizeof(int);
  e->work.ctrl = (int*)p; p += 16 * sizeof(int);
  p = (char*)(((uintptr_t)p + 15) & ~(uintptr_t)15);
  e->work.ws = (WaveState*)p;
  *out = e;
  return PB_OK;
}

extern "C" int pb_set_config(pb_engine e, const pb_config* cfg) {
  if (!e || !cfg) return PB_ERR_ARG;
  e->cfg = *cfg;
  derive_material(e);
  return PB_OK;
}

extern "C" void pb_destroy(pb_engine e) {
  if (!e) return;
  cudaSetDevice(e->device);
  for (void* p : e->owned) cudaFree(p);
  if (e->work_blob) cudaFree(e->work_blob);
  if (e->fixed_blob) cudaFree(e->fixed_blob);
  if (e->order_blob) cudaFree(e->order_blob);
  if (e->cand_blob) cudaFree(e->cand_blob);
  if (e->prim_mass) cudaFree(e->prim_mass);
  if (e->prim_stage) cudaFree(e->prim_stage);
  for (int i = 0; i < 2 * 8; ++i) if (e->ev[i]) cudaEventDestroy(e->ev[i]);
  for (int a = 0; a < pb_engine_s::LOOKAHEAD; ++a) for (int b = 0; b < 2; ++b) for (int c = 0; c < 2; ++c) if (e->evp[a][b][c]) cudaEventDestroy(e->evp[a][b][c]);
  if (e->h_ws) cudaFreeHost(e->h_ws);
  if (e->work.tlog) cudaFree(e->work.tlog);
  if (e->wave_exec) cudaGraphExecDestroy(e->wave_exec);
  if (e->wave_graph) cudaGraphDestroy(e->wave_graph);
  if (e->cap_stream) cudaStreamDestroy(e->cap_stream);
  if (e->aux_stream) cudaStreamDestroy(e->aux_stream);
  delete e;
}

// Coarse look-up table over an increasing grid of positive energies (struct Coarse): ub[b] = searchsorted_left(x, upper edge of bin
// key0 + b), optionally clipped to [lo_clip, hi_clip] (the interpolation tables use scipy's clip(., 1, n - 1)).
static int build_coarse(pb_engine e, const double* x, int n, int lo_clip, int hi_clip, Coarse* out) {
  *out = Coarse{nullptr, 0, 0};
  if (n < 1 || !(x[0] > 0.0)) return PB_OK;
  auto key = [](double v) { int64_t b; memcpy(&b, &v, 8); return (int)(b >> 32) >> COARSE_SHIFT; };
  const int key0 = key(x[0]), nb = key(x[n - 1]) - key0 + 1;
  std::vector<int> ub((size_t)nb);
  for (int b = 0; b < nb; ++b) {
    int64_t bits = (int64_t)(key0 + b + 1) << (32 + COARSE_SHIFT);       // smallest double of the next bin
    double edge; memcpy(&edge, &bits, 8);
    int s = (int)(std::lower_bound(x, x + n, edge) - x);
    ub[b] = std::min(std::max(s, lo_clip), hi_clip);
  }
  int* d = nullptr;
  PB_CUDA(e, cudaMalloc(&d, sizeof(int) * (size_t)nb));
  e->owned.push_back(d);
  PB_CUDA(e, cudaMemcpy(d, ub.data(), sizeof(int) * (size_t)nb, cudaMemcpyHostToDevice));
  *out = Coarse{d, key0, nb};
  return PB_OK;
}

extern "C" int pb_upload_nsigma(pb_engine e, int id, const double* E, const double* y, int n) {
  if (!e || id < 0 || id >= 16 || n < 0) return PB_ERR_ARG;
  PB_CUDA(e, cudaSetDevice(e->device));
  std::vector<double> node(4 * (size_t)std::max(n, 1), 0.0);
  for (int i = 0; i < n; ++i) {
    node[4 * i] = E[i]; node[4 * i + 1] = y[i];
    node[4 * i + 2] = (i + 1 < n) ? (y[i + 1] - y[i]) / (E[i + 1] - E[i]) : 0.0;    // scipy's per-call slope, hoisted
  }
  double* d = nullptr;
  PB_CUDA(e, cudaMalloc(&d, sizeof(double) * node.size()));
  e->owned.push_back(d);
  PB_CUDA(e, cudaMemcpy(d, node.data(), sizeof(double) * node.size(), cudaMemcpyHostToDevice));
  if (n >= 2 && !(E[0] > 0.0)) { e->err = "n*sigma tables need positive energies"; return PB_ERR_ARG; }
  for (int i = 1; i < n; ++i) if (!(E[i] >= E[i - 1])) { e->err = "n*sigma table energies must not decrease"; return PB_ERR_ARG; }
  Coarse c;
  { int rcc = build_coarse(e, E, n >= 2 ? n : 0, 1, n - 1, &c); if (rcc != PB_OK) return rcc; }
  e->tab.ns[id] = NSigmaTable{(const double4*)d, n ? E[0] : 0.0, n ? E[n - 1] : 0.0, c, n >= 2 ? n : 0, 0};
  e->ns_x[id].assign(E, E + n);
  e->ns_y[id].assign(y, y + n);
  e->species_dirty = true;
  return PB_OK;
}

// Total n*sigma(E) of each charged species as ONE piecewise-linear table on the union of its processes' grids.  A member
// table contributes to a union segment iff the segment lies inside its range (scipy's fill_value = 0 outside), with the
// value and slope of its own segment there; nodes store the right-hand limit, so the jump at a member's upper end is exact.
// (At a member's lower end the single point E == x_0 gets 0 instead of y_0.)  Sums differ from the reference's
// term-by-term sum by rounding only.
static int build_species_tables(pb_engine e) {
  static const int members[3][3] = {{P_BREM, P_MOLLER, -1}, {P_BREM, P_BHABHA, P_ANN}, {P_MUONBREM, P_MUONE, -1}};
  for (int sp = 0; sp < 3; ++sp) {
    std::vector<double> u;
    for (int k = 0; k < 3; ++k) {
      int id = members[sp][k];
      if (id >= 0 && e->ns_x[id].size() >= 2) u.insert(u.end(), e->ns_x[id].begin(), e->ns_x[id].end());
    }
    std::sort(u.begin(), u.end());
    u.erase(std::unique(u.begin(), u.end()), u.end());
    const int m = (int)u.size();
    std::vector<double> node(4 * (size_t)std::max(m, 1), 0.0);
    for (int j = 0; j < m; ++j) {
      double Y = 0.0, S = 0.0;
      if (j + 1 < m) {
        for (int k = 0; k < 3; ++k) {
          int id = members[sp][k];
          if (id < 0 || e->ns_x[id].size() < 2) continue;
          const std::vector<double>& x = e->ns_x[id];
          const std::vector<double>& y = e->ns_y[id];
          const int n = (int)x.size();
          if (!(x[0] <= u[j] && u[j + 1] <= x[n - 1])) continue;
          int i = (int)(std::upper_bound(x.begin(), x.end(), u[j]) - x.begin()) - 1;
          i = std::min(std::max(i, 0), n - 2);
          double slope = (y[i + 1] - y[i]) / (x[i + 1] - x[i]);
          Y += slope * (u[j] - x[i]) + y[i];
          S += slope;
        }
      }
      node[4 * j] = u[j]; node[4 * j + 1] = Y; node[4 * j + 2] = S;
    }
    double* d = nullptr;
    PB_CUDA(e, cudaMalloc(&d, sizeof(double) * node.size()));
    e->owned.push_back(d);
    PB_CUDA(e, cudaMemcpy(d, node.data(), sizeof(double) * node.size(), cudaMemcpyHostToDevice));
    Coarse c;
    { int rcc = build_coarse(e, u.data(), m >= 2 ? m : 0, 1, m - 1, &c); if (rcc != PB_OK) return rcc; }
    e->tab.sp[sp] = NSigmaTable{(const double4*)d, m ? u[0] : 0.0, m ? u[m - 1] : 0.0, c, m >= 2 ? m : 0, 0};
  }
  e->species_dirty = false;
  return PB_OK;
}

extern "C" int pb_upload_maps(pb_engine e, int process, const double* grid, int nE, int dim, const int32_t* ninc,
                              const double* E_inc, const double* max_F, int neval) {
  if (!e || process < 0 || process >= N_SAMPLED || dim < 1 || dim > 4 || nE < 1 || nE > LU_MAX) return PB_ERR_ARG;
  if (dim != proc_dim(process)) { e->err = "map dimension does not match process"; return PB_ERR_ARG; }
  PB_CUDA(e, cudaSetDevice(e->device));
  MapInfo mi{};
  int stride = 0;
  for (int d = 0; d < dim; ++d) { mi.ninc[d] = ninc[d]; mi.dninc[d] = (double)ninc[d]; mi.off[d] = stride; stride += ninc[d] + 1; }
  int padded = (stride + 15) / 16 * 16;          // rows start 128-byte aligned; TMA bulk size multiple of 16 B
  if (padded > GRID_SMEM_DOUBLES) { e->err = "map row larger than the shared-memory staging buffer"; return PB_ERR_ARG; }
  std::vector<double> host((size_t)padded * nE, 0.0);
  for (int r = 0; r < nE; ++r) memcpy(&host[(size_t)r * padded], grid + (size_t)r * stride, sizeof(double) * stride);
  double* d = nullptr;
  PB_CUDA(e, cudaMalloc(&d, sizeof(double) * ((size_t)padded * nE + 2 * (size_t)nE)));
  e->owned.push_back(d);
  PB_CUDA(e, cudaMemcpy(d, host.data(), sizeof(double) * host.size(), cudaMemcpyHostToDevice));
  double* dE = d + (size_t)padded * nE;
  PB_CUDA(e, cudaMemcpy(dE, E_inc, sizeof(double) * nE, cudaMemcpyHostToDevice));
  PB_CUDA(e, cudaMemcpy(dE + nE, max_F, sizeof(double) * nE, cudaMemcpyHostToDevice));
  mi.grid = d; mi.E = dE; mi.maxF = dE + nE; mi.nE = nE; mi.dim = dim; mi.stride = padded; mi.B = neval; mi.invB = 1.0 / (double)neval;
  if (!(E_inc[0] > 0.0)) { e->err = "map energies must be positive"; return PB_ERR_ARG; }
  for (int r = 1; r < nE; ++r) if (!(E_inc[r] >= E_inc[r - 1])) { e->err = "map energies must not decrease"; return PB_ERR_ARG; }
  { int rcc = build_coarse(e, E_inc, nE, 0, nE, &mi.c); if (rcc != PB_OK) return rcc; }
  e->tab.map[process] = mi;
  return PB_OK;
}

static int ensure_cand(pb_engine e, long long ncap);
// SM pass: a kernel instantiated without the dark integrands (fewer registers, one more CTA per SM); dark pass and
// stand-alone sampling: the generic instantiation.  (G, T) = lanes per sample x trials per lane and round.
template <int FAM>
static void launch_sample_fam(pb_engine e, int grid, const SampleIO& io, cudaStream_t stream, int family_mode) {
  const int G = (FAM == 2) ? e->sample_group : e->sample_group_dark;
  const int T = (FAM == 2) ? e->sample_trials : (FAM == 3 ? e->sample_trials_db : e->sample_trials_dark);
#define PB_LS(g, t) if (G == g && T == t) { k_sample<g, FAM, t><<<grid, SAMPLE_THREADS, 0, stream>>>(e->mat, e->tab, io, e->work, family_mode); return; }
  PB_LS(4, 1) PB_LS(2, 2) PB_LS(1, 2) PB_LS(4, 2) PB_LS(2, 1) PB_LS(8, 1) PB_LS(8, 2) PB_LS(1, 4) PB_LS(2, 4) PB_LS(1, 1)
#undef PB_LS
  k_sample<4, FAM, 1><<<grid, SAMPLE_THREADS, 0, stream>>>(e->mat, e->tab, io, e->work, family_mode);
}
static int sample_grid(pb_engine e, long long n, bool sm_pass, bool db = false) {
  const int T = db ? e->sample_trials_db : (sm_pass && !e->sample_generic ? e->sample_trials : e->sample_trials_dark);
  const int minb = db ? PB_SAMPLE_MINB_DB : (sm_pass ? (T > 1 ? PB_SAMPLE_MINB_SM_T : PB_SAMPLE_MINB_SM) : (T > 1 ? PB_SAMPLE_MINB_T : PB_SAMPLE_MINB));
  return (int)std::max<long long>(1, std::min<long long>((long long)e->n_sm * minb, (n + 31) / 32 + 1));
}
// SM pass: the FAM = 2 kernel (no dark integrands: fewer registers, one more CTA per SM).  Dark pass / stand-alone sampling: the
// dark-brem tiles (3-D folded form, ~70-300 trials per sample: 90 % of a dark pass) go to the FAM = 3 kernel and everything else to
// the generic one, two launches over the same tile table with their own cursors (PB_SAMPLE_SPLIT_DB=0: one generic launch).
static void launch_sample(pb_engine e, long long n, const SampleIO& io, cudaStream_t stream, bool sm_pass = false) {
  if (sm_pass && !e->sample_generic) { launch_sample_fam<2>(e, sample_grid(e, n, true), io, stream, 0); return; }
  if (sm_pass || !e->sample_split_db) { launch_sample_fam<-1>(e, sample_grid(e, n, sm_pass), io, stream, 0); return; }
  launch_sample_fam<3>(e, sample_grid(e, n, false, true), io, stream, 0);
  launch_sample_fam<-1>(e, sample_grid(e, n, false), io, stream, 1);
}

// index lists live apart from the other scratch: growing them must keep their contents (a paused wave still needs them)
static int ensure_order(pb_engine e, long long n_list, cudaStream_t stream) {
  if (n_list <= e->order_cap) return PB_OK;
  long long cap = std::max<long long>(n_list * 5 / 4, 1 << 16);
  int* blob = nullptr;
  PB_CUDA(e, cudaMalloc(&blob, sizeof(int) * 6 * (size_t)cap));
  if (e->order_blob) {
    for (int k = 0; k < 6; ++k)
      PB_CUDA(e, cudaMemcpyAsync(blob + k * cap, k < 4 ? e->work.list[k] : e->work.carry[k - 4], sizeof(int) * e->order_cap, cudaMemcpyDeviceToDevice, stream));
    PB_CUDA(e, cudaStreamSynchronize(stream));
    cudaFree(e->order_blob);
  }
  e->order_blob = blob;
  for (int k = 0; k < 4; ++k) e->work.list[k] = blob + k * cap;
  for (int k = 0; k < 2; ++k) e->work.carry[k] = blob + (4 + k) * cap;
  e->order_cap = cap;
  return PB_OK;
}

static int ensure_work(pb_engine e, long long n) {
  if (n <= e->work_n) return PB_OK;
  long long cap = std::max<long long>(n * 5 / 4, 1 << 16);
  if (e->work_blob) { cudaFree(e->work_blob); e->work_blob = nullptr; }
  long long max_tiles = cap / TILE_MIN + N_SAMPLED * LU_MAX + 1;
  size_t bytes = (size_t)cap * (4 + 8 + 32 + 8 + 8) + (size_t)max_tiles * 12 + 256;
  PB_CUDA(e, cudaMalloc(&e->work_blob, bytes));
  char* p = (char*)e->work_blob;
  e->work.xs = (double*)p; p += (size_t)cap * 32;
  e->work.sE = (double*)p; p += (size_t)cap * 8;
  e->work.skey = (uint2*)p; p += (size_t)cap * 8;
  e->work.sorted = (int2*)p; p += (size_t)cap * 8;
  e->work.bucket = (int*)p; p += (size_t)cap * 4;
  e->work.tile_bucket = (int*)p; p += (size_t)max_tiles * 4;
  e->work.tile_start = (int*)p; p += (size_t)max_tiles * 4;
  e->work.tile_count = (int*)p;
  e->work_n = cap;
  return PB_OK;
}

// The wave loop as ONE graph launch: WHILE(condition) { wave_begin (sets the condition); loop; finalize; scan; fill; sample; emit }.
// Every wave kernel is grid-stride / persistent over the device-side wave size, so the grids are fixed (a multiple of the SM
// count); kernel arguments are baked into the graph, which is rebuilt when any of them changes (new stack, grown scratch,
// new tables or configuration).  Measured: 1.45 us per kernel node vs 3.3 us per stream launch + 12 us per host
// synchronisation (profiles/r02_summary.md).
static int ensure_wave_graph(pb_engine e, const Stack& S, int ms_flag) {
  std::vector<unsigned char> sig(sizeof(Material) + sizeof(Tables) + sizeof(Stack) + sizeof(Work) + 5 * sizeof(long long));
  unsigned char* q = sig.data();
  memcpy(q, &e->mat, sizeof(Material)); q += sizeof(Material);
  memcpy(q, &e->tab, sizeof(Tables)); q += sizeof(Tables);
  memcpy(q, &S, sizeof(Stack)); q += sizeof(Stack);
  memcpy(q, &e->work, sizeof(Work)); q += sizeof(Work);
  long long extra[5] = {(long long)(uintptr_t)e->prim_mass, ms_flag, e->sample_group * 100 + e->sample_trials, e->sample_generic, e->emit_wave_order};
  memcpy(q, extra, sizeof(extra));
  if (e->wave_exec && sig == e->graph_sig) return PB_OK;
  if (e->wave_exec) { cudaGraphExecDestroy(e->wave_exec); e->wave_exec = nullptr; }
  if (e->wave_graph) { cudaGraphDestroy(e->wave_graph); e->wave_graph = nullptr; }
  if (!e->cap_stream) PB_CUDA(e, cudaStreamCreateWithFlags(&e->cap_stream, cudaStreamNonBlocking));
  PB_CUDA(e, cudaGraphCreate(&e->wave_graph, 0));
  cudaGraphConditionalHandle h;
  PB_CUDA(e, cudaGraphConditionalHandleCreate(&h, e->wave_graph, 1, cudaGraphCondAssignDefault));
  cudaGraphNodeParams np = {};
  np.type = cudaGraphNodeTypeConditional;
  np.conditional.handle = h;
  np.conditional.type = cudaGraphCondTypeWhile;
  np.conditional.size = 1;
  cudaGraphNode_t node;
  PB_CUDA(e, cudaGraphAddNode(&node, e->wave_graph, nullptr, 0, &np));
  cudaGraph_t body = np.conditional.phGraph_out[0];
  cudaStream_t cs = e->cap_stream;
  PB_CUDA(e, cudaStreamBeginCaptureToGraph(cs, body, nullptr, nullptr, 0, cudaStreamCaptureModeThreadLocal));
  // one full wave of resident CTAs per kernel (occupancy query): a partial second wave of a grid-stride kernel only adds a tail
  auto resident = [&](const void* fn, int threads) {
    int nb = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, fn, threads, 0) != cudaSuccess || nb < 1) nb = 1;
    return e->n_sm * nb;
  };
  const int g_loop = e->n_sm * PB_LOOP_MINB, g_fin = resident((const void*)k_finalize, 128), g_emit = resident((const void*)k_emit, 128),
            g_fill = resident((const void*)k_bucket_fill, 256);
  k_wave_begin_cond<<<1, 1, 0, cs>>>(e->work, h);
  k_loop<<<g_loop, 128, 0, cs>>>(e->mat, e->tab, S, e->work, e->prim_mass, ms_flag);
  k_finalize<<<g_fin, 128, 0, cs>>>(e->mat, e->tab, S, e->work, e->prim_mass, ms_flag);
  k_bucket_scan<<<1, 1024, 0, cs>>>(e->work);
  SampleIO io{S.pf, reinterpret_cast<const uint2*>(S.ids), 4, 2, nullptr, reinterpret_cast<int*>(S.aux), 2, e->work.ws};
  k_bucket_fill<<<g_fill, 256, 0, cs>>>(e->work, io, -1);
  launch_sample(e, 1LL << 40, io, cs, true);
  k_emit<<<g_emit, 128, 0, cs>>>(e->mat, e->tab, S, e->work, e->emit_wave_order, e->prim_mass + e->prim_cap);
  cudaError_t ce = cudaStreamEndCapture(cs, nullptr);
  if (ce != cudaSuccess) { e->err = std::string("wave graph capture: ") + cudaGetErrorString(ce); return PB_ERR_CUDA; }
  PB_CUDA(e, cudaGraphInstantiate(&e->wave_exec, e->wave_graph, 0));
  e->graph_sig.swap(sig);
  return PB_OK;
}

extern "C" int pb_run_showers(pb_engine e, const pb_primaries* prim, uint64_t seed, uint64_t first_id, int global_ms,
                              pb_stack* st, pb_counters* out, void* stream_) {
  if (!e || !prim || !st) return PB_ERR_ARG;
  cudaStream_t stream = (cudaStream_t)stream_;
  PB_CUDA(e, cudaSetDevice(e->device));
  long long n0 = prim->n;
  if (n0 <= 0 || n0 > st->capacity) { e->err = "primaries exceed stack capacity"; return PB_ERR_CAPACITY; }
  for (int p = 0; p < 8; ++p)
    if (e->tab.map[p].grid == nullptr || e->tab.ns[p].n == 0) { e->err = "tables not uploaded"; return PB_ERR_STATE; }
  if (e->species_dirty) { int rcs = build_species_tables(e); if (rcs != PB_OK) return rcs; }
  int B = e->tab.map[P_BREM].B;
  e->mat.max_trials = (long long)std::min<double>((double)e->cfg.max_sweeps * (double)B, 4.0e9);
  Stack S{st->p0, st->r0w, st->pf, st->rf, (int4*)st->ids, (int2*)st->aux, st->capacity};
  long long launches = 0;
  // ---- primaries: host SoA -> device staging -> stack records [0, n0)
  size_t stage_bytes = (size_t)n0 * (sizeof(double) * (4 + 3 + 1 + 1) + sizeof(int) * 2);
  if (stage_bytes > e->prim_stage_bytes) {
    if (e->prim_stage) cudaFree(e->prim_stage);
    PB_CUDA(e, cudaMalloc(&e->prim_stage, stage_bytes));
    e->prim_stage_bytes = stage_bytes;
  }
  if (n0 > e->prim_cap) {
    if (e->prim_mass) cudaFree(e->prim_mass);
    PB_CUDA(e, cudaMalloc(&e->prim_mass, sizeof(double) * 4 * n0));      // masses, then (decay path, two weight factors) per primary: k_init_primaries
    e->prim_cap = n0;
  }
  char* sp = (char*)e->prim_stage;
  const double* d_p = (double*)sp; sp += sizeof(double) * 4 * n0;
  const double* d_r = (double*)sp; sp += sizeof(double) * 3 * n0;
  const double* d_w = (double*)sp; sp += sizeof(double) * n0;
  const int* d_pid = (int*)sp; sp += sizeof(int) * n0;
  const int* d_fl = (int*)sp;
  const cudaMemcpyKind kind = prim->on_device ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice;
  PB_CUDA(e, cudaMemcpyAsync(e->prim_mass, prim->mass, sizeof(double) * n0, kind, stream));
  if (prim->on_device) {
    d_p = prim->p; d_r = prim->r; d_w = prim->weight; d_pid = prim->pid; d_fl = prim->flags;
  } else {
    PB_CUDA(e, cudaMemcpyAsync((void*)d_p, prim->p, sizeof(double) * 4 * n0, kind, stream));
    PB_CUDA(e, cudaMemcpyAsync((void*)d_r, prim->r, sizeof(double) * 3 * n0, kind, stream));
    PB_CUDA(e, cudaMemcpyAsync((void*)d_w, prim->weight, sizeof(double) * n0, kind, stream));
    PB_CUDA(e, cudaMemcpyAsync((void*)d_pid, prim->pid, sizeof(int) * n0, kind, stream));
    PB_CUDA(e, cudaMemcpyAsync((void*)d_fl, prim->flags, sizeof(int) * n0, kind, stream));
  }
  { int rc0 = ensure_order(e, std::max<long long>(2 * n0, 1 << 16), stream); if (rc0 != PB_OK) return rc0; }
  { int rc0 = ensure_work(e, std::max<long long>(2 * n0, 1 << 16)); if (rc0 != PB_OK) return rc0; }
  unsigned long long tail0[3] = {(unsigned long long)n0, 0ull, 0ull};
  PB_CUDA(e, cudaMemcpyAsync(e->work.tail, tail0, sizeof(tail0), cudaMemcpyHostToDevice, stream));
  PB_CUDA(e, cudaMemsetAsync(e->work.ctrl, 0, sizeof(int) * 8, stream));
  PB_CUDA(e, cudaMemsetAsync(e->work.counters, 0, sizeof(unsigned long long) * CNT_N, stream));
  PB_CUDA(e, cudaMemsetAsync(e->work.hist, 0, sizeof(int) * NBUCKET, stream));
  WaveState ws0{};
  ws0.capacity = st->capacity; ws0.work_cap = (int)std::min<long long>(e->work_n, 0x7fffffff);
  ws0.order_cap = (int)std::min<long long>(e->order_cap, 0x7fffffff);
  PB_CUDA(e, cudaMemcpyAsync(e->work.ws, &ws0, sizeof(ws0), cudaMemcpyHostToDevice, stream));
  const int plevel = e->profiling;
  memset(&e->prof, 0, sizeof(e->prof));
  // every-kernel timing (level 2) needs one event pair per launch: it runs one wave per synchronisation; level 1 times
  // k_loop and k_sample only, from a small pool, and keeps the lookahead
  bool recorded[8] = {false, false, false, false, false, false, false, false};
  bool rec_pool[pb_engine_s::LOOKAHEAD][2] = {};
  auto tick = [&](int k, int slot) {
    if (plevel >= 2) { cudaEventRecord(e->ev[2 * k], stream); recorded[k] = true; }
    else if (plevel == 1 && (k == PB_K_PROPAGATE || k == PB_K_SAMPLE)) { cudaEventRecord(e->evp[slot][k == PB_K_SAMPLE][0], stream); rec_pool[slot][k == PB_K_SAMPLE] = true; }
  };
  auto tock = [&](int k, int slot) {
    if (plevel >= 2) cudaEventRecord(e->ev[2 * k + 1], stream);
    else if (plevel == 1 && (k == PB_K_PROPAGATE || k == PB_K_SAMPLE)) cudaEventRecord(e->evp[slot][k == PB_K_SAMPLE][1], stream);
  };
  auto collect = [&]() {     // after a stream synchronise
    for (int k = 0; k < PB_K_N; ++k) {
      if (!recorded[k]) continue;
      float ms = 0.f;
      if (cudaEventElapsedTime(&ms, e->ev[2 * k], e->ev[2 * k + 1]) == cudaSuccess) e->prof.ms[k] += ms;
      recorded[k] = false;
    }
    for (int a = 0; a < pb_engine_s::LOOKAHEAD; ++a) for (int b = 0; b < 2; ++b) {
      if (!rec_pool[a][b]) continue;
      float ms = 0.f;
      if (cudaEventElapsedTime(&ms, e->evp[a][b][0], e->evp[a][b][1]) == cudaSuccess) e->prof.ms[b ? PB_K_SAMPLE : PB_K_PROPAGATE] += ms;
      rec_pool[a][b] = false;
    }
  };
  tick(PB_K_INIT, 0);
  k_init_primaries<<<(unsigned)((n0 + 255) / 256), 256, 0, stream>>>(e->mat, e->tab, S, e->work, d_p, d_r, d_w, e->prim_mass, d_pid, d_fl, n0, seed, first_id, e->prim_mass + e->prim_cap);
  tock(PB_K_INIT, 0);
  ++e->prof.launches[PB_K_INIT];
  ++launches;

  WaveState* hws = e->h_ws;
  unsigned long long* htail = reinterpret_cast<unsigned long long*>(reinterpret_cast<char*>(hws) + sizeof(WaveState));
  long long n_known = n0, n_prev = 0;
  const int ms_flag = global_ms ? 1 : 0;
  const bool graph = e->use_graph && plevel == 0 && !getenv("PB_LOG_WAVES");
  while (graph) {
    // host-free wave loop: the whole shower is one graph launch; the host only comes back if the per-wave scratch has to grow
    int rcg = ensure_wave_graph(e, S, ms_flag);
    if (rcg != PB_OK) return rcg;
    PB_CUDA(e, cudaGraphLaunch(e->wave_exec, stream));
    PB_CUDA(e, cudaMemcpyAsync(hws, e->work.ws, sizeof(WaveState), cudaMemcpyDeviceToHost, stream));
    PB_CUDA(e, cudaMemcpyAsync(htail, e->work.tail, 3 * sizeof(unsigned long long), cudaMemcpyDeviceToHost, stream));
    PB_CUDA(e, cudaStreamSynchronize(stream));
    if (hws->status == 3) {             // scratch too small for the next wave: grow (contents preserved) and resume it
      long long need = (long long)std::min<unsigned long long>(htail[0], (unsigned long long)st->capacity) - hws->begin + (long long)htail[2];
      if (need > 0x3fffffffLL) { e->err = "wave wider than 2^30"; return PB_ERR_CAPACITY; }
      int rc = ensure_work(e, 2 * need);
      if (rc != PB_OK) return rc;
      rc = ensure_order(e, 4 * need, stream);
      if (rc != PB_OK) return rc;
      int caps[2] = {(int)std::min<long long>(e->work_n, 0x7fffffff), (int)std::min<long long>(e->order_cap, 0x7fffffff)};
      PB_CUDA(e, cudaMemcpyAsync(&e->work.ws->work_cap, caps, sizeof(caps), cudaMemcpyHostToDevice, stream));
      continue;
    }
    break;
  }
  if (graph) launches += 7LL * hws->iters;
  while (!graph) {
    // grow phase (or full profiling): one wave per synchronisation; shrinking tail: LOOKAHEAD waves per synchronisation
    const int K = (plevel >= 2 || n_known > n_prev) ? 1 : pb_engine_s::LOOKAHEAD;
    const long long bound = std::max<long long>(K == 1 ? (n_known * 9) / 8 : 2 * n_known, 4096);
    // one thread per entry of the (upper-bounded) wave; the kernels are grid-stride, so a wave that outgrows the bound
    // inside a lookahead batch is still processed completely
    const unsigned g128 = (unsigned)((bound + 127) / 128);
    const unsigned g256 = (unsigned)((bound + 255) / 256);
    const int lg = (int)std::min<long long>((long long)e->n_sm * PB_LOOP_MINB, (bound + 4 * 8 - 1) / (4 * 8));
    const int sg = sample_grid(e, bound, true);
    for (int j = 0; j < K; ++j) {
      k_wave_begin<<<1, 1, 0, stream>>>(e->work);
      tick(PB_K_PROPAGATE, j);
      k_loop<<<lg, 128, 0, stream>>>(e->mat, e->tab, S, e->work, e->prim_mass, ms_flag);
      tock(PB_K_PROPAGATE, j); tick(PB_K_FINALIZE, j);
      k_finalize<<<g128, 128, 0, stream>>>(e->mat, e->tab, S, e->work, e->prim_mass, ms_flag);
      tock(PB_K_FINALIZE, j); tick(PB_K_SCAN, j);
      k_bucket_scan<<<1, 1024, 0, stream>>>(e->work);
      tock(PB_K_SCAN, j); tick(PB_K_FILL, j);
      SampleIO io{S.pf, reinterpret_cast<const uint2*>(S.ids), 4, 2, nullptr, reinterpret_cast<int*>(S.aux), 2, e->work.ws};
      k_bucket_fill<<<(g256 + 3) / 4, 256, 0, stream>>>(e->work, io, -1);          // a CTA ranks 1 024 entries per iteration
      tock(PB_K_FILL, j); tick(PB_K_SAMPLE, j);
      launch_sample(e, bound, io, stream, true);
      tock(PB_K_SAMPLE, j); tick(PB_K_EMIT, j);
      k_emit<<<g128, 128, 0, stream>>>(e->mat, e->tab, S, e->work, e->emit_wave_order, e->prim_mass + e->prim_cap);
      tock(PB_K_EMIT, j);
      launches += 7;
    }
    PB_CUDA(e, cudaMemcpyAsync(hws, e->work.ws, sizeof(WaveState), cudaMemcpyDeviceToHost, stream));
    PB_CUDA(e, cudaMemcpyAsync(htail, e->work.tail, 3 * sizeof(unsigned long long), cudaMemcpyDeviceToHost, stream));
    PB_CUDA(e, cudaStreamSynchronize(stream));
    double ms_before[PB_K_N];
    for (int k = 0; k < PB_K_N; ++k) ms_before[k] = e->prof.ms[k];
    collect();
    if (hws->status == 3) {             // scratch too small for the next wave: grow (contents preserved) and resume it
      long long need = (long long)std::min<unsigned long long>(htail[0], (unsigned long long)st->capacity) - hws->begin + (long long)htail[2];
      if (need > 0x3fffffffLL) { e->err = "wave wider than 2^30"; return PB_ERR_CAPACITY; }
      int rc = ensure_work(e, 2 * need);
      if (rc != PB_OK) return rc;
      rc = ensure_order(e, 4 * need, stream);
      if (rc != PB_OK) return rc;
      int caps[2] = {(int)std::min<long long>(e->work_n, 0x7fffffff), (int)std::min<long long>(e->order_cap, 0x7fffffff)};
      PB_CUDA(e, cudaMemcpyAsync(&e->work.ws->work_cap, caps, sizeof(caps), cudaMemcpyHostToDevice, stream));
      n_prev = 0; n_known = need;
      continue;
    }
    if (hws->status != 0) break;
    if (const char* lw = getenv("PB_LOG_WAVES")) {        // measurement aid: sizes of the waves the host sees
      if (FILE* f = fopen(lw, "a")) {
        fprintf(f, "wave %d n %d n_charged %d n_new %d n_carry %d", hws->waves, hws->n, hws->n_charged, hws->n_new, hws->n_carry);
        if (plevel >= 2) { fprintf(f, " ms"); for (int k = 1; k < PB_K_N; ++k) fprintf(f, " %.4f", e->prof.ms[k] - ms_before[k]); }   // this wave's kernels
        fprintf(f, "\n");
        fclose(f);
      }
    }
    n_prev = n_known;
    n_known = hws->n;
  }
  for (int k = 1; k < PB_K_N; ++k) e->prof.launches[k] = hws->waves;
  if (e->work.tlog && !e->tlog_path.empty()) {
    std::vector<unsigned long long> h(4 * (size_t)e->work.tlog_cap);
    if (cudaMemcpy(h.data(), e->work.tlog, h.size() * 8, cudaMemcpyDeviceToHost) == cudaSuccess)
      if (FILE* f = fopen(e->tlog_path.c_str(), "ab")) { fwrite(h.data(), 8, h.size(), f); fclose(f); }
    cudaMemset(e->work.tlog, 0, h.size() * 8);
  }
  long long end = hws->end, waves = hws->waves, max_wave = hws->max_wave, tot_charged = hws->tot_charged;
  if (hws->status == 2) {
    e->err = "particle stack capacity exhausted (wave " + std::to_string(waves) + ")";
    return PB_ERR_CAPACITY;
  }
  if (hws->status == 4) { e->err = "wave loop did not terminate"; return PB_ERR_STATE; }
  unsigned long long cnt[CNT_N];
  PB_CUDA(e, cudaMemcpyAsync(cnt, e->work.counters, sizeof(cnt), cudaMemcpyDeviceToHost, stream));
  PB_CUDA(e, cudaStreamSynchronize(stream));
  if (out) {
    out->n_particles = end; out->n_waves = waves; out->n_steps = (int64_t)cnt[CNT_STEPS];
    out->n_substeps = (int64_t)cnt[CNT_SUBSTEPS]; out->n_samples = (int64_t)cnt[CNT_SAMPLES];
    out->n_trials = (int64_t)cnt[CNT_TRIALS]; out->n_no_sample = (int64_t)cnt[CNT_NOSAMPLE];
    out->n_launches = launches; out->max_wave = max_wave; out->n_charged = tot_charged;
  }
  for (int p = 0; p < 16; ++p) { e->prof.trials[p] = (int64_t)cnt[CNT_PROC_TRIALS + p]; e->prof.samples[p] = (int64_t)cnt[CNT_PROC_SAMPLES + p]; }
  if (cnt[CNT_OVERFLOW]) { e->err = "particle stack overflow"; return PB_ERR_CAPACITY; }
  if (cnt[CNT_NOSAMPLE]) { e->err = "No Sample Found for " + std::to_string(cnt[CNT_NOSAMPLE]) + " particle(s)"; return PB_ERR_NO_SAMPLE; }
  return PB_OK;
}

static int upload_table(pb_engine e, const double* x, const double* y, int n, NSigmaTable* out) {
  std::vector<double> node(4 * (size_t)std::max(n, 1), 0.0);
  for (int i = 0; i < n; ++i) {
    node[4 * i] = x[i]; node[4 * i + 1] = y[i];
    node[4 * i + 2] = (i + 1 < n) ? (y[i + 1] - y[i]) / (x[i + 1] - x[i]) : 0.0;
  }
  double* d = nullptr;
  PB_CUDA(e, cudaMalloc(&d, sizeof(double) * node.size()));
  e->owned.push_back(d);
  PB_CUDA(e, cudaMemcpy(d, node.data(), sizeof(double) * node.size(), cudaMemcpyHostToDevice));
  *out = NSigmaTable{(const double4*)d, n ? x[0] : 0.0, n ? x[n - 1] : 0.0, Coarse{nullptr, 0, 0}, n, 0};   // dark tables (log10 nodes): binary search
  return PB_OK;
}

extern "C" int pb_upload_dark(pb_engine e, const pb_dark_tables* t) {
  if (!e || !t) return PB_ERR_ARG;
  PB_CUDA(e, cudaSetDevice(e->device));
  for (int k = 0; k < 4; ++k) {
    int rc = upload_table(e, t->w_E[k], t->w_y[k], t->w_n[k], &e->dark.w[k]);
    if (rc != PB_OK) return rc;
    int n = t->d_n[k];
    double* d = nullptr;
    PB_CUDA(e, cudaMalloc(&d, sizeof(double) * (size_t)std::max(n, 1) * 21));
    e->owned.push_back(d);
    PB_CUDA(e, cudaMemcpy(d, t->d_E[k], sizeof(double) * n, cudaMemcpyHostToDevice));
    PB_CUDA(e, cudaMemcpy(d + n, t->d_table[k], sizeof(double) * (size_t)n * 20, cudaMemcpyHostToDevice));
    e->dark.dE[k] = d; e->dark.dT[k] = d + n; e->dark.dn[k] = n;
    e->dark.min_E[k] = t->min_E[k];
  }
  int rc = upload_table(e, t->nsdark_comp_lx, t->nsdark_comp_ly, t->nsdark_comp_n, &e->dark.nsdark_comp);
  if (rc != PB_OK) return rc;
  e->dark_ready = true;
  return PB_OK;
}

extern "C" int pb_run_dark(pb_engine e, const pb_stack* sm, int64_t n_sm, uint32_t active, pb_stack* dk, pb_counters* out,
                           void* stream_) {
  if (!e || !sm || !dk || n_sm < 0 || n_sm > sm->capacity) return PB_ERR_ARG;
  cudaStream_t stream = (cudaStream_t)stream_;
  PB_CUDA(e, cudaSetDevice(e->device));
  if (!e->dark_ready) { e->err = "dark tables not uploaded"; return PB_ERR_STATE; }
  for (int p = P_DARKBREM; p <= P_DARKMUONBREM; ++p)
    if (((active >> p) & 1u) && e->tab.map[p].grid == nullptr) { e->err = "dark maps not uploaded for an active process"; return PB_ERR_STATE; }
  if (n_sm > 0x7fffffffLL) { e->err = "too many SM records for one dark pass"; return PB_ERR_CAPACITY; }
  int B = 300;
  for (int p = P_DARKBREM; p <= P_DARKMUONBREM; ++p) if (e->tab.map[p].grid) { B = e->tab.map[p].B; break; }
  e->mat.max_trials = (long long)std::min<double>((double)e->cfg.max_sweeps * (double)B, 4.0e9);
  Stack S{sm->p0, sm->r0w, sm->pf, sm->rf, (int4*)sm->ids, (int2*)sm->aux, sm->capacity};
  Stack O{dk->p0, dk->r0w, dk->pf, dk->rf, (int4*)dk->ids, (int2*)dk->aux, dk->capacity};
  // The SM records are processed in chunks of at most DARK_CHUNK records, so that the candidate scratch (2 candidates per record
  // in the worst case: 108 B per record) stays bounded whatever the size of the SM stack; dark vectors are appended chunk after chunk.
  const long long DARK_CHUNK = 1LL << 25;
  long long ncap = std::max<long long>(2 * std::min<long long>(n_sm, DARK_CHUNK), 1);
  int rc = ensure_work(e, ncap);
  if (rc != PB_OK) return rc;
  rc = ensure_cand(e, ncap);
  if (rc != PB_OK) return rc;
  const bool prof = e->profiling != 0;
  memset(&e->prof, 0, sizeof(e->prof));
  auto tick = [&](int k) { if (prof) cudaEventRecord(e->ev[2 * k], stream); };
  auto tock = [&](int k) { if (prof) { cudaEventRecord(e->ev[2 * k + 1], stream); } ++e->prof.launches[k]; };
  auto collect = [&](int k) {        // after a stream synchronise
    float ms = 0.f;
    if (prof && cudaEventElapsedTime(&ms, e->ev[2 * k], e->ev[2 * k + 1]) == cudaSuccess) e->prof.ms[k] += ms;
  };
  unsigned long long tail0[2] = {0ull, 0ull};
  PB_CUDA(e, cudaMemcpyAsync(e->work.tail, tail0, sizeof(tail0), cudaMemcpyHostToDevice, stream));
  PB_CUDA(e, cudaMemsetAsync(e->work.counters, 0, sizeof(unsigned long long) * CNT_N, stream));
  PB_CUDA(e, cudaMemsetAsync(e->work.hist, 0, sizeof(int) * NBUCKET, stream));
  long long launches = 0, cand_total = 0, max_cand = 0;
  for (long long first = 0; first < n_sm; first += DARK_CHUNK) {
    const long long last = std::min<long long>(n_sm, first + DARK_CHUNK);
    int n_cand = 0;
    PB_CUDA(e, cudaMemsetAsync(e->work.ctrl, 0, sizeof(int) * 8, stream));
    PB_CUDA(e, cudaMemsetAsync(e->cand.count, 0, sizeof(int), stream));
    tick(PB_K_FINALIZE);
    k_dark_prepare<<<(unsigned)((last - first + 127) / 128), 128, 0, stream>>>(e->mat, e->tab, e->dark, S, e->work, e->cand, first, last, active);
    tock(PB_K_FINALIZE);
    PB_CUDA(e, cudaMemcpyAsync(&n_cand, e->cand.count, sizeof(int), cudaMemcpyDeviceToHost, stream));
    PB_CUDA(e, cudaStreamSynchronize(stream));
    collect(PB_K_FINALIZE);
    ++launches;
    if (n_cand <= 0) continue;
    cand_total += n_cand;
    max_cand = std::max<long long>(max_cand, n_cand);
    if (cand_total > dk->capacity) { e->err = "dark stack capacity exhausted"; return PB_ERR_CAPACITY; }
    tick(PB_K_SCAN);
    k_bucket_scan<<<1, 1024, 0, stream>>>(e->work);
    tock(PB_K_SCAN); tick(PB_K_FILL);
    SampleIO io{e->cand.pf, reinterpret_cast<const uint2*>(S.ids), 4, 2, e->cand.slot, e->cand.ntr, 1, nullptr};
    k_bucket_fill<<<(unsigned)((n_cand + 1023) / 1024), 256, 0, stream>>>(e->work, io, n_cand);
    tock(PB_K_FILL); tick(PB_K_SAMPLE);
    launch_sample(e, n_cand, io, stream);
    tock(PB_K_SAMPLE); tick(PB_K_EMIT);
    k_dark_emit<<<(unsigned)((n_cand + 127) / 128), 128, 0, stream>>>(e->mat, S, O, e->work, e->cand, n_cand);
    tock(PB_K_EMIT);
    launches += 4;
    if (prof || first + DARK_CHUNK < n_sm) {          // the scratch is reused by the next chunk
      PB_CUDA(e, cudaStreamSynchronize(stream));
      collect(PB_K_SCAN); collect(PB_K_FILL); collect(PB_K_SAMPLE); collect(PB_K_EMIT);
    }
  }
  unsigned long long tl[2] = {0, 0};
  unsigned long long cnt[CNT_N];
  PB_CUDA(e, cudaMemcpyAsync(tl, e->work.tail, sizeof(tl), cudaMemcpyDeviceToHost, stream));
  PB_CUDA(e, cudaMemcpyAsync(cnt, e->work.counters, sizeof(cnt), cudaMemcpyDeviceToHost, stream));
  PB_CUDA(e, cudaStreamSynchronize(stream));
  for (int p = 0; p < 16; ++p) { e->prof.trials[p] = (int64_t)cnt[CNT_PROC_TRIALS + p]; e->prof.samples[p] = (int64_t)cnt[CNT_PROC_SAMPLES + p]; }
  if (out) {
    memset(out, 0, sizeof(*out));
    out->n_particles = (int64_t)std::min<unsigned long long>(tl[0], (unsigned long long)dk->capacity);
    out->n_steps = cand_total; out->n_samples = (int64_t)cnt[CNT_SAMPLES]; out->n_trials = (int64_t)cnt[CNT_TRIALS];
    out->n_no_sample = (int64_t)cnt[CNT_NOSAMPLE]; out->n_launches = launches; out->max_wave = max_cand; out->n_waves = 1;
  }
  if (cnt[CNT_OVERFLOW]) { e->err = "dark stack overflow"; return PB_ERR_CAPACITY; }
  return PB_OK;
}

static int ensure_cand(pb_engine e, long long ncap) {
  if (ncap <= e->cand_cap) return PB_OK;
  if (e->cand_blob) cudaFree(e->cand_blob);
  long long cap = ncap * 5 / 4 + 1024;
  PB_CUDA(e, cudaMalloc(&e->cand_blob, (size_t)cap * (4 + 4 + 8 + 32 + 4) + 64));
  char* p = (char*)e->cand_blob;
  e->cand.pf = (double*)p; p += (size_t)cap * 32;
  e->cand.wg = (double*)p; p += (size_t)cap * 8;
  e->cand.slot = (int*)p; p += (size_t)cap * 4;
  e->cand.proc = (int*)p; p += (size_t)cap * 4;
  e->cand.ntr = (int*)p; p += (size_t)cap * 4;
  e->cand.count = (int*)p;
  e->cand_cap = cap;
  return PB_OK;
}

extern "C" int pb_draw_samples(pb_engine e, int process, const double* E, int64_t n, int lu_key, uint64_t seed,
                               uint64_t first_id, double* x_out, int32_t* ntr_out, void* stream_) {
  if (!e || !E || !x_out || n <= 0 || n > 0x3fffffff || process < 0 || process >= N_SAMPLED) return PB_ERR_ARG;
  cudaStream_t stream = (cudaStream_t)stream_;
  PB_CUDA(e, cudaSetDevice(e->device));
  if (e->tab.map[process].grid == nullptr) { e->err = "maps not uploaded for this process"; return PB_ERR_STATE; }
  e->mat.max_trials = (long long)std::min<double>((double)e->cfg.max_sweeps * (double)e->tab.map[process].B, 4.0e9);
  int rc = ensure_work(e, n);
  if (rc != PB_OK) return rc;
  rc = ensure_cand(e, n);
  if (rc != PB_OK) return rc;
  double* dE = nullptr;
  PB_CUDA(e, cudaMalloc(&dE, sizeof(double) * n + sizeof(uint2) * n));
  uint2* dkeys = (uint2*)(dE + n);
  PB_CUDA(e, cudaMemcpyAsync(dE, E, sizeof(double) * n, cudaMemcpyHostToDevice, stream));
  PB_CUDA(e, cudaMemsetAsync(e->work.counters, 0, sizeof(unsigned long long) * CNT_N, stream));
  PB_CUDA(e, cudaMemsetAsync(e->work.hist, 0, sizeof(int) * NBUCKET, stream));
  PB_CUDA(e, cudaMemsetAsync(e->work.ctrl, 0, sizeof(int) * 8, stream));
  PB_CUDA(e, cudaMemsetAsync(e->work.xs, 0, sizeof(double) * 4 * n, stream));
  k_prepare_draws<<<(unsigned)((n + 255) / 256), 256, 0, stream>>>(e->tab, e->work, e->cand, dkeys, dE, (int)n, process, lu_key, seed, first_id);
  k_bucket_scan<<<1, 1024, 0, stream>>>(e->work);
  SampleIO io{e->cand.pf, dkeys, 1, 0, nullptr, e->cand.ntr, 1, nullptr};
  k_bucket_fill<<<(unsigned)((n + 1023) / 1024), 256, 0, stream>>>(e->work, io, (int)n);
  launch_sample(e, n, io, stream);
  cudaError_t c = cudaMemcpyAsync(x_out, e->work.xs, sizeof(double) * 4 * n, cudaMemcpyDeviceToHost, stream);
  if (c == cudaSuccess && ntr_out) c = cudaMemcpyAsync(ntr_out, e->cand.ntr, sizeof(int) * n, cudaMemcpyDeviceToHost, stream);
  if (c == cudaSuccess) c = cudaStreamSynchronize(stream);
  cudaFree(dE);
  if (c != cudaSuccess) { e->err = cudaGetErrorString(c); return PB_ERR_CUDA; }
  return PB_OK;
}

extern "C" int pb_find_max(pb_engine e, int process, int n_trials, uint64_t seed, double mT, double* max_out, double* sum_out) {
  if (!e || !max_out || !sum_out || process < 0 || process >= N_SAMPLED || n_trials < 1) return PB_ERR_ARG;
  PB_CUDA(e, cudaSetDevice(e->device));
  const MapInfo& mi = e->tab.map[process];
  if (mi.grid == nullptr) { e->err = "maps not uploaded for this process"; return PB_ERR_STATE; }
  Material m = e->mat;
  if (mT > 0) m.mT = mT;
  double* d = nullptr;
  PB_CUDA(e, cudaMalloc(&d, sizeof(double) * 2 * mi.nE));
  k_find_max<<<mi.nE, 128, 0, e->aux_stream>>>(m, e->tab, process, n_trials, seed, d, d + mi.nE);
  cudaError_t c = cudaMemcpyAsync(max_out, d, sizeof(double) * mi.nE, cudaMemcpyDeviceToHost, e->aux_stream);
  if (c == cudaSuccess) c = cudaMemcpyAsync(sum_out, d + mi.nE, sizeof(double) * mi.nE, cudaMemcpyDeviceToHost, e->aux_stream);
  if (c == cudaSuccess) c = cudaStreamSynchronize(e->aux_stream);
  cudaFree(d);
  if (c != cudaSuccess) { e->err = cudaGetErrorString(c); return PB_ERR_CUDA; }
  return PB_OK;
}

extern "C" int pb_detector_cut(pb_engine e, const pb_stack* st, int64_t first, int64_t n, const double* z_det, int n_det,
                               double radius, double inner, double E_lo, double E_hi, double* wpass, double* wall,
                               uint8_t* mask, void* stream_) {
  if (!e || !st || !z_det || n_det < 1 || n_det > MAX_DET || n < 0 || first < 0 || first + n > st->capacity) return PB_ERR_ARG;
  PB_CUDA(e, cudaSetDevice(e->device));
  cudaStream_t stream = (cudaStream_t)stream_;
  Stack S{st->p0, st->r0w, st->pf, st->rf, (int4*)st->ids, (int2*)st->aux, st->capacity};
  DetPlanes D{};
  D.n = n_det;
  for (int k = 0; k < n_det; ++k) D.z[k] = z_det[k];
  double* d = nullptr;
  PB_CUDA(e, cudaMalloc(&d, sizeof(double) * (MAX_DET + 1)));
  PB_CUDA(e, cudaMemsetAsync(d, 0, sizeof(double) * (MAX_DET + 1), stream));
  if (n > 0) {
    int grid = (int)std::min<long long>((n + 255) / 256, (long long)e->n_sm * 8);
    k_detector_cut<<<grid, 256, 0, stream>>>(S, first, n, D, radius, inner, E_lo, E_hi, d, mask);
  }
  double h[MAX_DET + 1];
  cudaError_t c = cudaMemcpyAsync(h, d, sizeof(h), cudaMemcpyDeviceToHost, stream);
  if (c == cudaSuccess) c = cudaStreamSynchronize(stream);
  cudaFree(d);
  if (c != cudaSuccess) { e->err = cudaGetErrorString(c); return PB_ERR_CUDA; }
  for (int k = 0; k < n_det; ++k) if (wpass) wpass[k] = h[k];
  if (wall) *wall = h[MAX_DET];
  return PB_OK;
}

extern "C" int pb_train_accumulate_p(pb_engine e, int process, const double* grid, int nE, int dim, const int32_t* ninc,
                                     const double* E_inc, int64_t n_points, uint64_t seed, double mT, double power, double* d_out,
                                     double* n_out, double* integral_out) {
  if (!(power > 0.0 && power <= 64.0)) return PB_ERR_ARG;
  if (!e || !grid || !ninc || !E_inc || !d_out || !n_out || !integral_out || nE < 1 || dim < 1 || dim > 4 || n_points < 1 ||
      process < 0 || process >= N_SAMPLED) return PB_ERR_ARG;
  PB_CUDA(e, cudaSetDevice(e->device));
  int stride = 0;
  int nn[4] = {0, 0, 0, 0};
  for (int d = 0; d < dim; ++d) { nn[d] = ninc[d]; stride += ninc[d] + 1; }
  size_t smem = (size_t)stride * (8 + 8 + 4);
  if (smem > 200 * 1024) { e->err = "map too large for the training kernel"; return PB_ERR_ARG; }
  Material m = e->mat;
  if (mT > 0) m.mT = mT;
  size_t rows = (size_t)nE * stride;
  double* d = nullptr;
  PB_CUDA(e, cudaMalloc(&d, sizeof(double) * (3 * rows + 2 * (size_t)nE)));
  double *d_grid = d, *d_d = d + rows, *d_n = d + 2 * rows, *d_E = d + 3 * rows, *d_I = d_E + nE;
  cudaStream_t st = e->aux_stream;
  cudaError_t c = cudaMemcpyAsync(d_grid, grid, sizeof(double) * rows, cudaMemcpyHostToDevice, st);
  if (c == cudaSuccess) c = cudaMemcpyAsync(d_E, E_inc, sizeof(double) * nE, cudaMemcpyHostToDevice, st);
  if (c == cudaSuccess) c = cudaMemsetAsync(d_d, 0, sizeof(double) * (2 * rows), st);
  if (c == cudaSuccess) c = cudaMemsetAsync(d_I, 0, sizeof(double) * nE, st);
  if (c == cudaSuccess) c = cudaFuncSetAttribute(k_train, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (c == cudaSuccess) {
    int bx = (int)std::max<long long>(1, std::min<long long>(64, n_points / 4096));
    k_train<<<dim3(bx, nE), 256, smem, st>>>(m, process, dim, stride, make_int4(nn[0], nn[1], nn[2], nn[3]), d_grid, d_E, n_points, seed, d_d, d_n, d_I, power);
    c = cudaMemcpyAsync(d_out, d_d, sizeof(double) * rows, cudaMemcpyDeviceToHost, st);
  }
  if (c == cudaSuccess) c = cudaMemcpyAsync(n_out, d_n, sizeof(double) * rows, cudaMemcpyDeviceToHost, st);
  if (c == cudaSuccess) c = cudaMemcpyAsync(integral_out, d_I, sizeof(double) * nE, cudaMemcpyDeviceToHost, st);
  if (c == cudaSuccess) c = cudaStreamSynchronize(st);
  cudaFree(d);
  if (c != cudaSuccess) { e->err = cudaGetErrorString(c); return PB_ERR_CUDA; }
  return PB_OK;
}

extern "C" int pb_train_accumulate(pb_engine e, int process, const double* grid, int nE, int dim, const int32_t* ninc,
                                   const double* E_inc, int64_t n_points, uint64_t seed, double mT, double* d_out,
                                   double* n_out, double* integral_out) {
  return pb_train_accumulate_p(e, process, grid, nE, dim, ninc, E_inc, n_points, seed, mT, 2.0, d_out, n_out, integral_out);
}

extern "C" int pb_set_profiling(pb_engine e, int on) {
  if (!e) return PB_ERR_ARG;
  e->profiling = on < 0 ? 0 : (on > 2 ? 2 : on);
  return PB_OK;
}
extern "C" int pb_get_profile(pb_engine e, pb_profile* out) {
  if (!e || !out) return PB_ERR_ARG;
  *out = e->prof;
  return PB_OK;
}

extern "C" int pb_tally(pb_engine e, const pb_stack* st, int64_t first, int64_t n, double* tally, void* stream_) {
  if (!e || !st || !tally || n < 0 || first < 0 || first + n > st->capacity) return PB_ERR_ARG;
  if (n == 0) return PB_OK;
  PB_CUDA(e, cudaSetDevice(e->device));
  Stack S{st->p0, st->r0w, st->pf, st->rf, (int4*)st->ids, (int2*)st->aux, st->capacity};
  int grid = (int)std::min<long long>((n + 255) / 256, (long long)e->n_sm * 3);
  cudaFuncSetAttribute(k_tally, cudaFuncAttributeMaxDynamicSharedMemorySize, 8 * PB_TALLY_SIZE * (int)sizeof(double));
  k_tally<<<grid, 256, 8 * PB_TALLY_SIZE * sizeof(double), (cudaStream_t)stream_>>>(S, first, n, tally);
  PB_CUDA(e, cudaGetLastError());
  return PB_OK;
}

extern "C" int pb_measure_fp64_peak(pb_engine e, double* tflops) {
  if (!e || !tflops) return PB_ERR_ARG;
  PB_CUDA(e, cudaSetDevice(e->device));
  double* d = nullptr;
  PB_CUDA(e, cudaMalloc(&d, 8));
  cudaEvent_t a, b;
  cudaEventCreate(&a); cudaEventCreate(&b);
  const int iters = 1 << 16, blocks = e->n_sm * 8, threads = 256;
  double best = 0.0;
  for (int rep = 0; rep < 4; ++rep) {
    cudaEventRecord(a, e->aux_stream);
    k_fp64_peak<<<blocks, threads, 0, e->aux_stream>>>(d, iters, 1.0000001, 1e-9);
    cudaEventRecord(b, e->aux_stream);
    cudaEventSynchronize(b);
    float ms = 0.f;
    cudaEventElapsedTime(&ms, a, b);
    double tf = 2.0 * 8.0 * (double)iters * blocks * threads / (ms * 1e-3) / 1e12;
    if (rep > 0) best = std::max(best, tf);
  }
  cudaEventDestroy(a); cudaEventDestroy(b); cudaFree(d);
  *tflops = best;
  return PB_OK;
}

extern "C" int pb_probe(pb_engine e, int what, int process, const double* in, int64_t n, int is, double* out, int os) {
  if (!e || !in || !out || n <= 0) return PB_ERR_ARG;
  PB_CUDA(e, cudaSetDevice(e->device));
  if (e->species_dirty) { int rcs = build_species_tables(e); if (rcs != PB_OK) return rcs; }     // PB_PROBE_SUBSTEP / PROPAGATE read Tables::sp
  double *din = nullptr, *dout = nullptr;
  PB_CUDA(e, cudaMalloc(&din, sizeof(double) * n * is));
  PB_CUDA(e, cudaMalloc(&dout, sizeof(double) * n * os));
  cudaError_t c = cudaMemcpyAsync(din, in, sizeof(double) * n * is, cudaMemcpyHostToDevice, e->aux_stream);
  if (c == cudaSuccess) c = cudaMemsetAsync(dout, 0, sizeof(double) * n * os, e->aux_stream);
  if (c == cudaSuccess) {
    k_probe<<<(unsigned)((n + 127) / 128), 128, 0, e->aux_stream>>>(e->mat, e->tab, what, process, din, n, is, dout, os);
    c = cudaMemcpyAsync(out, dout, sizeof(double) * n * os, cudaMemcpyDeviceToHost, e->aux_stream);
  }
  if (c == cudaSuccess) c = cudaStreamSynchronize(e->aux_stream);
  cudaFree(din); cudaFree(dout);
  if (c != cudaSuccess) { e->err = cudaGetErrorString(c); return PB_ERR_CUDA; }
  return PB_OK;
}

extern "C" int pb_replay(pb_engine e, int64_t n, const double* particles, const double* tape, const int64_t* tape_off, double* out) {
  if (!e || !particles || !tape_off || !out || n <= 0) return PB_ERR_ARG;
  PB_CUDA(e, cudaSetDevice(e->device));
  for (int p = 0; p < 8; ++p)
    if (e->tab.map[p].grid == nullptr || e->tab.ns[p].n == 0) { e->err = "tables not uploaded"; return PB_ERR_STATE; }
  if (e->species_dirty) { int rcs = build_species_tables(e); if (rcs != PB_OK) return rcs; }
  e->mat.max_trials = (long long)std::min<double>((double)e->cfg.max_sweeps * (double)e->tab.map[P_BREM].B, 4.0e9);
  const long long nt = tape_off[n];
  double *din = nullptr, *dtape = nullptr, *dout = nullptr;
  long long* doff = nullptr;
  PB_CUDA(e, cudaMalloc(&din, sizeof(double) * n * REPLAY_IN));
  PB_CUDA(e, cudaMalloc(&dtape, sizeof(double) * std::max<long long>(nt, 1)));
  PB_CUDA(e, cudaMalloc(&doff, sizeof(long long) * (n + 1)));
  PB_CUDA(e, cudaMalloc(&dout, sizeof(double) * n * REPLAY_OUT));
  cudaStream_t st;
  PB_CUDA(e, cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));
  cudaError_t c = cudaMemcpyAsync(din, particles, sizeof(double) * n * REPLAY_IN, cudaMemcpyHostToDevice, st);
  if (c == cudaSuccess && nt > 0) c = cudaMemcpyAsync(dtape, tape, sizeof(double) * nt, cudaMemcpyHostToDevice, st);
  if (c == cudaSuccess) c = cudaMemcpyAsync(doff, tape_off, sizeof(long long) * (n + 1), cudaMemcpyHostToDevice, st);
  if (c == cudaSuccess) {
    k_replay<<<(unsigned)((n + 63) / 64), 64, 0, st>>>(e->mat, e->tab, n, din, dtape, doff, dout);
    c = cudaMemcpyAsync(out, dout, sizeof(double) * n * REPLAY_OUT, cudaMemcpyDeviceToHost, st);
  }
  if (c == cudaSuccess) c = cudaStreamSynchronize(st);
  cudaStreamDestroy(st);
  cudaFree(din); cudaFree(dtape); cudaFree(doff); cudaFree(dout);
  if (c != cudaSuccess) { e->err = cudaGetErrorString(c); return PB_ERR_CUDA; }
  return PB_OK;
}

extern "C" int pb_quad_batch(pb_engine e, int n_tabs, const int32_t* tab_n, const double* const* tab_x, const double* const* tab_y,
                             const double* tab_fill, double dEdx_GeV_per_m, const pb_quad_call* calls, int64_t n, double* result,
                             double* abserr, int32_t* ier) {
  if (!e || n_tabs < 1 || !tab_n || !tab_x || !tab_y || !tab_fill || !calls || !result || n <= 0) return PB_ERR_ARG;
  for (int64_t i = 0; i < n; ++i) {
    const pb_quad_call& c = calls[i];
    bool ok = c.tab >= 0 && c.tab < n_tabs && (c.kind == 0 || c.kind == 1);
    for (int k = 0; k < 3 && ok; ++k) ok = c.surv[k] < n_tabs;
    if (!ok) { e->err = "pb_quad_batch: call " + std::to_string(i) + " refers to a table that does not exist"; return PB_ERR_ARG; }
  }
  PB_CUDA(e, cudaSetDevice(e->device));
  size_t total = 0;
  for (int k = 0; k < n_tabs; ++k) { if (tab_n[k] < 2) { e->err = "pb_quad_batch: tables need two nodes"; return PB_ERR_ARG; } total += (size_t)tab_n[k]; }
  std::vector<double> flat(2 * total);
  std::vector<QuadTab> tabs((size_t)n_tabs);
  double* d_flat = nullptr; QuadTab* d_tabs = nullptr; pb_quad_call* d_calls = nullptr; double* d_out = nullptr; int* d_ier = nullptr;
  cudaStream_t st = nullptr;
  auto cleanup = [&]() { cudaFree(d_flat); cudaFree(d_tabs); cudaFree(d_calls); cudaFree(d_out); cudaFree(d_ier); if (st) cudaStreamDestroy(st); };
#define PB_Q(call) do { cudaError_t _c = (call); if (_c != cudaSuccess) { e->err = std::string(#call) + ": " + cudaGetErrorString(_c); cleanup(); return PB_ERR_CUDA; } } while (0)
  PB_Q(cudaMalloc(&d_flat, sizeof(double) * flat.size()));
  size_t off = 0;
  for (int k = 0; k < n_tabs; ++k) {
    memcpy(&flat[off], tab_x[k], sizeof(double) * tab_n[k]);
    memcpy(&flat[off + tab_n[k]], tab_y[k], sizeof(double) * tab_n[k]);
    tabs[k] = QuadTab{d_flat + off, d_flat + off + tab_n[k], tab_n[k], 0, tab_fill[k]};
    off += 2 * (size_t)tab_n[k];
  }
  PB_Q(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));
  PB_Q(cudaMalloc(&d_tabs, sizeof(QuadTab) * tabs.size()));
  PB_Q(cudaMalloc(&d_calls, sizeof(pb_quad_call) * n));
  PB_Q(cudaMalloc(&d_out, sizeof(double) * 2 * n));
  PB_Q(cudaMalloc(&d_ier, sizeof(int) * n));
  PB_Q(cudaMemcpyAsync(d_flat, flat.data(), sizeof(double) * flat.size(), cudaMemcpyHostToDevice, st));
  PB_Q(cudaMemcpyAsync(d_tabs, tabs.data(), sizeof(QuadTab) * tabs.size(), cudaMemcpyHostToDevice, st));
  PB_Q(cudaMemcpyAsync(d_calls, calls, sizeof(pb_quad_call) * n, cudaMemcpyHostToDevice, st));
  k_quad<<<(unsigned)((n + 63) / 64), 64, 0, st>>>(QuadTabs{d_tabs, dEdx_GeV_per_m}, d_calls, n, d_out, d_out + n, d_ier);
  PB_Q(cudaGetLastError());
  PB_Q(cudaMemcpyAsync(result, d_out, sizeof(double) * n, cudaMemcpyDeviceToHost, st));
  if (abserr) PB_Q(cudaMemcpyAsync(abserr, d_out + n, sizeof(double) * n, cudaMemcpyDeviceToHost, st));
  if (ier) PB_Q(cudaMemcpyAsync(ier, d_ier, sizeof(int) * n, cudaMemcpyDeviceToHost, st));
  PB_Q(cudaStreamSynchronize(st));
#undef PB_Q
  cleanup();
  return PB_OK;
}
