/*
 * bc_oracle.h -- CPU ORACLE for the Blume-Capel checkerboard Metropolis path.
 *
 * THIS IS TEST INFRASTRUCTURE, NOT PRODUCT CODE.  Only tests/, __graft_entry__.smoke()
 * and bench.py's cpu_baseline / --impl reference legs may load it.  The shipped
 * engine (blume_capel_b200/libbc_engine.so) never links or calls anything here.
 *
 * What it restates (reference = /root/reference/main.cpp):
 *   - proposal rule            main.cpp:94
 *   - 4-neighbour sum          main.cpp:97-103 (periodic wrap via dx,dy main.cpp:43-44)
 *   - Blume-Capel energy diff  main.cpp:105-106
 *   - accept rule              main.cpp:109  (de <= 0 || u < exp(-B*de))
 *   - hot-start lattice        main.cpp:166-171
 * What is NEW relative to the reference (the reference is random-sequential with a
 * non-reproducible mt19937, so no trajectory of it can be replayed):
 *   - checkerboard schedule, Philox4x32-10 counter RNG, integer thresholds.
 *   See DESIGN.md "Bit-exact spec".
 *
 * Parity pinning: the reference has NO tests/golden vectors for this path
 * (SURVEY.md section 4).  The oracle is pinned by (i) Random123 Philox4x32-10 known
 * answers, (ii) the 54 thresholds against the literal reference expression,
 * (iii) exact enumeration of L=2,3,4 tori, (iv) statistics of the UNMODIFIED
 * reference built into oracle/_ref (tests/golden/reference_stats.json).
 */
#ifndef BC_ORACLE_H
#define BC_ORACLE_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define BC_ORACLE_PERIODIC 0
#define BC_ORACLE_OPEN 1

/* RNG stream ids (counter word 3, bits 1..3) */
#define BC_STREAM_COARSE 0u
#define BC_STREAM_FINE 1u
#define BC_STREAM_INIT 2u
#define BC_STREAM_BOND 3u

typedef struct {
    int64_t sweep; /* index of the sweep just completed (0-based, global counter) */
    int64_t M;     /* sum s           */
    int64_t Q;     /* sum s^2         */
    int64_t B;     /* sum over bonds s_i s_j (each stored bond once; L=2 periodic counts twice like main.cpp:43-44) */
} bc_oracle_obs;

/* Philox4x32-10 (Random123 / cuRAND constants). */
void bc_oracle_philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]);

/* 54-entry acceptance table, index 18*c + 9*flip + S with c = s+1, S = sum of the four
 * neighbour codes (= nb + 4).  thr31 in [0, 2^31]; accept iff u31 < thr31.
 * Follows main.cpp:94 (proposal), :105-106 (de), :109 (accept). */
void bc_oracle_build_table(double T, double D, double J, uint32_t thr31[54]);

/* The literal reference expression for one table entry (double), used by tests. */
double bc_oracle_reference_de(int cur, int flip, int nb, double D, double J, int* proposal);

/* iid uniform {-1,0,1} hot start (main.cpp:166-171), via Philox stream BC_STREAM_INIT. */
void bc_oracle_init_random(int L, uint64_t seed, uint32_t replica, int8_t* spins);

/* M, Q, B of a configuration. */
void bc_oracle_observables(int L, int boundary, const int8_t* spins, int64_t* M, int64_t* Q, int64_t* B);

/* Run n_sweeps checkerboard sweeps starting at global sweep counter t0.
 * spins: L*L int8 in {-1,0,1}, row-major idx = x + y*L (main.cpp:98-101), updated in place.
 * obs (may be NULL): one record after every sweep t with (t+1) % measure_every == 0.
 * Returns number of obs records written, or <0 on bad arguments. */
int64_t bc_oracle_sweep(int L, int boundary, int8_t* spins, const uint32_t thr31[54], uint64_t seed,
                        uint32_t replica, int64_t t0, int64_t n_sweeps, int64_t measure_every,
                        bc_oracle_obs* obs);

/* Same, restricted to rows [y0, y1) of the lattice (used to test slab decomposition on the CPU):
 * performs ONE half-sweep of the given colour on those rows only. */
void bc_oracle_half_sweep_rows(int L, int boundary, int8_t* spins, const uint32_t thr31[54], uint64_t seed,
                               uint32_t replica, int64_t t, int colour, int y0, int y1);

/* Cluster move (Swendsen-Wang on the +-1 spins; zeros never join, like wolff() main.cpp:113-154):
 * bond between equal non-zero neighbours open with probability p = 1 - exp(-2J/T) (main.cpp:115), every
 * cluster (identified by its smallest site index) flipped with probability 1/2.  Philox streams 3 (bonds) and 4
 * (flips); spec v2, bit-plane bond uniforms: see bond_open() / cluster_flips() in bc_oracle.c and DESIGN.md.
 * NEW relative to the reference (which grows one Wolff cluster at a time from a non-reproducible RNG); same
 * stationary distribution. */
#define BC_STREAM_FLIP 4u
uint32_t bc_oracle_bond_threshold(double T, double J);
void bc_oracle_cluster_update(int L, int boundary, int8_t* spins, uint32_t bond_thr, uint64_t seed, uint32_t replica,
                              int64_t cstep);

/* kind 0 = geometric "spin" clusters (the content of spin/ files, main.cpp:369-370), kind 1 = FK clusters (bond
 * probability p, main.cpp:371-372, bonds at counter 2^62 + mstep). */
void bc_oracle_cluster_moments(int L, int boundary, const int8_t* spins, int kind, uint32_t bond_thr, uint64_t seed,
                               uint32_t replica, int64_t mstep, int64_t out[3]);
/* labels[i] = smallest site index of the cluster of site i, -1 for a zero spin */
void bc_oracle_cluster_labels(int L, int boundary, const int8_t* spins, int kind, uint32_t bond_thr, uint64_t seed,
                              uint32_t replica, int64_t mstep, int32_t* labels);
/* gap-size histogram of one configuration (corner_contribution/gap_size_statistics.cpp:58-81,100-144); ADDS to hist:
 * variant 0: L/2 bins (the reference's, folded gaps + wrap-around gap), variant 1: L-1 bins (open variant) */
void bc_oracle_gap_histogram(int L, int boundary, const int8_t* spins, int kind, uint32_t bond_thr, uint64_t seed,
                             uint32_t replica, int64_t mstep, int variant, int64_t* hist);

/* Counts ties (coarse 15-bit draw equal to coarse threshold) seen since last reset; diagnostics. */
int64_t bc_oracle_tie_count(int reset);

/* Exact enumeration of the L x L torus (L <= 4): out = {E/N, <|m|>, <q>, U4, chi}.
 * Bonds = (+x, +y) per site with wrap (main.cpp:43-44), so L = 2 double-counts. */
int bc_oracle_enumerate(int L, double T, double D, double J, double out[5]);

/* Literal random-sequential Metropolis of main.cpp:89-112 (one attempt per call-equivalent),
 * driven by a 64-bit LCG-free splitmix/xoshiro stream; used for the CPU timing baseline of the
 * reference's per-attempt arithmetic and for statistics cross-checks.  Returns accepted count. */
int64_t bc_oracle_reference_metropolis(int L, int32_t* lattice, double T, double D, double J,
                                       int64_t n_attempts, uint64_t seed);

#ifdef __cplusplus
}
#endif
#endif
