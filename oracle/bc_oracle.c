/*
 * bc_oracle.c -- CPU oracle (TEST INFRASTRUCTURE ONLY; see bc_oracle.h).
 *
 * Plain C restatement of the reference's per-site Metropolis rule
 * (/root/reference/main.cpp:89-112) under the engine's checkerboard schedule and
 * Philox4x32-10 stream.  Written for clarity, one site at a time, no SIMD: it is
 * the thing the CUDA kernels must match bit for bit.
 */
#include "bc_oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>

/* ---------------------------------------------------------------- Philox4x32-10 */
#define PHILOX_M0 0xD2511F53u
#define PHILOX_M1 0xCD9E8D57u
#define PHILOX_W0 0x9E3779B9u
#define PHILOX_W1 0xBB67AE85u

void bc_oracle_philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]) {
    uint32_t c0 = ctr[0], c1 = ctr[1], c2 = ctr[2], c3 = ctr[3];
    uint32_t k0 = key[0], k1 = key[1];
    for (int r = 0; r < 10; ++r) {
        uint64_t p0 = (uint64_t)PHILOX_M0 * c0;
        uint64_t p1 = (uint64_t)PHILOX_M1 * c2;
        uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0;
        uint32_t n1 = (uint32_t)p1;
        uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1;
        uint32_t n3 = (uint32_t)p0;
        c0 = n0; c1 = n1; c2 = n2; c3 = n3;
        k0 += PHILOX_W0; k1 += PHILOX_W1;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

static void site_key(uint64_t seed, uint32_t replica, uint32_t key[2]) {
    key[0] = (uint32_t)seed;
    key[1] = (uint32_t)(seed >> 32) ^ replica;
}

static void site_ctr(uint32_t g, uint32_t y, int64_t t, uint32_t stream, int colour, uint32_t ctr[4]) {
    ctr[0] = g;
    ctr[1] = y;
    ctr[2] = (uint32_t)((uint64_t)t & 0xFFFFFFFFu);
    ctr[3] = ((uint32_t)((uint64_t)t >> 32) << 4) | (stream << 1) | (uint32_t)(colour & 1);
}

/* 16-bit draw for same-colour site number k of row y: Philox call per 8 sites (g = k>>3),
 * halfword j = k&7 -> word j>>1, low half if j even, high half if j odd. */
static uint32_t draw16(uint64_t seed, uint32_t replica, uint32_t k, uint32_t y, int64_t t, uint32_t stream,
                       int colour) {
    uint32_t key[2], ctr[4], w[4];
    site_key(seed, replica, key);
    site_ctr(k >> 3, y, t, stream, colour, ctr);
    bc_oracle_philox4x32_10(ctr, key, w);
    uint32_t j = k & 7u;
    uint32_t word = w[j >> 1];
    return (j & 1u) ? (word >> 16) : (word & 0xFFFFu);
}

/* ---------------------------------------------------------------- acceptance table */
double bc_oracle_reference_de(int cur, int flip, int nb, double D, double J, int* proposal) {
    /* main.cpp:94 */
    const int prop = (cur + 2 + flip) % 3 - 1;
    if (proposal) *proposal = prop;
    /* main.cpp:105-106 */
    const double de = -J * (prop - cur) * nb + D * (prop * prop - cur * cur);
    return de;
}

void bc_oracle_build_table(double T, double D, double J, uint32_t thr31[54]) {
    const double B = 1 / T; /* main.cpp:321 */
    for (int c = 0; c < 3; ++c)
        for (int f = 0; f < 2; ++f)
            for (int S = 0; S < 9; ++S) {
                const double de = bc_oracle_reference_de(c - 1, f, S - 4, D, J, NULL);
                uint32_t thr;
                if (de <= 0) { /* main.cpp:109 short-circuit: always accept */
                    thr = 0x80000000u;
                } else {
                    double p = exp(-B * de) * 2147483648.0;
                    long long q = llround(p);
                    if (q < 0) q = 0;
                    if (q > 2147483648LL) q = 2147483648LL;
                    thr = (uint32_t)q;
                }
                thr31[18 * c + 9 * f + S] = thr;
            }
}

/* ---------------------------------------------------------------- init / observables */
void bc_oracle_init_random(int L, uint64_t seed, uint32_t replica, int8_t* spins) {
    uint32_t key[2];
    site_key(seed, replica, key);
    const int64_t N = (int64_t)L * L;
    for (int64_t i = 0; i < N; ++i) {
        uint32_t ctr[4] = {(uint32_t)(i >> 2), (uint32_t)((uint64_t)i >> 34), 0u, BC_STREAM_INIT << 1};
        uint32_t w[4];
        bc_oracle_philox4x32_10(ctr, key, w);
        uint32_t r = w[i & 3];
        spins[i] = (int8_t)((int)(((uint64_t)r * 3u) >> 32) - 1);
    }
}

void bc_oracle_observables(int L, int boundary, const int8_t* spins, int64_t* M, int64_t* Q, int64_t* B) {
    int64_t m = 0, q = 0, b = 0;
    for (int y = 0; y < L; ++y)
        for (int x = 0; x < L; ++x) {
            const int s = spins[x + (int64_t)y * L];
            m += s;
            q += s * s;
            int r = 0, d = 0;
            if (x + 1 < L) r = spins[(x + 1) + (int64_t)y * L];
            else if (boundary == BC_ORACLE_PERIODIC) r = spins[0 + (int64_t)y * L];
            if (y + 1 < L) d = spins[x + (int64_t)(y + 1) * L];
            else if (boundary == BC_ORACLE_PERIODIC) d = spins[x];
            b += s * (r + d);
        }
    if (M) *M = m;
    if (Q) *Q = q;
    if (B) *B = b;
}

/* ---------------------------------------------------------------- the sweep */
static int64_t g_ties = 0;
int64_t bc_oracle_tie_count(int reset) {
    int64_t v = g_ties;
    if (reset) g_ties = 0;
    return v;
}

static int nb_code_sum(int L, int boundary, const int8_t* spins, int x, int y) {
    /* main.cpp:97-103: sum over the 4 neighbours; here in code space c = s+1 so S = nb + 4.
     * Open boundary: an off-lattice neighbour contributes spin 0 (code 1). */
    int S = 0;
    const int dxs[4] = {1, -1, 0, 0}, dys[4] = {0, 0, 1, -1};
    for (int d = 0; d < 4; ++d) {
        int nx = x + dxs[d], ny = y + dys[d];
        if (boundary == BC_ORACLE_PERIODIC) {
            nx = (nx + L) % L;
            ny = (ny + L) % L;
            S += spins[nx + (int64_t)ny * L] + 1;
        } else {
            if (nx < 0 || nx >= L || ny < 0 || ny >= L) S += 1;
            else S += spins[nx + (int64_t)ny * L] + 1;
        }
    }
    return S;
}

void bc_oracle_half_sweep_rows(int L, int boundary, int8_t* spins, const uint32_t thr31[54], uint64_t seed,
                               uint32_t replica, int64_t t, int colour, int y0, int y1) {
    uint32_t key[2];
    site_key(seed, replica, key);
    for (int y = y0; y < y1; ++y) {
        const int p = (y + colour) & 1;
        uint32_t k = 0;
        uint32_t w[4] = {0, 0, 0, 0};
        for (int x = p; x < L; x += 2, ++k) {
            const int64_t idx = x + (int64_t)y * L;
            const int c = spins[idx] + 1;
            /* coarse 16-bit draw: one Philox call per 8 same-colour sites of the row (== draw16) */
            const uint32_t j = k & 7u;
            if (j == 0) {
                uint32_t ctr[4];
                site_ctr(k >> 3, (uint32_t)y, t, BC_STREAM_COARSE, colour, ctr);
                bc_oracle_philox4x32_10(ctr, key, w);
            }
            const uint32_t h = (j & 1u) ? (w[j >> 1] >> 16) : (w[j >> 1] & 0xFFFFu);
            const int flip = (int)(h >> 15);
            const uint32_t u15 = h & 0x7FFFu;
            const int S = nb_code_sum(L, boundary, spins, x, y);
            const uint32_t thr = thr31[18 * c + 9 * flip + S];
            const uint32_t thr_c = thr >> 16;
            int accept;
            if (u15 < thr_c) accept = 1;
            else if (u15 > thr_c) accept = 0;
            else {
                /* tie on the 15 coarse bits: resolve with 16 fine bits; u31 = u15<<16 | f16 < thr31 */
                const uint32_t f16 = draw16(seed, replica, k, (uint32_t)y, t, BC_STREAM_FINE, colour);
                accept = f16 < (thr & 0xFFFFu);
                ++g_ties;
            }
            if (accept) spins[idx] = (int8_t)((c + 1 + flip) % 3 - 1); /* main.cpp:94 in code space */
        }
    }
}

int64_t bc_oracle_sweep(int L, int boundary, int8_t* spins, const uint32_t thr31[54], uint64_t seed,
                        uint32_t replica, int64_t t0, int64_t n_sweeps, int64_t measure_every,
                        bc_oracle_obs* obs) {
    if (L < 1 || n_sweeps < 0) return -1;
    if (boundary == BC_ORACLE_PERIODIC && (L & 1)) return -2; /* checkerboard needs even L on a torus */
    int64_t n_obs = 0;
    for (int64_t t = t0; t < t0 + n_sweeps; ++t) {
        bc_oracle_half_sweep_rows(L, boundary, spins, thr31, seed, replica, t, 0, 0, L);
        bc_oracle_half_sweep_rows(L, boundary, spins, thr31, seed, replica, t, 1, 0, L);
        if (obs && measure_every > 0 && ((t + 1) % measure_every) == 0) {
            obs[n_obs].sweep = t;
            bc_oracle_observables(L, boundary, spins, &obs[n_obs].M, &obs[n_obs].Q, &obs[n_obs].B);
            ++n_obs;
        }
    }
    return n_obs;
}

/* ---------------------------------------------------------------- cluster move */
uint32_t bc_oracle_bond_threshold(double T, double J) {
    const double p = 1 - exp(-2 * (1 / T) * J); /* main.cpp:115 */
    if (!(p > 0)) return 0u;
    double q = floor(p * 4294967296.0 + 0.5);
    if (q > 4294967295.0) q = 4294967295.0;
    return (uint32_t)q;
}

static int32_t uf_find(int32_t* parent, int32_t a) {
    while (parent[a] != a) { parent[a] = parent[parent[a]]; a = parent[a]; }
    return a;
}

/* Bond uniform of site (x, y) in direction dir (0 = +x, 1 = +y) at cluster counter cstep, spec v2: the 32 sites
 * x = 32g .. 32g+31 of a row share 32 "bit-plane" words; plane b (b = 0 is the most significant bit) is word b&3 of
 * Philox(counter = (y << 15 | g << 4 | dir << 3 | b >> 2, cstep_lo, cstep_hi, BC_STREAM_BOND << 1)), and
 * u(x) = sum_b bit_(x&31)(plane b) << (31 - b).  The bond is open iff u < thr.  (A GPU thread evaluates 32 bonds at
 * once, most significant plane first, and stops drawing planes when every bond is decided: about two Philox calls
 * per 32 bonds instead of sixteen.)  Evaluated here exactly like that definition, one bond at a time. */
static int bond_open(const uint32_t key[2], int x, int y, int dir, int64_t cstep, uint32_t thr) {
    const uint32_t c_lo = (uint32_t)((uint64_t)cstep & 0xFFFFFFFFu), c_hi = (uint32_t)((uint64_t)cstep >> 32);
    uint32_t w[4] = {0, 0, 0, 0};
    for (int b = 0; b < 32; ++b) {
        if ((b & 3) == 0) {
            uint32_t ctr[4] = {((uint32_t)y << 15) | ((uint32_t)(x >> 5) << 4) | ((uint32_t)dir << 3) | (uint32_t)(b >> 2),
                               c_lo, c_hi, BC_STREAM_BOND << 1};
            bc_oracle_philox4x32_10(ctr, key, w);
        }
        const uint32_t ubit = (w[b & 3] >> (x & 31)) & 1u, tbit = (thr >> (31 - b)) & 1u;
        if (ubit != tbit) return ubit < tbit;
    }
    return 0; /* u == thr */
}

/* flip decision of the cluster whose smallest site index is `root`: bit root&31 of word (root>>5)&3 of
 * Philox(counter = (root >> 7, cstep_lo, cstep_hi, BC_STREAM_FLIP << 1)) */
static int cluster_flips(const uint32_t key[2], uint32_t root, int64_t cstep) {
    uint32_t ctr[4] = {root >> 7, (uint32_t)((uint64_t)cstep & 0xFFFFFFFFu), (uint32_t)((uint64_t)cstep >> 32),
                       BC_STREAM_FLIP << 1}, w[4];
    bc_oracle_philox4x32_10(ctr, key, w);
    return (int)((w[(root >> 5) & 3u] >> (root & 31u)) & 1u);
}

/* union-find over the open bonds; parent[] ends up holding the smallest site index of each cluster */
static int32_t* label_clusters(int L, int boundary, const int8_t* spins, int always, uint32_t bond_thr, uint64_t seed,
                               uint32_t replica, int64_t cstep) {
    const int64_t N = (int64_t)L * L;
    int32_t* parent = (int32_t*)malloc(sizeof(int32_t) * (size_t)N);
    uint32_t key[2];
    site_key(seed, replica, key);
    for (int64_t i = 0; i < N; ++i) parent[i] = (int32_t)i;
    for (int y = 0; y < L; ++y)
        for (int x = 0; x < L; ++x) {
            const int64_t i = x + (int64_t)y * L;
            const int8_t s = spins[i];
            if (s == 0) continue; /* zeros never join a cluster (main.cpp:123-130,144) */
            for (int dir = 0; dir < 2; ++dir) {
                int nx = x + (dir == 0), ny = y + (dir == 1);
                if (nx == L || ny == L) {
                    if (boundary != BC_ORACLE_PERIODIC) continue;
                    nx %= L; ny %= L;
                }
                const int64_t j = nx + (int64_t)ny * L;
                if (j == i || spins[j] != s) continue;
                if (always || bond_open(key, x, y, dir, cstep, bond_thr)) {
                    int32_t a = uf_find(parent, (int32_t)i), b = uf_find(parent, (int32_t)j);
                    if (a != b) { if (a < b) parent[b] = a; else parent[a] = b; } /* root = smallest site */
                }
            }
        }
    for (int64_t i = 0; i < N; ++i) parent[i] = uf_find(parent, (int32_t)i);
    return parent;
}

void bc_oracle_cluster_update(int L, int boundary, int8_t* spins, uint32_t bond_thr, uint64_t seed, uint32_t replica,
                              int64_t cstep) {
    const int64_t N = (int64_t)L * L;
    int32_t* parent = label_clusters(L, boundary, spins, 0, bond_thr, seed, replica, cstep);
    uint32_t key[2];
    site_key(seed, replica, key);
    for (int64_t i = 0; i < N; ++i) {
        if (spins[i] == 0) continue;
        if (cluster_flips(key, (uint32_t)parent[i], cstep)) spins[i] = (int8_t)-spins[i];
    }
    free(parent);
}

void bc_oracle_cluster_labels(int L, int boundary, const int8_t* spins, int kind, uint32_t bond_thr, uint64_t seed,
                              uint32_t replica, int64_t mstep, int32_t* labels) {
    const int64_t N = (int64_t)L * L;
    int32_t* parent = label_clusters(L, boundary, spins, kind == 0, bond_thr, seed, replica, ((int64_t)1 << 62) + mstep);
    for (int64_t i = 0; i < N; ++i) labels[i] = spins[i] ? parent[i] : -1;
    free(parent);
}

/* Gap-size histogram of the clusters of one configuration, what corner_contribution/gap_size_statistics.cpp derives
 * from one cluster file (get_sample_gap_sizes, :128-144), restated on the labelled lattice.  A cluster file line is
 * the cluster's sites ascending (main.cpp:236-256); the tool walks it as follows, quirks included:
 *   - the walk starts at posn = first_site / L -- the ROW of the first site, not its column (:104) -- and the first
 *     site itself is never recorded (:106 starts at i = 1);
 *   - posn += delta; when posn >= L the collected "row" is closed (kept only if it holds more than one entry) and
 *     posn %= L (:108-118).  Equivalently: the sites s of the cluster are shifted to s' = s - first + first / L and
 *     grouped by s' / L, with in-row position s' % L;
 *   - per kept row: the gaps between consecutive entries, each folded to min(gap, L - gap), plus the wrap-around gap
 *     last - first; a row of exactly two entries contributes its one gap only (:58-81); bin gap - 1.
 * variant 0 = the reference's (periodic: hist has L/2 bins); variant 1 = open-boundary variant, new relative to the
 * reference (no folding and no wrap-around gap: hist has L - 1 bins). */
static void close_row(int64_t* hist, const int64_t* row, int64_t n, int L, int variant) {
    if (n < 2) return;
    if (variant == 1) {
        for (int64_t i = 0; i + 1 < n; ++i) ++hist[row[i + 1] - row[i] - 1];
        return;
    }
    for (int64_t i = 0; i + 1 < n; ++i) {
        int64_t gap = row[i + 1] - row[i];
        if (L - gap < gap) gap = L - gap;
        ++hist[gap - 1];
    }
    if (n > 2) {
        int64_t gap = row[n - 1] - row[0];
        if (L - gap < gap) gap = L - gap;
        ++hist[gap - 1];
    }
}

void bc_oracle_gap_histogram(int L, int boundary, const int8_t* spins, int kind, uint32_t bond_thr, uint64_t seed,
                             uint32_t replica, int64_t mstep, int variant, int64_t* hist) {
    const int64_t N = (int64_t)L * L;
    int32_t* parent = label_clusters(L, boundary, spins, kind == 0, bond_thr, seed, replica, ((int64_t)1 << 62) + mstep);
    /* members of every cluster in ascending order: counting sort by root (roots are smallest sites) */
    int64_t* start = (int64_t*)calloc((size_t)N + 1, sizeof(int64_t));
    for (int64_t i = 0; i < N; ++i) if (spins[i]) ++start[parent[i] + 1];
    for (int64_t i = 0; i < N; ++i) start[i + 1] += start[i];
    int64_t* fill = (int64_t*)malloc(sizeof(int64_t) * (size_t)N);
    int32_t* member = (int32_t*)malloc(sizeof(int32_t) * (size_t)(start[N] > 0 ? start[N] : 1));
    memcpy(fill, start, sizeof(int64_t) * (size_t)N);
    for (int64_t i = 0; i < N; ++i) if (spins[i]) member[fill[parent[i]]++] = (int32_t)i;
    int64_t* row = (int64_t*)malloc(sizeof(int64_t) * (size_t)L);
    for (int64_t r = 0; r < N; ++r) {
        const int64_t n = start[r + 1] - start[r];
        if (n == 0) continue;
        const int32_t* m = member + start[r];
        int64_t posn = m[0] / L, in_row = 0; /* gap_size_statistics.cpp:104 */
        for (int64_t i = 1; i < n; ++i) {
            posn += m[i] - m[i - 1];
            if (posn >= L) {
                close_row(hist, row, in_row, L, variant);
                in_row = 0;
                posn %= L;
            }
            row[in_row++] = posn;
        }
        close_row(hist, row, in_row, L, variant);
    }
    free(row); free(member); free(fill); free(start); free(parent);
}

/* sum size^2, max size, number of clusters (what gamma_nu.cpp:35-60 derives from a cluster file);
 * kind 0 = spin clusters (every equal-sign bond), kind 1 = FK clusters (bonds at counter 2^62 + mstep) */
void bc_oracle_cluster_moments(int L, int boundary, const int8_t* spins, int kind, uint32_t bond_thr, uint64_t seed,
                               uint32_t replica, int64_t mstep, int64_t out[3]) {
    const int64_t N = (int64_t)L * L;
    int32_t* parent = label_clusters(L, boundary, spins, kind == 0, bond_thr, seed, replica, ((int64_t)1 << 62) + mstep);
    int32_t* size = (int32_t*)calloc((size_t)N, sizeof(int32_t));
    for (int64_t i = 0; i < N; ++i) if (spins[i] != 0) ++size[parent[i]];
    out[0] = out[1] = out[2] = 0;
    for (int64_t i = 0; i < N; ++i) if (size[i]) {
        out[0] += (int64_t)size[i] * size[i];
        if (size[i] > out[1]) out[1] = size[i];
        ++out[2];
    }
    free(size);
    free(parent);
}

/* ---------------------------------------------------------------- exact enumeration */
int bc_oracle_enumerate(int L, double T, double D, double J, double out[5]) {
    if (L < 1 || L > 4) return -1;
    const int N = L * L;
    int8_t s[16];
    memset(s, 0, sizeof(s));
    for (int i = 0; i < N; ++i) s[i] = -1;
    double Z = 0, sE = 0, sAm = 0, sQ = 0, sM2 = 0, sM4 = 0;
    const double beta = 1.0 / T;
    /* energies are shifted by a running minimum Emin; accumulators are rescaled when it drops */
    double Emin = 1e300;
    for (;;) {
        int64_t M, Q, B;
        bc_oracle_observables(L, BC_ORACLE_PERIODIC, s, &M, &Q, &B);
        const double E = -J * (double)B + D * (double)Q;
        if (E < Emin) {
            const double f = (Emin > 1e299) ? 0.0 : exp(-beta * (Emin - E));
            Z *= f; sE *= f; sAm *= f; sQ *= f; sM2 *= f; sM4 *= f;
            Emin = E;
        }
        const double w = exp(-beta * (E - Emin));
        const double m = (double)M / N;
        Z += w;
        sE += w * E;
        sAm += w * fabs(m);
        sQ += w * (double)Q / N;
        sM2 += w * m * m;
        sM4 += w * m * m * m * m;
        /* next state (base-3 counter over {-1,0,1}) */
        int i = 0;
        while (i < N) {
            if (s[i] < 1) { ++s[i]; break; }
            s[i] = -1;
            ++i;
        }
        if (i == N) break;
    }
    out[0] = sE / Z / N;
    out[1] = sAm / Z;
    out[2] = sQ / Z;
    out[3] = 1.0 - (sM4 / Z) / (3.0 * (sM2 / Z) * (sM2 / Z));
    out[4] = N * (sM2 / Z - (sAm / Z) * (sAm / Z));
    return 0;
}

/* ---------------------------------------------------------------- literal reference attempt loop */
static inline uint64_t splitmix64(uint64_t* x) {
    uint64_t z = (*x += 0x9E3779B97F4A7C15ull);
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}

int64_t bc_oracle_reference_metropolis(int L, int32_t* lattice, double T, double D, double J,
                                       int64_t n_attempts, uint64_t seed) {
    /* main.cpp:89-112, one attempt per iteration: random site, 2-way proposal, periodic
     * neighbour sum through a modL table, double-precision dE, exp() only on uphill moves. */
    const int N = L * L;
    const double B = 1 / T;
    int* modL = (int*)malloc(sizeof(int) * (size_t)(2 * L));
    for (int i = 0; i < 2 * L; ++i) modL[i] = i % L;
    const int dx[4] = {1, L - 1, 0, 0}, dy[4] = {0, 0, 1, L - 1};
    uint64_t st = seed;
    int64_t accepted = 0;
    for (int64_t a = 0; a < n_attempts; ++a) {
        const uint64_t r = splitmix64(&st);
        const int posn = (int)(((r & 0xFFFFFFFFull) * (uint64_t)N) >> 32);
        const int flip = (int)((r >> 32) & 1);
        const int current = lattice[posn];
        const int proposal = (current + 2 + flip) % 3 - 1;
        int neighbors = 0;
        const int x = modL[posn % L + 0] , y = posn / L;
        for (int d = 0; d < 4; ++d) neighbors += lattice[modL[x + dx[d]] + modL[y + dy[d]] * L];
        const double de = -J * (proposal - current) * neighbors + D * (proposal * proposal - current * current);
        if (de <= 0 || (double)(splitmix64(&st) >> 11) * (1.0 / 9007199254740992.0) < exp(-B * de)) {
            lattice[posn] = proposal;
            ++accepted;
        }
    }
    free(modL);
    return accepted;
}
