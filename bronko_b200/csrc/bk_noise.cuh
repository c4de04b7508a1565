// bk_noise.cuh — the noise baseline of the reference (get_baseline_noise, src/call.rs:799-967; quirks: SURVEY.md
// Appendix C, Q12) as a parallel computation with bit-identical results.
//
// The reference is ONE sequential loop over i in [0, len+50) carrying three pieces of state:
//   s, s2   running FP64 sums of the minor-allele fractions (and their squares) of the last 100 positions,
//   maxes   a 10-entry "max table" with an evict-by-value quirk,
// followed per iteration by the modified Thompson-tau rejection loop, which only READS that state.
// Everything is split so that only genuinely sequential work stays sequential:
//
//   k_noise_fracs   one thread per position: sorted allele fractions (order-free).
//   k_noise_seq     block 0 / 1 of every sequence: the s / s2 chains, replicated EXACTLY by integer prefix sums
//                   (see "chains" below); blocks >= 2: the max table, one lane per chunk of 128 iterations,
//                   started SPECULATIVELY 256 iterations early from the ten largest fractions of the window.
//   k_noise_fix     verifies every speculative chunk against the true state at its boundary (the table after
//                   iteration i depends only on the table after i-1 and on the data, so equal states at one
//                   iteration prove everything after it) and replays from the true state where the warm-up
//                   had not converged yet, until it meets the speculative trajectory.
//   k_noise_tau     one thread per output position: n (a count), then the Thompson-tau loop on the snapshots.
//
// chains.  While s stays inside one binade [2^e, 2^(e+1)) its ulp u = 2^(e-52) is fixed; with S = s/u (an integer
// in [2^52, 2^53)) and X = x/u:  fl(s + x) = (S + RN(X))·u  as long as the result stays strictly inside the binade,
// and RN(X) depends on S only through its parity (exact ties X = a + 1/2 round to even).  Every operation is thus
// a map "parity → integer increment", those maps compose associatively, and 1536 consecutive operations (256
// iterations, one per thread) are one block-wide scan.  The first operation whose result leaves the binade (or whose
// operand is as large as the sum / subnormal) ends the accepted prefix and is executed in real FP64; stretches
// where that keeps happening (very sparse coverage) are executed serially like the reference.
//
// Written once for nvcc (kernels below) and g++ (tests/emul steps the same primitives on the CPU; tests only).
#pragma once
#include "bk_core.cuh"

namespace bk {

typedef long long i64;

#define BK_NOISE_WINDOW 100
#define BK_NOISE_HALF 50
#define BK_NOISE_TABLE 10
#define BK_NZ_CHUNK 128          // iterations per speculative table chunk
#define BK_NZ_WARM 256           // warm-up iterations in front of a chunk
#define BK_NZ_ROUND 256          // iterations per chain round (= threads of the chain block)
#define BK_NZ_TILE 1024          // iterations of fractions staged per shared-memory tile of a chain block
#ifndef BK_NZ_SERIAL
#define BK_NZ_SERIAL 8           // iterations executed serially per batch (after a stop, and for as long as the exponent keeps moving)
#endif
#define BK_NZ_PAD_LO 100         // zero positions in front of every sequence in the fraction array
#define BK_NZ_PAD (BK_NZ_PAD_LO + 150)   // total padding positions per sequence
#define BK_NZ_MASK52 0xFFFFFFFFFFFFFull

#if defined(__CUDACC__)
BK_HD double nz_d(u64 b) { return __longlong_as_double((long long)b); }
BK_HD u64 nz_b(double d) { return (u64)__double_as_longlong(d); }
BK_HD double nz_add(double a, double b) { return __dadd_rn(a, b); }
BK_HD double nz_sub(double a, double b) { return __dsub_rn(a, b); }
BK_HD double nz_mul(double a, double b) { return __dmul_rn(a, b); }
BK_HD double nz_div(double a, double b) { return __ddiv_rn(a, b); }
BK_HD double nz_sqrt(double a) { return __dsqrt_rn(a); }
BK_HD double nz_abs(double a) { return fabs(a); }
#else
}  // namespace bk
#include <cmath>
#include <cstring>
namespace bk {
BK_HD double nz_d(u64 b) { double d; memcpy(&d, &b, 8); return d; }
BK_HD u64 nz_b(double d) { u64 b; memcpy(&b, &d, 8); return b; }
BK_HD double nz_add(double a, double b) { return a + b; }      // built with -ffp-contract=off
BK_HD double nz_sub(double a, double b) { return a - b; }
BK_HD double nz_mul(double a, double b) { return a * b; }
BK_HD double nz_div(double a, double b) { return a / b; }
BK_HD double nz_sqrt(double a) { return std::sqrt(a); }
BK_HD double nz_abs(double a) { return std::fabs(a); }
#endif

// ---- fractions: src/call.rs:829-842.  counts sorted descending, "minor" = ranks 2..4 ------------------------
BK_HD void nz_fractions(const u32* f4, const u32* r4, double* m3) {
    u64 c0 = (u64)f4[0] + r4[0], c1 = (u64)f4[1] + r4[1], c2 = (u64)f4[2] + r4[2], c3 = (u64)f4[3] + r4[3];
    u64 t;
#define BK_CSWAP(a, b) if (a < b) { t = a; a = b; b = t; }
    BK_CSWAP(c0, c1) BK_CSWAP(c2, c3) BK_CSWAP(c0, c2) BK_CSWAP(c1, c3) BK_CSWAP(c1, c2)
#undef BK_CSWAP
    const u64 total = c0 + c1 + c2 + c3;
    m3[0] = m3[1] = m3[2] = 0.0;
    if (total != 0) {
        const double td = (double)total;
        m3[0] = nz_div((double)c1, td); m3[1] = nz_div((double)c2, td); m3[2] = nz_div((double)c3, td);
    }
}

// ---- the max table: src/call.rs:857-890 --------------------------------------------------------------------
// Entries are bit patterns of non-negative doubles (they order like their bit patterns), non-increasing, 0 = empty.
struct NzTable { u64 t[BK_NOISE_TABLE]; u64 skip_below; };
// skip_below: a leaving value whose bit pattern is below this (and above NZ_TINY) cannot be within 1e-12 of any entry —
// the smallest non-zero entry minus 4e-12 (everything if the table is empty, nothing if that entry is tiny itself)
#define BK_NZ_TINY_BITS 0x3D919799812DEA11ull      /* 4e-12 */

BK_HD void nz_table_relo(NzTable& T) {
    u64 lo = ~0ull;
#pragma unroll
    for (int q = 0; q < BK_NOISE_TABLE; q++) if (T.t[q] != 0) lo = T.t[q];     // non-increasing: the last non-zero
    if (lo == ~0ull) T.skip_below = ~0ull;
    else { const double d = nz_sub(nz_d(lo), 4e-12); T.skip_below = d > 4e-12 ? nz_b(d) : 0ull; }
}
BK_HD void nz_table_clear(NzTable& T) {
#pragma unroll
    for (int q = 0; q < BK_NOISE_TABLE; q++) T.t[q] = 0;
    T.skip_below = ~0ull;
}
// insert (src/call.rs:872-890): bubbles up while strictly greater → lands behind every entry >= nw
BK_HD void nz_table_insert(NzTable& T, u64 nwb) {
    if (!(nwb > T.t[BK_NOISE_TABLE - 1])) return;           // also rejects nw == 0
    u32 p = 0;
#pragma unroll
    for (int q = 0; q < BK_NOISE_TABLE; q++) p += (T.t[q] >= nwb) ? 1u : 0u;
#pragma unroll
    for (int q = BK_NOISE_TABLE - 1; q >= 1; q--) T.t[q] = ((u32)q > p) ? T.t[q - 1] : ((u32)q == p ? nwb : T.t[q]);
    if (p == 0) T.t[0] = nwb;
    nz_table_relo(T);
}
// evict (src/call.rs:857-869): the FIRST entry within 1e-12 of the leaving value is removed, nothing refills
BK_HD void nz_table_evict(NzTable& T, u64 oldb) {
    if (oldb == 0) return;                                  // fractions are never negative: old > 0.0 <=> bits != 0
    if (oldb > BK_NZ_TINY_BITS && oldb < T.skip_below) return;
    const double old = nz_d(oldb);
    u32 pos = BK_NOISE_TABLE;
#pragma unroll
    for (int q = BK_NOISE_TABLE - 1; q >= 0; q--) if (nz_abs(nz_sub(nz_d(T.t[q]), old)) < 1e-12) pos = (u32)q;
    if (pos == BK_NOISE_TABLE) return;
#pragma unroll
    for (int q = 0; q < BK_NOISE_TABLE - 1; q++) T.t[q] = ((u32)q >= pos) ? T.t[q + 1] : T.t[q];
    T.t[BK_NOISE_TABLE - 1] = 0;
    nz_table_relo(T);
}
BK_HD void nz_table_update(NzTable& T, double old, double nw) {
    nz_table_evict(T, nz_b(old));
    nz_table_insert(T, nz_b(nw));
}
BK_HD bool nz_table_equal(const NzTable& T, const double* snap) {
    bool eq = true;
#pragma unroll
    for (int q = 0; q < BK_NOISE_TABLE; q++) eq = eq && (T.t[q] == nz_b(snap[q]));
    return eq;
}

// ---- chains: one operand against the binade of the running sum ---------------------------------------------
// The increments (in ulps of the binade of s) that adding x produces for an even / an odd S come from real FP64
// additions against two constants of the same binade: fl(C + x) - C with C = 1.5·2^e (even in ulps) and C + ulp
// (odd).  The hardware rounds C + x exactly like s + x — to the binade's ulp, exact ties to even — so ties, tiny
// and subnormal operands need no special case.  Valid while |x| < 2^(e-1) (then C + x cannot leave C's binade).
struct NzBinade { double c_even, c_odd, scale, xmax; };
BK_HD NzBinade nz_binade(u32 ef) {                                  // ef = exponent field of s, 64 <= ef <= 0x7FE
    NzBinade b;
    b.c_even = nz_d(((u64)ef << 52) | (1ull << 51));
    b.c_odd = nz_d((((u64)ef << 52) | (1ull << 51)) + 1);
    b.scale = nz_d((u64)(2098u - ef) << 52);                         // 2^(52-e): ulps → integers
    b.xmax = nz_d((u64)(ef - 1) << 52);                              // 2^(e-1)
    return b;
}
BK_HD bool nz_incs(const NzBinade& b, double x, i64* i_even, i64* i_odd) {
    if (!(nz_abs(x) < b.xmax)) { *i_even = 0; *i_odd = 0; return false; }     // operand as large as the sum (or non-finite)
    *i_even = (i64)nz_mul(nz_sub(nz_add(b.c_even, x), b.c_even), b.scale);
    *i_odd = (i64)nz_mul(nz_sub(nz_add(b.c_odd, x), b.c_odd), b.scale);
    return true;
}
// running integer offset after the operation, for a chain whose offset before it was `run` from a start of parity p
BK_HD i64 nz_apply(i64 i_even, i64 i_odd, i64 run, u32 p) {
    return run + ((((u64)run + p) & 1ull) ? i_odd : i_even);
}
// (g then f): offsets for start parity 0 / 1
BK_HD void nz_compose(i64 g0, i64 g1, i64 f0, i64 f1, i64* h0, i64* h1) {
    *h0 = g0 + ((g0 & 1) ? f1 : f0);
    *h1 = g1 + (((g1 + 1) & 1) ? f1 : f0);
}
BK_HD bool nz_inside(i64 T) { return T > (1ll << 52) && T < (1ll << 53); }
BK_HD double nz_value(u32 ef, i64 T) { return nz_d(((u64)ef << 52) | ((u64)T & BK_NZ_MASK52)); }

// operation q (0..5) of iteration i: - old_j, + new_j for j = 0..2 (src/call.rs:845-895); M(p, j) = fraction j of position p
template <bool SQUARE, class LdM>
BK_HD double nz_operand(const LdM& M, i32 i, u32 q) {
    const u32 j = q >> 1;
    double v = (q & 1) ? M(i, j) : M(i - BK_NOISE_WINDOW, j);
    if (SQUARE) v = nz_mul(v, v);
    return (q & 1) ? v : -v;
}

// ---- Thompson tau: src/call.rs:898-961.  mx = table after iteration i, returns Noise.max ---------------------
template <class Tau>
BK_HD double nz_tau_loop(u32 cn0, double s0, double s20, const double* mx, const Tau& tau_of) {
    double mu = 0.0, var = 0.0;
    if (cn0 != 0) { mu = nz_div(s0, (double)cn0); var = nz_sub(nz_div(s20, (double)cn0), nz_mul(mu, mu)); }
    u32 idx = 0, cn = cn0;
    double cs = s0, cs2 = s20, cmu = mu, cvar = var;
    double cand = mx[0];
    while (cand != 0.0) {
        const double sd = nz_sqrt(cvar);
        const double tau = (cn > 2) ? tau_of(cn <= 300 ? cn : 300) : nz_d(0x7FF0000000000000ull);
        if (nz_abs(nz_sub(cand, cmu)) > nz_mul(tau, sd)) {
            cs = nz_sub(cs, cand);
            cs2 = nz_sub(cs2, cand);                          // sic: candidate, not its square (src/call.rs:936)
            cn -= 1;
            if (cn > 0) { cmu = nz_div(cs, (double)cn); cvar = nz_sub(nz_div(cs2, (double)cn), nz_mul(cmu, cmu)); }
            else { cmu = 0.0; cvar = 0.0; }
            idx += 1;
            cand = idx < BK_NOISE_TABLE ? mx[idx] : 0.0;      // the reference would panic at idx == 10
        } else break;
    }
    return cand;
}

// ---- one speculative table chunk (a single lane on the device) -----------------------------------------------
// M(p, j): fractions, zero outside [0, len).  Chunk c covers iterations [c*CHUNK, min(iters, (c+1)*CHUNK)).
// warm[10] receives the table after iteration c*CHUNK - 1, snap[i*10 ..] the table after iteration i.
template <class LdM>
BK_HD void nz_table_chunk(const LdM& M, u32 iters, u32 c, double* snap, double* warm) {
    const i32 i_begin = (i32)(c * BK_NZ_CHUNK);
    const i32 i_end = (i32)iters < i_begin + BK_NZ_CHUNK ? (i32)iters : i_begin + BK_NZ_CHUNK;
    i32 b = i_begin - BK_NZ_WARM;
    NzTable T;
    nz_table_clear(T);
    if (b <= 0) b = 0;                                         // exact start
    else {                                                     // speculative: the ten largest of the window
        for (i32 p = b - BK_NOISE_WINDOW; p < b; p++)
            for (u32 j = 0; j < 3; j++) { const double v = M(p, j); if (v > 0.0) nz_table_insert(T, nz_b(v)); }
    }
    for (i32 i = b; i < i_begin; i++)
        for (u32 j = 0; j < 3; j++) nz_table_update(T, M(i - BK_NOISE_WINDOW, j), M(i, j));
#pragma unroll
    for (int q = 0; q < BK_NOISE_TABLE; q++) warm[q] = nz_d(T.t[q]);
    for (i32 i = i_begin; i < i_end; i++) {
        for (u32 j = 0; j < 3; j++) nz_table_update(T, M(i - BK_NOISE_WINDOW, j), M(i, j));
#pragma unroll
        for (int q = 0; q < BK_NOISE_TABLE; q++) snap[(size_t)i * BK_NOISE_TABLE + q] = nz_d(T.t[q]);
    }
}

// Replay from the true state after iteration i0-1 until the speculative trajectory is met; returns the iteration
// at which the states agreed (iters if never).  snap is overwritten with the truth on the way.
template <class LdM>
BK_HD u32 nz_table_replay(const LdM& M, u32 iters, u32 i0, double* snap) {
    NzTable T;
#pragma unroll
    for (int q = 0; q < BK_NOISE_TABLE; q++) T.t[q] = nz_b(snap[(size_t)(i0 - 1) * BK_NOISE_TABLE + q]);
    nz_table_relo(T);
    for (u32 i = i0; i < iters; i++) {
        for (u32 j = 0; j < 3; j++) nz_table_update(T, M((i32)i - BK_NOISE_WINDOW, j), M((i32)i, j));
        double* sp = snap + (size_t)i * BK_NOISE_TABLE;
        if (nz_table_equal(T, sp)) return i;
#pragma unroll
        for (int q = 0; q < BK_NOISE_TABLE; q++) sp[q] = nz_d(T.t[q]);
    }
    return iters;
}

#if defined(__CUDACC__)
// ================================================================================================
// kernels
// ================================================================================================
struct NoiseView {
    const Counters* ctr;
    const u32* genome_row0; const u32* genome_seq_off; const u32* seq_row0;
    const u32* pile; u32 pile_stride;
    double* maf;          // (rows + BK_NZ_PAD * seqs) * 3
    double* snap_s;       // rows + 50 * seqs
    double* snap_s2;
    double* snap_tab;     // (rows + 50 * seqs) * 10
    double* warm;         // chunk slots * 10
    u8* flag;             // chunk slots: boundary check failed
    u32* stats;           // [0] chunks replayed, [1] iterations replayed, [2] chain rounds, [3] chain stops, [4] serial iterations,
                          // [5] / [6] cycles/16 of the s / s2 chain, [7] cycles/16 of the slowest table lane
    double* noise_max;    // rows
};

struct NzSeq { u32 r0, len, iters, mbase, ibase, cbase; bool ok; };
// sequence q of the selected genome: r0 = first row (genome relative), mbase = index of position 0 in the fraction
// array (in positions), ibase = index of iteration 0 in the snapshot arrays, cbase = first chunk slot
__device__ __forceinline__ NzSeq nz_seq(const NoiseView& nv, u32 q) {
    NzSeq s; s.ok = false; s.r0 = s.len = s.iters = s.mbase = s.ibase = s.cbase = 0;
    const i32 best = nv.ctr->best;
    if (best < 0) return s;
    const u32 sq = nv.genome_seq_off[best] + q;
    if (sq >= nv.genome_seq_off[best + 1]) return s;
    s.r0 = nv.seq_row0[sq] - nv.genome_row0[best];
    s.len = nv.seq_row0[sq + 1] - nv.seq_row0[sq];
    s.iters = s.len + BK_NOISE_HALF;
    s.mbase = s.r0 + q * BK_NZ_PAD + BK_NZ_PAD_LO;
    s.ibase = s.r0 + q * BK_NOISE_HALF;
    s.cbase = s.ibase / BK_NZ_CHUNK + q;
    s.ok = true;
    return s;
}

__constant__ double c_tau[301];

// grid (ceil((max_rows + BK_NZ_PAD) / 256), max_seqs)
__global__ void __launch_bounds__(256) k_noise_fracs(NoiseView nv) {
    const NzSeq s = nz_seq(nv, blockIdx.y);
    if (!s.ok) return;
    const u32 x = blockIdx.x * blockDim.x + threadIdx.x;          // position p = x - BK_NZ_PAD_LO
    if (x >= s.len + BK_NZ_PAD) return;
    double m3[3] = {0.0, 0.0, 0.0};
    if (x >= BK_NZ_PAD_LO && x < BK_NZ_PAD_LO + s.len) {
        const u32 row = s.r0 + x - BK_NZ_PAD_LO;
        const uint4 f = *reinterpret_cast<const uint4*>(nv.pile + (size_t)row * 4);
        const uint4 r = *reinterpret_cast<const uint4*>(nv.pile + nv.pile_stride + (size_t)row * 4);
        const u32 f4[4] = {f.x, f.y, f.z, f.w}, r4[4] = {r.x, r.y, r.z, r.w};
        nz_fractions(f4, r4, m3);
    }
    double* o = nv.maf + (size_t)(s.mbase - BK_NZ_PAD_LO + x) * 3;
    o[0] = m3[0]; o[1] = m3[1]; o[2] = m3[2];
}

// ---- chain block: 256 threads, one iteration per thread and round ---------------------------------------------
#define BK_NZ_SEQ_THREADS 256
#define BK_NZ_CHAIN_SMEM ((BK_NZ_TILE + BK_NOISE_WINDOW) * 3 * 8)
#define BK_NZ_TABLE_POS (BK_NOISE_WINDOW + BK_NZ_WARM + BK_NZ_CHUNK)          // positions a chunk lane touches
#define BK_NZ_TABLE_WARPS 8                                                   // chunk lanes per table block
#define BK_NZ_TABLE_SMEM (BK_NZ_TABLE_WARPS * BK_NZ_TABLE_POS * 3 * 8)
#define BK_NZ_SEQ_SMEM (BK_NZ_TABLE_SMEM > BK_NZ_CHAIN_SMEM ? BK_NZ_TABLE_SMEM : BK_NZ_CHAIN_SMEM)

// Developer builds only (tools/build_variant.sh ... -DBK_NZ_PHASES, run with BK_NOISE_DEBUG=1): cycles of the s² chain
// by phase of a round — 0 tile staging, 1 operands + increments + per-thread maps, 2 warp scan, 3 barrier + cross-warp
// combine, 4 sums + range check + reduction + barrier, 5 accepted prefix + barrier, 6 the stop's iteration, 7 serial
// batches — into stats[8..15] (cycles / 16).  The product build compiles none of it.
#ifdef BK_NZ_PHASES
#define BK_NZ_PH_DECL long long ph_t = clock64(); unsigned long long ph_acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
#define BK_NZ_PH(i) do { const long long ph_n = clock64(); ph_acc[i] += (unsigned long long)(ph_n - ph_t); ph_t = ph_n; } while (0)
#define BK_NZ_PH_STORE do { if (tid == 0 && nv.stats && SQUARE) for (int ph_i = 0; ph_i < 8; ph_i++) nv.stats[8 + ph_i] = (u32)(ph_acc[ph_i] >> 4); } while (0)
#else
#define BK_NZ_PH_DECL
#define BK_NZ_PH(i) do {} while (0)
#define BK_NZ_PH_STORE do {} while (0)
#endif

template <bool SQUARE>
__device__ __forceinline__ void nz_chain_block(const NoiseView& nv, const NzSeq& sq, double* mt) {
    __shared__ i64 wt0[2][8], wt1[2][8];
    __shared__ u32 wbad[2][8];
    __shared__ double sstate[2];
    const u32 tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const double* mafp = nv.maf + (size_t)(sq.mbase - BK_NZ_PAD_LO) * 3;       // position -100
    double* snap = (SQUARE ? nv.snap_s2 : nv.snap_s) + sq.ibase;
    const u32 iters = sq.iters;
    const u32 pos_total = sq.len + BK_NZ_PAD;                                   // positions present in mafp
    double s = 0.0;
    u32 i0 = 0, tile_lo = 0, tile_hi = 0;                                       // tile covers iterations [tile_lo, tile_hi)
    u32 round = 0, serial_left = 0;
    const u32 width = BK_NZ_ROUND;                                              // iterations tried per round
    u32 st_rounds = 0, st_stops = 0, st_serial = 0;
    const long long t_begin = clock64();
    BK_NZ_PH_DECL
    while (i0 < iters) {
        if (i0 + 1 > tile_hi || (i0 + BK_NZ_ROUND > tile_hi && tile_hi < iters)) {
            __syncthreads();
            tile_lo = i0; tile_hi = min(iters, i0 + BK_NZ_TILE);
            const u32 nx = (tile_hi - tile_lo + BK_NOISE_WINDOW) * 3;            // positions tile_lo-100 .. tile_hi-1
            const size_t x0 = (size_t)tile_lo * 3;                               // mafp index of position tile_lo-100
            for (u32 x = tid; x < nx; x += BK_NZ_SEQ_THREADS) mt[x] = (tile_lo + x / 3 < pos_total) ? mafp[x0 + x] : 0.0;
            __syncthreads();
            BK_NZ_PH(0);
        }
        const u32 tl = tile_lo;
        auto M = [mt, tl](i32 p, u32 j) { return mt[(u32)(p + BK_NOISE_WINDOW - (i32)tl) * 3 + j]; };
        const u32 n_it = min(width, tile_hi - i0);
        const u64 sb = nz_b(s);
        const u32 ef = (u32)(sb >> 52);
        if (ef < 64u || ef >= 0x7FFu || serial_left) {
            // s is zero / subnormal / negative / non-finite, or rounds stopped making progress: like the reference
            // A serial iteration costs ~1/30 of a round, and a sum that hovers at a binade border crosses it in
            // nearly every iteration for tens to hundreds of iterations: stay serial for as long as the exponent
            // keeps moving (a round would be stopped by its first operations again).
            const u32 n_ser = min(tile_hi - i0, (u32)BK_NZ_SERIAL);
            bool moved = false;
            for (u32 it = 0; it < n_ser; it++) {
#pragma unroll
                for (u32 q = 0; q < 6; q++) { s = nz_add(s, nz_operand<SQUARE>(M, (i32)(i0 + it), q)); moved = moved || (u32)(nz_b(s) >> 52) != ef; }
                if (tid == 0) snap[i0 + it] = s;
            }
            i0 += n_ser; st_serial += n_ser;
            serial_left = moved ? 1u : 0u;
            BK_NZ_PH(7);
            continue;
        }
        const u32 buf = round & 1;
        round++; st_rounds++;
        const NzBinade bin = nz_binade(ef);
        const i64 S0 = (i64)((sb & BK_NZ_MASK52) | (1ull << 52));
        const bool active = tid < n_it;
        i64 pre0[6], pre1[6];
        i64 x0 = 0, x1 = 0;                                                      // exclusive prefix inside the warp
        u32 bad = 6;
        if (wid * 32 < n_it) {                                                   // warps without an iteration skip the work
            i64 run0 = 0, run1 = 0;
#pragma unroll
            for (u32 q = 0; q < 6; q++) {
                const double x = active ? nz_operand<SQUARE>(M, (i32)(i0 + tid), q) : 0.0;
                i64 ie, io;
                const bool ok = nz_incs(bin, x, &ie, &io);
                run0 = nz_apply(ie, io, run0, 0); run1 = nz_apply(ie, io, run1, 1);
                pre0[q] = run0; pre1[q] = run1;
                if (!ok && bad == 6) bad = q;
            }
            BK_NZ_PH(1);
            // inclusive scan of the parity maps over the warp
            i64 f0 = run0, f1 = run1;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const i64 g0 = __shfl_up_sync(0xFFFFFFFFu, f0, o), g1 = __shfl_up_sync(0xFFFFFFFFu, f1, o);
                if (lane >= (u32)o) { i64 h0, h1; nz_compose(g0, g1, f0, f1, &h0, &h1); f0 = h0; f1 = h1; }
            }
            if (lane == 31) { wt0[buf][wid] = f0; wt1[buf][wid] = f1; }
            x0 = __shfl_up_sync(0xFFFFFFFFu, f0, 1); x1 = __shfl_up_sync(0xFFFFFFFFu, f1, 1);
            if (lane == 0) { x0 = 0; x1 = 0; }
        } else {
#pragma unroll
            for (u32 q = 0; q < 6; q++) { pre0[q] = 0; pre1[q] = 0; }
            if (lane == 31) { wt0[buf][wid] = 0; wt1[buf][wid] = 0; }
        }
        BK_NZ_PH(2);
        __syncthreads();
        i64 base = S0;
        for (u32 w = 0; w < wid; w++) base += (base & 1) ? wt1[buf][w] : wt0[buf][w];
        base += (base & 1) ? x1 : x0;
        BK_NZ_PH(3);
        const bool odd = (base & 1) != 0;
        i64 T[6];
#pragma unroll
        for (u32 q = 0; q < 6; q++) {
            T[q] = base + (odd ? pre1[q] : pre0[q]);
            if (!nz_inside(T[q]) && bad > q) bad = q;
        }
        const u32 total_ops = n_it * 6;
        u32 mine = (active && bad < 6) ? tid * 6 + bad : total_ops;
        mine = __reduce_min_sync(0xFFFFFFFFu, mine);
        if (lane == 0) wbad[buf][wid] = mine;
        __syncthreads();
        BK_NZ_PH(4);
        u32 n_ok = total_ops;
#pragma unroll
        for (u32 w = 0; w < 8; w++) n_ok = min(n_ok, wbad[buf][w]);
        if (active && tid * 6 + 5 < n_ok) snap[i0 + tid] = nz_value(ef, T[5]);
        if (n_ok > 0) {
            const u32 owner = (n_ok - 1) / 6, oq = (n_ok - 1) - owner * 6;
            if (tid == owner) {
                i64 Tl = T[0];
#pragma unroll
                for (u32 q = 1; q < 6; q++) if (q == oq) Tl = T[q];
                sstate[buf] = nz_value(ef, Tl);
            }
        }
        __syncthreads();
        if (n_ok > 0) s = sstate[buf];
        BK_NZ_PH(5);
        if (n_ok == total_ops) { i0 += n_it; continue; }
        // the operation that ended the accepted prefix and the rest of its iteration, in real FP64
        st_stops++;
        const u32 ib = n_ok / 6, qb = n_ok - ib * 6;
        for (u32 q = qb; q < 6; q++) s = nz_add(s, nz_operand<SQUARE>(M, (i32)(i0 + ib), q));
        if (tid == 0) snap[i0 + ib] = s;
        i0 += ib + 1;
        serial_left = 1;                               // stops come in bursts (the sum hovers at a binade border)
        BK_NZ_PH(6);
    }
    BK_NZ_PH_STORE;
    if (tid == 0 && nv.stats) {
        atomicAdd(nv.stats + 2, st_rounds); atomicAdd(nv.stats + 3, st_stops); atomicAdd(nv.stats + 4, st_serial);
        nv.stats[SQUARE ? 6 : 5] = (u32)((clock64() - t_begin) >> 4);            // cycles / 16
    }
}

// grid (2 + ceil(max_chunks / 8), max_seqs), 256 threads, BK_NZ_SEQ_SMEM dynamic shared memory
__global__ void __launch_bounds__(BK_NZ_SEQ_THREADS) k_noise_seq(NoiseView nv) {
    extern __shared__ __align__(16) u8 nz_sm[];
    const NzSeq s = nz_seq(nv, blockIdx.y);
    if (!s.ok || s.len < BK_NOISE_WINDOW) return;                 // the reference panics for len < 100 (zero noise reported)
    double* mt = reinterpret_cast<double*>(nz_sm);
    if (blockIdx.x == 0) { nz_chain_block<false>(nv, s, mt); return; }
    if (blockIdx.x == 1) { nz_chain_block<true>(nv, s, mt); return; }
    // table chunks: warp w of this block owns chunk (blockIdx.x - 2) * 8 + w; its lanes stage the fractions the chunk
    // touches into shared memory, lane 0 walks them
    const u32 lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    if (wid >= BK_NZ_TABLE_WARPS) return;
    const u32 c = (blockIdx.x - 2) * BK_NZ_TABLE_WARPS + wid;
    const u32 n_chunks = (s.iters + BK_NZ_CHUNK - 1) / BK_NZ_CHUNK;
    if (c >= n_chunks) return;
    double* wm = mt + (size_t)wid * BK_NZ_TABLE_POS * 3;
    const i32 i_begin = (i32)(c * BK_NZ_CHUNK);
    i32 p_lo = i_begin - BK_NZ_WARM - BK_NOISE_WINDOW;            // first position the chunk may touch
    if (p_lo < -BK_NOISE_WINDOW) p_lo = -BK_NOISE_WINDOW;
    const i32 p_hi = min((i32)s.iters, i_begin + BK_NZ_CHUNK);    // one past the last
    const double* mafp = nv.maf + (size_t)s.mbase * 3;            // position 0
    const u32 nx = (u32)(p_hi - p_lo) * 3;
    for (u32 x = lane; x < nx; x += 32) wm[x] = mafp[(i64)p_lo * 3 + (i64)x];
    __syncwarp();
    if (lane == 0) {
        const long long t_begin = clock64();
        auto M = [wm, p_lo](i32 p, u32 j) { return wm[(u32)(p - p_lo) * 3 + j]; };
        nz_table_chunk(M, s.iters, c, nv.snap_tab + (size_t)s.ibase * BK_NOISE_TABLE, nv.warm + (size_t)(s.cbase + c) * BK_NOISE_TABLE);
        if (nv.stats) atomicMax(nv.stats + 7, (u32)((clock64() - t_begin) >> 4));
    }
}

// grid (1, max_seqs), 256 threads: verify the chunk boundaries in parallel, replay the failed ones in order
__global__ void __launch_bounds__(256) k_noise_fix(NoiseView nv) {
    const NzSeq s = nz_seq(nv, blockIdx.y);
    if (!s.ok || s.len < BK_NOISE_WINDOW) return;
    u8* flag = nv.flag + s.cbase;                                 // n_chunks bytes
    const u32 n_chunks = (s.iters + BK_NZ_CHUNK - 1) / BK_NZ_CHUNK;
    double* snap = nv.snap_tab + (size_t)s.ibase * BK_NOISE_TABLE;
    const double* warm = nv.warm + (size_t)s.cbase * BK_NOISE_TABLE;
    for (u32 c = threadIdx.x; c < n_chunks; c += blockDim.x) {
        bool bad = false;
        if (c > 0) {
            const double* a = snap + (size_t)(c * BK_NZ_CHUNK - 1) * BK_NOISE_TABLE;
            const double* b = warm + (size_t)c * BK_NOISE_TABLE;
            for (int q = 0; q < BK_NOISE_TABLE; q++) bad = bad || (nz_b(a[q]) != nz_b(b[q]));
        }
        flag[c] = bad ? 1 : 0;
    }
    __syncthreads();
    if (threadIdx.x != 0) return;
    const double* mafp = nv.maf + (size_t)s.mbase * 3;
    auto M = [mafp](i32 p, u32 j) { return mafp[(i64)p * 3 + j]; };
    u32 n_replayed = 0, it_replayed = 0;
    u32 c = 1;
    while (c < n_chunks) {
        if (!flag[c]) { c++; continue; }
        const u32 i0 = c * BK_NZ_CHUNK;
        const u32 met = nz_table_replay(M, s.iters, i0, snap);
        n_replayed++; it_replayed += min(met, s.iters - 1) - i0 + 1;
        c = met / BK_NZ_CHUNK + 1;                                // boundaries inside the replayed span are settled
    }
    if (nv.stats && n_replayed) { atomicAdd(nv.stats + 0, n_replayed); atomicAdd(nv.stats + 1, it_replayed); }
}

// grid (ceil((max_rows + 50) / 256), max_seqs), 256 threads: n by counting, then the Thompson-tau loop
__global__ void __launch_bounds__(256) k_noise_tau(NoiseView nv) {
    __shared__ u8 cnt[256 + BK_NOISE_WINDOW];
    const NzSeq s = nz_seq(nv, blockIdx.y);
    if (!s.ok) return;
    const u32 i_base = blockIdx.x * 256;
    if (s.len < BK_NOISE_WINDOW) {                                // the reference panics here; report zero noise
        for (u32 i = i_base + threadIdx.x; i < min(i_base + 256, s.len); i += blockDim.x) nv.noise_max[s.r0 + i] = 0.0;
        return;
    }
    if (i_base >= s.iters) return;
    const double* mafp = nv.maf + (size_t)s.mbase * 3;
    // positive fractions per position p = i_base - 99 + x, x in [0, 355)
    for (u32 x = threadIdx.x; x < 256 + BK_NOISE_WINDOW - 1; x += blockDim.x) {
        const i64 p = (i64)i_base - (BK_NOISE_WINDOW - 1) + x;   // >= -99; positions past len+149 are never needed
        u32 n = 0;
        if (p < (i64)s.len) { const double* m = mafp + p * 3; n = (m[0] > 0.0 ? 1u : 0u) + (m[1] > 0.0 ? 1u : 0u) + (m[2] > 0.0 ? 1u : 0u); }
        cnt[x] = (u8)n;
    }
    __syncthreads();
    const u32 i = i_base + threadIdx.x;
    if (i >= s.iters || i < BK_NOISE_HALF) return;
    u32 cn0 = 0;                                                   // n after iteration i: positions [i-99, i]
    for (u32 x = 0; x < BK_NOISE_WINDOW; x++) cn0 += cnt[threadIdx.x + x];
    const double s0 = nv.snap_s[s.ibase + i], s20 = nv.snap_s2[s.ibase + i];
    const double* mxp = nv.snap_tab + (size_t)(s.ibase + i) * BK_NOISE_TABLE;
    double mx[BK_NOISE_TABLE];
#pragma unroll
    for (int q = 0; q < BK_NOISE_TABLE; q += 2) { const double2 v = *reinterpret_cast<const double2*>(mxp + q); mx[q] = v.x; mx[q + 1] = v.y; }
    auto tau_of = [](u32 n) { return c_tau[n]; };
    nv.noise_max[s.r0 + i - BK_NOISE_HALF] = nz_tau_loop(cn0, s0, s20, mx, tau_of);
}
#endif  // __CUDACC__

}  // namespace bk
