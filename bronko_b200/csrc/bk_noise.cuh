// bk_noise.cuh — the noise baseline of the reference (get_baseline_noise, src/call.rs:799-967; quirks: SURVEY.md
// Appendix C, Q12) as a parallel computation with bit-identical results.
//
// The reference is ONE sequential loop over i in [0, len+50) carrying three pieces of state:
//   s, s2   running FP64 sums of the minor-allele fractions (and their squares) of the last 100 positions,
//   maxes   a 10-entry "max table" with an evict-by-value quirk,
// followed per iteration by the modified Thompson-tau rejection loop, which only READS that state.
// Everything is split so that only genuinely sequential work stays sequential:
//
//   k_noise_fracs   one thread per position: sorted allele fractions (order-free); per iteration whether any of its
//                   six operands is non-zero ("active") and the APPROXIMATE window sums (the chains' look-ahead).
//   k_noise_seq     block 0 / 1 of every sequence: the s / s2 chains over the active iterations, replicated EXACTLY
//                   (see "chains" below); blocks >= 2: the max table, one lane per chunk of 128 iterations,
//                   started SPECULATIVELY 256 iterations early from the ten largest fractions of the window.
//   k_noise_fix     verifies every speculative chunk against the true state at its boundary (the table after
//                   iteration i depends only on the table after i-1 and on the data, so equal states at one
//                   iteration prove everything after it) and replays from the true state where the warm-up
//                   had not converged yet, until it meets the speculative trajectory.
//   k_noise_tau     one thread per output position: n (a count), then the Thompson-tau loop on the snapshots.
//
// chains.  An iteration whose six operands are zero changes nothing and is skipped.  The others are taken 256 at a time:
// a look-ahead over the approximate window sums decides whether the run of iterations ahead that fits three adjacent
// binades is long enough for a ROUND, or whether it and what follows are added up in real FP64 like the reference
// (sparse windows: the sum halves, doubles or cancels every few iterations).  A round works in units of the lowest
// binade's ulp u: with S = s/u and X = x/u = A + f, fl(s + x) = (S + A + r)·u where the rounding r (|r| <= 2) depends on
// S only through S mod 8 once the binade ("zone") of the operation is known — which the integer prefix sum of the A's
// tells.  Every operation is thus a map "S mod 8 -> r", those maps compose associatively, and 1536 consecutive
// operations (256 iterations, one per thread) are three block-wide scans.  The first operation that cannot be decided
// or leaves the zones ends the accepted prefix and is executed in real FP64.  ("three-zone rounds" and "look-ahead"
// below; nz_chain_block.)
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
#ifndef BK_NZ_IPT
#define BK_NZ_IPT 1              // iterations per thread and chain round (nz_chain_block is written for 1)
#endif
#define BK_NZ_OPT (6 * BK_NZ_IPT) // operations per thread and round
#ifndef BK_NZ_SEQ_THREADS
#define BK_NZ_SEQ_THREADS 256     // threads of a chain block = active iterations per pass (512: measured slower — a pass is dependent integer chains, and a stop throws more away)
#endif
#define BK_NZ_WARPS (BK_NZ_SEQ_THREADS / 32)
#define BK_NZ_ROUND (BK_NZ_SEQ_THREADS * BK_NZ_IPT)   // iterations per chain round
#ifndef BK_NZ_MIN_RUN
#define BK_NZ_MIN_RUN 64         // a run of iterations that fits three zones but is shorter than this is walked in real FP64 (a round costs ~100 such iterations)
#define BK_NZ_SERIAL_RUN 64      // ... for at least this many iterations
#endif
#ifndef BK_NZ_SERIAL
#define BK_NZ_SERIAL 8           // iterations executed serially after a round that accepted next to nothing although the look-ahead saw a long run
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
BK_HD long long nz_floor_ll(double a) { return __double2ll_rd(a); }
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
BK_HD long long nz_floor_ll(double a) { return (long long)std::floor(a); }
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

// ---- chains across binade borders: three-zone rounds -------------------------------------------------------------
// A sum that hovers at a power of two crosses it again and again (thin coverage: in nearly every iteration for tens to
// hundreds of iterations), an iSNV entering the window doubles it — and a round that only knows one binade stops at
// every crossing.  Three adjacent binades are therefore handled together, in units of the LOWEST one's ulp u: zone g
// (g = 0, 1, 2) holds the multiples of G = 2^g in [2^(52+g), 2^(53+g)).  With X = x/u = A + f (A = floor, 0 <= f < 1,
// both exact) and v = S + A, the IEEE result of S + x is v rounded to a multiple of G, G the zone v falls in:
//     low = v mod G, w = v - low;   up  = low + f > G/2;   tie = low + f = G/2
//     result = w + G * [up or (tie and w/G odd)]           (round to nearest, ties to even)
// i.e. S + A + r with |r| <= 2, r depending on S only through S mod 8 — once the zone of the operation is known.  The
// zone follows from the integer prefix sum of the A's: the roundings move the true S by at most two units per
// operation, so S0 + (A's so far) decides unless it is within that drift of a zone border (then the operation ends the
// accepted prefix and is executed in real FP64, like an operation that leaves the three zones).  Every operation is thus
// a map "S mod 8 -> r", those maps compose associatively, and a round is two block-wide scans: the A's, then the maps.
#define BK_NZ_ZONES 3
#define BK_NZ_CLASSES 8
struct NzZ2 { u32 el; double scale, xmax; };              // el = exponent field of the lowest zone
BK_HD bool nz2_zones(u64 sb, NzZ2* z) {                    // false: s is zero / tiny / non-finite / negative → serial
    const u32 ef = (u32)(sb >> 52);
    if (ef < 66u || ef >= 0x7FCu) return false;
    z->el = ef - 1;                                       // one binade of room below, one above
    z->scale = nz_d((u64)(2098u - z->el) << 52);          // 1 / ulp(lowest zone)
    z->xmax = nz_d((u64)(z->el + BK_NZ_ZONES) << 52);     // operands beyond the top of the zones cannot leave the sum inside them
    return true;
}
// Where to put the three zones: look-ahead.  The window sum after every iteration is known APPROXIMATELY before the
// chains start (k_noise_fracs adds the 300 fractions of the window in any order: the chain's own value differs from it
// by rounding noise only), and with it the binades the chain is going to visit.  A round picks the lowest zone so that
// the longest run of iterations ahead of it fits — one binade of room on either side of the start (the fixed choice of
// the first version) ended 60-80 % of the rounds of a thin sample after ~130 of 256 iterations, because the sum of a
// sparse window halves or doubles within a few dozen positions.  The choice is a hint: every operation is still checked
// against the zones it was computed for.
// key = (highest exponent field << 16) | (0xFFFF - lowest exponent field); merging two keys = per-half maximum.
#define BK_NZ_HINT_NEUTRAL 0u
BK_HD u32 nz_hint_key(u64 bits) {
    const u32 e = (u32)(bits >> 52);                       // (sign set: e >= 0x800 → breaks)
    if (bits == 0 || e < 66u || e >= 0x7FCu) return (0x7FFu << 16) | (0xFFFFu - 1u);     // a sum the rounds cannot carry: the run ends here
    const u64 m = bits & BK_NZ_MASK52;
    const u32 lo = e - (m < (1ull << 23) ? 1u : 0u), hi = e + (m > BK_NZ_MASK52 - (1ull << 23) ? 1u : 0u);    // next to a power of two: either side
    return (hi << 16) | (0xFFFFu - lo);
}
BK_HD u32 nz_hint_start(u32 ef) { return (ef << 16) | (0xFFFFu - ef); }
BK_HD u32 nz_hint_merge(u32 a, u32 b) {
#if defined(__CUDA_ARCH__)
    return __vmaxu2(a, b);
#else
    const u32 h = (a >> 16) > (b >> 16) ? (a >> 16) : (b >> 16), l = (a & 0xFFFFu) > (b & 0xFFFFu) ? (a & 0xFFFFu) : (b & 0xFFFFu);
    return (h << 16) | l;
#endif
}
BK_HD bool nz_hint_fits(u32 key) { return (key >> 16) - (0xFFFFu - (key & 0xFFFFu)) <= BK_NZ_ZONES - 1u; }
// lowest zone for the merged key of the run (it contains the start: ef - 2 <= result <= ef).  A spare zone goes BELOW the
// run: every iteration subtracts before it adds.
BK_HD u32 nz_hint_el(u32 key) {
    const u32 hi = key >> 16, lo = 0xFFFFu - (key & 0xFFFFu);
    return hi - lo == BK_NZ_ZONES - 1u ? lo : lo - 1u;
}
BK_HD void nz2_zones_at(u32 el, NzZ2* z) {
    z->el = el;
    z->scale = nz_d((u64)(2098u - el) << 52);
    z->xmax = nz_d((u64)(el + BK_NZ_ZONES) << 52);
}
BK_HD i64 nz2_start(const NzZ2& z, u64 sb) {
    const i64 m = (i64)((sb & BK_NZ_MASK52) | (1ull << 52));
    return m << ((u32)(sb >> 52) - z.el);
}
BK_HD bool nz2_split(const NzZ2& z, double x, i64* A, u32* fc) {
    if (!(nz_abs(x) < z.xmax)) { *A = 0; *fc = 0; return false; }
    const double X = nz_mul(x, z.scale);                  // exact: a power of two; |X| < 2^55
    const i64 a = nz_floor_ll(X);
    const double f = nz_sub(X, (double)a);                // exact, in [0, 1) (zero once |X| >= 2^52)
    *A = a;
    *fc = f == 0.0 ? 0u : (f < 0.5 ? 1u : (f == 0.5 ? 2u : 3u));
    return true;
}
// r = result - (S + A) for S mod 8 = m before the operation, the operation's result in zone g: a function of
// v = (S + A) mod 8 only, tabulated per (zone, fraction class) as eight signed bytes (nz2_round_slow is the formula the
// table is generated from; tests/emul checks one against the other for all 96 cases).
BK_HD i32 nz2_round_slow(u32 g, u32 fc, u32 v) {
    if (g == 0) return (fc == 3u || (fc == 2u && (v & 1u))) ? 1 : 0;
    const u32 G = 1u << g, low = v & (G - 1u), half = G >> 1;
    const bool up = low > half || (low == half && fc != 0u);
    const bool tie = low == half && fc == 0u;
    const bool odd = ((v - low) >> g) & 1u;
    return (i32)((up || (tie && odd)) ? G : 0u) - (i32)low;
}
BK_HD u64 nz2_round_table(u32 g, u32 fc) {                 // byte v = r for (S + A) mod 8 = v
    if (g == 0) return fc == 3u ? 0x0101010101010101ull : (fc == 2u ? 0x0100010001000100ull : 0ull);
    if (g == 1) return fc == 0u ? 0x0100FF000100FF00ull : 0x0100010001000100ull;
    return fc == 0u ? 0x0102FF0001FEFF00ull : 0x0102FF000102FF00ull;
}
BK_HD bool nz2_inside(i64 T) { return T > (1ll << 52) && T < (1ll << (52 + BK_NZ_ZONES)); }
BK_HD double nz2_value(const NzZ2& z, i64 T) {
    const u32 g = T >= (1ll << 54) ? 2u : (T >= (1ll << 53) ? 1u : 0u);
    return nz_d(((u64)(z.el + g) << 52) | (((u64)T >> g) & BK_NZ_MASK52));
}
// Eight classes side by side: a "class vector" holds one byte per start class m (two words: classes 0-3, 4-7).  Table
// look-ups for all classes are byte permutes (PRMT: four look-ups into an eight-byte table per instruction).
struct NzVec { u32 lo, hi; };
BK_HD NzVec nzvec_identity() { NzVec v; v.lo = 0x03020100u; v.hi = 0x07060504u; return v; }
BK_HD u32 nz_sel(u32 v) { const u32 t = v | (v >> 4); return (t & 0xFFu) | ((t >> 8) & 0xFF00u); }      // four bytes (each 0..7) → four selector nibbles
BK_HD u32 nz_vadd4(u32 a, u32 b) { return ((a & 0x7F7F7F7Fu) + (b & 0x7F7F7F7Fu)) ^ ((a ^ b) & 0x80808080u); }   // byte-wise a + b
BK_HD NzVec nzvec_lookup(u32 tlo, u32 thi, const NzVec& idx) {    // byte m = table[idx byte m]
    NzVec r; r.lo = BK_PRMT(tlo, thi, nz_sel(idx.lo)); r.hi = BK_PRMT(tlo, thi, nz_sel(idx.hi)); return r;
}
BK_HD u32 nzvec_get(const NzVec& v, u32 m) { return ((m < 4 ? v.lo : v.hi) >> (8 * (m & 3u))) & 0xFFu; }
BK_HD NzVec nzvec_compose(const NzVec& g, const NzVec& f) {       // transition maps: g first, then f
    return nzvec_lookup(f.lo, f.hi, g);
}
// one thread's operations (BK_NZ_OPT: six per iteration) for all eight start classes at once: pr[q] = roundings up to and
// including operation q (a signed byte per class), next = the class after them, and which operations cannot be decided
// (bad = first such operation, BK_NZ_OPT = none).  s_est = S0 + (A's of the round before this thread), first_idx = operations
// of the round before this thread.
struct NzThread { NzVec pr[BK_NZ_OPT]; i64 a_pre[BK_NZ_OPT]; NzVec next; u32 bad; };
BK_HD void nz2_thread(const i64* A, const u32* fc, const bool* ok, i64 s_est, u32 first_idx, NzThread* t) {
    i64 run = 0;
    NzVec cur = nzvec_identity(), tot; tot.lo = 0; tot.hi = 0;
    t->bad = BK_NZ_OPT;
#if defined(__CUDACC__)
#pragma unroll
#endif
    for (u32 q = 0; q < BK_NZ_OPT; q++) {
        const i64 v = s_est + run + A[q];                   // S + A without the roundings so far: off by at most `drift`
        const i64 drift = 2ll * (i64)(first_idx + q);         // (|r| <= 2 per operation)
        u32 g = 0; bool sure = true;
        for (u32 b = 1; b < BK_NZ_ZONES; b++) {
            const i64 E = v - (1ll << (52 + b));
            if (E >= drift) g = b; else if (!(E < -drift)) sure = false;
        }
        if ((!sure || !ok[q]) && t->bad == BK_NZ_OPT) t->bad = q;
        run += A[q];
        t->a_pre[q] = run;
        const u64 tab = nz2_round_table(g, fc[q]);
        const u32 arep = ((u32)(u64)A[q] & 7u) * 0x01010101u;
        NzVec vv; vv.lo = (cur.lo + arep) & 0x07070707u; vv.hi = (cur.hi + arep) & 0x07070707u;     // (S + A) mod 8 per class
        const NzVec r = nzvec_lookup((u32)tab, (u32)(tab >> 32), vv);
        tot.lo = nz_vadd4(tot.lo, r.lo); tot.hi = nz_vadd4(tot.hi, r.hi);
        t->pr[q] = tot;
        cur.lo = nz_vadd4(vv.lo, r.lo) & 0x07070707u; cur.hi = nz_vadd4(vv.hi, r.hi) & 0x07070707u;
    }
    t->next = cur;
}
BK_HD i32 nz2_pr(const NzVec& v, u32 m) { return (i32)(signed char)(unsigned char)nzvec_get(v, m); }

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
    // active iterations (nz_chain_block): an iteration whose six operands are all zero leaves both sums as they are
    u8* actflag;          // rows + 50 * seqs: iteration i has a non-zero operand (at ibase + i; k_noise_fracs)
    u32* act_list;        // rows + 50 * seqs: the active iterations of a sequence, ascending (at ibase)
    u32* act_rank;        // rows + 50 * seqs: active iterations <= i (at ibase + i)
    u32* act_n;           // seqs
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

// grid (ceil((max_rows + BK_NZ_PAD) / 256), max_seqs).  Also the look-ahead of the chains (nz_hint_key): the approximate
// window sums after every iteration go where the chains will write the exact ones (snap_s / snap_s2).
__global__ void __launch_bounds__(256) k_noise_fracs(NoiseView nv) {
    __shared__ double fr[(256 + BK_NOISE_WINDOW) * 3];              // positions x0 - 100 .. x0 + 255
    const NzSeq s = nz_seq(nv, blockIdx.y);
    if (!s.ok) return;
    const u32 x0 = blockIdx.x * blockDim.x;                         // position p = x - BK_NZ_PAD_LO
    if (x0 >= s.len + BK_NZ_PAD) return;
    for (u32 t = threadIdx.x; t < 256 + BK_NOISE_WINDOW; t += 256) {
        const i64 x = (i64)x0 - BK_NOISE_WINDOW + t;
        double m3[3] = {0.0, 0.0, 0.0};
        if (x >= BK_NZ_PAD_LO && x < (i64)BK_NZ_PAD_LO + s.len) {
            const u32 row = s.r0 + (u32)x - BK_NZ_PAD_LO;
            const uint4 f = *reinterpret_cast<const uint4*>(nv.pile + (size_t)row * 4);
            const uint4 r = *reinterpret_cast<const uint4*>(nv.pile + nv.pile_stride + (size_t)row * 4);
            const u32 f4[4] = {f.x, f.y, f.z, f.w}, r4[4] = {r.x, r.y, r.z, r.w};
            nz_fractions(f4, r4, m3);
        }
        fr[t * 3] = m3[0]; fr[t * 3 + 1] = m3[1]; fr[t * 3 + 2] = m3[2];
        if (t >= BK_NOISE_WINDOW && x < (i64)s.len + BK_NZ_PAD) {
            double* o = nv.maf + (size_t)(s.mbase - BK_NZ_PAD_LO + (u32)x) * 3;
            o[0] = m3[0]; o[1] = m3[1]; o[2] = m3[2];
        }
    }
    __syncthreads();
    const u32 x = x0 + threadIdx.x;
    if (x >= BK_NZ_PAD_LO && x - BK_NZ_PAD_LO < s.iters) {           // iteration i = position i: the window is positions i - 99 .. i
        double a = 0.0, b = 0.0;
        for (u32 w = 1; w <= BK_NOISE_WINDOW; w++)
#pragma unroll
            for (u32 j = 0; j < 3; j++) { const double v = fr[(threadIdx.x + w) * 3 + j]; a += v; b += v * v; }
        nv.snap_s[s.ibase + x - BK_NZ_PAD_LO] = a;
        nv.snap_s2[s.ibase + x - BK_NZ_PAD_LO] = b;
        const double* fo = fr + threadIdx.x * 3;                                 // position i - 100
        const double* fn = fr + (threadIdx.x + BK_NOISE_WINDOW) * 3;             // position i
        nv.actflag[s.ibase + x - BK_NZ_PAD_LO] = (fo[0] != 0.0 || fo[1] != 0.0 || fo[2] != 0.0 || fn[0] != 0.0 || fn[1] != 0.0 || fn[2] != 0.0) ? 1 : 0;
    }
}

// ---- chain block: 256 threads, one ACTIVE iteration per thread and pass ---------------------------------------
// An iteration whose six operands are zero (no minor allele at position i nor at i - 100) leaves s and s² exactly as
// they are, so the chains only walk the active ones: all of them at 10,000x, two thirds at 5,000x, 9 % at 2,000x, 90 of
// 29,953 at 400x.  The sums of the skipped iterations are read through act_rank / act_list (k_noise_tau).
// A pass takes the next 256 active iterations, loads their operands and looks ahead (nz_hint_key): if the run of
// iterations that fits three zones is long, the pass is a round (three block scans, below); if the sum is about to halve
// or double every few iterations — a window that holds a handful of fractions — the run and what follows are added up
// in real FP64, one after the other like the reference, which costs ~1 % of a round per iteration.
#define BK_NZ_CHAIN_SMEM (BK_NZ_ROUND * 6 * 8)
#define BK_NZ_TABLE_POS (BK_NOISE_WINDOW + BK_NZ_WARM + BK_NZ_CHUNK)          // positions a chunk lane touches
#define BK_NZ_TABLE_WARPS 8                                                   // chunk lanes per table block
#define BK_NZ_TABLE_SMEM (BK_NZ_TABLE_WARPS * BK_NZ_TABLE_POS * 3 * 8)
#define BK_NZ_SEQ_SMEM (BK_NZ_TABLE_SMEM > BK_NZ_CHAIN_SMEM ? BK_NZ_TABLE_SMEM : BK_NZ_CHAIN_SMEM)

// Developer builds only (tools/build_variant.sh ... -DBK_NZ_PHASES, run with BK_NOISE_DEBUG=1): cycles of the s² chain
// by phase of a pass — 0 compaction, 1 operands + look-ahead, 2 split + warp scan, 3 barrier + maps + cross-warp combine,
// 4 sums + range check + reduction + barrier, 5 accepted prefix + barrier, 6 the stop's iteration, 7 serial runs — into
// stats[8..15] (cycles / 16).  The product build compiles none of it.
#ifdef BK_NZ_PHASES
#define BK_NZ_PH_DECL long long ph_t = clock64(); unsigned long long ph_acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
#define BK_NZ_PH(i) do { const long long ph_n = clock64(); ph_acc[i] += (unsigned long long)(ph_n - ph_t); ph_t = ph_n; } while (0)
#define BK_NZ_PH_STORE do { if (tid == 0 && nv.stats && SQUARE) for (int ph_i = 0; ph_i < 8; ph_i++) nv.stats[8 + ph_i] = (u32)(ph_acc[ph_i] >> 4); } while (0)
#else
#define BK_NZ_PH_DECL
#define BK_NZ_PH(i) do {} while (0)
#define BK_NZ_PH_STORE do {} while (0)
#endif

__device__ __forceinline__ NzVec nzvec_shfl_up(const NzVec& f, int o) {
    NzVec g;
    g.lo = __shfl_up_sync(0xFFFFFFFFu, f.lo, o); g.hi = __shfl_up_sync(0xFFFFFFFFu, f.hi, o);
    return g;
}

// The active iterations of a sequence → act_list / act_rank / act_n.  Both chain blocks run it and write the same
// values; each reads what it needs after its own barrier.  A warp owns a contiguous span of iterations and walks it 32
// at a time (one coalesced load of the flags k_noise_fracs wrote, a ballot), sixteen loads in flight.
#define BK_NZ_CU 16
__device__ __forceinline__ u32 nz_compact_active(const NoiseView& nv, const NzSeq& sq, u32* wtmp /* BK_NZ_WARPS words of shared memory */) {
    const u32 tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const u8* fl = nv.actflag + sq.ibase;
    const u32 steps = ((sq.iters + 31) / 32 + BK_NZ_WARPS - 1) / BK_NZ_WARPS;   // 32-iteration steps per warp
    const u32 b = wid * steps * 32;
    u32 cnt = 0;
    for (u32 st = 0; st < steps; st += BK_NZ_CU) {
        u32 f[BK_NZ_CU];
#pragma unroll
        for (u32 u = 0; u < BK_NZ_CU; u++) { const u32 i = b + (st + u) * 32 + lane; f[u] = (st + u < steps && i < sq.iters) ? fl[i] : 0u; }
#pragma unroll
        for (u32 u = 0; u < BK_NZ_CU; u++) cnt += (u32)__popc(__ballot_sync(0xFFFFFFFFu, f[u] != 0));
    }
    if (lane == 0) wtmp[wid] = cnt;
    __syncthreads();
    u32 base = 0, total = 0;
#pragma unroll
    for (u32 w = 0; w < BK_NZ_WARPS; w++) { const u32 v = wtmp[w]; if (w < wid) base += v; total += v; }
    u32* list = nv.act_list + sq.ibase;
    u32* rank = nv.act_rank + sq.ibase;
    for (u32 st = 0; st < steps; st += BK_NZ_CU) {
        u32 f[BK_NZ_CU];
#pragma unroll
        for (u32 u = 0; u < BK_NZ_CU; u++) { const u32 i = b + (st + u) * 32 + lane; f[u] = (st + u < steps && i < sq.iters) ? fl[i] : 0u; }
#pragma unroll
        for (u32 u = 0; u < BK_NZ_CU; u++) {
            const u32 i = b + (st + u) * 32 + lane;
            const u32 m = __ballot_sync(0xFFFFFFFFu, f[u] != 0);
            const u32 incl = base + (u32)__popc(m & (0xFFFFFFFFu >> (31 - lane)));       // active iterations <= i
            if (f[u]) list[incl - 1] = i;
            if (st + u < steps && i < sq.iters) rank[i] = incl;
            base += (u32)__popc(m);
        }
    }
    if (tid == 0) nv.act_n[blockIdx.y] = total;
    __syncthreads();                                                            // (the list is read by other threads of this block)
    return total;
}

template <bool SQUARE>
__device__ __forceinline__ void nz_chain_block(const NoiseView& nv, const NzSeq& sq, double* xs) {
    __shared__ i64 wsumA[2][BK_NZ_WARPS];
    __shared__ u32 wnextL[2][BK_NZ_WARPS], wnextH[2][BK_NZ_WARPS];               // per warp: class transition of the whole warp
    __shared__ i32 wsumR[2][BK_NZ_WARPS];                              // per warp: roundings of the whole warp
    __shared__ u32 wbad[2][BK_NZ_WARPS];
    __shared__ u32 whint[2][BK_NZ_WARPS], wlast[BK_NZ_WARPS];           // look-ahead: run keys of the warps (per pass parity) / the warps' last fitting keys
    __shared__ u32 wfit[BK_NZ_WARPS + 1];                               // look-ahead: fitting iterations per warp (and scratch of the compaction)
    __shared__ double sstate[2];
    static_assert(BK_NZ_IPT == 1, "one active iteration per thread and pass");
    const u32 tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const long long t_begin = clock64();
    BK_NZ_PH_DECL
    const u32 n_act = nz_compact_active(nv, sq, wfit);
    BK_NZ_PH(0);
    const double* maf0 = nv.maf + (size_t)sq.mbase * 3;                         // position 0
    const u32* list = nv.act_list + sq.ibase;
    double* snap = (SQUARE ? nv.snap_s2 : nv.snap_s) + sq.ibase;
    double s = 0.0;
    u32 a0 = 0, round = 0, force_serial = 0;                                    // a0 = active iterations done
    u32 pf_a0 = 0xFFFFFFFFu, it_n = 0, hk_n = BK_NZ_HINT_NEUTRAL;               // the pass whose operands are in x_n / hb_n and whose keys are in whint
    u64 hb_n = 0;
    double x_n[6] = {0.0, 0.0, 0.0, 0.0, 0.0, 0.0};
    u32 st_rounds = 0, st_stops = 0, st_serial = 0;
    while (a0 < n_act) {
        const u32 n_it = min((u32)BK_NZ_ROUND, n_act - a0);
        const u32 buf = round & 1;
        round++;
        // this thread's iteration: its six operands (-old_j, +new_j; src/call.rs:845-895) from the fraction array, its look-ahead
        // key — loaded during the previous round if that round was expected to end where it did (pf_a0), else now
        u32 it = 0, hk = BK_NZ_HINT_NEUTRAL;
        double x[6];
        const bool prefetched = pf_a0 == a0;
        if (!prefetched) {
            if (tid < n_it) {
                it_n = list[a0 + tid];
                const double* mo = maf0 + ((i64)it_n - BK_NOISE_WINDOW) * 3;
                const double* mn = maf0 + (i64)it_n * 3;
                hb_n = nz_b(snap[it_n]);                                         // (k_noise_fracs left the approximate window sum here)
#pragma unroll
                for (u32 j = 0; j < 3; j++) { x_n[2 * j] = mo[j]; x_n[2 * j + 1] = mn[j]; }
            }
        }
        if (tid < n_it) {
            it = it_n;
#pragma unroll
            for (u32 j = 0; j < 3; j++) {
                double o = x_n[2 * j], n = x_n[2 * j + 1];
                if (SQUARE) { o = nz_mul(o, o); n = nz_mul(n, n); }
                x[2 * j] = -o; x[2 * j + 1] = n;
            }
            hk = nz_hint_key(hb_n);
        } else {
#pragma unroll
            for (u32 q = 0; q < 6; q++) x[q] = 0.0;
        }
#pragma unroll
        for (u32 q = 0; q < 6; q++) xs[tid * 6 + q] = x[q];
        // look-ahead: how many of the iterations ahead fit three zones, and where to put the zones
        const u64 sb = nz_b(s);
        const u32 ef = (u32)(sb >> 52);
        const bool capable = ef >= 66u && ef < 0x7FCu;                           // positive, normal, finite (a set sign bit makes ef >= 0x800)
        const u32 start = nz_hint_start(capable ? ef : 1023u);
        if (!prefetched) {
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) { const u32 g = __shfl_up_sync(0xFFFFFFFFu, hk, o); if (lane >= (u32)o) hk = nz_hint_merge(hk, g); }
            if (lane == 31) whint[buf][wid] = hk;
            hk_n = hk;
            __syncthreads();
        }
        hk = hk_n;                                                               // (prefetched: scanned and published before the last round's barriers)
        // the next pass, if this one turns out to be a round that accepts everything: its iteration index now, ...
        const u32 a_next = a0 + n_it, n_next = min((u32)BK_NZ_ROUND, n_act - a_next);
        u32 it_p = 0;
        if (tid < n_next) it_p = list[a_next + tid];
        u32 pre = start;
#pragma unroll
        for (u32 w = 0; w < BK_NZ_WARPS - 1; w++) { const u32 v = whint[buf][w]; if (w < wid) pre = nz_hint_merge(pre, v); }
        hk = nz_hint_merge(hk, pre);
        const bool fits = tid < n_it && nz_hint_fits(hk);                        // (the keys only grow along the run: fits is a prefix)
        const u32 fv = fits ? hk : start;
        const u32 vh = __reduce_max_sync(0xFFFFFFFFu, fv >> 16), vl = __reduce_max_sync(0xFFFFFFFFu, fv & 0xFFFFu);
        const u32 nf = (u32)__popc(__ballot_sync(0xFFFFFFFFu, fits));
        if (lane == 0) { wlast[wid] = (vh << 16) | vl; wfit[wid] = nf; }
        __syncthreads();
        u32 F = start, run = 0;
#pragma unroll
        for (u32 w = 0; w < BK_NZ_WARPS; w++) { F = nz_hint_merge(F, wlast[w]); run += wfit[w]; }
        BK_NZ_PH(1);
        if (!capable || force_serial || run < BK_NZ_MIN_RUN) {
            // the sum is zero / tiny / negative, or about to leave any three zones within a few iterations: like the reference
            const u32 n_ser = min(n_it, force_serial ? (u32)BK_NZ_SERIAL : (capable ? max(run + 1u, (u32)BK_NZ_SERIAL_RUN) : (u32)BK_NZ_SERIAL_RUN));
            double keep = 0.0;
            for (u32 t = 0; t < n_ser; t++) {
#pragma unroll
                for (u32 q = 0; q < 6; q++) s = nz_add(s, xs[t * 6 + q]);
                if (t == tid) keep = s;
            }
            if (tid < n_ser) snap[it] = keep;
            a0 += n_ser; st_serial += n_ser; force_serial = 0; pf_a0 = 0xFFFFFFFFu;
            __syncthreads();                                                     // (xs is rewritten by the next pass)
            BK_NZ_PH(7);
            continue;
        }
        st_rounds++;
        NzZ2 z;
        nz2_zones_at(nz_hint_el(F), &z);
        const i64 S0 = nz2_start(z, sb);
        i64 A[BK_NZ_OPT]; u32 fc[BK_NZ_OPT]; bool ok[BK_NZ_OPT];
        i64 PA = 0, incl = 0;
        const bool warp_on = wid * 32 < n_it;                                    // warps without an iteration skip the work
        if (warp_on) {
#pragma unroll
            for (u32 q = 0; q < BK_NZ_OPT; q++) { ok[q] = nz2_split(z, x[q], &A[q], &fc[q]); PA += A[q]; }
            incl = PA;                                                           // scan 1: the A's
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) { const i64 g = __shfl_up_sync(0xFFFFFFFFu, incl, o); if (lane >= (u32)o) incl += g; }
        } else {
#pragma unroll
            for (u32 q = 0; q < BK_NZ_OPT; q++) { A[q] = 0; fc[q] = 0; ok[q] = true; }
        }
        if (lane == 31) wsumA[buf][wid] = incl;
        BK_NZ_PH(2);
        __syncthreads();
        // ... its operands and its approximate sum while this round computes (nothing waits for them before the round's end)
        if (tid < n_next) {
            it_n = it_p;
            const double* mo = maf0 + ((i64)it_p - BK_NOISE_WINDOW) * 3;
            const double* mn = maf0 + (i64)it_p * 3;
            hb_n = nz_b(snap[it_p]);
#pragma unroll
            for (u32 j = 0; j < 3; j++) { x_n[2 * j] = mo[j]; x_n[2 * j + 1] = mn[j]; }
        }
        i64 Pex = incl - PA;
#pragma unroll
        for (u32 w = 0; w < BK_NZ_WARPS - 1; w++) { const i64 v = wsumA[buf][w]; if (w < wid) Pex += v; }      // (unrolled: the loads issue together)
        NzThread T;
        NzVec f = nzvec_identity();
        NzVec ex = f;
        if (warp_on) {
            nz2_thread(A, fc, ok, S0 + Pex, tid * BK_NZ_OPT, &T);
            f = T.next;                                                          // scan 2: the class transitions
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) { const NzVec g = nzvec_shfl_up(f, o); if (lane >= (u32)o) f = nzvec_compose(g, f); }
            ex = nzvec_shfl_up(f, 1);
            if (lane == 0) ex = nzvec_identity();
        } else {
            T.bad = BK_NZ_OPT;
#pragma unroll
            for (u32 q = 0; q < BK_NZ_OPT; q++) { T.pr[q].lo = 0; T.pr[q].hi = 0; T.a_pre[q] = 0; }
        }
        if (lane == 31) { wnextL[buf][wid] = f.lo; wnextH[buf][wid] = f.hi; }
        BK_NZ_PH(3);
        __syncthreads();
        // the class in front of this thread: the transitions of the warps before it and of the lanes before it, applied
        // to the start class one after the other
        u32 cs = (u32)(u64)S0 & 7u;                                              // S mod 8
#pragma unroll
        for (u32 w = 0; w < BK_NZ_WARPS - 1; w++) { NzVec g; g.lo = wnextL[buf][w]; g.hi = wnextH[buf][w]; if (w < wid) cs = nzvec_get(g, cs); }
        cs = nzvec_get(ex, cs);
        // scan 3: the roundings, now that every thread knows the class it starts from
        const i32 rt = nz2_pr(T.pr[BK_NZ_OPT - 1], cs);
        i32 rincl = rt;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const i32 g = __shfl_up_sync(0xFFFFFFFFu, rincl, o); if (lane >= (u32)o) rincl += g; }
        if (lane == 31) wsumR[buf][wid] = rincl;
        __syncthreads();
        {   // ... and the scan of its look-ahead keys, published by the barriers that end this round
            u32 k2 = tid < n_next ? nz_hint_key(hb_n) : BK_NZ_HINT_NEUTRAL;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) { const u32 g = __shfl_up_sync(0xFFFFFFFFu, k2, o); if (lane >= (u32)o) k2 = nz_hint_merge(k2, g); }
            if (lane == 31) whint[buf ^ 1][wid] = k2;
            hk_n = k2;
        }
        i32 Rex = rincl - rt;
#pragma unroll
        for (u32 w = 0; w < BK_NZ_WARPS - 1; w++) { const i32 v = wsumR[buf][w]; if (w < wid) Rex += v; }
        i64 Tq[BK_NZ_OPT];
        u32 bad = T.bad;
#pragma unroll
        for (u32 q = 0; q < BK_NZ_OPT; q++) {
            Tq[q] = S0 + Pex + T.a_pre[q] + (i64)Rex + (i64)nz2_pr(T.pr[q], cs);
            if (!nz2_inside(Tq[q]) && bad > q) bad = q;
        }
        const u32 total_ops = n_it * 6;
        u32 mine = bad < BK_NZ_OPT ? min(tid * BK_NZ_OPT + bad, total_ops) : total_ops;     // (operations past the pass's end are zeros)
        mine = __reduce_min_sync(0xFFFFFFFFu, mine);
        if (lane == 0) wbad[buf][wid] = mine;
        __syncthreads();
        BK_NZ_PH(4);
        u32 n_ok = total_ops;
#pragma unroll
        for (u32 w = 0; w < BK_NZ_WARPS; w++) n_ok = min(n_ok, wbad[buf][w]);
        if (tid < n_it && tid * 6 + 5 < n_ok) snap[it] = nz2_value(z, Tq[5]);    // iterations whose six operations were all accepted
        if (n_ok > 0) {
            const u32 owner = (n_ok - 1) / BK_NZ_OPT, oq = (n_ok - 1) - owner * BK_NZ_OPT;
            if (tid == owner) {
                i64 Tl = Tq[0];
#pragma unroll
                for (u32 q = 1; q < BK_NZ_OPT; q++) if (q == oq) Tl = Tq[q];
                sstate[buf] = nz2_value(z, Tl);
            }
        }
#ifdef BK_NZ_WHY                                   // developer builds: why the accepted prefix ended → stats[8..12]
        if (n_ok < total_ops && tid == n_ok / BK_NZ_OPT && nv.stats) {
            const u32 q = n_ok % BK_NZ_OPT;
            const u32 why = !ok[q] ? 0u : (T.bad == q ? 1u : (Tq[q] <= (1ll << 52) ? 3u : 2u));      // operand too large / near a border / above / below
            atomicAdd(nv.stats + 8 + why, 1u); atomicAdd(nv.stats + 12, n_ok / 6);
        }
#endif
        __syncthreads();
        if (n_ok > 0) s = sstate[buf];
        BK_NZ_PH(5);
        if (n_ok == total_ops) { a0 += n_it; pf_a0 = a0; continue; }
        pf_a0 = 0xFFFFFFFFu;
        // the operation that ended the accepted prefix and the rest of its iteration, in real FP64
        st_stops++;
        const u32 ib = n_ok / 6, qb = n_ok - ib * 6;
        for (u32 q = qb; q < 6; q++) s = nz_add(s, xs[ib * 6 + q]);
        if (tid == ib) snap[it] = s;
        a0 += ib + 1;
        force_serial = ib < BK_NZ_SERIAL ? 1u : 0u;    // (the look-ahead saw a long run and was wrong: a few iterations like the reference)
        __syncthreads();                               // (xs is rewritten by the next pass)
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
#ifdef BK_NZ_NO_TABLES                              // developer builds: profile the chain blocks alone (results are wrong)
    return;
#endif
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
    // the sums after iteration i are the sums after the last ACTIVE iteration up to i (nz_chain_block), zero before the first
    const u32 rk = nv.act_rank[s.ibase + i];
    const u32 ia = rk ? nv.act_list[s.ibase + rk - 1] : 0u;
    const double s0 = rk ? nv.snap_s[s.ibase + ia] : 0.0, s20 = rk ? nv.snap_s2[s.ibase + ia] : 0.0;
    const double* mxp = nv.snap_tab + (size_t)(s.ibase + i) * BK_NOISE_TABLE;
    double mx[BK_NOISE_TABLE];
#pragma unroll
    for (int q = 0; q < BK_NOISE_TABLE; q += 2) { const double2 v = *reinterpret_cast<const double2*>(mxp + q); mx[q] = v.x; mx[q + 1] = v.y; }
    auto tau_of = [](u32 n) { return c_tau[n]; };
    nv.noise_max[s.r0 + i - BK_NOISE_HALF] = nz_tau_loop(cn0, s0, s20, mx, tau_of);
}
#endif  // __CUDACC__

}  // namespace bk
