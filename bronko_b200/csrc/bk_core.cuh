// bk_core.cuh — the per-read / per-k-mer logic of the counting stage, written once and compiled
// (a) by nvcc into the sm_100a kernels of bk_device.cu and (b) by g++ into tests/emul (a tests-only
// shared object that steps the same code on the CPU so the logic can be checked without a GPU; the
// product library never contains or calls a host build of this file).
//
// Counting replaces the KMC3 subprocess (reference src/call.rs:1152-1226; contract in SURVEY.md
// Appendix B): exact counts of non-canonical k-mers, reads split at non-ACGT symbols.
//
// Design: a read is seeded with ONE exact lookup of its first k-mer in the table of
// reference k-mers, then compared 32 bases at a time against the 2-bit packed oriented reference.
// Every run of consecutive k-mers that equal consecutive reference k-mers is counted with two
// atomics on a difference array (+1 at the first raw slot, -1 after the last); a prefix sum gives
// per-slot counts later.  K-mers that do not extend (sequencing errors, variants, foreign reads) are
// queued as (byte offset, count) stretches and counted one by one in the leftover kernel: exact
// reference k-mer → difference array, otherwise → open-addressing table of novel k-mers.
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define BK_HD __device__ __forceinline__
#define BK_COLD __device__ __noinline__          /* rare paths: kept out of line so the hot loops stay small */
#define BK_CLZLL(x) __clzll((long long)(x))
#define BK_POPCLL(x) __popcll((unsigned long long)(x))
#define BK_POPC(x) __popc((unsigned)(x))
#define BK_FFS0(x) ((u32)__ffs((int)(x)) - 1u)
#define BK_ANY(p) __any_sync(0xFFFFFFFFu, (p))
#define BK_WARP_MAX(x) __reduce_max_sync(0xFFFFFFFFu, (unsigned)(x))
#define BK_SYNCWARP() __syncwarp()
#define BK_FUNNEL_R(lo, hi, sh) __funnelshift_r((lo), (hi), (sh))
#define BK_PRMT(x, y, sel) __byte_perm((x), (y), (sel))
#else
#define BK_HD static inline
#define BK_COLD static
#define BK_CLZLL(x) __builtin_clzll((unsigned long long)(x))
#define BK_POPCLL(x) __builtin_popcountll((unsigned long long)(x))
#define BK_POPC(x) __builtin_popcount((unsigned)(x))
#define BK_FFS0(x) ((u32)__builtin_ctz((unsigned)(x)))
#define BK_ANY(p) (p)
#define BK_WARP_MAX(x) ((unsigned)(x))
#define BK_SYNCWARP() do {} while (0)
#define BK_FUNNEL_R(lo, hi, sh) ((sh) ? (((lo) >> (sh)) | ((hi) << (32 - (sh)))) : (lo))
static inline unsigned BK_PRMT(unsigned x, unsigned y, unsigned sel) {      // selector nibbles 0..7 only
    const unsigned long long src = ((unsigned long long)y << 32) | x;
    unsigned r = 0;
    for (int i = 0; i < 4; i++) r |= (unsigned)((src >> (8 * ((sel >> (4 * i)) & 7))) & 0xFF) << (8 * i);
    return r;
}
struct uint2 { unsigned int x, y; };
static inline uint2 make_uint2(unsigned int x, unsigned int y) { uint2 r; r.x = x; r.y = y; return r; }
#endif

namespace bk {

typedef uint64_t u64;
typedef uint32_t u32;
typedef int32_t i32;
typedef uint8_t u8;

static const u64 BK_EMPTY = ~0ull;

struct GenSlot { u64 key; u32 cnt; u32 pad; };         // novel k-mer table slot, 16 B
struct W4 { u32 x, y, z, w; };                         // one 16-byte chunk (32 bases) of the 4-bit reference
struct ExactSlotD { u64 key; u32 gidx; u32 oseq; };    // reference k-mer table slot, 16 B

BK_HD u32 hash_slot(u64 x, u32 shift) { return (u32)(((x ^ (x >> 31)) * 0x9E3779B97F4A7C15ull) >> shift); }

// --- 4 ASCII bases (one little-endian u32, first base in the low byte) → 8 packed bits, first base
// in the two high bits, code A=0 C=1 G=2 T=3 (reference src/lcb.rs:47-55, lower case accepted);
// *err != 0 iff some byte is not one of ACGTacgt.
BK_HD u32 pack4(u32 w, u32* err) {
    const u32 u = w & 0xDFDFDFDFu;
    const u32 a = u >> 1, b = u >> 2, c = u >> 4;
    const u32 y = (a ^ b) & 0x03030303u;
    const u32 e1 = c ^ (b & ~a);
    const u32 e2 = e1 | ~(u ^ c);
    *err = (e2 & 0x01010101u) | ((u ^ 0x40404040u) & 0xC8C8C8C8u);
    return (y * 0x40100401u) >> 24;
}

// Word source: aligned u32 words of the ASCII base buffer (shared-memory tile or global memory).
// ld(i) returns word i of the buffer the byte offsets refer to.
template <class Ld>
BK_HD u32 load_unaligned(const Ld& ld, u32 byte_off) {
    const u32 wi = byte_off >> 2, sh = (byte_off & 3) * 8;
    const u32 lo = ld(wi);
    if (sh == 0) return lo;
    const u32 hi = ld(wi + 1);
    return (lo >> sh) | (hi << (32 - sh));
}

// 32 bases starting at byte_off, all of them inside the read → 64 packed bits MSB-first; *bad != 0 iff
// one of the 32 bytes is not ACGT.  Branch-free: 9 aligned word loads, 8 funnel shifts, 8 pack4.
// (Reads up to 7 bytes past the 32 bases: buffers carry >= 16 bytes of slack.)
template <class Ld>
BK_HD u64 pack32_full(const Ld& ld, u32 byte_off, u32* bad) {
    const u32 wi = byte_off >> 2, sh = (byte_off & 3) * 8;
    u32 w[9];
#pragma unroll
    for (u32 i = 0; i < 9; i++) w[i] = ld(wi + i);
    u32 hi32 = 0, lo32 = 0, acc = 0;
#pragma unroll
    for (u32 i = 0; i < 8; i++) {
        u32 err;
        const u32 p = pack4(BK_FUNNEL_R(w[i], w[i + 1], sh), &err);
        acc |= err;
        if (i < 4) hi32 |= p << (24 - 8 * i); else lo32 |= p << (56 - 8 * i);
    }
    *bad = acc;
    return ((u64)hi32 << 32) | lo32;
}

// Same for the last word of a read: nb (1..31) bases belong to the read; bits of bases >= nb are zero
// and bytes past the read never count as bad.
template <class Ld>
BK_HD u64 pack32_tail(const Ld& ld, u32 byte_off, u32 nb, u32* bad) {
    const u32 wi = byte_off >> 2, sh = (byte_off & 3) * 8;
    u32 prev = ld(wi);
    u32 hi32 = 0, lo32 = 0, acc = 0;
    const u32 nwords = (nb + 3) >> 2;
    for (u32 i = 0; i < nwords; i++) {
        const u32 nxt = ld(wi + i + 1);
        const u32 w = BK_FUNNEL_R(prev, nxt, sh);
        prev = nxt;
        u32 err;
        u32 p = pack4(w, &err);
        const u32 vb = nb - 4 * i;
        if (vb < 4) { err &= (1u << (8 * vb)) - 1u; p &= 0xFFu << (8 - 2 * vb); }
        acc |= err;
        if (i < 4) hi32 |= p << (24 - 8 * i); else lo32 |= p << (56 - 8 * i);
    }
    *bad = acc;
    return ((u64)hi32 << 32) | lo32;
}

template <class Ld>
BK_HD u64 pack32(const Ld& ld, u32 byte_off, u32 nb, u32* bad) {
    return nb >= 32 ? pack32_full(ld, byte_off, bad) : pack32_tail(ld, byte_off, nb, bad);
}

// k bases starting at byte_off → k-mer value (first base most significant, src/lcb.rs:67-74);
// returns false if a byte is not ACGT.
template <class Ld>
BK_HD bool pack_kmer(const Ld& ld, u32 byte_off, u32 k, u64* out) {
    u32 bad;
    const u64 v = pack32_tail(ld, byte_off, k, &bad);
    *out = v >> (64 - 2 * k);
    return bad == 0;
}
// Same when at least 32 bytes are readable from byte_off on (the branch-free packer; if any of the 32 bytes is
// not ACGT fall back to the exact k-byte version, the offending byte may lie beyond the k-mer).
template <class Ld>
BK_HD bool pack_kmer32(const Ld& ld, u32 byte_off, u32 k, u64* out) {
    u32 bad;
    const u64 v = pack32_full(ld, byte_off, &bad);
    if (bad) return pack_kmer(ld, byte_off, k, out);
    *out = v >> (64 - 2 * k);
    return true;
}

// Device-side view of everything the counting stage touches.
struct CountView {
    u32 k;
    const u32* refnib; u32 ref_chunks;    // oriented reference, one 4-bit code per base index (A0 C1 G2 T3, padding 4),
                                          // little-endian nibbles, in 16-byte chunks of 32 bases
    const u32* oseq_start; const u32* oseq_len;
    const ExactSlotD* exact; u32 exact_shift, exact_mask;
    u32* diff;                       // n_raw + 2
    GenSlot* gen; u32 gen_shift, gen_mask;
    u32* gen_full;                   // set to 1 if the novel table / list ran out of room
    u64* nov; u32 nov_cap; u32* nov_n;   // list mode (nov != null): novel k-mer occurrences are appended here instead of
                                         // being counted in `gen`; bk_bins.cuh counts the list afterwards
    u32* nov_w;                          // weights of the list entries (null: every entry counts once)
    uint2* desc; u32 desc_cap; u32* n_desc;
    // Mismatch lines (null = off): k-mers with exactly ONE mismatch against the diagonal their read follows are not
    // listed one by one.  A mismatch of read base e (reference base index r = g0 + e) to base b is seen by the k-mers
    // starting at read bases e - j, j = 0 .. k-1 (j = where the mismatch sits inside the k-mer); those that hold no other
    // bad base form a range jlo .. jhi and are counted with two atomics on line (r, b): dense[(r * 4 + b) * (k + 1) + jlo]
    // += 1, [.. + jhi + 1] -= 1 — a difference array along j, like `diff` along the diagonal.  Cell (line, j) after the
    // prefix sum = occurrences of the k-mer "reference k-mer at raw slot r - j with digit j replaced by b".
    u32* dense; u8* dense_flag;          // lines x (k + 1) counters; one byte per line: touched
};

#if defined(__CUDACC__)
BK_HD void add_u32(u32* p, u32 v) { atomicAdd(p, v); }
BK_HD u32 fetch_add_u32(u32* p, u32 v) { return atomicAdd(p, v); }
BK_HD u64 cas_u64(u64* p, u64 cmp, u64 val) { return atomicCAS((unsigned long long*)p, (unsigned long long)cmp, (unsigned long long)val); }
BK_HD u64 load_key(const u64* p) { return *(const volatile u64*)p; }
BK_HD ExactSlotD load_exact(const ExactSlotD* p) {
    const uint4 v = __ldg((const uint4*)p);
    ExactSlotD s; s.key = ((u64)v.y << 32) | v.x; s.gidx = v.z; s.oseq = v.w; return s;
}
#else
BK_HD void add_u32(u32* p, u32 v) { *p += v; }
BK_HD u32 fetch_add_u32(u32* p, u32 v) { u32 o = *p; *p += v; return o; }
BK_HD u64 cas_u64(u64* p, u64 cmp, u64 val) { u64 o = *p; if (o == cmp) *p = val; return o; }
BK_HD u64 load_key(const u64* p) { return *p; }
BK_HD ExactSlotD load_exact(const ExactSlotD* p) { return *p; }
#endif

BK_HD bool exact_lookup(const CountView& v, u64 kmer, u32* gidx, u32* oseq) {
    u32 h = hash_slot(kmer, v.exact_shift);
    for (;;) {
        const ExactSlotD s = load_exact(v.exact + h);
        if (s.key == kmer) { *gidx = s.gidx; *oseq = s.oseq; return true; }
        if (s.key == BK_EMPTY) return false;
        h = (h + 1) & v.exact_mask;
    }
}

// Count one k-mer occurrence that did not extend a run.  Returns 1 if it created a new novel key.
BK_HD u32 count_one(const CountView& v, u64 kmer) {
    u32 gidx, oseq;
    if (exact_lookup(v, kmer, &gidx, &oseq)) {
        add_u32(v.diff + gidx, 1u);
        add_u32(v.diff + gidx + 1, 0xFFFFFFFFu);
        return 0;
    }
    if (v.nov) {                                         // list mode, rare paths only (one atomic per k-mer)
        const u32 pos = fetch_add_u32(v.nov_n, 1u);
        if (pos < v.nov_cap) { v.nov[pos] = kmer; if (v.nov_w) v.nov_w[pos] = 1u; } else *v.gen_full = 1;
        return 0;
    }
    u32 h = hash_slot(kmer, v.gen_shift);
    for (u32 probe = 0; probe <= v.gen_mask; probe++) {
        u64 cur = load_key(&v.gen[h].key);
        if (cur == BK_EMPTY) {
            cur = cas_u64(&v.gen[h].key, BK_EMPTY, kmer);
            if (cur == BK_EMPTY) { add_u32(&v.gen[h].cnt, 1u); return 1; }
        }
        if (cur == kmer) { add_u32(&v.gen[h].cnt, 1u); return 0; }
        h = (h + 1) & v.gen_mask;
    }
    *v.gen_full = 1;
    return 0;
}

// Count the k-mers starting at bytes [byte_off, byte_off+cnt) one at a time, validating every base
// (k-mers that contain a non-ACGT byte are not counted: KMC splits reads there).  step/first let a
// warp interleave lanes.  Returns the number of new novel keys.
template <class Ld>
BK_COLD u32 count_stretch(const CountView& v, const Ld& ld, u32 byte_off, u32 cnt, u32 first, u32 step) {
    u32 created = 0;
    for (u32 j = first; j < cnt; j += step) {
        u64 km;
        if (pack_kmer(ld, byte_off + j, v.k, &km)) created += count_one(v, km);
    }
    return created;
}

// Leftover stretches of one read wait here until the caller flushes them (one warp-aggregated atomic
// per warp on the device).  More than two per read are rare and go to the queue directly.
struct Pending { u32 n; uint2 d0, d1; };

BK_HD void queue_push(const CountView& v, uint2 d, u32* overflow) {
    const u32 slot = fetch_add_u32(v.n_desc, 1u);
    if (slot < v.desc_cap) v.desc[slot] = d; else *overflow = 1;
}

// Leftover stretch: k-mers starting at read bases [a, a+cnt) of the read at byte offset o0 of the word
// source, which itself starts gofs bytes into the pushed buffer (descriptors use buffer offsets).
// Returns new novel keys if the queue was full and the stretch had to be counted in place.
template <class Ld>
BK_HD u32 emit_leftover(const CountView& v, const Ld& ld, u32 o0, u32 a, u32 cnt, u32 gofs, Pending& pend) {
    const uint2 d = make_uint2(gofs + o0 + a, cnt);
    if (pend.n == 0) { pend.d0 = d; pend.n = 1; return 0; }
    if (pend.n == 1) { pend.d1 = d; pend.n = 2; return 0; }
    u32 overflow = 0;
    queue_push(v, d, &overflow);
    return overflow ? count_stretch(v, ld, o0 + a, cnt, 0, 1) : 0;     // queue full: count in place (slow, exact)
}

BK_HD void emit_run(const CountView& v, i32 g0, u32 a, u32 cnt) {
    add_u32(v.diff + (u32)(g0 + (i32)a), 1u);
    add_u32(v.diff + (u32)(g0 + (i32)a) + cnt, 0xFFFFFFFFu);
}

// one read byte → 2-bit code (A0 C1 G2 T3, either case); false if it is not one of ACGTacgt
template <class Ld>
BK_HD bool base_code(const Ld& ld, u32 byte_off, u32* code) {
    const u32 up = ((ld(byte_off >> 2) >> (8 * (byte_off & 3))) & 0xFFu) & 0xDFu;
    *code = ((up >> 1) ^ (up >> 2)) & 3u;
    return up == 'A' || up == 'C' || up == 'G' || up == 'T';
}
BK_HD void emit_dense(const CountView& v, u32 refpos, u32 alt, u32 jlo, u32 jhi) {
    const u32 line = refpos * 4u + alt;
    u32* row = v.dense + (size_t)line * (v.k + 1);
    add_u32(row + jlo, 1u);
    add_u32(row + jhi + 1, 0xFFFFFFFFu);
    v.dense_flag[line] = 1;
}

#ifndef BK_MAX_SEEDS
#define BK_MAX_SEEDS 4      // seed attempts (spaced k apart) per diagonal search
#endif
#ifndef BK_MAX_DIAGS
#define BK_MAX_DIAGS 3      // diagonals tried per read (re-seed after an indel / wrong diagonal)
#endif
#ifndef BK_BAIL_MISMATCHES
#define BK_BAIL_MISMATCHES 8  // this many bad bases inside one 32-base word ends the diagonal
#endif

// Seed k-mer: k bytes at byte_off → k-mer value WITHOUT validation.  Bits 1..2 of an ASCII base are a 2-bit code
// (A 00, C 01, T 10, G 11; the same for lower case), gathered four at a time by a multiply and turned into the
// A0 C1 G2 T3 code by one xor.  Any other byte yields some code: a seed found through it is harmless, because the
// extension below compares every byte of the read with the letter the reference expects.
// Reads at most ceil(k/4) + 1 words (<= 15 bytes past the k-mer: inside the slack every buffer carries).
template <class Ld>
BK_HD u64 pack_seed(const Ld& ld, u32 byte_off, u32 k) {
    const u32 wi = byte_off >> 2, sh = (byte_off & 3) * 8;
    u32 hi32 = 0, lo32 = 0;
    u32 prev = ld(wi);
#pragma unroll
    for (u32 i = 0; i < 8; i++) {
        if (4 * i < k) {
            const u32 nxt = ld(wi + i + 1);
            const u32 y = (BK_FUNNEL_R(prev, nxt, sh) >> 1) & 0x03030303u;
            const u32 pk = (y * 0x40100401u) >> 24;                // first base in the two high bits
            if (i < 4) hi32 |= pk << (24 - 8 * i); else lo32 |= pk << (56 - 8 * i);
            prev = nxt;
        }
    }
    u64 v = (((u64)hi32 << 32) | lo32) >> (64 - 2 * k);
    return v ^ ((v >> 1) & 0x5555555555555555ull);                 // T 10 <-> G 11
}

// The 32 reference letters expected at global base index g .. g+31, as eight words of four upper-case ASCII letters
// ('#' in the padding): two 16-byte chunks of the 4-bit reference, aligned, byte-permuted.  cur / nxt are the chunks
// holding base g and the one after it.
BK_HD void ref_letters(const W4& cur, const W4& nxt, u32 g, u32* ex8) {
    const u32 wo = (g >> 3) & 3, bs = (g & 7) * 4;
    u32 d0 = cur.x, d1 = cur.y, d2 = cur.z, d3 = cur.w, d4 = nxt.x, d5 = nxt.y, d6 = nxt.z, d7 = nxt.w;
    if (wo & 1) { d0 = d1; d1 = d2; d2 = d3; d3 = d4; d4 = d5; d5 = d6; d6 = d7; }
    if (wo & 2) { d0 = d2; d1 = d3; d2 = d4; d3 = d5; d4 = d6; }
    const u32 e0 = BK_FUNNEL_R(d0, d1, bs), e1 = BK_FUNNEL_R(d1, d2, bs), e2 = BK_FUNNEL_R(d2, d3, bs), e3 = BK_FUNNEL_R(d3, d4, bs);
    ex8[0] = BK_PRMT(0x54474341u, 0x23232323u, e0); ex8[1] = BK_PRMT(0x54474341u, 0x23232323u, e0 >> 16);
    ex8[2] = BK_PRMT(0x54474341u, 0x23232323u, e1); ex8[3] = BK_PRMT(0x54474341u, 0x23232323u, e1 >> 16);
    ex8[4] = BK_PRMT(0x54474341u, 0x23232323u, e2); ex8[5] = BK_PRMT(0x54474341u, 0x23232323u, e2 >> 16);
    ex8[6] = BK_PRMT(0x54474341u, 0x23232323u, e3); ex8[7] = BK_PRMT(0x54474341u, 0x23232323u, e3 >> 16);
}
template <class LdRef4>
BK_HD W4 ref_chunk(const LdRef4& ldr4, i32 c, i32 cmax) { return ldr4((u32)c < (u32)cmax ? (u32)c : (u32)cmax); }   // (words inside the overlap never clamp)

// Fast filter: are the 32 read bytes at byte offset rb exactly the upper-case letters ex8?  (No case folding: a
// lower-case read only fails the filter and is then judged by word_mask().)
template <class Ld>
BK_HD u32 word_differs(const Ld& ld, u32 rb, const u32* ex8) {
    const u32 rwi = rb >> 2, rsh = (rb & 3) * 8;
    u32 acc = 0;
    u32 rp = ld(rwi);
#pragma unroll
    for (u32 i = 0; i < 8; i++) {
        const u32 rn = ld(rwi + i + 1);
        acc |= BK_FUNNEL_R(rp, rn, rsh) ^ ex8[i];
        rp = rn;
    }
    return acc;
}

// Exact judgement of one word: bit j set iff base j of the word (j in [lo, hi)) is NOT the expected letter in either
// case — a mismatch, a non-ACGT byte, or reference padding.
template <class Ld>
BK_HD u32 word_mask(const Ld& ld, u32 rb, const u32* ex8, i32 lo, i32 hi) {
    const u32 rwi = rb >> 2, rsh = (rb & 3) * 8;
    u32 t = 0;
    u32 rp = ld(rwi);
#pragma unroll
    for (u32 i = 0; i < 8; i++) {
        const u32 rn = ld(rwi + i + 1);
        const u32 x = (BK_FUNNEL_R(rp, rn, rsh) & 0xDFDFDFDFu) ^ ex8[i];
        rp = rn;
        u32 f = x | (x >> 4);
        f |= f >> 2;
        f |= f >> 1;
        f &= 0x01010101u;                                    // byte flags
        t |= ((f * 0x10204080u) >> 28) << (4 * i);           // 4 flags → 4 bits, first base lowest
    }
    u32 keep = 0xFFFFFFFFu;
    if (lo > 0) keep = lo >= 32 ? 0u : keep << lo;
    if (hi < 32) keep = hi <= 0 ? 0u : keep & (0xFFFFFFFFu >> (32 - hi));
    return t & keep;
}

// One read: bytes [o0, o0+len) of the word source.  On the device EVERY lane of the warp must call
// this (lanes without a read pass len = 0): the loops run a warp-uniform number of rounds and
// reconverge after every round, so a mismatch in one lane's read does not split the warp for the
// rest of the read.  Returns the number of novel keys created by in-place fallbacks (normally 0).
//
// State: k-mers are indexed by their start base.  `c` = first k-mer start not classified yet (every
// k-mer below c has been put in a run or a leftover stretch), `ms` = start of the current stretch of
// matching bases on the current diagonal.  A bad base at e (mismatch, non-ACGT byte, end of the
// overlap with the oriented sequence, end of read) closes the stretch [ms, e): if it holds >= k bases
// its k-mers [ms, e-k] become a run, everything pending before ms a leftover stretch.
//
// Extension: the words of the read (32 bases on the read's own grid) that lie fully inside the overlap with the
// oriented sequence go through word_differs(), a branch-free filter every lane executes; it only records WHICH words
// are off.  The partial last word is filtered through the full word that ends with it.  Words that failed the filter
// (and partial words that cannot be filtered) are then judged one by one with word_mask() and turned into events —
// the only divergent part, a few rounds per warp.
template <class Ld, class LdRef4>
BK_HD u32 scan_read(const CountView& v, const Ld& ld, const LdRef4& ldr4, u32 o0, u32 len, u32 gofs, Pending& pend) {
    const u32 k = v.k;
    const bool has = len >= k;                   // shorter than k: contributes no k-mer
    const u32 nk = has ? len - k + 1 : 0;
    u32 created = 0;
    i32 c = 0;
    i32 seed_from = 0;
    bool active = has;                           // still looking for / following a diagonal
    // mismatch lines: the last bad base of the current diagonal whose one-mismatch k-mers are not classified yet
    // (-1: none) and the first k-mer start that may hold it alone
    const bool dense_on = v.dense != nullptr && len < 32768u;
    i32 pe = -1, pe_lo = 0;
    const i32 cmax = (i32)v.ref_chunks - 1;
    for (u32 diag = 0; diag < BK_MAX_DIAGS; diag++) {
        if (!BK_ANY(active)) break;
        // ---- seed: first k-mer at or after seed_from that is an exact reference k-mer ----
        u32 gidx = 0, oseq = 0, q = (u32)seed_from;
        bool found = false;
        for (u32 t = 0; t < BK_MAX_SEEDS; t++) {
            const bool need = active && !found && q + k <= len;
            if (!BK_ANY(need)) break;
            if (need) {
                if (exact_lookup(v, pack_seed(ld, o0 + q, k), &gidx, &oseq)) found = true;
                else q += k;
            }
            BK_SYNCWARP();
        }
        if (!found) active = false;              // no seed: what is left of the read is leftover
        // ---- extend along the diagonal ----
        i32 g0 = 0, i_lo = 0, i_hi = 0, ms = 0, w = 0, w_end = 0;
        if (active) {
            g0 = (i32)gidx - (i32)q;             // global base index of read base 0 on this diagonal
            const i32 os = (i32)v.oseq_start[oseq], oe = os + (i32)v.oseq_len[oseq];
            i_lo = os > g0 ? os - g0 : 0;                              // read bases inside the
            i_hi = (oe - g0) < (i32)len ? (oe - g0) : (i32)len;        // oriented sequence
            if (i_lo < seed_from) i_lo = seed_from;
            ms = i_lo;
            w = i_lo >> 5;
            w_end = (i_hi + 31) >> 5;
        }
        bool bailed = false;
        pe = -1;                                 // (a mismatch left pending by a diagonal that was given up stays a leftover)
        // A bad base at e (real_: a mismatch / non-ACGT byte inside the overlap; else the end of the overlap).  First the
        // pending mismatch pe is settled: the k-mers that hold pe and no other bad base start in [pe_lo, min(pe, e - k)];
        // if pe is a clean substitution they go to its mismatch line, and what was pending before them to a leftover
        // stretch.  K-mers holding two bad bases, a non-ACGT byte or bases outside the overlap stay leftovers.
#define BK_EVENT(e_, real_)                                                                 \
    do {                                                                                    \
        const i32 e__ = (e_);                                                               \
        if (pe >= 0) {                                                                      \
            const i32 hi__ = (e__ - (i32)k) < pe ? (e__ - (i32)k) : pe;                     \
            u32 code__;                                                                     \
            if (pe_lo <= hi__ && base_code(ld, o0 + (u32)pe, &code__)) {                    \
                if (c < pe_lo) created += emit_leftover(v, ld, o0, (u32)c, (u32)(pe_lo - c), gofs, pend); \
                emit_dense(v, (u32)(g0 + pe), code__, (u32)(pe - hi__), (u32)(pe - pe_lo)); \
                c = hi__ + 1;                                                               \
            }                                                                               \
            pe = -1;                                                                        \
        }                                                                                   \
        if (e__ - ms >= (i32)k) {                                                           \
            if (c < ms) created += emit_leftover(v, ld, o0, (u32)c, (u32)(ms - c), gofs, pend); \
            emit_run(v, g0, (u32)ms, (u32)(e__ - (i32)k + 1 - ms));                         \
            c = e__ - (i32)k + 1;                                                           \
        }                                                                                   \
        if (dense_on && (real_)) { pe = e__; pe_lo = (e__ - (i32)k + 1) > ms ? (e__ - (i32)k + 1) : ms; } \
        ms = e__ + 1;                                                                       \
    } while (0)
        for (;;) {                               // segments of up to 32 words
            const bool seg = active && !bailed && w < w_end;
            if (!BK_ANY(seg)) break;
            const i32 seg_w0 = w;
            const i32 seg_end = seg ? (w_end < w + 32 ? w_end : w + 32) : w;
            u32 wm = 0;                          // bit (ww - seg_w0): word ww must be judged by word_mask()
            // -- filter pass: one reference chunk per word, the previous one is kept --
            i32 c4 = (g0 + 32 * w) >> 5;
            W4 cur = {0, 0, 0, 0};
            if (seg) cur = ref_chunk(ldr4, c4, cmax);
            const u32 n_rounds = BK_WARP_MAX(seg_end - w);          // warp-uniform trip count, lanes with fewer words idle
            for (u32 it = 0; it < n_rounds; it++) {
                const bool go = seg && w < seg_end;
                if (go) {
                    const i32 b0 = 32 * w;
                    const W4 nxt = ref_chunk(ldr4, c4 + 1, cmax);
                    if (b0 >= i_lo && b0 + 32 <= i_hi) {               // all 32 bases belong to the overlap
                        u32 ex8[8];
                        ref_letters(cur, nxt, (u32)(g0 + b0), ex8);
                        if (word_differs(ld, o0 + (u32)b0, ex8)) wm |= 1u << (w - seg_w0);
                    } else if (b0 >= i_lo && i_hi - 32 >= i_lo) {      // partial last word: the full word ending with it
                        const i32 bt = i_hi - 32;                      // (b0 - 32 < bt < b0: its bases lie in cur / nxt
                        u32 ex8[8];                                    //  of the previous word and in this word's)
                        const i32 ct = (g0 + bt) >> 5;
                        const W4 ca = ref_chunk(ldr4, ct, cmax), cb = ref_chunk(ldr4, ct + 1, cmax);
                        ref_letters(ca, cb, (u32)(g0 + bt), ex8);
                        if (word_differs(ld, o0 + (u32)bt, ex8)) wm |= 1u << (w - seg_w0);
                    } else wm |= 1u << (w - seg_w0);                   // partial and not filterable
                    cur = nxt;
                    c4++;
                    w++;
                }
            }
            // -- judge the words that failed, in order --
            for (;;) {
                const bool go = wm != 0;
                if (!BK_ANY(go)) break;
                if (go) {
                    const i32 ww = seg_w0 + (i32)BK_FFS0(wm);
                    wm &= wm - 1;
                    const i32 b0 = 32 * ww;
                    const i32 cw = (g0 + b0) >> 5;
                    const W4 ca = ref_chunk(ldr4, cw, cmax), cb = ref_chunk(ldr4, cw + 1, cmax);
                    u32 ex8[8];
                    ref_letters(ca, cb, (u32)(g0 + b0), ex8);
                    u32 t = word_mask(ld, o0 + (u32)b0, ex8, i_lo - b0, i_hi - b0);
                    if (BK_POPC(t) >= BK_BAIL_MISMATCHES) {            // wrong diagonal from here on: close, re-seed
                        const i32 e = b0 + (i32)BK_FFS0(t);
                        BK_EVENT(e, false);                            // (wrong diagonal: nothing around e is a one-mismatch k-mer of it)
                        seed_from = e + 1;
                        bailed = true;
                        wm = 0;
                    } else {
                        while (t) {
                            const u32 j = BK_FFS0(t);
                            t &= t - 1;
                            BK_EVENT(b0 + (i32)j, true);
                        }
                    }
                }
                BK_SYNCWARP();
            }
        }
        if (active && !bailed) { BK_EVENT(i_hi, false); active = false; }
#undef BK_EVENT
    }
    if (has && (i32)nk > c) created += emit_leftover(v, ld, o0, (u32)c, nk - (u32)c, gofs, pend);
    return created;
}

}  // namespace bk
