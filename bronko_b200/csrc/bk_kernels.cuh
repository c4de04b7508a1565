// bk_kernels.cuh — sm_100a kernels of the k-mer→pileup path.  Included only by bk_device.cu.
//
//  counting  : k_scan (pack + seed/extend, difference-array runs), k_leftover (per-k-mer counting),
//              k_diff_* (prefix sum + fold onto distinct reference k-mers), k_compact_* (the "KMC dump")
//  mapping   : k_map<STATS|PILEUP>   — reference src/call.rs:1257-1434
//  selection : k_select               — src/call.rs:422-502
//  scoring   : k_noise — src/call.rs:799-967 ; k_call — src/call.rs:969-1150
#pragma once
#include <cuda_runtime.h>

#include "bk_core.cuh"

namespace bk {

struct BucketSlotD { u64 key; u32 off; u32 len; };
struct BucketEntryD { u32 row; unsigned short file_id; u8 idx; u8 canonical; };

struct FileCounters { u32 n_counted; u32 gen_new; u32 unique; u32 nov_n; u64 total_kmers; u32 n_desc;
                      u32 n_dense; };   // n_dense: counted k-mers kept at the END of the counted list (mapped where they were counted, bk_dense.cuh)
struct Counters {
    u32 gen_full; u32 var_overflow; u32 pad0; u32 pad1;
    FileCounters f[2];
    i32 best; u32 n_var; u32 n_major; u32 n_minor;
    u64 pos_covered; u64 total_cov;
};

}  // namespace bk
#include "bk_noise.cuh"
namespace bk {

__device__ __forceinline__ u32 warp_sum_u32(u32 v) {
#pragma unroll
    for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xFFFFFFFFu, v, o);
    return v;
}
__device__ __forceinline__ u64 warp_sum_u64(u64 v) {
#pragma unroll
    for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xFFFFFFFFu, v, o);
    return v;
}
// slot for this lane in an append-only list; every lane of the warp must call it
__device__ __forceinline__ u32 warp_append(u32* counter, bool pred) {
    const u32 m = __ballot_sync(0xFFFFFFFFu, pred);
    if (m == 0) return 0;
    const u32 lane = threadIdx.x & 31, leader = __ffs(m) - 1;
    u32 base = 0;
    if (lane == leader) base = atomicAdd(counter, __popc(m));
    base = __shfl_sync(0xFFFFFFFFu, base, leader);
    return base + __popc(m & ((1u << lane) - 1));
}

// ------------------------------------------------------------------------------------------------
// k_scan: one thread per read.  A CTA stages the contiguous ASCII bytes of its tile of reads into
// shared memory with coalesced 16-byte loads, then every thread packs / seeds / extends its read
// from shared memory (bk_core.cuh: scan_read).  Tiles whose bytes do not fit (very long reads) are
// read straight from global memory.
//   bases + off_bias-relative offsets: read r = bytes [off[r]-off_bias, off[r+1]-off_bias) of `bases`.
// ------------------------------------------------------------------------------------------------
#define BK_SCAN_THREADS 256

// Flush the leftover stretches the lanes of this warp collected for their reads: one atomic per warp
// reserves the queue slots.  If the queue is full the stretch is counted in place (slow, still exact).
template <class Ld>
__device__ __forceinline__ u32 flush_pending(const CountView& v, const Ld& ld, const Pending& pend, u32 gofs) {
    const u32 lane = threadIdx.x & 31;
    u32 incl = pend.n;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const u32 t = __shfl_up_sync(0xFFFFFFFFu, incl, o); if (lane >= (u32)o) incl += t; }
    const u32 total = __shfl_sync(0xFFFFFFFFu, incl, 31);
    if (total == 0) return 0;
    u32 base = 0;
    if (lane == 31) base = atomicAdd(v.n_desc, total);
    base = __shfl_sync(0xFFFFFFFFu, base, 31) + incl - pend.n;
    u32 created = 0;
    if (pend.n > 0) { if (base < v.desc_cap) v.desc[base] = pend.d0; else created += count_stretch(v, ld, pend.d0.x - gofs, pend.d0.y, 0, 1); }
    if (pend.n > 1) { if (base + 1 < v.desc_cap) v.desc[base + 1] = pend.d1; else created += count_stretch(v, ld, pend.d1.x - gofs, pend.d1.y, 0, 1); }
    return created;
}

// TMA bulk copy (cp.async.bulk, 1-D) of one tile global → shared, completion on an mbarrier: one thread issues it,
// no thread spends instructions on moving bytes, and the same thread asks for the CTA's next tile to be pulled into L2.
__device__ __forceinline__ u32 smem_addr(const void* p) { return (u32)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long* bar, u32 count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_addr(bar)), "r"(count));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void tile_load(void* dst, const void* src, u32 bytes, unsigned long long* bar) {
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");      // generic-proxy reads of the old tile are done (barrier before)
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_addr(bar)), "r"(bytes) : "memory");
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_addr(dst)), "l"(src), "r"(bytes), "r"(smem_addr(bar)) : "memory");
}
__device__ __forceinline__ void tile_prefetch_l2(const void* src, u32 bytes) {
    asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(src), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, u32 parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "BK_WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra BK_DONE_%=;\n"
        "bra BK_WAIT_%=;\n"
        "BK_DONE_%=:\n"
        "}\n" ::"r"(smem_addr(bar)), "r"(parity) : "memory");
}

__global__ void __launch_bounds__(BK_SCAN_THREADS, 4)
k_scan(CountView v, const u8* __restrict__ bases, const u32* __restrict__ off, u32 off_bias, u32 r_begin, u32 r_end,
       u32 tile_reads, u32 tile_bytes, u32* gen_new) {
    extern __shared__ __align__(128) u8 smem[];
    __shared__ __align__(8) unsigned long long tile_bar;
    const u32 n_tiles = (r_end - r_begin + tile_reads - 1) / tile_reads;
    const u32* refnib = v.refnib;
    auto ldr4 = [refnib](u32 i4) { const uint4 q = __ldg(reinterpret_cast<const uint4*>(refnib) + i4); W4 r; r.x = q.x; r.y = q.y; r.z = q.z; r.w = q.w; return r; };
    if (threadIdx.x == 0) mbar_init(&tile_bar, 1);
    __syncthreads();
    u32 parity = 0;
    u32 created = 0;
    for (u32 tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const u32 r0 = r_begin + tile * tile_reads;
        const u32 r1 = min(r0 + tile_reads, r_end);
        const u32 a = (__ldg(off + r0) - off_bias) & ~15u;
        const u32 e = (__ldg(off + r1) - off_bias + 15u) & ~15u;
        const bool staged = (e - a) <= tile_bytes;          // uniform over the CTA
        const u32 r = r0 + threadIdx.x;
        u32 o0 = 0, len = 0;                                 // lanes without a read scan an empty one
        if (r < r1) { o0 = __ldg(off + r) - off_bias; len = __ldg(off + r + 1) - off_bias - o0; }
        Pending pend; pend.n = 0; pend.d0 = make_uint2(0, 0); pend.d1 = make_uint2(0, 0);
        if (staged) {
            __syncthreads();                                 // previous tile fully consumed
            if (threadIdx.x == 0) {
                if (e > a) tile_load(smem, bases + a, e - a, &tile_bar);
                const u32 nt = tile + gridDim.x;             // this CTA's next tile: into L2 while this one is scanned
                if (nt < n_tiles) {
                    const u32 na = (__ldg(off + r_begin + nt * tile_reads) - off_bias) & ~15u;
                    const u32 ne = (__ldg(off + min(r_begin + (nt + 1) * tile_reads, r_end)) - off_bias + 15u) & ~15u;
                    if (ne > na && ne - na <= tile_bytes) tile_prefetch_l2(bases + na, ne - na);
                }
            }
            if (e > a) { mbar_wait(&tile_bar, parity); parity ^= 1; }
            const u32* sw = reinterpret_cast<const u32*>(smem);
            auto ld = [sw](u32 i) { return sw[i]; };
            created += scan_read(v, ld, ldr4, r < r1 ? o0 - a : 0u, len, a, pend);
            created += flush_pending(v, ld, pend, a);
        } else {
            const u32* gw = reinterpret_cast<const u32*>(bases);
            auto ld = [gw](u32 i) { return __ldg(gw + i); };
            created += scan_read(v, ld, ldr4, o0, len, 0, pend);
            created += flush_pending(v, ld, pend, 0);
        }
    }
    created = warp_sum_u32(created);
    if ((threadIdx.x & 31) == 0 && created) atomicAdd(gen_new, created);
}

// k_leftover: one LANE per queued stretch.  The lane rolls the k-mers of its stretch (one byte per
// k-mer, non-ACGT bytes reset the roll: KMC splits reads there) and counts each with count_one().
// Stretches longer than BK_LEFT_LONG k-mers (foreign reads, wrong diagonals) are handled by the whole
// warp afterwards, lane j taking k-mers j, j+32, ...
// LIST: the k-mers are not counted here but written to the list v.nov (one slot per k-mer of the stretch, BK_HOLE
// where the k-mer contained a non-ACGT byte); a warp reserves the slots of its 32 stretches with one atomic, and the
// kernel touches no table at all.  bk_bins.cuh counts the list.
#define BK_LEFT_LONG 96
#define BK_HOLE (~0ull)
template <int LIST>
__global__ void __launch_bounds__(256)
k_leftover(CountView v, const u8* __restrict__ bases, u32* gen_new) {
    const u32 n = min(*v.n_desc, v.desc_cap);
    const u32 lane = threadIdx.x & 31;
    const u32 stride = gridDim.x * blockDim.x;
    const u32* gw = reinterpret_cast<const u32*>(bases);
    auto ld = [gw](u32 i) { return __ldg(gw + i); };
    const u32 k = v.k;
    const u64 kmask = (k == 32) ? ~0ull : ((1ull << (2 * k)) - 1);
    u32 created = 0;
    const u32 n_round = (n + 31) & ~31u;
    for (u32 i = blockIdx.x * blockDim.x + threadIdx.x; i < n_round; i += stride) {
        uint2 d = make_uint2(0, 0);
        if (i < n) d = v.desc[i];
        const bool is_long = d.y > BK_LEFT_LONG;
        const bool mine = d.y != 0 && !is_long;
        u32 base = 0;
        if (LIST) {                                          // slots of this warp's stretches: one atomic
            const u32 want = mine ? d.y : 0u;
            u32 incl = want;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) { const u32 t = __shfl_up_sync(0xFFFFFFFFu, incl, o); if (lane >= (u32)o) incl += t; }
            const u32 total = __shfl_sync(0xFFFFFFFFu, incl, 31);
            if (lane == 31 && total) base = atomicAdd(v.nov_n, total);
            base = __shfl_sync(0xFFFFFFFFu, base, 31) + incl - want;
        }
        if (mine) {
            const u32 n_bytes = d.y + k - 1;
            u64 km = 0; u32 run = 0;
            u32 word = 0;
            for (u32 b = 0; b < n_bytes; b++) {
                const u32 addr = d.x + b;
                if (b == 0 || (addr & 3) == 0) word = ld(addr >> 2);      // (aligned 16-byte blocks with the next one requested ahead: measured slower, 0.158 -> 0.174 ms per sample)
                const u32 c = (word >> (8 * (addr & 3))) & 0xFFu;
                const u32 up = c & 0xDFu;
                const bool ok = up == 'A' || up == 'C' || up == 'G' || up == 'T';
                const u32 code = ((up >> 1) ^ (up >> 2)) & 3u;
                km = ((km << 2) | code) & kmask;
                run = ok ? run + 1 : 0;
                if (LIST) {
                    if (b + 1 >= k) {                        // k-mer number b + 1 - k of the stretch ends at this byte
                        const u32 pos = base + (b + 1 - k);                // reference k-mers among them are
                        if (pos < v.nov_cap) { v.nov[pos] = run >= k ? km : BK_HOLE; if (v.nov_w) v.nov_w[pos] = 1u; }   // recognised when the bins are counted
                        else *v.gen_full = 1;
                    }
                } else if (run >= k) created += count_one(v, km);
            }
        }
        u32 lm = __ballot_sync(0xFFFFFFFFu, is_long);
        while (lm) {
            const u32 src = __ffs(lm) - 1;
            lm &= lm - 1;
            const u32 off = __shfl_sync(0xFFFFFFFFu, d.x, src), cnt = __shfl_sync(0xFFFFFFFFu, d.y, src);
            created += count_stretch(v, ld, off, cnt, lane, 32);
        }
    }
    created = warp_sum_u32(created);
    if (lane == 0 && created) atomicAdd(gen_new, created);
}

// ------------------------------------------------------------------------------------------------
// prefix sum of the difference array (u32, wrap-around arithmetic) + fold onto distinct k-mer ids
// ------------------------------------------------------------------------------------------------
#define BK_PS_THREADS 256
#define BK_PS_PER_THREAD 16
#define BK_PS_BLOCK (BK_PS_THREADS * BK_PS_PER_THREAD)

__device__ __forceinline__ u32 block_excl_scan_256(u32 v, u32* total) {
    __shared__ u32 wsum[8];
    const u32 lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    u32 inc = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const u32 t = __shfl_up_sync(0xFFFFFFFFu, inc, o); if (lane >= (u32)o) inc += t; }
    if (lane == 31) wsum[w] = inc;
    __syncthreads();
    u32 woff = 0, tot = 0;
#pragma unroll
    for (u32 i = 0; i < 8; i++) { const u32 s = wsum[i]; if (i < w) woff += s; tot += s; }
    __syncthreads();
    if (total) *total = tot;
    return woff + inc - v;
}

__global__ void __launch_bounds__(BK_PS_THREADS) k_diff_blocksum(const u32* __restrict__ diff, u32 n, u32* bsum) {
    const u32 base = blockIdx.x * BK_PS_BLOCK + threadIdx.x * BK_PS_PER_THREAD;
    u32 s = 0;
#pragma unroll
    for (u32 i = 0; i < BK_PS_PER_THREAD; i++) if (base + i < n) s += diff[base + i];
    u32 tot;
    (void)block_excl_scan_256(s, &tot);
    if (threadIdx.x == 0) bsum[blockIdx.x] = tot;
}

__global__ void __launch_bounds__(BK_PS_THREADS) k_diff_scan_bsum(u32* bsum, u32 nb) {
    u32 carry = 0;
    for (u32 b0 = 0; b0 < nb; b0 += BK_PS_THREADS) {
        const u32 i = b0 + threadIdx.x;
        const u32 v = i < nb ? bsum[i] : 0;
        u32 tot;
        const u32 ex = block_excl_scan_256(v, &tot);
        if (i < nb) bsum[i] = carry + ex;
        carry += tot;
    }
}

__global__ void __launch_bounds__(BK_PS_THREADS)
k_diff_apply(const u32* __restrict__ diff, u32 n, const u32* __restrict__ bsum, const u32* __restrict__ slot2id, u32* idcnt) {
    const u32 base = blockIdx.x * BK_PS_BLOCK + threadIdx.x * BK_PS_PER_THREAD;
    u32 d[BK_PS_PER_THREAD];
    u32 s = 0;
#pragma unroll
    for (u32 i = 0; i < BK_PS_PER_THREAD; i++) { d[i] = (base + i < n) ? diff[base + i] : 0; s += d[i]; }
    u32 run = bsum[blockIdx.x] + block_excl_scan_256(s, nullptr);
#pragma unroll
    for (u32 i = 0; i < BK_PS_PER_THREAD; i++) {
        run += d[i];
        if (run != 0 && base + i < n) {
            const u32 id = slot2id[base + i];
            if (id != 0xFFFFFFFFu) atomicAdd(idcnt + id, run);
        }
    }
}

// ------------------------------------------------------------------------------------------------
// compaction = the KMC dump: keep ci <= count <= cx (cx = 1e9), stored count = min(count, cs)
// ------------------------------------------------------------------------------------------------
struct CompactArgs { u32 ci, cs; u32 rank, n_ranks; u64* out_kmers; u32* out_counts; u32 out_cap; FileCounters* fc; };

__device__ __forceinline__ void compact_emit(const CompactArgs& a, bool have, u64 kmer, u32 c, bool owned, u32& uniq, u64& total) {
    if (have && owned) { uniq++; total += c; }
    const bool keep = have && owned && c >= a.ci && c <= 1000000000u;
    const u32 slot = warp_append(&a.fc->n_counted, keep);
    if (keep && slot < a.out_cap) { a.out_kmers[slot] = kmer; a.out_counts[slot] = min(c, a.cs); }
}

__global__ void __launch_bounds__(256) k_compact_ids(CompactArgs a, const u32* __restrict__ idcnt, const u64* __restrict__ id_kmer, u32 n_ids) {
    const u32 stride = gridDim.x * blockDim.x;
    u32 uniq = 0; u64 total = 0;
    const u32 n_round = (n_ids + 31) & ~31u;
    for (u32 i = blockIdx.x * blockDim.x + threadIdx.x; i < n_round; i += stride) {
        const bool in = i < n_ids;
        const u32 c = in ? idcnt[i] : 0;
        compact_emit(a, in && c != 0, in ? id_kmer[i] : 0, c, (i % a.n_ranks) == a.rank, uniq, total);
    }
    uniq = warp_sum_u32(uniq); total = warp_sum_u64(total);
    if ((threadIdx.x & 31) == 0 && uniq) { atomicAdd(&a.fc->unique, uniq); atomicAdd((unsigned long long*)&a.fc->total_kmers, (unsigned long long)total); }
}

// ------------------------------------------------------------------------------------------------
// k_map — reference map_kmers (src/call.rs:1257-1434).  One warp per counted k-mer, one lane per
// bucket: lane i computes bucket id i of the canonical k-mer in closed form (src/lcb.rs:1-45,
// SURVEY.md Appendix G; u64 arithmetic wraps like release Rust), probes the bucket table and walks
// the entries.  STATS pass: per-genome hit counts → perfect / variant / unique-perfect tallies.
// PILEUP pass (after selection): only the selected genome's entries update the four arrays —
// support += 1, depth = max(depth, count); identical to the reference because only that genome's
// arrays are ever read (src/call.rs:252-255, 344-347).
// ------------------------------------------------------------------------------------------------
struct MapView {
    u32 k; u32 b0, b1;                     // queried bucket indices [b0, b1) (src/call.rs:1291-1300)
    const BucketSlotD* slots; u32 shift, mask;
    const BucketEntryD* entries;
    u32 n_genomes; const u32* genome_row0;
    // grouped form of the re-keyed table (bk_host.h: group_slots / group_centers / group_buckets); null when not available
    const BucketSlotD* gslots; u32 gshift, gmask; const BucketSlotD* gcenters; const uint2* gbuckets; u32 gmid;
};

__device__ __forceinline__ u64 revcomp_dev(u64 v, u32 k) {
    u64 x = ~v;
    x = ((x >> 2) & 0x3333333333333333ull) | ((x & 0x3333333333333333ull) << 2);
    x = ((x >> 4) & 0x0F0F0F0F0F0F0F0Full) | ((x & 0x0F0F0F0F0F0F0F0Full) << 4);
    const u32 lo = (u32)x, hi = (u32)(x >> 32);
    x = ((u64)__byte_perm(lo, 0, 0x0123) << 32) | __byte_perm(hi, 0, 0x0123);
    return x >> (64 - 2 * k);
}

// REKEY: the table is keyed by (bucket index << 58) | (canonical k-mer with that digit zeroed) — the same map as the
// bucket ids (bijection), without the id arithmetic; the host verifies every key before enabling it (derive_index).
template <int PILEUP, int REKEY>
__global__ void __launch_bounds__(256)
k_map(MapView m, const u64* __restrict__ kmers, const u32* __restrict__ counts, const u32* n_ptr, u32 n_cap,
      u32* gstats, const i32* best_ptr, u32* pile, u32 pile_stride) {
    extern __shared__ u32 sm[];            // STATS: [n_genomes*4] CTA tallies, then 8 x [n_genomes] per-warp hits
    const u32 lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const u32 warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, n_warps = (gridDim.x * blockDim.x) >> 5;
    const u32 n = min(*n_ptr, n_cap);
    const u32 k = m.k;
    u32* cta = sm;
    u32* hits = sm + m.n_genomes * 4 + wib * m.n_genomes;
    __shared__ u32 s_wpre[8][33], s_woff[8][32];       // per warp: exclusive prefix of the entry-list lengths (+ sentinel), list starts
    u32* wpre = s_wpre[wib]; u32* woff = s_woff[wib];
    if (lane == 0) wpre[32] = 0xFFFFFFFFu;
    i32 best = -1; u32 g_row0 = 0;
    if (PILEUP) {
        best = *best_ptr;
        if (best < 0) return;
        g_row0 = m.genome_row0[best];
    } else {
        for (u32 i = threadIdx.x; i < m.n_genomes * 12; i += blockDim.x) sm[i] = 0;
        __syncthreads();
    }
    const u32 nb = m.b1 - m.b0;
    for (u32 t = warp; t < n; t += n_warps) {
        const u64 fwd = kmers[t];
        const u32 cnt = counts[t];
        const u64 rev = revcomp_dev(fwd, k);
        const bool rc = !(fwd < rev);                                 // src/lcb.rs:87-95
        const u64 kb = rc ? rev : fwd;
        // bucket id of index `lane`
        const bool valid = lane < k;
        const u32 sh = valid ? 2 * (k - 1 - lane) : 0;
        const u64 w = 1ull << sh;
        const u64 dgt = (kb >> sh) & 3;
        const u64 cur = dgt << sh;
        const u64 val = kb & (w - 1);
        const u64 mu = valid ? (dgt ? w + (cur >> 2) * (u64)(k - 1 - lane) : val) : 0;
        const u32 zmask = __ballot_sync(0xFFFFFFFFu, valid && dgt == 0);
        const u64 num_a = __popc(zmask & ((1u << lane) - 1));
        u64 bucket;
        if (REKEY) bucket = ((u64)lane << 58) | (kb & ~(3ull << sh));
        else { const u64 sum_mu = warp_sum_u64(mu); bucket = sum_mu - mu + val - num_a * cur + 1 + num_a; }
        u32 off = 0, len = 0;
        if (lane >= m.b0 && lane < m.b1) {
            u32 h = hash_slot(bucket, m.shift);
            for (;;) {
                const uint4 s = __ldg(reinterpret_cast<const uint4*>(m.slots) + h);
                const u64 key = ((u64)s.y << 32) | s.x;
                if (key == bucket) { off = s.z; len = s.w; break; }
                if (key == BK_EMPTY) break;
                h = (h + 1) & m.mask;
            }
        }
        // entry lists: a few entries per bucket (databases of a few genomes) are walked by the lane that probed the
        // bucket, all lanes in parallel; long lists (a 200-strain database holds ~150 entries per bucket, 2,400 per
        // exact k-mer) are flattened over the warp: item t of the concatenated lists goes to lane t mod 32, so
        // consecutive lanes read consecutive entries
        u32 incl = len;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const u32 v = __shfl_up_sync(0xFFFFFFFFu, incl, o); if (lane >= (u32)o) incl += v; }
        const u32 total = __shfl_sync(0xFFFFFFFFu, incl, 31);
        const bool flat = total > 64;
        if (flat) { wpre[lane] = incl - len; woff[lane] = off; __syncwarp(); }
        const u32 n_it = flat ? (total + 31) / 32 : len;
        for (u32 it = 0; it < n_it; it++) {
            u32 e_at;
            if (flat) {
                const u32 t = it * 32 + lane;
                if (t >= total) break;
                u32 lo = 0;                                            // last lane whose exclusive prefix is <= t
#pragma unroll
                for (u32 step = 16; step; step >>= 1) if (wpre[lo + step] <= t) lo += step;
                e_at = woff[lo] + (t - wpre[lo]);
            } else e_at = off + it;
            {
                const uint2 raw = __ldg(reinterpret_cast<const uint2*>(m.entries) + e_at);
                const u32 row = raw.x, file_id = raw.y & 0xFFFFu, idx = (raw.y >> 16) & 0xFFu, canon = raw.y >> 24;
                if (row == 0xFFFFFFFFu) continue;
                if (!PILEUP) {
                    atomicAdd(hits + file_id, 1u);                    // src/call.rs:1316-1318
                } else if ((i32)file_id == best) {
                    u32 bit; bool to_fwd;
                    if (canon) { bit = (u32)((kb >> (2 * idx)) & 3) ^ 3u; to_fwd = rc; }          // src/call.rs:1330-1357
                    else { bit = (u32)((kb >> (2 * (k - idx - 1))) & 3); to_fwd = !rc; }       // src/call.rs:1358-1384
                    const u32 cell = (row + idx - g_row0) * 4 + bit;
                    atomicAdd(pile + (to_fwd ? 2u : 3u) * pile_stride + cell, 1u);
                    atomicMax(pile + (to_fwd ? 0u : 1u) * pile_stride + cell, cnt);
                }
            }
        }
        __syncwarp();
        if (!PILEUP) {
            __syncwarp();
            u32 n_perfect = 0, perfect_g = 0;
            for (u32 g0 = 0; g0 < m.n_genomes; g0 += 32) {            // src/call.rs:1389-1419
                const u32 g = g0 + lane;
                const u32 h = g < m.n_genomes ? hits[g] : 0;
                bool perfect = false;
                if (h) {
                    hits[g] = 0;
                    cta[g * 4 + 3] = 1;
                    perfect = (h == nb);
                    atomicAdd(cta + g * 4 + (perfect ? 0 : 1), 1u);
                }
                const u32 pm = __ballot_sync(0xFFFFFFFFu, perfect);
                if (pm) { n_perfect += __popc(pm); perfect_g = g0 + __ffs(pm) - 1; }
            }
            if (n_perfect == 1 && lane == 0) atomicAdd(cta + perfect_g * 4 + 2, 1u);
            __syncwarp();
        }
    }
    if (!PILEUP) {
        __syncthreads();
        for (u32 i = threadIdx.x; i < m.n_genomes * 4; i += blockDim.x) {
            const u32 v = cta[i];
            if (v) { if ((i & 3) == 3) gstats[i] = 1; else atomicAdd(gstats + i, v); }
        }
    }
}

// k_map_small — the same computation with ONE THREAD per counted k-mer, for databases of at most four
// genomes (per-genome hit counts fit four 16-bit fields of a register).  The bucket ids come from the
// incremental form of src/lcb.rs:1-45 (two passes over the k digits, u64 wrap-around), each queried
// bucket is probed as soon as its id is known.  ~25x fewer warp instructions per k-mer than the
// warp-per-k-mer kernel; the counted list keeps reference k-mers and novel k-mers in separate runs,
// so warps stay homogeneous.
// MODE 0: tallies only; MODE 1: pileup of the selected genome only; MODE 2 (databases of at most four genomes, one
// pass): tallies AND the pileup of EVERY genome — genome g's four arrays start at pile + g * 4 * pile_stride, rows
// relative to the genome; k_pile_pick copies the selected genome's slice afterwards.  Same results as the reference,
// which also updates every genome (src/call.rs:1324-1385) and reads only the selected one.
template <int MODE, int REKEY>
__global__ void __launch_bounds__(256)
k_map_small(MapView m, const u64* __restrict__ kmers, const u32* __restrict__ counts, const u32* n_ptr, u32 n_cap,
            u32* gstats, const i32* best_ptr, u32* pile, u32 pile_stride) {
    const u32 n = min(*n_ptr, n_cap);
    const u32 k = m.k;
    const u32 lane = threadIdx.x & 31;
    i32 best = -1; u32 g_row0 = 0;
    if (MODE == 1) {
        best = *best_ptr;
        if (best < 0) return;
        g_row0 = m.genome_row0[best];
    }
    const u32 nb = m.b1 - m.b0;
    u32 acc[12];                      // lane 0: [g*3 + {perfect, variant, unique}] ; presence = any of them
#pragma unroll
    for (u32 i = 0; i < 12; i++) acc[i] = 0;
    const u32 stride = gridDim.x * blockDim.x;
    const u32 n_round = (n + 31) & ~31u;
    for (u32 t = blockIdx.x * blockDim.x + threadIdx.x; t < n_round; t += stride) {
        const bool live = t < n;
        u64 hits4 = 0;                // 4 x 16-bit per-genome hit counts (saturating is unnecessary: <= 31*entries)
        if (live) {
            const u64 fwd = kmers[t];
            const u32 cnt = counts[t];
            const u64 rev = revcomp_dev(fwd, k);
            const bool rc = !(fwd < rev);
            const u64 kb = rc ? rev : fwd;
            // pass 1: sum of mu
            u64 mask = 3ull << (2 * (k - 1)), p = 1ull << (2 * (k - 1));
            u64 val = kb, sum_mu = 0;
            if (!REKEY) {
                for (u32 i = 0; i < k; i++) {
                    const u64 cur = kb & mask;
                    val -= cur;
                    sum_mu += cur ? p + (cur >> 2) * (u64)(k - 1 - i) : val;
                    mask >>= 2; p >>= 2;
                }
            }
            // pass 2: ids of the queried buckets, probe, walk entries
            mask = 3ull << (2 * (k - 1)); p = 1ull << (2 * (k - 1));
            val = kb;
            u64 num_a = 0;
            for (u32 i = REKEY ? m.b0 : 0u; i < m.b1; i++) {
                u64 cur = 0;
                if (!REKEY) { cur = kb & mask; val -= cur; }
                if (i >= m.b0) {
                    u64 bucket;
                    if (REKEY) bucket = ((u64)i << 58) | (kb & ~(3ull << (2 * (k - 1 - i))));
                    else { const u64 mu = cur ? p + (cur >> 2) * (u64)(k - 1 - i) : val; bucket = sum_mu - mu + val - num_a * cur + 1 + num_a; }
                    u32 h = hash_slot(bucket, m.shift);
                    u32 off = 0, len = 0;
                    for (;;) {
                        const uint4 sl = __ldg(reinterpret_cast<const uint4*>(m.slots) + h);
                        const u64 key = ((u64)sl.y << 32) | sl.x;
                        if (key == bucket) { off = sl.z; len = sl.w; break; }
                        if (key == BK_EMPTY) break;
                        h = (h + 1) & m.mask;
                    }
                    for (u32 j = 0; j < len; j++) {
                        const uint2 raw = __ldg(reinterpret_cast<const uint2*>(m.entries) + off + j);
                        const u32 row = raw.x, file_id = raw.y & 0xFFFFu, idx = (raw.y >> 16) & 0xFFu, canon = raw.y >> 24;
                        if (row == 0xFFFFFFFFu) continue;
                        if (MODE != 1) hits4 += 1ull << (16 * file_id);       // src/call.rs:1316-1318
                        if (MODE == 2 || (MODE == 1 && (i32)file_id == best)) {
                            u32 bit; bool to_fwd;
                            if (canon) { bit = (u32)((kb >> (2 * idx)) & 3) ^ 3u; to_fwd = rc; }        // src/call.rs:1330-1357
                            else { bit = (u32)((kb >> (2 * (k - idx - 1))) & 3); to_fwd = !rc; }     // src/call.rs:1358-1384
                            u32* gp = pile;
                            u32 r0 = g_row0;
                            if (MODE == 2) { gp = pile + (size_t)file_id * 4 * pile_stride; r0 = __ldg(m.genome_row0 + file_id); }
                            const u32 cell = (row + idx - r0) * 4 + bit;
                            atomicAdd(gp + (to_fwd ? 2u : 3u) * pile_stride + cell, 1u);
                            atomicMax(gp + (to_fwd ? 0u : 1u) * pile_stride + cell, cnt);
                        }
                    }
                }
                num_a += (cur == 0) ? 1 : 0;
                mask >>= 2; p >>= 2;
            }
        }
        if (MODE != 1) {                                                     // src/call.rs:1389-1419
            u32 n_perfect = 0;
#pragma unroll
            for (u32 g = 0; g < 4; g++) n_perfect += (((hits4 >> (16 * g)) & 0xFFFFu) == nb && nb != 0) ? 1u : 0u;
#pragma unroll
            for (u32 g = 0; g < 4; g++) {
                const u32 h = (u32)((hits4 >> (16 * g)) & 0xFFFFu);
                const bool perfect = h != 0 && h == nb;
                const u32 mp = __ballot_sync(0xFFFFFFFFu, perfect);
                const u32 mv = __ballot_sync(0xFFFFFFFFu, h != 0 && !perfect);
                const u32 mu_ = __ballot_sync(0xFFFFFFFFu, perfect && n_perfect == 1);
                if (lane == 0) { acc[g * 3] += __popc(mp); acc[g * 3 + 1] += __popc(mv); acc[g * 3 + 2] += __popc(mu_); }
            }
        }
    }
    if (MODE != 1 && lane == 0) {
        for (u32 g = 0; g < m.n_genomes && g < 4; g++) {
            if (acc[g * 3]) atomicAdd(gstats + g * 4, acc[g * 3]);
            if (acc[g * 3 + 1]) atomicAdd(gstats + g * 4 + 1, acc[g * 3 + 1]);
            if (acc[g * 3 + 2]) atomicAdd(gstats + g * 4 + 2, acc[g * 3 + 2]);
            if (acc[g * 3] | acc[g * 3 + 1]) gstats[g * 4 + 3] = 1;
        }
    }
}

// k_map_grp — the thread-per-k-mer map on the GROUPED table (bk_host.h: group_slots / group_centers / group_buckets).
// The reference k-mers that can share a bucket with the query share its low half (bucket index < k/2) or its high half,
// so two group probes (issued together) find the one or two "centers" per side; a center equal to the query hits every
// bucket of the side, a center that differs in exactly digit j hits bucket j, and each hit costs one {off, len} load.
// This replaces the 16 independent random probes of k_map_small (bound by the outstanding L2 misses an SM can hold)
// and the scan over ~10 bucket records per side of the first grouped layout.  Modes as k_map_small.
// entries [off, off+len) of one bucket hit: tallies and / or pileup updates (src/call.rs:1309-1385)
#define BK_MAP_QUEUE 4
#ifndef BK_MAP_BATCH
#define BK_MAP_BATCH 4          // independent record / entry loads issued together
#endif
#ifndef BK_MAP_GRP_CTAS
#define BK_MAP_GRP_CTAS 3       // resident CTAs per SM the register budget is set for
#endif
template <int MODE>
__device__ __forceinline__ void map_walk(const MapView& m, u32 off, u32 len, u64 kb, bool rc, u32 cnt, i32 best, u32 g_row0,
                                         u32* pile, u32 pile_stride, u64& hits4) {
    const u32 k = m.k;
    // the entries of a hit are independent loads: BK_MAP_BATCH of them are in flight before the first is used (walked
    // one by one, waiting for each entry was a fifth of the kernel's stall samples)
    for (u32 j0 = 0; j0 < len; j0 += BK_MAP_BATCH) {
      uint2 rawb[BK_MAP_BATCH];
#pragma unroll
      for (u32 jj = 0; jj < BK_MAP_BATCH; jj++)
          rawb[jj] = j0 + jj < len ? __ldg(reinterpret_cast<const uint2*>(m.entries) + off + j0 + jj) : make_uint2(0xFFFFFFFFu, 0u);
#pragma unroll
      for (u32 jj = 0; jj < BK_MAP_BATCH; jj++) {
        if (jj && j0 + jj >= len) break;
        const uint2 raw = rawb[jj];
        const u32 row = raw.x, file_id = raw.y & 0xFFFFu, idx = (raw.y >> 16) & 0xFFu, canon = raw.y >> 24;
        if (row == 0xFFFFFFFFu) continue;
        if (MODE != 1) hits4 += 1ull << (16 * file_id);       // src/call.rs:1316-1318
        if (MODE == 2 || (MODE == 1 && (i32)file_id == best)) {
            u32 bit; bool to_fwd;
            if (canon) { bit = (u32)((kb >> (2 * idx)) & 3) ^ 3u; to_fwd = rc; }        // src/call.rs:1330-1357
            else { bit = (u32)((kb >> (2 * (k - idx - 1))) & 3); to_fwd = !rc; }     // src/call.rs:1358-1384
            u32* gp = pile;
            u32 r0 = g_row0;
            if (MODE == 2) { gp = pile + (size_t)file_id * 4 * pile_stride; r0 = __ldg(m.genome_row0 + file_id); }
            const u32 cell = (row + idx - r0) * 4 + bit;
            atomicAdd(gp + (to_fwd ? 2u : 3u) * pile_stride + cell, 1u);
            atomicMax(gp + (to_fwd ? 0u : 1u) * pile_stride + cell, cnt);
        }
      }
    }
}

template <int MODE>
__global__ void __launch_bounds__(256, BK_MAP_GRP_CTAS)
k_map_grp(MapView m, const u64* __restrict__ kmers, const u32* __restrict__ counts, const u32* n_ptr, u32 n_cap,
          u32* gstats, const i32* best_ptr, u32* pile, u32 pile_stride) {
    const u32 n = min(*n_ptr, n_cap);
    const u32 k = m.k;
    const u32 lane = threadIdx.x & 31;
    i32 best = -1; u32 g_row0 = 0;
    if (MODE == 1) {
        best = *best_ptr;
        if (best < 0) return;
        g_row0 = m.genome_row0[best];
    }
    const u32 nb = m.b1 - m.b0;
    const u32 lo_bits = 2 * (k - m.gmid);
    u32 acc[12];
#pragma unroll
    for (u32 i = 0; i < 12; i++) acc[i] = 0;
    const u32 stride = gridDim.x * blockDim.x;
    const u32 n_round = (n + 31) & ~31u;
    for (u32 t = blockIdx.x * blockDim.x + threadIdx.x; t < n_round; t += stride) {
        const bool live = t < n && nb != 0;
        u64 hits4 = 0;
        u32 qo0 = 0, qo1 = 0, qo2 = 0, qo3 = 0, ql0 = 0, ql1 = 0, ql2 = 0, ql3 = 0, nq = 0;     // hits waiting
        u64 kb_keep = 0; bool rc_keep = false; u32 cnt_keep = 0;
        if (live) {
            const u64 fwd = __ldg(kmers + t);
            const u32 cnt = __ldg(counts + t);
            const u64 rev = revcomp_dev(fwd, k);
            const bool rc = !(fwd < rev);
            const u64 kb = rc ? rev : fwd;
            // the two groups: probe both tables before looking at either
            u64 gk[2]; u32 gh[2]; uint4 gs[2];
            gk[0] = kb & ((1ull << lo_bits) - 1);
            gk[1] = (1ull << 62) | (kb >> lo_bits);
#pragma unroll
            for (u32 sd = 0; sd < 2; sd++) { gh[sd] = hash_slot(gk[sd], m.gshift); gs[sd] = __ldg(reinterpret_cast<const uint4*>(m.gslots) + gh[sd]); }
#pragma unroll
            for (u32 sd = 0; sd < 2; sd++) {
                // side 0 serves bucket indices [b0, min(b1, mid)), side 1 [max(b0, mid), b1)
                const u32 i_lo = sd == 0 ? m.b0 : max(m.b0, m.gmid), i_hi = sd == 0 ? min(m.b1, m.gmid) : m.b1;
                if (i_lo >= i_hi) continue;
                u32 h = gh[sd];
                uint4 sl = gs[sd];
                u32 first = 0, count = 0;
                for (;;) {
                    const u64 key = ((u64)sl.y << 32) | sl.x;
                    if (key == gk[sd]) { first = sl.z; count = sl.w; break; }
                    if (key == BK_EMPTY) break;
                    h = (h + 1) & m.gmask;
                    sl = __ldg(reinterpret_cast<const uint4*>(m.gslots) + h);
                }
                const u32 side_lo = sd == 0 ? 0u : m.gmid;
                const u32 range = ((1u << i_hi) - 1u) & ~((1u << i_lo) - 1u);              // queried indices of this side (k <= 29)
                u32 done = 0;                                                             // indices already taken
                for (u32 cb = 0; cb < count; cb += 2) {                                    // centers, two loads in flight
                    uint4 cen[2];
                    cen[0] = __ldg(reinterpret_cast<const uint4*>(m.gcenters) + first + cb);
                    cen[1] = cb + 1 < count ? __ldg(reinterpret_cast<const uint4*>(m.gcenters) + first + cb + 1) : make_uint4(0u, 0u, 0u, 0u);
#pragma unroll
                    for (u32 cc = 0; cc < 2; cc++) {
                        const u64 x = kb ^ (((u64)cen[cc].y << 32) | cen[cc].x);
                        u32 cand = cen[cc].w & range;                                      // center == query: every bucket present
                        if (x) {
                            const u64 nz = (x | (x >> 1)) & 0x5555555555555555ull;       // one bit per differing digit
                            cand = (nz & (nz - 1)) ? 0u : cand & (1u << (k - 1 - ((63u - (u32)__clzll((long long)nz)) >> 1)));
                        }
                        cand &= ~done;
                        done |= cand;
                        const u32 b_first = cen[cc].z;                                     // bucket of index side_lo
                        while (cand) {                                                     // hits: {off, len} of up to BK_MAP_BATCH buckets together
                            u32 bi[BK_MAP_BATCH]; uint2 ol[BK_MAP_BATCH];
#pragma unroll
                            for (u32 jj = 0; jj < BK_MAP_BATCH; jj++) {
                                bi[jj] = cand ? (u32)__ffs((int)cand) - 1u : 0xFFFFFFFFu;
                                cand &= cand - 1;
                                ol[jj] = bi[jj] != 0xFFFFFFFFu ? __ldg(m.gbuckets + (b_first + bi[jj] - side_lo)) : make_uint2(0u, 0u);
                            }
#pragma unroll
                            for (u32 jj = 0; jj < BK_MAP_BATCH; jj++) {
                                if (bi[jj] == 0xFFFFFFFFu) break;
                                // a hit: remember it; the entries are walked after the lookups, all lanes of the warp together
                                // (walked lane by lane as the hits turned up, that part ran with 1.7 of 32 lanes active).
                                // Shift-in queue: no dynamic indexing; a fifth hit evicts the oldest, which is walked at once.
                                if (nq == BK_MAP_QUEUE) { map_walk<MODE>(m, qo3, ql3, kb, rc, cnt, best, g_row0, pile, pile_stride, hits4); nq--; }
                                qo3 = qo2; ql3 = ql2; qo2 = qo1; ql2 = ql1; qo1 = qo0; ql1 = ql0; qo0 = ol[jj].x; ql0 = ol[jj].y; nq++;
                            }
                        }
                    }
                }
            }
            kb_keep = kb; rc_keep = rc; cnt_keep = cnt;
        }
        // the queued hits, slot by slot, in lockstep
        const u32 qo[BK_MAP_QUEUE] = {qo0, qo1, qo2, qo3}, ql[BK_MAP_QUEUE] = {ql0, ql1, ql2, ql3};
#pragma unroll
        for (u32 q = 0; q < BK_MAP_QUEUE; q++) {
            if (!__any_sync(0xFFFFFFFFu, q < nq)) break;
            if (q < nq) map_walk<MODE>(m, qo[q], ql[q], kb_keep, rc_keep, cnt_keep, best, g_row0, pile, pile_stride, hits4);
        }
        if (MODE != 1) {                                                     // src/call.rs:1389-1419
            u32 n_perfect = 0;
#pragma unroll
            for (u32 g = 0; g < 4; g++) n_perfect += (((hits4 >> (16 * g)) & 0xFFFFu) == nb && nb != 0) ? 1u : 0u;
#pragma unroll
            for (u32 g = 0; g < 4; g++) {
                const u32 h = (u32)((hits4 >> (16 * g)) & 0xFFFFu);
                const bool perfect = h != 0 && h == nb;
                const u32 mp = __ballot_sync(0xFFFFFFFFu, perfect);
                const u32 mv = __ballot_sync(0xFFFFFFFFu, h != 0 && !perfect);
                const u32 mu_ = __ballot_sync(0xFFFFFFFFu, perfect && n_perfect == 1);
                if (lane == 0) { acc[g * 3] += __popc(mp); acc[g * 3 + 1] += __popc(mv); acc[g * 3 + 2] += __popc(mu_); }
            }
        }
    }
    if (MODE != 1 && lane == 0) {
        for (u32 g = 0; g < m.n_genomes && g < 4; g++) {
            if (acc[g * 3]) atomicAdd(gstats + g * 4, acc[g * 3]);
            if (acc[g * 3 + 1]) atomicAdd(gstats + g * 4 + 1, acc[g * 3 + 1]);
            if (acc[g * 3 + 2]) atomicAdd(gstats + g * 4 + 2, acc[g * 3 + 2]);
            if (acc[g * 3] | acc[g * 3 + 1]) gstats[g * 4 + 3] = 1;
        }
    }
}

// after a MODE 2 map: the selected genome's four arrays → the front of the buffer every later stage reads
__global__ void __launch_bounds__(256) k_pile_pick(const u32* __restrict__ pile_all, u32* __restrict__ pile, u32 pile_stride, const i32* best_ptr) {
    const i32 best = *best_ptr;
    if (best < 0) return;
    const uint4* src = reinterpret_cast<const uint4*>(pile_all + (size_t)best * 4 * pile_stride);
    uint4* dst = reinterpret_cast<uint4*>(pile);
    const u32 n4 = pile_stride;                       // 4 arrays * pile_stride u32 = pile_stride uint4
    for (u32 i = blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += gridDim.x * blockDim.x) dst[i] = src[i];
}

// k_select — pick_best_genome / pick_best_genome_paired (src/call.rs:422-502): argmax of
// perfect / genome_len / 2.0 with strict '>' from 0.0; ties keep the lowest file index (the
// reference's tie order is FxHashMap iteration order, unpinned).
__global__ void k_select(const u32* gstats0, const u32* gstats1, u32 n_files, u32 n_genomes, const u64* genome_len, Counters* c) {
    if (threadIdx.x || blockIdx.x) return;
    i32 best = -1; double best_score = 0.0;
    for (u32 g = 0; g < n_genomes; g++) {
        u64 perfect = gstats0[g * 4]; bool present = gstats0[g * 4 + 3] != 0;
        if (n_files > 1) { perfect += gstats1[g * 4]; present = present || gstats1[g * 4 + 3] != 0; }
        if (!present) continue;
        const double score = (double)perfect / (double)genome_len[g] / 2.0;
        if (score > best_score) { best_score = score; best = (i32)g; }
    }
    c->best = best;
}

// ------------------------------------------------------------------------------------------------
// Noise baseline, src/call.rs:799-967: bk_noise.cuh (k_noise_fracs / k_noise_seq / k_noise_fix / k_noise_tau).
// ------------------------------------------------------------------------------------------------
struct ScoreView {
    u32 n_genomes; const u32* genome_row0; const u32* genome_seq_off; const u32* seq_row0;
    const u8* ref_code; const Counters* ctr;
    const u32* pile; u32 pile_stride;
};

// ------------------------------------------------------------------------------------------------
// k_call — call_variants, src/call.rs:969-1150 (one thread per position of the selected genome).
// ------------------------------------------------------------------------------------------------
struct CallParams {
    u32 k; u32 no_end_filter, no_strand_filter, no_strand_balance_filter; u32 n_per_strand;
    u64 min_depth, min_variant_depth;
    double min_af, strand_balance_ratio, strand_odds_max, variant_multiplier;
};

__global__ void __launch_bounds__(256)
k_call(ScoreView sv, CallParams p, const double* __restrict__ noise_max, bk_variant* vars, u32 var_cap, Counters* ctr) {
    const i32 best = ctr->best;
    if (best < 0) return;
    const u32 g_row0 = sv.genome_row0[best];
    const u32 rows = sv.genome_row0[best + 1] - g_row0;
    const u32 row = blockIdx.x * blockDim.x + threadIdx.x;
    u32 covered = 0; u64 coverage = 0;
    if (row < rows) {
        u32 s = sv.genome_seq_off[best];
        while (sv.seq_row0[s + 1] - g_row0 <= row) s++;
        const u32 seq_r0 = sv.seq_row0[s] - g_row0;
        const u32 len = sv.seq_row0[s + 1] - sv.seq_row0[s];
        const u32 i = row - seq_r0;
        u32 start = 0, end = len;
        if (!p.no_end_filter) { start = p.k; end = len >= p.k ? len - p.k : 0; }     // src/call.rs:1013-1016
        if (i >= start && i < end) {
            const uint4 f4 = *reinterpret_cast<const uint4*>(sv.pile + row * 4);
            const uint4 r4 = *reinterpret_cast<const uint4*>(sv.pile + sv.pile_stride + row * 4);
            const uint4 cf4 = *reinterpret_cast<const uint4*>(sv.pile + 2 * sv.pile_stride + row * 4);
            const uint4 cr4 = *reinterpret_cast<const uint4*>(sv.pile + 3 * sv.pile_stride + row * 4);
            const u32 fw[4] = {f4.x, f4.y, f4.z, f4.w}, rv[4] = {r4.x, r4.y, r4.z, r4.w};
            const u32 cf[4] = {cf4.x, cf4.y, cf4.z, cf4.w}, cr[4] = {cr4.x, cr4.y, cr4.z, cr4.w};
            const u32 ref_base = sv.ref_code[g_row0 + row];
            u64 row_total[4]; u64 total_depth = 0;
#pragma unroll
            for (int b = 0; b < 4; b++) { row_total[b] = (u64)fw[b] + rv[b]; total_depth += row_total[b]; }
            if (total_depth != 0) {
                covered = 1; coverage = total_depth;
                for (u32 alt = 0; alt < 4; alt++) {
                    if (alt == ref_base || row_total[alt] == 0) continue;
                    double sor = p.strand_odds_max + 1.0;
                    if (!p.no_strand_filter) {
                        const double a = (double)fw[ref_base] + 1.0, b = (double)rv[ref_base] + 1.0;
                        const double c = (double)fw[alt] + 1.0, d = (double)rv[alt] + 1.0;
                        const double ref_total = a + b + c + d;
                        const double min_strand_percent = fmin(a + c, b + d) / ref_total;
                        if (!p.no_strand_balance_filter || min_strand_percent >= p.strand_balance_ratio) {
                            const double r = (a * d) / (b * c);
                            const double ref_ratio = fmin(a, b) / fmax(a, b);
                            const double alt_ratio = fmin(c, d) / fmax(c, d);
                            sor = log(r + (1.0 / r)) + log(ref_ratio) - log(alt_ratio);
                            if (sor > p.strand_odds_max) continue;
                            if (cf[alt] < p.n_per_strand && cr[alt] < p.n_per_strand) continue;
                        } else sor = -1.0;
                    }
                    const u64 alt_count = row_total[alt];
                    const double af = (double)alt_count / (double)total_depth;
                    const double y0 = p.variant_multiplier;
                    const double factor = y0 + 0.5 * pow(0.03, 100.0 * af);
                    if (af < p.min_af || af < (fmax(factor, y0) * noise_max[row])) continue;
                    if (af >= 0.5) atomicAdd(&ctr->n_major, 1u);
                    else {
                        if (total_depth < p.min_depth) continue;
                        if (alt_count < p.min_variant_depth) continue;
                        atomicAdd(&ctr->n_minor, 1u);
                    }
                    const u32 slot = atomicAdd(&ctr->n_var, 1u);
                    if (slot < var_cap) {
                        bk_variant v;
                        v.seq = s - sv.genome_seq_off[best]; v.pos = i + 1; v.ref_base = (u8)ref_base; v.alt_base = (u8)alt;
                        for (int q = 0; q < 6; q++) v._pad[q] = 0;
                        v.fwd_ref = fw[ref_base]; v.rev_ref = rv[ref_base]; v.fwd_alt = fw[alt]; v.rev_alt = rv[alt];
                        v.depth = total_depth; v.af = af; v.sor = sor;
                        vars[slot] = v;
                    } else ctr->var_overflow = 1;
                }
            }
        }
    }
    covered = warp_sum_u32(covered); coverage = warp_sum_u64(coverage);
    if ((threadIdx.x & 31) == 0 && covered) {
        atomicAdd((unsigned long long*)&ctr->pos_covered, (unsigned long long)covered);
        atomicAdd((unsigned long long*)&ctr->total_cov, (unsigned long long)coverage);
    }
}

}  // namespace bk

#include "bk_bins.cuh"
#include "bk_dense.cuh"
