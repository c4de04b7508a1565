// bk_bins.cuh — counting the novel k-mers of one file without a DRAM-resident hash table.
//
// k_leftover<1> leaves every k-mer occurrence that did not extend a run in a flat list (8 bytes each, BK_HOLE for
// k-mers with a non-ACGT byte).  A hash table for them would be larger than L2 (2.9 M distinct k-mers per file at C2) and every one of
// the 6 M insertions a dependent random DRAM access.  Instead the list is radix-partitioned by the high bits of the
// k-mer's hash into P bins, and each bin — a few hundred to a few thousand occurrences — is counted by one CTA in a
// SHARED-MEMORY table and written straight to the counted list (the "KMC dump": ci <= count <= 1e9, stored
// min(count, cs); reference src/call.rs:1166-1177).  All global traffic is streaming:
//
//   k_bin_hist     G CTAs, one contiguous chunk of the list each: per-CTA bin histogram → cnt[bin * G + cta]
//   (prefix sum)   exclusive scan of cnt in that order = where every CTA's share of every bin starts
//   k_bin_scatter  same chunks: k-mers to their bin (shared-memory cursors, no global atomics)
//   k_bin_count    one CTA per bin: shared-memory open addressing.  A bin is first tried in ONE round whatever its
//                  size (a deep sample repeats the same k-mers: 40,000 occurrences of a bin at 10^6x are ~900 distinct
//                  k-mers); only if the table fills up is it walked in rounds by a second hash, sized for the worst
//                  case (every occurrence distinct).  The distinct k-mers of the table are then looked
//                  up ONCE each in the table of reference k-mers: those add their count to idcnt (the few per cent of
//                  leftover k-mers that are reference k-mers on another diagonal), the others are the novel k-mers.
//
// The same kernels serve the read-sharded deep sample (W = weighted entries, MODE of k_bin_count):
//   MODE 0  occurrences → counted list (cut-offs applied): one rank holds the whole file
//   MODE 1  occurrences → this rank's distinct (k-mer, partial count) pairs, written at the bin's own offset and
//           compacted by k_pairs_compact in bin order, so the pairs of an owner rank (a contiguous range of bins)
//           are contiguous; reference k-mers still go to idcnt (all-reduced afterwards)
//   MODE 2  weighted pairs received from every rank → counted list (cut-offs applied to the merged counts,
//           src/call.rs:1172-1173): the owner's final pass
#pragma once
#include "bk_core.cuh"

namespace bk {

#define BK_BIN_G_THREADS 1024
#define BK_BIN_G_PER_SM 1                    // CTAs of k_bin_hist / k_bin_scatter per SM (2 measured slower: shorter runs per bin)
#define BK_BIN_SLOTS 4096                    // shared-memory table of k_bin_count: 4096 x (8 + 4 + 2) bytes
#define BK_BIN_ROUND 2048                    // occurrences one worst-case round may hold (distinct <= occurrences <= half the slots, in expectation)
#define BK_BIN_FILL 3072                     // distinct k-mers at which an optimistic single round gives up
#define BK_BIN_SMEM (BK_BIN_SLOTS * 14)       // keys, counts, list of occupied slots
#ifndef BK_BIN_AGG_MIN
#define BK_BIN_AGG_MIN (2 * BK_BIN_ROUND)    // bins with more occurrences than this merge equal keys per warp before the shared-memory atomic
#endif
#define BK_OWNER_UNITS_LOG2 6                // owner ranks split the hash space in 64 units (a bin never straddles one: P >= 64)

__device__ __forceinline__ u64 bin_hash(u64 x) { return (x ^ (x >> 31)) * 0x9E3779B97F4A7C15ull; }
__device__ __forceinline__ u64 bloom_mask_dev(u64 h, u32 log2w) {            // bk_host.h: bloom_mask
    const u64 r = h >> (64 - log2w - 18);
    return (1ull << (r & 63)) | (1ull << ((r >> 6) & 63)) | (1ull << ((r >> 12) & 63));
}
// first hash unit (of 64) owned by rank r of n: r owns units [unit_lo(r), unit_lo(r + 1))
__host__ __device__ __forceinline__ u32 owner_unit_lo(u32 r, u32 n) { return (r * (1u << BK_OWNER_UNITS_LOG2) + n - 1) / n; }

struct BinView {
    const u64* nov; const u32* nov_w;                    // the list; weights (W kernels) or null
    const u32* nov_n; u32 nov_cap;                       // n = min(*nov_n, nov_cap)
    u64* sorted; u32* sorted_w;                          // the list grouped by bin
    u32* cnt;                                            // P * G + 1 counters, then their exclusive prefix
    u32 log2p; u32 G;
    const ExactSlotD* exact; u32 exact_shift, exact_mask;   // reference k-mer → representative raw slot
    const u32* slot2id; u32* idcnt;                          // raw slot → distinct reference k-mer id → its count
    u64* pair_k; u32* pair_c; u32* dcount;               // MODE 1: pairs of bin b at [cnt[b * G], + dcount[b])
    // mismatch lines (bk_dense.cuh; null = off): (j << 58 | k-mer without digit j) → id of the reference k-mer that equals
    // the k-mer everywhere but at digit j; a distinct k-mer of the list that is the string of an unambiguous cell joins it
    const ExactSlotD* nb; u32 nb_shift, nb_mask; u32 k;
    const u64* nb_bloom; u32 nb_bloom_log2;              // blocked bit set over the keys of nb (bk_host.h: bloom_mask; 16 MB that stay in L2): most k-mers of the list have no neighbour at all
    const u32* id_amb; const u32* id_rep; u32* dense; u8* dense_flag;
};

__device__ __forceinline__ void bin_chunk(const BinView& b, u32* lo, u32* hi) {
    const u32 n = min(*b.nov_n, b.nov_cap);
    const u32 per = (n + b.G - 1) / b.G;
    *lo = min(n, blockIdx.x * per);
    *hi = min(n, *lo + per);
}

// grid G, BK_BIN_G_THREADS threads, (1 << log2p) * 4 bytes of shared memory
__global__ void __launch_bounds__(BK_BIN_G_THREADS) k_bin_hist(BinView b) {
    extern __shared__ u32 bh[];
    const u32 P = 1u << b.log2p;
    for (u32 i = threadIdx.x; i < P; i += blockDim.x) bh[i] = 0;
    __syncthreads();
    u32 lo, hi;
    bin_chunk(b, &lo, &hi);
    for (u32 i0 = lo; i0 < hi; i0 += 4 * BK_BIN_G_THREADS) {           // four loads in flight per thread
        u64 key[4];
#pragma unroll
        for (u32 j = 0; j < 4; j++) { const u32 i = i0 + j * BK_BIN_G_THREADS + threadIdx.x; key[j] = i < hi ? __ldg(b.nov + i) : BK_HOLE; }
#pragma unroll
        for (u32 j = 0; j < 4; j++) if (key[j] != BK_HOLE) atomicAdd(bh + (u32)(bin_hash(key[j]) >> (64 - b.log2p)), 1u);
    }
    __syncthreads();
    for (u32 i = threadIdx.x; i < P; i += blockDim.x) b.cnt[(size_t)i * b.G + blockIdx.x] = bh[i];
}

// cnt now holds exclusive prefixes: this CTA's share of bin i starts at cnt[i * G + cta]
template <bool W>
__global__ void __launch_bounds__(BK_BIN_G_THREADS) k_bin_scatter(BinView b) {
    extern __shared__ u32 bh[];
    const u32 P = 1u << b.log2p;
    for (u32 i = threadIdx.x; i < P; i += blockDim.x) bh[i] = b.cnt[(size_t)i * b.G + blockIdx.x];
    __syncthreads();
    u32 lo, hi;
    bin_chunk(b, &lo, &hi);
    for (u32 i0 = lo; i0 < hi; i0 += 4 * BK_BIN_G_THREADS) {
        u64 key[4]; u32 wt[4];
#pragma unroll
        for (u32 j = 0; j < 4; j++) {
            const u32 i = i0 + j * BK_BIN_G_THREADS + threadIdx.x;
            key[j] = i < hi ? __ldg(b.nov + i) : BK_HOLE;
            wt[j] = (W && i < hi) ? __ldg(b.nov_w + i) : 1u;
        }
#pragma unroll
        for (u32 j = 0; j < 4; j++) if (key[j] != BK_HOLE) {
            const u32 pos = atomicAdd(bh + (u32)(bin_hash(key[j]) >> (64 - b.log2p)), 1u);
            b.sorted[pos] = key[j];
            if (W) b.sorted_w[pos] = wt[j];
        }
    }
}

// grid P, 256 threads, BK_BIN_SMEM bytes of shared memory.
// A round: clear the table; insert (the keys of up to eight iterations are loaded before the first is inserted; the
// thread that claims an empty slot also appends it to the list of occupied slots); then, over that dense list:
// look the distinct k-mers up in the reference table (first probes of four k-mers together; a reference k-mer adds
// its count to idcnt and drops out) and compact what is left to the output of the MODE.
template <int MODE, bool W>
__global__ void __launch_bounds__(256, 3) k_bin_count(BinView b, CompactArgs a, u32* full) {
    extern __shared__ __align__(16) u8 bsm[];
    __shared__ u32 s_base, s_nocc, s_abort, s_out;
    u64* keys = reinterpret_cast<u64*>(bsm);
    u32* cnts = reinterpret_cast<u32*>(bsm + BK_BIN_SLOTS * 8);
    unsigned short* occ = reinterpret_cast<unsigned short*>(bsm + BK_BIN_SLOTS * 12);     // occupied slots, in claim order
    const u32 lane = threadIdx.x & 31;
    const u32 s = b.cnt[(size_t)blockIdx.x * b.G], e = b.cnt[(size_t)(blockIdx.x + 1) * b.G];
    if (e == s) { if (MODE == 1 && threadIdx.x == 0) b.dcount[blockIdx.x] = 0; return; }
    const u32 rounds_safe = (e - s + BK_BIN_ROUND - 1) / BK_BIN_ROUND;
    u32 rounds = 1;                                                            // optimistic: everything in one round
    u32 uniq = 0; u64 total = 0;
    if (threadIdx.x == 0) s_out = 0;
    for (u32 r = 0; r < rounds; r++) {
        {                                                                      // clear: 16-byte stores
            uint4* k4 = reinterpret_cast<uint4*>(keys); uint4* c4 = reinterpret_cast<uint4*>(cnts);
            const uint4 ones = make_uint4(0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu), zero = make_uint4(0u, 0u, 0u, 0u);
#pragma unroll
            for (u32 i = 0; i < BK_BIN_SLOTS / 2 / 256; i++) k4[i * 256 + threadIdx.x] = ones;
#pragma unroll
            for (u32 i = 0; i < BK_BIN_SLOTS / 4 / 256; i++) c4[i * 256 + threadIdx.x] = zero;
        }
        if (threadIdx.x == 0) { s_nocc = 0; s_abort = 0; }
        __syncthreads();
        const bool optimistic = rounds == 1 && rounds_safe > 1;                // may give up; a worst-case-sized round never does
        // a k-mer repeated a million times must not serialise on one shared-memory word: in bins that are far larger
        // than average (that is where such a k-mer lands) the lanes holding the key of the first taking lane let that
        // lane add for all of them.  Ordinary bins skip the two ballots and the shuffles: ~45 % of their keys are distinct.
        const bool merge_equal = (e - s) > BK_BIN_AGG_MIN;
        const u32 n_round = ((e - s + 31) & ~31u);
        for (u32 i0 = 0; i0 < n_round; i0 += 8 * 256) {
            u64 kq[8]; u32 wq[8];
#pragma unroll
            for (u32 j = 0; j < 8; j++) {
                const u32 i = i0 + j * 256 + threadIdx.x;
                kq[j] = (s + i < e) ? __ldg(b.sorted + s + i) : BK_HOLE;
                wq[j] = (W && s + i < e) ? __ldg(b.sorted_w + s + i) : 1u;
            }
#pragma unroll
            for (u32 j = 0; j < 8; j++) {
                if (i0 + j * 256 + (threadIdx.x & ~31u) >= n_round) break;        // warp-uniform
                if (optimistic && __any_sync(0xFFFFFFFFu, *(volatile u32*)&s_abort != 0)) break;   // warp-uniform: the table is filling up
                const u64 key = kq[j];
                const u64 h = bin_hash(key);
                bool take = key != BK_HOLE;
                if (rounds > 1) take = take && ((u32)((h * 0xD6E8FEB86659FD93ull) >> 40) % rounds) == r;     // (uniform branch: one round is the rule)
                u32 w = wq[j];
                const u32 tm = merge_equal ? __ballot_sync(0xFFFFFFFFu, take) : 0u;
                if (tm) {
                    const u32 first = (u32)__ffs(tm) - 1;
                    const u64 key0 = __shfl_sync(0xFFFFFFFFu, key, first);       // (every lane: no short-circuit around it)
                    const u32 same = __ballot_sync(0xFFFFFFFFu, take && key == key0);
                    if (__popc(same) >= 4) {
                        u32 ws = (same >> lane) & 1 ? w : 0u;                    // sum of the weights of the lanes holding key0
#pragma unroll
                        for (int o = 16; o; o >>= 1) ws += __shfl_xor_sync(0xFFFFFFFFu, ws, o);
                        if (lane == first) w = ws; else if ((same >> lane) & 1) take = false;
                    }
                }
                if (take) {
                    u32 slot = (u32)(h >> (52 - b.log2p)) & (BK_BIN_SLOTS - 1);
                    u32 probes = 0;
                    for (;;) {
                        const u64 old = atomicCAS(reinterpret_cast<unsigned long long*>(keys + slot), (unsigned long long)BK_HOLE, (unsigned long long)key);
                        if (old == BK_HOLE) {                                    // (one atomic per warp through a ballot: measured slower)
                            const u32 at = atomicAdd(&s_nocc, 1u);
                            if (at < BK_BIN_SLOTS) occ[at] = (unsigned short)slot;
                            if (at >= BK_BIN_FILL) s_abort = 1;
                        }
                        if (old == BK_HOLE || old == key) { atomicAdd(cnts + slot, w); break; }
                        slot = (slot + 1) & (BK_BIN_SLOTS - 1);
                        if (++probes >= BK_BIN_SLOTS) { s_abort = 1; break; }
                    }
                }
            }
        }
        __syncthreads();
        if (s_abort) {                                     // block-uniform
            __syncthreads();
            if (optimistic) { rounds = rounds_safe; r = 0xFFFFFFFFu; continue; }        // nothing was emitted yet: start over in worst-case rounds
            if (threadIdx.x == 0) *full = 1;               // a worst-case round overflowed (adversarial input): reported, results invalid
            return;
        }
        // the occupied slots, four per thread and pass: reference k-mers leave (their count goes to idcnt), what fails
        // the cut-offs is dropped, the rest is compacted to the counted list
        const u32 n_occ = s_nocc;
        __syncthreads();                                 // (the next round resets the counter)
        for (u32 p0 = 0; p0 < n_occ; p0 += 4 * 256) {
            u64 kk[4]; u32 cc[4], hh[4]; ExactSlotD e0[4];
            u32 mine = 0;
#pragma unroll
            for (u32 j = 0; j < 4; j++) {
                const u32 p = p0 + j * 256 + threadIdx.x;
                kk[j] = BK_HOLE; cc[j] = 0; hh[j] = 0;
                e0[j].key = BK_EMPTY; e0[j].gidx = 0; e0[j].oseq = 0;
                if (p < n_occ) {
                    const u32 slot = occ[p];
                    kk[j] = keys[slot]; cc[j] = cnts[slot];
                    if (MODE != 2) {
                        hh[j] = hash_slot(kk[j], b.exact_shift);
                        e0[j] = load_exact(b.exact + hh[j]);
                    }
                }
            }
            bool is_ref[4];
#pragma unroll
            for (u32 j = 0; j < 4; j++) {
                is_ref[j] = false;
                if (kk[j] == BK_HOLE) continue;
                if (MODE != 2) {                         // (MODE 2: the senders already took the reference k-mers out)
                    u32 h = hh[j];
                    ExactSlotD sl = e0[j];
                    for (;;) {                           // same probe sequence as bk_core.cuh: exact_lookup
                        if (sl.key == kk[j]) { atomicAdd(b.idcnt + __ldg(b.slot2id + sl.gidx), cc[j]); is_ref[j] = true; break; }
                        if (sl.key == BK_EMPTY) break;
                        h = (h + 1) & b.exact_mask;
                        sl = load_exact(b.exact + h);
                    }
                }
            }
            if (MODE == 0 && b.nb) {
                // Is the k-mer the string of an unambiguous mismatch-line cell (bk_dense.cuh)?  k candidate cells — digit jj
                // replaced — probed seven at a time: the loads of a batch are independent, a thread waits for three or four
                // memory round trips per k-mer instead of k.
#pragma unroll 1
                for (u32 j = 0; j < 4; j++) {
                    if (kk[j] == BK_HOLE || is_ref[j]) continue;
                    const u64 K = kk[j];
                    u32 cand = 0;                                              // bit jj: the bit set does not rule cell jj out
#pragma unroll 1
                    for (u32 j0 = 0; j0 < b.k; j0 += 7) {
                        u64 bw[7], bm[7];
#pragma unroll
                        for (u32 t = 0; t < 7; t++) {                          // seven independent loads that hit L2
                            const u32 jj = min(j0 + t, b.k - 1);
                            const u64 hb = bin_hash(((u64)jj << 58) | (K & ~(3ull << (2 * (b.k - 1 - jj)))));
                            bm[t] = bloom_mask_dev(hb, b.nb_bloom_log2);
                            bw[t] = __ldg(b.nb_bloom + (hb >> (64 - b.nb_bloom_log2)));
                        }
#pragma unroll
                        for (u32 t = 0; t < 7; t++) if (j0 + t < b.k && (bw[t] & bm[t]) == bm[t]) cand |= 1u << (j0 + t);
                    }
                    // (one k-mer in ~250 gets here without being a neighbour: a warp no longer walks the big table for
                    // every cell because one of its lanes might have to)
                    for (; cand; cand &= cand - 1) {
                        const u32 jj = (u32)__ffs((int)cand) - 1u;
                        const u64 key = ((u64)jj << 58) | (K & ~(3ull << (2 * (b.k - 1 - jj))));
                        u32 h = hash_slot(key, b.nb_shift);
                        ExactSlotD s1 = load_exact(b.nb + h);
                        while (s1.key != key && s1.key != BK_EMPTY) { h = (h + 1) & b.nb_mask; s1 = load_exact(b.nb + h); }
                        if (s1.key != key) continue;
                        // a neighbour: if it is ambiguous every neighbour is, the k-mer stays here
                        const u32 id = s1.gidx;
                        if (((__ldg(b.id_amb + id) >> jj) & 1u) == 0) {
                            const u32 line = (__ldg(b.id_rep + id) + jj) * 4u + (u32)((K >> (2 * (b.k - 1 - jj))) & 3);
                            atomicAdd(b.dense + (size_t)line * (b.k + 1) + jj, cc[j]);
                            b.dense_flag[line] = 1;
                            is_ref[j] = true;                                  // leaves the bins: counted with its cell
                        }
                        break;
                    }
                }
            }
#pragma unroll
            for (u32 j = 0; j < 4; j++) {
                if (kk[j] == BK_HOLE) continue;
                bool keep = !is_ref[j];
                if (MODE != 1) {
                    if (!is_ref[j]) { uniq++; total += cc[j]; }
                    keep = keep && cc[j] >= a.ci && cc[j] <= 1000000000u;
                }
                if (keep) mine++; else kk[j] = BK_HOLE;
            }
            u32 tot;
            u32 o = block_excl_scan_256(mine, &tot);
            if (threadIdx.x == 0) {
                if (MODE == 1) { s_base = s + s_out; s_out += tot; }
                else s_base = tot ? atomicAdd(&a.fc->n_counted, tot) : 0u;
            }
            __syncthreads();
            o += s_base;
#pragma unroll
            for (u32 j = 0; j < 4; j++)
                if (kk[j] != BK_HOLE) {
                    if (MODE == 1) { b.pair_k[o] = kk[j]; b.pair_c[o] = cc[j]; }            // (o < e: distinct <= occurrences)
                    else if (o < a.out_cap) { a.out_kmers[o] = kk[j]; a.out_counts[o] = min(cc[j], a.cs); }
                    o++;
                }
            __syncthreads();
        }
    }
    if (MODE == 1) { if (threadIdx.x == 0) b.dcount[blockIdx.x] = s_out; return; }
    // one pair of atomics per CTA (thousands of CTAs, one address each)
    __shared__ u32 s_uniq; __shared__ unsigned long long s_total;
    if (threadIdx.x == 0) { s_uniq = 0; s_total = 0; }
    __syncthreads();
    uniq = warp_sum_u32(uniq); total = warp_sum_u64(total);
    if (lane == 0 && uniq) { atomicAdd(&s_uniq, uniq); atomicAdd(&s_total, (unsigned long long)total); }
    __syncthreads();
    if (threadIdx.x == 0 && s_uniq) { atomicAdd(&a.fc->unique, s_uniq); atomicAdd((unsigned long long*)&a.fc->total_kmers, s_total); }
}

// exclusive prefix in place (after k_diff_blocksum / k_diff_scan_bsum over the same array); element n receives the total
__global__ void __launch_bounds__(BK_PS_THREADS) k_excl_apply(u32* a, u32 n, const u32* __restrict__ bsum) {
    const u32 base = blockIdx.x * BK_PS_BLOCK + threadIdx.x * BK_PS_PER_THREAD;
    u32 d[BK_PS_PER_THREAD];
    u32 s = 0;
#pragma unroll
    for (u32 i = 0; i < BK_PS_PER_THREAD; i++) { d[i] = (base + i < n) ? a[base + i] : 0; s += d[i]; }
    u32 run = bsum[blockIdx.x] + block_excl_scan_256(s, nullptr);
#pragma unroll
    for (u32 i = 0; i < BK_PS_PER_THREAD; i++) {
        if (base + i < n) a[base + i] = run;
        run += d[i];
        if (base + i + 1 == n) a[n] = run;
    }
}

// MODE 1 epilogue (read-sharded sample): dcount holds the exclusive prefix of the per-bin pair counts (P + 1 entries);
// CTA b moves the pairs of bin b from the bin's own offset to their dense position, and CTA 0 writes where every
// owner rank's pairs start (own_off[r], r = 0..n_ranks; an owner's bins are contiguous).
__global__ void __launch_bounds__(256) k_pairs_compact(BinView b, const u32* __restrict__ dprefix, u64* __restrict__ out_k, u32* __restrict__ out_c,
                                                       u32 n_ranks, u32* own_off) {
    const u32 P = 1u << b.log2p;
    if (blockIdx.x == 0)
        for (u32 r = threadIdx.x; r <= n_ranks; r += blockDim.x)
            own_off[r] = r == n_ranks ? dprefix[P] : dprefix[owner_unit_lo(r, n_ranks) << (b.log2p - BK_OWNER_UNITS_LOG2)];
    const u32 src = b.cnt[(size_t)blockIdx.x * b.G];
    const u32 dst = dprefix[blockIdx.x], n = dprefix[blockIdx.x + 1] - dst;
    for (u32 i = threadIdx.x; i < n; i += blockDim.x) { out_k[dst + i] = b.pair_k[src + i]; out_c[dst + i] = b.pair_c[src + i]; }
}

}  // namespace bk
