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
//   k_bin_count    one CTA per bin: shared-memory open addressing; bins with more occurrences than the table can
//                  safely take are walked in rounds by a second hash.  The distinct k-mers of the table are then looked
//                  up ONCE each in the table of reference k-mers: those add their count to idcnt (the few per cent of
//                  leftover k-mers that are reference k-mers on another diagonal), the others are the novel k-mers
#pragma once
#include "bk_core.cuh"

namespace bk {

#define BK_BIN_G_THREADS 1024
#define BK_BIN_G_PER_SM 1                    // CTAs of k_bin_hist / k_bin_scatter per SM (2 measured slower: shorter runs per bin)
#define BK_BIN_SLOTS 4096                    // shared-memory table of k_bin_count: 4096 x (8 + 4 + 2) bytes
#define BK_BIN_ROUND 2048                    // occurrences one round may hold (distinct <= occurrences <= half the slots, in expectation)
#define BK_BIN_SMEM (BK_BIN_SLOTS * 14)       // keys, counts, list of occupied slots

__device__ __forceinline__ u64 bin_hash(u64 x) { return (x ^ (x >> 31)) * 0x9E3779B97F4A7C15ull; }

struct BinView {
    const u64* nov; const u32* nov_n; u32 nov_cap;      // the list (n = min(*nov_n, nov_cap))
    u64* sorted;                                         // the list grouped by bin
    u32* cnt;                                            // P * G + 1 counters, then their exclusive prefix
    u32 log2p; u32 G;
    const ExactSlotD* exact; u32 exact_shift, exact_mask;   // reference k-mer → representative raw slot
    const u32* slot2id; u32* idcnt;                          // raw slot → distinct reference k-mer id → its count
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
__global__ void __launch_bounds__(BK_BIN_G_THREADS) k_bin_scatter(BinView b) {
    extern __shared__ u32 bh[];
    const u32 P = 1u << b.log2p;
    for (u32 i = threadIdx.x; i < P; i += blockDim.x) bh[i] = b.cnt[(size_t)i * b.G + blockIdx.x];
    __syncthreads();
    u32 lo, hi;
    bin_chunk(b, &lo, &hi);
    for (u32 i0 = lo; i0 < hi; i0 += 4 * BK_BIN_G_THREADS) {
        u64 key[4];
#pragma unroll
        for (u32 j = 0; j < 4; j++) { const u32 i = i0 + j * BK_BIN_G_THREADS + threadIdx.x; key[j] = i < hi ? __ldg(b.nov + i) : BK_HOLE; }
#pragma unroll
        for (u32 j = 0; j < 4; j++) if (key[j] != BK_HOLE) b.sorted[atomicAdd(bh + (u32)(bin_hash(key[j]) >> (64 - b.log2p)), 1u)] = key[j];
    }
}

// grid P, 256 threads, BK_BIN_SMEM bytes of shared memory.  Appends to the counted list of the file.
// A round: clear the table; insert (the keys of up to eight iterations are loaded before the first is inserted; the
// thread that claims an empty slot also appends it to the list of occupied slots); then, over that dense list:
// look the distinct k-mers up in the reference table (first probes of four k-mers together; a reference k-mer adds
// its count to idcnt and drops out) and compact what passes the KMC cut-offs to the counted list.
__global__ void __launch_bounds__(256, 3) k_bin_count(BinView b, CompactArgs a, u32* full) {
    extern __shared__ __align__(16) u8 bsm[];
    __shared__ u32 s_base, s_nocc;
    u64* keys = reinterpret_cast<u64*>(bsm);
    u32* cnts = reinterpret_cast<u32*>(bsm + BK_BIN_SLOTS * 8);
    unsigned short* occ = reinterpret_cast<unsigned short*>(bsm + BK_BIN_SLOTS * 12);     // occupied slots, in claim order
    const u32 lane = threadIdx.x & 31;
    const u32 s = b.cnt[(size_t)blockIdx.x * b.G], e = b.cnt[(size_t)(blockIdx.x + 1) * b.G];
    if (e == s) return;
    const u32 rounds = (e - s + BK_BIN_ROUND - 1) / BK_BIN_ROUND;
    u32 uniq = 0; u64 total = 0;
    for (u32 r = 0; r < rounds; r++) {
        {                                                                      // clear: 16-byte stores
            uint4* k4 = reinterpret_cast<uint4*>(keys); uint4* c4 = reinterpret_cast<uint4*>(cnts);
            const uint4 ones = make_uint4(0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu), zero = make_uint4(0u, 0u, 0u, 0u);
#pragma unroll
            for (u32 i = 0; i < BK_BIN_SLOTS / 2 / 256; i++) k4[i * 256 + threadIdx.x] = ones;
#pragma unroll
            for (u32 i = 0; i < BK_BIN_SLOTS / 4 / 256; i++) c4[i * 256 + threadIdx.x] = zero;
        }
        if (threadIdx.x == 0) s_nocc = 0;
        __syncthreads();
        const u32 n_round = ((e - s + 31) & ~31u);
        for (u32 i0 = 0; i0 < n_round; i0 += 8 * 256) {
            u64 kq[8];
#pragma unroll
            for (u32 j = 0; j < 8; j++) { const u32 i = i0 + j * 256 + threadIdx.x; kq[j] = (s + i < e) ? __ldg(b.sorted + s + i) : BK_HOLE; }
#pragma unroll
            for (u32 j = 0; j < 8; j++) {
                if (i0 + j * 256 + (threadIdx.x & ~31u) >= n_round) break;        // warp-uniform
                const u64 key = kq[j];
                const u64 h = bin_hash(key);
                bool take = key != BK_HOLE;
                if (rounds > 1) take = take && ((u32)((h * 0xD6E8FEB86659FD93ull) >> 40) % rounds) == r;     // (uniform branch: one round is the rule)
                // a k-mer repeated a million times must not serialise on one shared-memory word: if many lanes hold
                // the key of the first taking lane, that lane adds for all of them
                u32 w = 1;
                const u32 tm = __ballot_sync(0xFFFFFFFFu, take);
                if (tm) {
                    const u32 first = (u32)__ffs(tm) - 1;
                    const u64 key0 = __shfl_sync(0xFFFFFFFFu, key, first);       // (every lane: no short-circuit around it)
                    const u32 same = __ballot_sync(0xFFFFFFFFu, take && key == key0);
                    if (__popc(same) >= 4) { if (lane == first) w = __popc(same); else if ((same >> lane) & 1) take = false; }
                }
                if (take) {
                    u32 slot = (u32)(h >> (52 - b.log2p)) & (BK_BIN_SLOTS - 1);
                    u32 probes = 0;
                    for (;;) {
                        const u64 old = atomicCAS(reinterpret_cast<unsigned long long*>(keys + slot), (unsigned long long)BK_HOLE, (unsigned long long)key);
                        if (old == BK_HOLE) occ[atomicAdd(&s_nocc, 1u)] = (unsigned short)slot;     // (one atomic per warp through a ballot: measured slower)
                        if (old == BK_HOLE || old == key) { atomicAdd(cnts + slot, w); break; }
                        slot = (slot + 1) & (BK_BIN_SLOTS - 1);
                        if (++probes >= BK_BIN_SLOTS) { *full = 1; break; }
                    }
                }
            }
        }
        __syncthreads();
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
                    hh[j] = hash_slot(kk[j], b.exact_shift);
                    e0[j] = load_exact(b.exact + hh[j]);
                }
            }
#pragma unroll
            for (u32 j = 0; j < 4; j++) {
                if (kk[j] == BK_HOLE) continue;
                u32 h = hh[j];
                ExactSlotD sl = e0[j];
                bool is_ref = false;
                for (;;) {                               // same probe sequence as bk_core.cuh: exact_lookup
                    if (sl.key == kk[j]) { atomicAdd(b.idcnt + __ldg(b.slot2id + sl.gidx), cc[j]); is_ref = true; break; }
                    if (sl.key == BK_EMPTY) break;
                    h = (h + 1) & b.exact_mask;
                    sl = load_exact(b.exact + h);
                }
                if (!is_ref) { uniq++; total += cc[j]; }
                const bool keep = !is_ref && cc[j] >= a.ci && cc[j] <= 1000000000u;
                if (keep) mine++; else kk[j] = BK_HOLE;
            }
            u32 tot;
            u32 o = block_excl_scan_256(mine, &tot);
            if (threadIdx.x == 0) s_base = tot ? atomicAdd(&a.fc->n_counted, tot) : 0u;
            __syncthreads();
            o += s_base;
#pragma unroll
            for (u32 j = 0; j < 4; j++)
                if (kk[j] != BK_HOLE) { if (o < a.out_cap) { a.out_kmers[o] = kk[j]; a.out_counts[o] = min(cc[j], a.cs); } o++; }
            __syncthreads();
        }
    }
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

}  // namespace bk
