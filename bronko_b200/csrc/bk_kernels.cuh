// bk_kernels.cuh — sm_100a kernels of the k-mer→pileup path.  Included only by bk_device.cu.
//
//  counting  : k_scan (pack + seed/extend, difference-array runs), k_leftover (per-k-mer counting),
//              k_diff_* (prefix sum + fold onto distinct reference k-mers), k_compact_* (the "KMC dump")
//  mapping   : k_map<STATS|PILEUP>   — reference src/call.rs:1257-1434
//  selection : k_select               — src/call.rs:422-502
//  scoring   : k_noise_prep / k_noise_seq / k_noise_tau — src/call.rs:799-967 ; k_call — src/call.rs:969-1150
#pragma once
#include <cuda_runtime.h>

#include "bk_core.cuh"

namespace bk {

struct BucketSlotD { u64 key; u32 off; u32 len; };
struct BucketEntryD { u32 row; unsigned short file_id; u8 idx; u8 canonical; };

struct FileCounters { u32 n_counted; u32 gen_new; u32 unique; u32 pad; u64 total_kmers; };
struct Counters {
    u32 n_desc; u32 gen_full; u32 var_overflow; u32 pad0;
    FileCounters f[2];
    i32 best; u32 n_var; u32 n_major; u32 n_minor;
    u64 pos_covered; u64 total_cov;
};

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
__global__ void __launch_bounds__(BK_SCAN_THREADS)
k_scan(CountView v, const u8* __restrict__ bases, const u32* __restrict__ off, u32 off_bias, u32 r_begin, u32 r_end,
       u32 tile_reads, u32 tile_bytes, u32* gen_new) {
    extern __shared__ __align__(16) u8 smem[];
    const u32 n_tiles = (r_end - r_begin + tile_reads - 1) / tile_reads;
    const u64* refpk = v.refpk;
    auto ldr = [refpk](u32 i) { return __ldg(refpk + i); };
    u32 created = 0;
    for (u32 tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const u32 r0 = r_begin + tile * tile_reads;
        const u32 r1 = min(r0 + tile_reads, r_end);
        const u32 a = (__ldg(off + r0) - off_bias) & ~15u;
        const u32 e = (__ldg(off + r1) - off_bias + 15u) & ~15u;
        const bool staged = (e - a) <= tile_bytes;          // uniform over the CTA
        const u32 r = r0 + threadIdx.x;
        u32 o0 = 0, len = 0;
        if (r < r1) { o0 = __ldg(off + r) - off_bias; len = __ldg(off + r + 1) - off_bias - o0; }
        if (staged) {
            __syncthreads();                                 // previous tile fully consumed
            const uint4* src = reinterpret_cast<const uint4*>(bases + a);
            uint4* dst = reinterpret_cast<uint4*>(smem);
            const u32 n16 = (e - a) >> 4;
            for (u32 i = threadIdx.x; i < n16; i += BK_SCAN_THREADS) {
                uint4 q;
                asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
                             : "=r"(q.x), "=r"(q.y), "=r"(q.z), "=r"(q.w) : "l"(src + i));
                dst[i] = q;
            }
            __syncthreads();
            if (r < r1) {
                const u32* sw = reinterpret_cast<const u32*>(smem);
                auto ld = [sw](u32 i) { return sw[i]; };
                created += scan_read(v, ld, ldr, o0 - a, len, a);
            }
        } else if (r < r1) {
            const u32* gw = reinterpret_cast<const u32*>(bases);
            auto ld = [gw](u32 i) { return __ldg(gw + i); };
            created += scan_read(v, ld, ldr, o0, len, 0);
        }
    }
    created = warp_sum_u32(created);
    if ((threadIdx.x & 31) == 0 && created) atomicAdd(gen_new, created);
}

// k_leftover: one warp per queued stretch; lane j counts k-mers j, j+32, ... of the stretch.
__global__ void __launch_bounds__(256)
k_leftover(CountView v, const u8* __restrict__ bases, u32* gen_new) {
    const u32 n = min(*v.n_desc, v.desc_cap);
    const u32 lane = threadIdx.x & 31;
    const u32 warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, n_warps = (gridDim.x * blockDim.x) >> 5;
    const u32* gw = reinterpret_cast<const u32*>(bases);
    auto ld = [gw](u32 i) { return __ldg(gw + i); };
    u32 created = 0;
    for (u32 i = warp; i < n; i += n_warps) {
        const uint2 d = v.desc[i];
        created += count_stretch(v, ld, d.x, d.y, lane, 32);
    }
    created = warp_sum_u32(created);
    if (lane == 0 && created) atomicAdd(gen_new, created);
}

__global__ void k_gen_init(GenSlot* gen, u64 n) {
    const u64 stride = (u64)gridDim.x * blockDim.x;
    uint4* p = reinterpret_cast<uint4*>(gen);
    const uint4 e = make_uint4(0xFFFFFFFFu, 0xFFFFFFFFu, 0u, 0u);
    for (u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) p[i] = e;
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
    if (have) { uniq++; total += c; }
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

__global__ void __launch_bounds__(256) k_compact_gen(CompactArgs a, const GenSlot* __restrict__ gen, u32 n_slots) {
    const u32 stride = gridDim.x * blockDim.x;
    u32 uniq = 0; u64 total = 0;
    for (u32 i = blockIdx.x * blockDim.x + threadIdx.x; i < n_slots; i += stride) {   // n_slots is a power of two >= 1024
        const uint4 s = __ldg(reinterpret_cast<const uint4*>(gen) + i);
        const u64 key = ((u64)s.y << 32) | s.x;
        compact_emit(a, key != BK_EMPTY, key, s.z, true, uniq, total);
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
};

__device__ __forceinline__ u64 revcomp_dev(u64 v, u32 k) {
    u64 x = ~v;
    x = ((x >> 2) & 0x3333333333333333ull) | ((x & 0x3333333333333333ull) << 2);
    x = ((x >> 4) & 0x0F0F0F0F0F0F0F0Full) | ((x & 0x0F0F0F0F0F0F0F0Full) << 4);
    const u32 lo = (u32)x, hi = (u32)(x >> 32);
    x = ((u64)__byte_perm(lo, 0, 0x0123) << 32) | __byte_perm(hi, 0, 0x0123);
    return x >> (64 - 2 * k);
}

template <int PILEUP>
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
        const u64 sum_mu = warp_sum_u64(mu);
        const u64 bucket = sum_mu - mu + val - num_a * cur + 1 + num_a;
        if (lane >= m.b0 && lane < m.b1) {
            u32 h = hash_slot(bucket, m.shift);
            u32 off = 0, len = 0;
            for (;;) {
                const uint4 s = __ldg(reinterpret_cast<const uint4*>(m.slots) + h);
                const u64 key = ((u64)s.y << 32) | s.x;
                if (key == bucket) { off = s.z; len = s.w; break; }
                if (key == BK_EMPTY) break;
                h = (h + 1) & m.mask;
            }
            for (u32 j = 0; j < len; j++) {
                const uint2 raw = __ldg(reinterpret_cast<const uint2*>(m.entries) + off + j);
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
// Noise baseline, src/call.rs:799-967 (quirks: SURVEY.md Appendix C, Q12).  The reference is one
// sequential loop; here it is split so that only what is inherently sequential runs on one thread:
//   k_noise_prep : per position, sorted minor-allele fractions (and their squares)      [parallel]
//   k_noise_seq  : the running n / s / s2 sums (exact same addition order, FP64, no FMA) on one
//                  thread and the 10-entry max table (with its evict-by-value quirk) on another,
//                  snapshotting both for every output position                          [sequential]
//   k_noise_tau  : the Thompson-tau rejection loop per position from the snapshots      [parallel]
// ------------------------------------------------------------------------------------------------
struct ScoreView {
    u32 n_genomes; const u32* genome_row0; const u32* genome_seq_off; const u32* seq_row0;
    const u8* ref_code; const Counters* ctr;
    const u32* pile; u32 pile_stride;
};

__global__ void __launch_bounds__(256) k_noise_prep(ScoreView sv, double* maf, double* msq) {
    const i32 best = sv.ctr->best;
    if (best < 0) return;
    const u32 rows = sv.genome_row0[best + 1] - sv.genome_row0[best];
    const u32 row = blockIdx.x * blockDim.x + threadIdx.x;
    if (row >= rows) return;
    const uint4 f = *reinterpret_cast<const uint4*>(sv.pile + row * 4);
    const uint4 r = *reinterpret_cast<const uint4*>(sv.pile + sv.pile_stride + row * 4);
    u64 c0 = (u64)f.x + r.x, c1 = (u64)f.y + r.y, c2 = (u64)f.z + r.z, c3 = (u64)f.w + r.w;
    u64 t;
#define BK_CSWAP(a, b) if (a < b) { t = a; a = b; b = t; }
    BK_CSWAP(c0, c1) BK_CSWAP(c2, c3) BK_CSWAP(c0, c2) BK_CSWAP(c1, c3) BK_CSWAP(c1, c2)
#undef BK_CSWAP
    const u64 total = c0 + c1 + c2 + c3;
    double m1 = 0.0, m2 = 0.0, m3 = 0.0;
    if (total != 0) { const double td = (double)total; m1 = (double)c1 / td; m2 = (double)c2 / td; m3 = (double)c3 / td; }
    maf[row * 3 + 0] = m1; maf[row * 3 + 1] = m2; maf[row * 3 + 2] = m3;
    msq[row * 3 + 0] = __dmul_rn(m1, m1); msq[row * 3 + 1] = __dmul_rn(m2, m2); msq[row * 3 + 2] = __dmul_rn(m3, m3);
}

#define BK_NOISE_WINDOW 100
#define BK_NOISE_HALF 50
#define BK_NOISE_TABLE 10

__global__ void __launch_bounds__(64)
k_noise_seq(ScoreView sv, const double* __restrict__ maf, const double* __restrict__ msq,
            u32* st_n, double* st_s, double* st_s2, double* st_max) {
    const i32 best = sv.ctr->best;
    if (best < 0) return;
    const u32 s = sv.genome_seq_off[best] + blockIdx.x;
    if (s >= sv.genome_seq_off[best + 1]) return;
    if (threadIdx.x != 0 && threadIdx.x != 32) return;
    const u32 r0 = sv.seq_row0[s] - sv.genome_row0[best];
    const u32 len = sv.seq_row0[s + 1] - sv.seq_row0[s];
    const double* mf = maf + (size_t)r0 * 3;
    const double* mq = msq + (size_t)r0 * 3;
    if (len < BK_NOISE_WINDOW) {     // the reference indexes out of bounds (panics) here; report zero noise
        if (threadIdx.x == 0) for (u32 i = 0; i < len; i++) { st_n[r0 + i] = 0; st_s[r0 + i] = 0.0; st_s2[r0 + i] = 0.0; }
        else for (u32 i = 0; i < len * BK_NOISE_TABLE; i++) st_max[(size_t)r0 * BK_NOISE_TABLE + i] = 0.0;
        return;
    }
    const u32 iters = len + BK_NOISE_HALF;
    if (threadIdx.x == 0) {
        // running sums: identical operation order to src/call.rs:845-895 (adding/subtracting an
        // exact 0.0 where the reference skips the update leaves every bit unchanged)
        u32 n = 0; double sum = 0.0, sum2 = 0.0;
        for (u32 i = 0; i < iters; i++) {
            const bool has_old = i >= BK_NOISE_WINDOW;          // position i-100 < len always holds
            const bool has_new = i < len;
#pragma unroll
            for (u32 j = 0; j < 3; j++) {
                const double old = has_old ? mf[(size_t)(i - BK_NOISE_WINDOW) * 3 + j] : 0.0;
                const double old2 = has_old ? mq[(size_t)(i - BK_NOISE_WINDOW) * 3 + j] : 0.0;
                const double nw = has_new ? mf[(size_t)i * 3 + j] : 0.0;
                const double nw2 = has_new ? mq[(size_t)i * 3 + j] : 0.0;
                if (old > 0.0) { n -= 1; sum = __dsub_rn(sum, old); sum2 = __dsub_rn(sum2, old2); }
                if (nw > 0.0) { n += 1; sum = __dadd_rn(sum, nw); sum2 = __dadd_rn(sum2, nw2); }
            }
            if (i >= BK_NOISE_HALF) { const u32 w = r0 + i - BK_NOISE_HALF; st_n[w] = n; st_s[w] = sum; st_s2[w] = sum2; }
        }
    } else {
        // max table, src/call.rs:857-892: evict the first entry within 1e-12 of the value leaving the
        // window (never refilled), insert by bubbling up with strict '>'.
        double m[BK_NOISE_TABLE];
#pragma unroll
        for (u32 q = 0; q < BK_NOISE_TABLE; q++) m[q] = 0.0;
        for (u32 i = 0; i < iters; i++) {
            const bool has_old = i >= BK_NOISE_WINDOW;
            const bool has_new = i < len;
#pragma unroll
            for (u32 j = 0; j < 3; j++) {
                const double old = has_old ? mf[(size_t)(i - BK_NOISE_WINDOW) * 3 + j] : 0.0;
                const double nw = has_new ? mf[(size_t)i * 3 + j] : 0.0;
                if (old > 0.0) {
                    // exact-safe shortcut: every entry is either >= m[9]+... ; if old is well below the
                    // smallest entry nothing can be within 1e-12 of it
                    if (!(m[BK_NOISE_TABLE - 1] > 0.0 && old < m[BK_NOISE_TABLE - 1] - 1e-9)) {
                        u32 pos = BK_NOISE_TABLE;
#pragma unroll
                        for (u32 q = BK_NOISE_TABLE; q-- > 0;) if (fabs(__dsub_rn(m[q], old)) < 1e-12) pos = q;
                        if (pos < BK_NOISE_TABLE) {
#pragma unroll
                            for (u32 q = 0; q < BK_NOISE_TABLE - 1; q++) if (q >= pos) m[q] = m[q + 1];
                            m[BK_NOISE_TABLE - 1] = 0.0;
                        }
                    }
                }
                if (nw > 0.0 && nw > m[BK_NOISE_TABLE - 1]) {
                    u32 gt = 0;
#pragma unroll
                    for (u32 q = 0; q < BK_NOISE_TABLE; q++) gt += (nw > m[q]) ? 1u : 0u;
                    const u32 pos = BK_NOISE_TABLE - gt;      // table is non-increasing: '>' holds on a suffix
#pragma unroll
                    for (u32 q = BK_NOISE_TABLE; q-- > 1;) if (q > pos) m[q] = m[q - 1];
#pragma unroll
                    for (u32 q = 0; q < BK_NOISE_TABLE; q++) if (q == pos) m[q] = nw;
                }
            }
            if (i >= BK_NOISE_HALF) {
                double* o = st_max + (size_t)(r0 + i - BK_NOISE_HALF) * BK_NOISE_TABLE;
#pragma unroll
                for (u32 q = 0; q < BK_NOISE_TABLE; q++) o[q] = m[q];
            }
        }
    }
}

__constant__ double c_tau[301];

__global__ void __launch_bounds__(256)
k_noise_tau(ScoreView sv, const u32* __restrict__ st_n, const double* __restrict__ st_s, const double* __restrict__ st_s2,
            const double* __restrict__ st_max, double* noise_max) {
    const i32 best = sv.ctr->best;
    if (best < 0) return;
    const u32 rows = sv.genome_row0[best + 1] - sv.genome_row0[best];
    const u32 row = blockIdx.x * blockDim.x + threadIdx.x;
    if (row >= rows) return;
    const u32 n = st_n[row];
    const double s = st_s[row], s2 = st_s2[row];
    const double* mx = st_max + (size_t)row * BK_NOISE_TABLE;
    double mu = 0.0, var = 0.0;
    if (n != 0) { mu = __ddiv_rn(s, (double)n); var = __dsub_rn(__ddiv_rn(s2, (double)n), __dmul_rn(mu, mu)); }
    u32 idx = 0, cn = n;
    double cs = s, cs2 = s2, cmu = mu, cvar = var;
    while (idx < BK_NOISE_TABLE && mx[idx] != 0.0) {              // src/call.rs:915-950
        const double cand = mx[idx];
        const double sd = sqrt(cvar);
        const double tau = (cn > 2) ? c_tau[cn <= 300 ? cn : 300] : __longlong_as_double(0x7FF0000000000000ll);
        if (fabs(__dsub_rn(cand, cmu)) > __dmul_rn(tau, sd)) {
            cs = __dsub_rn(cs, cand);
            cs2 = __dsub_rn(cs2, cand);                           // sic: candidate, not its square (src/call.rs:936)
            cn -= 1;
            if (cn > 0) { cmu = __ddiv_rn(cs, (double)cn); cvar = __dsub_rn(__ddiv_rn(cs2, (double)cn), __dmul_rn(cmu, cmu)); }
            else { cmu = 0.0; cvar = 0.0; }
            idx += 1;
        } else break;
    }
    noise_max[row] = idx < BK_NOISE_TABLE ? mx[idx] : 0.0;        // the reference would panic at idx == 10
}

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
