// bk_dense.cuh — mismatch lines: counting the k-mers that hold exactly ONE mismatch against the diagonal of their read
// without listing them (bk_core.cuh: emit_dense writes them; bk_host.h: slot2rep / id_amb / nb_slots).
//
// At 0.2 % sequencing error ~4 % of all k-mer occurrences are such k-mers (21 per error) — they were 95 % of the novel
// k-mer list, and that list (k_leftover → k_bin_hist → k_bin_scatter → k_bin_count) was a third of a sample's kernel time.
// The scan now spends two atomics per sequencing ERROR on line (reference base r, read base b); after the scan
//
//   k_dense_prefix   touched lines: prefix sum along j → cell (line, j) = occurrences of "reference k-mer at raw slot r - j
//                    with digit j replaced by b"
//   k_dense_fold     cells on a raw slot that is not the representative slot of its reference k-mer (another strain
//                    holding the same k-mer, a repeat) move to the representative's cell: one cell per k-mer STRING
//   k_dense_emit<0>  cells whose string another cell or a reference k-mer can spell too (id_amb) leave as weighted
//                    entries for the exact bins, with whatever the list holds (MODE 0; read-sharded ranks send ALL cells
//                    that way, their counts are partial: MODE 2)
//   (bins)           k_bin_count looks every distinct k-mer that reached it another way (two mismatches on its own
//                    diagonal, a foreign diagonal, reads without a seed ...) up in nb_slots: if it is the string of an
//                    unambiguous cell, its count joins the cell
//   k_dense_emit<1>  the remaining cells are final: KMC cut-offs (src/call.rs:1172-1173), counted list, line zeroed
//
// so the array is all zero again when the sample ends (it is cleared once, when it is allocated).
#pragma once
#include "bk_kernels.cuh"

namespace bk {

struct DenseView {
    u32* dense; u8* flag; u32 n_lines; u32 k;                // n_lines = n_raw * 4; (k + 1) counters per line
    const u32* slot2rep; const u32* slot2id; const u64* id_kmer; const u32* id_amb;
    const u32* line_amb; const u32* line_fold;               // per reference base: bit j = cell j of its lines is ambiguous / must be folded
    u64* nov; u32* nov_w; u32* nov_n; u32 nov_cap; u32* full;   // the weighted list of the file (ambiguous cells join it)
    // map shortcut (k_dense_emit<1, true>; bk_host.h): the one bucket an unambiguous cell can hit, tallies, all-genome pileups
    const uint2* id_bucket; MapView m; u32* gstats; u32* pile; u32 pile_stride;
};

// One THREAD per line (a warp looks at 32 flags with one coalesced load): the k + 1 counters of a line are 8-byte aligned
// and read with independent 8-byte loads, so a thread keeps a whole line in flight and a warp 32 of them — one line
// after the other per warp left these kernels waiting on one memory round trip per line.
#define BK_DENSE_MAXK 29
#define BK_DENSE_ROW ((BK_DENSE_MAXK + 2) / 2 * 2)       // counters held per thread (k + 1 rounded up to even)

__device__ __forceinline__ void dense_load(const u32* row, u32 k, u32* c) {
    const uint2* r2 = reinterpret_cast<const uint2*>(row);
#pragma unroll
    for (u32 i = 0; i < BK_DENSE_ROW / 2; i++) {
        uint2 v = make_uint2(0u, 0u);
        if (2 * i <= k) v = r2[i];                           // (a row is (k + 1) counters, k odd: a whole number of pairs)
        c[2 * i] = v.x; c[2 * i + 1] = v.y;
    }
}
__device__ __forceinline__ void dense_store(u32* row, u32 k, const u32* c) {
    uint2* r2 = reinterpret_cast<uint2*>(row);
#pragma unroll
    for (u32 i = 0; i < BK_DENSE_ROW / 2; i++) if (2 * i <= k) r2[i] = make_uint2(c[2 * i], c[2 * i + 1]);
}

// One WARP per 32 lines: the flags with one coalesced load, then the touched lines four at a time, lane j on counter j
// (a row is 88 contiguous bytes at k = 21: three or four sectors per warp instruction instead of one per lane).
__global__ void __launch_bounds__(256) k_dense_prefix(DenseView dv) {
    const u32 lane = threadIdx.x & 31, kk = dv.k + 1;
    const u32 n_warps = (gridDim.x * blockDim.x) >> 5;
    for (u32 w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; w * 32u < dv.n_lines; w += n_warps) {
        const u32 mine = w * 32u + lane;
        u32 fm = __ballot_sync(0xFFFFFFFFu, mine < dv.n_lines && dv.flag[mine] != 0);
        while (fm) {
            u32* row[4]; u32 c[4];
#pragma unroll
            for (u32 u = 0; u < 4; u++) {
                row[u] = fm ? dv.dense + (size_t)(w * 32u + (u32)__ffs((int)fm) - 1u) * kk : nullptr;
                fm &= fm - 1;
                c[u] = (row[u] && lane < kk) ? row[u][lane] : 0u;
            }
#pragma unroll
            for (u32 u = 0; u < 4; u++) {
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) { const u32 x = __shfl_up_sync(0xFFFFFFFFu, c[u], o); if (lane >= (u32)o) c[u] += x; }
                if (row[u] && lane < kk) row[u][lane] = c[u];                  // (the counter at k ends up as the line's sum: zero)
            }
        }
    }
}

__global__ void __launch_bounds__(256) k_dense_fold(DenseView dv) {
    for (u32 line = blockIdx.x * blockDim.x + threadIdx.x; line < dv.n_lines; line += gridDim.x * blockDim.x) {
        if (!dv.flag[line]) continue;
        const u32 refpos = line >> 2, alt = line & 3u;
        u32 fm = __ldg(dv.line_fold + refpos);                 // (most lines run along representative slots: nothing to move, the row is not read)
        if (!fm) continue;
        u32* row = dv.dense + (size_t)line * (dv.k + 1);
        for (; fm; fm &= fm - 1) {
            const u32 j = (u32)__ffs((int)fm) - 1u;
            const u32 c = row[j];
            if (!c) continue;
            const u32 slot = refpos - j, rep = __ldg(dv.slot2rep + slot);      // (a representative's cell is never moved: no one adds to a cell that leaves)
            const u32 to = (rep + j) * 4u + alt;
            atomicAdd(dv.dense + (size_t)to * (dv.k + 1) + j, c);
            dv.flag[to] = 1;
            row[j] = 0;
        }
    }
}

// MODE 0: ambiguous cells → weighted list; MODE 2: every cell → weighted list (and the line is cleared);
// MODE 1: every remaining cell → cut-offs → counted list, line cleared.
// The slots of what a CTA's lines append are reserved with ONE atomic per CTA and round of 256 lines (a reservation
// per line would be ~200,000 atomics on one address per file).
// MAP (MODE 1, databases the one-pass map serves): the kept cells are mapped right here — map_kmers for a k-mer that is
// one digit away from exactly one reference k-mer is one bucket (src/call.rs:1302-1385 with a single hit), no hashing —
// and go to the END of the counted list (FileCounters.n_dense), where the map kernel does not look.
template <int MODE, bool MAP>
__global__ void __launch_bounds__(256) k_dense_emit(DenseView dv, CompactArgs a) {
    __shared__ u32 s_base, s_n;
    __shared__ u32 s_tally[12];                              // [genome * 3 + {perfect, variant, unique}]
    __shared__ u16 s_cell[256 * BK_DENSE_MAXK];              // the leaving cells of the round's 256 lines: (thread << 5) | j
    if (MAP) { if (threadIdx.x < 12) s_tally[threadIdx.x] = 0; __syncthreads(); }
    u32 uniq = 0; u64 total = 0;
    const u32 n_round = (dv.n_lines + 255u) & ~255u;
    const u32 lane = threadIdx.x & 31, kk = dv.k + 1;
    for (u32 line = blockIdx.x * blockDim.x + threadIdx.x; line < n_round; line += gridDim.x * blockDim.x) {
        // phase 1, a warp per 32 lines: the flags with one load, then the live lines four at a time with lane j on cell j
        bool live = line < dv.n_lines && dv.flag[line] != 0;
        u32 ambm = 0;
        if (MODE == 0 && live) { ambm = __ldg(dv.line_amb + (line >> 2)); live = ambm != 0; }     // (no ambiguous cell on the line: the row is not read)
        if (threadIdx.x == 0) s_n = 0;
        __syncthreads();
        const u32 lm = __ballot_sync(0xFFFFFFFFu, live);
        const u32 wline0 = line - lane;
        for (u32 fm = lm; fm;) {
            u32 li[4], c[4];
#pragma unroll
            for (u32 u = 0; u < 4; u++) {
                li[u] = fm ? (u32)__ffs((int)fm) - 1u : 0xFFFFFFFFu;
                fm &= fm - 1;
                const u32 refpos = (wline0 + li[u]) >> 2;
                c[u] = (li[u] != 0xFFFFFFFFu && lane < dv.k && lane <= refpos) ? dv.dense[(size_t)(wline0 + li[u]) * kk + lane] : 0u;
            }
#pragma unroll
            for (u32 u = 0; u < 4; u++) {
                if (li[u] == 0xFFFFFFFFu) break;                               // warp-uniform
                bool out = c[u] != 0;
                if (MODE == 1) { if (out) { uniq++; total += c[u]; } out = out && c[u] >= a.ci && c[u] <= 1000000000u; }
                else if (MODE == 0) { const u32 am = __shfl_sync(0xFFFFFFFFu, ambm, li[u]); out = out && ((am >> lane) & 1u) != 0; }   // (every lane shuffles)
                const u32 om = __ballot_sync(0xFFFFFFFFu, out);
                if (!om) continue;
                u32 base = 0;
                if (lane == 0) base = atomicAdd(&s_n, (u32)__popc(om));
                base = __shfl_sync(0xFFFFFFFFu, base, 0);
                if (out) s_cell[base + __popc(om & ((1u << lane) - 1u))] = (u16)((((threadIdx.x & ~31u) + li[u]) << 5) | lane);
            }
        }
        __syncthreads();
        const u32 tot = s_n;
        if (threadIdx.x == 0) s_base = tot ? atomicAdd(MODE == 1 ? (MAP ? &a.fc->n_dense : &a.fc->n_counted) : dv.nov_n, tot) : 0u;
        __syncthreads();
        // the cells of the round, one per thread: a line with 21 cells (a real variant) no longer holds its warp
        // for 21 dependent look-ups while the other lanes idle
        const u32 line0 = line - threadIdx.x;
        for (u32 i = threadIdx.x; i < tot; i += 256) {
            const u32 cell = s_cell[i], cl = line0 + (cell >> 5), j = cell & 31u;
            const u32 crefpos = cl >> 2, alt = cl & 3u, at_i = s_base + i;
            u32* row = dv.dense + (size_t)cl * (dv.k + 1);
            const u32 cj = row[j];
            const u32 id = __ldg(dv.slot2id + (crefpos - j));                  // (after the fold every live cell sits on a representative slot)
            const u32 sh = 2 * (dv.k - 1 - j);
            const u64 km = (__ldg(dv.id_kmer + id) & ~(3ull << sh)) | ((u64)alt << sh);
            if (MODE != 1) {
                if (at_i < dv.nov_cap) { dv.nov[at_i] = km; dv.nov_w[at_i] = cj; } else *dv.full = 1;
                if (MODE == 0) row[j] = 0;
            } else if (!MAP) {
                if (at_i < a.out_cap) { a.out_kmers[at_i] = km; a.out_counts[at_i] = min(cj, a.cs); }
            } else {
                const u32 cnt = min(cj, a.cs);
                if (at_i < a.out_cap) { a.out_kmers[a.out_cap - 1 - at_i] = km; a.out_counts[a.out_cap - 1 - at_i] = cnt; }
                const u32 amb_id = __ldg(dv.id_amb + id);
                const u64 rev = revcomp_dev(km, dv.k);
                const bool rc = !(km < rev);                                   // src/lcb.rs:87-95
                const u32 jc = rc ? dv.k - 1 - j : j;
                u64 hits4 = 0;
                if (rc == ((amb_id >> 31) != 0) && jc >= dv.m.b0 && jc < dv.m.b1) {
                    const uint2 ol = __ldg(dv.id_bucket + (size_t)id * dv.k + j);
                    if (ol.y) map_walk<2>(dv.m, ol.x, ol.y, rc ? rev : km, rc, cnt, -1, 0u, dv.pile, dv.pile_stride, hits4);
                }
                if (hits4) {                                                   // src/call.rs:1389-1419
                    const u32 nb = dv.m.b1 - dv.m.b0;
                    u32 n_perfect = 0;
#pragma unroll
                    for (u32 g = 0; g < 4; g++) n_perfect += (((hits4 >> (16 * g)) & 0xFFFFu) == nb) ? 1u : 0u;
#pragma unroll
                    for (u32 g = 0; g < 4; g++) {
                        const u32 h = (u32)((hits4 >> (16 * g)) & 0xFFFFu);
                        if (!h) continue;
                        if (h == nb) { atomicAdd(s_tally + g * 3, 1u); if (n_perfect == 1) atomicAdd(s_tally + g * 3 + 2, 1u); }
                        else atomicAdd(s_tally + g * 3 + 1, 1u);
                    }
                }
            }
        }
        __syncthreads();                                                       // (the cells are read: s_base / s_cell are free, the lines can be cleared)
        if (MODE != 0) {                                                       // nothing is left on the lines
            for (u32 fm = lm; fm; fm &= fm - 1) {
                const u32 ln = wline0 + (u32)__ffs((int)fm) - 1u;
                if (lane < kk) dv.dense[(size_t)ln * kk + lane] = 0u;
            }
            if (live) dv.flag[line] = 0;
        }
    }
    if (MODE == 1) {
        uniq = warp_sum_u32(uniq); total = warp_sum_u64(total);
        if ((threadIdx.x & 31) == 0 && uniq) { atomicAdd(&a.fc->unique, uniq); atomicAdd((unsigned long long*)&a.fc->total_kmers, (unsigned long long)total); }
    }
    if (MAP) {
        __syncthreads();
        if (threadIdx.x < 12 && s_tally[threadIdx.x]) {
            const u32 g = threadIdx.x / 3, w = threadIdx.x % 3;
            if (g < dv.m.n_genomes) { atomicAdd(dv.gstats + g * 4 + w, s_tally[threadIdx.x]); if (w < 2) dv.gstats[g * 4 + 3] = 1; }
        }
    }
}

// k_compact_ids_map — the reference k-mers of a file (k_compact_ids) mapped where they are compacted, for the databases
// k_dense_emit<1, true> serves.  A reference k-mer hits every queried bucket of its own canonical form, and those are
// the buckets id_bucket already names: no hashing, no group probes.  They are most of the map's work (each one walks
// ~k buckets x the strains that share it; a k-mer with a sequencing error hits one bucket), and with a thread per k-mer
// 67,000 threads walked ~80 entries each one after the other.  Here a CTA takes 64 ids per round and spreads their
// (id, bucket index) pairs over its 256 threads; the hits of an id meet in shared memory (16-bit fields: hits16_ok).
// The k-mers go to the END of the counted list, like the cells (FileCounters.n_dense).
#define BK_IDMAP_IDS 64
__global__ void __launch_bounds__(256) k_compact_ids_map(DenseView dv, CompactArgs a, const u32* __restrict__ idcnt, u32 n_ids) {
    __shared__ u32 s_id[BK_IDMAP_IDS], s_cnt[BK_IDMAP_IDS], s_hlo[BK_IDMAP_IDS], s_hhi[BK_IDMAP_IDS];
    __shared__ u32 s_n, s_base, s_tally[12];
    const u32 k = dv.k, nb = dv.m.b1 - dv.m.b0;
    u32 uniq = 0; u64 total = 0;
    if (threadIdx.x < 12) s_tally[threadIdx.x] = 0;
    for (u32 base = blockIdx.x * BK_IDMAP_IDS; base < n_ids; base += gridDim.x * BK_IDMAP_IDS) {
        if (threadIdx.x == 0) s_n = 0;
        __syncthreads();
        if (threadIdx.x < BK_IDMAP_IDS) {
            const u32 i = base + threadIdx.x;
            const u32 c = i < n_ids ? idcnt[i] : 0u;
            if (c) {
                uniq++; total += c;
                if (c >= a.ci && c <= 1000000000u) {                           // src/call.rs:1172-1173
                    const u32 p = atomicAdd(&s_n, 1u);
                    s_id[p] = i; s_cnt[p] = min(c, a.cs); s_hlo[p] = 0; s_hhi[p] = 0;
                }
            }
        }
        __syncthreads();
        const u32 n = s_n;
        if (threadIdx.x == 0 && n) s_base = atomicAdd(&a.fc->n_dense, n);
        for (u32 item = threadIdx.x; item < n * k; item += 256) {
            const u32 p = item / k, j = item - p * k;
            const u32 id = s_id[p];
            const u64 km = __ldg(dv.id_kmer + id), rev = revcomp_dev(km, k);
            const bool rc = !(km < rev);                                       // src/lcb.rs:87-95
            const u32 jc = rc ? k - 1 - j : j;
            if (jc < dv.m.b0 || jc >= dv.m.b1) continue;
            const uint2 ol = __ldg(dv.id_bucket + (size_t)id * k + j);
            if (!ol.y) continue;
            u64 hits4 = 0;
            map_walk<2>(dv.m, ol.x, ol.y, rc ? rev : km, rc, s_cnt[p], -1, 0u, dv.pile, dv.pile_stride, hits4);
            if ((u32)hits4) atomicAdd(s_hlo + p, (u32)hits4);
            if ((u32)(hits4 >> 32)) atomicAdd(s_hhi + p, (u32)(hits4 >> 32));
        }
        __syncthreads();
        if (threadIdx.x < n) {
            const u32 at = s_base + threadIdx.x;
            if (at < a.out_cap) { a.out_kmers[a.out_cap - 1 - at] = __ldg(dv.id_kmer + s_id[threadIdx.x]); a.out_counts[a.out_cap - 1 - at] = s_cnt[threadIdx.x]; }
            const u64 hits4 = ((u64)s_hhi[threadIdx.x] << 32) | s_hlo[threadIdx.x];
            if (hits4) {                                                       // src/call.rs:1389-1419
                u32 n_perfect = 0;
#pragma unroll
                for (u32 g = 0; g < 4; g++) n_perfect += (((hits4 >> (16 * g)) & 0xFFFFu) == nb) ? 1u : 0u;
#pragma unroll
                for (u32 g = 0; g < 4; g++) {
                    const u32 h = (u32)((hits4 >> (16 * g)) & 0xFFFFu);
                    if (!h) continue;
                    if (h == nb) { atomicAdd(s_tally + g * 3, 1u); if (n_perfect == 1) atomicAdd(s_tally + g * 3 + 2, 1u); }
                    else atomicAdd(s_tally + g * 3 + 1, 1u);
                }
            }
        }
        __syncthreads();
    }
    uniq = warp_sum_u32(uniq); total = warp_sum_u64(total);
    if ((threadIdx.x & 31) == 0 && uniq) { atomicAdd(&a.fc->unique, uniq); atomicAdd((unsigned long long*)&a.fc->total_kmers, (unsigned long long)total); }
    __syncthreads();
    if (threadIdx.x < 12 && s_tally[threadIdx.x]) {
        const u32 g = threadIdx.x / 3, w = threadIdx.x % 3;
        if (g < dv.m.n_genomes) { atomicAdd(dv.gstats + g * 4 + w, s_tally[threadIdx.x]); if (w < 2) dv.gstats[g * 4 + 3] = 1; }
    }
}

}  // namespace bk
