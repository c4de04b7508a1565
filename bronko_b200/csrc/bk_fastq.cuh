// bk_fastq.cuh — FASTQ text → (bases, read offsets) on the GPU: the decode stage in front of the path (KMC's FASTQ
// reader inside count_kmers_kmc, reference src/call.rs:1166-1181; contract in SURVEY.md Appendix B: 4-line records, only
// the sequence line is used, '\r' before the line end dropped, a last line without '\n' counts).
//
// The text sits in device memory — inflated there by the GPU's decompression engine (BGZF), copied there after a host
// inflate (plain gzip) or as it is (plain text).  Newline j (0-based) ends line j; the sequence line of record r is
// line 4r + 1, i.e. it starts after newline 4r and ends at newline 4r + 1:
//
//   k_fq_count    one warp per 4 KiB tile: number of newlines of the tile
//   (prefix sum)  → global index of the first newline of every tile
//   k_fq_index    same tiles: every newline gets its global index; index = 0 (mod 4) writes the start of a read,
//                 1 (mod 4) its end, and the newline that closes the last complete record marks where the next
//                 segment of a long file continues
//   k_fq_lengths  read lengths (+ the longest) → prefix sum = read offsets
//   k_fq_gather   one warp per read: sequence bytes to the contiguous buffer the scan kernel streams
#pragma once
#include "bk_kernels.cuh"

namespace bk {

#define BK_FQ_TILE 4096u
#define BK_FQ_PIECES 8u            // 16-byte pieces per lane and tile: piece q of lane l covers bytes q*512 + l*16 ..

// 16 bytes → bit b set iff byte b is '\n'
__device__ __forceinline__ u32 fq_nl_mask(uint4 v) {
    const u32 w[4] = {v.x, v.y, v.z, v.w};
    u32 m = 0;
#pragma unroll
    for (u32 i = 0; i < 4; i++) {
        const u32 f = __vcmpeq4(w[i], 0x0A0A0A0Au) & 0x01010101u;
        m |= ((f * 0x10204080u) >> 28) << (4 * i);
    }
    return m;
}

__device__ __forceinline__ uint4 fq_load(const u8* txt, u32 n, u32 at) {       // 16 bytes at `at` (16-aligned), zero past n
    if (at + 16 <= n) return __ldg(reinterpret_cast<const uint4*>(txt + at));
    u32 w[4] = {0, 0, 0, 0};
    for (u32 i = 0; i < 16 && at + i < n; i++) w[i >> 2] |= (u32)txt[at + i] << (8 * (i & 3));
    return make_uint4(w[0], w[1], w[2], w[3]);
}

// grid-stride over tiles, one warp per tile
__global__ void __launch_bounds__(256) k_fq_count(const u8* __restrict__ txt, u32 n, u32 n_tiles, u32* __restrict__ tile_nl) {
    const u32 lane = threadIdx.x & 31;
    const u32 warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, n_warps = (gridDim.x * blockDim.x) >> 5;
    for (u32 t = warp; t < n_tiles; t += n_warps) {
        const u32 base = t * BK_FQ_TILE;
        u32 c = 0;
#pragma unroll
        for (u32 q = 0; q < BK_FQ_PIECES; q++) {
            const u32 at = base + q * 512u + lane * 16u;
            if (at < n) c += __popc(fq_nl_mask(fq_load(txt, n, at)));
        }
        c = warp_sum_u32(c);
        if (lane == 0) tile_nl[t] = c;
    }
}

// tile_first[t] = global index of the first newline of tile t.  n_emit reads are written (read r: bytes
// [seq_start[r], seq_end[r])); the byte after newline number `tail_after` (if it exists) is where the unconsumed tail of the
// segment starts.
__global__ void __launch_bounds__(256) k_fq_index(const u8* __restrict__ txt, u32 n, u32 n_tiles, const u32* __restrict__ tile_first,
                                                  u32 n_emit, u32* __restrict__ seq_start, u32* __restrict__ seq_end, u32 tail_after, u32* tail_start) {
    const u32 lane = threadIdx.x & 31;
    const u32 warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, n_warps = (gridDim.x * blockDim.x) >> 5;
    for (u32 t = warp; t < n_tiles; t += n_warps) {
        const u32 base = t * BK_FQ_TILE;
        u32 run = tile_first[t];
#pragma unroll 1
        for (u32 q = 0; q < BK_FQ_PIECES; q++) {
            const u32 at = base + q * 512u + lane * 16u;
            u32 m = at < n ? fq_nl_mask(fq_load(txt, n, at)) : 0u;
            const u32 c = __popc(m);
            u32 incl = c;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) { const u32 v = __shfl_up_sync(0xFFFFFFFFu, incl, o); if (lane >= (u32)o) incl += v; }
            u32 j = run + incl - c;
            run += __shfl_sync(0xFFFFFFFFu, incl, 31);
            while (m) {
                const u32 p = at + (u32)__ffs((int)m) - 1u;
                m &= m - 1;
                const u32 r = j >> 2, ph = j & 3u;
                if (ph == 0) { if (r < n_emit) seq_start[r] = p + 1; }
                else if (ph == 1) { if (r < n_emit) seq_end[r] = (p > 0 && txt[p - 1] == '\r') ? p - 1 : p; }
                if (j == tail_after) *tail_start = p + 1;
                j++;
            }
        }
    }
}

// len[r] (in place of a prefix-sum input), the longest read; a read whose start was never written (a file that begins
// with its sequence line cannot exist: line 0 is a header) has length 0
__global__ void __launch_bounds__(256) k_fq_lengths(const u32* __restrict__ seq_start, const u32* __restrict__ seq_end, u32 n_reads, u32* __restrict__ len, u32* max_len) {
    u32 mx = 0;
    for (u32 r = blockIdx.x * blockDim.x + threadIdx.x; r < n_reads; r += gridDim.x * blockDim.x) {
        const u32 s = seq_start[r], e = seq_end[r];
        const u32 l = e > s ? e - s : 0u;
        len[r] = l;
        mx = max(mx, l);
    }
    mx = __reduce_max_sync(0xFFFFFFFFu, mx);
    if ((threadIdx.x & 31) == 0 && mx) atomicMax(max_len, mx);
}

// off = exclusive prefix of the lengths (n_reads + 1 entries)
__global__ void __launch_bounds__(256) k_fq_gather(const u8* __restrict__ txt, const u32* __restrict__ seq_start, const u32* __restrict__ off, u32 n_reads, u8* __restrict__ bases) {
    const u32 lane = threadIdx.x & 31;
    const u32 warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, n_warps = (gridDim.x * blockDim.x) >> 5;
    for (u32 r = warp; r < n_reads; r += n_warps) {
        const u32 s = seq_start[r], o = off[r], l = off[r + 1] - o;
        for (u32 i = lane; i < l; i += 32) bases[o + i] = txt[s + i];
    }
}

// The byte behind the text: '\n' if the text does not end with one (a last line without '\n' counts),
// else a byte that starts a line nobody ends (an empty last line must NOT appear: a header that ends the file is not
// followed by an empty read).
__global__ void k_fq_terminate(u8* txt, u32 n) { txt[n] = (n == 0 || txt[n - 1] == '\n') ? (u8)'@' : (u8)'\n'; }

// 2-bit packed reads (bk_reads_push_packed) → the ASCII bytes the scan kernel streams: one thread per packed word
// (16 bases), one 16-byte store.  Base i of the push sits at bits 2 * (i % 16) of word i / 16, code A0 C1 G2 T3.
__global__ void __launch_bounds__(256) k_unpack2(const u32* __restrict__ packed, u64 n_words, u8* __restrict__ bases) {
    const u64 stride = (u64)gridDim.x * blockDim.x;
    for (u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x; i < n_words; i += stride) {
        const u32 w = __ldg(packed + i);
        u32 o[4];
#pragma unroll
        for (u32 j = 0; j < 4; j++) {
            const u32 x = (w >> (8 * j)) & 0xFFu;                                  // four codes
            const u32 sel = (x & 3u) | ((x & 0xCu) << 2) | ((x & 0x30u) << 4) | ((x & 0xC0u) << 6);
            o[j] = __byte_perm(0x54474341u, 0u, sel);                             // "ACGT"[code]
        }
        reinterpret_cast<uint4*>(bases)[i] = make_uint4(o[0], o[1], o[2], o[3]);
    }
}

// after a batch of engine inflates: every member must have produced the size its trailer announced
__global__ void __launch_bounds__(256) k_fq_check_sizes(const u32* __restrict__ got, const u32* __restrict__ want, u32 n, u32* bad) {
    for (u32 i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) if (got[i] != want[i]) atomicAdd(bad, 1u);
}

}  // namespace bk
