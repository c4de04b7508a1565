// bk_shard.cuh — kernels of the read-sharded deep sample (BASELINE config C3) that are not part of
// the single-rank path: packing the per-rank statistics for one SUM all-reduce, and the in-process transport (several
// ranks on ONE device, used by the single-GPU tests: NCCL refuses two ranks on one device, the kernels of the
// sharded path are the same under either transport).
//
// Why the path is not "all-reduce the pileups" (north_star) — the reference's depth is a MAX over k-mer counts that
// were summed over the whole file, thresholded (>= --min-kmers) and saturated (<= 10^6) first
// (src/call.rs:1172-1173, 1341-1345), and R1 / R2 are counted separately (302-317): counts are merged across ranks
// BEFORE the cut-offs, every k-mer is then mapped by exactly one owner rank, and only then are the depth arrays
// combined with MAX and the support arrays / tallies with SUM.
#pragma once
#include "bk_kernels.cuh"

namespace bk {

#define BK_SHARD_MAX_LOCAL 16
struct PtrList { void* p[BK_SHARD_MAX_LOCAL]; };

// in-process all-reduce: element i of every member's buffer becomes the reduction over the members
template <class T, int OP /*0 sum, 1 max*/>
__global__ void __launch_bounds__(256) k_local_allreduce(PtrList bufs, u32 n_members, u64 count) {
    const u64 stride = (u64)gridDim.x * blockDim.x;
    for (u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x; i < count; i += stride) {
        T acc = reinterpret_cast<const T*>(bufs.p[0])[i];
        for (u32 m = 1; m < n_members; m++) {
            const T v = reinterpret_cast<const T*>(bufs.p[m])[i];
            acc = OP == 0 ? (T)(acc + v) : (v > acc ? v : acc);
        }
        for (u32 m = 0; m < n_members; m++) reinterpret_cast<T*>(bufs.p[m])[i] = acc;
    }
}

// Per-rank statistics → one u64 vector (SUM all-reduced in place):
//   [file * stride + 0..3]  total_reads, total_kmers, unique_kmers, unique_counted of this rank's share
//   [file * stride + 4 + i] tallies of the file (genome-major: perfect, variant, unique-perfect, present)
// Exact because every distinct k-mer (reference id or novel) is owned — counted, thresholded, mapped — by one rank.
__global__ void __launch_bounds__(256) k_shard_pack(const Counters* c, const u32* g0, const u32* g1, u32 n_g4, u64 reads0, u64 reads1, u64* out) {
    const u32 stride = 4 + n_g4;
    for (u32 f = 0; f < 2; f++) {
        const u32* g = f ? g1 : g0;
        for (u32 i = threadIdx.x; i < stride; i += blockDim.x) {
            u64 v = 0;
            if (i == 0) v = f ? reads1 : reads0;
            else if (i == 1) v = c->f[f].total_kmers;
            else if (i == 2) v = c->f[f].unique;
            else if (i == 3) v = c->f[f].n_counted;
            else if (g) v = g[i - 4];
            out[f * stride + i] = v;
        }
    }
}
__global__ void __launch_bounds__(256) k_shard_unpack(const u64* in, u32* g0, u32* g1, u32 n_g4) {
    const u32 stride = 4 + n_g4;
    for (u32 f = 0; f < 2; f++) {
        u32* g = f ? g1 : g0;
        if (!g) continue;
        for (u32 i = threadIdx.x; i < n_g4; i += blockDim.x) {
            const u64 v = in[f * stride + 4 + i];
            g[i] = (i & 3) == 3 ? (v ? 1u : 0u) : (u32)v;
        }
    }
}

}  // namespace bk
