// bk_device.cu — bk_ctx, device memory, kernel launches and the C ABI of libbronko_b200.so
// (include/bronko_b200.h).  There is no CPU fallback anywhere in this file: every stage of a sample
// runs as a CUDA kernel from bk_kernels.cuh, and bk_create fails without an sm_100 device.
#include <cuda.h>
#include <dlfcn.h>
#include <nccl.h>

#include <algorithm>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <new>
#include <string>
#include <vector>

#include "bk_host.h"
#define BK_VARS_EAGER 1024
#include "bk_shard.cuh"
#include "bk_fastq.cuh"

using namespace bk;

static std::string g_create_error;

#define BK_CUDA(call)                                                                         \
    do {                                                                                      \
        cudaError_t e__ = (call);                                                             \
        if (e__ != cudaSuccess) return ctx->fail(BK_ERR_CUDA, "%s: %s", #call, cudaGetErrorString(e__)); \
    } while (0)

namespace {

template <class T>
struct DevBuf {
    T* p = nullptr; size_t cap = 0;
    cudaError_t reserve(size_t n) {
        if (n <= cap) return cudaSuccess;
        if (p) cudaFree(p);
        p = nullptr; cap = 0;
        cudaError_t e = cudaMalloc((void**)&p, std::max<size_t>(n, 1) * sizeof(T));
        if (e == cudaSuccess) cap = n;
        return e;
    }
    template <class V>
    cudaError_t upload(const V& v, cudaStream_t st) {
        cudaError_t e = reserve(v.size());
        if (e != cudaSuccess || v.empty()) return e;
        return cudaMemcpyAsync(p, v.data(), v.size() * sizeof(T), cudaMemcpyHostToDevice, st);
    }
    void release() { if (p) cudaFree(p); p = nullptr; cap = 0; }
};

enum Stage { ST_SCAN = 0, ST_LEFTOVER, ST_FINALIZE, ST_MAP, ST_SCORE, ST_COLL, ST_DECODE, ST_N };

// Everything one reads file (R1 or R2) owns.  The two files of a pair are independent until the genome is selected
// (src/call.rs:302-317 counts and maps them one after the other), so each has its own scratch and its own streams.
struct FileState {
    DevBuf<u32> diff, idcnt;
    DevBuf<u64> ckmers; DevBuf<u32> ccounts;   // the counted list: the "KMC dump" of the file
    DevBuf<u32> gstats;
    DevBuf<uint2> desc;                        // leftover stretches queued by the scan of one push
    DevBuf<u32> bsum;                          // prefix-sum scratch
    DevBuf<u64> nov, sorted;                   // novel k-mer occurrences of the file, and the same grouped by bin (bk_bins.cuh)
    DevBuf<u32> bin_cnt;
    DevBuf<u32> dense; DevBuf<u8> dflag;       // mismatch lines (bk_dense.cuh): all zero between samples
    bool dense_dirty = true;                   // not known to be all zero (fresh allocation, or a sample that did not finish)
    // read-sharded mode: pair counts next to nov / sorted, per-bin pair counts, pairs received from the other ranks
    DevBuf<u32> wa, wb, dcount, own_off;
    DevBuf<u64> rk, rsk; DevBuf<u32> rc, rsw;
    size_t ck_cap = 0;                         // capacity the counted list of this sample was given (its back end holds the mismatch-line k-mers)
    u64 nov_ub = 0;                            // upper bound of list entries the pushes so far were given room for
    u64 nov_limit = 0;                         // entries the kernels may use (the allocation can be larger: it is reused)
    u32 bin_log2p = 8;
    bool used = false, folded = false, finalized = false;
    u64 total_reads = 0, total_bases = 0;
    void release() {
        diff.release(); idcnt.release(); ckmers.release(); ccounts.release(); gstats.release(); desc.release(); bsum.release();
        nov.release(); sorted.release(); bin_cnt.release(); dense.release(); dflag.release(); wa.release(); wb.release(); dcount.release(); own_off.release();
        rk.release(); rsk.release(); rc.release(); rsw.release();
    }
};

// The index on one device: host form, derived tables and their device copies.  Read-only once uploaded, so the
// contexts of one GPU (one per sample in flight) share a single copy (bk_index_share): the bucket table stays
// L2-resident for all of them instead of once per context.
struct IndexDev {
    int device = 0;
    HostIndex ix;
    DerivedIndex d;
    DevBuf<BucketSlotD> d_bucket_slots; DevBuf<BucketEntryD> d_bucket_entries;
    DevBuf<BucketSlotD> d_group_slots, d_group_centers; DevBuf<uint2> d_group_buckets;
    DevBuf<u32> d_refnib; DevBuf<u32> d_oseq_start, d_oseq_len;
    DevBuf<ExactSlotD> d_exact;
    DevBuf<u32> d_slot2id; DevBuf<u64> d_id_kmer;
    DevBuf<u32> d_slot2rep, d_id_amb, d_id_rep; DevBuf<ExactSlotD> d_nb;       // mismatch lines (bk_dense.cuh); empty when !d.dense_ok
    DevBuf<u32> d_line_amb, d_line_fold; DevBuf<u64> d_nb_bloom;
    DevBuf<uint2> d_id_bucket;                                                   // map shortcut (bk_host.h); empty when !d.map_shortcut_ok
    DevBuf<u32> d_genome_row0, d_genome_seq_off, d_seq_row0; DevBuf<u64> d_genome_len; DevBuf<u8> d_ref_code;
    u32 max_seqs_per_genome = 1;
    ~IndexDev() {
        cudaSetDevice(device);
        d_bucket_slots.release(); d_bucket_entries.release(); d_group_slots.release(); d_group_centers.release(); d_group_buckets.release(); d_refnib.release(); d_oseq_start.release(); d_oseq_len.release();
        d_exact.release(); d_slot2id.release(); d_id_kmer.release(); d_genome_row0.release(); d_genome_seq_off.release();
        d_seq_row0.release(); d_genome_len.release(); d_ref_code.release();
        d_slot2rep.release(); d_id_amb.release(); d_id_rep.release(); d_nb.release(); d_id_bucket.release(); d_line_amb.release(); d_line_fold.release(); d_nb_bloom.release();
    }
};

// The ranks of one read-sharded sample (bk_shard_* in include/bronko_b200.h).  NCCL: one rank per process and GPU,
// `members` holds this process's context only.  Local: all ranks are contexts of this process on one device (tests).
struct ShardGroup {
    u32 n = 1;
    bool local = false;
    std::vector<bk_ctx*> members;
    void* comm = nullptr;                   // ncclComm_t
    ~ShardGroup();
};

}  // namespace

struct FqState;        // bk_fastq.inc: buffers of the FASTQ decode stage of one file slot
extern "C" {
static void fq_destroy(FqState* q);
static void fq_begin_sample(FqState* q);
}

struct bk_ctx {
    int device = 0;
    // Streams.  The two files of a sample run concurrently: file f counts (scan / leftover) on s_count[f] and is
    // finalized and mapped on s_fin[f]; selection + scoring run on s_score.  The three levels have rising CUDA
    // priority.  Every big kernel fills the SMs' register files, so with several samples in flight (one context each) a
    // kernel's CTAs only start as CTAs of other kernels retire, and without priorities the handful of CTAs of a sample's
    // last, latency-bound stages (the noise chains) queue behind thousands of pending CTAs of other samples' first
    // stages.  BK_PRIO=0 puts everything on one stream.
    cudaStream_t s_count[2] = {nullptr, nullptr}, s_fin[2] = {nullptr, nullptr}, s_score = nullptr, copy_stream = nullptr;
    std::vector<cudaStream_t> owned_streams;
    cudaEvent_t ev_chain = nullptr;
    int prio_mode = 2;
    // order `dst` after everything issued to `src` so far (one scratch event: a wait captures the record before it)
    void chain(cudaStream_t src, cudaStream_t dst) {
        if (src == dst) return;
        cudaEventRecord(ev_chain, src);
        cudaStreamWaitEvent(dst, ev_chain, 0);
    }
    std::string err;
    int sm_count = 148;
    bool spin_wait = false;

    std::shared_ptr<IndexDev> I;            // null until an index is loaded / built / shared

    // per-sample state
    bk_params params;
    bool in_sample = false, finished = false;
    FileState file[2];
    DevBuf<Counters> d_ctr;
    Counters* h_ctr = nullptr;              // pinned
    // Everything a sample copies back sits in PINNED memory: a cudaMemcpyAsync into pageable memory makes the calling
    // thread wait for the stream inside the driver, spinning — a core per context for as long as its sample runs.
    u32* h_gstats = nullptr; size_t h_gstats_cap = 0;            // 2 files x genomes x 4
    bk_variant* h_vars = nullptr;                                // the first BK_VARS_EAGER variants travel with the counters
    u32* h_sizes = nullptr; size_t h_sizes_cap = 0;              // read-sharded sample: the all-gathered pair counts
    cudaEvent_t ev_wait = nullptr;                               // host_wait
    int pinned_u32(u32** p, size_t* cap, size_t n) {
        if (n <= *cap) return BK_OK;
        if (*p) cudaFreeHost(*p);
        *p = nullptr; *cap = 0;
        if (cudaMallocHost((void**)p, n * sizeof(u32)) != cudaSuccess) return BK_ERR_NOMEM;
        *cap = n;
        return BK_OK;
    }
    // the host thread sleeps until the stream has drained (BK_SPIN=1: spins, ~50 us sooner)
    cudaError_t host_wait(cudaStream_t st) {
        cudaError_t e = cudaEventRecord(ev_wait, st);
        return e != cudaSuccess ? e : cudaEventSynchronize(ev_wait);
    }
    u64* h_stats = nullptr;                 // pinned: globally reduced KMC numbers of a sharded sample
    DevBuf<u32> d_pile;                     // 4 arrays x max_genome_rows x 4
    DevBuf<u32> d_pile_all;                 // one such block per genome (databases of at most four genomes: one-pass map)
    DevBuf<double> d_noise;                 // Noise.max per row
    DevBuf<double> d_nz_maf, d_nz_s, d_nz_s2, d_nz_tab, d_nz_warm;   // noise scratch (bk_noise.cuh: NoiseView)
    DevBuf<u8> d_nz_nzflag; DevBuf<u32> d_nz_list, d_nz_rank, d_nz_nact;
    DevBuf<u8> d_nz_flag; DevBuf<u32> d_nz_stats;
    u32 nz_max_chunks = 1;
    DevBuf<bk_variant> d_vars;
    // staging for host pushes (double buffered)
    DevBuf<u8> d_stage[2]; DevBuf<u32> d_stage_off;
    cudaEvent_t stage_free[2] = {nullptr, nullptr}, stage_copied[2] = {nullptr, nullptr};
    cudaEvent_t off_free = nullptr, off_copied = nullptr;   // d_stage_off: last kernels that read it / its H2D copy
    int stage_next = 0;
    FqState* fq[2] = {nullptr, nullptr};    // FASTQ decode stage (bk_fastq.inc), created on first use
    // read-sharded deep sample
    std::shared_ptr<ShardGroup> shard;      // null: this context holds whole samples
    u32 shard_rank = 0;
    DevBuf<u64> d_shard_stats;
    DevBuf<u32> d_shard_sizes;
    bool noise_debug = false;
    bool force_warp_map = false;            // tests: exercise the many-genome map kernel on a small db
    bool no_fused_map = false;              // tests: BK_NO_FUSED_MAP keeps the two-pass map on small databases
    bool no_group_map = false;              // tests: BK_NO_GROUP_MAP probes the per-bucket table instead of the grouped one

    // results
    bk_sample_result result;
    std::vector<bk_variant> variants;
    std::vector<bk_genome_stats> gstats[2];

    // timing
    struct Span { cudaEvent_t a, b; int stage; cudaStream_t st; };
    std::vector<Span> spans; size_t spans_used = 0;
    cudaEvent_t ev_begin = nullptr, ev_end = nullptr;
    u32 launches = 0, scan_launches = 0, coll_calls = 0;
    bk_stage_times times;

    u32 shard_n() const { return shard ? shard->n : 1u; }
    int fail(int code, const char* fmt, ...) {
        char buf[1024];
        va_list ap; va_start(ap, fmt); vsnprintf(buf, sizeof buf, fmt, ap); va_end(ap);
        err = buf;
        return code;
    }
    bool stage_timing = true;               // bk_stage_timing
    int span_begin(int stage, cudaStream_t st) {
        if (!stage_timing) return -1;
        if (spans_used == spans.size()) {
            Span s; s.stage = stage; s.st = st;
            if (cudaEventCreate(&s.a) != cudaSuccess || cudaEventCreate(&s.b) != cudaSuccess) return -1;
            spans.push_back(s);
        }
        spans[spans_used].stage = stage; spans[spans_used].st = st;
        cudaEventRecord(spans[spans_used].a, st);
        return (int)spans_used++;
    }
    void span_end(int id) { if (id >= 0) cudaEventRecord(spans[id].b, spans[id].st); }
};

static int grid_for(const bk_ctx* ctx, u64 items, u32 per_block, u32 max_waves = 8) {
    u64 g = (items + per_block - 1) / per_block;
    const u64 cap = (u64)ctx->sm_count * max_waves;
    if (g > cap) g = cap;
    if (g < 1) g = 1;
    return (int)g;
}

// Nothing throws across the ABI: allocation failures and anything else unexpected become a status.
template <class F>
static int guarded(bk_ctx* ctx, F&& f) {
    try { return f(); }
    catch (const std::bad_alloc&) { return ctx ? ctx->fail(BK_ERR_NOMEM, "out of host memory") : (int)BK_ERR_NOMEM; }
    catch (const std::exception& e) { return ctx ? ctx->fail(BK_ERR_ARG, "internal error: %s", e.what()) : (int)BK_ERR_ARG; }
    catch (...) { return ctx ? ctx->fail(BK_ERR_ARG, "internal error") : (int)BK_ERR_ARG; }
}

extern "C" {

const char* bk_version(void) { return "bronko_b200 0.2.0 (sm_100a)"; }

int bk_create(bk_ctx** out, int device) {
    if (!out) return BK_ERR_ARG;
    *out = nullptr;
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n <= 0) {
        g_create_error = std::string("no CUDA device: ") + (e != cudaSuccess ? cudaGetErrorString(e) : "device count is 0") +
                         " (libbronko_b200 has no CPU path)";
        return BK_ERR_NO_DEVICE;
    }
    if (device < 0 || device >= n) { g_create_error = "device index out of range"; return BK_ERR_ARG; }
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, device) != cudaSuccess || prop.major != 10) {
        g_create_error = "device is not sm_100 (B200); the kernels are built for sm_100a only";
        return BK_ERR_NO_DEVICE;
    }
    if (cudaSetDevice(device) != cudaSuccess) { g_create_error = "cudaSetDevice failed"; return BK_ERR_CUDA; }
    bk_ctx* ctx = new (std::nothrow) bk_ctx();
    if (!ctx) { g_create_error = "out of host memory"; return BK_ERR_NOMEM; }
    ctx->device = device;
    ctx->sm_count = prop.multiProcessorCount;
    int prio_least = 0, prio_greatest = 0;
    cudaDeviceGetStreamPriorityRange(&prio_least, &prio_greatest);              // numerically lower = higher priority
    if (const char* e = getenv("BK_PRIO")) ctx->prio_mode = atoi(e);
    if (prio_greatest >= prio_least) ctx->prio_mode = 0;
    ctx->spin_wait = getenv("BK_SPIN") != nullptr;
    auto mk = [&](cudaStream_t* s, int prio) {
        if (cudaStreamCreateWithPriority(s, cudaStreamNonBlocking, prio) != cudaSuccess) return false;
        ctx->owned_streams.push_back(*s);
        return true;
    };
    bool ok = mk(&ctx->s_count[0], prio_least) && mk(&ctx->copy_stream, prio_least) &&
              cudaEventCreateWithFlags(&ctx->ev_chain, cudaEventDisableTiming) == cudaSuccess &&
              cudaMallocHost((void**)&ctx->h_ctr, sizeof(Counters)) == cudaSuccess &&
              cudaMallocHost((void**)&ctx->h_vars, BK_VARS_EAGER * sizeof(bk_variant)) == cudaSuccess &&
              cudaEventCreateWithFlags(&ctx->ev_wait, (ctx->spin_wait ? 0u : (unsigned)cudaEventBlockingSync) | cudaEventDisableTiming) == cudaSuccess &&
              cudaMallocHost((void**)&ctx->h_stats, 8 * sizeof(u64)) == cudaSuccess &&
              cudaEventCreate(&ctx->ev_begin) == cudaSuccess &&
              cudaEventCreateWithFlags(&ctx->ev_end, ctx->spin_wait ? cudaEventDefault : cudaEventBlockingSync) == cudaSuccess;
    if (ok && ctx->prio_mode != 0) {
        const int mid = std::max(prio_greatest, prio_least - 1);
        ok = mk(&ctx->s_count[1], prio_least) && mk(&ctx->s_fin[0], mid) && mk(&ctx->s_fin[1], mid) && mk(&ctx->s_score, prio_greatest);
    } else if (ok) {
        ctx->s_count[1] = ctx->s_fin[0] = ctx->s_fin[1] = ctx->s_score = ctx->s_count[0];
    }
    for (int i = 0; i < 2 && ok; i++)
        ok = cudaEventCreateWithFlags(&ctx->stage_free[i], cudaEventDisableTiming) == cudaSuccess &&
             cudaEventCreateWithFlags(&ctx->stage_copied[i], cudaEventDisableTiming) == cudaSuccess;
    if (ok) ok = cudaEventCreateWithFlags(&ctx->off_free, cudaEventDisableTiming) == cudaSuccess &&
                 cudaEventCreateWithFlags(&ctx->off_copied, cudaEventDisableTiming) == cudaSuccess;
    if (ok) {
        double tau[301];
        tau_table(tau);
        ok = cudaMemcpyToSymbol(c_tau, tau, sizeof tau) == cudaSuccess;
    }
    if (ok) ok = cudaFuncSetAttribute(k_scan, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024) == cudaSuccess &&
                 cudaFuncSetAttribute(k_map<0, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024) == cudaSuccess &&
                 cudaFuncSetAttribute(k_map<0, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024) == cudaSuccess &&
                 cudaFuncSetAttribute(k_noise_seq, cudaFuncAttributeMaxDynamicSharedMemorySize, BK_NZ_SEQ_SMEM) == cudaSuccess &&
                 cudaFuncSetAttribute(k_bin_hist, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024) == cudaSuccess &&
                 cudaFuncSetAttribute(k_bin_scatter<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024) == cudaSuccess &&
                 cudaFuncSetAttribute(k_bin_scatter<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024) == cudaSuccess &&
                 cudaFuncSetAttribute(k_bin_count<0, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, BK_BIN_SMEM) == cudaSuccess &&
                 cudaFuncSetAttribute(k_bin_count<0, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, BK_BIN_SMEM) == cudaSuccess &&
                 cudaFuncSetAttribute(k_bin_count<1, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, BK_BIN_SMEM) == cudaSuccess &&
                 cudaFuncSetAttribute(k_bin_count<1, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, BK_BIN_SMEM) == cudaSuccess &&
                 cudaFuncSetAttribute(k_bin_count<2, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, BK_BIN_SMEM) == cudaSuccess;
    if (!ok) { g_create_error = std::string("context setup failed: ") + cudaGetErrorString(cudaGetLastError()); delete ctx; return BK_ERR_CUDA; }
    memset(&ctx->times, 0, sizeof ctx->times);
    memset(&ctx->result, 0, sizeof ctx->result);
    bk_params_default(&ctx->params);
    ctx->force_warp_map = getenv("BK_FORCE_WARP_MAP") != nullptr;
    ctx->noise_debug = getenv("BK_NOISE_DEBUG") != nullptr;
    ctx->no_fused_map = getenv("BK_NO_FUSED_MAP") != nullptr;
    ctx->no_group_map = getenv("BK_NO_GROUP_MAP") != nullptr;
    *out = ctx;
    return BK_OK;
}

void bk_destroy(bk_ctx* ctx) {
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    cudaDeviceSynchronize();
    if (ctx->shard) {                       // leave the group: the other members must not be driven through it any more
        for (bk_ctx*& m : ctx->shard->members) if (m == ctx) m = nullptr;
        ctx->shard.reset();
    }
    ctx->I.reset();                         // the last context sharing an index frees its device copies
    for (FileState& f : ctx->file) f.release();
    for (FqState*& q : ctx->fq) { fq_destroy(q); q = nullptr; }
    ctx->d_ctr.release(); ctx->d_pile.release(); ctx->d_pile_all.release();
    ctx->d_noise.release(); ctx->d_vars.release();
    ctx->d_nz_maf.release(); ctx->d_nz_s.release(); ctx->d_nz_s2.release(); ctx->d_nz_tab.release(); ctx->d_nz_warm.release();
    ctx->d_nz_nzflag.release(); ctx->d_nz_list.release(); ctx->d_nz_rank.release(); ctx->d_nz_nact.release();
    ctx->d_nz_flag.release(); ctx->d_nz_stats.release();
    ctx->d_stage[0].release(); ctx->d_stage[1].release(); ctx->d_stage_off.release();
    ctx->d_shard_stats.release(); ctx->d_shard_sizes.release();
    for (auto& s : ctx->spans) { cudaEventDestroy(s.a); cudaEventDestroy(s.b); }
    for (int i = 0; i < 2; i++) { if (ctx->stage_free[i]) cudaEventDestroy(ctx->stage_free[i]); if (ctx->stage_copied[i]) cudaEventDestroy(ctx->stage_copied[i]); }
    if (ctx->off_free) cudaEventDestroy(ctx->off_free);
    if (ctx->off_copied) cudaEventDestroy(ctx->off_copied);
    if (ctx->ev_begin) cudaEventDestroy(ctx->ev_begin);
    if (ctx->ev_end) cudaEventDestroy(ctx->ev_end);
    if (ctx->h_ctr) cudaFreeHost(ctx->h_ctr);
    if (ctx->h_vars) cudaFreeHost(ctx->h_vars);
    if (ctx->h_gstats) cudaFreeHost(ctx->h_gstats);
    if (ctx->h_sizes) cudaFreeHost(ctx->h_sizes);
    if (ctx->ev_wait) cudaEventDestroy(ctx->ev_wait);
    if (ctx->h_stats) cudaFreeHost(ctx->h_stats);
    for (cudaStream_t st : ctx->owned_streams) cudaStreamDestroy(st);
    if (ctx->ev_chain) cudaEventDestroy(ctx->ev_chain);
    delete ctx;
}

const char* bk_last_error(bk_ctx* ctx) { return ctx ? ctx->err.c_str() : g_create_error.c_str(); }
void* bk_stream(bk_ctx* ctx) { return ctx ? (void*)ctx->s_count[0] : nullptr; }
void* bk_stream_slot(bk_ctx* ctx, int file_slot) { return (ctx && file_slot >= 0 && file_slot < 2) ? (void*)ctx->s_count[file_slot] : nullptr; }

void* bk_host_alloc(uint64_t bytes) { void* p = nullptr; return cudaMallocHost(&p, bytes ? bytes : 1) == cudaSuccess ? p : nullptr; }
void bk_host_free(void* p) { if (p) cudaFreeHost(p); }

void bk_params_default(bk_params* p) {    // reference src/consts.rs:1-20, src/call.rs:1173
    memset(p, 0, sizeof *p);
    p->k = 21; p->min_kmers = 3; p->counter_max = 1000000; p->use_full_kmer = 0; p->n_fixed = 2;
    p->n_per_strand = 2; p->table_log2 = 0; p->min_depth = 300; p->min_variant_depth = 3;
    p->min_af = 0.03; p->strand_balance_ratio = 0.1; p->strand_odds_max = 6.0; p->variant_multiplier = 1.5;
}

uint64_t bk_clean_sample_id(const char* path, char* buf, uint64_t cap) {
    const std::string t = clean_sample_id(path ? path : "");
    if (buf && cap) { const u64 n = std::min<u64>(cap - 1, t.size()); memcpy(buf, t.data(), n); buf[n] = 0; }
    return t.size() + 1;
}

// ---------------------------------------------------------------------------------------------
// index
// ---------------------------------------------------------------------------------------------
static int size_for_index(bk_ctx* ctx);

// device copies of a freshly filled IndexDev (I->ix); on success it becomes the context's index
static int upload_index(bk_ctx* ctx, std::shared_ptr<IndexDev> fresh) {
    cudaSetDevice(ctx->device);
    fresh->device = ctx->device;
    if (fresh->ix.k < 15 || fresh->ix.k > 31 || (fresh->ix.k & 1) == 0)
        return ctx->fail(BK_ERR_ARG, "Invalid kmer size, must be odd and between [15-31]");
    derive_index(fresh->ix, fresh->d, getenv("BK_NO_REKEY") == nullptr);
    ctx->I = fresh;
    const DerivedIndex& d = ctx->I->d;
    if (d.n_genomes == 0 || d.n_genomes > 4096) return ctx->fail(BK_ERR_ARG, "index holds %u genomes (supported: 1..4096)", d.n_genomes);
    if ((u64)d.n_raw + 2 >= 0x7FFFFFFFull) return ctx->fail(BK_ERR_ARG, "reference set too large for 32-bit slot indices");
    cudaStream_t st = ctx->s_count[0];
    static_assert(sizeof(BucketSlot) == sizeof(BucketSlotD) && sizeof(BucketEntry) == sizeof(BucketEntryD) && sizeof(ExactSlot) == sizeof(ExactSlotD), "layout");
    BK_CUDA(ctx->I->d_bucket_slots.reserve(d.bucket_slots.size()));
    BK_CUDA(cudaMemcpyAsync(ctx->I->d_bucket_slots.p, d.bucket_slots.data(), d.bucket_slots.size() * 16, cudaMemcpyHostToDevice, st));
    BK_CUDA(ctx->I->d_bucket_entries.reserve(d.bucket_entries.size()));
    if (!d.bucket_entries.empty())
        BK_CUDA(cudaMemcpyAsync(ctx->I->d_bucket_entries.p, d.bucket_entries.data(), d.bucket_entries.size() * 8, cudaMemcpyHostToDevice, st));
    if (!d.group_slots.empty()) {
        static_assert(sizeof(OffLen) == sizeof(uint2), "layout");
        BK_CUDA(ctx->I->d_group_slots.reserve(d.group_slots.size()));
        BK_CUDA(cudaMemcpyAsync(ctx->I->d_group_slots.p, d.group_slots.data(), d.group_slots.size() * 16, cudaMemcpyHostToDevice, st));
        BK_CUDA(ctx->I->d_group_centers.reserve(std::max<size_t>(d.group_centers.size(), 1)));
        if (!d.group_centers.empty())
            BK_CUDA(cudaMemcpyAsync(ctx->I->d_group_centers.p, d.group_centers.data(), d.group_centers.size() * 16, cudaMemcpyHostToDevice, st));
        BK_CUDA(ctx->I->d_group_buckets.reserve(d.group_buckets.size() + 4));     // (+4: batched loads may run past a side's last bucket)
        if (!d.group_buckets.empty())
            BK_CUDA(cudaMemcpyAsync(ctx->I->d_group_buckets.p, d.group_buckets.data(), d.group_buckets.size() * 8, cudaMemcpyHostToDevice, st));
    }
    BK_CUDA(ctx->I->d_exact.reserve(d.exact_slots.size()));
    BK_CUDA(cudaMemcpyAsync(ctx->I->d_exact.p, d.exact_slots.data(), d.exact_slots.size() * 16, cudaMemcpyHostToDevice, st));
    BK_CUDA(ctx->I->d_refnib.upload(d.refnib, st));
    BK_CUDA(ctx->I->d_oseq_start.upload(d.oseq_start, st));
    BK_CUDA(ctx->I->d_oseq_len.upload(d.oseq_len, st));
    BK_CUDA(ctx->I->d_slot2id.upload(d.slot2id, st));
    BK_CUDA(ctx->I->d_id_kmer.upload(d.id_kmer, st));
    if (d.dense_ok) {
        BK_CUDA(ctx->I->d_slot2rep.upload(d.slot2rep, st));
        BK_CUDA(ctx->I->d_id_amb.upload(d.id_amb, st));
        BK_CUDA(ctx->I->d_id_rep.upload(d.id_rep, st));
        BK_CUDA(ctx->I->d_line_amb.upload(d.line_amb, st));
        BK_CUDA(ctx->I->d_line_fold.upload(d.line_fold, st));
        BK_CUDA(ctx->I->d_nb_bloom.upload(d.nb_bloom, st));
        BK_CUDA(ctx->I->d_nb.reserve(d.nb_slots.size()));
        BK_CUDA(cudaMemcpyAsync(ctx->I->d_nb.p, d.nb_slots.data(), d.nb_slots.size() * 16, cudaMemcpyHostToDevice, st));
        if (d.map_shortcut_ok) {
            static_assert(sizeof(OffLen) == sizeof(uint2), "layout");
            BK_CUDA(ctx->I->d_id_bucket.reserve(d.id_bucket.size()));
            BK_CUDA(cudaMemcpyAsync(ctx->I->d_id_bucket.p, d.id_bucket.data(), d.id_bucket.size() * 8, cudaMemcpyHostToDevice, st));
        }
    }
    BK_CUDA(ctx->I->d_genome_row0.upload(d.genome_row0, st));
    BK_CUDA(ctx->I->d_genome_seq_off.upload(d.genome_seq_off, st));
    BK_CUDA(ctx->I->d_seq_row0.upload(d.seq_row0, st));
    BK_CUDA(ctx->I->d_genome_len.upload(d.genome_len, st));
    BK_CUDA(ctx->I->d_ref_code.upload(d.ref_code, st));
    ctx->I->max_seqs_per_genome = 1;
    for (u32 g = 0; g < d.n_genomes; g++) ctx->I->max_seqs_per_genome = std::max(ctx->I->max_seqs_per_genome, d.genome_seq_off[g + 1] - d.genome_seq_off[g]);
    BK_CUDA(cudaStreamSynchronize(st));
    return size_for_index(ctx);
}

// per-context buffers whose size follows the index (pileups, noise scratch, per-file counters)
static int size_for_index(bk_ctx* ctx) {
    cudaSetDevice(ctx->device);
    const DerivedIndex& d = ctx->I->d;
    const size_t rows = std::max<u32>(d.max_genome_rows, 1);
    BK_CUDA(ctx->d_pile.reserve(rows * 16));
    if (d.n_genomes <= 4) BK_CUDA(ctx->d_pile_all.reserve(rows * 16 * d.n_genomes));
    BK_CUDA(ctx->d_noise.reserve(rows));
    {   // noise scratch: fractions with padding per sequence, per-iteration snapshots, chunk slots
        const size_t seqs = ctx->I->max_seqs_per_genome;
        const size_t it_slots = rows + BK_NOISE_HALF * seqs;
        const size_t chunk_slots = it_slots / BK_NZ_CHUNK + seqs + 2;
        u32 max_len = 0;
        for (size_t q = 0; q + 1 < d.seq_row0.size(); q++) max_len = std::max(max_len, d.seq_row0[q + 1] - d.seq_row0[q]);
        ctx->nz_max_chunks = (max_len + BK_NOISE_HALF + BK_NZ_CHUNK - 1) / BK_NZ_CHUNK;
        BK_CUDA(ctx->d_nz_maf.reserve((rows + BK_NZ_PAD * seqs) * 3));
        BK_CUDA(ctx->d_nz_s.reserve(it_slots)); BK_CUDA(ctx->d_nz_s2.reserve(it_slots));
        BK_CUDA(ctx->d_nz_nzflag.reserve(it_slots)); BK_CUDA(ctx->d_nz_list.reserve(it_slots)); BK_CUDA(ctx->d_nz_rank.reserve(it_slots));
        BK_CUDA(ctx->d_nz_nact.reserve(seqs));
        BK_CUDA(ctx->d_nz_tab.reserve(it_slots * BK_NOISE_TABLE));
        BK_CUDA(ctx->d_nz_warm.reserve(chunk_slots * BK_NOISE_TABLE));
        BK_CUDA(ctx->d_nz_flag.reserve(chunk_slots));
        BK_CUDA(ctx->d_nz_stats.reserve(16));
    }
    BK_CUDA(ctx->d_vars.reserve(rows * 3));
    BK_CUDA(ctx->d_ctr.reserve(1));
    for (FileState& f : ctx->file) {
        BK_CUDA(f.diff.reserve((size_t)d.n_raw + 2));
        BK_CUDA(f.idcnt.reserve(d.id_kmer.size()));
        BK_CUDA(f.gstats.reserve((size_t)d.n_genomes * 4));
        BK_CUDA(f.bsum.reserve(std::max<size_t>((size_t)d.n_raw + 2, (size_t)16384 * ctx->sm_count * BK_BIN_G_PER_SM) / BK_PS_BLOCK + 2));
        if (d.dense_ok) {
            BK_CUDA(f.dense.reserve((size_t)d.n_raw * 4 * (d.k + 1)));
            BK_CUDA(f.dflag.reserve((size_t)d.n_raw * 4));
        }
        f.dense_dirty = true;
    }
    ctx->in_sample = false; ctx->finished = false;
    return BK_OK;
}

int bk_index_share(bk_ctx* ctx, bk_ctx* owner) {
    if (!ctx || !owner) return BK_ERR_ARG;
    return guarded(ctx, [&]() -> int {
    if (!owner->I) return ctx->fail(BK_ERR_ARG, "bk_index_share: the other context holds no index");
    if (owner->device != ctx->device) return ctx->fail(BK_ERR_ARG, "bk_index_share: contexts live on different devices");
    if (ctx->in_sample && !ctx->finished) return ctx->fail(BK_ERR_ARG, "bk_index_share: call between samples");
    ctx->I = owner->I;
    return size_for_index(ctx);
    });
}

int bk_index_load(bk_ctx* ctx, uint32_t k, uint64_t n_keys, const uint64_t* keys, const uint64_t* entry_off,
                  const bk_bucket_info* entries, uint32_t n_genomes, const uint32_t* genome_seq_off,
                  const uint64_t* seq_len, const uint64_t* seq_base_off, const uint8_t* ref_bases) {
    if (!ctx) return BK_ERR_ARG;
    if (!keys || !entry_off || !entries || !genome_seq_off || !seq_len || !seq_base_off || !ref_bases)
        return ctx->fail(BK_ERR_ARG, "bk_index_load: null argument");
    return guarded(ctx, [&]() -> int {
    auto fresh = std::make_shared<IndexDev>();
    HostIndex& ix = fresh->ix;
    ix.k = k; ix.meta_k = k;
    std::vector<KeyedEntry> pairs;
    pairs.reserve(entry_off[n_keys]);
    for (u64 i = 0; i < n_keys; i++)
        for (u64 j = entry_off[i]; j < entry_off[i + 1]; j++) pairs.push_back(KeyedEntry{keys[i], entries[j]});
    index_from_pairs(ix, pairs);
    for (u32 g = 0; g < n_genomes; g++) {
        HostGenome hg;
        hg.name = "genome" + std::to_string(g);
        for (u32 s = genome_seq_off[g]; s < genome_seq_off[g + 1]; s++) {
            HostSeq q;
            q.name = "seq" + std::to_string(s - genome_seq_off[g]);
            q.len = seq_len[s];
            q.bases.assign(ref_bases + seq_base_off[s], ref_bases + seq_base_off[s] + seq_len[s]);
            hg.seqs.push_back(std::move(q));
        }
        ix.genomes.push_back(std::move(hg));
    }
    return upload_index(ctx, fresh);
    });
}

int bk_index_load_file(bk_ctx* ctx, const char* path) {
    if (!ctx || !path) return BK_ERR_ARG;
    return guarded(ctx, [&]() -> int {
    std::string err;
    auto fresh = std::make_shared<IndexDev>();
    if (!bkdb_read(path, fresh->ix, err)) return ctx->fail(BK_ERR_IO, "%s", err.c_str());
    return upload_index(ctx, fresh);
    });
}

int bk_index_build(bk_ctx* ctx, uint32_t k, uint32_t n_files, const char* const* fasta_paths) {
    if (!ctx || !fasta_paths || n_files == 0) return BK_ERR_ARG;
    if (k < 15 || k > 31 || (k & 1) == 0) return ctx->fail(BK_ERR_ARG, "Invalid kmer size, must be odd and between [15-31]");
    return guarded(ctx, [&]() -> int {
    std::vector<std::string> paths(fasta_paths, fasta_paths + n_files);
    std::string err;
    auto fresh = std::make_shared<IndexDev>();
    if (!index_build_from_fasta(k, paths, fresh->ix, err)) return ctx->fail(BK_ERR_IO, "%s", err.c_str());
    return upload_index(ctx, fresh);
    });
}

int bk_index_save(bk_ctx* ctx, const char* path) {
    if (!ctx || !path || !(ctx->I != nullptr)) return BK_ERR_ARG;
    return guarded(ctx, [&]() -> int {
    std::string err;
    if (!bkdb_write(path, ctx->I->ix, err)) return ctx->fail(BK_ERR_IO, "%s", err.c_str());
    return BK_OK;
    });
}

int bk_index_info(bk_ctx* ctx, uint32_t* k, uint64_t* n_keys, uint64_t* n_entries, uint32_t* n_genomes) {
    if (!ctx || !(ctx->I != nullptr)) return BK_ERR_ARG;
    if (k) *k = ctx->I->ix.k;
    if (n_keys) *n_keys = ctx->I->ix.keys.size();
    if (n_entries) *n_entries = ctx->I->ix.entries.size();
    if (n_genomes) *n_genomes = (u32)ctx->I->ix.genomes.size();
    return BK_OK;
}
const char* bk_genome_name(bk_ctx* ctx, uint32_t g) { return (ctx && ctx->I && g < ctx->I->ix.genomes.size()) ? ctx->I->ix.genomes[g].name.c_str() : nullptr; }
uint32_t bk_genome_n_seqs(bk_ctx* ctx, uint32_t g) { return (ctx && ctx->I && g < ctx->I->ix.genomes.size()) ? (u32)ctx->I->ix.genomes[g].seqs.size() : 0; }
const char* bk_seq_name(bk_ctx* ctx, uint32_t g, uint32_t s) {
    return (ctx && ctx->I && g < ctx->I->ix.genomes.size() && s < ctx->I->ix.genomes[g].seqs.size()) ? ctx->I->ix.genomes[g].seqs[s].name.c_str() : nullptr;
}
uint64_t bk_seq_len(bk_ctx* ctx, uint32_t g, uint32_t s) {
    return (ctx && ctx->I && g < ctx->I->ix.genomes.size() && s < ctx->I->ix.genomes[g].seqs.size()) ? ctx->I->ix.genomes[g].seqs[s].len : 0;
}
const uint8_t* bk_seq_bases(bk_ctx* ctx, uint32_t g, uint32_t s) {
    return (ctx && ctx->I && g < ctx->I->ix.genomes.size() && s < ctx->I->ix.genomes[g].seqs.size()) ? ctx->I->ix.genomes[g].seqs[s].bases.data() : nullptr;
}
int bk_index_export(bk_ctx* ctx, uint64_t* keys, uint64_t* entry_off, bk_bucket_info* entries) {
    if (!ctx || !(ctx->I != nullptr) || !keys || !entry_off || !entries) return BK_ERR_ARG;
    memcpy(keys, ctx->I->ix.keys.data(), ctx->I->ix.keys.size() * 8);
    memcpy(entry_off, ctx->I->ix.entry_off.data(), ctx->I->ix.entry_off.size() * 8);
    memcpy(entries, ctx->I->ix.entries.data(), ctx->I->ix.entries.size() * sizeof(bk_bucket_info));
    return BK_OK;
}

// ---------------------------------------------------------------------------------------------
// sample
// ---------------------------------------------------------------------------------------------
static bool can_fuse_map(const bk_ctx* ctx);

int bk_sample_begin(bk_ctx* ctx, const bk_params* params) {
    if (!ctx) return BK_ERR_ARG;
    if (!(ctx->I != nullptr)) return ctx->fail(BK_ERR_ARG, "bk_sample_begin: no index loaded");
    if (!params) return ctx->fail(BK_ERR_ARG, "bk_sample_begin: null params");
    if (params->k != ctx->I->ix.k)
        return ctx->fail(BK_ERR_ARG, "Database k is not the same as provided, please set -k to %u or build a new index", ctx->I->ix.k);
    if (params->table_log2 != 0 && (params->table_log2 < 10 || params->table_log2 > 31))
        return ctx->fail(BK_ERR_ARG, "table_log2 must be 0 (auto) or in [10, 31]");
    cudaSetDevice(ctx->device);
    ctx->params = *params;
    ctx->in_sample = true; ctx->finished = false;
    for (FileState& f : ctx->file) { f.used = false; f.folded = false; f.finalized = false; f.total_reads = 0; f.total_bases = 0; f.nov_ub = 0; }
    ctx->spans_used = 0; ctx->launches = 0; ctx->scan_launches = 0; ctx->coll_calls = 0;
    ctx->variants.clear();
    memset(&ctx->result, 0, sizeof ctx->result);
    ctx->result.best_genome = -1;
    for (FqState* q : ctx->fq) fq_begin_sample(q);
    // (every stream of the context is idle here: the previous sample ended with a host wait on its last event)
    cudaStream_t st = ctx->s_count[0];
    cudaEventRecord(ctx->ev_begin, st);
    BK_CUDA(cudaMemsetAsync(ctx->d_ctr.p, 0, sizeof(Counters), st));
    const size_t pile_bytes = (size_t)ctx->I->d.max_genome_rows * 4 * 4 * 4;
    if (can_fuse_map(ctx)) BK_CUDA(cudaMemsetAsync(ctx->d_pile_all.p, 0, pile_bytes * ctx->I->d.n_genomes, st));
    else BK_CUDA(cudaMemsetAsync(ctx->d_pile.p, 0, pile_bytes, st));
    ctx->chain(st, ctx->s_count[1]);
    return BK_OK;
}

static CountView make_count_view(bk_ctx* ctx, int slot) {
    FileState& f = ctx->file[slot];
    CountView v;
    v.k = ctx->I->ix.k;
    v.refnib = ctx->I->d_refnib.p; v.ref_chunks = (u32)(ctx->I->d.refnib.size() / 4);
    v.oseq_start = ctx->I->d_oseq_start.p; v.oseq_len = ctx->I->d_oseq_len.p;
    v.exact = ctx->I->d_exact.p; v.exact_shift = 64 - ctx->I->d.exact_log2; v.exact_mask = (1u << ctx->I->d.exact_log2) - 1;
    v.diff = f.diff.p;
    v.gen = nullptr; v.gen_shift = 0; v.gen_mask = 0;      // (the open-addressing table of bk_core.cuh is only used by the CPU stepping of the tests)
    v.gen_full = &ctx->d_ctr.p->gen_full;
    v.nov = f.nov.p; v.nov_cap = (u32)std::min<u64>(f.nov.cap, f.nov_limit);
    v.nov_n = &ctx->d_ctr.p->f[slot].nov_n;
    const bool dense = ctx->I->d.dense_ok;
    v.nov_w = dense ? f.wa.p : nullptr;                    // (with mismatch lines the list is weighted: ambiguous cells join it)
    v.dense = dense ? f.dense.p : nullptr; v.dense_flag = dense ? f.dflag.p : nullptr;
    v.desc = f.desc.p; v.desc_cap = (u32)std::min<size_t>(f.desc.cap, 0xFFFFFFFFu); v.n_desc = &ctx->d_ctr.p->f[slot].n_desc;
    return v;
}

// CTAs per SM of the grid-stride kernels (leftover, map).  More and shorter-lived CTAs let the CTAs of higher-priority
// stages of other samples in flight start sooner (a CTA slot only frees when a CTA retires).
#ifndef BK_STRIDE_WAVES
#define BK_STRIDE_WAVES 8
#endif

// Developer builds only (tools/build_variant.sh ... -DBK_ABLATE): BK_ABLATE=scan,leftover,bins,map,noise skips the kernels of a
// stage so that its marginal cost with several samples in flight can be measured.  Results are garbage; the product
// build compiles this to `false`.
#ifdef BK_ABLATE
static bool ablated(const char* stage) { const char* e = getenv("BK_ABLATE"); return e && strstr(e, stage); }
#else
static inline bool ablated(const char*) { return false; }
#endif

// first use of a file slot in this sample: zero its difference array and per-id counts
static int file_prepare(bk_ctx* ctx, int slot) {
    FileState& f = ctx->file[slot];
    if (f.used) return BK_OK;
    cudaStream_t st = ctx->s_count[slot];
    BK_CUDA(cudaMemsetAsync(f.diff.p, 0, ((size_t)ctx->I->d.n_raw + 2) * 4, st));
    BK_CUDA(cudaMemsetAsync(f.idcnt.p, 0, std::max<size_t>(ctx->I->d.id_kmer.size(), 1) * 4, st));
    if (ctx->I->d.dense_ok) {
        if (f.dense_dirty) {                               // (normally the finish leaves the lines all zero)
            BK_CUDA(cudaMemsetAsync(f.dense.p, 0, f.dense.cap * 4, st));
            BK_CUDA(cudaMemsetAsync(f.dflag.p, 0, f.dflag.cap, st));
        }
        f.dense_dirty = true;
    }
    f.used = true;
    return BK_OK;
}

// Room in the file's list for the novel k-mer occurrences of a push of n_bases bases.  They cannot outnumber the
// bases, but at 0.2 % error only ~5 % of them are novel: the bound is kept per push, and when it no longer fits the
// allocation the real fill of the list is read back first (one sync per such push), so a deep sample pushed in chunks
// holds what it uses — not 8 bytes per base.  bk_params.table_log2 != 0 fixes the capacity instead.
static int novel_room(bk_ctx* ctx, int slot, u64 n_bases) {
    FileState& f = ctx->file[slot];
    cudaStream_t st = ctx->s_count[slot];
    const u64 LIMIT = 0xFFFFFFF0ull;
    if (ctx->params.table_log2) {
        const u64 need = 1ull << ctx->params.table_log2;
        if (f.nov_ub == 0) { BK_CUDA(f.nov.reserve(need)); if (ctx->I->d.dense_ok) BK_CUDA(f.wa.reserve(need)); }
        f.nov_ub = need; f.nov_limit = need;
        return BK_OK;
    }
    u64 need = f.nov_ub + n_bases + 64;
    if (need > f.nov.cap && f.nov_ub != 0) {
        u32 used = 0;
        BK_CUDA(cudaMemcpyAsync(&used, &ctx->d_ctr.p->f[slot].nov_n, 4, cudaMemcpyDeviceToHost, st));
        BK_CUDA(cudaStreamSynchronize(st));
        f.nov_ub = std::min<u64>(used, f.nov_limit);
        need = f.nov_ub + n_bases + 64;
    }
    if (need > LIMIT) return ctx->fail(BK_ERR_OVERFLOW, "more than 2^32 novel k-mer occurrences in one file");
    if (need > f.nov.cap) {
        const u64 want = std::min<u64>(need + need / 4, LIMIT);
        if (f.nov_ub != 0 && f.nov.p) {                  // not the first push of the file: keep the entries
            DevBuf<u64> bigger;
            BK_CUDA(bigger.reserve(want));
            BK_CUDA(cudaMemcpyAsync(bigger.p, f.nov.p, f.nov_ub * 8, cudaMemcpyDeviceToDevice, st));
            BK_CUDA(cudaStreamSynchronize(st));
            f.nov.release();
            f.nov = bigger;
            if (ctx->I->d.dense_ok) {
                DevBuf<u32> bw;
                BK_CUDA(bw.reserve(want));
                BK_CUDA(cudaMemcpyAsync(bw.p, f.wa.p, f.nov_ub * 4, cudaMemcpyDeviceToDevice, st));
                BK_CUDA(cudaStreamSynchronize(st));
                f.wa.release();
                f.wa = bw;
            }
        } else { BK_CUDA(f.nov.reserve(want)); if (ctx->I->d.dense_ok) BK_CUDA(f.wa.reserve(want)); }
    }
    f.nov_ub = need; f.nov_limit = need;
    return BK_OK;
}

// scan + leftover kernels over reads [r_begin, r_end) whose bytes live in d_bases (offset by off_bias)
static int launch_count(bk_ctx* ctx, int slot, const u8* d_bases, const u32* d_off, u32 off_bias, u32 r_begin, u32 r_end, u32 max_len) {
    FileState& f = ctx->file[slot];
    cudaStream_t st = ctx->s_count[slot];
    const u32 n = r_end - r_begin;
    if (n == 0) return BK_OK;
    BK_CUDA(f.desc.reserve((size_t)n * 2 + 4096));
    BK_CUDA(cudaMemsetAsync(&ctx->d_ctr.p->f[slot].n_desc, 0, 4, st));
    CountView v = make_count_view(ctx, slot);
    const u32 tile_bytes = 40 * 1024;
    u32 tile_reads = BK_SCAN_THREADS;
    if (max_len > 0) tile_reads = std::min<u32>(BK_SCAN_THREADS, std::max<u32>(1, tile_bytes / (max_len + 16)));
    if (tile_reads < 32) tile_reads = BK_SCAN_THREADS;   // long reads: tiles will not fit; kernel reads global memory
    const u32 n_tiles = (n + tile_reads - 1) / tile_reads;
    int sp = ctx->span_begin(ST_SCAN, st);
    if (!ablated("scan")) k_scan<<<grid_for(ctx, n_tiles, 1, 16), BK_SCAN_THREADS, tile_bytes + 64, st>>>(
        v, d_bases, d_off, off_bias, r_begin, r_end, tile_reads, tile_bytes, &ctx->d_ctr.p->f[slot].gen_new);
    ctx->span_end(sp);
    ctx->launches++; ctx->scan_launches++;
    BK_CUDA(cudaGetLastError());
    sp = ctx->span_begin(ST_LEFTOVER, st);
    if (!ablated("leftover")) k_leftover<1><<<ctx->sm_count * BK_STRIDE_WAVES, 256, 0, st>>>(v, d_bases, &ctx->d_ctr.p->f[slot].gen_new);
    ctx->launches++;
    ctx->span_end(sp);
    BK_CUDA(cudaGetLastError());
    return BK_OK;
}

static int check_push(bk_ctx* ctx, int slot) {
    if (!ctx) return BK_ERR_ARG;
    if (!ctx->in_sample || ctx->finished) return ctx->fail(BK_ERR_ARG, "bk_reads_push: call bk_sample_begin first");
    if (slot < 0 || slot > 1) return ctx->fail(BK_ERR_ARG, "bk_reads_push: file_slot must be 0 or 1");
    if (ctx->file[slot].folded) return ctx->fail(BK_ERR_ARG, "bk_reads_push: file already finalized");
    cudaSetDevice(ctx->device);
    return BK_OK;
}

int bk_reads_push_device(bk_ctx* ctx, int slot, const uint8_t* d_bases, const uint32_t* d_off, uint64_t n_reads,
                         uint64_t n_bases, uint32_t max_read_len) {
    int rc = check_push(ctx, slot);
    if (rc) return rc;
    if (n_reads == 0) { return file_prepare(ctx, slot); }
    if (!d_bases || !d_off) return ctx->fail(BK_ERR_ARG, "bk_reads_push_device: null buffer");
    if (((uintptr_t)d_bases & 15) != 0) return ctx->fail(BK_ERR_ARG, "bk_reads_push_device: bases must be 16-byte aligned");
    if (n_reads >= 0xFFFFFFFFull || n_bases >= 0xFFFFFFF0ull) return ctx->fail(BK_ERR_ARG, "bk_reads_push_device: push at most 2^32-16 bases / reads at a time");
    return guarded(ctx, [&]() -> int {
        int rc;
        if ((rc = file_prepare(ctx, slot))) return rc;
        if ((rc = novel_room(ctx, slot, n_bases))) return rc;
        ctx->file[slot].total_reads += n_reads; ctx->file[slot].total_bases += n_bases;
        return launch_count(ctx, slot, d_bases, d_off, 0, 0, (u32)n_reads, max_read_len);
    });
}

int bk_reads_push(bk_ctx* ctx, int slot, const uint8_t* bases, const uint32_t* read_off, uint64_t n_reads) {
    int rc = check_push(ctx, slot);
    if (rc) return rc;
    if (n_reads == 0) return file_prepare(ctx, slot);
    if (!bases || !read_off) return ctx->fail(BK_ERR_ARG, "bk_reads_push: null buffer");
    if (n_reads >= 0xFFFFFFFFull) return ctx->fail(BK_ERR_ARG, "bk_reads_push: too many reads in one push");
    return guarded(ctx, [&]() -> int {
    int rc;
    cudaStream_t st = ctx->s_count[slot];
    const u64 n_bases = read_off[n_reads];
    if ((rc = file_prepare(ctx, slot))) return rc;
    if ((rc = novel_room(ctx, slot, n_bases))) return rc;
    ctx->file[slot].total_reads += n_reads; ctx->file[slot].total_bases += n_bases;
    // offsets once, bases in chunks through two staging buffers so H2D overlaps the kernels.  Every copy from the
    // caller's buffers is issued on copy_stream, which is synchronised before returning: the caller may refill its
    // (pinned) buffers at once.  The offsets buffer is reused by every push: its copy waits for the kernels of the
    // previous push that still read it (off_free); a growing reserve() frees only after the device is idle (cudaFree).
    BK_CUDA(cudaStreamWaitEvent(ctx->copy_stream, ctx->off_free, 0));
    BK_CUDA(ctx->d_stage_off.reserve(n_reads + 1));
    BK_CUDA(cudaMemcpyAsync(ctx->d_stage_off.p, read_off, (n_reads + 1) * 4, cudaMemcpyHostToDevice, ctx->copy_stream));
    BK_CUDA(cudaEventRecord(ctx->off_copied, ctx->copy_stream));
    BK_CUDA(cudaStreamWaitEvent(st, ctx->off_copied, 0));
    const u64 CHUNK = 64ull << 20;
    u64 r = 0;
    while (r < n_reads) {
        u64 r_end = r;
        const u32 c_begin = read_off[r] & ~15u;          // chunk starts at a 16-byte boundary of the host buffer
        u32 max_len = 0;
        while (r_end < n_reads && (u64)read_off[r_end + 1] - c_begin <= CHUNK) {
            max_len = std::max(max_len, read_off[r_end + 1] - read_off[r_end]);
            r_end++;
        }
        if (r_end == r) { max_len = read_off[r + 1] - read_off[r]; r_end = r + 1; }   // a single read larger than CHUNK
        const u64 c_bytes = (u64)read_off[r_end] - c_begin;
        const int b = ctx->stage_next; ctx->stage_next ^= 1;
        BK_CUDA(ctx->d_stage[b].reserve(std::max<u64>(c_bytes, CHUNK) + 256));
        BK_CUDA(cudaStreamWaitEvent(ctx->copy_stream, ctx->stage_free[b], 0));
        BK_CUDA(cudaMemcpyAsync(ctx->d_stage[b].p, bases + c_begin, c_bytes, cudaMemcpyHostToDevice, ctx->copy_stream));
        BK_CUDA(cudaEventRecord(ctx->stage_copied[b], ctx->copy_stream));
        BK_CUDA(cudaStreamWaitEvent(st, ctx->stage_copied[b], 0));
        if ((rc = launch_count(ctx, slot, ctx->d_stage[b].p, ctx->d_stage_off.p, c_begin, (u32)r, (u32)r_end, max_len))) return rc;
        BK_CUDA(cudaEventRecord(ctx->stage_free[b], st));
        r = r_end;
    }
    BK_CUDA(cudaEventRecord(ctx->off_free, st));
    // the caller may reuse its buffers once we return: wait for the copies (not for the kernels)
    BK_CUDA(cudaStreamSynchronize(ctx->copy_stream));
    return BK_OK;
    });
}

int bk_reads_push_decoded(bk_ctx* ctx, int slot, const bk_reads* reads) {
    int rc = check_push(ctx, slot);
    if (rc) return rc;
    if (!reads) return ctx->fail(BK_ERR_ARG, "bk_reads_push_decoded: null reads");
    const uint64_t n_chunks = bk_reads_n_chunks(reads);
    bool any = false;
    for (uint64_t i = 0; i < n_chunks; i++) {
        const uint8_t* bases; const uint32_t* off; uint64_t n_reads, n_bases;
        bk_reads_chunk(reads, i, &bases, &off, &n_reads, &n_bases);
        if (n_reads == 0) continue;
        any = true;
        if ((rc = bk_reads_push(ctx, slot, bases, off, n_reads))) return rc;
        BK_CUDA(cudaStreamSynchronize(ctx->s_count[slot]));    // pageable source: the caller may free `reads` when this returns
    }
    if (!any) return file_prepare(ctx, slot);             // an empty file is still a file of the sample
    return BK_OK;
}

// ---- stages of bk_sample_finish (the read-sharded mode drives the same stages, bk_shard.inc) ---------------

static void launch_prefix(bk_ctx* ctx, u32* a, u32 n, u32* bsum, cudaStream_t st) {     // a[0..n] := exclusive prefix, a[n] = total
    const u32 nb = (n + BK_PS_BLOCK - 1) / BK_PS_BLOCK;
    k_diff_blocksum<<<nb, BK_PS_THREADS, 0, st>>>(a, n, bsum);
    k_diff_scan_bsum<<<1, BK_PS_THREADS, 0, st>>>(bsum, nb);
    k_excl_apply<<<nb, BK_PS_THREADS, 0, st>>>(a, n, bsum);
    ctx->launches += 3;
}

// stage 1: prefix sum of the difference array + fold onto distinct reference k-mers → idcnt
static int stage_fold(bk_ctx* ctx, int slot, cudaStream_t st) {
    FileState& f = ctx->file[slot];
    if (f.folded) return BK_OK;
    const DerivedIndex& d = ctx->I->d;
    const u32 n = d.n_raw;
    const u32 nb = (n + BK_PS_BLOCK - 1) / BK_PS_BLOCK;
    int sp = ctx->span_begin(ST_FINALIZE, st);
    BK_CUDA(cudaMemsetAsync(f.gstats.p, 0, (size_t)d.n_genomes * 16, st));      // (tallies: the mismatch-line cells add theirs before the map kernel does)
    k_diff_blocksum<<<nb, BK_PS_THREADS, 0, st>>>(f.diff.p, n, f.bsum.p);
    k_diff_scan_bsum<<<1, BK_PS_THREADS, 0, st>>>(f.bsum.p, nb);
    k_diff_apply<<<nb, BK_PS_THREADS, 0, st>>>(f.diff.p, n, f.bsum.p, ctx->I->d_slot2id.p, f.idcnt.p);
    ctx->span_end(sp);
    ctx->launches += 3;
    BK_CUDA(cudaGetLastError());
    f.folded = true;
    return BK_OK;
}

static CompactArgs make_compact_args(bk_ctx* ctx, int slot, size_t out_cap) {
    FileState& f = ctx->file[slot];
    CompactArgs a;
    a.ci = ctx->params.min_kmers; a.cs = ctx->params.counter_max; a.rank = ctx->shard_rank; a.n_ranks = ctx->shard_n();
    a.out_kmers = f.ckmers.p; a.out_counts = f.ccounts.p; a.out_cap = (u32)std::min<size_t>(out_cap, 0xFFFFFFFFu);
    a.fc = &ctx->d_ctr.p->f[slot];
    return a;
}

// the list of `ub` (an upper bound) entries → bins: P = 2^lp bins sized so that a worst-case round of a bin is ~1/32 of
// the bound at 3 % of it novel; histogram, prefix, scatter.  W: entries carry weights.
static int launch_bins(bk_ctx* ctx, int slot, BinView& b, u64 ub, bool weighted, cudaStream_t st, bool weighted_pairs = false) {
    FileState& f = ctx->file[slot];
    const DerivedIndex& d = ctx->I->d;
    // (with mismatch lines the list holds ~1/20 of what it held: fewer, fuller bins; a list that is large all the same —
    // foreign reads — is walked in rounds)
    if (d.dense_ok && !weighted_pairs) ub /= 16;
    u32 lp = 6;
    while (lp < 14 && ((u64)BK_BIN_ROUND << lp) < ub / 32) lp++;
    f.bin_log2p = lp;
    const u32 P = 1u << lp, G = (u32)ctx->sm_count * BK_BIN_G_PER_SM, PG = P * G;
    BK_CUDA(f.bin_cnt.reserve((size_t)PG + 1));
    b.cnt = f.bin_cnt.p; b.log2p = lp; b.G = G;
    b.exact = ctx->I->d_exact.p; b.exact_shift = 64 - d.exact_log2; b.exact_mask = (1u << d.exact_log2) - 1;
    b.slot2id = ctx->I->d_slot2id.p; b.idcnt = f.idcnt.p;
    b.k = d.k;
    if (d.dense_ok) {
        b.nb = ctx->I->d_nb.p; b.nb_shift = 64 - d.nb_log2; b.nb_mask = (1u << d.nb_log2) - 1;
        b.nb_bloom = ctx->I->d_nb_bloom.p; b.nb_bloom_log2 = d.nb_bloom_log2;
        b.id_amb = ctx->I->d_id_amb.p; b.id_rep = ctx->I->d_id_rep.p; b.dense = f.dense.p; b.dense_flag = f.dflag.p;
    }
    k_bin_hist<<<G, BK_BIN_G_THREADS, P * 4, st>>>(b);
    launch_prefix(ctx, f.bin_cnt.p, PG, f.bsum.p, st);
    if (weighted) k_bin_scatter<true><<<G, BK_BIN_G_THREADS, P * 4, st>>>(b);
    else k_bin_scatter<false><<<G, BK_BIN_G_THREADS, P * 4, st>>>(b);
    ctx->launches += 2;
    BK_CUDA(cudaGetLastError());
    return BK_OK;
}

static MapView make_map_view(bk_ctx* ctx);
static bool dense_maps(const bk_ctx* ctx) { return ctx->I->d.dense_ok && ctx->I->d.map_shortcut_ok && can_fuse_map(ctx) && !ctx->shard; }
static DenseView make_dense_view(bk_ctx* ctx, int slot) {
    FileState& f = ctx->file[slot];
    const DerivedIndex& d = ctx->I->d;
    DenseView dv; memset(&dv, 0, sizeof dv);
    if (!d.dense_ok) return dv;
    dv.dense = f.dense.p; dv.flag = f.dflag.p; dv.n_lines = d.n_raw * 4; dv.k = d.k;
    dv.slot2rep = ctx->I->d_slot2rep.p; dv.slot2id = ctx->I->d_slot2id.p; dv.id_kmer = ctx->I->d_id_kmer.p; dv.id_amb = ctx->I->d_id_amb.p;
    dv.line_amb = ctx->I->d_line_amb.p; dv.line_fold = ctx->I->d_line_fold.p;
    dv.nov = f.nov.p; dv.nov_w = f.wa.p; dv.nov_n = &ctx->d_ctr.p->f[slot].nov_n; dv.nov_cap = (u32)std::min<u64>(f.nov.cap, f.nov_limit);
    dv.full = &ctx->d_ctr.p->gen_full;
    if (dense_maps(ctx)) {
        dv.id_bucket = ctx->I->d_id_bucket.p; dv.m = make_map_view(ctx); dv.gstats = f.gstats.p;
        dv.pile = ctx->d_pile_all.p; dv.pile_stride = d.max_genome_rows * 4;
    }
    return dv;
}

// mismatch-line cells that join the list (the ambiguous ones; all of them on a read-sharded rank) need room behind the
// occurrences the pushes were given room for: at most one entry per cell, and no more cells than k-mers
static u64 dense_cells_max(const bk_ctx* ctx, const FileState& f) {
    const DerivedIndex& d = ctx->I->d;
    return d.dense_ok ? std::min<u64>((u64)d.n_raw * 4 * d.k, f.total_bases) : 0;
}
static int dense_list_room(bk_ctx* ctx, int slot, cudaStream_t st) {
    FileState& f = ctx->file[slot];
    if (!ctx->I->d.dense_ok || ctx->params.table_log2) return BK_OK;      // (a caller-fixed capacity is what it is: overflow is reported)
    const u64 need = std::min<u64>(f.nov_ub + dense_cells_max(ctx, f) + 64, 0xFFFFFFF0ull);
    if (need > f.nov.cap) {
        DevBuf<u64> bk_; DevBuf<u32> bw;
        BK_CUDA(bk_.reserve(need)); BK_CUDA(bw.reserve(need));
        if (f.nov.p && f.nov_ub) {
            BK_CUDA(cudaMemcpyAsync(bk_.p, f.nov.p, std::min<u64>(f.nov_ub, f.nov.cap) * 8, cudaMemcpyDeviceToDevice, st));
            BK_CUDA(cudaMemcpyAsync(bw.p, f.wa.p, std::min<u64>(f.nov_ub, f.wa.cap) * 4, cudaMemcpyDeviceToDevice, st));
            BK_CUDA(cudaStreamSynchronize(st));
        }
        f.nov.release(); f.wa.release();
        f.nov = bk_; f.wa = bw;
    }
    f.nov_ub = need; f.nov_limit = need;
    return BK_OK;
}

// stage 2: the KMC dump of one file (threshold, cap, the four stdout numbers): novel k-mers through the bins, then the
// reference k-mers (after the bins: they add to idcnt)
static int stage_compact(bk_ctx* ctx, int slot, cudaStream_t st) {
    FileState& f = ctx->file[slot];
    if (f.finalized) return BK_OK;
    const DerivedIndex& d = ctx->I->d;
    const u32 n_ids = (u32)d.id_kmer.size();
    // the novel part of the list: a kept k-mer stands for >= ci occurrences; mismatch-line cells (kept where they are, or
    // through the list if ambiguous): at most one k-mer each
    const size_t novel_cap = std::min<u64>(f.nov.cap, std::max<u64>(f.nov_ub, 1)) / std::max<u32>(ctx->params.min_kmers, 1) + 1;
    const size_t out_cap = (size_t)n_ids + novel_cap + dense_cells_max(ctx, f);
    { int rc = dense_list_room(ctx, slot, st); if (rc) return rc; }
    BK_CUDA(f.ckmers.reserve(out_cap)); BK_CUDA(f.ccounts.reserve(out_cap));
    f.ck_cap = out_cap;
    BK_CUDA(f.sorted.reserve(std::max<size_t>(std::min<u64>(f.nov.cap, std::max<u64>(f.nov_ub, 1)), 1)));
    int sp = ctx->span_begin(ST_FINALIZE, st);
    const CompactArgs a = make_compact_args(ctx, slot, out_cap);
    const bool dense = d.dense_ok;
    DenseView dv = make_dense_view(ctx, slot);
    const int dgrid = dense ? grid_for(ctx, (u64)dv.n_lines, 256, 4) : 0;
    if (dense) {                                           // mismatch lines: counts per cell, one cell per string, ambiguous ones to the list
        BK_CUDA(f.wb.reserve(f.sorted.cap));
        k_dense_prefix<<<dgrid, 256, 0, st>>>(dv);
        k_dense_fold<<<dgrid, 256, 0, st>>>(dv);
        k_dense_emit<0, false><<<dgrid, 256, 0, st>>>(dv, a);
        ctx->launches += 3;
    }
    BinView b; memset(&b, 0, sizeof b);
    b.nov = f.nov.p; b.nov_n = &ctx->d_ctr.p->f[slot].nov_n; b.nov_cap = (u32)std::min<u64>(f.nov.cap, f.nov_limit);
    b.sorted = f.sorted.p;
    if (dense) { b.nov_w = f.wa.p; b.sorted_w = f.wb.p; }
    if (!ablated("bins")) {
        int rc = launch_bins(ctx, slot, b, f.nov_ub, dense, st);
        if (rc) return rc;
        if (ablated("bincount")) {}
        else if (dense) k_bin_count<0, true><<<1u << b.log2p, 256, BK_BIN_SMEM, st>>>(b, a, &ctx->d_ctr.p->gen_full);
        else k_bin_count<0, false><<<1u << b.log2p, 256, BK_BIN_SMEM, st>>>(b, a, &ctx->d_ctr.p->gen_full);
        ctx->launches++;
    }
    if (dense) {                                           // what is left on the lines is final: cut-offs, counted list, lines back to zero
        if (dense_maps(ctx)) k_dense_emit<1, true><<<dgrid, 256, 0, st>>>(dv, a);       // ... and mapped on the spot: tallies + all-genome pileups
        else k_dense_emit<1, false><<<dgrid, 256, 0, st>>>(dv, a);
        ctx->launches++;
        f.dense_dirty = false;
    }
    if (dense && dense_maps(ctx) && !getenv("BK_NO_ID_MAP"))       // the reference k-mers: compacted and mapped in one go
        k_compact_ids_map<<<std::max(1u, std::min((n_ids + BK_IDMAP_IDS - 1) / BK_IDMAP_IDS, (u32)ctx->sm_count * 16u)), 256, 0, st>>>(dv, a, f.idcnt.p, n_ids);
    else k_compact_ids<<<grid_for(ctx, n_ids, 256), 256, 0, st>>>(a, f.idcnt.p, ctx->I->d_id_kmer.p, n_ids);
    ctx->launches++;
    ctx->span_end(sp);
    BK_CUDA(cudaGetLastError());
    f.finalized = true;
    return BK_OK;
}

static MapView make_map_view(bk_ctx* ctx) {
    MapView m;
    const u32 k = ctx->I->ix.k;
    m.k = k;
    // src/call.rs:1291-1300: buckets[n_fixed .. k - n_fixed - 1) unless --use-full-kmer
    if (ctx->params.use_full_kmer) { m.b0 = 0; m.b1 = k; }
    else if ((u64)ctx->params.n_fixed * 2 + 1 >= k) { m.b0 = 0; m.b1 = 0; }
    else { m.b0 = ctx->params.n_fixed; m.b1 = k - ctx->params.n_fixed - 1; }
    m.slots = ctx->I->d_bucket_slots.p; m.shift = 64 - ctx->I->d.bucket_log2; m.mask = (1u << ctx->I->d.bucket_log2) - 1;
    m.entries = ctx->I->d_bucket_entries.p;
    m.n_genomes = ctx->I->d.n_genomes; m.genome_row0 = ctx->I->d_genome_row0.p;
    const bool grouped = ctx->I->d.rekeyed && !ctx->I->d.group_slots.empty() && !ctx->no_group_map;
    m.gslots = grouped ? ctx->I->d_group_slots.p : nullptr; m.gcenters = ctx->I->d_group_centers.p; m.gbuckets = ctx->I->d_group_buckets.p;
    m.gshift = 64 - ctx->I->d.group_log2; m.gmask = (1u << ctx->I->d.group_log2) - 1; m.gmid = ctx->I->d.group_mid;
    return m;
}

static int n_files_used(bk_ctx* ctx) { return ctx->file[1].used ? 2 : 1; }

// per-genome hit counts of one k-mer fit the 16-bit fields of the thread-per-k-mer kernels (a query hits at most k buckets)
static bool hits16_ok(const bk_ctx* ctx) { return (u64)ctx->I->d.max_key_entries * ctx->I->ix.k < 65536ull; }
static bool small_db(const bk_ctx* ctx) { return ctx->I->d.n_genomes <= 4 && !ctx->force_warp_map && hits16_ok(ctx); }
// tallies and the pileups of all genomes in ONE pass per file (databases of at most four genomes, re-keyed table),
// selection afterwards, then the selected genome's arrays are moved to d_pile
static bool can_fuse_map(const bk_ctx* ctx) { return small_db(ctx) && ctx->I->d.rekeyed && !ctx->no_fused_map; }

// one-pass map of one file (can_fuse_map): tallies + the pileups of every genome (src/call.rs:1316-1385)
static int stage_map_fused(bk_ctx* ctx, int f, cudaStream_t st) {
    const DerivedIndex& d = ctx->I->d;
    const MapView m = make_map_view(ctx);
    const u32 pile_stride = d.max_genome_rows * 4;
    Counters* dc = ctx->d_ctr.p;
    FileState& fs = ctx->file[f];
    int sp = ctx->span_begin(ST_MAP, st);
    const u32 ccap = (u32)std::min<size_t>(fs.ckmers.cap, 0xFFFFFFFFu);
    if (ablated("map")) {}
    else if (m.gslots) k_map_grp<2><<<ctx->sm_count * BK_STRIDE_WAVES, 256, 0, st>>>(m, fs.ckmers.p, fs.ccounts.p, &dc->f[f].n_counted, ccap, fs.gstats.p, nullptr, ctx->d_pile_all.p, pile_stride);
    else k_map_small<2, 1><<<ctx->sm_count * BK_STRIDE_WAVES, 256, 0, st>>>(m, fs.ckmers.p, fs.ccounts.p, &dc->f[f].n_counted, ccap, fs.gstats.p, nullptr, ctx->d_pile_all.p, pile_stride);
    ctx->launches++;
    ctx->span_end(sp);
    BK_CUDA(cudaGetLastError());
    return BK_OK;
}

// pick_best_genome(_paired) (src/call.rs:422-502); after a fused map also the move of the selected genome's arrays
static int stage_select(bk_ctx* ctx, bool fused, cudaStream_t st) {
    const DerivedIndex& d = ctx->I->d;
    const int n_files = n_files_used(ctx);
    Counters* dc = ctx->d_ctr.p;
    int sp = ctx->span_begin(ST_MAP, st);
    k_select<<<1, 32, 0, st>>>(ctx->file[0].gstats.p, ctx->file[n_files - 1].gstats.p, n_files, d.n_genomes, ctx->I->d_genome_len.p, dc);
    ctx->launches++;
    if (fused) { k_pile_pick<<<ctx->sm_count, 256, 0, st>>>(ctx->d_pile_all.p, ctx->d_pile.p, d.max_genome_rows * 4, &dc->best); ctx->launches++; }
    ctx->span_end(sp);
    BK_CUDA(cudaGetLastError());
    return BK_OK;
}

// two-pass map, pass 1: map_kmers tallies of one file (src/call.rs:1389-1430)
static int stage_map_stats(bk_ctx* ctx, int f, cudaStream_t st) {
    const DerivedIndex& d = ctx->I->d;
    const MapView m = make_map_view(ctx);
    const size_t map_smem = (size_t)d.n_genomes * 12 * 4;
    const bool small = small_db(ctx);
    Counters* dc = ctx->d_ctr.p;
    FileState& fs = ctx->file[f];
    int sp = ctx->span_begin(ST_MAP, st);
    const u32 ccap = (u32)std::min<size_t>(fs.ckmers.cap, 0xFFFFFFFFu);
    if (small && m.gslots) k_map_grp<0><<<ctx->sm_count * BK_STRIDE_WAVES, 256, 0, st>>>(m, fs.ckmers.p, fs.ccounts.p, &dc->f[f].n_counted, ccap, fs.gstats.p, nullptr, nullptr, 0);
    else if (small) (d.rekeyed ? k_map_small<0, 1> : k_map_small<0, 0>)<<<ctx->sm_count * BK_STRIDE_WAVES, 256, 0, st>>>(m, fs.ckmers.p, fs.ccounts.p, &dc->f[f].n_counted, ccap, fs.gstats.p, nullptr, nullptr, 0);
    else (d.rekeyed ? k_map<0, 1> : k_map<0, 0>)<<<ctx->sm_count * BK_STRIDE_WAVES, 256, map_smem, st>>>(m, fs.ckmers.p, fs.ccounts.p, &dc->f[f].n_counted, ccap, fs.gstats.p, nullptr, nullptr, 0);
    ctx->launches++;
    ctx->span_end(sp);
    BK_CUDA(cudaGetLastError());
    return BK_OK;
}

// two-pass map, pass 2: the selected genome's pileup from one file (src/call.rs:1324-1385)
static int stage_map_pileup(bk_ctx* ctx, int f, cudaStream_t st) {
    const DerivedIndex& d = ctx->I->d;
    const MapView m = make_map_view(ctx);
    const bool small = small_db(ctx);
    const u32 pile_stride = d.max_genome_rows * 4;
    Counters* dc = ctx->d_ctr.p;
    FileState& fs = ctx->file[f];
    int sp = ctx->span_begin(ST_MAP, st);
    const u32 ccap = (u32)std::min<size_t>(fs.ckmers.cap, 0xFFFFFFFFu);
    if (small && m.gslots) k_map_grp<1><<<ctx->sm_count * BK_STRIDE_WAVES, 256, 0, st>>>(m, fs.ckmers.p, fs.ccounts.p, &dc->f[f].n_counted, ccap, nullptr, &dc->best, ctx->d_pile.p, pile_stride);
    else if (small) (d.rekeyed ? k_map_small<1, 1> : k_map_small<1, 0>)<<<ctx->sm_count * BK_STRIDE_WAVES, 256, 0, st>>>(m, fs.ckmers.p, fs.ccounts.p, &dc->f[f].n_counted, ccap, nullptr, &dc->best, ctx->d_pile.p, pile_stride);
    else (d.rekeyed ? k_map<1, 1> : k_map<1, 0>)<<<ctx->sm_count * BK_STRIDE_WAVES, 256, 0, st>>>(m, fs.ckmers.p, fs.ccounts.p, &dc->f[f].n_counted, ccap, nullptr, &dc->best, ctx->d_pile.p, pile_stride);
    ctx->launches++;
    ctx->span_end(sp);
    BK_CUDA(cudaGetLastError());
    return BK_OK;
}

// last stage: noise baseline + call_variants, read everything back, fill bk_sample_result
static int stage_score(bk_ctx* ctx, bk_sample_result* out, cudaStream_t st) {
    const DerivedIndex& d = ctx->I->d;
    const int n_files = n_files_used(ctx);
    const u32 pile_stride = d.max_genome_rows * 4;
    Counters* dc = ctx->d_ctr.p;
    int sp = ctx->span_begin(ST_SCORE, st);
    ScoreView sv;
    sv.n_genomes = d.n_genomes; sv.genome_row0 = ctx->I->d_genome_row0.p; sv.genome_seq_off = ctx->I->d_genome_seq_off.p;
    sv.seq_row0 = ctx->I->d_seq_row0.p; sv.ref_code = ctx->I->d_ref_code.p; sv.ctr = dc; sv.pile = ctx->d_pile.p; sv.pile_stride = pile_stride;
    const u32 row_blocks = (d.max_genome_rows + 255) / 256;
    NoiseView nv;
    nv.ctr = dc; nv.genome_row0 = ctx->I->d_genome_row0.p; nv.genome_seq_off = ctx->I->d_genome_seq_off.p; nv.seq_row0 = ctx->I->d_seq_row0.p;
    nv.pile = ctx->d_pile.p; nv.pile_stride = pile_stride;
    nv.maf = ctx->d_nz_maf.p; nv.snap_s = ctx->d_nz_s.p; nv.snap_s2 = ctx->d_nz_s2.p; nv.snap_tab = ctx->d_nz_tab.p;
    nv.warm = ctx->d_nz_warm.p; nv.flag = ctx->d_nz_flag.p; nv.stats = ctx->noise_debug ? ctx->d_nz_stats.p : nullptr;
    nv.noise_max = ctx->d_noise.p;
    nv.actflag = ctx->d_nz_nzflag.p; nv.act_list = ctx->d_nz_list.p; nv.act_rank = ctx->d_nz_rank.p; nv.act_n = ctx->d_nz_nact.p;
    const u32 nseq = ctx->I->max_seqs_per_genome;
    if (ctx->noise_debug) cudaMemsetAsync(ctx->d_nz_stats.p, 0, 64, st);
    k_noise_fracs<<<dim3((d.max_genome_rows + BK_NZ_PAD + 255) / 256, nseq), 256, 0, st>>>(nv);
    if (!ablated("noise")) k_noise_seq<<<dim3(2 + (ctx->nz_max_chunks + 7) / 8, nseq), BK_NZ_SEQ_THREADS, BK_NZ_SEQ_SMEM, st>>>(nv);
    k_noise_fix<<<dim3(1, nseq), 256, 0, st>>>(nv);
    k_noise_tau<<<dim3((d.max_genome_rows + BK_NOISE_HALF + 255) / 256, nseq), 256, 0, st>>>(nv);
    if (ctx->noise_debug) {        // BK_NOISE_DEBUG=1
        u32 h[16];
        cudaMemcpyAsync(h, ctx->d_nz_stats.p, 64, cudaMemcpyDeviceToHost, st);
        cudaStreamSynchronize(st);
        fprintf(stderr, "[noise] table chunks replayed %u (%u iterations); chain rounds %u, stops %u, serial iterations %u; kcycles: s %u, s2 %u, slowest table lane %u\n",
                h[0], h[1], h[2], h[3], h[4], h[5] / 64, h[6] / 64, h[7] / 64);
#ifdef BK_NZ_WHY
        fprintf(stderr, "[noise] stops by cause: operand too large %u, near a zone border %u, above the zones %u, below the zones %u; iterations accepted in stopped rounds %u\n", h[8], h[9], h[10], h[11], h[12]);
#endif
#ifdef BK_NZ_PHASES
        fprintf(stderr, "[noise] s2 chain kcycles by phase: tile %u, operands+maps %u, warp scan %u, combine %u, check+reduce %u, prefix %u, stop %u, serial %u\n",
                h[8] / 64, h[9] / 64, h[10] / 64, h[11] / 64, h[12] / 64, h[13] / 64, h[14] / 64, h[15] / 64);
#endif
    }
    CallParams cp;
    const bk_params& p = ctx->params;
    cp.k = p.k; cp.no_end_filter = p.no_end_filter; cp.no_strand_filter = p.no_strand_filter;
    cp.no_strand_balance_filter = p.no_strand_balance_filter; cp.n_per_strand = p.n_per_strand;
    cp.min_depth = p.min_depth; cp.min_variant_depth = p.min_variant_depth; cp.min_af = p.min_af;
    cp.strand_balance_ratio = p.strand_balance_ratio; cp.strand_odds_max = p.strand_odds_max; cp.variant_multiplier = p.variant_multiplier;
    k_call<<<row_blocks, 256, 0, st>>>(sv, cp, ctx->d_noise.p, ctx->d_vars.p, (u32)std::min<size_t>(ctx->d_vars.cap, 0xFFFFFFFFu), dc);
    ctx->span_end(sp);
    ctx->launches += 5;
    BK_CUDA(cudaGetLastError());

    BK_CUDA(cudaMemcpyAsync(ctx->h_ctr, dc, sizeof(Counters), cudaMemcpyDeviceToHost, st));
    if (ctx->pinned_u32(&ctx->h_gstats, &ctx->h_gstats_cap, (size_t)d.n_genomes * 8)) return ctx->fail(BK_ERR_NOMEM, "out of pinned host memory");
    const u32* hg[2] = {ctx->h_gstats, ctx->h_gstats + (size_t)d.n_genomes * 4};
    for (int f = 0; f < n_files; f++)
        BK_CUDA(cudaMemcpyAsync(ctx->h_gstats + (size_t)f * d.n_genomes * 4, ctx->file[f].gstats.p, (size_t)d.n_genomes * 16, cudaMemcpyDeviceToHost, st));
    const size_t n_eager = std::min<size_t>(BK_VARS_EAGER, ctx->d_vars.cap);
    BK_CUDA(cudaMemcpyAsync(ctx->h_vars, ctx->d_vars.p, n_eager * sizeof(bk_variant), cudaMemcpyDeviceToHost, st));   // (the variants of nearly every sample: no second round trip)
    if (ctx->shard) {                                    // the globally reduced KMC numbers of both files (bk_shard.cuh: k_shard_pack)
        const u32 stride = 4 + d.n_genomes * 4;
        BK_CUDA(cudaMemcpyAsync(ctx->h_stats, ctx->d_shard_stats.p, 4 * sizeof(u64), cudaMemcpyDeviceToHost, st));
        BK_CUDA(cudaMemcpyAsync(ctx->h_stats + 4, ctx->d_shard_stats.p + stride, 4 * sizeof(u64), cudaMemcpyDeviceToHost, st));
    }
    BK_CUDA(cudaEventRecord(ctx->ev_end, st));
    BK_CUDA(cudaEventSynchronize(ctx->ev_end));          // blocking wait (BK_SPIN=1 spins): the host thread sleeps while the GPU works
    const Counters& c = *ctx->h_ctr;
    ctx->finished = true;
    if (getenv("BK_DEBUG_COUNTS"))
        for (int f = 0; f < n_files; f++)
            fprintf(stderr, "[counts] file %d: reads %llu, leftover stretches (last push) %u, list entries %u, counted front %u + back %u, distinct %u, total k-mers %llu\n", f,
                    (unsigned long long)ctx->file[f].total_reads, c.f[f].n_desc, c.f[f].nov_n, c.f[f].n_counted, c.f[f].n_dense, c.f[f].unique, (unsigned long long)c.f[f].total_kmers);
    if (c.gen_full) return ctx->fail(BK_ERR_OVERFLOW, "no room left for novel k-mers (list / bin table); set bk_params.table_log2 higher");
    if (c.var_overflow) return ctx->fail(BK_ERR_OVERFLOW, "variant buffer overflow");
    for (int f = 0; f < n_files; f++) {
        if ((size_t)c.f[f].n_counted + c.f[f].n_dense > ctx->file[f].ckmers.cap) return ctx->fail(BK_ERR_OVERFLOW, "counted k-mer list overflow");
        ctx->gstats[f].assign(d.n_genomes, bk_genome_stats());
        for (u32 g = 0; g < d.n_genomes; g++) {
            bk_genome_stats& s = ctx->gstats[f][g];
            s.perfect = hg[f][g * 4]; s.variant = hg[f][g * 4 + 1]; s.unique_perfect = hg[f][g * 4 + 2]; s.present = hg[f][g * 4 + 3] ? 1 : 0; s._pad = 0;
        }
    }
    bk_sample_result& r = ctx->result;
    memset(&r, 0, sizeof r);
    r.best_genome = c.best; r.n_files = n_files;
    for (int f = 0; f < n_files; f++) {
        if (ctx->shard) {
            r.kmc[f].total_reads = ctx->h_stats[f * 4]; r.kmc[f].total_kmers = ctx->h_stats[f * 4 + 1];
            r.kmc[f].unique_kmers = ctx->h_stats[f * 4 + 2]; r.kmc[f].unique_counted = ctx->h_stats[f * 4 + 3];
        } else {
            r.kmc[f].total_reads = ctx->file[f].total_reads; r.kmc[f].total_kmers = c.f[f].total_kmers;
            r.kmc[f].unique_kmers = c.f[f].unique; r.kmc[f].unique_counted = (u64)c.f[f].n_counted + c.f[f].n_dense;
        }
    }
    if (c.best < 0) {
        if (out) *out = r;
        return ctx->fail(BK_ERR_NO_GENOME, "Unable to pick a best genome");
    }
    r.n_variants = c.n_var; r.num_major_variants = c.n_major; r.num_minor_variants = c.n_minor;
    u64 total_positions = 0;
    for (const HostSeq& q : ctx->I->ix.genomes[c.best].seqs) total_positions += q.bases.size();
    r.breadth_coverage = (double)c.pos_covered / (double)total_positions;      // src/call.rs:1144-1145
    r.depth_coverage = (double)c.total_cov / (double)c.pos_covered;
    u64 uc = 0, pv = 0;
    for (int f = 0; f < n_files; f++) {
        uc += r.kmc[f].unique_counted;
        r.num_perfect_kmers += ctx->gstats[f][c.best].perfect; r.num_variant_kmers += ctx->gstats[f][c.best].variant;
    }
    pv = r.num_perfect_kmers + r.num_variant_kmers;
    r.num_unmapped_kmers = uc - pv;                                             // src/call.rs:242, 336 (usize arithmetic)
    ctx->variants.resize(c.n_var);
    if (c.n_var) {
        memcpy(ctx->variants.data(), ctx->h_vars, std::min<size_t>(c.n_var, n_eager) * sizeof(bk_variant));
        if (c.n_var > n_eager) {                         // (thousands of variants: the rest, through a pinned bounce buffer)
            bk_variant* more = nullptr;
            if (cudaMallocHost((void**)&more, (size_t)(c.n_var - n_eager) * sizeof(bk_variant)) != cudaSuccess) return ctx->fail(BK_ERR_NOMEM, "out of pinned host memory");
            cudaError_t e = cudaMemcpyAsync(more, ctx->d_vars.p + n_eager, (size_t)(c.n_var - n_eager) * sizeof(bk_variant), cudaMemcpyDeviceToHost, st);
            if (e == cudaSuccess) e = ctx->host_wait(st);
            if (e == cudaSuccess) memcpy(ctx->variants.data() + n_eager, more, (size_t)(c.n_var - n_eager) * sizeof(bk_variant));
            cudaFreeHost(more);
            BK_CUDA(e);
        }
        std::sort(ctx->variants.begin(), ctx->variants.end(), [](const bk_variant& a, const bk_variant& b) {
            if (a.seq != b.seq) return a.seq < b.seq;
            if (a.pos != b.pos) return a.pos < b.pos;
            return a.alt_base < b.alt_base;
        });
    }
    bk_stage_times& t = ctx->times;
    memset(&t, 0, sizeof t);
    float* acc[ST_N] = {&t.scan_ms, &t.leftover_ms, &t.finalize_ms, &t.map_ms, &t.score_ms, &t.coll_ms, &t.decode_ms};
    for (size_t i = 0; i < ctx->spans_used; i++) {
        float ms = 0;
        if (cudaEventElapsedTime(&ms, ctx->spans[i].a, ctx->spans[i].b) == cudaSuccess) *acc[ctx->spans[i].stage] += ms;
    }
    cudaEventElapsedTime(&t.total_ms, ctx->ev_begin, ctx->ev_end);
    t.launches = ctx->launches; t.scan_launches = ctx->scan_launches; t.coll_calls = ctx->coll_calls;
    if (out) *out = r;
    return BK_OK;
}

static int check_finish(bk_ctx* ctx, const char* who) {
    if (!ctx) return BK_ERR_ARG;
    if (!ctx->in_sample || ctx->finished) return ctx->fail(BK_ERR_ARG, "%s: no sample in progress", who);
    if (!ctx->file[0].used) return ctx->fail(BK_ERR_ARG, "%s: no reads were pushed to file slot 0", who);
    cudaSetDevice(ctx->device);
    return BK_OK;
}

static int shard_finish(ShardGroup& G, std::vector<bk_ctx*>& ms, bk_sample_result* out);      // bk_shard.inc

int bk_sample_finish(bk_ctx* ctx, bk_sample_result* out) {
    int rc = check_finish(ctx, "bk_sample_finish");
    if (rc) return rc;
    return guarded(ctx, [&]() -> int {
    int rc;
    if (ctx->shard) {
        if (ctx->shard->local) return ctx->fail(BK_ERR_ARG, "bk_sample_finish: context belongs to an in-process shard group; call bk_shard_finish_local");
        std::vector<bk_ctx*> me(1, ctx);
        return shard_finish(*ctx->shard, me, out);
    }
    const int n_files = n_files_used(ctx);
    const bool fused = can_fuse_map(ctx);
    for (int f = 0; f < n_files; f++) {                    // the files run side by side until the selection
        cudaStream_t st = ctx->s_fin[f];
        ctx->chain(ctx->s_count[f], st);
        if ((rc = stage_fold(ctx, f, st))) return rc;
        if ((rc = stage_compact(ctx, f, st))) return rc;
        if ((rc = fused ? stage_map_fused(ctx, f, st) : stage_map_stats(ctx, f, st))) return rc;
        ctx->chain(st, ctx->s_score);
    }
    if ((rc = stage_select(ctx, fused, ctx->s_score))) return rc;
    if (!fused) {
        for (int f = 0; f < n_files; f++) {
            ctx->chain(ctx->s_score, ctx->s_fin[f]);
            if ((rc = stage_map_pileup(ctx, f, ctx->s_fin[f]))) return rc;
        }
        for (int f = 0; f < n_files; f++) ctx->chain(ctx->s_fin[f], ctx->s_score);
    }
    return stage_score(ctx, out, ctx->s_score);
    });
}

#include "bk_shard.inc"
#include "bk_fastq.inc"

int bk_stage_timing(bk_ctx* ctx, int on) {
    if (!ctx) return BK_ERR_ARG;
    ctx->stage_timing = on != 0;
    return BK_OK;
}

int bk_stage_times_get(bk_ctx* ctx, bk_stage_times* out) {
    if (!ctx || !out) return BK_ERR_ARG;
    *out = ctx->times;
    return BK_OK;
}

int bk_sample_result_get(bk_ctx* ctx, bk_sample_result* out) {
    if (!ctx || !ctx->finished || !out) return BK_ERR_ARG;
    *out = ctx->result;
    return BK_OK;
}

int bk_sample_variants(bk_ctx* ctx, bk_variant* out, uint64_t cap) {
    if (!ctx || !ctx->finished) return BK_ERR_ARG;
    if (cap < ctx->variants.size()) return ctx->fail(BK_ERR_ARG, "bk_sample_variants: buffer too small");
    if (!ctx->variants.empty()) memcpy(out, ctx->variants.data(), ctx->variants.size() * sizeof(bk_variant));
    return BK_OK;
}

int bk_sample_genome_stats(bk_ctx* ctx, int slot, bk_genome_stats* out) {
    if (!ctx || !ctx->finished || slot < 0 || slot > 1 || !out) return BK_ERR_ARG;
    if (ctx->gstats[slot].size() != ctx->I->d.n_genomes) return ctx->fail(BK_ERR_ARG, "no stats for file slot %d", slot);
    memcpy(out, ctx->gstats[slot].data(), ctx->gstats[slot].size() * sizeof(bk_genome_stats));
    return BK_OK;
}

int bk_sample_pileup(bk_ctx* ctx, int arr, uint64_t* out, uint64_t cap_rows) {
    if (!ctx || !ctx->finished || arr < 0 || arr > 3 || !out) return BK_ERR_ARG;
    const int best = ctx->result.best_genome;
    if (best < 0) return ctx->fail(BK_ERR_NO_GENOME, "no genome selected");
    cudaSetDevice(ctx->device);
    const u32 rows = ctx->I->d.genome_row0[best + 1] - ctx->I->d.genome_row0[best];
    if (cap_rows < rows) return ctx->fail(BK_ERR_ARG, "bk_sample_pileup: buffer too small (%u rows)", rows);
    std::vector<u32> tmp((size_t)rows * 4);
    BK_CUDA(cudaMemcpyAsync(tmp.data(), ctx->d_pile.p + (size_t)arr * ctx->I->d.max_genome_rows * 4, tmp.size() * 4, cudaMemcpyDeviceToHost, ctx->s_score));
    BK_CUDA(cudaStreamSynchronize(ctx->s_score));
    for (size_t i = 0; i < tmp.size(); i++) out[i] = tmp[i];
    return BK_OK;
}

int bk_sample_noise_max(bk_ctx* ctx, double* out, uint64_t cap_rows) {
    if (!ctx || !ctx->finished || !out) return BK_ERR_ARG;
    const int best = ctx->result.best_genome;
    if (best < 0) return ctx->fail(BK_ERR_NO_GENOME, "no genome selected");
    cudaSetDevice(ctx->device);
    const u32 rows = ctx->I->d.genome_row0[best + 1] - ctx->I->d.genome_row0[best];
    if (cap_rows < rows) return ctx->fail(BK_ERR_ARG, "bk_sample_noise_max: buffer too small");
    BK_CUDA(cudaMemcpyAsync(out, ctx->d_noise.p, (size_t)rows * 8, cudaMemcpyDeviceToHost, ctx->s_score));
    BK_CUDA(cudaStreamSynchronize(ctx->s_score));
    return BK_OK;
}

int bk_kmer_counts_get(bk_ctx* ctx, int slot, uint64_t* kmers, uint32_t* counts, uint64_t* n) {
    if (!ctx || !ctx->finished || slot < 0 || slot > 1 || !n) return BK_ERR_ARG;
    if (!ctx->file[slot].finalized) return ctx->fail(BK_ERR_ARG, "file slot %d has no counts", slot);
    cudaSetDevice(ctx->device);
    const u64 have = ctx->result.kmc[slot].unique_counted;
    if (!kmers || !counts) { *n = have; return BK_OK; }
    if (*n < have) return ctx->fail(BK_ERR_ARG, "bk_kmer_counts_get: buffer too small");
    *n = have;
    if (!have) return BK_OK;
    std::vector<u64> hk(have); std::vector<u32> hc(have);
    // the list has a front (k-mers the map kernel handled) and a back (mismatch-line cells mapped where they were counted)
    const u64 n_back = ctx->shard ? 0 : ctx->h_ctr->f[slot].n_dense, n_front = have - n_back;
    const size_t cap = ctx->file[slot].ck_cap;
    if (n_front) {
        BK_CUDA(cudaMemcpyAsync(hk.data(), ctx->file[slot].ckmers.p, n_front * 8, cudaMemcpyDeviceToHost, ctx->s_score));
        BK_CUDA(cudaMemcpyAsync(hc.data(), ctx->file[slot].ccounts.p, n_front * 4, cudaMemcpyDeviceToHost, ctx->s_score));
    }
    if (n_back) {
        BK_CUDA(cudaMemcpyAsync(hk.data() + n_front, ctx->file[slot].ckmers.p + (cap - n_back), n_back * 8, cudaMemcpyDeviceToHost, ctx->s_score));
        BK_CUDA(cudaMemcpyAsync(hc.data() + n_front, ctx->file[slot].ccounts.p + (cap - n_back), n_back * 4, cudaMemcpyDeviceToHost, ctx->s_score));
    }
    BK_CUDA(cudaStreamSynchronize(ctx->s_score));
    std::vector<u32> order(have);
    for (u64 i = 0; i < have; i++) order[i] = (u32)i;
    std::sort(order.begin(), order.end(), [&](u32 a, u32 b) { return hk[a] < hk[b]; });
    for (u64 i = 0; i < have; i++) { kmers[i] = hk[order[i]]; counts[i] = hc[order[i]]; }
    return BK_OK;
}

// ---------------------------------------------------------------------------------------------
// writers — reference src/call.rs:735-774 (VCF) and 648-695 (pileup TSV)
// ---------------------------------------------------------------------------------------------
static bool write_all(const char* path, const std::string& s) {
    FILE* f = fopen(path, "wb");
    if (!f) return false;
    const size_t w = fwrite(s.data(), 1, s.size(), f);
    fclose(f);
    return w == s.size();
}

int bk_write_vcf(bk_ctx* ctx, const char* reads_path, const char* out_path) {
    if (!ctx || !ctx->finished || !reads_path || !out_path) return BK_ERR_ARG;
    const int best = ctx->result.best_genome;
    if (best < 0) return ctx->fail(BK_ERR_NO_GENOME, "no genome selected");
    const HostGenome& g = ctx->I->ix.genomes[best];
    std::string o;
    o += "##fileformat=VCFv4.5\n##source=bronko-v0.1.0\n";
    o += std::string("##reference=file://") + reads_path + "\n";
    for (const HostSeq& q : g.seqs) o += "##contig=<ID=" + first_token(q.name) + ",length=" + std::to_string(q.len) + ">\n";
    o += "##INFO=<ID=DP,Number=1,Type=Integer,Description=\"Total Depth\">\n";
    o += "##INFO=<ID=AF,Number=1,Type=Float,Description=\"Allele Frequency\">\n";
    o += "##INFO=<ID=DP4,Number=4,Type=Integer,Description=\"Fwd_ref,Rev_ref,Fwd_alt,Rev_alt\">\n";
    o += "##INFO=<ID=SOR,Number=4,Type=Float,Description=\"SOR\">\n";
    o += "#CHROM\tPOS\tID\tREF\tALT\tQUAL\tFILTER\tINFO\n";
    for (const bk_variant& v : ctx->variants) {
        o += first_token(g.seqs[v.seq].name) + "\t" + std::to_string(v.pos) + "\t.\t";
        o += "ACGT"[v.ref_base & 3]; o += "\t"; o += "ACGT"[v.alt_base & 3];
        o += "\t.\tPASS\tDP=" + std::to_string(v.depth) + ";AF=" + fmt_fixed(v.af, 3) + ";DP4=" + std::to_string(v.fwd_ref) + "," +
             std::to_string(v.rev_ref) + "," + std::to_string(v.fwd_alt) + "," + std::to_string(v.rev_alt) + ";SOR=" + fmt_fixed(v.sor, 3) + "\n";
    }
    if (!write_all(out_path, o)) return ctx->fail(BK_ERR_IO, "Failed to create vcf output file %s", out_path);
    return BK_OK;
}

int bk_write_pileup(bk_ctx* ctx, const char* out_path) {
    if (!ctx || !ctx->finished || !out_path) return BK_ERR_ARG;
    const int best = ctx->result.best_genome;
    if (best < 0) return ctx->fail(BK_ERR_NO_GENOME, "no genome selected");
    const HostGenome& g = ctx->I->ix.genomes[best];
    const u32 rows = ctx->I->d.genome_row0[best + 1] - ctx->I->d.genome_row0[best];
    std::vector<u64> fw((size_t)rows * 4), rv((size_t)rows * 4);
    int rc;
    if ((rc = bk_sample_pileup(ctx, 0, fw.data(), rows)) || (rc = bk_sample_pileup(ctx, 1, rv.data(), rows))) return rc;
    std::string o = "reference\tindex\tref\tA\tC\tG\tT\ta\tc\tg\tt\n";
    size_t row = 0;
    for (const HostSeq& q : g.seqs) {
        for (size_t i = 0; i < q.bases.size(); i++, row++) {
            o += q.name + "\t" + std::to_string(i + 1) + "\t"; o += (char)q.bases[i];
            for (int b = 0; b < 4; b++) o += "\t" + std::to_string(fw[row * 4 + b]);
            for (int b = 0; b < 4; b++) o += "\t" + std::to_string(rv[row * 4 + b]);
            o += "\n";
        }
    }
    if (!write_all(out_path, o)) return ctx->fail(BK_ERR_IO, "Failed to create tsv pileup file %s", out_path);
    return BK_OK;
}

}  // extern "C"
