// bk_device.cu — bk_ctx, device memory, kernel launches and the C ABI of libbronko_b200.so
// (include/bronko_b200.h).  There is no CPU fallback anywhere in this file: every stage of a sample
// runs as a CUDA kernel from bk_kernels.cuh, and bk_create fails without an sm_100 device.
#include <algorithm>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <string>
#include <vector>

#include "bk_host.h"
#include "bk_kernels.cuh"

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

enum Stage { ST_SCAN = 0, ST_LEFTOVER, ST_FINALIZE, ST_MAP, ST_SCORE, ST_N };

struct FileState {
    DevBuf<u32> diff, idcnt;
    DevBuf<GenSlot> gen;
    DevBuf<u64> ckmers;
    DevBuf<u32> ccounts;
    DevBuf<u32> gstats;
    u32 gen_log2 = 0;
    DevBuf<u64> xk; DevBuf<u32> xc;          // sharded mode: novel (k-mer, count) pairs grouped by owner rank
    DevBuf<u64> nov;                         // list mode (bk_bins.cuh): novel k-mer occurrences of the file (grouped by bin in ctx->d_nov_sorted)
    DevBuf<u32> bin_cnt;
    bool list_mode = false;                  // novel k-mers through the list + bins instead of the gen table
    u64 nov_ub = 0;                          // upper bound of list entries the pushes so far were given room for
    u64 nov_limit = 0;                       // entries the kernels may use (the allocation can be larger: it is reused)
    u32 bin_log2p = 8;
    bool used = false, folded = false, finalized = false;
    u64 total_reads = 0, total_bases = 0;
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
    DevBuf<u32> d_genome_row0, d_genome_seq_off, d_seq_row0; DevBuf<u64> d_genome_len; DevBuf<u8> d_ref_code;
    u32 max_seqs_per_genome = 1;
    ~IndexDev() {
        cudaSetDevice(device);
        d_bucket_slots.release(); d_bucket_entries.release(); d_group_slots.release(); d_group_centers.release(); d_group_buckets.release(); d_refnib.release(); d_oseq_start.release(); d_oseq_len.release();
        d_exact.release(); d_slot2id.release(); d_id_kmer.release(); d_genome_row0.release(); d_genome_seq_off.release();
        d_seq_row0.release(); d_genome_len.release(); d_ref_code.release();
    }
};

}  // namespace

struct bk_ctx {
    int device = 0;
    cudaStream_t stream = nullptr, copy_stream = nullptr;
    // Stage priorities.  `stream` is the stream the sample's work is currently issued to: stage_stream[0] while reads are
    // pushed (scan / leftover), [1] for finalize + map, [2] for the score stage, each with a higher CUDA priority than the
    // one before and chained by an event.  Every big kernel fills the SMs' register files, so with several samples in
    // flight (one context each) a kernel's CTAs only start as CTAs of other kernels retire, and without priorities the
    // handful of CTAs of a sample's last, latency-bound stages (the noise chains) queue behind thousands of pending CTAs
    // of other samples' first stages.  BK_PRIO=0 keeps everything on one stream.
    cudaStream_t stage_stream[3] = {nullptr, nullptr, nullptr};
    cudaEvent_t ev_chain = nullptr;
    int level = 0, prio_mode = 2;
    void use_level(int lvl) {
        if (prio_mode == 0) return;
        if (prio_mode == 1) lvl = lvl == 2 ? 2 : 0;
        if (lvl == level) return;
        cudaEventRecord(ev_chain, stage_stream[level]);
        cudaStreamWaitEvent(stage_stream[lvl], ev_chain, 0);
        level = lvl; stream = stage_stream[lvl];
    }
    std::string err;
    int sm_count = 148;

    std::shared_ptr<IndexDev> I;            // null until an index is loaded / built / shared

    // per-sample state
    bk_params params;
    bool in_sample = false, finished = false;
    FileState file[2];
    DevBuf<Counters> d_ctr;
    Counters* h_ctr = nullptr;              // pinned
    DevBuf<uint2> d_desc; DevBuf<u32> d_bsum;
    DevBuf<u64> d_nov_sorted;               // list mode: one file's novel k-mers grouped by bin (files are finalized one after the other)
    DevBuf<u32> d_pile;                     // 4 arrays x max_genome_rows x 4
    DevBuf<u32> d_pile_all;                 // one such block per genome (databases of at most four genomes: one-pass map)
    DevBuf<double> d_noise;                 // Noise.max per row
    DevBuf<double> d_nz_maf, d_nz_s, d_nz_s2, d_nz_tab, d_nz_warm;   // noise scratch (bk_noise.cuh: NoiseView)
    DevBuf<u8> d_nz_flag; DevBuf<u32> d_nz_stats;
    u32 nz_max_chunks = 1;
    DevBuf<bk_variant> d_vars;
    // staging for host pushes (double buffered)
    DevBuf<u8> d_stage[2]; DevBuf<u32> d_stage_off;
    cudaEvent_t stage_free[2] = {nullptr, nullptr}, stage_copied[2] = {nullptr, nullptr};
    int stage_next = 0;
    u32 shard_rank = 0, shard_n = 1;
    bk_kmc_stats shard_kmc[2];
    DevBuf<u32> d_part;
    bool noise_debug = false;
    bool force_warp_map = false;            // tests: exercise the many-genome map kernel on a small db
    bool no_fused_map = false;              // tests: BK_NO_FUSED_MAP keeps the two-pass map on small databases
    bool no_group_map = false;              // tests: BK_NO_GROUP_MAP probes the per-bucket table instead of the grouped one
    bool novel_table = false;               // tests: BK_NOVEL_TABLE counts novel k-mers in the global hash table (what the sharded mode uses)

    // results
    bk_sample_result result;
    std::vector<bk_variant> variants;
    std::vector<bk_genome_stats> gstats[2];

    // timing
    struct Span { cudaEvent_t a, b; int stage; };
    std::vector<Span> spans; size_t spans_used = 0;
    cudaEvent_t ev_begin = nullptr, ev_end = nullptr;
    u32 launches = 0, scan_launches = 0;
    bk_stage_times times;

    int fail(int code, const char* fmt, ...) {
        char buf[1024];
        va_list ap; va_start(ap, fmt); vsnprintf(buf, sizeof buf, fmt, ap); va_end(ap);
        err = buf;
        return code;
    }
    int span_begin(int stage) {
        if (spans_used == spans.size()) {
            Span s; s.stage = stage;
            if (cudaEventCreate(&s.a) != cudaSuccess || cudaEventCreate(&s.b) != cudaSuccess) return -1;
            spans.push_back(s);
        }
        spans[spans_used].stage = stage;
        cudaEventRecord(spans[spans_used].a, stream);
        return (int)spans_used++;
    }
    void span_end(int id) { if (id >= 0) cudaEventRecord(spans[id].b, stream); }
};

static int grid_for(const bk_ctx* ctx, u64 items, u32 per_block, u32 max_waves = 8) {
    u64 g = (items + per_block - 1) / per_block;
    const u64 cap = (u64)ctx->sm_count * max_waves;
    if (g > cap) g = cap;
    if (g < 1) g = 1;
    return (int)g;
}

extern "C" {

const char* bk_version(void) { return "bronko_b200 0.1.0 (sm_100a)"; }

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
    bk_ctx* ctx = new bk_ctx();
    ctx->device = device;
    ctx->sm_count = prop.multiProcessorCount;
    int prio_least = 0, prio_greatest = 0;
    cudaDeviceGetStreamPriorityRange(&prio_least, &prio_greatest);              // numerically lower = higher priority
    if (const char* e = getenv("BK_PRIO")) ctx->prio_mode = atoi(e);
    if (prio_greatest >= prio_least) ctx->prio_mode = 0;
    bool ok = cudaStreamCreateWithPriority(&ctx->stage_stream[0], cudaStreamNonBlocking, prio_least) == cudaSuccess &&
              cudaStreamCreateWithFlags(&ctx->copy_stream, cudaStreamNonBlocking) == cudaSuccess &&
              cudaEventCreateWithFlags(&ctx->ev_chain, cudaEventDisableTiming) == cudaSuccess &&
              cudaMallocHost((void**)&ctx->h_ctr, sizeof(Counters)) == cudaSuccess &&
              cudaEventCreate(&ctx->ev_begin) == cudaSuccess && cudaEventCreate(&ctx->ev_end) == cudaSuccess;
    if (ok && ctx->prio_mode != 0)
        ok = cudaStreamCreateWithPriority(&ctx->stage_stream[1], cudaStreamNonBlocking, std::max(prio_greatest, prio_least - 1)) == cudaSuccess &&
             cudaStreamCreateWithPriority(&ctx->stage_stream[2], cudaStreamNonBlocking, prio_greatest) == cudaSuccess;
    ctx->stream = ctx->stage_stream[0];
    for (int i = 0; i < 2 && ok; i++)
        ok = cudaEventCreateWithFlags(&ctx->stage_free[i], cudaEventDisableTiming) == cudaSuccess &&
             cudaEventCreateWithFlags(&ctx->stage_copied[i], cudaEventDisableTiming) == cudaSuccess;
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
                 cudaFuncSetAttribute(k_bin_scatter, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024) == cudaSuccess &&
                 cudaFuncSetAttribute(k_bin_count, cudaFuncAttributeMaxDynamicSharedMemorySize, BK_BIN_SMEM) == cudaSuccess;
    if (!ok) { g_create_error = std::string("context setup failed: ") + cudaGetErrorString(cudaGetLastError()); delete ctx; return BK_ERR_CUDA; }
    memset(&ctx->times, 0, sizeof ctx->times);
    memset(&ctx->result, 0, sizeof ctx->result);
    bk_params_default(&ctx->params);
    ctx->force_warp_map = getenv("BK_FORCE_WARP_MAP") != nullptr;
    ctx->noise_debug = getenv("BK_NOISE_DEBUG") != nullptr;
    ctx->no_fused_map = getenv("BK_NO_FUSED_MAP") != nullptr;
    ctx->novel_table = getenv("BK_NOVEL_TABLE") != nullptr;
    ctx->no_group_map = getenv("BK_NO_GROUP_MAP") != nullptr;
    *out = ctx;
    return BK_OK;
}

void bk_destroy(bk_ctx* ctx) {
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    cudaDeviceSynchronize();
    ctx->I.reset();                         // the last context sharing an index frees its device copies
    for (FileState& f : ctx->file) { f.diff.release(); f.idcnt.release(); f.gen.release(); f.ckmers.release(); f.ccounts.release(); f.gstats.release(); f.xk.release(); f.xc.release(); f.nov.release(); f.bin_cnt.release(); }
    ctx->d_part.release();
    ctx->d_ctr.release(); ctx->d_desc.release(); ctx->d_bsum.release(); ctx->d_nov_sorted.release(); ctx->d_pile.release(); ctx->d_pile_all.release();
    ctx->d_noise.release(); ctx->d_vars.release();
    ctx->d_nz_maf.release(); ctx->d_nz_s.release(); ctx->d_nz_s2.release(); ctx->d_nz_tab.release(); ctx->d_nz_warm.release();
    ctx->d_nz_flag.release(); ctx->d_nz_stats.release();
    ctx->d_stage[0].release(); ctx->d_stage[1].release(); ctx->d_stage_off.release();
    for (auto& s : ctx->spans) { cudaEventDestroy(s.a); cudaEventDestroy(s.b); }
    for (int i = 0; i < 2; i++) { if (ctx->stage_free[i]) cudaEventDestroy(ctx->stage_free[i]); if (ctx->stage_copied[i]) cudaEventDestroy(ctx->stage_copied[i]); }
    if (ctx->ev_begin) cudaEventDestroy(ctx->ev_begin);
    if (ctx->ev_end) cudaEventDestroy(ctx->ev_end);
    if (ctx->h_ctr) cudaFreeHost(ctx->h_ctr);
    for (cudaStream_t st : ctx->stage_stream) if (st) cudaStreamDestroy(st);
    if (ctx->ev_chain) cudaEventDestroy(ctx->ev_chain);
    if (ctx->copy_stream) cudaStreamDestroy(ctx->copy_stream);
    delete ctx;
}

const char* bk_last_error(bk_ctx* ctx) { return ctx ? ctx->err.c_str() : g_create_error.c_str(); }
void* bk_stream(bk_ctx* ctx) { return ctx ? (void*)ctx->stage_stream[0] : nullptr; }     // the stream pushes are issued to

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
    cudaStream_t st = ctx->stream;
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
        BK_CUDA(ctx->d_nz_tab.reserve(it_slots * BK_NOISE_TABLE));
        BK_CUDA(ctx->d_nz_warm.reserve(chunk_slots * BK_NOISE_TABLE));
        BK_CUDA(ctx->d_nz_flag.reserve(chunk_slots));
        BK_CUDA(ctx->d_nz_stats.reserve(16));
    }
    BK_CUDA(ctx->d_vars.reserve(rows * 3));
    BK_CUDA(ctx->d_ctr.reserve(1));
    BK_CUDA(ctx->d_bsum.reserve(std::max<size_t>((size_t)d.n_raw + 2, (size_t)16384 * ctx->sm_count * BK_BIN_G_PER_SM) / BK_PS_BLOCK + 2));
    for (FileState& f : ctx->file) {
        BK_CUDA(f.diff.reserve((size_t)d.n_raw + 2));
        BK_CUDA(f.idcnt.reserve(d.id_kmer.size()));
        BK_CUDA(f.gstats.reserve((size_t)d.n_genomes * 4));
    }
    ctx->in_sample = false; ctx->finished = false;
    return BK_OK;
}

int bk_index_share(bk_ctx* ctx, bk_ctx* owner) {
    if (!ctx || !owner) return BK_ERR_ARG;
    if (!owner->I) return ctx->fail(BK_ERR_ARG, "bk_index_share: the other context holds no index");
    if (owner->device != ctx->device) return ctx->fail(BK_ERR_ARG, "bk_index_share: contexts live on different devices");
    if (ctx->in_sample && !ctx->finished) return ctx->fail(BK_ERR_ARG, "bk_index_share: call between samples");
    ctx->I = owner->I;
    return size_for_index(ctx);
}

int bk_index_load(bk_ctx* ctx, uint32_t k, uint64_t n_keys, const uint64_t* keys, const uint64_t* entry_off,
                  const bk_bucket_info* entries, uint32_t n_genomes, const uint32_t* genome_seq_off,
                  const uint64_t* seq_len, const uint64_t* seq_base_off, const uint8_t* ref_bases) {
    if (!ctx) return BK_ERR_ARG;
    if (!keys || !entry_off || !entries || !genome_seq_off || !seq_len || !seq_base_off || !ref_bases)
        return ctx->fail(BK_ERR_ARG, "bk_index_load: null argument");
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
}

int bk_index_load_file(bk_ctx* ctx, const char* path) {
    if (!ctx || !path) return BK_ERR_ARG;
    std::string err;
    auto fresh = std::make_shared<IndexDev>();
    if (!bkdb_read(path, fresh->ix, err)) return ctx->fail(BK_ERR_IO, "%s", err.c_str());
    return upload_index(ctx, fresh);
}

int bk_index_build(bk_ctx* ctx, uint32_t k, uint32_t n_files, const char* const* fasta_paths) {
    if (!ctx || !fasta_paths || n_files == 0) return BK_ERR_ARG;
    if (k < 15 || k > 31 || (k & 1) == 0) return ctx->fail(BK_ERR_ARG, "Invalid kmer size, must be odd and between [15-31]");
    std::vector<std::string> paths(fasta_paths, fasta_paths + n_files);
    std::string err;
    auto fresh = std::make_shared<IndexDev>();
    if (!index_build_from_fasta(k, paths, fresh->ix, err)) return ctx->fail(BK_ERR_IO, "%s", err.c_str());
    return upload_index(ctx, fresh);
}

int bk_index_save(bk_ctx* ctx, const char* path) {
    if (!ctx || !path || !(ctx->I != nullptr)) return BK_ERR_ARG;
    std::string err;
    if (!bkdb_write(path, ctx->I->ix, err)) return ctx->fail(BK_ERR_IO, "%s", err.c_str());
    return BK_OK;
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
    ctx->spans_used = 0; ctx->launches = 0; ctx->scan_launches = 0;
    ctx->use_level(0);                    // (after the previous sample's last stage, by the event chain)
    ctx->variants.clear();
    memset(&ctx->result, 0, sizeof ctx->result);
    ctx->result.best_genome = -1;
    cudaEventRecord(ctx->ev_begin, ctx->stream);
    BK_CUDA(cudaMemsetAsync(ctx->d_ctr.p, 0, sizeof(Counters), ctx->stream));
    return BK_OK;
}

static CountView make_count_view(bk_ctx* ctx, FileState& f) {
    CountView v;
    v.k = ctx->I->ix.k;
    v.refnib = ctx->I->d_refnib.p; v.ref_chunks = (u32)(ctx->I->d.refnib.size() / 4);
    v.oseq_start = ctx->I->d_oseq_start.p; v.oseq_len = ctx->I->d_oseq_len.p;
    v.exact = ctx->I->d_exact.p; v.exact_shift = 64 - ctx->I->d.exact_log2; v.exact_mask = (1u << ctx->I->d.exact_log2) - 1;
    v.diff = f.diff.p;
    v.gen = f.gen.p; v.gen_shift = 64 - f.gen_log2; v.gen_mask = (u32)((1ull << f.gen_log2) - 1);
    v.gen_full = &ctx->d_ctr.p->gen_full;
    v.nov = f.list_mode ? f.nov.p : nullptr; v.nov_cap = (u32)std::min<u64>(f.nov.cap, f.nov_limit);
    v.nov_n = &ctx->d_ctr.p->f[&f - ctx->file].nov_n;
    v.desc = ctx->d_desc.p; v.desc_cap = (u32)std::min<size_t>(ctx->d_desc.cap, 0xFFFFFFFFu); v.n_desc = &ctx->d_ctr.p->n_desc;
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

// first use of a file slot in this sample: zero its difference array and (re)initialise its table
static int file_prepare(bk_ctx* ctx, int slot, u64 bases_hint) {
    FileState& f = ctx->file[slot];
    if (f.used) return BK_OK;
    f.list_mode = ctx->shard_n == 1 && !ctx->novel_table;
    BK_CUDA(cudaMemsetAsync(f.diff.p, 0, ((size_t)ctx->I->d.n_raw + 2) * 4, ctx->stream));
    BK_CUDA(cudaMemsetAsync(f.idcnt.p, 0, std::max<size_t>(ctx->I->d.id_kmer.size(), 1) * 4, ctx->stream));
    if (!f.list_mode) {
        u32 lg = ctx->params.table_log2;
        if (lg == 0) {            // auto: ~1 slot per 32 read bases of this first push, clamped to [2^22, 2^27]
            lg = 22;
            while (lg < 27 && (1ull << lg) < bases_hint / 32) lg++;
        }
        f.gen_log2 = lg;
        BK_CUDA(f.gen.reserve(1ull << lg));
        k_gen_init<<<grid_for(ctx, 1ull << lg, 256 * 8), 256, 0, ctx->stream>>>(f.gen.p, 1ull << lg);
        ctx->launches++;
        BK_CUDA(cudaGetLastError());
    }
    f.used = true;
    return BK_OK;
}

// list mode: room for the novel k-mer occurrences of a push of n_bases bases (they cannot outnumber the bases);
// what earlier pushes of the file wrote is kept.  bk_params.table_log2 != 0 fixes the capacity instead.
static int novel_room(bk_ctx* ctx, int slot, u64 n_bases) {
    FileState& f = ctx->file[slot];
    if (!f.list_mode) return BK_OK;
    u64 need;
    const u64 before = f.nov_ub;
    if (ctx->params.table_log2) { need = 1ull << ctx->params.table_log2; f.nov_ub = need; }
    else { f.nov_ub += n_bases + 64; need = f.nov_ub; }
    need = std::min<u64>(need, 0xFFFFFFF0ull);
    f.nov_limit = need;
    if (need <= f.nov.cap) return BK_OK;
    if (before != 0 && f.nov.p) {                      // not the first push of the file: keep the entries
        DevBuf<u64> bigger;
        BK_CUDA(bigger.reserve(need + need / 2));
        BK_CUDA(cudaMemcpyAsync(bigger.p, f.nov.p, f.nov.cap * 8, cudaMemcpyDeviceToDevice, ctx->stream));
        BK_CUDA(cudaStreamSynchronize(ctx->stream));
        f.nov.release();
        f.nov = bigger;
    } else BK_CUDA(f.nov.reserve(need));
    return BK_OK;
}

// scan + leftover kernels over reads [r_begin, r_end) whose bytes live in d_bases (offset by off_bias)
static int launch_count(bk_ctx* ctx, int slot, const u8* d_bases, const u32* d_off, u32 off_bias, u32 r_begin, u32 r_end, u32 max_len) {
    FileState& f = ctx->file[slot];
    const u32 n = r_end - r_begin;
    if (n == 0) return BK_OK;
    BK_CUDA(ctx->d_desc.reserve((size_t)n * 2 + 4096));
    BK_CUDA(cudaMemsetAsync(&ctx->d_ctr.p->n_desc, 0, 4, ctx->stream));
    CountView v = make_count_view(ctx, f);
    const u32 tile_bytes = 40 * 1024;
    u32 tile_reads = BK_SCAN_THREADS;
    if (max_len > 0) tile_reads = std::min<u32>(BK_SCAN_THREADS, std::max<u32>(1, tile_bytes / (max_len + 16)));
    if (tile_reads < 32) tile_reads = BK_SCAN_THREADS;   // long reads: tiles will not fit; kernel reads global memory
    const u32 n_tiles = (n + tile_reads - 1) / tile_reads;
    int sp = ctx->span_begin(ST_SCAN);
    if (!ablated("scan")) k_scan<<<grid_for(ctx, n_tiles, 1, 16), BK_SCAN_THREADS, tile_bytes + 64, ctx->stream>>>(
        v, d_bases, d_off, off_bias, r_begin, r_end, tile_reads, tile_bytes, &ctx->d_ctr.p->f[slot].gen_new);
    ctx->span_end(sp);
    ctx->launches++; ctx->scan_launches++;
    BK_CUDA(cudaGetLastError());
    sp = ctx->span_begin(ST_LEFTOVER);
    if (ablated("leftover")) {}
    else if (f.list_mode) k_leftover<1><<<ctx->sm_count * BK_STRIDE_WAVES, 256, 0, ctx->stream>>>(v, d_bases, &ctx->d_ctr.p->f[slot].gen_new);
    else k_leftover<0><<<ctx->sm_count * BK_STRIDE_WAVES, 256, 0, ctx->stream>>>(v, d_bases, &ctx->d_ctr.p->f[slot].gen_new);
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
    if (n_reads == 0) { return file_prepare(ctx, slot, 0); }
    if (!d_bases || !d_off) return ctx->fail(BK_ERR_ARG, "bk_reads_push_device: null buffer");
    if (((uintptr_t)d_bases & 15) != 0) return ctx->fail(BK_ERR_ARG, "bk_reads_push_device: bases must be 16-byte aligned");
    if (n_reads >= 0xFFFFFFFFull || n_bases >= 0xFFFFFFF0ull) return ctx->fail(BK_ERR_ARG, "bk_reads_push_device: push at most 2^32-16 bases / reads at a time");
    if ((rc = file_prepare(ctx, slot, n_bases))) return rc;
    if ((rc = novel_room(ctx, slot, n_bases))) return rc;
    ctx->file[slot].total_reads += n_reads; ctx->file[slot].total_bases += n_bases;
    return launch_count(ctx, slot, d_bases, d_off, 0, 0, (u32)n_reads, max_read_len);
}

int bk_reads_push(bk_ctx* ctx, int slot, const uint8_t* bases, const uint32_t* read_off, uint64_t n_reads) {
    int rc = check_push(ctx, slot);
    if (rc) return rc;
    if (n_reads == 0) return file_prepare(ctx, slot, 0);
    if (!bases || !read_off) return ctx->fail(BK_ERR_ARG, "bk_reads_push: null buffer");
    if (n_reads >= 0xFFFFFFFFull) return ctx->fail(BK_ERR_ARG, "bk_reads_push: too many reads in one push");
    const u64 n_bases = read_off[n_reads];
    if ((rc = file_prepare(ctx, slot, n_bases))) return rc;
    if ((rc = novel_room(ctx, slot, n_bases))) return rc;
    ctx->file[slot].total_reads += n_reads; ctx->file[slot].total_bases += n_bases;
    // offsets once, bases in chunks through two staging buffers so H2D overlaps the kernels
    BK_CUDA(ctx->d_stage_off.reserve(n_reads + 1));
    BK_CUDA(cudaMemcpyAsync(ctx->d_stage_off.p, read_off, (n_reads + 1) * 4, cudaMemcpyHostToDevice, ctx->stream));
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
        BK_CUDA(cudaStreamWaitEvent(ctx->stream, ctx->stage_copied[b], 0));
        if ((rc = launch_count(ctx, slot, ctx->d_stage[b].p, ctx->d_stage_off.p, c_begin, (u32)r, (u32)r_end, max_len))) return rc;
        BK_CUDA(cudaEventRecord(ctx->stage_free[b], ctx->stream));
        r = r_end;
    }
    // the caller may reuse its buffers once we return: wait for the copies (not for the kernels)
    BK_CUDA(cudaStreamSynchronize(ctx->copy_stream));
    BK_CUDA(cudaEventSynchronize(ctx->stage_copied[ctx->stage_next ^ 1]));
    return BK_OK;
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
        BK_CUDA(cudaStreamSynchronize(ctx->stream));    // pageable source: the caller may free `reads` when this returns
    }
    if (!any) return file_prepare(ctx, slot, 0);          // an empty file is still a file of the sample
    return BK_OK;
}

int bk_reads_push_fastq(bk_ctx* ctx, int slot, const char* path) {
    int rc = check_push(ctx, slot);
    if (rc) return rc;
    if (!path) return ctx->fail(BK_ERR_ARG, "bk_reads_push_fastq: null path");
    bk_reads* reads = nullptr;
    char err[512] = "";
    if (bk_fastq_decode(path, &reads, err, sizeof err) != BK_OK) return ctx->fail(BK_ERR_IO, "%s", err);
    rc = bk_reads_push_decoded(ctx, slot, reads);
    bk_reads_free(reads);
    return rc;
}

// ---- stages of bk_sample_finish (also driven one by one in the read-sharded mode) ---------------

// stage 1: prefix sum of the difference array + fold onto distinct reference k-mers → idcnt
static int stage_fold(bk_ctx* ctx, int slot) {
    ctx->use_level(1);
    FileState& f = ctx->file[slot];
    if (f.folded) return BK_OK;
    const DerivedIndex& d = ctx->I->d;
    const u32 n = d.n_raw;
    const u32 nb = (n + BK_PS_BLOCK - 1) / BK_PS_BLOCK;
    int sp = ctx->span_begin(ST_FINALIZE);
    k_diff_blocksum<<<nb, BK_PS_THREADS, 0, ctx->stream>>>(f.diff.p, n, ctx->d_bsum.p);
    k_diff_scan_bsum<<<1, BK_PS_THREADS, 0, ctx->stream>>>(ctx->d_bsum.p, nb);
    k_diff_apply<<<nb, BK_PS_THREADS, 0, ctx->stream>>>(f.diff.p, n, ctx->d_bsum.p, ctx->I->d_slot2id.p, f.idcnt.p);
    ctx->span_end(sp);
    ctx->launches += 3;
    BK_CUDA(cudaGetLastError());
    f.folded = true;
    return BK_OK;
}

// stage 2: compaction = the KMC dump of one file (threshold, cap, the four stdout numbers)
static int stage_compact(bk_ctx* ctx, int slot) {
    FileState& f = ctx->file[slot];
    if (f.finalized) return BK_OK;
    const DerivedIndex& d = ctx->I->d;
    const u32 n_ids = (u32)d.id_kmer.size();
    // the novel part of the list: a kept k-mer stands for >= ci occurrences (list mode) / occupies a table slot
    const size_t novel_cap = f.list_mode ? std::min<u64>(f.nov.cap, std::max<u64>(f.nov_ub, 1)) / std::max<u32>(ctx->params.min_kmers, 1) + 1
                                         : (size_t)(1ull << f.gen_log2);
    const size_t out_cap = (size_t)n_ids + novel_cap;
    BK_CUDA(f.ckmers.reserve(out_cap)); BK_CUDA(f.ccounts.reserve(out_cap));
    int sp = ctx->span_begin(ST_FINALIZE);
    CompactArgs a;
    a.ci = ctx->params.min_kmers; a.cs = ctx->params.counter_max; a.rank = ctx->shard_rank; a.n_ranks = ctx->shard_n;
    a.out_kmers = f.ckmers.p; a.out_counts = f.ccounts.p; a.out_cap = (u32)std::min<size_t>(out_cap, 0xFFFFFFFFu);
    a.fc = &ctx->d_ctr.p->f[slot];
    if (!f.list_mode) {
        k_compact_ids<<<grid_for(ctx, n_ids, 256), 256, 0, ctx->stream>>>(a, f.idcnt.p, ctx->I->d_id_kmer.p, n_ids);
        k_compact_gen<<<grid_for(ctx, 1ull << f.gen_log2, 256), 256, 0, ctx->stream>>>(a, f.gen.p, (u32)(1ull << f.gen_log2));
        ctx->launches += 2;
    } else {
        // bins sized for the room the pushes were given (≈ 5 % of it is used at 0.2 % error: a few hundred k-mers per bin)
        BinView b;
        u32 lp = 6;                          // a round of a bin is sized for 1/32 of the room: 3 % of the bases novel
        while (lp < 14 && ((u64)BK_BIN_ROUND << lp) < f.nov_ub / 32) lp++;
        f.bin_log2p = lp;
        const u32 P = 1u << lp, G = (u32)ctx->sm_count * BK_BIN_G_PER_SM, PG = P * G;
        const u32 nb = (PG + BK_PS_BLOCK - 1) / BK_PS_BLOCK;
        BK_CUDA(f.bin_cnt.reserve((size_t)PG + 1));
        BK_CUDA(ctx->d_nov_sorted.reserve(std::max<size_t>(f.nov.cap, 1)));
        b.nov = f.nov.p; b.nov_n = &ctx->d_ctr.p->f[slot].nov_n; b.nov_cap = (u32)std::min<u64>(f.nov.cap, f.nov_limit);
        b.sorted = ctx->d_nov_sorted.p; b.cnt = f.bin_cnt.p; b.log2p = lp; b.G = G;
        b.exact = ctx->I->d_exact.p; b.exact_shift = 64 - d.exact_log2; b.exact_mask = (1u << d.exact_log2) - 1;
        b.slot2id = ctx->I->d_slot2id.p; b.idcnt = f.idcnt.p;
        if (!ablated("bins")) {
        k_bin_hist<<<G, BK_BIN_G_THREADS, P * 4, ctx->stream>>>(b);
        k_diff_blocksum<<<nb, BK_PS_THREADS, 0, ctx->stream>>>(f.bin_cnt.p, PG, ctx->d_bsum.p);
        k_diff_scan_bsum<<<1, BK_PS_THREADS, 0, ctx->stream>>>(ctx->d_bsum.p, nb);
        k_excl_apply<<<nb, BK_PS_THREADS, 0, ctx->stream>>>(f.bin_cnt.p, PG, ctx->d_bsum.p);
        k_bin_scatter<<<G, BK_BIN_G_THREADS, P * 4, ctx->stream>>>(b);
        if (!ablated("bincount")) k_bin_count<<<P, 256, BK_BIN_SMEM, ctx->stream>>>(b, a, &ctx->d_ctr.p->gen_full);
        }
        k_compact_ids<<<grid_for(ctx, n_ids, 256), 256, 0, ctx->stream>>>(a, f.idcnt.p, ctx->I->d_id_kmer.p, n_ids);   // after the bins: they add to idcnt
        ctx->launches += 7;
    }
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

// stages 3 + 4 in one pass per file for databases of at most four genomes (re-keyed table, not sharded): tallies and
// the pileups of all genomes together, selection, then the selected genome's arrays are moved to d_pile
static bool can_fuse_map(const bk_ctx* ctx) {
    return ctx->I->d.n_genomes <= 4 && ctx->I->d.rekeyed && !ctx->force_warp_map && ctx->shard_n == 1 && !ctx->no_fused_map;
}
static int stage_map_fused(bk_ctx* ctx) {
    ctx->use_level(1);
    const DerivedIndex& d = ctx->I->d;
    const MapView m = make_map_view(ctx);
    const u32 pile_stride = d.max_genome_rows * 4;
    const int n_files = n_files_used(ctx);
    Counters* dc = ctx->d_ctr.p;
    cudaStream_t st = ctx->stream;
    int sp = ctx->span_begin(ST_MAP);
    BK_CUDA(cudaMemsetAsync(ctx->d_pile_all.p, 0, (size_t)pile_stride * 4 * 4 * d.n_genomes, st));
    BK_CUDA(cudaMemsetAsync(ctx->d_pile.p, 0, (size_t)pile_stride * 4 * 4, st));
    for (int f = 0; f < n_files; f++) {
        FileState& fs = ctx->file[f];
        BK_CUDA(cudaMemsetAsync(fs.gstats.p, 0, (size_t)d.n_genomes * 16, st));
        const u32 ccap = (u32)std::min<size_t>(fs.ckmers.cap, 0xFFFFFFFFu);
        if (ablated("map")) {}
        else if (m.gslots) k_map_grp<2><<<ctx->sm_count * BK_STRIDE_WAVES, 256, 0, st>>>(m, fs.ckmers.p, fs.ccounts.p, &dc->f[f].n_counted, ccap, fs.gstats.p, nullptr, ctx->d_pile_all.p, pile_stride);
        else k_map_small<2, 1><<<ctx->sm_count * BK_STRIDE_WAVES, 256, 0, st>>>(m, fs.ckmers.p, fs.ccounts.p, &dc->f[f].n_counted, ccap, fs.gstats.p, nullptr, ctx->d_pile_all.p, pile_stride);
        ctx->launches++;
    }
    k_select<<<1, 32, 0, st>>>(ctx->file[0].gstats.p, ctx->file[n_files - 1].gstats.p, n_files, d.n_genomes, ctx->I->d_genome_len.p, dc);
    k_pile_pick<<<ctx->sm_count, 256, 0, st>>>(ctx->d_pile_all.p, ctx->d_pile.p, pile_stride, &dc->best);
    ctx->launches += 2;
    ctx->span_end(sp);
    BK_CUDA(cudaGetLastError());
    return BK_OK;
}

// stage 3: map_kmers tallies of every file (src/call.rs:1389-1430)
static int stage_map_stats(bk_ctx* ctx) {
    ctx->use_level(1);
    const DerivedIndex& d = ctx->I->d;
    const MapView m = make_map_view(ctx);
    const size_t map_smem = (size_t)d.n_genomes * 12 * 4;
    const bool small_db = d.n_genomes <= 4 && !ctx->force_warp_map;
    Counters* dc = ctx->d_ctr.p;
    cudaStream_t st = ctx->stream;
    int sp = ctx->span_begin(ST_MAP);
    for (int f = 0; f < n_files_used(ctx); f++) {
        FileState& fs = ctx->file[f];
        BK_CUDA(cudaMemsetAsync(fs.gstats.p, 0, (size_t)d.n_genomes * 16, st));
        const u32 ccap = (u32)std::min<size_t>(fs.ckmers.cap, 0xFFFFFFFFu);
        if (small_db && m.gslots) k_map_grp<0><<<ctx->sm_count * BK_STRIDE_WAVES, 256, 0, st>>>(m, fs.ckmers.p, fs.ccounts.p, &dc->f[f].n_counted, ccap, fs.gstats.p, nullptr, nullptr, 0);
        else if (small_db) (d.rekeyed ? k_map_small<0, 1> : k_map_small<0, 0>)<<<ctx->sm_count * BK_STRIDE_WAVES, 256, 0, st>>>(m, fs.ckmers.p, fs.ccounts.p, &dc->f[f].n_counted, ccap, fs.gstats.p, nullptr, nullptr, 0);
        else (d.rekeyed ? k_map<0, 1> : k_map<0, 0>)<<<ctx->sm_count * BK_STRIDE_WAVES, 256, map_smem, st>>>(m, fs.ckmers.p, fs.ccounts.p, &dc->f[f].n_counted, ccap, fs.gstats.p, nullptr, nullptr, 0);
        ctx->launches++;
    }
    ctx->span_end(sp);
    BK_CUDA(cudaGetLastError());
    return BK_OK;
}

// stage 4: pick_best_genome(_paired) + the selected genome's pileup (src/call.rs:1324-1385)
static int stage_select_pileup(bk_ctx* ctx) {
    ctx->use_level(1);
    const DerivedIndex& d = ctx->I->d;
    const MapView m = make_map_view(ctx);
    const bool small_db = d.n_genomes <= 4 && !ctx->force_warp_map;
    const u32 pile_stride = d.max_genome_rows * 4;
    const int n_files = n_files_used(ctx);
    Counters* dc = ctx->d_ctr.p;
    cudaStream_t st = ctx->stream;
    int sp = ctx->span_begin(ST_MAP);
    BK_CUDA(cudaMemsetAsync(ctx->d_pile.p, 0, (size_t)pile_stride * 4 * 4, st));
    k_select<<<1, 32, 0, st>>>(ctx->file[0].gstats.p, ctx->file[n_files - 1].gstats.p, n_files, d.n_genomes, ctx->I->d_genome_len.p, dc);
    ctx->launches++;
    for (int f = 0; f < n_files; f++) {
        FileState& fs = ctx->file[f];
        const u32 ccap = (u32)std::min<size_t>(fs.ckmers.cap, 0xFFFFFFFFu);
        if (small_db && m.gslots) k_map_grp<1><<<ctx->sm_count * BK_STRIDE_WAVES, 256, 0, st>>>(m, fs.ckmers.p, fs.ccounts.p, &dc->f[f].n_counted, ccap, nullptr, &dc->best, ctx->d_pile.p, pile_stride);
        else if (small_db) (d.rekeyed ? k_map_small<1, 1> : k_map_small<1, 0>)<<<ctx->sm_count * BK_STRIDE_WAVES, 256, 0, st>>>(m, fs.ckmers.p, fs.ccounts.p, &dc->f[f].n_counted, ccap, nullptr, &dc->best, ctx->d_pile.p, pile_stride);
        else (d.rekeyed ? k_map<1, 1> : k_map<1, 0>)<<<ctx->sm_count * BK_STRIDE_WAVES, 256, 0, st>>>(m, fs.ckmers.p, fs.ccounts.p, &dc->f[f].n_counted, ccap, nullptr, &dc->best, ctx->d_pile.p, pile_stride);
        ctx->launches++;
    }
    ctx->span_end(sp);
    BK_CUDA(cudaGetLastError());
    return BK_OK;
}

// stage 5: noise baseline + call_variants, read everything back, fill bk_sample_result
static int stage_score(bk_ctx* ctx, bk_sample_result* out) {
    ctx->use_level(2);
    const DerivedIndex& d = ctx->I->d;
    const int n_files = n_files_used(ctx);
    const u32 pile_stride = d.max_genome_rows * 4;
    Counters* dc = ctx->d_ctr.p;
    cudaStream_t st = ctx->stream;
    int sp = ctx->span_begin(ST_SCORE);
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
    std::vector<u32> hg[2];
    for (int f = 0; f < n_files; f++) {
        hg[f].resize((size_t)d.n_genomes * 4);
        BK_CUDA(cudaMemcpyAsync(hg[f].data(), ctx->file[f].gstats.p, hg[f].size() * 4, cudaMemcpyDeviceToHost, st));
    }
    BK_CUDA(cudaStreamSynchronize(st));
    const Counters& c = *ctx->h_ctr;
    ctx->finished = true;
    if (c.gen_full) return ctx->fail(BK_ERR_OVERFLOW, "no room left for novel k-mers (table / list / bin table); set bk_params.table_log2 higher");
    if (c.var_overflow) return ctx->fail(BK_ERR_OVERFLOW, "variant buffer overflow");
    for (int f = 0; f < n_files; f++) {
        if ((size_t)c.f[f].n_counted > ctx->file[f].ckmers.cap) return ctx->fail(BK_ERR_OVERFLOW, "counted k-mer list overflow");
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
        if (ctx->shard_n > 1) r.kmc[f] = ctx->shard_kmc[f];      // globally reduced numbers supplied by the host
        else {
            r.kmc[f].total_reads = ctx->file[f].total_reads; r.kmc[f].total_kmers = c.f[f].total_kmers;
            r.kmc[f].unique_kmers = c.f[f].unique; r.kmc[f].unique_counted = c.f[f].n_counted;
        }
    }
    cudaEventRecord(ctx->ev_end, st);
    if (c.best < 0) {
        if (out) *out = r;
        cudaEventSynchronize(ctx->ev_end);
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
        BK_CUDA(cudaMemcpyAsync(ctx->variants.data(), ctx->d_vars.p, (size_t)c.n_var * sizeof(bk_variant), cudaMemcpyDeviceToHost, st));
        cudaEventRecord(ctx->ev_end, st);
        BK_CUDA(cudaStreamSynchronize(st));
        std::sort(ctx->variants.begin(), ctx->variants.end(), [](const bk_variant& a, const bk_variant& b) {
            if (a.seq != b.seq) return a.seq < b.seq;
            if (a.pos != b.pos) return a.pos < b.pos;
            return a.alt_base < b.alt_base;
        });
    } else {
        cudaEventSynchronize(ctx->ev_end);
    }
    bk_stage_times& t = ctx->times;
    memset(&t, 0, sizeof t);
    float* acc[ST_N] = {&t.scan_ms, &t.leftover_ms, &t.finalize_ms, &t.map_ms, &t.score_ms};
    for (size_t i = 0; i < ctx->spans_used; i++) {
        float ms = 0;
        if (cudaEventElapsedTime(&ms, ctx->spans[i].a, ctx->spans[i].b) == cudaSuccess) *acc[ctx->spans[i].stage] += ms;
    }
    cudaEventElapsedTime(&t.total_ms, ctx->ev_begin, ctx->ev_end);
    t.launches = ctx->launches; t.scan_launches = ctx->scan_launches;
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

int bk_sample_finish(bk_ctx* ctx, bk_sample_result* out) {
    int rc = check_finish(ctx, "bk_sample_finish");
    if (rc) return rc;
    if (ctx->shard_n > 1) return ctx->fail(BK_ERR_ARG, "bk_sample_finish: context is in sharded mode; drive the bk_shard_* stages");
    for (int f = 0; f < n_files_used(ctx); f++) {
        if ((rc = stage_fold(ctx, f))) return rc;
        if ((rc = stage_compact(ctx, f))) return rc;
    }
    if (can_fuse_map(ctx)) { if ((rc = stage_map_fused(ctx))) return rc; }
    else {
        if ((rc = stage_map_stats(ctx))) return rc;
        if ((rc = stage_select_pileup(ctx))) return rc;
    }
    return stage_score(ctx, out);
}

// ---------------------------------------------------------------------------------------------
// read-sharded deep sample.  Every rank scans its share of the reads of ONE sample;
// k-mer counts are merged across ranks before the threshold / cap / max-pileup:
//   bk_shard_begin            per-rank partial counts are folded; novel k-mers are compacted and
//                             partitioned by owner rank
//   (host)                    all-reduce(SUM) the dense reference-k-mer counts; all-to-all the novel
//                             (k-mer, count) pairs to their owners
//   bk_shard_import_novel     the owner rebuilds its novel table from the merged pairs
//   bk_shard_map_stats        threshold + cap on the merged counts (each reference k-mer id and each
//                             novel k-mer is owned by exactly one rank), per-genome tallies
//   (host)                    all-reduce(SUM) the tallies and the KMC numbers
//   bk_shard_select_pileup    selection (same on every rank) + this rank's pileup contribution
//   (host)                    all-reduce(MAX) the two depth arrays, all-reduce(SUM) the two support arrays
//   bk_shard_score            noise + variants (every rank gets the same result)
// ---------------------------------------------------------------------------------------------
int bk_shard_config(bk_ctx* ctx, uint32_t rank, uint32_t n_ranks) {
    if (!ctx) return BK_ERR_ARG;
    if (ctx->in_sample && !ctx->finished) return ctx->fail(BK_ERR_ARG, "bk_shard_config: call between samples");
    if (n_ranks == 0 || rank >= n_ranks || n_ranks > 1024) return ctx->fail(BK_ERR_ARG, "bk_shard_config: bad rank / n_ranks");
    ctx->shard_rank = rank; ctx->shard_n = n_ranks;
    return BK_OK;
}

int bk_shard_begin(bk_ctx* ctx, int file_slot, void** d_ref_counts, uint64_t* n_ref_counts, void** d_novel_kmers,
                   void** d_novel_counts, uint64_t* part_off) {
    int rc = check_finish(ctx, "bk_shard_begin");
    if (rc) return rc;
    if (file_slot < 0 || file_slot > 1 || !ctx->file[file_slot].used) return ctx->fail(BK_ERR_ARG, "bk_shard_begin: file slot %d has no reads", file_slot);
    if (!d_ref_counts || !n_ref_counts || !d_novel_kmers || !d_novel_counts || !part_off) return ctx->fail(BK_ERR_ARG, "bk_shard_begin: null argument");
    FileState& f = ctx->file[file_slot];
    if ((rc = stage_fold(ctx, file_slot))) return rc;
    const u32 n_slots = (u32)(1ull << f.gen_log2);
    const u32 nr = ctx->shard_n;
    BK_CUDA(ctx->d_part.reserve((size_t)nr * 2));
    BK_CUDA(cudaMemsetAsync(ctx->d_part.p, 0, (size_t)nr * 2 * 4, ctx->stream));
    k_novel_partition<0><<<grid_for(ctx, n_slots, 256), 256, nr * 4, ctx->stream>>>(f.gen.p, n_slots, nr, ctx->d_part.p, nullptr, nullptr, nullptr);
    std::vector<u32> cnt(nr);
    BK_CUDA(cudaMemcpyAsync(cnt.data(), ctx->d_part.p, nr * 4, cudaMemcpyDeviceToHost, ctx->stream));
    BK_CUDA(cudaStreamSynchronize(ctx->stream));
    std::vector<u32> start(nr);
    u64 total = 0;
    for (u32 r = 0; r < nr; r++) { part_off[r] = total; start[r] = (u32)total; total += cnt[r]; }
    part_off[nr] = total;
    BK_CUDA(f.xk.reserve(total + 1)); BK_CUDA(f.xc.reserve(total + 1));
    BK_CUDA(cudaMemsetAsync(ctx->d_part.p, 0, (size_t)nr * 4, ctx->stream));
    BK_CUDA(cudaMemcpyAsync(ctx->d_part.p + nr, start.data(), nr * 4, cudaMemcpyHostToDevice, ctx->stream));
    k_novel_partition<1><<<grid_for(ctx, n_slots, 256), 256, nr * 4, ctx->stream>>>(f.gen.p, n_slots, nr, ctx->d_part.p, ctx->d_part.p + nr, f.xk.p, f.xc.p);
    ctx->launches += 2;
    BK_CUDA(cudaGetLastError());
    BK_CUDA(cudaStreamSynchronize(ctx->stream));
    *d_ref_counts = f.idcnt.p; *n_ref_counts = ctx->I->d.id_kmer.size();
    *d_novel_kmers = f.xk.p; *d_novel_counts = f.xc.p;
    return BK_OK;
}

int bk_shard_import_novel(bk_ctx* ctx, int file_slot, const void* d_kmers, const void* d_counts, uint64_t n) {
    int rc = check_finish(ctx, "bk_shard_import_novel");
    if (rc) return rc;
    if (file_slot < 0 || file_slot > 1 || !ctx->file[file_slot].used) return ctx->fail(BK_ERR_ARG, "bk_shard_import_novel: bad file slot");
    FileState& f = ctx->file[file_slot];
    u32 lg = 10;
    while ((1ull << lg) < 2 * n + 16) lg++;
    if (lg > 31) return ctx->fail(BK_ERR_OVERFLOW, "too many novel k-mers for one rank");
    f.gen_log2 = lg;
    BK_CUDA(f.gen.reserve(1ull << lg));
    k_gen_init<<<grid_for(ctx, 1ull << lg, 256 * 8), 256, 0, ctx->stream>>>(f.gen.p, 1ull << lg);
    if (n) k_novel_insert<<<grid_for(ctx, n, 256), 256, 0, ctx->stream>>>(f.gen.p, 64 - lg, (u32)((1ull << lg) - 1), (const u64*)d_kmers, (const u32*)d_counts, n,
                                                                           &ctx->d_ctr.p->gen_full);
    ctx->launches += 2;
    BK_CUDA(cudaGetLastError());
    BK_CUDA(cudaStreamSynchronize(ctx->stream));     // the caller may free / reuse its buffers now
    return BK_OK;
}

int bk_shard_map_stats(bk_ctx* ctx, void** d_tallies0, void** d_tallies1, uint64_t* n_tallies, bk_kmc_stats* partial) {
    int rc = check_finish(ctx, "bk_shard_map_stats");
    if (rc) return rc;
    const int n_files = n_files_used(ctx);
    for (int f = 0; f < n_files; f++) {
        if (!ctx->file[f].folded) return ctx->fail(BK_ERR_ARG, "bk_shard_map_stats: call bk_shard_begin for file %d first", f);
        if ((rc = stage_compact(ctx, f))) return rc;
    }
    if ((rc = stage_map_stats(ctx))) return rc;
    BK_CUDA(cudaMemcpyAsync(ctx->h_ctr, ctx->d_ctr.p, sizeof(Counters), cudaMemcpyDeviceToHost, ctx->stream));
    BK_CUDA(cudaStreamSynchronize(ctx->stream));
    if (ctx->h_ctr->gen_full) return ctx->fail(BK_ERR_OVERFLOW, "novel k-mer table is full; set bk_params.table_log2 higher");
    for (int f = 0; f < 2; f++) {
        bk_kmc_stats z; memset(&z, 0, sizeof z);
        if (f < n_files) {
            z.total_reads = ctx->file[f].total_reads; z.total_kmers = ctx->h_ctr->f[f].total_kmers;
            z.unique_kmers = ctx->h_ctr->f[f].unique; z.unique_counted = ctx->h_ctr->f[f].n_counted;
        }
        if (partial) partial[f] = z;
    }
    if (d_tallies0) *d_tallies0 = ctx->file[0].gstats.p;
    if (d_tallies1) *d_tallies1 = n_files > 1 ? ctx->file[1].gstats.p : nullptr;
    if (n_tallies) *n_tallies = (u64)ctx->I->d.n_genomes * 4;
    return BK_OK;
}

int bk_shard_select_pileup(bk_ctx* ctx, const bk_kmc_stats* global_kmc, void** d_pile, uint64_t* n_per_array) {
    int rc = check_finish(ctx, "bk_shard_select_pileup");
    if (rc) return rc;
    if (global_kmc) { ctx->shard_kmc[0] = global_kmc[0]; ctx->shard_kmc[1] = global_kmc[1]; }
    if ((rc = stage_select_pileup(ctx))) return rc;
    BK_CUDA(cudaStreamSynchronize(ctx->stream));
    if (d_pile) *d_pile = ctx->d_pile.p;
    if (n_per_array) *n_per_array = (u64)ctx->I->d.max_genome_rows * 4;
    return BK_OK;
}

int bk_shard_score(bk_ctx* ctx, bk_sample_result* out) {
    int rc = check_finish(ctx, "bk_shard_score");
    if (rc) return rc;
    return stage_score(ctx, out);
}

int bk_stage_times_get(bk_ctx* ctx, bk_stage_times* out) {
    if (!ctx || !out) return BK_ERR_ARG;
    *out = ctx->times;
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
    BK_CUDA(cudaMemcpyAsync(tmp.data(), ctx->d_pile.p + (size_t)arr * ctx->I->d.max_genome_rows * 4, tmp.size() * 4, cudaMemcpyDeviceToHost, ctx->stream));
    BK_CUDA(cudaStreamSynchronize(ctx->stream));
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
    BK_CUDA(cudaMemcpyAsync(out, ctx->d_noise.p, (size_t)rows * 8, cudaMemcpyDeviceToHost, ctx->stream));
    BK_CUDA(cudaStreamSynchronize(ctx->stream));
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
    BK_CUDA(cudaMemcpyAsync(hk.data(), ctx->file[slot].ckmers.p, have * 8, cudaMemcpyDeviceToHost, ctx->stream));
    BK_CUDA(cudaMemcpyAsync(hc.data(), ctx->file[slot].ccounts.p, have * 4, cudaMemcpyDeviceToHost, ctx->stream));
    BK_CUDA(cudaStreamSynchronize(ctx->stream));
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
