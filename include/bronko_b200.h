/* bronko_b200.h — C ABI of libbronko_b200.so: the B200 (sm_100a) k-mer→pileup path of bronko.
 *
 * The reference (treangenlab/bronko v0.1.0, Rust) has no FFI/plugin interface: the hot path sits
 * behind an internal Rust seam inside call::call (src/call.rs:151-402).  Each entry point below
 * names the reference function(s) it replaces (paths relative to the reference tree).  A Rust
 * `extern "C"` crate binds these unchanged (INTEGRATION.md shows the stub); in this image (no
 * rustc) the callers are the C++ `bronko` CLI (bronko_b200/csrc/bronko_main.cpp) and the Python
 * ctypes mirror (bronko_b200/api.py).
 *
 * Conventions: every call returns 0 on success or a negative bk_status; nothing ever exits or
 * throws across the ABI (the reference logs `error!` and calls std::process::exit(1); the host maps
 * a negative status to exactly that).  bk_last_error() gives the message.  The caller owns every
 * input buffer (valid until the call returns); the library owns all device memory.  One bk_ctx per
 * GPU, not thread-safe, samples are sequential per ctx (mirrors the sequential sample loop at
 * src/call.rs:213,298).  There is NO CPU fallback: bk_create fails if no sm_100 device is usable.
 */
#ifndef BRONKO_B200_H
#define BRONKO_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct bk_ctx bk_ctx;

typedef enum {
    BK_OK = 0,
    BK_ERR_ARG = -1,      /* invalid argument / call order */
    BK_ERR_CUDA = -2,     /* CUDA runtime error (message has the cudaError string) */
    BK_ERR_IO = -3,       /* file could not be opened / parsed */
    BK_ERR_NO_GENOME = -4,/* pick_best_genome returned None (src/call.rs:230-233): host must exit(1) */
    BK_ERR_OVERFLOW = -5, /* a device table overflowed its capacity; re-run with a larger table */
    BK_ERR_NO_DEVICE = -6,
    BK_ERR_NOMEM = -7     /* host allocation failed */
} bk_status;

/* #[repr(C)] BucketInfo, src/build.rs:52-60 — 12 bytes including padding. */
typedef struct {
    uint16_t file_id;
    uint8_t seq_id;
    uint8_t _pad0;
    uint32_t location;
    uint8_t idx;
    uint8_t canonical;
    uint8_t _pad1[2];
} bk_bucket_info;

/* The CallArgs fields the path consumes (src/cli.rs:61-166, defaults src/consts.rs:1-20). */
typedef struct {
    uint32_t k;                       /* -k / --kmer-size (odd, 15..31) */
    uint32_t min_kmers;               /* --min-kmers: KMC -ci */
    uint32_t counter_max;             /* KMC -cs (src/call.rs:1173: 1000000) */
    uint32_t use_full_kmer;           /* --use-full-kmer */
    uint32_t n_fixed;                 /* --n-fixed */
    uint32_t no_end_filter;           /* --no-end-filter */
    uint32_t no_strand_filter;        /* --no-strand-filter */
    uint32_t no_strand_balance_filter;/* --no-strand-balance-filter */
    uint32_t n_per_strand;            /* --n-per-strand */
    uint32_t table_log2;              /* log2 slots of the novel-k-mer hash table; 0 = auto */
    uint64_t min_depth;               /* --min-depth */
    uint64_t min_variant_depth;       /* --min-variant-depth */
    double min_af;                    /* --min-af */
    double strand_balance_ratio;      /* --balance-ratio */
    double strand_odds_max;           /* --strand_odds */
    double variant_multiplier;        /* --noise-multiplier */
} bk_params;

/* The four numbers bronko parses from KMC's stdout (src/call.rs:1190-1200). */
typedef struct {
    uint64_t total_reads, total_kmers, unique_kmers, unique_counted;
} bk_kmc_stats;

/* Value of FxHashMap<u16,(usize,usize,usize)> returned by map_kmers (src/call.rs:1257). */
typedef struct {
    uint64_t perfect, variant, unique_perfect;
    uint32_t present;                 /* genome has an entry in the reference's map (>=1 bucket hit) */
    uint32_t _pad;
} bk_genome_stats;

/* VCFRecord, src/call.rs:776-789 (seq = index of the sequence inside the selected genome). */
typedef struct {
    uint32_t seq;
    uint32_t pos;                     /* 1-based */
    uint8_t ref_base, alt_base;
    uint8_t _pad[6];
    uint64_t fwd_ref, rev_ref, fwd_alt, rev_alt, depth;
    double af, sor;
} bk_variant;

/* What call() keeps per sample: OutputInfo (src/call.rs:138-149) + the call_variants tuple. */
typedef struct {
    int32_t best_genome;              /* index into the db's files */
    uint32_t n_files;                 /* 1 = single-end, 2 = paired */
    uint64_t n_variants;
    uint64_t num_major_variants, num_minor_variants;
    double breadth_coverage, depth_coverage;
    uint64_t num_perfect_kmers, num_variant_kmers, num_unmapped_kmers;
    bk_kmc_stats kmc[2];
} bk_sample_result;

/* Per-stage device time of the last sample (CUDA events on the ctx stream), milliseconds. */
typedef struct {
    float scan_ms;      /* pack + seed/extend scan kernel(s)        */
    float leftover_ms;  /* novel / mismatching k-mer counting        */
    float finalize_ms;  /* prefix-sum, fold, compaction              */
    float map_ms;       /* bucket lookup, stats and pileup           */
    float score_ms;     /* select + noise + variant kernels          */
    float total_ms;     /* bk_sample_begin → end of bk_sample_finish */
    uint32_t launches;  /* kernels launched for this sample          */
    uint32_t scan_launches;
    float coll_ms;      /* read-sharded sample: collectives (all-reduces, size all-gather, pair all-to-all) */
    uint32_t coll_calls;
    float decode_ms;    /* bk_reads_push_fastq*: H2D of the file, inflate, FASTQ parse (device time)         */
} bk_stage_times;

/* What the decode stage did with the last file pushed to a slot by bk_reads_push_fastq / _fastq_mem. */
typedef struct {
    uint32_t mode;      /* 1 plain text, 2 gzip inflated by zlib on the host, 3 BGZF inflated by the GPU's decompression engine */
    uint32_t segments;  /* device passes (long files are decoded a segment of text at a time)                */
    uint64_t compressed_bytes, text_bytes, n_reads, n_bases;
} bk_decode_info;

/* ---- context ---------------------------------------------------------------------------- */
int bk_create(bk_ctx** out, int device);
void bk_destroy(bk_ctx* ctx);
const char* bk_last_error(bk_ctx* ctx);      /* ctx may be NULL: last bk_create error */
void* bk_stream(bk_ctx* ctx);                /* = bk_stream_slot(ctx, 0) */
void* bk_stream_slot(bk_ctx* ctx, int file_slot);
                                             /* cudaStream_t the pushes (counting kernels) of one file slot are enqueued on:
                                              * order device buffers handed to bk_reads_push_device against it.  The two
                                              * files of a pair run side by side on their own streams; later stages of a
                                              * sample run on internal streams of higher priority chained to them by
                                              * events; every call that returns results synchronises. */
const char* bk_version(void);

/* ---- index: replaces the bincode decode at src/call.rs:179-200 / build_indexes at 170-178 ---- */
/* From flat arrays (what a Rust host holds after decoding BronkoIndex, src/build.rs:23-60):
 * keys[n_keys], entry_off[n_keys+1], entries[entry_off[n_keys]] in per-key order;
 * genome g owns sequences [genome_seq_off[g], genome_seq_off[g+1]); sequence s has seq_len[s] bases
 * at ref_bases + seq_base_off[s] (raw FASTA bytes, SeqMeta.seq). */
int bk_index_load(bk_ctx* ctx, uint32_t k, uint64_t n_keys, const uint64_t* keys,
                  const uint64_t* entry_off, const bk_bucket_info* entries, uint32_t n_genomes,
                  const uint32_t* genome_seq_off, const uint64_t* seq_len,
                  const uint64_t* seq_base_off, const uint8_t* ref_bases);
/* .bkdb file (bincode 2 standard config). */
int bk_index_load_file(bk_ctx* ctx, const char* bkdb_path);
/* build_indexes (src/build.rs:145-231) from FASTA(.gz) files, then load it. */
int bk_index_build(bk_ctx* ctx, uint32_t k, uint32_t n_files, const char* const* fasta_paths);
/* save_index (src/build.rs:122-143); keys are written in ascending order. */
int bk_index_save(bk_ctx* ctx, const char* bkdb_path);
/* Use the index already loaded into `owner` (same device) without copying it: the contexts of one GPU — one per
 * sample in flight — read one set of tables.  The reference holds one index per process (src/call.rs:170-200). */
int bk_index_share(bk_ctx* ctx, bk_ctx* owner);
int bk_index_info(bk_ctx* ctx, uint32_t* k, uint64_t* n_keys, uint64_t* n_entries, uint32_t* n_genomes);
const char* bk_genome_name(bk_ctx* ctx, uint32_t genome);
uint32_t bk_genome_n_seqs(bk_ctx* ctx, uint32_t genome);
const char* bk_seq_name(bk_ctx* ctx, uint32_t genome, uint32_t seq);
uint64_t bk_seq_len(bk_ctx* ctx, uint32_t genome, uint32_t seq);
const uint8_t* bk_seq_bases(bk_ctx* ctx, uint32_t genome, uint32_t seq);
/* parity hook: the decoded map, keys ascending (same layout as bk_index_load's inputs). */
int bk_index_export(bk_ctx* ctx, uint64_t* keys, uint64_t* entry_off, bk_bucket_info* entries);

/* ---- one sample: replaces src/call.rs:213-292 (SE) / 298-387 (PE) ------------------------- */
void bk_params_default(bk_params* p);                       /* src/consts.rs defaults, k = 21 */
int bk_sample_begin(bk_ctx* ctx, const bk_params* params);
/* get_kmers / count_kmers_kmc (src/call.rs:630-646, 1152-1226): feed decoded reads of one file.
 * file_slot 0 = the -r file or R1, 1 = R2 (counted and thresholded separately, src/call.rs:302-307).
 * bases = concatenated sequence lines (ASCII), read r = bases[read_off[r] .. read_off[r+1]).
 * Host buffers (pinned recommended: bk_host_alloc); may be called repeatedly per file (chunks).  Both buffers have been
 * copied when the call returns: the caller may overwrite them at once. */
int bk_reads_push(bk_ctx* ctx, int file_slot, const uint8_t* bases, const uint32_t* read_off,
                  uint64_t n_reads);
/* Same, buffers already in device memory: d_bases 16-byte aligned and readable for 64 bytes past n_bases (the scan
 * loads whole words around a read; bk_fastq_decode chunks carry that slack).  The kernels read the buffers
 * asynchronously: they must stay valid and unmodified until bk_sample_finish returns (or the context's stream,
 * bk_stream(), has been synchronised). */
int bk_reads_push_device(bk_ctx* ctx, int file_slot, const uint8_t* d_bases,
                         const uint32_t* d_read_off, uint64_t n_reads, uint64_t n_bases,
                         uint32_t max_read_len);
/* 2-bit packed reads — a quarter of the bytes across PCIe, which is what bounds a sample pushed from host memory (the
 * GPU path runs ~6x faster than 16 PCIe Gen5 lanes deliver ASCII).  packed: 16 bases per u32, base i of the push at
 * bits 2 * (i % 16) of word i / 16, codes A0 C1 G2 T3; reads contiguous in base space, read r = bases
 * [read_off[r], read_off[r + 1]).  Only reads made of ACGT / acgt can be packed (KMC counts lower case like upper case
 * and splits reads at any other byte): bk_reads_pack — a host helper of the decode stage — splits a chunk into its
 * packable reads and the rest, which goes through bk_reads_push as ASCII.  Counting does not depend on the order of the
 * reads, so the two pushes together equal the ASCII push of the chunk.  Buffers of bk_reads_pack are the caller's:
 * packed >= n_bases / 16 + 2 words, packed_off and rest_off >= n_reads + 1 entries, rest_bases >= n_bases + 64 bytes. */
int bk_reads_push_packed(bk_ctx* ctx, int file_slot, const uint32_t* packed, const uint32_t* read_off, uint64_t n_reads);
int bk_reads_pack(const uint8_t* bases, const uint32_t* read_off, uint64_t n_reads, uint32_t* packed, uint32_t* packed_off,
                  uint64_t* n_packed, uint8_t* rest_bases, uint32_t* rest_off, uint64_t* n_rest);
/* A FASTQ(.gz) file (KMC reader contract) decoded ON THE DEVICE and pushed: the host only moves
 * bytes.  BGZF files (bgzip / htslib: independent <= 64 KiB gzip members) are inflated by the B200's hardware
 * decompression engine; other gzip files by zlib on the host (a single deflate stream cannot be split); the text is
 * parsed by kernels either way.  _mem takes the file's bytes from host memory (pinned: bk_host_alloc — copied
 * asynchronously; pageable works).  bk_decode_info_get tells which way the last file of a slot went. */
int bk_reads_push_fastq(bk_ctx* ctx, int file_slot, const char* fastq_path);
int bk_reads_push_fastq_mem(bk_ctx* ctx, int file_slot, const uint8_t* file_bytes, uint64_t n_bytes);
int bk_decode_info_get(bk_ctx* ctx, int file_slot, bk_decode_info* out);
/* The host decode stage on its own — the step in front of the path (KMC's FASTQ reader inside count_kmers_kmc,
 * src/call.rs:1166-1181; contract in SURVEY.md Appendix B: 4-line records, only the sequence line is used, gz detected).
 * Needs no context and no GPU and is thread-safe: decode the files of the next samples on other host threads while
 * the GPU works (inflate runs at a few hundred MB/s per thread, the path at hundreds of GB/s).  Decoded reads are
 * chunks of at most 2^30 bases (offsets are u32), each with 64 readable bytes past its last base. */
typedef struct bk_reads bk_reads;
int bk_fastq_decode(const char* fastq_path, bk_reads** out, char* err, uint64_t err_cap);
uint64_t bk_reads_n_chunks(const bk_reads* reads);
int bk_reads_chunk(const bk_reads* reads, uint64_t i, const uint8_t** bases, const uint32_t** read_off,
                   uint64_t* n_reads, uint64_t* n_bases);
int bk_reads_push_decoded(bk_ctx* ctx, int file_slot, const bk_reads* reads);   /* every chunk, then returns */
void bk_reads_free(bk_reads* reads);
/* map_kmers ×n_files → pick_best_genome(_paired) → call_variants (src/call.rs:224-268 / 314-360).
 * Returns BK_ERR_NO_GENOME where the reference exits with "Unable to pick a best genome". */
int bk_sample_finish(bk_ctx* ctx, bk_sample_result* out);
/* Results of the last finished sample. */
int bk_sample_result_get(bk_ctx* ctx, bk_sample_result* out);               /* what bk_sample_finish returned */
int bk_sample_variants(bk_ctx* ctx, bk_variant* out, uint64_t cap);          /* sorted seq,pos,alt */
int bk_sample_genome_stats(bk_ctx* ctx, int file_slot, bk_genome_stats* out);/* n_genomes entries */
/* OutputData.counts of the selected genome (src/call.rs:1235-1239), widened to u64:
 * arr 0 = output (fwd depth), 1 = output_rev, 2 = output_counts (fwd support), 3 = output_rev_counts;
 * out = rows*4 u64 where rows = sum of the genome's sequence lengths. */
int bk_sample_pileup(bk_ctx* ctx, int arr, uint64_t* out, uint64_t cap_rows);
int bk_sample_noise_max(bk_ctx* ctx, double* out, uint64_t cap_rows);        /* Noise.max per row */
/* parity hook: the KMC dump of one file (kept k-mers, capped counts), ascending by k-mer.
 * Pass NULL pointers to query the count. */
int bk_kmer_counts_get(bk_ctx* ctx, int file_slot, uint64_t* kmers, uint32_t* counts, uint64_t* n);
int bk_stage_times_get(bk_ctx* ctx, bk_stage_times* out);
/* Per-stage CUDA events on or off (on when a context is created).  Off: a sample records only its begin / end events —
 * ~20 fewer driver calls per sample on the host thread, which is what a GPU that is fed by several contexts of one
 * process runs out of first; bk_stage_times then holds total_ms and the launch counts only.  No effect on results. */
int bk_stage_timing(bk_ctx* ctx, int on);

/* ---- writers: src/call.rs:648-774 (formats in SURVEY.md Appendix D) ------------------------ */
int bk_write_vcf(bk_ctx* ctx, const char* reads_path, const char* out_path);
int bk_write_pileup(bk_ctx* ctx, const char* out_path);
/* clean_sample_id (src/util.rs:30-50) → buf; returns needed size incl. NUL */
uint64_t bk_clean_sample_id(const char* path, char* buf, uint64_t cap);

/* ---- read-sharded deep sample (BASELINE config C3) -----------------------------
 * Every rank scans its share of the reads of ONE sample with bk_reads_push* (every rank must use the same file
 * slots; a push of zero reads counts).  The pileup is a MAX over globally summed, thresholded (>= min_kmers),
 * saturated counts (src/call.rs:1172-1173, 1341-1345; R1 / R2 separately, 302-317), so the library merges the k-mer
 * counts across ranks BEFORE the cut-offs — all-reduce(SUM) of the dense reference-k-mer counts, all-to-all of the
 * novel (k-mer, partial count) pairs to an owner rank — lets every k-mer be thresholded and mapped by exactly one
 * rank, and only then combines depth (MAX), support and tallies (SUM).  bk_sample_finish does all of it, collectives
 * included (NCCL, enqueued on the context's stream), must be called by every rank, and returns the same result on
 * every rank.  One rank per process and GPU:
 *   rank 0: bk_shard_unique_id(id) ; the host hands the 128 bytes to every rank (MPI / torch.distributed / a file) ;
 *   every rank: bk_shard_init(ctx, rank, n_ranks, id), between samples ; n_ranks <= 1 returns to whole samples.
 * libnccl.so.2 is opened when the first of these is called; a process that never shards does not need it. */
int bk_shard_unique_id(uint8_t* out128);
int bk_shard_init(bk_ctx* ctx, uint32_t rank, uint32_t n_ranks, const uint8_t* id128);
int bk_shard_info(bk_ctx* ctx, uint32_t* rank, uint32_t* n_ranks);
/* The same sharded path with all ranks in ONE process on ONE device (NCCL refuses two ranks on a device): contexts
 * ctxs[0..n) sharing one index (bk_index_share) become ranks 0..n-1, every one gets its share of the reads, and
 * bk_shard_finish_local runs the sharded finish for all of them (collectives are kernels / copies).  The single-GPU
 * parity tests of the sharded kernels use this; n = 1 dissolves the group. */
int bk_shard_local(bk_ctx** ctxs, uint32_t n);
int bk_shard_finish_local(bk_ctx** ctxs, uint32_t n, bk_sample_result* out);

/* ---- pinned host memory helpers ---------------------------------------------------------- */
void* bk_host_alloc(uint64_t bytes);
void bk_host_free(void* p);

#ifdef __cplusplus
}
#endif
#endif /* BRONKO_B200_H */
