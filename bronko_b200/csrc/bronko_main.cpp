// bronko_main.cpp — the `bronko` host CLI on top of the C ABI (include/bronko_b200.h).
//
// Mirrors the reference's command line (treangenlab/bronko src/cli.rs:15-166, defaults src/consts.rs)
// and the driver loop call::call (src/call.rs:151-402): `bronko build` writes a .bkdb, `bronko call`
// processes single-end files then pairs sequentially and writes <stem>.vcf, <stem>.tsv (--pileup),
// bronko_overview.tsv and <genome>.mfa (--alignment).  Everything between "reads decoded" and "variant
// records" runs on the GPU through libbronko_b200.so; a negative status is turned into the
// reference's behaviour (error line + exit code 1).  The reference is Rust; there is no Rust toolchain
// in this image, so this host is C++ and touches the device code only through the C ABI a Rust FFI
// crate would bind.
#include <sys/stat.h>
#include <unistd.h>

#include <chrono>
#include <deque>
#include <future>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <set>
#include <string>
#include <thread>
#include <vector>

#include "bk_host.h"

static const char* VERSION = "0.1.0";

[[noreturn]] static void die(const std::string& msg) {
    fprintf(stderr, "ERROR [bronko] %s\n", msg.c_str());
    exit(1);
}
static void info(const std::string& msg) { fprintf(stderr, "INFO  [bronko] %s\n", msg.c_str()); }
static void warn(const std::string& msg) { fprintf(stderr, "WARN  [bronko] %s\n", msg.c_str()); }

static bool ends_with(const std::string& s, const char* suf) {
    const size_t m = strlen(suf);
    return s.size() >= m && s.compare(s.size() - m, m, suf) == 0;
}
// src/util.rs:4-28
static bool check_fastq(const std::string& f) {
    return ends_with(f, ".fq") || ends_with(f, ".fastq") || ends_with(f, ".fq.gz") || ends_with(f, "fastq.gz") || ends_with(f, "fnq") || ends_with(f, "fnq.gz");
}
static bool check_fasta(const std::string& f) {
    return ends_with(f, ".fa") || ends_with(f, ".fasta") || ends_with(f, ".fa.gz") || ends_with(f, "fasta.gz") || ends_with(f, "fna") || ends_with(f, "fna.gz");
}

struct Args {
    std::string mode;
    std::vector<std::string> genomes, reads, first_pairs, second_pairs;
    std::string db, output;
    bool have_genomes = false, have_db = false;
    long kmer = 21, min_kmers = 3, n_fixed = 2, n_per_strand = 2, min_depth = 300, min_variant_depth = 3, threads = 4;
    double min_af = 0.03, balance_ratio = 0.1, strand_odds = 6.0, noise_multiplier = 1.5;
    bool use_full_kmer = false, no_end_filter = false, no_strand_filter = false, no_strand_balance_filter = false;
    bool pileup = false, alignment = false, keep_kmer_info = false, debug = false, verbose = false;
    int device = 0;
    long table_log2 = 0;
};

static void usage(const char* mode) {
    if (!mode || !strcmp(mode, "top"))
        printf("Usage: bronko <COMMAND>\n\nCommands:\n  build  Create an bronko index of existing viral references for a given species\n"
               "  call   Perform rapid viral variant calling of viral sequencing data\n  help   Print this message\n\nOptions:\n  -h, --help     Print help\n  -V, --version  Print version\n");
    else if (!strcmp(mode, "build"))
        printf("Usage: bronko build [OPTIONS]\n\nREFERENCE INPUT:\n  -g, --genomes <GENOMES>...  Genome files to be built into index (fasta/gzip)\n\nKMER:\n"
               "  -k, --kmer-size <KMER>      Kmer size [default: 21]\n\nOUTPUT:\n  -o, --output <OUTPUT>       Name of index file (.bkdb will be added) [default: bronko]\n\n"
               "Options:\n  -t, --threads <THREADS>     Number of threads [default: 4]\n      --debug                 Debug output\n      --verbose               Verbose output\n");
    else
        printf("Usage: bronko call [OPTIONS]\n\nREFERENCE INPUT:\n  -g, --genomes <GENOMES>...   Genome fasta(.gz) files to use as references (bronko build will be called)\n"
               "  -d, --db <DB>                Use a prebuilt bronko db (.bkdb) of genomes of interest\n\nREADS INPUT:\n  -r, --reads <READS>...       Input single-end reads (fastq/gzip)\n"
               "  -1, --first-pairs <..>...    First pairs for raw paired-end reads (fastq/gzip)\n  -2, --second-pairs <..>...   Second pairs for raw paired-end reads (fastq/gzip)\n\nALGORITHM:\n"
               "  -k, --kmer-size <KMER>       Kmer size used for analysis [default: 21]\n      --min-kmers <N>          Minimum times a kmer must occur in sequencing data to be used [default: 3]\n"
               "      --use-full-kmer          Use the entire kmer length for variant positions\n      --n-fixed <N>            Number of fixed positions at the end of each kmer [default: 2]\n\n"
               "VARIANT CALLING PARAMETERS:\n      --min-af <F> [0.03]  --no-end-filter  --no-strand-filter  --no-strand-balance-filter\n"
               "      --balance-ratio <F> [0.1]  --n-per-strand <N> [2]  --strand_odds <F> [6]  --min-depth <N> [300]\n      --min-variant-depth <N> [3]  --noise-multiplier <F> [1.5]\n\n"
               "OUTPUT:\n  -o, --output <DIR>           Folder to output all resulting files [default: bronko_output]\n      --pileup  --alignment  --keep-kmer-info\n\n"
               "Options:\n  -t, --threads <THREADS>      Number of threads [default: 4]\n      --device <N>             CUDA device to run on [default: 0]\n      --table-log2 <N>         log2 slots of the novel k-mer table (0 = auto)\n      --debug  --verbose\n");
}

static Args parse(int argc, char** argv) {
    Args a;
    if (argc < 2) { usage("top"); exit(2); }
    a.mode = argv[1];
    if (a.mode == "-h" || a.mode == "--help" || a.mode == "help") { usage("top"); exit(0); }
    if (a.mode == "-V" || a.mode == "--version") { printf("bronko %s\n", VERSION); exit(0); }
    if (a.mode != "build" && a.mode != "call") { fprintf(stderr, "error: unrecognized subcommand '%s'\n", a.mode.c_str()); exit(2); }
    if (argc == 2) { usage(a.mode.c_str()); exit(2); }                 // arg_required_else_help
    a.output = a.mode == "build" ? "bronko" : "bronko_output";
    auto is_flag = [](const char* s) { return s[0] == '-' && s[1] != 0 && !(s[1] >= '0' && s[1] <= '9' && s[2] == 0 && false); };
    for (int i = 2; i < argc; i++) {
        std::string f = argv[i];
        auto multi = [&](std::vector<std::string>& dst) {
            int n = 0;
            while (i + 1 < argc && !(argv[i + 1][0] == '-' && strlen(argv[i + 1]) > 1)) { dst.push_back(argv[++i]); n++; }
            if (!n) { fprintf(stderr, "error: a value is required for '%s' but none was supplied\n", f.c_str()); exit(2); }
        };
        auto val = [&]() -> std::string {
            if (i + 1 >= argc) { fprintf(stderr, "error: a value is required for '%s' but none was supplied\n", f.c_str()); exit(2); }
            return argv[++i];
        };
        (void)is_flag;
        if (f == "-h" || f == "--help") { usage(a.mode.c_str()); exit(0); }
        else if (f == "-g" || f == "--genomes") { multi(a.genomes); a.have_genomes = true; }
        else if (f == "-k" || f == "--kmer-size") a.kmer = atol(val().c_str());
        else if (f == "-o" || f == "--output") a.output = val();
        else if (f == "-t" || f == "--threads") a.threads = atol(val().c_str());
        else if (f == "--debug") a.debug = true;
        else if (f == "--verbose") a.verbose = true;
        else if (a.mode == "call" && (f == "-d" || f == "--db")) { a.db = val(); a.have_db = true; }
        else if (a.mode == "call" && (f == "-r" || f == "--reads")) multi(a.reads);
        else if (a.mode == "call" && (f == "-1" || f == "--first-pairs")) multi(a.first_pairs);
        else if (a.mode == "call" && (f == "-2" || f == "--second-pairs")) multi(a.second_pairs);
        else if (a.mode == "call" && f == "--min-kmers") a.min_kmers = atol(val().c_str());
        else if (a.mode == "call" && f == "--use-full-kmer") a.use_full_kmer = true;
        else if (a.mode == "call" && f == "--n-fixed") a.n_fixed = atol(val().c_str());
        else if (a.mode == "call" && f == "--min-af") a.min_af = atof(val().c_str());
        else if (a.mode == "call" && f == "--no-end-filter") a.no_end_filter = true;
        else if (a.mode == "call" && f == "--no-strand-filter") a.no_strand_filter = true;
        else if (a.mode == "call" && f == "--no-strand-balance-filter") a.no_strand_balance_filter = true;
        else if (a.mode == "call" && f == "--balance-ratio") a.balance_ratio = atof(val().c_str());
        else if (a.mode == "call" && f == "--n-per-strand") a.n_per_strand = atol(val().c_str());
        else if (a.mode == "call" && f == "--strand_odds") a.strand_odds = atof(val().c_str());
        else if (a.mode == "call" && f == "--min-depth") a.min_depth = atol(val().c_str());
        else if (a.mode == "call" && f == "--min-variant-depth") a.min_variant_depth = atol(val().c_str());
        else if (a.mode == "call" && f == "--noise-multiplier") a.noise_multiplier = atof(val().c_str());
        else if (a.mode == "call" && f == "--pileup") a.pileup = true;
        else if (a.mode == "call" && f == "--alignment") a.alignment = true;
        else if (a.mode == "call" && f == "--keep-kmer-info") a.keep_kmer_info = true;
        else if (a.mode == "call" && f == "--device") a.device = atoi(val().c_str());
        else if (a.mode == "call" && f == "--table-log2") a.table_log2 = atol(val().c_str());
        else { fprintf(stderr, "error: unexpected argument '%s' found\n", f.c_str()); exit(2); }
    }
    return a;
}

static void check_common(const Args& a) {
    if (a.kmer % 2 != 1 || a.kmer > 31 || a.kmer < 15) die("Invalid kmer size, must be odd and between [15-31]");
    const long avail = (long)std::thread::hardware_concurrency();
    if (a.threads <= 0) die("Number of threads must be greater than 0");
    if (avail > 0 && a.threads > avail)
        die("You requested " + std::to_string(a.threads) + " threads but only have " + std::to_string(avail) + " available on your system");
}

// ---- bronko build (src/build.rs:62-120) ---------------------------------------------------------
static int run_build(const Args& a) {
    check_common(a);
    if (a.genomes.empty()) die("Please provide the genomes you would like to index.");
    for (const std::string& f : a.genomes)
        if (!check_fasta(f)) die(f + " does not appear to be a fasta file (must be .fa(.gz)/.fasta(.gz)/.fna(.gz))");
    info("Building indexes from fasta files");
    bk::HostIndex ix;
    std::string err;
    // (a file that does not parse ends the reference inside build_indexes with its own message, src/build.rs:156-159)
    if (!bk::index_build_from_fasta((uint32_t)a.kmer, a.genomes, ix, err)) die(err.find("Failed to parse fasta file") != std::string::npos ? err : err + " | Reference failed to build");
    const std::string out = a.output + ".bkdb";
    info("Saving index to " + out);
    if (!bk::bkdb_write(out, ix, err)) die(err + " | Unable to save index");
    return 0;
}

// ---- bronko call (src/call.rs:30-402) -----------------------------------------------------------
struct OutputInfo {      // src/call.rs:138-149
    std::string filename, selected_genome;
    uint64_t num_major = 0, num_minor = 0;
    double breadth = 0, depth = 0;
    uint64_t perfect = 0, variant = 0, unmapped = 0;
    std::vector<bk_variant> variants;
    int best = -1;
};

static void check_call_args(const Args& a) {
    check_common(a);
    for (const std::string& f : a.reads)
        if (!check_fastq(f)) die(f + " does not appear to be a fastq file (must be .fq(.gz)/.fastq(.gz)/.fnq(.gz))");
    if (a.have_genomes && a.have_db) die("Please provide either a db or the genomes you would like to index, not both.");
    if (!a.have_genomes && !a.have_db) die("Please provide either a db or the genomes you would like to index.");
    for (const std::string& f : a.genomes)
        if (!check_fasta(f)) die(f + " does not appear to be a fasta file (must be .fa(.gz)/.fasta(.gz)/.fna(.gz))");
    if (a.min_af < 0.01) warn("Minimum allele frequency set below 0.01, more false positive variants will be returned. We suggest setting this to a more realistic threshold (0.01-0.05)");
    else if (a.min_af > 1.0) die("Minimum allele frequency set above 1, please set between 0-1 (recommended between 0.01-0.05)");
    else if (a.min_af >= 0.5) warn("Minimum allele frequency set equal to or greater than 0.5, no minor variants will be returned");
    if (a.n_per_strand <= 0) warn("Number of kmers per strand set to 0, this is equivalent to no strand filtering");
    else if (a.n_per_strand >= a.kmer) die("Number of kmers per strand set >= k, please set lower value (recommended 2-4, default 2)");
    else if (a.n_per_strand >= 5) warn("Number of kmers per strand set very high, only strongly supported variants will be returned");
    if (a.balance_ratio < 0.0) die("Strand balance ratio is set to below 0, must be between 0.0 and 1.0");
    else if (a.balance_ratio > 1.0) die("Strand balance ratio is set above 1, must be between 0.0 and 1.0");
    else if (a.balance_ratio == 1.0) warn("Strand balance ratio is set to 1, all variants will pass this filter");
    if (a.noise_multiplier < 1.0) die("Noise multiplier for variant detection is set to below 1.0, must be greater than 1.0 (recommended between 1.3-2.0)");
    else if (a.noise_multiplier > 2.0) warn("Strand balance ratio is set above 2, may experience a drop in recall (we recommend ~1.5)");
    else if (a.noise_multiplier == 1.0) warn("Noise multiplier for variant detection set to 1.0, all variants will pass this filter");
    if (a.first_pairs.size() != a.second_pairs.size()) die("Number of paired end sequences do not match, exiting.");
    if (a.min_kmers < 0 || a.n_fixed < 0 || a.min_depth < 0 || a.min_variant_depth < 0) die("negative values are not valid for unsigned options");
}

static void mkdirs(const std::string& path) {
    std::string cur;
    for (size_t i = 0; i <= path.size(); i++) {
        if (i == path.size() || path[i] == '/') {
            if (!cur.empty() && mkdir(cur.c_str(), 0777) != 0 && errno != EEXIST) die(std::string(strerror(errno)) + " | Unable to create outputs in output directory 2");
        }
        if (i < path.size()) cur += path[i];
    }
}

static std::string sample_id(const std::string& path) {
    char buf[4096];
    bk_clean_sample_id(path.c_str(), buf, sizeof buf);
    return buf;
}

#define CK(call) do { int rc__ = (call); if (rc__ != 0) die(bk_last_error(ctx)); } while (0)

static void write_counts_txt(bk_ctx* ctx, int slot, uint32_t k, const std::string& path) {   // the KMC dump kept by --keep-kmer-info
    uint64_t n = 0;
    CK(bk_kmer_counts_get(ctx, slot, nullptr, nullptr, &n));
    std::vector<uint64_t> km(n); std::vector<uint32_t> ct(n);
    if (n) CK(bk_kmer_counts_get(ctx, slot, km.data(), ct.data(), &n));
    FILE* f = fopen(path.c_str(), "wb");
    if (!f) die("Failed to create " + path);
    std::string line(k, 'A');
    for (uint64_t i = 0; i < n; i++) {
        for (uint32_t j = 0; j < k; j++) line[j] = "ACGT"[(km[i] >> (2 * (k - 1 - j))) & 3];
        fprintf(f, "%s\t%u\n", line.c_str(), ct[i]);
    }
    fclose(f);
}

// The host decode stage (FASTQ(.gz) → bases; KMC's reader in the reference).  gzip inflate runs at a few hundred MB/s
// per thread and the GPU path at hundreds of GB/s, so the files of the next samples are decoded on other threads
// while the GPU works on this one: R1 and R2 of a sample concurrently, up to `lookahead` samples ahead (-t bounds it).
struct DecodedSample {
    bk_reads* r[2] = {nullptr, nullptr};
    bool on_device[2] = {false, false};      // BGZF: nothing to do on the host — the GPU's decompression engine inflates it (bk_reads_push_fastq)
    std::string err;
};
// BGZF = gzip whose first member carries the 'BC' extra subfield (bgzip / htslib)
static bool looks_bgzf(const std::string& path) {
    unsigned char h[18];
    FILE* f = fopen(path.c_str(), "rb");
    if (!f) return false;
    const size_t n = fread(h, 1, sizeof h, f);
    fclose(f);
    return n == 18 && h[0] == 0x1f && h[1] == 0x8b && h[2] == 8 && (h[3] & 4) && h[12] == 'B' && h[13] == 'C';
}
static bk_reads* decode_file(const std::string& path, std::string* err) {
    bk_reads* r = nullptr;
    char msg[512] = "";
    if (bk_fastq_decode(path.c_str(), &r, msg, sizeof msg) != 0) { *err = msg[0] ? msg : ("Failed to read reads file: " + path); return nullptr; }
    return r;
}
static DecodedSample decode_sample(std::string r1, std::string r2, bool paired) {
    DecodedSample d;
    std::string e2;
    std::future<bk_reads*> second;
    d.on_device[0] = looks_bgzf(r1);
    d.on_device[1] = paired && looks_bgzf(r2);
    if (paired && !d.on_device[1]) second = std::async(std::launch::async, decode_file, r2, &e2);
    if (!d.on_device[0]) d.r[0] = decode_file(r1, &d.err);
    if (paired && !d.on_device[1]) { d.r[1] = second.get(); if (d.err.empty()) d.err = e2; }
    return d;
}

static OutputInfo process_sample(bk_ctx* ctx, const Args& a, const bk_params& p, const std::string& r1, const std::string* r2, DecodedSample d) {
    if (!d.err.empty()) die(d.err);
    CK(bk_sample_begin(ctx, &p));
    if (d.on_device[0]) CK(bk_reads_push_fastq(ctx, 0, r1.c_str())); else CK(bk_reads_push_decoded(ctx, 0, d.r[0]));
    if (r2) { if (d.on_device[1]) CK(bk_reads_push_fastq(ctx, 1, r2->c_str())); else CK(bk_reads_push_decoded(ctx, 1, d.r[1])); }
    bk_reads_free(d.r[0]); bk_reads_free(d.r[1]);
    bk_sample_result res;
    const int rc = bk_sample_finish(ctx, &res);
    if (rc == BK_ERR_NO_GENOME) die("Unable to pick a best genome");       // src/call.rs:230-233, 320-323
    if (rc != 0) die(bk_last_error(ctx));
    const int nf = r2 ? 2 : 1;
    uint64_t reads = 0, uc = 0, uk = 0, tk = 0;
    for (int f = 0; f < nf; f++) { reads += res.kmc[f].total_reads; uc += res.kmc[f].unique_counted; uk += res.kmc[f].unique_kmers; tk += res.kmc[f].total_kmers; }
    info(std::to_string(reads) + " reads counted from " + r1);
    info(std::to_string(uc) + " unique kmers above " + std::to_string(a.min_kmers) + " count, " + std::to_string(uk) + " total unique kmers, " +
         std::to_string(tk) + " total kmers (~" + std::to_string(tk * (uint64_t)a.kmer) + " basepairs)");
    OutputInfo oi;
    oi.filename = r1;
    oi.best = res.best_genome;
    oi.selected_genome = bk_genome_name(ctx, res.best_genome);
    info("Selected a representative genome: " + oi.selected_genome);
    info("Mapped " + std::to_string(res.num_perfect_kmers) + "/" + std::to_string(uc) + " kmers perfectly, " + std::to_string(res.num_variant_kmers) + "/" +
         std::to_string(uc) + " had a variant, " + std::to_string(res.num_unmapped_kmers) + " unmapped");
    if (((double)(res.num_variant_kmers + res.num_perfect_kmers) / (double)uc) < 0.2)
        warn("Percent of kmers found is very low for this reference, suggesting lack of a representative reference, a bad sequencing run, contamination in sample, or some other issue");
    oi.num_major = res.num_major_variants; oi.num_minor = res.num_minor_variants;
    oi.breadth = res.breadth_coverage; oi.depth = res.depth_coverage;
    oi.perfect = res.num_perfect_kmers; oi.variant = res.num_variant_kmers; oi.unmapped = res.num_unmapped_kmers;
    oi.variants.resize(res.n_variants);
    if (res.n_variants) CK(bk_sample_variants(ctx, oi.variants.data(), res.n_variants));
    info("Called " + std::to_string(oi.num_major) + " major variants, " + std::to_string(oi.num_minor) + " minor above maf = " + bk::fmt_fixed(a.min_af, 2));
    const std::string stem = sample_id(r1);
    if (a.pileup) CK(bk_write_pileup(ctx, (a.output + "/" + stem + ".tsv").c_str()));
    CK(bk_write_vcf(ctx, r1.c_str(), (a.output + "/" + stem + ".vcf").c_str()));
    if (a.keep_kmer_info) {
        write_counts_txt(ctx, 0, (uint32_t)a.kmer, a.output + "/" + stem + "_counts.txt");
        if (r2) write_counts_txt(ctx, 1, (uint32_t)a.kmer, a.output + "/" + sample_id(*r2) + "_counts.txt");
    }
    return oi;
}

// src/call.rs:698-732
static void print_output_info(const Args& a, const std::vector<OutputInfo>& infos) {
    const std::string path = a.output + "/bronko_overview.tsv";
    FILE* f = fopen(path.c_str(), "wb");
    if (!f) die("Failed to create tsv file");
    fprintf(f, "filename\tselected_genome\tnum_major_variants\tnum_minor_variants\tbreadth_coverage\tdepth_coverage\tnum_perfect_kmers\tnum_variant_kmers\tnum_unmapped_kmers\n");
    for (const OutputInfo& o : infos)
        fprintf(f, "%s\t%s\t%llu\t%llu\t%s\t%s\t%llu\t%llu\t%llu\n", o.filename.c_str(), o.selected_genome.c_str(), (unsigned long long)o.num_major,
                (unsigned long long)o.num_minor, bk::fmt_fixed(o.breadth, 4).c_str(), bk::fmt_fixed(o.depth, 4).c_str(), (unsigned long long)o.perfect,
                (unsigned long long)o.variant, (unsigned long long)o.unmapped);
    fclose(f);
}

// src/call.rs:504-628.  Samples appear in processing order (the reference iterates an FxHashMap there).
static void build_alignments(bk_ctx* ctx, const Args& a, const std::vector<OutputInfo>& infos) {
    std::map<int, std::vector<const OutputInfo*>> by_genome;
    for (const OutputInfo& o : infos) {
        if (o.breadth < 0.90) { info("Skipping " + o.filename + " (breadth of coverage = " + std::to_string(o.breadth) + ")"); continue; }
        by_genome[o.best].push_back(&o);
    }
    for (auto& kv : by_genome) {
        const std::string gname = bk_genome_name(ctx, kv.first);
        if (kv.second.size() < 3) { info("Skipping " + gname + " (only " + std::to_string(kv.second.size()) + " samples)"); continue; }
        info("Building alignment for genome " + gname + " with " + std::to_string(kv.second.size()) + " samples");
        typedef std::pair<std::string, uint64_t> Key;            // (sequence name, 1-based pos), sorted like the reference
        std::map<Key, uint8_t> all_pos;
        std::vector<std::map<Key, uint8_t>> per_sample(kv.second.size());
        for (size_t s = 0; s < kv.second.size(); s++)
            for (const bk_variant& v : kv.second[s]->variants)
                if (v.af >= 0.5) {
                    const Key key(bk_seq_name(ctx, kv.first, v.seq), v.pos);
                    all_pos[key] = v.ref_base;
                    per_sample[s][key] = v.alt_base;
                }
        FILE* f = fopen((a.output + "/" + gname + ".mfa").c_str(), "wb");
        if (!f) die("Failed to create mfa alignment file");
        std::string line;
        for (auto& pk : all_pos) line += "ACGT"[pk.second & 3];
        fprintf(f, ">%s\n%s\n", gname.c_str(), line.c_str());
        for (size_t s = 0; s < kv.second.size(); s++) {
            line.clear();
            for (auto& pk : all_pos) {
                auto it = per_sample[s].find(pk.first);
                line += "ACGT"[(it != per_sample[s].end() ? it->second : pk.second) & 3];
            }
            fprintf(f, ">%s\n%s\n", sample_id(kv.second[s]->filename).c_str(), line.c_str());
        }
        fclose(f);
    }
}

static int run_call(const Args& a) {
    check_call_args(a);
    mkdirs(a.output);
    bk_ctx* ctx = nullptr;
    if (bk_create(&ctx, a.device) != 0) die(bk_last_error(nullptr));
    if (a.have_genomes) {
        info("Creating bronko index from provided reference genomes");
        std::vector<const char*> ps;
        for (const std::string& g : a.genomes) ps.push_back(g.c_str());
        if (bk_index_build(ctx, (uint32_t)a.kmer, (uint32_t)ps.size(), ps.data()) != 0) {
            const std::string err = bk_last_error(ctx);
            die(err.find("Failed to parse fasta file") != std::string::npos ? err : err + " | Reference failed to build");
        }
    } else {
        info("Reading in provided bronko index");
        if (bk_index_load_file(ctx, a.db.c_str()) != 0) die(bk_last_error(ctx));
        uint32_t k = 0;
        bk_index_info(ctx, &k, nullptr, nullptr, nullptr);
        if ((long)k != a.kmer) die("Database k is not the same as provided, please set -k to " + std::to_string(k) + " or build a new index");
    }
    bk_params p;
    bk_params_default(&p);
    p.k = (uint32_t)a.kmer; p.min_kmers = (uint32_t)a.min_kmers; p.use_full_kmer = a.use_full_kmer; p.n_fixed = (uint32_t)a.n_fixed;
    p.no_end_filter = a.no_end_filter; p.no_strand_filter = a.no_strand_filter; p.no_strand_balance_filter = a.no_strand_balance_filter;
    p.n_per_strand = (uint32_t)a.n_per_strand; p.min_depth = (uint64_t)a.min_depth; p.min_variant_depth = (uint64_t)a.min_variant_depth;
    p.min_af = a.min_af; p.strand_balance_ratio = a.balance_ratio; p.strand_odds_max = a.strand_odds; p.variant_multiplier = a.noise_multiplier;
    p.table_log2 = (uint32_t)a.table_log2;
    std::vector<OutputInfo> infos;
    // samples in the reference's order (single-end files, then pairs: src/call.rs:213, 298), decoded ahead of the GPU
    struct Job { std::string r1, r2; bool paired; };
    std::vector<Job> jobs;
    for (const std::string& r : a.reads) jobs.push_back(Job{r, "", false});
    for (size_t i = 0; i < a.first_pairs.size(); i++) jobs.push_back(Job{a.first_pairs[i], a.second_pairs[i], true});
    const size_t lookahead = (size_t)std::min<long>(3, std::max<long>(0, a.threads / 2 - 1));    // two decode threads per sample in the window
    std::deque<std::future<DecodedSample>> ahead;
    size_t next = 0;
    for (size_t i = 0; i < jobs.size(); i++) {
        while (next < jobs.size() && ahead.size() < lookahead + 1) {
            ahead.push_back(std::async(std::launch::async, decode_sample, jobs[next].r1, jobs[next].r2, jobs[next].paired));
            next++;
        }
        const Job& j = jobs[i];
        if (j.paired) info("Processing paired reads " + j.r1 + ", " + j.r2); else info("Processing " + j.r1);
        DecodedSample d = ahead.front().get();
        ahead.pop_front();
        infos.push_back(process_sample(ctx, a, p, j.r1, j.paired ? &j.r2 : nullptr, d));
    }
    info("Printing overview");
    print_output_info(a, infos);
    info("All samples processed successfully");
    if (a.alignment) { info("Building alignment(s)"); build_alignments(ctx, a, infos); }
    info("");
    info("bronko complete!");
    bk_destroy(ctx);
    return 0;
}

int main(int argc, char** argv) {
    printf("bronko v%s\nB200 (sm_100a) k-mer->pileup path; CLI compatible with treangenlab/bronko v0.1.0\n\n", VERSION);
    fflush(stdout);
    const auto t0 = std::chrono::steady_clock::now();
    const Args a = parse(argc, argv);
    const int rc = a.mode == "build" ? run_build(a) : run_call(a);
    const double s = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    fprintf(stderr, "\nbronko v%s finished in %gs\n", VERSION, s);
    return rc;
}
