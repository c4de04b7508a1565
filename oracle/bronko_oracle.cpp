// bronko_oracle.cpp — CPU restatement of treangenlab/bronko's k-mer→pileup path.
//
// TEST INFRASTRUCTURE ONLY.  Nothing under bronko_b200/ (the product) may link, import or
// execute this file.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
// `--impl reference` legs use it, as the checker / the reported CPU baseline.
//
// The reference is Rust (no cargo/rustc in this image) and shells out to KMC3 (absent), so it
// cannot be compiled or run here; this file follows the reference source line by line instead.
// Every function cites the reference file:line (relative to /root/reference) it restates.
//
// Pinning status:
//   * assign_buckets        — pinned by the two vectors at src/lcb.rs:146-154 (tests/test_oracle_golden.py)
//   * bincode .bkdb reader, build_indexes — pinned by test_data/hpv.bkdb (parse to EOF; rebuild from
//     HPV16.fa gives the same key→entries map with the same per-key order)
//   * KMC3 counting contract (external binary, version unpinned, README.md:28), map_kmers,
//     pick_best_genome, get_baseline_noise (statrs 0.18 StudentsT::inverse_cdf), call_variants and
//     the writers — PARITY UNPINNED: the reference has no test or fixture for `bronko call`.
//
// Build: see oracle/Makefile  (g++ -O2 -std=c++17 -ffp-contract=off -shared -fPIC ... -lz -lpthread)

#include <algorithm>
#include <array>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <limits>
#include <map>
#include <string>
#include <thread>
#include <unordered_map>
#include <vector>
#include <zlib.h>

typedef uint64_t u64;
typedef uint32_t u32;
typedef uint16_t u16;
typedef uint8_t u8;

// ----------------------------------------------------------------------------------------------
// src/lcb.rs
// ----------------------------------------------------------------------------------------------

// src/lcb.rs:1-45 — all arithmetic is u64 and wraps (release-mode Rust), which only matters for k=31.
static void assign_buckets(u64 kmer, int k, u64* buckets) {
    u64 num_a[32] = {0}, val[32] = {0}, mu[32] = {0};
    u64 mask = 3ull << ((k - 1) * 2);
    u64 p = 1ull << ((k - 1) * 2);
    u64 cur = kmer & mask;
    val[0] = kmer - cur;
    mu[0] = (cur != 0) ? p + ((cur >> 2) * (u64)(k - 1)) : val[0];
    u64 sum_mu = mu[0];
    for (int i = 1; i < k; i++) {
        num_a[i] = num_a[i - 1] + (cur == 0 ? 1 : 0);
        mask >>= 2;
        cur = kmer & mask;
        p >>= 2;
        val[i] = val[i - 1] - cur;
        mu[i] = (cur != 0) ? p + ((cur >> 2) * (u64)(k - i - 1)) : val[i];
        sum_mu += mu[i];
    }
    mask = 3ull << ((k - 1) * 2);
    for (int i = 0; i < k; i++) {
        cur = kmer & mask;
        mask >>= 2;
        buckets[i] = sum_mu - mu[i] + val[i] - num_a[i] * cur + 1 + num_a[i];
    }
}

// src/lcb.rs:47-55 — anything that is not ACGT/acgt encodes as 0 ('A').
static inline u8 nt_to_bits(u8 nt) {
    switch (nt) {
        case 'A': case 'a': return 0;
        case 'C': case 'c': return 1;
        case 'G': case 'g': return 2;
        case 'T': case 't': return 3;
        default: return 0;
    }
}

// src/lcb.rs:57-65
static inline char nucleotide_bits_to_char(u64 bits) {
    switch (bits) { case 0: return 'A'; case 1: return 'C'; case 2: return 'G'; case 3: return 'T'; default: return 'N'; }
}

// src/lcb.rs:67-74
static inline u64 kmer_to_u64(const u8* kmer, int k) {
    u64 v = 0;
    for (int i = 0; i < k; i++) { v <<= 2; v |= nt_to_bits(kmer[i]); }
    return v;
}

// src/lcb.rs:76-85
static inline u64 reverse_complement_u64(u64 kmer_val, int k) {
    u64 rc = 0;
    for (int i = 0; i < k; i++) {
        u64 two = (kmer_val >> (2 * i)) & 3;
        rc <<= 2;
        rc |= (3 ^ two);
    }
    return rc;
}

// src/lcb.rs:87-95 — (canonical value, true iff the canonical form is the reverse complement)
static inline u64 canonical_kmer(const u8* kmer, int k, bool* rc) {
    u64 fwd = kmer_to_u64(kmer, k);
    u64 rev = reverse_complement_u64(fwd, k);
    if (fwd < rev) { *rc = false; return fwd; }
    *rc = true; return rev;
}

// ----------------------------------------------------------------------------------------------
// src/build.rs types (23-60)
// ----------------------------------------------------------------------------------------------

// #[repr(C)] BucketInfo, build.rs:52-60: u16, u8, (pad) u32, u8, bool, (pad) = 12 bytes.
struct BucketInfo {
    u16 file_id;
    u8 seq_id;
    u32 location;
    u8 idx;
    u8 canonical;
};
static_assert(sizeof(BucketInfo) == 12, "BucketInfo must match #[repr(C)] layout");

struct SeqMeta { std::string name; u64 len; std::vector<u8> seq; };
struct FileMeta { std::string name; std::vector<SeqMeta> sequences; };
struct ViralMetadata { std::vector<FileMeta> files; u64 k; };

struct Index {
    u64 k = 0;
    std::unordered_map<u64, std::vector<BucketInfo>> global_index;
    ViralMetadata metadata;
    std::string err;
};

// ----------------------------------------------------------------------------------------------
// bincode 2 `config::standard()` (varint, little endian) — build.rs:134-141, call.rs:186-191
// ----------------------------------------------------------------------------------------------

struct Reader {
    const u8* p; const u8* end; bool ok = true;
    u8 byte() { if (p >= end) { ok = false; return 0; } return *p++; }
    u64 varint() {
        u8 b = byte();
        if (b < 251) return b;
        int n = (b == 251) ? 2 : (b == 252) ? 4 : (b == 253) ? 8 : -1;
        if (n < 0) { ok = false; return 0; }
        u64 v = 0;
        for (int i = 0; i < n; i++) v |= (u64)byte() << (8 * i);
        return v;
    }
    std::string str() {
        u64 n = varint();
        if (!ok || (u64)(end - p) < n) { ok = false; return ""; }
        std::string s((const char*)p, n); p += n; return s;
    }
};

struct Writer {
    std::vector<u8> out;
    void byte(u8 b) { out.push_back(b); }
    void varint(u64 v) {
        if (v < 251) { byte((u8)v); return; }
        int n; u8 tag;
        if (v <= 0xFFFF) { n = 2; tag = 251; } else if (v <= 0xFFFFFFFFull) { n = 4; tag = 252; } else { n = 8; tag = 253; }
        byte(tag);
        for (int i = 0; i < n; i++) byte((u8)(v >> (8 * i)));
    }
    void str(const std::string& s) { varint(s.size()); out.insert(out.end(), s.begin(), s.end()); }
};

static bool read_file(const char* path, std::vector<u8>& buf) {
    FILE* f = fopen(path, "rb");
    if (!f) return false;
    fseek(f, 0, SEEK_END); long n = ftell(f); fseek(f, 0, SEEK_SET);
    buf.resize(n);
    size_t got = n ? fread(buf.data(), 1, n, f) : 0;
    fclose(f);
    return got == (size_t)n;
}

// Decode of BronkoIndex (build.rs:23-28), field order as declared; SURVEY Appendix A.
static Index* index_decode(const std::vector<u8>& buf, u64* consumed) {
    Index* ix = new Index();
    Reader r{buf.data(), buf.data() + buf.size()};
    ix->k = r.varint();
    u64 n_keys = r.varint();
    ix->global_index.reserve(n_keys * 2);
    for (u64 i = 0; i < n_keys && r.ok; i++) {
        u64 key = r.varint();
        u64 n = r.varint();
        std::vector<BucketInfo>& v = ix->global_index[key];
        v.reserve(n);
        for (u64 j = 0; j < n && r.ok; j++) {
            BucketInfo b; memset(&b, 0, sizeof b);
            b.file_id = (u16)r.varint();
            b.seq_id = r.byte();
            b.location = (u32)r.varint();
            b.idx = r.byte();
            b.canonical = r.byte();
            v.push_back(b);
        }
    }
    u64 n_files = r.varint();
    for (u64 f = 0; f < n_files && r.ok; f++) {
        FileMeta fm;
        fm.name = r.str();
        u64 n_seq = r.varint();
        for (u64 s = 0; s < n_seq && r.ok; s++) {
            SeqMeta sm;
            sm.name = r.str();
            sm.len = r.varint();
            u64 n = r.varint();
            if ((u64)(r.end - r.p) < n) { r.ok = false; break; }
            sm.seq.assign(r.p, r.p + n); r.p += n;
            fm.sequences.push_back(std::move(sm));
        }
        ix->metadata.files.push_back(std::move(fm));
    }
    ix->metadata.k = r.varint();
    if (!r.ok) { delete ix; return nullptr; }
    if (consumed) *consumed = (u64)(r.p - buf.data());
    return ix;
}

// Encode (build.rs:122-143).  The reference writes the map in hashbrown iteration order, which is
// not reproducible without emulating hashbrown+FxHash; this writer emits keys in ascending order.
// A reference `bronko call` decodes either order into the same FxHashMap.
static void index_encode(const Index* ix, Writer& w) {
    w.varint(ix->k);
    std::vector<u64> keys; keys.reserve(ix->global_index.size());
    for (auto& kv : ix->global_index) keys.push_back(kv.first);
    std::sort(keys.begin(), keys.end());
    w.varint(keys.size());
    for (u64 key : keys) {
        const auto& v = ix->global_index.at(key);
        w.varint(key);
        w.varint(v.size());
        for (const BucketInfo& b : v) {
            w.varint(b.file_id); w.byte(b.seq_id); w.varint(b.location); w.byte(b.idx); w.byte(b.canonical);
        }
    }
    w.varint(ix->metadata.files.size());
    for (const FileMeta& fm : ix->metadata.files) {
        w.str(fm.name);
        w.varint(fm.sequences.size());
        for (const SeqMeta& sm : fm.sequences) {
            w.str(sm.name); w.varint(sm.len);
            w.varint(sm.seq.size()); w.out.insert(w.out.end(), sm.seq.begin(), sm.seq.end());
        }
    }
    w.varint(ix->metadata.k);
}

// ----------------------------------------------------------------------------------------------
// FASTA / FASTQ text readers (needletail parse_fastx_file semantics for FASTA, build.rs:156-189;
// KMC's FASTQ reader semantics for reads).  gz is auto-detected by zlib.
// ----------------------------------------------------------------------------------------------

static bool slurp_gz(const char* path, std::string& out) {
    gzFile g = gzopen(path, "rb");
    if (!g) return false;
    gzbuffer(g, 1 << 20);
    char buf[1 << 16];
    int n;
    while ((n = gzread(g, buf, sizeof buf)) > 0) out.append(buf, n);
    gzclose(g);
    return n == 0;
}

struct FastaRec { std::string id; std::vector<u8> seq; };

static bool read_fasta(const char* path, std::vector<FastaRec>& recs) {
    std::string txt;
    if (!slurp_gz(path, txt)) return false;
    size_t i = 0, n = txt.size();
    while (i < n) {
        size_t e = txt.find('\n', i);
        if (e == std::string::npos) e = n;
        size_t le = e;
        if (le > i && txt[le - 1] == '\r') le--;
        if (le > i && txt[i] == '>') {
            recs.push_back(FastaRec{txt.substr(i + 1, le - i - 1), {}});
        } else if (!recs.empty()) {
            recs.back().seq.insert(recs.back().seq.end(), txt.begin() + i, txt.begin() + le);
        }
        i = e + 1;
    }
    return true;
}

// file_stem of a path (build.rs:161-165): final component, minus the last extension.
static std::string file_stem(const std::string& path) {
    size_t s = path.find_last_of('/');
    std::string f = (s == std::string::npos) ? path : path.substr(s + 1);
    size_t d = f.find_last_of('.');
    if (d == std::string::npos || d == 0) return f;
    return f.substr(0, d);
}

static std::string first_token(const std::string& s) {
    size_t i = 0;
    while (i < s.size() && isspace((unsigned char)s[i])) i++;
    size_t j = i;
    while (j < s.size() && !isspace((unsigned char)s[j])) j++;
    return s.substr(i, j - i);
}

// src/build.rs:145-231 — per file, per record, every k-mer → canonical → k buckets; per-file maps
// merged in file order so entries within a key are (file order, ascending location).
static Index* build_indexes(int k, int n_files, const char** paths) {
    Index* ix = new Index();
    ix->k = k; ix->metadata.k = k;
    for (int file_id = 0; file_id < n_files; file_id++) {
        std::vector<FastaRec> recs;
        if (!read_fasta(paths[file_id], recs)) { ix->err = std::string("Failed to parse fasta file: ") + paths[file_id]; return ix; }
        FileMeta fm;
        fm.name = file_stem(paths[file_id]);
        u8 seq_id = 0;
        for (FastaRec& rec : recs) {
            SeqMeta sm;
            sm.name = first_token(rec.id);
            sm.len = rec.seq.size();
            sm.seq = rec.seq;
            size_t seq_len = rec.seq.size();
            // `for i in 0..=seq_len.saturating_sub(k)` then `&seq[i..i+k]` panics when seq_len < k;
            // the oracle skips such sequences' k-mers instead of aborting.
            if (seq_len >= (size_t)k) {
                u64 buckets[32];
                for (size_t i = 0; i + k <= seq_len; i++) {
                    bool canonical;
                    u64 kb = canonical_kmer(rec.seq.data() + i, k, &canonical);
                    assign_buckets(kb, k, buckets);
                    for (int j = 0; j < k; j++) {
                        BucketInfo b; memset(&b, 0, sizeof b);
                        b.file_id = (u16)file_id; b.seq_id = seq_id; b.location = (u32)i; b.idx = (u8)j; b.canonical = canonical;
                        ix->global_index[buckets[j]].push_back(b);
                    }
                }
            }
            fm.sequences.push_back(std::move(sm));
            seq_id += 1;  // u8 wraps after 255 as in build.rs:170,207
        }
        ix->metadata.files.push_back(std::move(fm));
    }
    return ix;
}

// ----------------------------------------------------------------------------------------------
// KMC3 contract (call.rs:1152-1226):  kmc -k{K} -b -ci{min} -cs1000000
//   non-canonical exact counts; reads split at non-ACGT symbols; acgt == ACGT; keep ci <= count <= cx
//   (cx default 1e9); stored counter = min(count, cs).
// ----------------------------------------------------------------------------------------------

struct FlatCounter {
    std::vector<u64> keys; std::vector<u64> vals; u64 mask = 0, n = 0;
    static constexpr u64 EMPTY = ~0ull;
    void init(u64 cap_pow2) { keys.assign(cap_pow2, EMPTY); vals.assign(cap_pow2, 0); mask = cap_pow2 - 1; n = 0; }
    static inline u64 mix(u64 x) { x ^= x >> 31; x *= 0x7fb5d329728ea185ull; x ^= x >> 27; x *= 0x81dadef4bc2dd44dull; x ^= x >> 33; return x; }
    void grow() {
        std::vector<u64> ok, ov; ok.swap(keys); ov.swap(vals);
        init((mask + 1) * 2);
        for (size_t i = 0; i < ok.size(); i++) if (ok[i] != EMPTY) add(ok[i], ov[i]);
    }
    void add(u64 key, u64 c) {
        if ((n + 1) * 10 > (mask + 1) * 7) grow();
        u64 h = mix(key) & mask;
        while (true) {
            if (keys[h] == key) { vals[h] += c; return; }
            if (keys[h] == EMPTY) { keys[h] = key; vals[h] = c; n++; return; }
            h = (h + 1) & mask;
        }
    }
};

struct Counts {
    std::vector<u64> kmers;   // kept k-mers (ci <= count <= cx), ascending
    std::vector<u64> counts;  // min(count, cs)
    u64 total_reads = 0, total_kmers = 0, unique_kmers = 0, unique_counted = 0;
};

static inline int base_code(u8 c) {
    switch (c) {
        case 'A': case 'a': return 0; case 'C': case 'c': return 1;
        case 'G': case 'g': return 2; case 'T': case 't': return 3; default: return -1;
    }
}

static Counts* count_kmers(int k, const u8* bases, const u64* off, u64 n_reads, u64 ci, u64 cs, int threads) {
    if (threads < 1) threads = 1;
    const int P = threads;
    const u64 kmask = (k == 32) ? ~0ull : ((1ull << (2 * k)) - 1);
    // pass 1: each thread rolls the k-mers of its slice of reads into P hash-partitioned buffers
    std::vector<std::vector<std::vector<u64>>> part(P, std::vector<std::vector<u64>>(P));
    std::vector<u64> tot(P, 0);
    auto scan = [&](int t) {
        u64 r0 = n_reads * t / P, r1 = n_reads * (t + 1) / P;
        for (u64 r = r0; r < r1; r++) {
            const u8* s = bases + off[r]; u64 len = off[r + 1] - off[r];
            u64 cur = 0; int valid = 0;
            for (u64 i = 0; i < len; i++) {
                int c = base_code(s[i]);
                if (c < 0) { valid = 0; cur = 0; continue; }
                cur = ((cur << 2) | (u64)c) & kmask;
                if (++valid >= k) { part[t][FlatCounter::mix(cur) % P].push_back(cur); tot[t]++; }
            }
        }
    };
    {
        std::vector<std::thread> th;
        for (int t = 0; t < P; t++) th.emplace_back(scan, t);
        for (auto& x : th) x.join();
    }
    // pass 2: thread p owns partition p
    std::vector<FlatCounter> tabs(P);
    auto merge = [&](int p) {
        u64 n = 0; for (int t = 0; t < P; t++) n += part[t][p].size();
        u64 cap = 1024; while (cap < n / 4 + 16) cap <<= 1;
        tabs[p].init(cap);
        for (int t = 0; t < P; t++) { for (u64 km : part[t][p]) tabs[p].add(km, 1); std::vector<u64>().swap(part[t][p]); }
    };
    {
        std::vector<std::thread> th;
        for (int p = 0; p < P; p++) th.emplace_back(merge, p);
        for (auto& x : th) x.join();
    }
    Counts* c = new Counts();
    c->total_reads = n_reads;
    for (int t = 0; t < P; t++) c->total_kmers += tot[t];
    const u64 cx = 1000000000ull;
    std::vector<std::pair<u64, u64>> kept;
    for (int p = 0; p < P; p++) {
        c->unique_kmers += tabs[p].n;
        for (size_t i = 0; i < tabs[p].keys.size(); i++) {
            if (tabs[p].keys[i] == FlatCounter::EMPTY) continue;
            u64 v = tabs[p].vals[i];
            if (v >= ci && v <= cx) kept.emplace_back(tabs[p].keys[i], std::min(v, cs));
        }
    }
    std::sort(kept.begin(), kept.end());
    c->unique_counted = kept.size();
    c->kmers.reserve(kept.size()); c->counts.reserve(kept.size());
    for (auto& kv : kept) { c->kmers.push_back(kv.first); c->counts.push_back(kv.second); }
    return c;
}

// FASTQ: 4-line records, sequence = 2nd line (CR stripped).  KMC contract, SURVEY Appendix B.
struct Reads { std::vector<u8> bases; std::vector<u64> off; };
static Reads* read_fastq(const char* path) {
    std::string txt;
    if (!slurp_gz(path, txt)) return nullptr;
    Reads* rd = new Reads();
    rd->off.push_back(0);
    size_t i = 0, n = txt.size(); int line = 0;
    while (i < n) {
        size_t e = txt.find('\n', i);
        if (e == std::string::npos) e = n;
        size_t le = e;
        if (le > i && txt[le - 1] == '\r') le--;
        if ((line & 3) == 1) {
            rd->bases.insert(rd->bases.end(), txt.begin() + i, txt.begin() + le);
            rd->off.push_back(rd->bases.size());
        }
        line++;
        i = e + 1;
    }
    return rd;
}

// ----------------------------------------------------------------------------------------------
// Pileups (call.rs:1235-1239, 1437-1480): per genome, per sequence, four Vec<[u64;4]>:
//   0 = output (fwd depth), 1 = output_rev (rev depth), 2 = output_counts (fwd support), 3 = output_rev_counts.
// Flattened: arr[a][row][base], row = genome_row0[g] + seq_row0 + position.
// ----------------------------------------------------------------------------------------------

struct Params {
    u64 k, min_kmers; int use_full_kmer; u64 n_fixed;
    double min_af; int no_end_filter, no_strand_filter, no_strand_balance_filter;
    double strand_balance_ratio; u64 n_per_strand; double strand_odds_max;
    u64 min_depth, min_variant_depth; double variant_multiplier;
};

struct VCFRecord {      // call.rs:776-789 (seq kept as an index into the best genome's sequences)
    u32 seq; u32 pos; u8 ref_base, alt_base; u8 pad[6];
    u64 fwd_ref, rev_ref, fwd_alt, rev_alt, depth;
    double af, sor;
};

struct GenomeStats { u64 perfect = 0, variant = 0, unique = 0; bool present = false; };

struct Sample {
    const Index* ix = nullptr;
    std::vector<u64> genome_row0;             // first row of each genome (+ total at the end)
    std::vector<std::vector<u64>> seq_row0;   // [g][s] row offset inside the genome
    std::vector<std::array<u64, 4>> arr[4];
    std::vector<GenomeStats> stats[2];        // per reads file
    int n_files = 0;
    int best = -1;
    std::vector<VCFRecord> variants;
    u64 num_major = 0, num_minor = 0;
    double breadth = 0, depth_cov = 0;
    std::vector<double> noise_max;            // per row of the best genome
    u64 unique_counted[2] = {0, 0};
};

// call.rs:1437-1480
static void initialize_output_maps(Sample* s) {
    const ViralMetadata& md = s->ix->metadata;
    u64 rows = 0;
    s->genome_row0.clear(); s->seq_row0.clear();
    for (const FileMeta& fm : md.files) {
        s->genome_row0.push_back(rows);
        std::vector<u64> so; u64 r = 0;
        for (const SeqMeta& sm : fm.sequences) { so.push_back(r); r += sm.len; }
        s->seq_row0.push_back(so);
        rows += r;
    }
    s->genome_row0.push_back(rows);
    for (int a = 0; a < 4; a++) s->arr[a].assign(rows, std::array<u64, 4>{0, 0, 0, 0});
}

// call.rs:1257-1434 — one counted k-mer at a time (the rayon chunking only affects scheduling;
// all updates are max / += and commute).
static void map_kmers(Sample* s, const Counts* kc, const Params& pr, std::vector<GenomeStats>& out) {
    const Index* ix = s->ix;
    const int k = (int)pr.k;
    const size_t n_genomes = ix->metadata.files.size();
    out.assign(n_genomes, GenomeStats());
    std::vector<u64> hits(n_genomes, 0);
    std::vector<u16> touched;
    u64 buckets[32];
    u8 kmer_txt[32];
    for (size_t t = 0; t < kc->kmers.size(); t++) {
        u64 kv = kc->kmers[t];
        const u64 n = kc->counts[t];
        // the reference holds the k-mer as text (KMC dump) and re-encodes it: call.rs:1288
        for (int i = 0; i < k; i++) kmer_txt[i] = "ACGT"[(kv >> (2 * (k - 1 - i))) & 3];
        bool rc;
        u64 kmer_bin = canonical_kmer(kmer_txt, k, &rc);
        assign_buckets(kmer_bin, k, buckets);
        // call.rs:1291-1300 — asymmetric slice [n_fixed, k - n_fixed - 1)
        int b0 = 0, b1 = k;
        if (!pr.use_full_kmer) {
            if (pr.n_fixed * 2 + 1 >= (u64)k) { b0 = b1 = 0; }
            else { b0 = (int)pr.n_fixed; b1 = k - (int)pr.n_fixed - 1; }
        }
        const u64 num_buckets_perfect = (u64)(b1 - b0);
        touched.clear();
        for (int bi = b0; bi < b1; bi++) {
            auto it = ix->global_index.find(buckets[bi]);
            if (it == ix->global_index.end()) continue;
            for (const BucketInfo& info : it->second) {
                if (hits[info.file_id]++ == 0) touched.push_back(info.file_id);   // call.rs:1316-1318
                const u64 genome_pos = info.location;
                const u64 nuc_x = info.idx;
                const u64 row = s->genome_row0[info.file_id] + s->seq_row0[info.file_id][info.seq_id] + genome_pos + nuc_x;
                u64 bit_idx; int depth_arr, count_arr;
                if (info.canonical) {                                             // call.rs:1330-1357
                    u64 pos = k - nuc_x - 1;
                    bit_idx = ((kmer_bin >> (2 * (k - pos - 1))) & 3) ^ 3;
                    if (rc) { count_arr = 2; depth_arr = 0; } else { count_arr = 3; depth_arr = 1; }
                } else {                                                          // call.rs:1358-1384
                    u64 pos = nuc_x;
                    bit_idx = (kmer_bin >> (2 * (k - pos - 1))) & 3;
                    if (rc) { count_arr = 3; depth_arr = 1; } else { count_arr = 2; depth_arr = 0; }
                }
                s->arr[count_arr][row][bit_idx] += 1;
                if (s->arr[depth_arr][row][bit_idx] < n) s->arr[depth_arr][row][bit_idx] = n;
            }
        }
        // call.rs:1389-1419
        int n_perfect = 0; u16 uniq = 0;
        for (u16 g : touched) if (hits[g] == num_buckets_perfect) { n_perfect++; uniq = g; }
        for (u16 g : touched) {
            out[g].present = true;
            if (hits[g] == num_buckets_perfect) out[g].perfect++;
            else if (hits[g] > 0) out[g].variant++;
        }
        if (n_perfect == 1) out[uniq].unique++;
        for (u16 g : touched) hits[g] = 0;
    }
}

// call.rs:422-450 / 452-502 — strict '>' from 0.0.  The reference iterates an FxHashMap, so exact
// ties resolve by hash-table order (unpinned); here ties keep the lowest file index.
static int pick_best_genome(const Sample* s) {
    const ViralMetadata& md = s->ix->metadata;
    int best = -1; double best_score = 0.0;
    for (size_t g = 0; g < md.files.size(); g++) {
        bool present = false; u64 perfect = 0;
        for (int f = 0; f < s->n_files; f++) if (s->stats[f][g].present) { present = true; perfect += s->stats[f][g].perfect; }
        if (!present) continue;
        u64 genome_len = 0;
        for (const SeqMeta& sm : md.files[g].sequences) genome_len += sm.len;
        double score = (double)perfect / (double)genome_len / 2.0;
        if (score > best_score) { best_score = score; best = (int)g; }
    }
    return best;
}

// ----------------------------------------------------------------------------------------------
// statrs 0.18 StudentsT::inverse_cdf (call.rs:924-925).  statrs is not in the tree; this restates
// its published algorithm (distribution/students_t.rs inverse_cdf → function/beta.rs inv_beta_reg,
// a port of AS 109, + beta_reg continued fraction + Lanczos ln_gamma) from the upstream source as
// remembered.  PARITY UNPINNED (no reference test); anchored against scipy in tests to 1e-10.
// ----------------------------------------------------------------------------------------------

static double ln_gamma(double x) {
    static const double R = 10.900511;
    static const double DK[11] = {
        2.48574089138753565546e-5, 1.05142378581721974210, -3.45687097222016235469,
        4.51227709466894823700, -2.98285225323576655721, 1.05639711577126713077,
        -1.95428773191645869583e-1, 1.70970543404441224307e-2, -5.71926117404305781283e-4,
        4.63399473359905636708e-6, -2.71994908488607703910e-9};
    static const double LN_2_SQRT_E_OVER_PI = 0.6207822376352452223455184457816472122518527279025978;
    static const double LN_PI = 1.1447298858494001741434273513530587116472948129153;
    if (x < 0.5) {
        double s = DK[0];
        for (int i = 1; i < 11; i++) s += DK[i] / ((double)i - x);
        return LN_PI - std::log(std::sin(M_PI * x)) - std::log(s) - LN_2_SQRT_E_OVER_PI - (0.5 - x) * std::log((0.5 - x + R) / M_E);
    }
    double s = DK[0];
    for (int i = 1; i < 11; i++) s += DK[i] / (x + (double)i - 1.0);
    return std::log(s) + LN_2_SQRT_E_OVER_PI + (x - 0.5) * std::log((x - 0.5 + R) / M_E);
}

static double beta_reg(double a, double b, double x) {
    double bt = (x == 0.0 || x == 1.0) ? 0.0
        : std::exp(ln_gamma(a + b) - ln_gamma(a) - ln_gamma(b) + a * std::log(x) + b * std::log(1.0 - x));
    bool symm = x >= (a + 1.0) / (a + b + 2.0);
    const double eps = 1.1102230246251565e-16;
    const double fpmin = std::numeric_limits<double>::min() / eps;
    if (symm) { std::swap(a, b); x = 1.0 - x; }
    double qab = a + b, qap = a + 1.0, qam = a - 1.0;
    double c = 1.0, d = 1.0 - qab * x / qap;
    if (std::fabs(d) < fpmin) d = fpmin;
    d = 1.0 / d;
    double h = d;
    for (int mi = 1; mi < 141; mi++) {
        double m = mi, m2 = m * 2.0;
        double aa = m * (b - m) * x / ((qam + m2) * (a + m2));
        d = 1.0 + aa * d; if (std::fabs(d) < fpmin) d = fpmin;
        c = 1.0 + aa / c; if (std::fabs(c) < fpmin) c = fpmin;
        d = 1.0 / d;
        h = h * d * c;
        aa = -(a + m) * (qab + m) * x / ((a + m2) * (qap + m2));
        d = 1.0 + aa * d; if (std::fabs(d) < fpmin) d = fpmin;
        c = 1.0 + aa / c; if (std::fabs(c) < fpmin) c = fpmin;
        d = 1.0 / d;
        double del = d * c;
        h *= del;
        if (std::fabs(del - 1.0) <= eps) break;
    }
    return symm ? 1.0 - bt * h / a : bt * h / a;
}

static double inv_beta_reg(double a, double b, double x) {
    const double SAE = -30.0, FPU = 1e-30;
    if (x == 0.0) return 0.0;
    if (x == 1.0) return 1.0;
    double beta = ln_gamma(a) + ln_gamma(b) - ln_gamma(a + b);
    bool flip = false;
    if (x > 0.5) { std::swap(a, b); x = 1.0 - x; flip = true; }
    double r = std::sqrt(-std::log(x * x));
    double y = r - (2.30753 + 0.27061 * r) / (1.0 + (0.99229 + 0.04481 * r) * r);
    double p;
    if (a > 1.0 && b > 1.0) {
        r = (y * y - 3.0) / 6.0;
        double s = 1.0 / (a + a - 1.0), t = 1.0 / (b + b - 1.0);
        double h = 2.0 / (s + t);
        double w = y * std::sqrt(h + r) / h - (t - s) * (r + 5.0 / 6.0 - 2.0 / (3.0 * h));
        p = a / (a + b * std::exp(w + w));
    } else {
        r = b + b;
        double t = 1.0 / (9.0 * b);
        t = r * std::pow(1.0 - t + y * std::sqrt(t), 3.0);
        if (t <= 0.0) {
            p = 1.0 - std::exp((std::log((1.0 - x) * b) + beta) / b);
        } else {
            t = (4.0 * a + r - 2.0) / t;
            if (t <= 1.0) p = std::exp((std::log(x * a) + beta) / a);
            else p = 1.0 - 2.0 / (t + 1.0);
        }
    }
    r = 1.0 - a;
    double t = 1.0 - b;
    double yprev = 0.0, sq = 1.0, prev = 1.0;
    if (p < 0.0001) p = 0.0001;
    if (p > 0.9999) p = 0.9999;
    double iex = std::max(-5.0 / a / a - 1.0 / std::pow(x, 0.2) - 13.0, SAE);
    double acu = std::pow(10.0, iex);
    double tx = p;
    for (int guard = 0; guard < 1000; guard++) {
        y = beta_reg(a, b, p);
        y = (y - x) * std::exp(beta + r * std::log(p) + t * std::log(1.0 - p));
        if (y * yprev <= 0.0) prev = std::max(sq, FPU);
        double g = 1.0;
        bool done = false;
        while (true) {
            while (true) {
                double adj = g * y;
                sq = adj * adj;
                if (sq < prev) {
                    tx = p - adj;
                    if (tx >= 0.0 && tx <= 1.0) break;
                }
                g /= 3.0;
            }
            if (prev <= acu || y * y <= acu) { p = tx; done = true; break; }
            if (tx != 0.0 && tx != 1.0) break;
            g /= 3.0;
        }
        if (done) break;
        if (tx == p) break;
        p = tx;
        yprev = y;
    }
    return flip ? 1.0 - p : p;
}

// StudentsT::new(0,1,df).inverse_cdf(x)
static double students_t_inverse_cdf(double df, double x) {
    double x1 = (x >= 0.5) ? 1.0 - x : x;
    double y = inv_beta_reg(0.5 * df, 0.5, 2.0 * x1);
    y = std::sqrt(df * (1.0 - y) / y);
    return (x >= 0.5) ? y : -y;
}

// tau as computed at call.rs:922-929 for a given curr_n (> 2)
static double thompson_tau(u64 curr_n) {
    const double alpha = 0.001;
    double n = (double)curr_n;
    double df = (double)(curr_n - 2);
    double t_crit = students_t_inverse_cdf(df, 1.0 - alpha / n);
    return (t_crit * (n - 1.0)) / (std::sqrt(n) * std::sqrt(n - 2.0 + t_crit * t_crit));
}

// ----------------------------------------------------------------------------------------------
// call.rs:799-967 — streaming modified Thompson-tau baseline; only `.max` is consumed downstream
// (call.rs:1107) but mean/std are produced too for inspection.
// ----------------------------------------------------------------------------------------------
struct Noise { double max, mean, std; };

static bool get_baseline_noise(const std::array<u64, 4>* fwd, const std::array<u64, 4>* rev, size_t len, std::vector<Noise>& out) {
    const size_t window_size = 100;
    const size_t max_table_len = window_size / 10;
    static double tau_tab[301]; static bool tau_init = false;
    if (!tau_init) { for (u64 n = 3; n <= 300; n++) tau_tab[n] = thompson_tau(n); tau_init = true; }

    out.assign(len, Noise{0.0, 0.0, 0.0});
    // `vec![0.0; len*3]` indexed up to 299: the reference panics (index out of bounds) for len < 100.
    if (len * 3 < window_size * 3) return false;
    std::vector<double> window_counts(len * 3, 0.0);
    std::vector<int> in_max(len * 3, 0);
    double maxes[10]; for (size_t i = 0; i < max_table_len; i++) maxes[i] = 0.0;
    size_t n = 0; double s = 0.0, s2 = 0.0, mu, var;
    const size_t half_window = window_size / 2;

    for (size_t i = 0; i < len + half_window; i++) {
        size_t base_pos = (i % window_size) * 3;
        double freqs[4] = {0.0, 0.0, 0.0, 0.0};
        if (i < len) {
            u64 counts[4];
            for (int b = 0; b < 4; b++) counts[b] = fwd[i][b] + rev[i][b];
            std::sort(counts, counts + 4, [](u64 a, u64 b) { return a > b; });
            u64 total_depth = counts[0] + counts[1] + counts[2] + counts[3];
            if (total_depth != 0) for (int b = 0; b < 4; b++) freqs[b] = (double)counts[b] / (double)total_depth;
        }
        for (size_t j = 1; j < 4; j++) {
            size_t idx = base_pos + (j - 1);
            double old = window_counts[idx];
            if (old > 0.0) {
                n -= 1; s -= old; s2 -= old * old;
                if (in_max[idx] == 1) {
                    size_t pos = max_table_len;
                    for (size_t q = 0; q < max_table_len; q++) if (std::fabs(maxes[q] - old) < 1e-12) { pos = q; break; }
                    if (pos < max_table_len) {
                        for (size_t q = pos; q < max_table_len - 1; q++) maxes[q] = maxes[q + 1];
                        maxes[max_table_len - 1] = 0.0;
                    }
                    in_max[idx] = 0;
                }
            }
            double maf = freqs[j];
            if (maf > 0.0) {
                n += 1; s += maf; s2 += maf * maf;
                for (size_t q = max_table_len; q-- > 0;) {
                    if (maf > maxes[q]) {
                        if (q + 1 < max_table_len) maxes[q + 1] = maxes[q];
                        maxes[q] = maf;
                    } else break;
                }
                in_max[idx] = 1;
            } else {
                in_max[idx] = 0;
                window_counts[idx] = 0.0;
            }
            window_counts[idx] = maf;
        }
        if (n != 0) { mu = s / (double)n; var = (s2 / (double)n) - mu * mu; } else { mu = 0.0; var = 0.0; }

        size_t curr_max_idx = 0; size_t curr_n = n;
        double curr_s = s, curr_s2 = s2, curr_mu = mu, curr_var = var;
        while (curr_max_idx < max_table_len && maxes[curr_max_idx] != 0.0) {
            double candidate = maxes[curr_max_idx];
            double sd = std::sqrt(curr_var);
            double tau = (curr_n > 2) ? tau_tab[curr_n] : std::numeric_limits<double>::infinity();
            if (std::fabs(candidate - curr_mu) > tau * sd) {
                curr_s -= candidate;
                curr_s2 -= candidate;          // sic: not candidate^2 (call.rs:936)
                curr_n -= 1;
                if (curr_n > 0) { curr_mu = curr_s / (double)curr_n; curr_var = (curr_s2 / (double)curr_n) - curr_mu * curr_mu; }
                else { curr_mu = 0.0; curr_var = 0.0; }
                curr_max_idx += 1;
            } else break;
        }
        if (i >= half_window) {
            size_t w = i - half_window;
            // maxes[curr_max_idx] with curr_max_idx == 10 would be an out-of-bounds panic in the
            // reference (all ten table entries rejected); reported as 0 here.
            if (w < len) out[w] = Noise{curr_max_idx < max_table_len ? maxes[curr_max_idx] : 0.0, curr_mu, std::sqrt(curr_var)};
        }
    }
    return true;
}

// call.rs:969-1150.  Sequences are visited in metadata order (the reference iterates a DashMap,
// whose order is random per run — Q19; single-contig genomes are unaffected).
static void call_variants(Sample* s, const Params& pr) {
    const FileMeta& fm = s->ix->metadata.files[s->best];
    const u64 g0 = s->genome_row0[s->best];
    u64 positions_covered = 0, total_positions = 0, total_coverage = 0;
    s->variants.clear(); s->num_major = s->num_minor = 0;
    s->noise_max.assign(s->genome_row0[s->best + 1] - g0, 0.0);
    const bool filter_end_seq = !pr.no_end_filter, strand_filter = !pr.no_strand_filter;
    for (size_t si = 0; si < fm.sequences.size(); si++) {
        const SeqMeta& sm = fm.sequences[si];
        const u64 r0 = g0 + s->seq_row0[s->best][si];
        const std::array<u64, 4>* fwd = s->arr[0].data() + r0;
        const std::array<u64, 4>* rev = s->arr[1].data() + r0;
        const std::array<u64, 4>* fwd_counts = s->arr[2].data() + r0;
        const std::array<u64, 4>* rev_counts = s->arr[3].data() + r0;
        const size_t len = sm.len;
        std::vector<Noise> baseline;
        bool ok = get_baseline_noise(fwd, rev, len, baseline);
        (void)ok;
        for (size_t i = 0; i < len; i++) s->noise_max[r0 - g0 + i] = baseline[i].max;
        size_t start = 0, end = len;
        if (filter_end_seq) { start = pr.k; end = (len >= pr.k) ? len - pr.k : 0; }
        total_positions += len;
        for (size_t i = start; i < end; i++) {
            const std::array<u64, 4>& row = fwd[i]; const std::array<u64, 4>& row_rev = rev[i];
            const std::array<u64, 4>& count = fwd_counts[i]; const std::array<u64, 4>& count_rev = rev_counts[i];
            u8 ref_base = nt_to_bits(sm.seq[i]);
            u64 row_total[4]; u64 total_depth = 0;
            for (int b = 0; b < 4; b++) { row_total[b] = row[b] + row_rev[b]; total_depth += row_total[b]; }
            if (total_depth == 0) continue;
            positions_covered += 1; total_coverage += total_depth;
            for (u8 alt = 0; alt < 4; alt++) {
                if (alt == ref_base || row_total[alt] == 0) continue;
                double sor = pr.strand_odds_max + 1.0;
                if (strand_filter) {
                    double a = (double)row[ref_base] + 1.0, b = (double)row_rev[ref_base] + 1.0;
                    double c = (double)row[alt] + 1.0, d = (double)row_rev[alt] + 1.0;
                    double ref_total = a + b + c + d;
                    double min_strand_depth = std::fmin(a + c, b + d);
                    double min_strand_percent = min_strand_depth / ref_total;
                    if ((!pr.no_strand_balance_filter) | ((pr.no_strand_balance_filter != 0) & (min_strand_percent >= pr.strand_balance_ratio))) {
                        double r = (a * d) / (b * c);
                        double ref_ratio = std::fmin(a, b) / std::fmax(a, b);
                        double alt_ratio = std::fmin(c, d) / std::fmax(c, d);
                        sor = std::log(r + (1.0 / r)) + std::log(ref_ratio) - std::log(alt_ratio);
                        if (sor > pr.strand_odds_max) continue;
                        u64 c_k = count[alt], d_k = count_rev[alt];
                        if (c_k < pr.n_per_strand && d_k < pr.n_per_strand) continue;
                    } else sor = -1.0;
                }
                u64 alt_count = row_total[alt];
                double af = (double)alt_count / (double)total_depth;
                double y0 = pr.variant_multiplier, p0 = 0.5, aa = 0.03;
                double factor = y0 + p0 * std::pow(aa, 100.0 * af);
                if (af < pr.min_af || af < (std::fmax(factor, y0) * baseline[i].max)) continue;
                if (af >= 0.5) s->num_major += 1;
                else {
                    if (total_depth < pr.min_depth) continue;
                    if (alt_count < pr.min_variant_depth) continue;
                    s->num_minor += 1;
                }
                VCFRecord v; memset(&v, 0, sizeof v);
                v.seq = (u32)si; v.pos = (u32)(i + 1); v.ref_base = ref_base; v.alt_base = alt;
                v.fwd_ref = row[ref_base]; v.rev_ref = row_rev[ref_base]; v.fwd_alt = row[alt]; v.rev_alt = row_rev[alt];
                v.depth = total_depth; v.af = af; v.sor = sor;
                s->variants.push_back(v);
            }
        }
    }
    s->breadth = (double)positions_covered / (double)total_positions;
    s->depth_cov = (double)total_coverage / (double)positions_covered;
}

// ----------------------------------------------------------------------------------------------
// util.rs:30-50 and writers call.rs:648-774, 504-628
// ----------------------------------------------------------------------------------------------

static bool ends_with(const std::string& s, const std::string& suf) {
    return s.size() >= suf.size() && s.compare(s.size() - suf.size(), suf.size(), suf) == 0;
}

static std::string clean_sample_id(const std::string& path) {
    size_t sl = path.find_last_of('/');
    std::string filename = (sl == std::string::npos) ? path : path.substr(sl + 1);
    static const char* suffixes[] = {".fastq.gz", ".fasta.gz", "fna.gz", "fnq.gz", ".fq.gz", ".fastq", ".fasta", ".fnq", ".fna", ".fa", ".fq"};
    for (const char* suf : suffixes) {
        if (ends_with(filename, suf)) {
            std::string r = filename;      // trim_end_matches: strip the suffix repeatedly
            std::string sf = suf;
            while (ends_with(r, sf)) r.erase(r.size() - sf.size());
            return r;
        }
    }
    return file_stem(filename);
}

// Rust `{:.N}` for f64 (NaN → "NaN", ±inf → "inf"/"-inf"); finite values agree with printf.
static std::string fmt_fixed(double v, int prec) {
    if (std::isnan(v)) return "NaN";
    if (std::isinf(v)) return v > 0 ? "inf" : "-inf";
    char buf[64]; snprintf(buf, sizeof buf, "%.*f", prec, v); return buf;
}

static std::string vcf_text(const Sample* s, const std::string& reads_path) {
    const FileMeta& fm = s->ix->metadata.files[s->best];
    std::string o;
    o += "##fileformat=VCFv4.5\n##source=bronko-v0.1.0\n";
    o += "##reference=file://" + reads_path + "\n";
    for (const SeqMeta& sm : fm.sequences) o += "##contig=<ID=" + first_token(sm.name) + ",length=" + std::to_string(sm.len) + ">\n";
    o += "##INFO=<ID=DP,Number=1,Type=Integer,Description=\"Total Depth\">\n";
    o += "##INFO=<ID=AF,Number=1,Type=Float,Description=\"Allele Frequency\">\n";
    o += "##INFO=<ID=DP4,Number=4,Type=Integer,Description=\"Fwd_ref,Rev_ref,Fwd_alt,Rev_alt\">\n";
    o += "##INFO=<ID=SOR,Number=4,Type=Float,Description=\"SOR\">\n";
    o += "#CHROM\tPOS\tID\tREF\tALT\tQUAL\tFILTER\tINFO\n";
    for (const VCFRecord& v : s->variants) {
        o += first_token(fm.sequences[v.seq].name) + "\t" + std::to_string(v.pos) + "\t.\t";
        o += nucleotide_bits_to_char(v.ref_base); o += "\t"; o += nucleotide_bits_to_char(v.alt_base);
        o += "\t.\tPASS\tDP=" + std::to_string(v.depth) + ";AF=" + fmt_fixed(v.af, 3) + ";DP4=" +
             std::to_string(v.fwd_ref) + "," + std::to_string(v.rev_ref) + "," + std::to_string(v.fwd_alt) + "," + std::to_string(v.rev_alt) +
             ";SOR=" + fmt_fixed(v.sor, 3) + "\n";
    }
    return o;
}

static std::string pileup_text(const Sample* s) {
    const FileMeta& fm = s->ix->metadata.files[s->best];
    std::string o = "reference\tindex\tref\tA\tC\tG\tT\ta\tc\tg\tt\n";
    for (size_t si = 0; si < fm.sequences.size(); si++) {
        const SeqMeta& sm = fm.sequences[si];
        u64 r0 = s->genome_row0[s->best] + s->seq_row0[s->best][si];
        for (size_t i = 0; i < sm.seq.size(); i++) {
            o += sm.name + "\t" + std::to_string(i + 1) + "\t"; o += (char)sm.seq[i];
            for (int a = 0; a < 2; a++) for (int b = 0; b < 4; b++) o += "\t" + std::to_string(s->arr[a][r0 + i][b]);
            o += "\n";
        }
    }
    return o;
}

static bool write_text(const char* path, const std::string& t) {
    FILE* f = fopen(path, "wb"); if (!f) return false;
    fwrite(t.data(), 1, t.size(), f); fclose(f); return true;
}

// ----------------------------------------------------------------------------------------------
// C API (ctypes)
// ----------------------------------------------------------------------------------------------
extern "C" {

void orc_assign_buckets(u64 kmer, int k, u64* out) { assign_buckets(kmer, k, out); }
u64 orc_canonical_kmer(const char* kmer, int k, int* rc) { bool r; u64 v = canonical_kmer((const u8*)kmer, k, &r); *rc = r; return v; }
u64 orc_reverse_complement(u64 v, int k) { return reverse_complement_u64(v, k); }
double orc_students_t_inverse_cdf(double df, double x) { return students_t_inverse_cdf(df, x); }
double orc_thompson_tau(u64 n) { return thompson_tau(n); }

void* orc_index_build(int k, int n_files, const char** paths) {
    Index* ix = build_indexes(k, n_files, paths);
    if (!ix->err.empty()) { fprintf(stderr, "oracle: %s\n", ix->err.c_str()); delete ix; return nullptr; }
    return ix;
}
void* orc_index_load(const char* path, u64* consumed, u64* file_size) {
    std::vector<u8> buf;
    if (!read_file(path, buf)) return nullptr;
    if (file_size) *file_size = buf.size();
    return index_decode(buf, consumed);
}
void* orc_index_decode(const u8* data, u64 n, u64* consumed) {
    std::vector<u8> buf(data, data + n);
    return index_decode(buf, consumed);
}
int orc_index_save(void* h, const char* path) {
    Writer w; index_encode((Index*)h, w);
    FILE* f = fopen(path, "wb"); if (!f) return -1;
    fwrite(w.out.data(), 1, w.out.size(), f); fclose(f); return 0;
}
void orc_index_free(void* h) { delete (Index*)h; }
u64 orc_index_k(void* h) { return ((Index*)h)->k; }
u64 orc_index_meta_k(void* h) { return ((Index*)h)->metadata.k; }
u64 orc_index_n_keys(void* h) { return ((Index*)h)->global_index.size(); }
u64 orc_index_n_entries(void* h) { u64 n = 0; for (auto& kv : ((Index*)h)->global_index) n += kv.second.size(); return n; }
// keys ascending; entry_off has n_keys+1 elements; entries keep their per-key order
void orc_index_export(void* h, u64* keys, u64* entry_off, BucketInfo* entries) {
    Index* ix = (Index*)h;
    std::vector<u64> ks; ks.reserve(ix->global_index.size());
    for (auto& kv : ix->global_index) ks.push_back(kv.first);
    std::sort(ks.begin(), ks.end());
    u64 o = 0;
    for (size_t i = 0; i < ks.size(); i++) {
        keys[i] = ks[i]; entry_off[i] = o;
        for (const BucketInfo& b : ix->global_index[ks[i]]) entries[o++] = b;
    }
    entry_off[ks.size()] = o;
}
u64 orc_index_n_genomes(void* h) { return ((Index*)h)->metadata.files.size(); }
const char* orc_genome_name(void* h, u64 g) { return ((Index*)h)->metadata.files[g].name.c_str(); }
u64 orc_genome_n_seqs(void* h, u64 g) { return ((Index*)h)->metadata.files[g].sequences.size(); }
const char* orc_seq_name(void* h, u64 g, u64 s) { return ((Index*)h)->metadata.files[g].sequences[s].name.c_str(); }
u64 orc_seq_len(void* h, u64 g, u64 s) { return ((Index*)h)->metadata.files[g].sequences[s].len; }
u64 orc_seq_nbases(void* h, u64 g, u64 s) { return ((Index*)h)->metadata.files[g].sequences[s].seq.size(); }
const u8* orc_seq_bases(void* h, u64 g, u64 s) { return ((Index*)h)->metadata.files[g].sequences[s].seq.data(); }

void* orc_fastq_read(const char* path) { return read_fastq(path); }
u64 orc_reads_n(void* h) { return ((Reads*)h)->off.size() - 1; }
u64 orc_reads_nbases(void* h) { return ((Reads*)h)->bases.size(); }
const u8* orc_reads_bases(void* h) { return ((Reads*)h)->bases.data(); }
const u64* orc_reads_off(void* h) { return ((Reads*)h)->off.data(); }
void orc_reads_free(void* h) { delete (Reads*)h; }

void* orc_count(int k, const u8* bases, const u64* off, u64 n_reads, u64 ci, u64 cs, int threads) {
    return count_kmers(k, bases, off, n_reads, ci, cs, threads);
}
u64 orc_counts_n(void* h) { return ((Counts*)h)->kmers.size(); }
void orc_counts_get(void* h, u64* kmers, u64* counts) {
    Counts* c = (Counts*)h;
    memcpy(kmers, c->kmers.data(), c->kmers.size() * 8); memcpy(counts, c->counts.data(), c->counts.size() * 8);
}
void orc_counts_stats(void* h, u64* out4) {
    Counts* c = (Counts*)h; out4[0] = c->total_reads; out4[1] = c->total_kmers; out4[2] = c->unique_kmers; out4[3] = c->unique_counted;
}
// a Counts object from an explicit (k-mer, count) list — lets tests drive map_kmers directly
void* orc_counts_from_list(const u64* kmers, const u64* counts, u64 n) {
    Counts* c = new Counts();
    c->kmers.assign(kmers, kmers + n); c->counts.assign(counts, counts + n); c->unique_counted = n;
    return c;
}
void orc_counts_free(void* h) { delete (Counts*)h; }

// Runs initialize_output_maps → map_kmers (per file) → pick_best_genome(_paired) → call_variants,
// exactly the sequence at call.rs:224-268 (SE) / 314-360 (PE).  Returns NULL if no genome is picked
// (the reference exits 1 there, call.rs:230-233).
void* orc_sample_run(void* index, const Params* pr, int n_files, void** counts) {
    Sample* s = new Sample();
    s->ix = (Index*)index; s->n_files = n_files;
    initialize_output_maps(s);
    for (int f = 0; f < n_files; f++) {
        map_kmers(s, (Counts*)counts[f], *pr, s->stats[f]);
        s->unique_counted[f] = ((Counts*)counts[f])->unique_counted;
    }
    s->best = pick_best_genome(s);
    if (s->best >= 0) call_variants(s, *pr);
    return s;
}
// Stages for the read-sharded protocol tests (tests/test_dist_cpu.py): map only (no selection), then
// call_variants on externally combined pileups of a given genome.
void* orc_sample_map_only(void* index, const Params* pr, int n_files, void** counts) {
    Sample* s = new Sample();
    s->ix = (Index*)index; s->n_files = n_files;
    initialize_output_maps(s);
    for (int f = 0; f < n_files; f++) {
        map_kmers(s, (Counts*)counts[f], *pr, s->stats[f]);
        s->unique_counted[f] = ((Counts*)counts[f])->unique_counted;
    }
    return s;
}
// arrays: 4 x rows(best) x 4 u64 replacing the genome's pileups; stats: per file n_genomes x 4 (as orc_sample_stats)
void orc_sample_call_with(void* h, const Params* pr, int best, const u64* arrays, const u64* stats0, const u64* stats1,
                          u64 unique_counted0, u64 unique_counted1) {
    Sample* s = (Sample*)h;
    const u64 r0 = s->genome_row0[best], rows = s->genome_row0[best + 1] - r0;
    for (int a = 0; a < 4; a++) memcpy(s->arr[a].data() + r0, arrays + (size_t)a * rows * 4, rows * 32);
    const u64* st[2] = {stats0, stats1};
    for (int f = 0; f < s->n_files; f++)
        for (size_t g = 0; g < s->stats[f].size(); g++) {
            s->stats[f][g].perfect = st[f][g * 4]; s->stats[f][g].variant = st[f][g * 4 + 1];
            s->stats[f][g].unique = st[f][g * 4 + 2]; s->stats[f][g].present = st[f][g * 4 + 3] != 0;
        }
    s->unique_counted[0] = unique_counted0; s->unique_counted[1] = unique_counted1;
    s->best = best;
    if (best >= 0) call_variants(s, *pr);
}
int orc_pick_best(void* h) { return pick_best_genome((Sample*)h); }
void orc_sample_free(void* h) { delete (Sample*)h; }
int orc_sample_best(void* h) { return ((Sample*)h)->best; }
// out[g*4 + {0,1,2,3}] = perfect, variant, unique, present
void orc_sample_stats(void* h, int file, u64* out) {
    Sample* s = (Sample*)h;
    for (size_t g = 0; g < s->stats[file].size(); g++) {
        out[g * 4] = s->stats[file][g].perfect; out[g * 4 + 1] = s->stats[file][g].variant;
        out[g * 4 + 2] = s->stats[file][g].unique; out[g * 4 + 3] = s->stats[file][g].present;
    }
}
u64 orc_sample_genome_rows(void* h, int g) { Sample* s = (Sample*)h; return s->genome_row0[g + 1] - s->genome_row0[g]; }
// arr: 0 fwd depth, 1 rev depth, 2 fwd support, 3 rev support; out = rows*4 u64
void orc_sample_pileup(void* h, int g, int arr, u64* out) {
    Sample* s = (Sample*)h;
    u64 r0 = s->genome_row0[g], r1 = s->genome_row0[g + 1];
    memcpy(out, s->arr[arr].data() + r0, (r1 - r0) * 32);
}
u64 orc_sample_n_variants(void* h) { return ((Sample*)h)->variants.size(); }
void orc_sample_variants(void* h, VCFRecord* out) { Sample* s = (Sample*)h; memcpy(out, s->variants.data(), s->variants.size() * sizeof(VCFRecord)); }
void orc_sample_summary(void* h, u64* major, u64* minor, double* breadth, double* depth) {
    Sample* s = (Sample*)h; *major = s->num_major; *minor = s->num_minor; *breadth = s->breadth; *depth = s->depth_cov;
}
void orc_sample_noise_max(void* h, double* out) { Sample* s = (Sample*)h; memcpy(out, s->noise_max.data(), s->noise_max.size() * 8); }
// num_unmapped as at call.rs:242 / 336 (usize arithmetic; wraps like release Rust if negative)
u64 orc_sample_unmapped(void* h) {
    Sample* s = (Sample*)h; u64 uc = 0, pv = 0;
    for (int f = 0; f < s->n_files; f++) { uc += s->unique_counted[f]; pv += s->stats[f][s->best].perfect + s->stats[f][s->best].variant; }
    return uc - pv;
}
int orc_write_vcf(void* h, const char* reads_path, const char* out_path) { return write_text(out_path, vcf_text((Sample*)h, reads_path)) ? 0 : -1; }
int orc_write_pileup(void* h, const char* out_path) { return write_text(out_path, pileup_text((Sample*)h)) ? 0 : -1; }
// text into caller buffer; returns needed size
u64 orc_vcf_text(void* h, const char* reads_path, char* buf, u64 cap) {
    std::string t = vcf_text((Sample*)h, reads_path);
    if (buf && cap) { u64 n = std::min<u64>(cap - 1, t.size()); memcpy(buf, t.data(), n); buf[n] = 0; }
    return t.size() + 1;
}
u64 orc_pileup_text(void* h, char* buf, u64 cap) {
    std::string t = pileup_text((Sample*)h);
    if (buf && cap) { u64 n = std::min<u64>(cap - 1, t.size()); memcpy(buf, t.data(), n); buf[n] = 0; }
    return t.size() + 1;
}
u64 orc_clean_sample_id(const char* path, char* buf, u64 cap) {
    std::string t = clean_sample_id(path);
    if (buf && cap) { u64 n = std::min<u64>(cap - 1, t.size()); memcpy(buf, t.data(), n); buf[n] = 0; }
    return t.size() + 1;
}
// stand-alone noise (tests): fwd/rev are len*4 u64; out_max/out_mean/out_std are len doubles
int orc_baseline_noise(const u64* fwd, const u64* rev, u64 len, double* out_max, double* out_mean, double* out_std) {
    std::vector<Noise> o;
    bool ok = get_baseline_noise((const std::array<u64, 4>*)fwd, (const std::array<u64, 4>*)rev, len, o);
    for (u64 i = 0; i < len; i++) { out_max[i] = o[i].max; if (out_mean) out_mean[i] = o[i].mean; if (out_std) out_std[i] = o[i].std; }
    return ok ? 0 : -1;
}

}  // extern "C"
