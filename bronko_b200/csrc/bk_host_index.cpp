// bk_host_index.cpp — .bkdb (bincode 2) reader/writer, FASTA reader, the
// host index builder (build_indexes, reference src/build.rs:145-231) and the tables the device
// kernels need, derived once per index load.
#include "bk_host.h"

#include <algorithm>
#include <atomic>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <thread>
#include <zlib.h>

namespace bk {

// ---------------------------------------------------------------------------------------------
// lcb.rs primitives, closed form.  u64 arithmetic wraps on purpose: for
// k = 31 the reference's ids exceed 2^64 and wrap too (release build), src/lcb.rs:8-41.
// ---------------------------------------------------------------------------------------------
void assign_buckets_host(u64 kmer, int k, u64* out) {
    u64 mu[32], val[32], cur[32];
    u32 zero_before[32];
    u64 sum_mu = 0;
    u32 zeros = 0;
    for (int i = 0; i < k; i++) {
        const int sh = 2 * (k - 1 - i);
        const u64 w = 1ull << sh;
        const u64 d = (kmer >> sh) & 3;
        cur[i] = d << sh;
        val[i] = kmer & (w - 1);
        mu[i] = d ? w + (cur[i] >> 2) * (u64)(k - 1 - i) : val[i];
        zero_before[i] = zeros;
        zeros += (d == 0);
        sum_mu += mu[i];
    }
    for (int i = 0; i < k; i++)
        out[i] = sum_mu - mu[i] + val[i] - (u64)zero_before[i] * cur[i] + 1 + zero_before[i];
}

u64 revcomp_host(u64 v, int k) {
    u64 x = ~v;
    x = ((x >> 2) & 0x3333333333333333ull) | ((x & 0x3333333333333333ull) << 2);
    x = ((x >> 4) & 0x0F0F0F0F0F0F0F0Full) | ((x & 0x0F0F0F0F0F0F0F0Full) << 4);
    x = __builtin_bswap64(x);
    return x >> (64 - 2 * k);
}

// ---------------------------------------------------------------------------------------------
// bincode varint
// ---------------------------------------------------------------------------------------------
namespace {
struct Cursor {
    const u8* p; const u8* end; bool ok;
    u8 byte() { if (p >= end) { ok = false; return 0; } return *p++; }
    u64 varint() {
        u8 t = byte();
        if (t < 251) return t;
        int n = t == 251 ? 2 : t == 252 ? 4 : t == 253 ? 8 : 0;
        if (!n) { ok = false; return 0; }
        if (end - p < n) { ok = false; return 0; }
        u64 v = 0;
        memcpy(&v, p, n);   // little endian host
        p += n;
        return v;
    }
    bool bytes(u64 n, const u8** out) {
        if ((u64)(end - p) < n) { ok = false; return false; }
        *out = p; p += n; return true;
    }
};
struct Sink {
    std::vector<u8> buf;
    void byte(u8 b) { buf.push_back(b); }
    void varint(u64 v) {
        if (v < 251) { buf.push_back((u8)v); return; }
        int n = v <= 0xFFFFull ? 2 : v <= 0xFFFFFFFFull ? 4 : 8;
        buf.push_back(n == 2 ? 251 : n == 4 ? 252 : 253);
        for (int i = 0; i < n; i++) buf.push_back((u8)(v >> (8 * i)));
    }
    void blob(const void* p, u64 n) { varint(n); buf.insert(buf.end(), (const u8*)p, (const u8*)p + n); }
};
}  // namespace

// ---------------------------------------------------------------------------------------------
// host threads: a 200-strain database is 1.25e8 (bucket id, entry) pairs — one stable sort of them on one thread was
// 20 of the 28 s of building or loading it
// ---------------------------------------------------------------------------------------------
static unsigned host_threads() {
    const unsigned hw = std::thread::hardware_concurrency();
    return std::min(16u, std::max(1u, hw));
}
// fn(i) for i in [0, n) on up to host_threads() threads (dynamic distribution)
static void parallel_for(size_t n, const std::function<void(size_t)>& fn) {
    const unsigned T = (unsigned)std::min<size_t>(host_threads(), n);
    if (T <= 1) { for (size_t i = 0; i < n; i++) fn(i); return; }
    std::atomic<size_t> next{0};
    std::vector<std::thread> th;
    for (unsigned t = 0; t < T; t++)
        th.emplace_back([&] { for (size_t i = next.fetch_add(1); i < n; i = next.fetch_add(1)) fn(i); });
    for (std::thread& t : th) t.join();
}
// stable sort: chunks sorted in parallel, then rounds of pairwise std::merge (stable: ties keep the left run first)
template <class T, class Less>
static void parallel_stable_sort(std::vector<T>& v, Less less) {
    const size_t n = v.size();
    unsigned P = host_threads();
    while (P & (P - 1)) P &= P - 1;                           // power of two
    if (n < (1u << 20) || P == 1) { std::stable_sort(v.begin(), v.end(), less); return; }
    std::vector<size_t> cut(P + 1);
    for (unsigned i = 0; i <= P; i++) cut[i] = n / P * i + std::min<size_t>(i, n % P);
    parallel_for(P, [&](size_t i) { std::stable_sort(v.begin() + cut[i], v.begin() + cut[i + 1], less); });
    std::vector<T> tmp(n);
    T* src = v.data();
    T* dst = tmp.data();
    for (unsigned width = 1; width < P; width *= 2) {
        parallel_for(P / (2 * width), [&](size_t j) {
            const size_t a = cut[2 * width * j], m = cut[2 * width * j + width], b = cut[2 * width * j + 2 * width];
            std::merge(src + a, src + m, src + m, src + b, dst + a, less);
        });
        std::swap(src, dst);
    }
    if (src != v.data()) v.swap(tmp);
}

void index_from_pairs(HostIndex& ix, std::vector<KeyedEntry>& pairs) {
    parallel_stable_sort(pairs, [](const KeyedEntry& a, const KeyedEntry& b) { return a.key < b.key; });
    ix.keys.clear(); ix.entry_off.clear(); ix.entries.clear();
    ix.entries.reserve(pairs.size());
    for (size_t i = 0; i < pairs.size(); i++) {
        if (i == 0 || pairs[i].key != pairs[i - 1].key) { ix.keys.push_back(pairs[i].key); ix.entry_off.push_back(i); }
        ix.entries.push_back(pairs[i].e);
    }
    ix.entry_off.push_back(pairs.size());
}

bool bkdb_decode(const u8* data, u64 n, HostIndex& ix, std::string& err) {
    Cursor c{data, data + n, true};
    ix = HostIndex();
    ix.k = (u32)c.varint();
    const u64 n_keys = c.varint();
    // the file holds the map in hash-table order: entries are kept as they come and only one record per key is sorted
    struct KeyRec { u64 key; u64 first; u64 count; };
    std::vector<KeyRec> recs;
    std::vector<bk_bucket_info> raw;
    recs.reserve((size_t)std::min<u64>(n_keys, n / 2));
    raw.reserve((size_t)std::min<u64>(n_keys + n_keys / 8, n / 5));
    for (u64 i = 0; i < n_keys && c.ok; i++) {
        const u64 key = c.varint();
        const u64 m = c.varint();
        recs.push_back(KeyRec{key, (u64)raw.size(), 0});
        for (u64 j = 0; j < m && c.ok; j++) {
            bk_bucket_info e; memset(&e, 0, sizeof e);
            e.file_id = (u16)c.varint();
            e.seq_id = c.byte();
            e.location = (u32)c.varint();
            e.idx = c.byte();
            e.canonical = c.byte();
            raw.push_back(e);
        }
        recs.back().count = (u64)raw.size() - recs.back().first;
    }
    const u64 n_files = c.varint();
    for (u64 f = 0; f < n_files && c.ok; f++) {
        HostGenome g;
        const u8* s; u64 len = c.varint();
        if (!c.bytes(len, &s)) break;
        g.name.assign((const char*)s, len);
        const u64 n_seq = c.varint();
        for (u64 q = 0; q < n_seq && c.ok; q++) {
            HostSeq hs;
            len = c.varint();
            if (!c.bytes(len, &s)) break;
            hs.name.assign((const char*)s, len);
            hs.len = c.varint();
            len = c.varint();
            if (!c.bytes(len, &s)) break;
            hs.bases.assign(s, s + len);
            g.seqs.push_back(std::move(hs));
        }
        ix.genomes.push_back(std::move(g));
    }
    ix.meta_k = c.varint();
    if (!c.ok) { err = "truncated or malformed .bkdb"; return false; }
    // a map never holds one key twice; should a file do so anyway, the stable sort keeps the file's order and the
    // entries of equal keys are joined (what sorting the (key, entry) pairs themselves would give)
    parallel_stable_sort(recs, [](const KeyRec& a, const KeyRec& b) { return a.key < b.key; });
    ix.keys.clear(); ix.entry_off.clear(); ix.entries.clear();
    ix.entries.reserve(raw.size());
    for (size_t i = 0; i < recs.size(); i++) {
        if (i == 0 || recs[i].key != recs[i - 1].key) { ix.keys.push_back(recs[i].key); ix.entry_off.push_back(ix.entries.size()); }
        ix.entries.insert(ix.entries.end(), raw.begin() + recs[i].first, raw.begin() + recs[i].first + recs[i].count);
    }
    ix.entry_off.push_back(ix.entries.size());
    return true;
}

bool slurp_maybe_gz(const std::string& path, std::string& out) {
    gzFile g = gzopen(path.c_str(), "rb");
    if (!g) return false;
    gzbuffer(g, 1 << 20);
    std::vector<char> buf(1 << 20);
    int n;
    while ((n = gzread(g, buf.data(), (unsigned)buf.size())) > 0) out.append(buf.data(), n);
    gzclose(g);
    return n == 0;
}

bool bkdb_read(const std::string& path, HostIndex& ix, std::string& err) {
    FILE* f = fopen(path.c_str(), "rb");
    if (!f) { err = "Failed to open file '" + path + "'"; return false; }
    fseek(f, 0, SEEK_END);
    long n = ftell(f);
    fseek(f, 0, SEEK_SET);
    std::vector<u8> buf(n > 0 ? n : 0);
    size_t got = n > 0 ? fread(buf.data(), 1, n, f) : 0;
    fclose(f);
    if ((long)got != n) { err = "short read on '" + path + "'"; return false; }
    if (!bkdb_decode(buf.data(), buf.size(), ix, err)) { err = "Failed to read Bronko Index from '" + path + "': " + err; return false; }
    return true;
}

bool bkdb_write(const std::string& path, const HostIndex& ix, std::string& err) {
    Sink s;
    s.varint(ix.k);
    s.varint(ix.keys.size());
    for (size_t i = 0; i < ix.keys.size(); i++) {
        s.varint(ix.keys[i]);
        s.varint(ix.entry_off[i + 1] - ix.entry_off[i]);
        for (u64 j = ix.entry_off[i]; j < ix.entry_off[i + 1]; j++) {
            const bk_bucket_info& e = ix.entries[j];
            s.varint(e.file_id); s.byte(e.seq_id); s.varint(e.location); s.byte(e.idx); s.byte(e.canonical);
        }
    }
    s.varint(ix.genomes.size());
    for (const HostGenome& g : ix.genomes) {
        s.blob(g.name.data(), g.name.size());
        s.varint(g.seqs.size());
        for (const HostSeq& q : g.seqs) {
            s.blob(q.name.data(), q.name.size());
            s.varint(q.len);
            s.blob(q.bases.data(), q.bases.size());
        }
    }
    s.varint(ix.meta_k);
    FILE* f = fopen(path.c_str(), "wb");
    if (!f) { err = "File path " + path + " not valid"; return false; }
    size_t w = fwrite(s.buf.data(), 1, s.buf.size(), f);
    fclose(f);
    if (w != s.buf.size()) { err = "short write on " + path; return false; }
    return true;
}

std::string first_token(const std::string& s) {
    size_t a = 0;
    while (a < s.size() && isspace((unsigned char)s[a])) a++;
    size_t b = a;
    while (b < s.size() && !isspace((unsigned char)s[b])) b++;
    return s.substr(a, b - a);
}

static std::string path_stem(const std::string& path) {
    size_t sl = path.find_last_of('/');
    std::string f = sl == std::string::npos ? path : path.substr(sl + 1);
    size_t dot = f.find_last_of('.');
    return (dot == std::string::npos || dot == 0) ? f : f.substr(0, dot);
}

// FASTA text → records; header = text after '>' up to end of line, sequence lines concatenated
// with CR/LF removed (needletail semantics, reference src/build.rs:156-189).
bool index_build_from_fasta(u32 k, const std::vector<std::string>& paths, HostIndex& ix, std::string& err) {
    ix = HostIndex();
    ix.k = k; ix.meta_k = k;
    // one file = one genome (build.rs:145-231): the files are parsed and their (bucket id, entry) pairs generated on
    // several threads, then joined in file order — the order the reference merges its per-file maps in (223-227)
    struct FileOut { HostGenome g; std::vector<KeyedEntry> pairs; bool ok = true; std::string why; };
    std::vector<FileOut> outs(paths.size());
    parallel_for(paths.size(), [&](size_t file_id) {
        FileOut& o = outs[file_id];
        u64 ids[32];
        std::string txt;
        if (!slurp_maybe_gz(paths[file_id], txt)) { o.ok = false; o.why = "Failed to open or inflate the file"; return; }
        // needletail (parse_fastx_file, src/build.rs:156-159) rejects an empty file and a file that does not start with
        // a record marker; the reference logs "<error> | Failed to parse fasta file: <path>" and exits
        if (txt.empty()) { o.ok = false; o.why = "Failed to read the first two bytes. Is the file empty?"; return; }
        if (txt[0] != '>') { o.ok = false; o.why = std::string("Expected '>' at the start of the file but found '") + txt[0] + "'"; return; }
        HostGenome& g = o.g;
        g.name = path_stem(paths[file_id]);
        size_t pos = 0;
        while (pos < txt.size()) {
            size_t eol = txt.find('\n', pos);
            if (eol == std::string::npos) eol = txt.size();
            size_t le = eol;
            if (le > pos && txt[le - 1] == '\r') le--;
            if (le > pos && txt[pos] == '>') {
                HostSeq q;
                q.name = first_token(txt.substr(pos + 1, le - pos - 1));
                g.seqs.push_back(std::move(q));
            } else if (!g.seqs.empty()) {
                g.seqs.back().bases.insert(g.seqs.back().bases.end(), txt.begin() + pos, txt.begin() + le);
            }
            pos = eol + 1;
        }
        u8 seq_id = 0;   // u8 counter wraps after 255 sequences exactly as build.rs:170,207
        for (HostSeq& q : g.seqs) {
            q.len = q.bases.size();
            const u64 L = q.len;
            // src/build.rs:191-192: `for i in 0..=seq_len.saturating_sub(k) { &seq[i..i + k] }` panics for a sequence
            // shorter than k; here the build fails with a message instead
            if (L < k) { o.ok = false; o.why = "sequence '" + q.name + "' is shorter than k (the reference panics on it)"; return; }
            if (L >= k) {
                const u64 kmask = (1ull << (2 * k)) - 1;
                u64 fwd = 0;
                o.pairs.reserve(o.pairs.size() + (size_t)(L - k + 1) * k);
                for (u64 i = 0; i < L; i++) {
                    fwd = ((fwd << 2) | nt_to_bits_host(q.bases[i])) & kmask;
                    if (i + 1 < k) continue;
                    const u64 start = i + 1 - k;
                    const u64 rev = revcomp_host(fwd, (int)k);
                    const bool rc = !(fwd < rev);               // lcb.rs:87-95
                    assign_buckets_host(rc ? rev : fwd, (int)k, ids);
                    for (u32 j = 0; j < k; j++) {
                        KeyedEntry ke; memset(&ke, 0, sizeof ke);
                        ke.key = ids[j];
                        ke.e.file_id = (u16)file_id; ke.e.seq_id = seq_id; ke.e.location = (u32)start;
                        ke.e.idx = (u8)j; ke.e.canonical = rc;
                        o.pairs.push_back(ke);
                    }
                }
            }
            seq_id++;
        }
    });
    size_t total = 0;
    for (size_t file_id = 0; file_id < paths.size(); file_id++) {
        if (!outs[file_id].ok) { err = outs[file_id].why + " | Failed to parse fasta file: " + paths[file_id]; return false; }
        total += outs[file_id].pairs.size();
    }
    std::vector<KeyedEntry> pairs;
    pairs.reserve(total);
    for (FileOut& o : outs) {
        pairs.insert(pairs.end(), o.pairs.begin(), o.pairs.end());
        std::vector<KeyedEntry>().swap(o.pairs);
        ix.genomes.push_back(std::move(o.g));
    }
    index_from_pairs(ix, pairs);
    return true;
}

// ---------------------------------------------------------------------------------------------
// Derived device tables
// ---------------------------------------------------------------------------------------------
static u32 log2_cap(u64 n_items) {
    u32 l = 10;
    while ((1ull << l) < 2 * n_items + 16) l++;
    return l;
}

void derive_index(const HostIndex& ix, DerivedIndex& d, bool allow_rekey) {
    d = DerivedIndex();
    d.k = ix.k;
    const u32 k = ix.k;
    d.n_genomes = (u32)ix.genomes.size();
    u32 row = 0, seq = 0;
    for (const HostGenome& g : ix.genomes) {
        d.genome_row0.push_back(row);
        d.genome_seq_off.push_back(seq);
        u64 glen = 0;
        for (const HostSeq& q : g.seqs) {
            d.seq_row0.push_back(row);
            for (u64 i = 0; i < q.bases.size(); i++) d.ref_code.push_back(nt_to_bits_host(q.bases[i]));
            row += (u32)q.bases.size();
            glen += q.len;
            seq++;
        }
        d.genome_len.push_back(glen);
        d.max_genome_rows = std::max(d.max_genome_rows, row - d.genome_row0.back());
    }
    d.genome_row0.push_back(row);
    d.genome_seq_off.push_back(seq);
    d.seq_row0.push_back(row);
    d.n_seqs = seq;

    // bucket table
    d.bucket_log2 = log2_cap(ix.keys.size());
    d.bucket_slots.assign(1ull << d.bucket_log2, BucketSlot{~0ull, 0, 0});
    d.bucket_entries.resize(ix.entries.size());
    const size_t ECHUNK = 1u << 20;
    parallel_for((ix.entries.size() + ECHUNK - 1) / ECHUNK, [&](size_t c) {
    for (size_t i = c * ECHUNK, i_end = std::min(ix.entries.size(), (c + 1) * ECHUNK); i < i_end; i++) {
        const bk_bucket_info& e = ix.entries[i];
        BucketEntry be;
        u32 r = 0xFFFFFFFFu;
        if (e.file_id < d.n_genomes) {
            const u32 s0 = d.genome_seq_off[e.file_id], s1 = d.genome_seq_off[e.file_id + 1];
            if (s0 + e.seq_id < s1) {
                const u32 sr0 = d.seq_row0[s0 + e.seq_id], sr1 = d.seq_row0[s0 + e.seq_id + 1];
                // the reference would index out of bounds (panic) on such an entry; it is skipped here
                if ((u64)sr0 + e.location + e.idx < sr1) r = sr0 + e.location;
            }
        }
        be.row = r; be.file_id = e.file_id; be.idx = e.idx; be.canonical = e.canonical ? 1 : 0;
        d.bucket_entries[i] = be;
    }
    });
    for (size_t i = 0; i + 1 < ix.entry_off.size(); i++) d.max_key_entries = std::max<u32>(d.max_key_entries, (u32)std::min<u64>(ix.entry_off[i + 1] - ix.entry_off[i], 0xFFFFFFFFull));
    // Re-key: bucket id j of a canonical k-mer is a perfect rank of (j, k-mer without digit j), so the device can probe
    // with that pair directly and skip the id arithmetic.  Every key is verified against its first entry (the
    // reference k-mer at `location`, canonicalised, digit idx zeroed must rank to exactly this key); one failure (an
    // index not produced by `bronko build`) keeps the id-keyed table.
    std::vector<u64> slot_key(ix.keys.begin(), ix.keys.end());
    d.rekeyed = allow_rekey && k <= 29;
    const bool want_groups = d.n_genomes <= 4;                // the grouped form serves the thread-per-k-mer map kernels only
    std::vector<u64> center;                                  // per key: the canonical k-mer its first entry was verified with
    if (d.rekeyed) {
        if (want_groups) center.resize(ix.keys.size());
        std::atomic<bool> failed{false};
        const size_t KCHUNK = 1u << 18;
        parallel_for((ix.keys.size() + KCHUNK - 1) / KCHUNK, [&](size_t c) {
            u64 ids[32];
            for (size_t i = c * KCHUNK, i_end = std::min(ix.keys.size(), (c + 1) * KCHUNK); i < i_end && !failed.load(std::memory_order_relaxed); i++) {
                if (ix.entry_off[i + 1] == ix.entry_off[i]) { failed = true; break; }
                const bk_bucket_info& e = ix.entries[ix.entry_off[i]];
                bool ok = e.file_id < d.n_genomes && e.idx < k;
                const HostSeq* q = nullptr;
                if (ok) { const HostGenome& g = ix.genomes[e.file_id]; ok = e.seq_id < g.seqs.size(); if (ok) q = &g.seqs[e.seq_id]; }
                if (ok) ok = (u64)e.location + k <= q->bases.size();
                if (!ok) { failed = true; break; }
                u64 fwd = 0;
                for (u32 b = 0; b < k; b++) fwd = (fwd << 2) | nt_to_bits_host(q->bases[e.location + b]);
                const u64 rev = revcomp_host(fwd, (int)k);
                const u64 kb = fwd < rev ? fwd : rev;                       // src/lcb.rs:87-95
                assign_buckets_host(kb, (int)k, ids);
                if (ids[e.idx] != ix.keys[i]) { failed = true; break; }
                slot_key[i] = ((u64)e.idx << 58) | (kb & ~(3ull << (2 * (k - 1 - e.idx))));
                if (want_groups) center[i] = kb;
            }
        });
        if (failed) { d.rekeyed = false; slot_key.assign(ix.keys.begin(), ix.keys.end()); }
    }
    const u64 bmask = (1ull << d.bucket_log2) - 1;
    for (size_t i = 0; i < ix.keys.size(); i++) {
        u64 h = hash_slot_host(slot_key[i], 64 - d.bucket_log2);
        while (d.bucket_slots[h].key != ~0ull) h = (h + 1) & bmask;
        d.bucket_slots[h] = BucketSlot{slot_key[i], (u32)ix.entry_off[i], (u32)(ix.entry_off[i + 1] - ix.entry_off[i])};
    }

    if (d.rekeyed && want_groups) {           // grouped form (bk_host.h)
        const u32 mid = k / 2;
        d.group_mid = mid;
        const u32 lo_bits = 2 * (k - mid);                                   // digits mid..k-1
        struct Cen { u64 g, c; };
        std::vector<Cen> cens(ix.keys.size());
        for (size_t i = 0; i < ix.keys.size(); i++) {
            const u32 idx = (u32)(slot_key[i] >> 58);
            const u64 c = center[i];
            const u64 half = idx < mid ? (c & ((1ull << lo_bits) - 1)) : (c >> lo_bits);
            cens[i] = Cen{((u64)(idx < mid ? 0 : 1) << 62) | half, c};
        }
        std::sort(cens.begin(), cens.end(), [](const Cen& a, const Cen& b) { return a.g != b.g ? a.g < b.g : a.c < b.c; });
        cens.erase(std::unique(cens.begin(), cens.end(), [](const Cen& a, const Cen& b) { return a.g == b.g && a.c == b.c; }), cens.end());
        size_t n_groups = 0;
        for (size_t i = 0; i < cens.size(); i++) if (i == 0 || cens[i].g != cens[i - 1].g) n_groups++;
        d.group_log2 = log2_cap(n_groups);
        d.group_slots.assign(1ull << d.group_log2, BucketSlot{~0ull, 0, 0});
        d.group_centers.resize(cens.size());
        d.group_buckets.clear();
        const u64 gmask = (1ull << d.group_log2) - 1;
        auto find_bucket = [&](u64 key, u32* off, u32* len) {            // the probe sequence of the device
            u64 h = hash_slot_host(key, 64 - d.bucket_log2);
            for (;;) {
                const BucketSlot& sl = d.bucket_slots[h];
                if (sl.key == key) { *off = sl.off; *len = sl.len; return true; }
                if (sl.key == ~0ull) return false;
                h = (h + 1) & bmask;
            }
        };
        for (size_t i = 0; i < cens.size();) {
            size_t j = i;
            while (j < cens.size() && cens[j].g == cens[i].g) {
                // the buckets of this center on this side: index i2 in [lo, hi) → (off, len) of bucket (i2, center without digit i2)
                const u32 side = (u32)(cens[j].g >> 62), lo = side ? mid : 0, hi = side ? k : mid;
                const u32 boff = (u32)d.group_buckets.size();
                u32 present = 0;
                for (u32 i2 = lo; i2 < hi; i2++) {
                    u32 off = 0, len = 0;
                    if (find_bucket(((u64)i2 << 58) | (cens[j].c & ~(3ull << (2 * (k - 1 - i2)))), &off, &len)) present |= 1u << i2;
                    d.group_buckets.push_back(OffLen{off, len});
                }
                d.group_centers[j] = BucketSlot{cens[j].c, boff, present};
                j++;
            }
            u64 h = hash_slot_host(cens[i].g, 64 - d.group_log2);
            while (d.group_slots[h].key != ~0ull) h = (h + 1) & gmask;
            d.group_slots[h] = BucketSlot{cens[i].g, (u32)i, (u32)(j - i)};
            i = j;
        }
    }

    // oriented reference store
    struct Occ { u64 kmer; u32 gidx; u32 oseq; };
    std::vector<Occ> occ;
    u32 g_base = REF_PAD_BASES;
    std::vector<u8> codes;   // global base index → 2-bit code
    codes.assign(REF_PAD_BASES, 0);
    const u64 kmask = (1ull << (2 * k)) - 1;
    for (const HostGenome& g : ix.genomes) {
        for (const HostSeq& q : g.seqs) {
            const u32 L = (u32)q.bases.size();
            for (int orient = 0; orient < 2; orient++) {
                const u32 oseq = (u32)d.oseq_start.size();
                d.oseq_start.push_back(g_base);
                d.oseq_len.push_back(L);
                u64 cur = 0;
                for (u32 i = 0; i < L; i++) {
                    const u8 c = orient == 0 ? nt_to_bits_host(q.bases[i]) : (u8)(3 ^ nt_to_bits_host(q.bases[L - 1 - i]));
                    codes.push_back(c);
                    cur = ((cur << 2) | c) & kmask;
                    if (i + 1 >= k) occ.push_back(Occ{cur, g_base + i + 1 - k, oseq});
                }
                g_base += L;
            }
        }
    }
    codes.resize(codes.size() + REF_PAD_BASES, 0);
    d.n_raw = (u32)codes.size();
    d.refpk.assign((codes.size() + 31) / 32 + 2, 0);
    for (size_t i = 0; i < codes.size(); i++) d.refpk[i >> 5] |= (u64)codes[i] << (62 - 2 * (i & 31));
    // 4-bit copy for the scan kernel (little-endian nibbles, base i in nibble i); padding nibbles are 4, which the
    // kernel's byte-permute turns into '#', a byte no case-folded read byte can equal
    const size_t n_chunks = (codes.size() + 31) / 32 + 2;
    d.refnib.assign(n_chunks * 4, 0x44444444u);
    for (size_t i = REF_PAD_BASES; i + REF_PAD_BASES < codes.size(); i++) {
        u32& wd = d.refnib[i >> 3];
        const u32 sh = 4 * (u32)(i & 7);
        wd = (wd & ~(0xFu << sh)) | ((u32)codes[i] << sh);
    }

    parallel_stable_sort(occ, [](const Occ& a, const Occ& b) { return a.kmer != b.kmer ? a.kmer < b.kmer : a.gidx < b.gidx; });      // (a total order)
    d.slot2id.assign(d.n_raw, 0xFFFFFFFFu);
    for (size_t i = 0; i < occ.size(); i++) {
        if (i == 0 || occ[i].kmer != occ[i - 1].kmer) d.id_kmer.push_back(occ[i].kmer);
        d.slot2id[occ[i].gidx] = (u32)d.id_kmer.size() - 1;
    }
    d.exact_log2 = log2_cap(d.id_kmer.size());
    d.exact_slots.assign(1ull << d.exact_log2, ExactSlot{~0ull, 0, 0});
    const u64 emask = (1ull << d.exact_log2) - 1;
    for (size_t i = 0; i < occ.size(); i++) {
        if (i != 0 && occ[i].kmer == occ[i - 1].kmer) continue;
        u64 h = hash_slot_host(occ[i].kmer, 64 - d.exact_log2);
        while (d.exact_slots[h].key != ~0ull) h = (h + 1) & emask;
        d.exact_slots[h] = ExactSlot{occ[i].kmer, occ[i].gidx, occ[i].oseq};
    }

    // ---- mismatch lines (bk_host.h) ----
    const size_t n_ids = d.id_kmer.size();
    d.dense_ok = k <= 29 && n_ids > 0 && n_ids <= 400000 && (u64)d.n_raw * 4 * (k + 1) * 4 <= (512ull << 20) && getenv("BK_NO_DENSE") == nullptr;
    if (!d.dense_ok) return;
    {
        d.id_rep.assign(n_ids, 0);
        for (size_t i = 0; i < occ.size(); i++) if (i == 0 || occ[i].kmer != occ[i - 1].kmer) d.id_rep[d.slot2id[occ[i].gidx]] = occ[i].gidx;
        d.slot2rep.assign(d.n_raw, 0xFFFFFFFFu);
        for (u32 sl = 0; sl < d.n_raw; sl++) if (d.slot2id[sl] != 0xFFFFFFFFu) d.slot2rep[sl] = d.id_rep[d.slot2id[sl]];
    }
    // pairs of reference k-mers within Hamming distance 2 share at least one third of their digits exactly: group the
    // ids by each third in turn and compare inside the groups
    d.id_amb.assign(n_ids, 0);
    {
        const u32 cut[4] = {0, k / 3, 2 * k / 3, k};
        std::vector<std::pair<u64, u32>> byp(n_ids);
        for (u32 part = 0; part < 3; part++) {
            const u32 d0 = cut[part], d1 = cut[part + 1];                        // digits [d0, d1), digit 0 = most significant
            const u32 sh = 2 * (k - d1);
            const u64 mask = ((1ull << (2 * (d1 - d0))) - 1);
            for (size_t i = 0; i < n_ids; i++) byp[i] = std::make_pair((d.id_kmer[i] >> sh) & mask, (u32)i);
            std::sort(byp.begin(), byp.end());
            for (size_t a = 0; a < n_ids;) {
                size_t b = a;
                while (b < n_ids && byp[b].first == byp[a].first) b++;
                if (b - a > 4096) {                                              // low complexity: do not compare 10^7 pairs, give up on the group
                    for (size_t x = a; x < b; x++) d.id_amb[byp[x].second] = (1u << k) - 1u;
                } else {
                    for (size_t x = a; x < b; x++)
                        for (size_t y = x + 1; y < b; y++) {
                            const u64 df = d.id_kmer[byp[x].second] ^ d.id_kmer[byp[y].second];
                            u64 nz = (df | (df >> 1)) & 0x5555555555555555ull;
                            const int cnt = __builtin_popcountll(nz);
                            if (cnt < 1 || cnt > 2) continue;
                            while (nz) {
                                const u32 bit = (u32)__builtin_ctzll(nz);
                                nz &= nz - 1;
                                const u32 j = k - 1 - bit / 2;
                                d.id_amb[byp[x].second] |= 1u << j;
                                d.id_amb[byp[y].second] |= 1u << j;
                            }
                        }
                }
                a = b;
            }
        }
    }
    d.nb_log2 = log2_cap(n_ids * k);
    d.nb_slots.assign(1ull << d.nb_log2, ExactSlot{~0ull, 0, 0});
    d.nb_bloom_log2 = d.nb_log2 > 8 ? d.nb_log2 - 2 : 6;                     // 64-bit words: slots >= 2 x keys → >= 32 bits per key, three of them set
    d.nb_bloom.assign(1ull << d.nb_bloom_log2, 0);
    const u64 nmask = (1ull << d.nb_log2) - 1;
    for (size_t i = 0; i < n_ids; i++)
        for (u32 j = 0; j < k; j++) {
            const u64 key = ((u64)j << 58) | (d.id_kmer[i] & ~(3ull << (2 * (k - 1 - j))));
            u64 h = hash_slot_host(key, 64 - d.nb_log2);
            while (d.nb_slots[h].key != ~0ull && d.nb_slots[h].key != key) h = (h + 1) & nmask;
            if (d.nb_slots[h].key == ~0ull) d.nb_slots[h] = ExactSlot{key, (u32)i, 0};      // (a second id with this key: both are flagged at j)
            const u64 hb = hash64_host(key);
            d.nb_bloom[hb >> (64 - d.nb_bloom_log2)] |= bloom_mask(hb, d.nb_bloom_log2);
        }
    // per line: which cells are ambiguous / must be folded (most lines: none — the kernels skip them without reading the row)
    d.line_amb.assign(d.n_raw, 0); d.line_fold.assign(d.n_raw, 0);
    for (u32 r = 0; r < d.n_raw; r++)
        for (u32 j = 0; j < k && j <= r; j++) {
            const u32 sl = r - j;
            if (d.slot2id[sl] == 0xFFFFFFFFu) continue;
            if ((d.id_amb[d.slot2id[sl]] >> j) & 1u) d.line_amb[r] |= 1u << j;
            if (d.slot2rep[sl] != sl) d.line_fold[r] |= 1u << j;
        }
    // map shortcut
    d.map_shortcut_ok = d.rekeyed && getenv("BK_NO_MAP_SHORTCUT") == nullptr;
    if (d.map_shortcut_ok) {
        d.id_bucket.assign(n_ids * k, OffLen{0, 0});
        const u64 bmask2 = (1ull << d.bucket_log2) - 1;
        parallel_for((n_ids + 4095) / 4096, [&](size_t c) {
            for (size_t i = c * 4096, i_end = std::min(n_ids, (c + 1) * 4096); i < i_end; i++) {
                const u64 S = d.id_kmer[i], R = revcomp_host(S, (int)k);
                const bool rcS = !(S < R);                                          // src/lcb.rs:87-95
                const u64 kb = rcS ? R : S;
                d.id_amb[i] = (d.id_amb[i] & 0x7FFFFFFFu) | (rcS ? 0x80000000u : 0u);
                for (u32 j = 0; j < k; j++) {
                    const u32 jc = rcS ? k - 1 - j : j;
                    const u64 key = ((u64)jc << 58) | (kb & ~(3ull << (2 * (k - 1 - jc))));
                    u64 h = hash_slot_host(key, 64 - d.bucket_log2);
                    for (;;) {
                        const BucketSlot& sl = d.bucket_slots[h];
                        if (sl.key == key) { d.id_bucket[i * k + j] = OffLen{sl.off, sl.len}; break; }
                        if (sl.key == ~0ull) break;
                        h = (h + 1) & bmask2;
                    }
                }
            }
        });
    }
}

}  // namespace bk
