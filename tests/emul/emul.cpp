// tests/emul/emul.cpp — TESTS ONLY.  Steps bronko_b200/csrc/bk_core.cuh (the counting-stage logic the
// CUDA kernels are built from) on the CPU, single-threaded, so that seed/extend/run/leftover/fold
// logic can be checked against the oracle in the `-m "not gpu"` suite.  Not part of libbronko_b200.so.
#include <algorithm>
#include <cstring>
#include <string>
#include <vector>

#include "../../bronko_b200/csrc/bk_core.cuh"
#include "../../bronko_b200/csrc/bk_host.h"

using namespace bk;

struct Emul {
    HostIndex ix;
    DerivedIndex d;
    std::vector<u64> out_kmers;
    std::vector<u32> out_counts;
    std::string err;
};

extern "C" {

void* emul_create_bkdb(const char* path) {
    Emul* e = new Emul();
    if (!bkdb_read(path, e->ix, e->err)) { delete e; return nullptr; }
    derive_index(e->ix, e->d);
    return e;
}
void* emul_create_fasta(u32 k, u32 n, const char** paths) {
    Emul* e = new Emul();
    std::vector<std::string> p(paths, paths + n);
    if (!index_build_from_fasta(k, p, e->ix, e->err)) { delete e; return nullptr; }
    derive_index(e->ix, e->d);
    return e;
}
void emul_free(void* h) { delete (Emul*)h; }
u64 emul_n_keys(void* h) { return ((Emul*)h)->ix.keys.size(); }
u64 emul_n_entries(void* h) { return ((Emul*)h)->ix.entries.size(); }
void emul_export(void* h, u64* keys, u64* off, bk_bucket_info* ent) {
    Emul* e = (Emul*)h;
    memcpy(keys, e->ix.keys.data(), e->ix.keys.size() * 8);
    memcpy(off, e->ix.entry_off.data(), e->ix.entry_off.size() * 8);
    memcpy(ent, e->ix.entries.data(), e->ix.entries.size() * sizeof(bk_bucket_info));
}
int emul_save(void* h, const char* path) { Emul* e = (Emul*)h; return bkdb_write(path, e->ix, e->err) ? 0 : -1; }
void emul_assign_buckets(u64 kmer, int k, u64* out) { assign_buckets_host(kmer, k, out); }
u64 emul_revcomp(u64 v, int k) { return revcomp_host(v, k); }
void emul_tau_table(double* t301) { tau_table(t301); }
u64 emul_clean_sample_id(const char* path, char* buf, u64 cap) {
    std::string t = clean_sample_id(path);
    if (buf && cap) { u64 n = std::min<u64>(cap - 1, t.size()); memcpy(buf, t.data(), n); buf[n] = 0; }
    return t.size() + 1;
}

// Counting stage on the CPU with the device logic: scan every read, drain the leftover queue, prefix
// sum + fold, compaction.  stats4 = total_reads, total_kmers, unique_kmers, unique_counted.
// dbg3 = number of leftover descriptors, leftover k-mers, novel keys.
u64 emul_count(void* h, const u8* bases, const u32* off, u64 n_reads, u32 gen_log2, u32 desc_cap,
               u32 ci, u32 cs, u64* stats4, u64* dbg3) {
    Emul* e = (Emul*)h;
    const DerivedIndex& d = e->d;
    const u64 n_bases = off[n_reads];
    std::vector<u32> words((n_bases + 64) / 4 + 4, 0x2A2A2A2Au);   // padding bytes are '*' (invalid)
    memcpy(words.data(), bases, n_bases);
    std::vector<u32> diff(d.n_raw + 2, 0);
    std::vector<GenSlot> gen(1ull << gen_log2, GenSlot{BK_EMPTY, 0, 0});
    std::vector<uint2> desc(desc_cap);
    u32 n_desc = 0, gen_full = 0;
    CountView v;
    v.k = d.k;
    v.refnib = d.refnib.data(); v.ref_chunks = (u32)(d.refnib.size() / 4);
    v.oseq_start = d.oseq_start.data(); v.oseq_len = d.oseq_len.data();
    v.exact = (const ExactSlotD*)d.exact_slots.data(); v.exact_shift = 64 - d.exact_log2; v.exact_mask = (1u << d.exact_log2) - 1;
    v.diff = diff.data();
    v.gen = gen.data(); v.gen_shift = 64 - gen_log2; v.gen_mask = (1u << gen_log2) - 1;
    v.gen_full = &gen_full;
    v.desc = desc.data(); v.desc_cap = desc_cap; v.n_desc = &n_desc;
    auto ld = [&](u32 i) { return words[i]; };
    auto ldr4 = [&](u32 i4) { W4 r; r.x = d.refnib[4 * i4]; r.y = d.refnib[4 * i4 + 1]; r.z = d.refnib[4 * i4 + 2]; r.w = d.refnib[4 * i4 + 3]; return r; };
    u64 novel = 0, left_kmers = 0;
    for (u64 r = 0; r < n_reads; r++) {
        Pending pend; pend.n = 0; pend.d0 = make_uint2(0, 0); pend.d1 = make_uint2(0, 0);
        novel += scan_read(v, ld, ldr4, off[r], off[r + 1] - off[r], 0, pend);
        // same policy as flush_pending() on the device: a full queue means counting in place
        for (u32 q = 0; q < pend.n; q++) {
            const uint2 d = q == 0 ? pend.d0 : pend.d1;
            u32 overflow = 0;
            queue_push(v, d, &overflow);
            if (overflow) novel += count_stretch(v, ld, d.x, d.y, 0, 1);
        }
    }
    const u32 nd = std::min(n_desc, desc_cap);
    for (u32 i = 0; i < nd; i++) { left_kmers += desc[i].y; novel += count_stretch(v, ld, desc[i].x, desc[i].y, 0, 1); }
    if (gen_full) return ~0ull;
    // prefix sum + fold
    std::vector<u32> idcnt(d.id_kmer.size(), 0);
    u32 run = 0;
    for (u32 i = 0; i < d.n_raw; i++) {
        run += diff[i];
        if (run != 0) {
            if (d.slot2id[i] == 0xFFFFFFFFu) return ~0ull - 1;   // count on an invalid slot: logic error
            idcnt[d.slot2id[i]] += run;
        }
    }
    if (run != 0) return ~0ull - 2;
    const u64 cx = 1000000000ull;
    std::vector<std::pair<u64, u32>> kept;
    u64 total = 0, uniq = 0;
    for (size_t id = 0; id < idcnt.size(); id++) {
        const u32 c = idcnt[id];
        if (!c) continue;
        uniq++; total += c;
        if (c >= ci && c <= cx) kept.emplace_back(d.id_kmer[id], std::min(c, cs));
    }
    for (const GenSlot& s : gen) {
        if (s.key == BK_EMPTY) continue;
        uniq++; total += s.cnt;
        if (s.cnt >= ci && s.cnt <= cx) kept.emplace_back(s.key, std::min(s.cnt, cs));
    }
    std::sort(kept.begin(), kept.end());
    e->out_kmers.clear(); e->out_counts.clear();
    for (auto& kv : kept) { e->out_kmers.push_back(kv.first); e->out_counts.push_back(kv.second); }
    stats4[0] = n_reads; stats4[1] = total; stats4[2] = uniq; stats4[3] = kept.size();
    if (dbg3) { dbg3[0] = n_desc; dbg3[1] = left_kmers; dbg3[2] = novel; }
    return kept.size();
}
void emul_count_get(void* h, u64* kmers, u32* counts) {
    Emul* e = (Emul*)h;
    memcpy(kmers, e->out_kmers.data(), e->out_kmers.size() * 8);
    memcpy(counts, e->out_counts.data(), e->out_counts.size() * 4);
}

}  // extern "C"
