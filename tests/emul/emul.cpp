// tests/emul/emul.cpp — TESTS ONLY.  Steps bronko_b200/csrc/bk_core.cuh (the counting-stage logic the
// CUDA kernels are built from) on the CPU, single-threaded, so that seed/extend/run/leftover/fold
// logic can be checked against the oracle in the `-m "not gpu"` suite.  Not part of libbronko_b200.so.
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "../../bronko_b200/csrc/bk_core.cuh"
#include "../../bronko_b200/csrc/bk_noise.cuh"
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

// The lookup of k_map_grp (bk_kernels.cuh) stepped on the host tables: for every canonical query k-mer the set of
// (bucket index, off, len) found through group_slots / group_centers / group_buckets must equal what k direct probes
// of bucket_slots find.  Returns the number of queries that disagree (0 = the grouped form is an exact re-indexing);
// ~0 if the index has no grouped form.
u64 emul_group_check(void* h, const u64* queries, u64 n, u32 b0, u32 b1) {
    const DerivedIndex& d = ((Emul*)h)->d;
    if (!d.rekeyed || d.group_slots.empty()) return ~0ull;
    const u32 k = d.k, mid = d.group_mid, lo_bits = 2 * (k - mid);
    const u64 bmask = (1ull << d.bucket_log2) - 1, gmask = (1ull << d.group_log2) - 1;
    u64 bad = 0;
    for (u64 q = 0; q < n; q++) {
        const u64 kb = queries[q];
        std::vector<u64> direct, grouped;                       // (index << 40) ^ off ^ (len << 32) is enough to compare sets
        for (u32 i = b0; i < b1; i++) {
            const u64 key = ((u64)i << 58) | (kb & ~(3ull << (2 * (k - 1 - i))));
            u64 hh = hash_slot_host(key, 64 - d.bucket_log2);
            for (;;) {
                const BucketSlot& sl = d.bucket_slots[hh];
                if (sl.key == key) { direct.push_back(((u64)i << 56) | ((u64)sl.len << 32) | sl.off); break; }
                if (sl.key == ~0ull) break;
                hh = (hh + 1) & bmask;
            }
        }
        for (u32 sd = 0; sd < 2; sd++) {
            const u32 i_lo = sd == 0 ? b0 : std::max(b0, mid), i_hi = sd == 0 ? std::min(b1, mid) : b1;
            if (i_lo >= i_hi) continue;
            const u64 gk = sd == 0 ? (kb & ((1ull << lo_bits) - 1)) : ((1ull << 62) | (kb >> lo_bits));
            u64 hh = hash_slot_host(gk, 64 - d.group_log2);
            u32 first = 0, count = 0;
            for (;;) {
                const BucketSlot& sl = d.group_slots[hh];
                if (sl.key == gk) { first = sl.off; count = sl.len; break; }
                if (sl.key == ~0ull) break;
                hh = (hh + 1) & gmask;
            }
            const u32 side_lo = sd == 0 ? 0u : mid;
            const u32 range = ((1u << i_hi) - 1u) & ~((1u << i_lo) - 1u);
            u32 done = 0;
            for (u32 c = 0; c < count; c++) {
                const BucketSlot& cen = d.group_centers[first + c];
                const u64 x = kb ^ cen.key;
                u32 cand = cen.len & range;
                if (x) {
                    const u64 nz = (x | (x >> 1)) & 0x5555555555555555ull;
                    cand = (nz & (nz - 1)) ? 0u : cand & (1u << (k - 1 - ((63u - (u32)BK_CLZLL(nz)) >> 1)));
                }
                cand &= ~done;
                done |= cand;
                for (u32 i = 0; i < 32; i++)
                    if ((cand >> i) & 1) {
                        const OffLen& ol = d.group_buckets[cen.off + i - side_lo];
                        grouped.push_back(((u64)i << 56) | ((u64)ol.len << 32) | ol.off);
                    }
            }
        }
        std::sort(direct.begin(), direct.end());
        std::sort(grouped.begin(), grouped.end());
        if (direct != grouped) bad++;
    }
    return bad;
}
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
u64 emul_count(void* h, const u8* bases, const u32* off, u64 n_reads, u32 gen_log2_flags, u32 desc_cap,
               u32 ci, u32 cs, u64* stats4, u64* dbg3) {
    Emul* e = (Emul*)h;
    const bool use_dense = (gen_log2_flags >> 31) != 0;
    const u32 gen_log2 = gen_log2_flags & 0xFFu;
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
    v.nov = nullptr; v.nov_cap = 0; v.nov_n = nullptr;
    v.desc = desc.data(); v.desc_cap = desc_cap; v.n_desc = &n_desc;
    // mismatch lines (bk_core.cuh: emit_dense), switched on by bit 31 of gen_log2's caller-side flag word
    std::vector<u32> dense; std::vector<u8> dflag;
    v.dense = nullptr; v.dense_flag = nullptr;
    if (use_dense) { dense.assign((size_t)d.n_raw * 4 * (d.k + 1), 0); dflag.assign((size_t)d.n_raw * 4, 0); v.dense = dense.data(); v.dense_flag = dflag.data(); }
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
    // mismatch lines: prefix sum along j; cell (line, j) = occurrences of the reference k-mer at raw slot r - j with digit j
    // replaced by b.  Here every such k-mer simply joins the other counts (exact reference k-mer → its slot, else the novel
    // table): the device's shortcuts for unambiguous cells (bk_dense.cuh) must give the same totals.
    u64 dense_cells = 0, dense_kmers = 0;
    if (use_dense) {
        const u32 k = d.k;
        for (size_t line = 0; line < dflag.size(); line++) {
            if (!dflag[line]) continue;
            const u32 refpos = (u32)(line >> 2), alt = (u32)(line & 3);
            u32 run = 0;
            for (u32 j = 0; j <= k; j++) {
                run += dense[line * (k + 1) + j];
                if (j == k) { if (run != 0) return ~0ull - 3; break; }             // a line must sum to zero
                if (!run) continue;
                if (refpos < j) return ~0ull - 4;
                const u32 slot = refpos - j;
                if (slot >= d.n_raw || d.slot2id[slot] == 0xFFFFFFFFu) return ~0ull - 5;      // cell on an invalid slot: logic error
                const u64 ref = d.id_kmer[d.slot2id[slot]];
                const u32 sh = 2 * (k - 1 - j);
                if (((ref >> sh) & 3) == alt) return ~0ull - 6;                     // not a mismatch: logic error
                const u64 km = (ref & ~(3ull << sh)) | ((u64)alt << sh);
                dense_cells++; dense_kmers += run;
                u32 gidx, oseq;
                if (exact_lookup(v, km, &gidx, &oseq)) { diff[gidx] += run; diff[gidx + 1] -= run; }
                else {
                    u32 hh = hash_slot(km, v.gen_shift);
                    for (;;) {
                        if (gen[hh].key == BK_EMPTY) { gen[hh].key = km; gen[hh].cnt = run; novel++; break; }
                        if (gen[hh].key == km) { gen[hh].cnt += run; break; }
                        hh = (hh + 1) & v.gen_mask;
                    }
                }
            }
        }
    }
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
    if (dbg3) { dbg3[0] = n_desc; dbg3[1] = left_kmers; dbg3[2] = use_dense ? dense_kmers : novel; }
    return kept.size();
}
void emul_count_get(void* h, u64* kmers, u32* counts) {
    Emul* e = (Emul*)h;
    memcpy(kmers, e->out_kmers.data(), e->out_kmers.size() * 8);
    memcpy(counts, e->out_counts.data(), e->out_counts.size() * 4);
}

// ---------------------------------------------------------------------------------------------------------
// Noise baseline with the device logic (bk_noise.cuh), orchestrated like the kernels: fractions, the two exact
// chains in block-wide rounds (threads stepped one after the other, the parity maps composed in thread order like
// the scan does), speculative table chunks + boundary verification + replay, Thompson tau.
//   fwd/rev: len*4 u32 depth arrays.  stats5: chunks replayed, iterations replayed, chain rounds, stops, serial its.
// ---------------------------------------------------------------------------------------------------------
}  // extern "C"
template <bool SQUARE, class LdM>
static void emul_chain(const LdM& M, u32 iters, double* snap, u32* stats) {
    // nz_chain_block of bk_noise.cuh, stepped in thread order (the block-wide scans of the device are replaced by
    // sequential sums / compositions of the same primitives: both are associative).  snap holds the approximate window
    // sums on entry (k_noise_fracs); the sums of the iterations that are not active are filled in at the end (the device
    // reads them through act_rank / act_list in k_noise_tau).
    std::vector<u32> list;
    for (u32 i = 0; i < iters; i++) {
        bool any = false;
        for (u32 j = 0; j < 3; j++) any = any || M((i32)i, j) != 0.0 || M((i32)i - BK_NOISE_WINDOW, j) != 0.0;
        if (any) list.push_back(i);
    }
    const u32 n_act = (u32)list.size();
    auto X = [&](u32 a, u32 q) { return nz_operand<SQUARE>(M, (i32)list[a], q); };
    double s = 0.0;
    u32 a0 = 0, force_serial = 0;
    while (a0 < n_act) {
        const u32 n_it = std::min<u32>(BK_NZ_ROUND, n_act - a0);
        const u64 sb = nz_b(s);
        const u32 ef = (u32)(sb >> 52);
        const bool capable = ef >= 66u && ef < 0x7FCu;
        const u32 start = nz_hint_start(capable ? ef : 1023u);
        u32 runkey = start, F = start, run = 0;
        for (u32 t = 0; t < n_it; t++) {
            runkey = nz_hint_merge(runkey, nz_hint_key(nz_b(snap[list[a0 + t]])));
            if (nz_hint_fits(runkey)) { F = nz_hint_merge(F, runkey); run++; }
        }
        if (!capable || force_serial || run < BK_NZ_MIN_RUN) {
            const u32 n_ser = std::min<u32>(n_it, force_serial ? (u32)BK_NZ_SERIAL : (capable ? std::max<u32>(run + 1u, BK_NZ_SERIAL_RUN) : (u32)BK_NZ_SERIAL_RUN));
            for (u32 t = 0; t < n_ser; t++) {
                for (u32 q = 0; q < 6; q++) s = nz_add(s, X(a0 + t, q));
                snap[list[a0 + t]] = s;
            }
            a0 += n_ser; stats[4] += n_ser; force_serial = 0;
            continue;
        }
        stats[2]++;
        NzZ2 z;
        {
            const u32 el = nz_hint_el(F);
            if (el + 2 < ef || el > ef) { stats[0] = 0xBAD; return; }        // the zones must hold the start
            nz2_zones_at(el, &z);
        }
        const i64 S0 = nz2_start(z, sb);
        const u32 n_thr = n_it;                              // one active iteration per thread
        std::vector<NzThread> T(n_thr);
        std::vector<i64> Pex(n_thr);
        i64 runA = 0;
        for (u32 t = 0; t < n_thr; t++) {                    // scan 1 + the thread maps
            i64 A[BK_NZ_OPT]; u32 fc[BK_NZ_OPT]; bool ok[BK_NZ_OPT]; i64 PA = 0;
            for (u32 q = 0; q < BK_NZ_OPT; q++) { ok[q] = nz2_split(z, X(a0 + t, q), &A[q], &fc[q]); PA += A[q]; }
            Pex[t] = runA;
            nz2_thread(A, fc, ok, S0 + runA, t * BK_NZ_OPT, &T[t]);
            runA += PA;
        }
        for (u32 g = 0; g < 3; g++) for (u32 fcc = 0; fcc < 4; fcc++) for (u32 v = 0; v < 8; v++)      // the table against its formula
            if ((i32)(signed char)(unsigned char)(nz2_round_table(g, fcc) >> (8 * v)) != nz2_round_slow(g, fcc, v)) { stats[0] = 0xBAD; return; }
        u32 cs = (u32)(u64)S0 & 7u;                          // scans 2 + 3 in thread order: class and roundings in front of every thread
        i64 Rex = 0;
        NzVec composed = nzvec_identity();                   // (and the composition of the transitions, as the device's warp scan builds it)
        u32 n_ok = n_it * 6;
        std::vector<i64> Tq((size_t)n_thr * BK_NZ_OPT);
        for (u32 t = 0; t < n_thr; t++) {
            if (nzvec_get(composed, (u32)(u64)S0 & 7u) != cs) { stats[0] = 0xBAD; return; }
            u32 bad = T[t].bad;
            for (u32 q = 0; q < BK_NZ_OPT; q++) {
                Tq[t * BK_NZ_OPT + q] = S0 + Pex[t] + T[t].a_pre[q] + Rex + (i64)nz2_pr(T[t].pr[q], cs);
                if (!nz2_inside(Tq[t * BK_NZ_OPT + q]) && bad > q) bad = q;
            }
            if (bad < BK_NZ_OPT) n_ok = std::min(n_ok, t * BK_NZ_OPT + bad);
            Rex += nz2_pr(T[t].pr[BK_NZ_OPT - 1], cs);
            cs = nzvec_get(T[t].next, cs);
            composed = nzvec_compose(composed, T[t].next);
        }
        for (u32 t = 0; t < n_it; t++) if (t * 6 + 5 < n_ok) snap[list[a0 + t]] = nz2_value(z, Tq[t * 6 + 5]);
        if (n_ok > 0) s = nz2_value(z, Tq[n_ok - 1]);
        if (n_ok == n_it * 6) { a0 += n_it; continue; }
        stats[3]++;
        const u32 ib = n_ok / 6, qb = n_ok - ib * 6;
        if (getenv("EMUL_DEBUG")) {
            const double x = X(a0 + ib, qb);
            i64 A; u32 fc; const bool ok = nz2_split(z, x, &A, &fc);
            fprintf(stderr, "[stop] sq=%d i=%u op=%u accepted=%u s=%.6g x=%.6g ok=%d el=%u T=%lld (2^53=%lld) threadbad=%u\n", (int)SQUARE, list[a0 + ib], qb, n_ok, s, x, (int)ok, z.el,
                    (long long)Tq[n_ok], (long long)(1ll << 53), T[ib].bad);
        }
        for (u32 q = qb; q < 6; q++) s = nz_add(s, X(a0 + ib, q));
        snap[list[a0 + ib]] = s;
        a0 += ib + 1;
        force_serial = ib < BK_NZ_SERIAL ? 1 : 0;
    }
    double last = 0.0;
    u32 a = 0;
    for (u32 i = 0; i < iters; i++) {
        if (a < n_act && list[a] == i) { last = snap[i]; a++; }
        else snap[i] = last;
    }
}

extern "C" {
// the look-ahead of nz_chain_block on its own: s = the sum a pass starts from, ahead[n] = approximate window sums of the
// iterations ahead → out[0] = iterations that fit three zones, out[1] = exponent field of the lowest zone (0: serial)
void emul_hint(double s, const double* ahead, u32 n, u32* out) {
    const u64 sb = nz_b(s);
    const u32 ef = (u32)(sb >> 52);
    const bool capable = ef >= 66u && ef < 0x7FCu;
    const u32 start = nz_hint_start(capable ? ef : 1023u);
    u32 runkey = start, F = start, run = 0;
    for (u32 t = 0; t < n; t++) {
        runkey = nz_hint_merge(runkey, nz_hint_key(nz_b(ahead[t])));
        if (nz_hint_fits(runkey)) { F = nz_hint_merge(F, runkey); run++; }
    }
    out[0] = run;
    out[1] = capable ? nz_hint_el(F) : 0u;
}

int emul_noise(const u32* fwd, const u32* rev, u32 len, double* out_max, u32* stats5) {
    for (int i = 0; i < 5; i++) stats5[i] = 0;
    if (len < BK_NOISE_WINDOW) { for (u32 i = 0; i < len; i++) out_max[i] = 0.0; return 1; }
    const u32 iters = len + BK_NOISE_HALF;
    std::vector<double> maf((size_t)(len + BK_NZ_PAD) * 3, 0.0);
    for (u32 p = 0; p < len; p++) nz_fractions(fwd + (size_t)p * 4, rev + (size_t)p * 4, &maf[(size_t)(p + BK_NZ_PAD_LO) * 3]);
    const double* m0 = maf.data() + (size_t)BK_NZ_PAD_LO * 3;
    auto M = [m0](i32 p, u32 j) { return m0[(long long)p * 3 + j]; };
    std::vector<double> snap_s(iters), snap_s2(iters), snap_tab((size_t)iters * BK_NOISE_TABLE);
    for (u32 i = 0; i < iters; i++) {                          // k_noise_fracs: approximate window sums where the exact ones will go
        double a = 0.0, b = 0.0;
        for (i32 p2 = (i32)i - (BK_NOISE_WINDOW - 1); p2 <= (i32)i; p2++)
            for (u32 j = 0; j < 3; j++) { const double v = M(p2, j); a += v; b += v * v; }
        snap_s[i] = a; snap_s2[i] = b;
    }
    emul_chain<false>(M, iters, snap_s.data(), stats5);
    emul_chain<true>(M, iters, snap_s2.data(), stats5);
    const u32 n_chunks = (iters + BK_NZ_CHUNK - 1) / BK_NZ_CHUNK;
    std::vector<double> warm((size_t)n_chunks * BK_NOISE_TABLE);
    for (u32 c = n_chunks; c-- > 0;)          // any order: chunks are independent
        nz_table_chunk(M, iters, c, snap_tab.data(), &warm[(size_t)c * BK_NOISE_TABLE]);
    std::vector<u8> flag(n_chunks, 0);
    for (u32 c = 1; c < n_chunks; c++)
        for (int q = 0; q < BK_NOISE_TABLE; q++)
            if (nz_b(snap_tab[(size_t)(c * BK_NZ_CHUNK - 1) * BK_NOISE_TABLE + q]) != nz_b(warm[(size_t)c * BK_NOISE_TABLE + q])) flag[c] = 1;
    for (u32 c = 1; c < n_chunks;) {
        if (!flag[c]) { c++; continue; }
        const u32 i0 = c * BK_NZ_CHUNK;
        const u32 met = nz_table_replay(M, iters, i0, snap_tab.data());
        stats5[0]++; stats5[1] += std::min(met, iters - 1) - i0 + 1;
        c = met / BK_NZ_CHUNK + 1;
    }
    double tau[301];
    tau_table(tau);
    auto tau_of = [&tau](u32 n) { return tau[n]; };
    for (u32 i = BK_NOISE_HALF; i < iters; i++) {
        u32 cn0 = 0;
        for (i32 p = (i32)i - (BK_NOISE_WINDOW - 1); p <= (i32)i; p++)
            if (p < (i32)len) cn0 += (M(p, 0) > 0.0) + (M(p, 1) > 0.0) + (M(p, 2) > 0.0);
        out_max[i - BK_NOISE_HALF] = nz_tau_loop(cn0, snap_s[i], snap_s2[i], &snap_tab[(size_t)i * BK_NOISE_TABLE], tau_of);
    }
    return 0;
}

}  // extern "C"
