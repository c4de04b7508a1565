// bk_host.h — host-side structures of libbronko_b200 (index, I/O, derived device tables).
#pragma once
#include <cstdint>
#include <string>
#include <vector>

#include "../../include/bronko_b200.h"

namespace bk {

typedef uint64_t u64;
typedef uint32_t u32;
typedef uint16_t u16;
typedef uint8_t u8;

// ---- the decoded BronkoIndex (src/build.rs:23-60) as flat arrays -----------------------------
struct HostSeq { std::string name; u64 len; std::vector<u8> bases; };
struct HostGenome { std::string name; std::vector<HostSeq> seqs; };

struct HostIndex {
    u32 k = 0;
    u64 meta_k = 0;
    std::vector<u64> keys;                 // ascending
    std::vector<u64> entry_off;            // n_keys + 1
    std::vector<bk_bucket_info> entries;   // per-key order preserved (file order, then location)
    std::vector<HostGenome> genomes;
};

// (bucket id, entry) pair used while building / decoding before the sort by key
struct KeyedEntry { u64 key; bk_bucket_info e; };

// Sort pairs by key (stable) and fill ix.keys / entry_off / entries.
void index_from_pairs(HostIndex& ix, std::vector<KeyedEntry>& pairs);

bool bkdb_decode(const u8* data, u64 n, HostIndex& ix, std::string& err);
bool bkdb_read(const std::string& path, HostIndex& ix, std::string& err);
bool bkdb_write(const std::string& path, const HostIndex& ix, std::string& err);
bool index_build_from_fasta(u32 k, const std::vector<std::string>& paths, HostIndex& ix, std::string& err);

// lcb.rs primitives (host copies used by the builder)
void assign_buckets_host(u64 kmer, int k, u64* out);
u64 revcomp_host(u64 v, int k);
inline u8 nt_to_bits_host(u8 c) {
    switch (c) { case 'C': case 'c': return 1; case 'G': case 'g': return 2; case 'T': case 't': return 3; default: return 0; }
}

// ---- text I/O ------------------------------------------------------------------------------
struct HostReads { std::vector<u8> bases; std::vector<u32> off; u32 max_len = 0; };
bool slurp_maybe_gz(const std::string& path, std::string& out);
// FASTQ(.gz), 4-line records; splits into chunks of at most max_chunk_bases so offsets fit u32.
bool fastq_read(const std::string& path, std::vector<HostReads>& chunks, u64 max_chunk_bases, std::string& err);
std::string clean_sample_id(const std::string& path);
std::string first_token(const std::string& s);
std::string fmt_fixed(double v, int prec);

// ---- tables derived from the index for the device ------------------------------------------
struct BucketSlot { u64 key; u32 off; u32 len; };                 // 16 B; key == ~0 → empty
struct BucketEntry { u32 row; u16 file_id; u8 idx; u8 canonical; }; // 8 B; row = global pileup row of `location`
struct ExactSlot { u64 key; u32 gidx; u32 oseq; };                // 16 B; key == ~0 → empty
struct OffLen { u32 off; u32 len; };                              // 8 B; entries [off, off + len) of one bucket

struct DerivedIndex {
    u32 k = 0, n_genomes = 0, n_seqs = 0;
    // pileup row space: all sequences of all genomes concatenated
    std::vector<u32> genome_row0;      // n_genomes + 1
    std::vector<u32> genome_seq_off;   // n_genomes + 1 (into the flat sequence list)
    std::vector<u32> seq_row0;         // n_seqs + 1, global row of each sequence
    std::vector<u64> genome_len;       // n_genomes
    std::vector<u8> ref_code;          // per row: nt_to_bits(base)
    u32 max_genome_rows = 0;
    // bucket-id → entries.  rekeyed: slots are keyed by (bucket index << 58) | (canonical k-mer with that digit
    // zeroed) instead of the bucket id — the same map, because assign_buckets is a bijection (src/lcb.rs:1-45);
    // only for k <= 29 (ids wrap for k = 31, src/lcb.rs:8-41) and only if every key of the index verifies.
    bool rekeyed = false;
    u32 bucket_log2 = 0;
    std::vector<BucketSlot> bucket_slots;
    // Grouped form of the re-keyed table (only when rekeyed and n_genomes <= 4: the thread-per-k-mer map kernels).
    // A query k-mer Q hits bucket (i, Q without digit i) iff some reference k-mer equals Q everywhere except possibly
    // at digit i.  Such a reference k-mer shares Q's low half (digits mid..k-1) when i < mid and Q's high half
    // otherwise, so the canonical reference k-mers ("centers") are grouped by (side, shared half):
    //   group key = (side << 62) | shared half value;  group_slots: open addressing {key, first center, count};
    //   group_centers: {center k-mer C, first of its buckets in group_buckets, bit mask of the bucket indices present};
    //   group_buckets: per center and side, for every index i of that side (0..mid-1 / mid..k-1): {off, len} of the
    //                  entries of bucket (i, C without digit i) — the same (off, len) bucket_slots holds.
    // A query needs two group probes, its one or two centers per side, and then one {off, len} per hit: C == Q hits
    // every bucket of the side, C differing from Q in exactly digit j hits bucket j, anything else nothing.  Centers are
    // the verified first-entry k-mers of the keys: every existing bucket (i, M) has such a center C with C without
    // digit i == M, so a query that hits the bucket is within one digit (i) of C and finds it; several centers can
    // lead to the same bucket (they differ in digit i only), the kernel takes each index once.
    u32 group_mid = 0;                       // bucket indices < group_mid are on side 0 (shared half = digits mid..k-1)
    u32 group_log2 = 0;
    std::vector<BucketSlot> group_slots;     // key == ~0 → empty; off = first center, len = count
    std::vector<BucketSlot> group_centers;   // key = center, off = first bucket, len = present mask
    std::vector<OffLen> group_buckets;
    std::vector<BucketEntry> bucket_entries;
    u32 max_key_entries = 0;                 // longest entry list of one key (the thread-per-k-mer map kernels count hits per
                                             // genome in 16-bit fields: usable while max_key_entries * k < 65536)
    // oriented reference store (forward and reverse-complement of every sequence), 2-bit packed,
    // MSB-first, 32 bases per u64; global base index space with REF_PAD_BASES of padding in front
    std::vector<u64> refpk;
    std::vector<u32> refnib;                 // the same store, one 4-bit code per base index (padding = 4), 16-byte chunks
    std::vector<u32> oseq_start, oseq_len;   // per oriented sequence, in global base indices
    u32 n_raw = 0;                           // size of the global base index space (incl. padding)
    // exact (non-canonical) reference k-mer → one representative raw slot
    u32 exact_log2 = 0;
    std::vector<ExactSlot> exact_slots;
    // raw slot → id of the distinct k-mer string (0xFFFFFFFF for invalid slots); id → k-mer
    std::vector<u32> slot2id;
    std::vector<u64> id_kmer;
    // Mismatch lines (bk_core.cuh: emit_dense; bk_dense.cuh).  A k-mer with exactly one mismatch against the diagonal of
    // its read is counted in cell (raw slot, j, b).  Cells are folded onto the representative raw slot of the reference
    // k-mer (slot2rep), and a cell's k-mer string "id's k-mer with digit j replaced by b" is unique to the cell unless
    // another reference k-mer lies within Hamming distance 2 of id's and differs from it at digit j (then two cells —
    // or a cell and a reference k-mer — can spell the same string): bit j of id_amb[id] marks those, they are
    // counted through the exact bins instead.  nb_slots: (j << 58 | k-mer without digit j) → id, for every id and j:
    // a k-mer that reaches the bins another way (two mismatches on its own diagonal, foreign diagonal, ...) finds the
    // cell it belongs to.  dense_ok = false (k > 29, large databases): everything goes through the bins as before.
    bool dense_ok = false;
    std::vector<u32> slot2rep;               // raw slot → representative raw slot of its k-mer (0xFFFFFFFF: invalid slot)
    std::vector<u32> id_rep;                 // id → its representative raw slot
    std::vector<u32> line_amb, line_fold;    // per reference base r: bit j = cell (slot r - j, j) is ambiguous / sits on a non-representative slot
    std::vector<u32> id_amb;                 // per id: bit j set = cells (id, j, *) are ambiguous
    u32 nb_log2 = 0;
    std::vector<ExactSlot> nb_slots;         // key, gidx = id (oseq unused); key == ~0 → empty
    u32 nb_bloom_log2 = 0;
    std::vector<u64> nb_bloom;               // blocked bit set over the keys of nb_slots: word = top nb_bloom_log2 bits of the hash, three bits of it
                                             // from the next 18 (bloom_mask) — >= 32 bits per key, one miss in ~5,000 look-ups, 16 MB that stay in L2
    // Map shortcut (rekeyed tables only).  The k-mer of an unambiguous cell (id, j, b) is within one digit of exactly one
    // reference k-mer — id's — so map_kmers can hit one bucket at most: index j of id's canonical form, and only if
    // replacing the digit does not flip which strand of the k-mer is canonical (bit 31 of id_amb[id]: id's k-mer is
    // the reverse complement of its canonical form).  id_bucket[id * k + j] = {off, len} of that bucket's entries.
    bool map_shortcut_ok = false;
    std::vector<OffLen> id_bucket;
};

static const u32 REF_PAD_BASES = 64;

void derive_index(const HostIndex& ix, DerivedIndex& d, bool allow_rekey = true);

// must match bk::hash_slot in bk_core.cuh (the device probes with the same function)
inline u32 hash_slot_host(u64 x, u32 shift) { return (u32)(((x ^ (x >> 31)) * 0x9E3779B97F4A7C15ull) >> shift); }
// the neighbour bit set (nb_bloom; bk_bins.cuh: bloom_mask_dev is the same function): the word is the top `log2w` bits
// of hash64, its three bits are the next 3 x 6 bits
inline u64 hash64_host(u64 x) { return (x ^ (x >> 31)) * 0x9E3779B97F4A7C15ull; }
inline u64 bloom_mask(u64 h, u32 log2w) {
    const u64 r = h >> (64 - log2w - 18);
    return (1ull << (r & 63)) | (1ull << ((r >> 6) & 63)) | (1ull << ((r >> 12) & 63));
}

// ---- Student-t / Thompson tau table (call.rs:922-929) --------------------------------------
void tau_table(double* tab301);   // tab[n] for n in 3..300 (others 0)

}  // namespace bk
