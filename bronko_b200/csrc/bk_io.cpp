// bk_io.cpp — host text I/O around the path: FASTQ(.gz) decode (KMC reader contract, SURVEY.md
// Appendix B), sample-id derivation (reference src/util.rs:30-50), Rust-compatible float formatting,
// and the Thompson-tau table (reference src/call.rs:922-929; statrs 0.18 StudentsT::inverse_cdf).
#include "bk_host.h"

#include <zlib.h>

#include <cmath>
#include <cstdio>
#include <cstring>
#include <limits>

namespace bk {

// Streaming decode: the file is inflated block by block (gzread also passes plain text through) and only the
// sequence lines (line index 1 of every 4-line record) are kept; a line may span blocks.  '\r' before the line end is
// dropped; a last line without '\n' counts.
bool fastq_read(const std::string& path, std::vector<HostReads>& chunks, u64 max_chunk_bases, std::string& err) {
    gzFile g = gzopen(path.c_str(), "rb");
    if (!g) { err = "Failed to read reads file: " + path; return false; }
    gzbuffer(g, 1 << 20);
    chunks.clear();
    chunks.emplace_back();
    chunks.back().off.push_back(0);
    std::vector<char> buf(8u << 20);
    std::string carry;                        // the part of a sequence line that ended the previous block
    u64 line = 0;
    auto end_read = [&](const char* p, size_t len) {             // a complete sequence line (without its '\n')
        const char* q = p; size_t n = len;
        if (!carry.empty()) { carry.append(p, len); q = carry.data(); n = carry.size(); }
        if (n > 0 && q[n - 1] == '\r') n--;
        if (chunks.back().bases.size() + n > max_chunk_bases && chunks.back().off.size() > 1) {
            chunks.emplace_back();
            chunks.back().off.push_back(0);
        }
        HostReads& c = chunks.back();
        c.bases.insert(c.bases.end(), (const u8*)q, (const u8*)q + n);
        c.off.push_back((u32)c.bases.size());
        if (n > c.max_len) c.max_len = (u32)n;
        carry.clear();
    };
    int got;
    bool open_line = false;                   // bytes of an unterminated line have been seen
    while ((got = gzread(g, buf.data(), (unsigned)buf.size())) > 0) {
        const char* p = buf.data();
        const char* const e = p + got;
        while (p < e) {
            const char* nl = (const char*)memchr(p, '\n', (size_t)(e - p));
            if (!nl) {                        // the line continues in the next block
                if ((line & 3) == 1) carry.append(p, (size_t)(e - p));
                open_line = true;
                break;
            }
            if ((line & 3) == 1) end_read(p, (size_t)(nl - p));
            line++;
            open_line = false;
            p = nl + 1;
        }
    }
    const bool ok = got == 0;
    gzclose(g);
    if (!ok) { err = "Failed to read reads file: " + path; return false; }
    if (open_line && (line & 3) == 1) end_read("", 0);           // last line without a newline
    return true;
}

static bool ends_with(const std::string& s, const char* suf) {
    const size_t m = strlen(suf);
    return s.size() >= m && memcmp(s.data() + s.size() - m, suf, m) == 0;
}

// util.rs:30-50: first matching suffix in list order, stripped repeatedly (trim_end_matches);
// otherwise drop the final extension.
std::string clean_sample_id(const std::string& path) {
    size_t sl = path.find_last_of('/');
    std::string name = sl == std::string::npos ? path : path.substr(sl + 1);
    static const char* const suffixes[] = {".fastq.gz", ".fasta.gz", "fna.gz", "fnq.gz", ".fq.gz", ".fastq",
                                           ".fasta", ".fnq", ".fna", ".fa", ".fq"};
    for (const char* suf : suffixes) {
        if (!ends_with(name, suf)) continue;
        const size_t m = strlen(suf);
        while (ends_with(name, suf)) name.resize(name.size() - m);
        return name;
    }
    size_t dot = name.find_last_of('.');
    return (dot == std::string::npos || dot == 0) ? name : name.substr(0, dot);
}

// Rust `{:.N}` on f64: same digits as printf for finite values; NaN / inf spelled Rust's way.
std::string fmt_fixed(double v, int prec) {
    if (std::isnan(v)) return "NaN";
    if (std::isinf(v)) return v > 0 ? "inf" : "-inf";
    char b[64];
    snprintf(b, sizeof b, "%.*f", prec, v);
    return b;
}

// ---------------------------------------------------------------------------------------------
// Student-t quantile the way statrs 0.18 computes it: x1 = 1 - x (for x >= 0.5),
// y = inv_beta_reg(df/2, 1/2, 2*x1), t = sqrt(df*(1-y)/y).  inv_beta_reg is AS 109 (Cran, Martin &
// Thomas 1977) on top of a Lentz continued fraction for the regularised incomplete beta and a
// Lanczos (g = 10.900511, 11 terms) ln-gamma.  Evaluated once per context: tau depends only on
// curr_n in [3, 300] (window 100 x 3 minor alleles).
// ---------------------------------------------------------------------------------------------
namespace {
const double kLanczosR = 10.900511;
const double kLanczos[11] = {2.48574089138753565546e-5, 1.05142378581721974210, -3.45687097222016235469,
                             4.51227709466894823700, -2.98285225323576655721, 1.05639711577126713077,
                             -1.95428773191645869583e-1, 1.70970543404441224307e-2, -5.71926117404305781283e-4,
                             4.63399473359905636708e-6, -2.71994908488607703910e-9};
const double kLn2SqrtEOverPi = 0.6207822376352452223455184457816472122518527279025978;
const double kLnPi = 1.1447298858494001741434273513530587116472948129153;

double lgam(double x) {
    if (x < 0.5) {
        double acc = kLanczos[0];
        for (int i = 1; i < 11; i++) acc += kLanczos[i] / ((double)i - x);
        return kLnPi - std::log(std::sin(M_PI * x)) - std::log(acc) - kLn2SqrtEOverPi - (0.5 - x) * std::log((0.5 - x + kLanczosR) / M_E);
    }
    double acc = kLanczos[0];
    for (int i = 1; i < 11; i++) acc += kLanczos[i] / (x + (double)i - 1.0);
    return std::log(acc) + kLn2SqrtEOverPi + (x - 0.5) * std::log((x - 0.5 + kLanczosR) / M_E);
}

double ibeta(double a, double b, double x) {
    const double front = (x == 0.0 || x == 1.0) ? 0.0
        : std::exp(lgam(a + b) - lgam(a) - lgam(b) + a * std::log(x) + b * std::log(1.0 - x));
    const bool mirror = x >= (a + 1.0) / (a + b + 2.0);
    const double eps = 1.1102230246251565e-16;
    const double tiny = std::numeric_limits<double>::min() / eps;
    if (mirror) { double t = a; a = b; b = t; x = 1.0 - x; }
    const double qab = a + b, qap = a + 1.0, qam = a - 1.0;
    double c = 1.0, d = 1.0 - qab * x / qap;
    if (std::fabs(d) < tiny) d = tiny;
    d = 1.0 / d;
    double h = d;
    for (int it = 1; it < 141; it++) {
        const double m = it, m2 = m * 2.0;
        double aa = m * (b - m) * x / ((qam + m2) * (a + m2));
        d = 1.0 + aa * d; if (std::fabs(d) < tiny) d = tiny;
        c = 1.0 + aa / c; if (std::fabs(c) < tiny) c = tiny;
        d = 1.0 / d;
        h = h * d * c;
        aa = -(a + m) * (qab + m) * x / ((a + m2) * (qap + m2));
        d = 1.0 + aa * d; if (std::fabs(d) < tiny) d = tiny;
        c = 1.0 + aa / c; if (std::fabs(c) < tiny) c = tiny;
        d = 1.0 / d;
        const double del = d * c;
        h *= del;
        if (std::fabs(del - 1.0) <= eps) break;
    }
    return mirror ? 1.0 - front * h / a : front * h / a;
}

double inv_ibeta(double a, double b, double x) {
    if (x == 0.0) return 0.0;
    if (x == 1.0) return 1.0;
    const double lbeta = lgam(a) + lgam(b) - lgam(a + b);
    bool mirrored = false;
    if (x > 0.5) { double t = a; a = b; b = t; x = 1.0 - x; mirrored = true; }
    double r = std::sqrt(-std::log(x * x));
    double y = r - (2.30753 + 0.27061 * r) / (1.0 + (0.99229 + 0.04481 * r) * r);
    double p;
    if (a > 1.0 && b > 1.0) {
        r = (y * y - 3.0) / 6.0;
        const double s = 1.0 / (a + a - 1.0), t = 1.0 / (b + b - 1.0);
        const double h = 2.0 / (s + t);
        const double w = y * std::sqrt(h + r) / h - (t - s) * (r + 5.0 / 6.0 - 2.0 / (3.0 * h));
        p = a / (a + b * std::exp(w + w));
    } else {
        r = b + b;
        double t = 1.0 / (9.0 * b);
        t = r * std::pow(1.0 - t + y * std::sqrt(t), 3.0);
        if (t <= 0.0) p = 1.0 - std::exp((std::log((1.0 - x) * b) + lbeta) / b);
        else {
            t = (4.0 * a + r - 2.0) / t;
            p = t <= 1.0 ? std::exp((std::log(x * a) + lbeta) / a) : 1.0 - 2.0 / (t + 1.0);
        }
    }
    r = 1.0 - a;
    const double t1 = 1.0 - b;
    double yprev = 0.0, sq = 1.0, prev = 1.0;
    if (p < 0.0001) p = 0.0001;
    if (p > 0.9999) p = 0.9999;
    const double acu = std::pow(10.0, std::fmax(-5.0 / a / a - 1.0 / std::pow(x, 0.2) - 13.0, -30.0));
    double tx = p;
    for (int guard = 0; guard < 1000; guard++) {
        y = ibeta(a, b, p);
        y = (y - x) * std::exp(lbeta + r * std::log(p) + t1 * std::log(1.0 - p));
        if (y * yprev <= 0.0) prev = std::fmax(sq, 1e-30);
        double g = 1.0;
        bool converged = false;
        for (;;) {
            for (;;) {
                const double adj = g * y;
                sq = adj * adj;
                if (sq < prev) {
                    tx = p - adj;
                    if (tx >= 0.0 && tx <= 1.0) break;
                }
                g /= 3.0;
            }
            if (prev <= acu || y * y <= acu) { p = tx; converged = true; break; }
            if (tx != 0.0 && tx != 1.0) break;
            g /= 3.0;
        }
        if (converged || tx == p) break;
        p = tx;
        yprev = y;
    }
    return mirrored ? 1.0 - p : p;
}
}  // namespace

void tau_table(double* tab301) {
    for (int i = 0; i < 301; i++) tab301[i] = 0.0;
    const double alpha = 0.001;
    for (int n = 3; n <= 300; n++) {
        const double dn = (double)n, df = (double)(n - 2);
        const double q = 1.0 - alpha / dn;
        const double x1 = q >= 0.5 ? 1.0 - q : q;
        double y = inv_ibeta(0.5 * df, 0.5, 2.0 * x1);
        y = std::sqrt(df * (1.0 - y) / y);
        const double t_crit = q >= 0.5 ? y : -y;
        tab301[n] = (t_crit * (dn - 1.0)) / (std::sqrt(dn) * std::sqrt(dn - 2.0 + t_crit * t_crit));
    }
}

}  // namespace bk

// ---- the host decode stage of the C ABI (include/bronko_b200.h): no context, no GPU, thread-safe ----------------
struct bk_reads { std::vector<bk::HostReads> chunks; };

extern "C" {

int bk_fastq_decode(const char* path, bk_reads** out, char* err, uint64_t err_cap) {
    if (out) *out = nullptr;
    if (!path || !out) return BK_ERR_ARG;
    bk_reads* r = new bk_reads();
    std::string e;
    if (!bk::fastq_read(path, r->chunks, 1ull << 30, e)) {
        if (err && err_cap) { snprintf(err, (size_t)err_cap, "%s", e.c_str()); }
        delete r;
        return BK_ERR_IO;
    }
    for (bk::HostReads& c : r->chunks) c.bases.resize(c.bases.size() + 64, '*');      // readable slack past the last read
    *out = r;
    return BK_OK;
}

uint64_t bk_reads_n_chunks(const bk_reads* r) { return r ? r->chunks.size() : 0; }

int bk_reads_chunk(const bk_reads* r, uint64_t i, const uint8_t** bases, const uint32_t** read_off, uint64_t* n_reads, uint64_t* n_bases) {
    if (!r || i >= r->chunks.size()) return BK_ERR_ARG;
    const bk::HostReads& c = r->chunks[i];
    if (bases) *bases = c.bases.data();
    if (read_off) *read_off = c.off.data();
    if (n_reads) *n_reads = c.off.size() - 1;
    if (n_bases) *n_bases = c.off.back();
    return BK_OK;
}

void bk_reads_free(bk_reads* r) { delete r; }

}  // extern "C"
