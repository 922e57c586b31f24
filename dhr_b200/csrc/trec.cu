// trec.cu -- host-side TREC run writer: replaces the Python string formatting of
// castorini/dhr retrieval/gip_retrieval.py:329-342 (7 M lines at MS MARCO dev scale).
//
//   "{qid} Q0 {docid} {rank+1} {score} {run_name}\n"
// * rows whose docid equals the query id are skipped and ranks are NOT renumbered (:339-341);
// * {score} is Python's repr() of the float that `.tolist()` made from the fp32 score (:158-159), i.e. the
//   shortest round-trip decimal of the double with repr's fixed/exponent rule -- py_float_repr below.
// Queries are formatted by a pool of threads into per-thread buffers and written in query order.
#include <algorithm>
#include <charconv>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <string>
#include <string_view>
#include <thread>
#include <unordered_map>
#include <vector>

#include "../../include/dhr_b200.h"

namespace {

// CPython float_repr_style 'short', format code 'r' (Python/pystrtod.c format_float_short): shortest digits that
// round-trip; exponent form when decpt > 16 or decpt < -3; at least two exponent digits; ".0" added to integers.
size_t py_float_repr(double x, char* out) {
    if (std::isnan(x)) { memcpy(out, "nan", 3); return 3; }
    if (std::isinf(x)) { if (x < 0) { memcpy(out, "-inf", 4); return 4; } memcpy(out, "inf", 3); return 3; }
    char sci[64];
    auto r = std::to_chars(sci, sci + sizeof(sci), x, std::chars_format::scientific);   // [-]d[.ddd]e[+-]XX, shortest
    *r.ptr = 0;
    const char* p = sci;
    char* o = out;
    if (*p == '-') { *o++ = '-'; ++p; }
    char digits[32]; int nd = 0;
    while (*p && *p != 'e') { if (*p != '.') digits[nd++] = *p; ++p; }
    const int exp10 = atoi(p + 1);
    const int decpt = exp10 + 1;                       // value = 0.d1d2... x 10^decpt
    if (decpt > 16 || decpt < -3) {
        *o++ = digits[0];
        if (nd > 1) { *o++ = '.'; memcpy(o, digits + 1, nd - 1); o += nd - 1; }
        *o++ = 'e';
        int e = decpt - 1;
        *o++ = e < 0 ? '-' : '+';
        if (e < 0) e = -e;
        char eb[8]; int ne = 0;
        do { eb[ne++] = (char)('0' + e % 10); e /= 10; } while (e);
        if (ne < 2) eb[ne++] = '0';
        while (ne) *o++ = eb[--ne];
    } else if (decpt <= 0) {
        *o++ = '0'; *o++ = '.';
        for (int i = 0; i < -decpt; ++i) *o++ = '0';
        memcpy(o, digits, nd); o += nd;
    } else if (decpt >= nd) {
        memcpy(o, digits, nd); o += nd;
        for (int i = nd; i < decpt; ++i) *o++ = '0';
        *o++ = '.'; *o++ = '0';
    } else {
        memcpy(o, digits, decpt); o += decpt;
        *o++ = '.';
        memcpy(o, digits + decpt, nd - decpt); o += nd - decpt;
    }
    return (size_t)(o - out);
}

struct IdTable {                       // ids as int64 values or as strings packed in one buffer with [n + 1] offsets
    const int64_t* ints; const char* str; const int64_t* off;
    size_t put(int64_t i, char* out) const {
        if (ints) { auto r = std::to_chars(out, out + 24, ints[i]); return (size_t)(r.ptr - out); }
        const size_t n = (size_t)(off[i + 1] - off[i]);
        memcpy(out, str + off[i], n);
        return n;
    }
    size_t max_len(int64_t i) const { return ints ? 24 : (size_t)(off[i + 1] - off[i]); }
};

bool ids_equal(const IdTable& a, int64_t i, const IdTable& b, int64_t j) {
    if (a.ints && b.ints) return a.ints[i] == b.ints[j];
    char ba[32], bb[32];
    const char* pa; const char* pb; size_t na, nb;
    if (a.ints) { na = a.put(i, ba); pa = ba; } else { pa = a.str + a.off[i]; na = (size_t)(a.off[i + 1] - a.off[i]); }
    if (b.ints) { nb = b.put(j, bb); pb = bb; } else { pb = b.str + b.off[j]; nb = (size_t)(b.off[j + 1] - b.off[j]); }
    return na == nb && memcmp(pa, pb, na) == 0;
}

}  // namespace

extern "C" int dhr_write_trec(const char* path, int append, int n_queries, int k, const int32_t* counts, const int64_t* rows,
                              const float* scores, const int64_t* qid_int, const char* qid_str, const int64_t* qid_off,
                              int64_t n_docids, const int64_t* docid_int, const char* docid_str, const int64_t* docid_off,
                              int skip_equal, const char* run_name, int n_threads, int64_t* lines_written) {
    if (!path || n_queries < 0 || k < 0 || !rows || !scores || !run_name) return DHR_ERR_INVALID;
    if ((!qid_int && !(qid_str && qid_off)) || (!docid_int && !(docid_str && docid_off))) return DHR_ERR_INVALID;
    const IdTable qt{qid_int, qid_str, qid_off}, dt{docid_int, docid_str, docid_off};
    const size_t run_len = strlen(run_name);
    constexpr size_t kMaxQid = 512, kMaxDocid = 400, kMaxRun = 256;
    if (run_len > kMaxRun) return DHR_ERR_INVALID;                       // checked once, before any worker formats a line
    if (n_threads <= 0) n_threads = (int)std::thread::hardware_concurrency();
    if (n_threads <= 0) n_threads = 1;
    if (n_threads > 64) n_threads = 64;
    if (n_threads > n_queries) n_threads = n_queries > 0 ? n_queries : 1;
    std::vector<std::string> bufs((size_t)n_threads);
    std::vector<int64_t> lines((size_t)n_threads, 0);
    std::vector<int> bad((size_t)n_threads, 0);
    auto work = [&](int t) {
        const int q0 = (int)((int64_t)n_queries * t / n_threads), q1 = (int)((int64_t)n_queries * (t + 1) / n_threads);
        std::string& b = bufs[(size_t)t];
        b.reserve((size_t)(q1 - q0) * (size_t)k * 48);
        char line[kMaxQid + kMaxDocid + kMaxRun + 96];                   // ids + run name + " Q0 ", rank, score repr, separators
        for (int q = q0; q < q1; ++q) {
            const int n = counts ? counts[q] : k;
            char qbuf[kMaxQid];
            if (qt.max_len(q) > sizeof(qbuf)) { bad[(size_t)t] = 1; return; }
            const size_t ql = qt.put(q, qbuf);
            for (int r = 0; r < n && r < k; ++r) {
                const int64_t row = rows[(size_t)q * k + r];
                if (row < 0) continue;                                   // padding (k > rows in the shard)
                if (row >= n_docids || dt.max_len(row) > kMaxDocid) { bad[(size_t)t] = 1; return; }
                if (skip_equal && ids_equal(dt, row, qt, q)) continue;   // gip_retrieval.py:340
                char* o = line;
                memcpy(o, qbuf, ql); o += ql;
                memcpy(o, " Q0 ", 4); o += 4;
                o += dt.put(row, o);
                *o++ = ' ';
                o = std::to_chars(o, o + 12, r + 1).ptr;
                *o++ = ' ';
                o += py_float_repr((double)scores[(size_t)q * k + r], o);
                *o++ = ' ';
                memcpy(o, run_name, run_len); o += run_len;
                *o++ = '\n';
                b.append(line, (size_t)(o - line));
                ++lines[(size_t)t];
            }
        }
    };
    std::vector<std::thread> pool;
    for (int t = 1; t < n_threads; ++t) pool.emplace_back(work, t);
    work(0);
    for (auto& th : pool) th.join();
    for (int t = 0; t < n_threads; ++t) if (bad[(size_t)t]) return DHR_ERR_INVALID;
    FILE* f = fopen(path, append ? "ab" : "wb");
    if (!f) return DHR_ERR_INVALID;
    int64_t total = 0;
    bool ok = true;
    for (int t = 0; t < n_threads; ++t) {
        if (!bufs[(size_t)t].empty() && fwrite(bufs[(size_t)t].data(), 1, bufs[(size_t)t].size(), f) != bufs[(size_t)t].size()) ok = false;
        total += lines[(size_t)t];
    }
    if (fclose(f) != 0) ok = false;
    if (lines_written) *lines_written = total;
    return ok ? DHR_OK : DHR_ERR_INVALID;
}


// ---- shard merge on the host ------------------------------------------------------------------
// Replaces retrieval/merge.result.py:20-43: concatenate the shards' (docid, score) per query id in file order, keep the
// top-k by score, rewrite the ranks.  The reference sorts with `argsort()[::-1]`, which leaves the order of equal scores
// to numpy's unstable sort; here ties are ordered by position in the concatenated lists (shard, then the shard's own
// rank), i.e. the single-shard order.  Scores are re-printed as Python would print float(text).
namespace {

struct TrecItem { const char* docid; uint32_t docid_len; uint32_t pos; double score; };
struct TrecGroup { std::string_view qid; std::vector<TrecItem> items; };

bool read_file(const char* path, std::string& out) {
    FILE* f = fopen(path, "rb");
    if (!f) return false;
    fseek(f, 0, SEEK_END);
    const long n = ftell(f);
    fseek(f, 0, SEEK_SET);
    out.resize(n > 0 ? (size_t)n : 0);
    const bool ok = n <= 0 || fread(&out[0], 1, (size_t)n, f) == (size_t)n;
    fclose(f);
    return ok;
}

}  // namespace

extern "C" int dhr_merge_trec(int n_paths, const char* const* paths, const char* out_path, int topk, const char* run_name,
                              int n_threads, int64_t* lines_written) {
    if (n_paths < 0 || (n_paths > 0 && !paths) || !out_path || topk < 0 || !run_name) return DHR_ERR_INVALID;
    std::vector<std::string> files((size_t)n_paths);
    for (int i = 0; i < n_paths; ++i)
        if (!paths[i] || !read_file(paths[i], files[(size_t)i])) return DHR_ERR_INVALID;
    std::vector<TrecGroup> groups;                                   // in order of first appearance (dict order of the reference)
    std::unordered_map<std::string_view, size_t> index;
    for (const std::string& text : files) {
        const char* p = text.data();
        const char* end = p + text.size();
        while (p < end) {
            const char* eol = (const char*)memchr(p, '\n', (size_t)(end - p));
            if (!eol) eol = end;
            const char* le = eol;
            while (le > p && (le[-1] == '\r' || le[-1] == ' ' || le[-1] == '\t')) --le;     // line.strip()
            const char* lb = p;
            while (lb < le && (*lb == ' ' || *lb == '\t')) ++lb;
            if (lb < le) {
                const char* fld[7]; int nf = 0;                        // split(' '): exactly six fields
                fld[nf++] = lb;
                for (const char* c = lb; c < le && nf < 7; ++c) if (*c == ' ') fld[nf++] = c + 1;
                if (nf != 6) return DHR_ERR_INVALID;
                const std::string_view qid(fld[0], (size_t)(fld[1] - 1 - fld[0]));
                TrecItem it;
                it.docid = fld[2]; it.docid_len = (uint32_t)(fld[3] - 1 - fld[2]);
                const char* sb = fld[4]; const char* se = fld[5] - 1;
                if (sb < se && *sb == '+') ++sb;
                auto r = std::from_chars(sb, se, it.score);
                if (r.ec != std::errc() || r.ptr != se) {               // inf / nan spellings of Python
                    const std::string_view sv(sb, (size_t)(se - sb));
                    if (sv == "inf" || sv == "Infinity") it.score = INFINITY;
                    else if (sv == "-inf" || sv == "-Infinity") it.score = -INFINITY;
                    else if (sv == "nan") it.score = NAN;
                    else return DHR_ERR_INVALID;
                }
                auto f = index.find(qid);
                size_t gi;
                if (f == index.end()) { gi = groups.size(); groups.push_back(TrecGroup{qid, {}}); index.emplace(qid, gi); }
                else gi = f->second;
                it.pos = (uint32_t)groups[gi].items.size();
                groups[gi].items.push_back(it);
            }
            p = eol < end ? eol + 1 : end;
        }
    }
    const int nq = (int)groups.size();
    const size_t run_len = strlen(run_name);
    if (run_len > 256) return DHR_ERR_INVALID;
    if (n_threads <= 0) n_threads = (int)std::thread::hardware_concurrency();
    if (n_threads <= 0) n_threads = 1;
    if (n_threads > 64) n_threads = 64;
    if (n_threads > nq) n_threads = nq > 0 ? nq : 1;
    std::vector<std::string> bufs((size_t)n_threads);
    std::vector<int64_t> lines((size_t)n_threads, 0);
    std::vector<int> bad((size_t)n_threads, 0);
    auto work = [&](int t) {
        const int q0 = (int)((int64_t)nq * t / n_threads), q1 = (int)((int64_t)nq * (t + 1) / n_threads);
        std::string& b = bufs[(size_t)t];
        char line[1200];
        for (int q = q0; q < q1; ++q) {
            std::vector<TrecItem>& v = groups[(size_t)q].items;
            const size_t keep = std::min<size_t>((size_t)topk, v.size());
            auto before = [](const TrecItem& a, const TrecItem& c) { return a.score > c.score || (a.score == c.score && a.pos < c.pos); };
            std::partial_sort(v.begin(), v.begin() + (long)keep, v.end(), before);
            const std::string_view qid = groups[(size_t)q].qid;
            for (size_t r = 0; r < keep; ++r) {
                if (qid.size() + v[r].docid_len + run_len + 80 > sizeof(line)) { bad[(size_t)t] = 1; return; }
                char* o = line;
                memcpy(o, qid.data(), qid.size()); o += qid.size();
                memcpy(o, " Q0 ", 4); o += 4;
                memcpy(o, v[r].docid, v[r].docid_len); o += v[r].docid_len;
                *o++ = ' ';
                o = std::to_chars(o, o + 12, (int)r + 1).ptr;
                *o++ = ' ';
                o += py_float_repr(v[r].score, o);
                *o++ = ' ';
                memcpy(o, run_name, run_len); o += run_len;
                *o++ = '\n';
                b.append(line, (size_t)(o - line));
                ++lines[(size_t)t];
            }
        }
    };
    std::vector<std::thread> pool;
    for (int t = 1; t < n_threads; ++t) pool.emplace_back(work, t);
    if (nq > 0) work(0);
    for (auto& th : pool) th.join();
    for (int t = 0; t < n_threads; ++t) if (bad[(size_t)t]) return DHR_ERR_INVALID;
    FILE* f = fopen(out_path, "wb");
    if (!f) return DHR_ERR_INVALID;
    int64_t total = 0;
    bool ok = true;
    for (int t = 0; t < n_threads; ++t) {
        if (!bufs[(size_t)t].empty() && fwrite(bufs[(size_t)t].data(), 1, bufs[(size_t)t].size(), f) != bufs[(size_t)t].size()) ok = false;
        total += lines[(size_t)t];
    }
    if (fclose(f) != 0) ok = false;
    if (lines_written) *lines_written = total;
    return ok ? DHR_OK : DHR_ERR_INVALID;
}
