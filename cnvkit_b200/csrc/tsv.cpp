// Host-side reader / writer of CNVkit's tab-separated tables (.cnn / .cnr / .cns): the boundary of
// the hot path.  Restates what skgenome/tabio/tab.py:18-33 (pandas.read_csv
// with sep="\t") and skgenome/tabio/__init__.py:175-195 (DataFrame.to_csv(sep="\t",
// float_format="%.6g", index=False)) do for these files, without pandas: one pass to index the
// fields, then per-column conversion with std::from_chars (correctly rounded, like pandas' parser on
// the <= 15-digit decimals these files hold) and printf-exact "%.6g" on the way out.
//
// No CUDA here: the columns land in caller-provided host buffers (pinned, if the caller wants to
// DMA them straight to HBM).
#include <algorithm>
#include <atomic>
#include <charconv>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <deque>
#include <limits>
#include <string>
#include <string_view>
#include <thread>
#include <unordered_map>
#include <vector>

#include "../../include/cnvkit_b200.h"
namespace cnvk {
void set_last_error(const char* fmt, ...);  // capi.cu
}

namespace {

struct Field {
    uint64_t off;
    uint32_t len;
    uint32_t quoted;  // field was written as "..." (inner "" are unescaped lazily)
};

struct Table {
    std::string data;
    std::vector<std::string> names;
    std::vector<Field> fields;  // row-major [nrows][ncols]
    int64_t nrows = 0;
    int ncols = 0;
};

// pandas' default na_values for read_csv
const char* const kNa[] = {"",     "#N/A", "#N/A N/A", "#NA",  "-1.#IND", "-1.#QNAN", "-NaN", "-nan", "1.#IND",
                           "1.#QNAN", "<NA>", "N/A",   "NA",   "NULL",    "NaN",      "None", "n/a",  "nan", "null"};

bool is_na(const char* p, size_t n) {
    if (n == 0) return true;
    if (n > 8) return false;
    for (const char* s : kNa)
        if (strlen(s) == n && memcmp(s, p, n) == 0) return true;
    return false;
}

// Splits one line into fields; returns the offset of the next line.  Minimal RFC-4180 quoting, as
// pandas' C tokenizer applies to sep="\t".
size_t split_line(const std::string& d, size_t pos, std::vector<Field>& out) {
    const size_t n = d.size();
    for (;;) {
        Field f{pos, 0, 0};
        if (pos < n && d[pos] == '"') {
            f.quoted = 1;
            size_t q = pos + 1;
            f.off = q;
            while (q < n) {
                if (d[q] == '"') {
                    if (q + 1 < n && d[q + 1] == '"') { q += 2; continue; }
                    break;
                }
                ++q;
            }
            f.len = (uint32_t)(q - f.off);
            pos = q < n ? q + 1 : q;
        } else {
            size_t q = pos;
            while (q < n && d[q] != '\t' && d[q] != '\n' && d[q] != '\r') ++q;
            f.len = (uint32_t)(q - pos);
            pos = q;
        }
        out.push_back(f);
        if (pos < n && d[pos] == '\t') { ++pos; continue; }
        if (pos < n && d[pos] == '\r') ++pos;
        if (pos < n && d[pos] == '\n') ++pos;
        return pos;
    }
}

std::string field_text(const Table& t, const Field& f) {
    std::string s(t.data.data() + f.off, f.len);
    if (f.quoted) {
        size_t k;
        size_t from = 0;
        while ((k = s.find("\"\"", from)) != std::string::npos) { s.erase(k, 1); from = k + 1; }
    }
    return s;
}

bool parse_i64(const char* p, size_t n, int64_t* v) {
    while (n && (*p == ' ')) { ++p; --n; }
    while (n && (p[n - 1] == ' ')) --n;
    if (n && *p == '+') { ++p; --n; }
    if (!n) return false;
    auto r = std::from_chars(p, p + n, *v);
    return r.ec == std::errc() && r.ptr == p + n;
}

bool parse_f64(const char* p, size_t n, double* v) {
    while (n && (*p == ' ')) { ++p; --n; }
    while (n && (p[n - 1] == ' ')) --n;
    if (n && *p == '+') { ++p; --n; }
    if (!n) return false;
    auto r = std::from_chars(p, p + n, *v);
    if (r.ec == std::errc() && r.ptr == p + n) return true;
    // "inf", "-inf", "Infinity" spellings pandas also accepts
    std::string s(p, n);
    for (auto& c : s) c = (char)tolower(c);
    if (s == "inf" || s == "infinity") { *v = INFINITY; return true; }
    if (s == "-inf" || s == "-infinity") { *v = -INFINITY; return true; }
    return false;
}

int fail(const char* fmt, const char* a) {
    cnvk::set_last_error(fmt, a);
    return -1;
}

}  // namespace

struct cnvk_tsv {
    Table t;
};

// one string field, quoted the way csv.QUOTE_MINIMAL does (tab, quote, CR or LF inside)
static void append_field(std::string& out, const char* s, size_t len) {
    bool q = false;
    for (size_t k = 0; k < len; ++k)
        if (s[k] == '\t' || s[k] == '"' || s[k] == '\n' || s[k] == '\r') { q = true; break; }
    if (!q) { out.append(s, len); return; }
    out += '"';
    for (size_t k = 0; k < len; ++k) { if (s[k] == '"') out += '"'; out += s[k]; }
    out += '"';
}

extern "C" {

int cnvk_tsv_open(const char* path, cnvk_tsv** out) {
    *out = nullptr;
    FILE* fp = fopen(path, "rb");
    if (!fp) return fail("cnvk_tsv_open: cannot open %s", path);
    auto* h = new cnvk_tsv();
    Table& t = h->t;
    char buf[1 << 16];
    size_t got;
    while ((got = fread(buf, 1, sizeof(buf), fp)) > 0) t.data.append(buf, got);
    fclose(fp);
    if (t.data.size() >= 3 && memcmp(t.data.data(), "\xEF\xBB\xBF", 3) == 0) t.data.erase(0, 3);
    size_t pos = 0;
    std::vector<Field> row;
    if (t.data.empty()) { *out = h; return 0; }  // blank file: zero columns (pandas EmptyDataError)
    pos = split_line(t.data, 0, row);
    t.ncols = (int)row.size();
    for (const Field& f : row) t.names.push_back(field_text(t, f));
    t.fields.reserve((size_t)(t.data.size() / 24) * (size_t)std::max(t.ncols, 1) / 2);
    while (pos < t.data.size()) {
        // skip blank lines (pandas skip_blank_lines=True)
        if (t.data[pos] == '\n') { ++pos; continue; }
        if (t.data[pos] == '\r' && pos + 1 < t.data.size() && t.data[pos + 1] == '\n') { pos += 2; continue; }
        row.clear();
        pos = split_line(t.data, pos, row);
        // short rows are padded with missing fields, as pandas' tokenizer does; long rows are an error
        while ((int)row.size() < t.ncols) row.push_back(Field{0, 0, 0});
        if ((int)row.size() != t.ncols) {
            char msg[160];
            snprintf(msg, sizeof(msg), "row %lld has %d fields, header has %d", (long long)t.nrows + 2, (int)row.size(),
                     t.ncols);
            delete h;
            return fail("cnvk_tsv_open: %s", msg);
        }
        t.fields.insert(t.fields.end(), row.begin(), row.end());
        ++t.nrows;
    }
    *out = h;
    return 0;
}

void cnvk_tsv_close(cnvk_tsv* h) { delete h; }

int cnvk_tsv_shape(const cnvk_tsv* h, int64_t* nrows, int32_t* ncols) {
    *nrows = h->t.nrows;
    *ncols = h->t.ncols;
    return 0;
}

const char* cnvk_tsv_colname(const cnvk_tsv* h, int32_t col) {
    return (col >= 0 && col < h->t.ncols) ? h->t.names[col].c_str() : nullptr;
}

// pandas' dtype inference for one column: 1 = int64 (every field an integer, none missing),
// 2 = float64 (numbers and missing values), 0 = object/str.
int cnvk_tsv_col_kind(const cnvk_tsv* h, int32_t col) {
    const Table& t = h->t;
    if (col < 0 || col >= t.ncols) return -1;
    bool all_int = true;
    for (int64_t r = 0; r < t.nrows; ++r) {
        const Field& f = t.fields[(size_t)r * t.ncols + col];
        const char* p = t.data.data() + f.off;
        if (f.quoted) return 0;
        int64_t iv;
        if (all_int && f.len && parse_i64(p, f.len, &iv)) continue;
        all_int = false;
        double dv;
        if (f.len && parse_f64(p, f.len, &dv)) continue;
        if (!is_na(p, f.len)) return 0;
    }
    return all_int ? 1 : 2;
}

int cnvk_tsv_read_f64(const cnvk_tsv* h, int32_t col, double* out) {
    const Table& t = h->t;
    if (col < 0 || col >= t.ncols) return fail("cnvk_tsv_read_f64: %s", "bad column");
    for (int64_t r = 0; r < t.nrows; ++r) {
        const Field& f = t.fields[(size_t)r * t.ncols + col];
        const char* p = t.data.data() + f.off;
        double v;
        if (f.len && parse_f64(p, f.len, &v)) { out[r] = v; continue; }   // a number, the common case
        if (!is_na(p, f.len)) return fail("cnvk_tsv_read_f64: not a number in column %s", t.names[col].c_str());
        out[r] = std::numeric_limits<double>::quiet_NaN();
    }
    return 0;
}

int cnvk_tsv_read_i64(const cnvk_tsv* h, int32_t col, int64_t* out) {
    const Table& t = h->t;
    if (col < 0 || col >= t.ncols) return fail("cnvk_tsv_read_i64: %s", "bad column");
    for (int64_t r = 0; r < t.nrows; ++r) {
        const Field& f = t.fields[(size_t)r * t.ncols + col];
        if (!parse_i64(t.data.data() + f.off, f.len, &out[r]))
            return fail("cnvk_tsv_read_i64: not an integer in column %s", t.names[col].c_str());
    }
    return 0;
}

// String column as one byte buffer + offsets[nrows+1].  With buf == NULL only *nbytes is filled.
int cnvk_tsv_read_str(const cnvk_tsv* h, int32_t col, int64_t* offsets, char* buf, int64_t cap, int64_t* nbytes) {
    const Table& t = h->t;
    if (col < 0 || col >= t.ncols) return fail("cnvk_tsv_read_str: %s", "bad column");
    int64_t tot = 0;
    for (int64_t r = 0; r < t.nrows; ++r) {
        const Field& f = t.fields[(size_t)r * t.ncols + col];
        if (buf) {
            if (f.quoted) {
                const std::string s = field_text(t, f);
                if (tot + (int64_t)s.size() > cap) return fail("cnvk_tsv_read_str: %s", "buffer too small");
                memcpy(buf + tot, s.data(), s.size());
                offsets[r] = tot;
                tot += (int64_t)s.size();
                continue;
            }
            if (tot + f.len > cap) return fail("cnvk_tsv_read_str: %s", "buffer too small");
            memcpy(buf + tot, t.data.data() + f.off, f.len);
            offsets[r] = tot;
        }
        tot += f.len;
    }
    if (buf) offsets[t.nrows] = tot;
    *nbytes = tot;
    return 0;
}

// String column dictionary-coded in order of first appearance: codes[nrows] plus the distinct values as
// offsets[n_unique + 1] into one byte buffer.  With codes == NULL only *n_unique / *nbytes are filled
// (sizes for the second call).  A table has few distinct chromosomes and gene labels, so the caller
// builds a handful of Python strings instead of one per row.
int cnvk_tsv_read_str_dict(const cnvk_tsv* h, int32_t col, int32_t* codes, int64_t* offsets, char* buf, int64_t cap,
                           int64_t* n_unique, int64_t* nbytes) {
    const Table& t = h->t;
    if (col < 0 || col >= t.ncols) return fail("cnvk_tsv_read_str_dict: %s", "bad column");
    std::unordered_map<std::string_view, int32_t> seen;
    std::deque<std::string> owned;           // unescaped text of quoted fields (a deque keeps addresses stable)
    std::vector<std::string_view> uniq;
    int64_t tot = 0;
    std::string_view prev;
    int32_t prev_code = -1;
    for (int64_t r = 0; r < t.nrows; ++r) {
        const Field& f = t.fields[(size_t)r * t.ncols + col];
        std::string_view v(t.data.data() + f.off, f.len);
        int32_t code;
        if (!f.quoted && prev_code >= 0 && v == prev) {
            code = prev_code;                // runs of equal labels (chromosome, gene) skip the hash
        } else {
            if (f.quoted) {
                std::string s = field_text(t, f);
                auto it0 = seen.find(std::string_view(s));
                if (it0 == seen.end()) {
                    owned.push_back(std::move(s));
                    v = std::string_view(owned.back());
                } else {
                    v = it0->first;
                }
            }
            auto it = seen.find(v);
            if (it == seen.end()) {
                code = (int32_t)uniq.size();
                seen.emplace(v, code);
                uniq.push_back(v);
                tot += (int64_t)v.size();
            } else {
                code = it->second;
            }
            if (!f.quoted) { prev = v; prev_code = code; }
        }
        if (codes) codes[r] = code;
    }
    *n_unique = (int64_t)uniq.size();
    *nbytes = tot;
    if (codes) {
        if (tot > cap) return fail("cnvk_tsv_read_str_dict: %s", "buffer too small");
        int64_t pos = 0;
        for (size_t u = 0; u < uniq.size(); ++u) {
            offsets[u] = pos;
            memcpy(buf + pos, uniq[u].data(), uniq[u].size());
            pos += (int64_t)uniq[u].size();
        }
        offsets[uniq.size()] = pos;
    }
    return 0;
}

// FNV-1a over the raw bytes of the given columns, row by row: lets a cohort reader check that a
// sample's bin columns equal the panel's without materialising any string.
uint64_t cnvk_tsv_hash_cols(const cnvk_tsv* h, const int32_t* cols, int32_t n) {
    const Table& t = h->t;
    uint64_t x = 1469598103934665603ull;
    for (int64_t r = 0; r < t.nrows; ++r)
        for (int k = 0; k < n; ++k) {
            const Field& f = t.fields[(size_t)r * t.ncols + cols[k]];
            const unsigned char* p = (const unsigned char*)t.data.data() + f.off;
            for (uint32_t b = 0; b < f.len; ++b) { x ^= p[b]; x *= 1099511628211ull; }
            x ^= 0x1f; x *= 1099511628211ull;
        }
    return x;
}

// DataFrame.to_csv(path, sep="\t", index=False, float_format="%.6g"): kinds[c] = 0 str
// (ptrs[c] = int64 offsets[nrows+1], ptrs2[c] = bytes), 1 int64, 2 float64 (NaN -> empty field),
// 3 dictionary-coded str (ptrs[c] = int32 codes[nrows], -1 = missing; ptrs2[c] = cnvk_tsv_dict*).
// Rows are formatted in parallel (contiguous row ranges, one buffer each) and written in order.
int cnvk_tsv_write(const char* path, int32_t ncols, const char* const* names, const int32_t* kinds,
                   const void* const* ptrs, const void* const* ptrs2, int64_t nrows) {
    // dictionary columns: render (quote) every distinct value once
    std::vector<std::vector<std::string>> dict((size_t)ncols);
    for (int c = 0; c < ncols; ++c) {
        if (kinds[c] != 3) continue;
        const cnvk_tsv_dict* d = (const cnvk_tsv_dict*)ptrs2[c];
        dict[c].resize((size_t)d->n);
        for (int64_t u = 0; u < d->n; ++u)
            append_field(dict[c][(size_t)u], d->blob + d->offsets[u], (size_t)(d->offsets[u + 1] - d->offsets[u]));
    }
    for (int c = 0; c < ncols; ++c) {
        if (kinds[c] != 3) continue;
        const int32_t* codes = (const int32_t*)ptrs[c];
        const int64_t nu = (int64_t)dict[c].size();
        for (int64_t r = 0; r < nrows; ++r)
            if (codes[r] >= nu) return fail("cnvk_tsv_write: dictionary code out of range in column %s", names[c]);
    }
    auto format_rows = [&](int64_t r0, int64_t r1, std::string& out) {
        char num[64];
        out.reserve((size_t)(r1 - r0) * (size_t)ncols * 9 + 64);
        for (int64_t r = r0; r < r1; ++r) {
            for (int c = 0; c < ncols; ++c) {
                if (c) out += '\t';
                if (kinds[c] == 1) {
                    auto res = std::to_chars(num, num + sizeof(num), ((const int64_t*)ptrs[c])[r]);
                    out.append(num, res.ptr - num);
                } else if (kinds[c] == 2) {
                    const double v = ((const double*)ptrs[c])[r];
                    if (v == v) {
                        if (std::isinf(v)) out += (v > 0 ? "inf" : "-inf");
                        else {   // "as if by printf %.6g in the C locale"
                            auto res = std::to_chars(num, num + sizeof(num), v, std::chars_format::general, 6);
                            out.append(num, res.ptr - num);
                        }
                    }
                } else if (kinds[c] == 3) {
                    const int32_t code = ((const int32_t*)ptrs[c])[r];
                    if (code >= 0) out += dict[c][(size_t)code];
                } else {
                    const int64_t* off = (const int64_t*)ptrs[c];
                    append_field(out, (const char*)ptrs2[c] + off[r], (size_t)(off[r + 1] - off[r]));
                }
            }
            out += '\n';
        }
    };
    FILE* fp = fopen(path, "wb");
    if (!fp) return fail("cnvk_tsv_write: cannot open %s", path);
    std::string head;
    for (int c = 0; c < ncols; ++c) { if (c) head += '\t'; head += names[c]; }
    head += '\n';
    fwrite(head.data(), 1, head.size(), fp);
    const int64_t chunk = 32768;
    int nthreads = (int)std::min<int64_t>(std::min<unsigned>(std::max(1u, std::thread::hardware_concurrency()), 4u),
                                          (nrows + chunk - 1) / chunk);
    if (const char* env = getenv("CNVK_TSV_THREADS")) nthreads = std::max(1, std::min(nthreads, atoi(env)));
    if (nthreads <= 1) {
        std::string out;
        for (int64_t r0 = 0; r0 < nrows; r0 += chunk) {
            out.clear();
            format_rows(r0, std::min(nrows, r0 + chunk), out);
            fwrite(out.data(), 1, out.size(), fp);
        }
    } else {
        const int64_t nchunks = (nrows + chunk - 1) / chunk;
        std::vector<std::string> parts((size_t)nchunks);
        std::atomic<int64_t> next{0};
        std::vector<std::thread> pool;
        for (int t = 0; t < nthreads; ++t)
            pool.emplace_back([&]() {
                for (;;) {
                    const int64_t k = next.fetch_add(1);
                    if (k >= nchunks) break;
                    format_rows(k * chunk, std::min(nrows, (k + 1) * chunk), parts[(size_t)k]);
                }
            });
        for (auto& th : pool) th.join();
        for (auto& p : parts) fwrite(p.data(), 1, p.size(), fp);
    }
    if (fclose(fp) != 0) return fail("cnvk_tsv_write: write failed for %s", path);
    return 0;
}

// skgenome/combiners.py:58-94 join_strings over bin ranges, the host half of transfer_fields
// (cnvlib/segmentation/__init__.py:467-501): for every row range [lo[s], hi[s]) the distinct gene codes in order
// of first appearance, skipping codes < 0 (missing gene) and codes without a kept name (keep[code] == 0: the
// ignored placeholder / antitarget names).  CSR output: out_off[nseg + 1], out_codes[<= capacity].  One pass over
// the rows with a per-code "last segment seen" stamp -- O(rows), no hashing.  Optional nested-bin mask
// (intersect.py _irange_nested): a row counts only if mask_end[row] > mask_start[s].
int cnvk_ordered_unique_ranges(const int64_t* codes, const int64_t* lo, const int64_t* hi, int32_t nseg, int64_t ncodes,
                               const uint8_t* keep, const int64_t* mask_end, const int64_t* mask_start, int64_t* out_off,
                               int64_t* out_codes, int64_t capacity) {
    std::vector<int32_t> stamp((size_t)std::max<int64_t>(ncodes, 1), -1);
    int64_t pos = 0;
    for (int32_t s = 0; s < nseg; ++s) {
        out_off[s] = pos;
        for (int64_t r = lo[s]; r < hi[s]; ++r) {
            const int64_t c = codes[r];
            if (c < 0 || c >= ncodes) continue;
            if (keep && !keep[c]) continue;
            if (mask_end && !(mask_end[r] > mask_start[s])) continue;
            if (stamp[(size_t)c] == s) continue;
            stamp[(size_t)c] = s;
            if (pos >= capacity) return fail("cnvk_ordered_unique_ranges: %s", "output capacity exceeded");
            out_codes[pos++] = c;
        }
    }
    out_off[nseg] = pos;
    return 0;
}

// The same pass, emitting the joined names: name_blob / name_off[ncodes + 1] hold one kept name per code (empty for
// codes without one), the result is "a,b,c" per range ("-" when nothing remains) in out_blob / out_off[nseg + 1].
// Only valid when every code carries at most one name and no name is shared between codes (the caller checks).
int cnvk_join_names_ranges(const int64_t* codes, const int64_t* lo, const int64_t* hi, int32_t nseg, int64_t ncodes,
                           const char* name_blob, const int64_t* name_off, const int64_t* mask_end,
                           const int64_t* mask_start, char* out_blob, int64_t* out_off, int64_t capacity) {
    std::vector<int32_t> stamp((size_t)std::max<int64_t>(ncodes, 1), -1);
    int64_t pos = 0;
    for (int32_t s = 0; s < nseg; ++s) {
        out_off[s] = pos;
        bool first = true;
        for (int64_t r = lo[s]; r < hi[s]; ++r) {
            const int64_t c = codes[r];
            if (c < 0 || c >= ncodes) continue;
            const int64_t len = name_off[c + 1] - name_off[c];
            if (len == 0) continue;
            if (mask_end && !(mask_end[r] > mask_start[s])) continue;
            if (stamp[(size_t)c] == s) continue;
            stamp[(size_t)c] = s;
            if (pos + len + 1 > capacity) return fail("cnvk_join_names_ranges: %s", "output capacity exceeded");
            if (!first) out_blob[pos++] = ',';
            memcpy(out_blob + pos, name_blob + name_off[c], (size_t)len);
            pos += len;
            first = false;
        }
        if (first) {
            if (pos + 1 > capacity) return fail("cnvk_join_names_ranges: %s", "output capacity exceeded");
            out_blob[pos++] = '-';
        }
    }
    out_off[nseg] = pos;
    return 0;
}

}  // extern "C"
