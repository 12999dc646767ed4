// latok_reader.cpp -- host-side ingest of the C ABI: csv / csv.gz rows -> packed batch (flat UTF-8 + int64 offsets).
//
// Counterpart of the reference's per-row Python loop in scripts/timing/time_tokenizer.py:25-40
// (`for row in csv.reader(infile): text = json.loads(row[1]).strip()`), SURVEY 8 f3: the rows are parsed, JSON-decoded,
// stripped and packed here, straight into the caller's (ideally pinned, latok_b200_host_alloc) buffers, so that a batch
// can go to latok_b200_submit without touching a Python object per row.  Pure host code; no tokenization happens here.
#include "../../include/latok_b200.h"

#include <zlib.h>

#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <new>
#include <string>

extern "C" int latok_b200_set_error_(int code, const char *fmt, ...);   // latok_capi.cu

struct latok_b200_reader {
    gzFile f = nullptr;              // gzopen reads plain files transparently
    int column = 1;
    unsigned char buf[1 << 16];
    int len = 0, pos = 0;
    bool eof = false;
    long long row = 0;               // records consumed so far (for messages)
    std::string pending;             // a decoded row that did not fit the previous call's buffer
    bool has_pending = false;
    std::string field, text;
};

namespace {

int next_byte(latok_b200_reader *r)
{
    if (r->pos == r->len) {
        if (r->eof) return -1;
        const int n = gzread(r->f, r->buf, sizeof r->buf);
        if (n <= 0) { r->eof = true; return n < 0 ? -2 : -1; }
        r->len = n; r->pos = 0;
    }
    return r->buf[r->pos++];
}

// One csv record (dialect 'excel': ',' delimiter, '"' quote, doubled quote inside quotes, records end at \n, \r\n or
// \r outside quotes).  Returns 1 and the wanted column in r->field, 0 at end of file, <0 on error.
int read_record(latok_b200_reader *r, bool &has_column)
{
    r->field.clear();
    has_column = false;
    int col = 0, c = next_byte(r);
    if (c == -2) return -2;
    if (c < 0) return 0;
    bool any = false;
    for (;;) {
        // ---- one field
        const bool want = col == r->column;
        if (c == '"') {
            for (;;) {
                c = next_byte(r);
                if (c < 0) return c == -2 ? -2 : -3;            // end of file inside quotes
                if (c == '"') {
                    c = next_byte(r);
                    if (c == '"') { if (want) r->field.push_back('"'); continue; }
                    break;
                }
                if (want) r->field.push_back((char)c);
            }
        }
        while (c >= 0 && c != ',' && c != '\n' && c != '\r') {
            if (want) r->field.push_back((char)c);
            c = next_byte(r);
        }
        any = true;
        if (want) has_column = true;
        if (c == -2) return -2;
        if (c == ',') { ++col; c = next_byte(r); if (c == -2) return -2; if (c < 0) { if (col == r->column) has_column = true; break; } continue; }
        if (c == '\r') {                                        // \r\n counts once
            c = next_byte(r);
            if (c == -2) return -2;
            if (c >= 0 && c != '\n') --r->pos;
        }
        break;
    }
    return any ? 1 : 0;
}

void put_utf8(std::string &s, unsigned cp)
{
    if (cp < 0x80) s.push_back((char)cp);
    else if (cp < 0x800) { s.push_back((char)(0xC0 | (cp >> 6))); s.push_back((char)(0x80 | (cp & 0x3F))); }
    else if (cp < 0x10000) {       // lone surrogates included (Python keeps them; the packer uses 'surrogatepass')
        s.push_back((char)(0xE0 | (cp >> 12))); s.push_back((char)(0x80 | ((cp >> 6) & 0x3F))); s.push_back((char)(0x80 | (cp & 0x3F)));
    } else {
        s.push_back((char)(0xF0 | (cp >> 18))); s.push_back((char)(0x80 | ((cp >> 12) & 0x3F)));
        s.push_back((char)(0x80 | ((cp >> 6) & 0x3F))); s.push_back((char)(0x80 | (cp & 0x3F)));
    }
}

int hex4(const std::string &s, size_t i, unsigned &v)
{
    if (i + 4 > s.size()) return -1;
    v = 0;
    for (int k = 0; k < 4; ++k) {
        const unsigned char c = (unsigned char)s[i + k];
        unsigned d;
        if (c >= '0' && c <= '9') d = c - '0';
        else if (c >= 'a' && c <= 'f') d = c - 'a' + 10;
        else if (c >= 'A' && c <= 'F') d = c - 'A' + 10;
        else return -1;
        v = v * 16 + d;
    }
    return 0;
}

// json.loads of one JSON string literal (surrounding JSON whitespace allowed) -> UTF-8
int json_string(const std::string &in, std::string &out)
{
    out.clear();
    size_t i = 0, n = in.size();
    auto ws = [](char c) { return c == ' ' || c == '\t' || c == '\n' || c == '\r'; };
    while (i < n && ws(in[i])) ++i;
    if (i >= n || in[i] != '"') return -1;
    ++i;
    for (;;) {
        if (i >= n) return -1;
        const unsigned char c = (unsigned char)in[i++];
        if (c == '"') break;
        if (c < 0x20) return -1;                                // control characters must be escaped
        if (c != '\\') { out.push_back((char)c); continue; }
        if (i >= n) return -1;
        const char e = in[i++];
        switch (e) {
        case '"': out.push_back('"'); break;
        case '\\': out.push_back('\\'); break;
        case '/': out.push_back('/'); break;
        case 'b': out.push_back('\b'); break;
        case 'f': out.push_back('\f'); break;
        case 'n': out.push_back('\n'); break;
        case 'r': out.push_back('\r'); break;
        case 't': out.push_back('\t'); break;
        case 'u': {
            unsigned v;
            if (hex4(in, i, v)) return -1;
            i += 4;
            if (v >= 0xD800 && v < 0xDC00 && i + 6 <= n && in[i] == '\\' && in[i + 1] == 'u') {
                unsigned lo;
                if (hex4(in, i + 2, lo) == 0 && lo >= 0xDC00 && lo < 0xE000) { v = 0x10000 + ((v - 0xD800) << 10) + (lo - 0xDC00); i += 6; }
            }
            put_utf8(out, v);
            break;
        }
        default: return -1;
        }
    }
    while (i < n && ws(in[i])) ++i;
    return i == n ? 0 : -1;
}

// length of the whitespace character (str.isspace(): SURVEY Q9, the tokenizer's SPACE class) that begins at p, 0: none
int space_at(const unsigned char *p, size_t n)
{
    if (n == 0) return 0;
    const unsigned c = p[0];
    if ((c >= 0x09 && c <= 0x0D) || (c >= 0x1C && c <= 0x20)) return 1;
    if (c == 0xC2 && n >= 2 && (p[1] == 0x85 || p[1] == 0xA0)) return 2;
    if (n >= 3) {
        if (c == 0xE1 && p[1] == 0x9A && p[2] == 0x80) return 3;                                      // U+1680
        if (c == 0xE2 && p[1] == 0x80 && ((p[2] >= 0x80 && p[2] <= 0x8A) || p[2] == 0xA8 || p[2] == 0xA9 || p[2] == 0xAF)) return 3;
        if (c == 0xE2 && p[1] == 0x81 && p[2] == 0x9F) return 3;                                      // U+205F
        if (c == 0xE3 && p[1] == 0x80 && p[2] == 0x80) return 3;                                      // U+3000
    }
    return 0;
}

void strip(const std::string &s, size_t &a, size_t &b)        // str.strip(): [a, b) of s
{
    const unsigned char *p = (const unsigned char *)s.data();
    a = 0; b = s.size();
    for (;;) { const int k = space_at(p + a, b - a); if (!k) break; a += k; }
    for (;;) {
        if (b == a) break;
        size_t q = b - 1;                                       // start of the last character
        while (q > a && (p[q] & 0xC0) == 0x80) --q;
        const int k = space_at(p + q, b - q);
        if (!k || q + k != b) break;
        b = q;
    }
}

}  // namespace

extern "C" {

int latok_b200_reader_open(const char *path, int column, latok_b200_reader **out)
{
    if (!out) return latok_b200_set_error_(LATOK_B200_EINVAL, "out is NULL");
    *out = nullptr;
    if (!path || column < 0) return latok_b200_set_error_(LATOK_B200_EINVAL, "must specify a path and a column >= 0");
    latok_b200_reader *r = new (std::nothrow) latok_b200_reader();
    if (!r) return latok_b200_set_error_(LATOK_B200_ENOMEM, "out of host memory");
    r->f = gzopen(path, "rb");
    if (!r->f) { delete r; return latok_b200_set_error_(LATOK_B200_EINVAL, "cannot open %s", path); }
    gzbuffer(r->f, 1 << 18);
    r->column = column;
    *out = r;
    return LATOK_B200_OK;
}

int latok_b200_reader_next(latok_b200_reader *r, int64_t max_rows, uint8_t *utf8, int64_t utf8_cap, int64_t *offsets,
                           int64_t *n_rows)
{
    if (!r || !offsets || !n_rows) return latok_b200_set_error_(LATOK_B200_EINVAL, "reader, offsets or n_rows is NULL");
    if (max_rows < 0 || utf8_cap < 0 || (utf8_cap > 0 && !utf8)) return latok_b200_set_error_(LATOK_B200_EINVAL, "bad buffer arguments");
    int64_t rows = 0, bytes = 0;
    offsets[0] = 0;
    while (rows < max_rows) {
        if (!r->has_pending) {
            bool has_column = false;
            const int rc = read_record(r, has_column);
            if (rc == 0) break;
            ++r->row;
            if (rc == -2) return latok_b200_set_error_(LATOK_B200_EINVAL, "read error near row %lld", r->row);
            if (rc < 0) return latok_b200_set_error_(LATOK_B200_EINVAL, "row %lld: end of file inside a quoted field", r->row);
            if (!has_column) return latok_b200_set_error_(LATOK_B200_EINVAL, "row %lld has no column %d", r->row, r->column);
            if (json_string(r->field, r->text)) return latok_b200_set_error_(LATOK_B200_EINVAL, "row %lld: column %d is not a JSON string", r->row, r->column);
            size_t a, b;
            strip(r->text, a, b);
            r->pending.assign(r->text, a, b - a);
            r->has_pending = true;
        }
        const int64_t n = (int64_t)r->pending.size();
        if (bytes + n > utf8_cap) {
            if (rows == 0) return latok_b200_set_error_(LATOK_B200_EINVAL, "row %lld (%lld bytes) does not fit the buffer", r->row, (long long)n);
            break;                                              // keep it for the next call
        }
        if (n) memcpy(utf8 + bytes, r->pending.data(), (size_t)n);
        bytes += n;
        offsets[++rows] = bytes;
        r->has_pending = false;
    }
    *n_rows = rows;
    return LATOK_B200_OK;
}

int latok_b200_reader_close(latok_b200_reader *r)
{
    if (!r) return LATOK_B200_OK;
    if (r->f) gzclose(r->f);
    delete r;
    return LATOK_B200_OK;
}

}  // extern "C"
