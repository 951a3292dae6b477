// Host FASTA ingest with the reference's parsing semantics (replaces readfasta, PolyFastA.py:227-250).
//
// The reference iterates text-mode lines (universal newlines: "\n", "\r\n" and a lone "\r" all end a
// line); a line whose first character is '>' opens a record whose header is the rest of the line,
// right-stripped (:233,:242); the same header seen again restarts that record but keeps its position in
// the dict (:234,:243); any other line is right-stripped, upper-cased and appended to the current record
// iff the current header is non-empty (:235-236,:244-245).  If the LAST header seen is empty (or none was
// seen) the file "is not FASTA" (:246-248).  Upper-casing is deferred to the device encoder (its LUT folds
// case) and to pfa_fasta_copy_row.
//
// Output: rows in first-seen order in one arena; when all rows have the same length L the arena is the
// row-major matrix text[row*L + col] that pfa_aln_from_fasta uploads.
#include <fcntl.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>

#include <algorithm>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <string>
#include <thread>
#include <unordered_map>
#include <vector>

#include "pfa_host.h"

namespace {
// str.rstrip() with no argument strips characters for which str.isspace() is true; in ASCII these are
// \t \n \v \f \r, the separators 0x1c-0x1f and the blank.
inline bool py_space(unsigned char c) { return (c >= 0x09 && c <= 0x0d) || (c >= 0x1c && c <= 0x20); }
}  // namespace

// line scan: records with their line slices (pointing into buf), lengths, and whether all rows have one length
int pfa_parse_lines(const unsigned char* buf, size_t len, PfaParsed* out) {
    std::vector<PfaRecord>& recs = out->recs;
    recs.clear();
    std::unordered_map<std::string, size_t> index;
    bool seen_header = false, head_nonempty = false;
    size_t cur = 0;
    size_t i = 0;
    // a '\r' anywhere switches to the byte-wise scan (a lone '\r' ends a line under universal newlines); otherwise lines
    // end at '\n' only and memchr finds them at memory speed
    const bool has_cr = len && memchr(buf, '\r', len) != nullptr;
    while (i < len) {
        size_t e;
        if (has_cr) {
            e = i;
            while (e < len && buf[e] != '\n' && buf[e] != '\r') ++e;
        } else {
            const void* nl = memchr(buf + i, '\n', len - i);
            e = nl ? (size_t)(static_cast<const unsigned char*>(nl) - buf) : len;
        }
        size_t next = e;
        if (next < len) next += (buf[next] == '\r' && next + 1 < len && buf[next + 1] == '\n') ? 2 : 1;
        // the line is buf[i, e) plus its terminator; never empty as a Python line, but may be empty here
        size_t r = e;
        while (r > i && py_space(buf[r - 1])) --r;
        if (e > i && buf[i] == '>') {
            std::string h(reinterpret_cast<const char*>(buf + i + 1), r > i + 1 ? r - (i + 1) : 0);
            seen_header = true;
            head_nonempty = !h.empty();
            auto it = index.find(h);
            if (it == index.end()) {
                index.emplace(h, recs.size());
                cur = recs.size();
                recs.push_back(PfaRecord{h, {}, 0});
            } else {
                cur = it->second;
                recs[cur].parts.clear();
                recs[cur].len = 0;
            }
        } else if (seen_header && head_nonempty && r > i) {
            recs[cur].parts.push_back(PfaSlice{buf + i, r - i});
            recs[cur].len += (int64_t)(r - i);
        }
        i = next;
    }
    if (!seen_header || !head_nonempty) return PFA_ERR_NOT_FASTA;
    out->total = 0;
    out->same = true;
    for (const PfaRecord& rec : recs) {
        out->total += rec.len;
        if (rec.len != recs[0].len) out->same = false;
    }
    out->seqlen = out->same ? recs[0].len : -1;
    return PFA_OK;
}

// copy the line slices of one record to dst; returns false when a byte >= 0x80 was seen
bool pfa_copy_record(const PfaRecord& rec, unsigned char* dst) {
    unsigned char acc = 0;
    for (const PfaSlice& s : rec.parts) {
        memcpy(dst, s.p, s.len);
        for (size_t b = 0; b < s.len; ++b) acc |= s.p[b];
        dst += s.len;
    }
    return !(acc & 0x80);
}

static int parse_impl(const unsigned char* buf, size_t len, pfa_fasta** out) {
    PfaParsed parsed;
    int rc = pfa_parse_lines(buf, len, &parsed);
    if (rc) return rc;
    const std::vector<PfaRecord>& recs = parsed.recs;
    pfa_fasta* f = new pfa_fasta();
    f->n = (int64_t)recs.size();
    f->row_len.resize(recs.size());
    f->row_off.resize(recs.size() + 1);
    int64_t total = 0;
    for (size_t k = 0; k < recs.size(); ++k) {
        f->row_len[k] = recs[k].len;
        f->row_off[k] = total;
        total += recs[k].len;
        f->header_off.push_back((int64_t)f->headers.size());
        f->headers += recs[k].header;
    }
    f->row_off[recs.size()] = total;
    f->header_off.push_back((int64_t)f->headers.size());
    f->seqlen = parsed.seqlen;
    f->data_bytes = (size_t)std::max<int64_t>(total, 1);
    f->data = (unsigned char*)malloc(f->data_bytes);
    if (!f->data) {
        delete f;
        return PFA_ERR_NOMEM;
    }
    // rows are independent, so large files are copied by several threads
    unsigned nthreads = total > (64ll << 20) ? std::min<unsigned>(16, std::max(1u, std::thread::hardware_concurrency())) : 1;
    std::vector<char> bad(nthreads, 0);
    auto work = [&](unsigned t) {
        for (size_t k = t; k < recs.size(); k += nthreads)
            if (!pfa_copy_record(recs[k], f->data + f->row_off[k])) bad[t] = 1;
    };
    if (nthreads == 1) {
        work(0);
    } else {
        std::vector<std::thread> th;
        for (unsigned t = 0; t < nthreads; ++t) th.emplace_back(work, t);
        for (auto& x : th) x.join();
    }
    bool non_ascii = false;
    for (char b : bad) non_ascii |= (b != 0);
    if (non_ascii) {
        pfa_fasta_free(f);
        return PFA_ERR_NON_ASCII;
    }
    *out = f;
    return PFA_OK;
}

// whole file into a reusable buffer (read() rather than mmap: no page fault per 4 KB for the many small files of --dir)
int pfa_read_file(const char* path, std::vector<unsigned char>* buf, size_t* len) {
    int fd = open(path, O_RDONLY);
    if (fd < 0) return PFA_ERR_IO;
    struct stat st;
    if (fstat(fd, &st) != 0 || S_ISDIR(st.st_mode)) {
        close(fd);
        return PFA_ERR_IO;
    }
    const size_t size = (size_t)st.st_size;
    if (buf->size() < size) buf->resize(size);
    size_t got = 0;
    while (got < size) {
        const ssize_t r = read(fd, buf->data() + got, size - got);
        if (r < 0) {
            close(fd);
            return PFA_ERR_IO;
        }
        if (r == 0) break;
        got += (size_t)r;
    }
    close(fd);
    *len = got;  // the buffer keeps its size so that the next, equally large file does not zero-fill it again
    return PFA_OK;
}

extern "C" {

int pfa_fasta_parse_buffer(const void* buf, size_t len, pfa_fasta** out) {
    if (!out || (!buf && len)) return PFA_ERR_ARG;
    *out = nullptr;
    return parse_impl(static_cast<const unsigned char*>(buf), len, out);
}

int pfa_fasta_parse_file(const char* path, pfa_fasta** out) {
    if (!out || !path) return PFA_ERR_ARG;
    *out = nullptr;
    static thread_local std::vector<unsigned char> buf;  // warm across the files one thread parses
    size_t n = 0;
    int rc = pfa_read_file(path, &buf, &n);
    if (rc) return rc;
    return parse_impl(buf.data(), n, out);
}

int pfa_fasta_parse_files(const char* const* paths, int count, int threads, pfa_fasta** out, int* status) {
    if (count < 0 || (count > 0 && (!paths || !out || !status))) return PFA_ERR_ARG;
    if (threads < 1) threads = (int)std::max(1u, std::thread::hardware_concurrency());
    threads = std::min(threads, std::max(count, 1));
    auto work = [&](int t) {
        for (int i = t; i < count; i += threads) {
            out[i] = nullptr;
            status[i] = pfa_fasta_parse_file(paths[i], &out[i]);
        }
    };
    if (threads == 1) {
        work(0);
    } else {
        std::vector<std::thread> th;
        for (int t = 0; t < threads; ++t) th.emplace_back(work, t);
        for (auto& x : th) x.join();
    }
    return PFA_OK;
}

int64_t pfa_fasta_match_mask(const pfa_fasta* f, const char* key, int64_t key_len, uint32_t* mask, int64_t mask_words) {
    if (!f || !mask || key_len < 0 || (key_len > 0 && !key) || mask_words * 32 < f->n) return -1;
    memset(mask, 0, sizeof(uint32_t) * (size_t)mask_words);
    int64_t hits = 0;
    for (int64_t r = 0; r < f->n; ++r) {
        const char* h = f->headers.data() + f->header_off[(size_t)r];
        const size_t hl = (size_t)(f->header_off[(size_t)r + 1] - f->header_off[(size_t)r]);
        const bool hit = key_len == 0 || (hl >= (size_t)key_len && memmem(h, hl, key, (size_t)key_len) != nullptr);
        if (hit) {
            mask[r >> 5] |= 1u << (r & 31);
            ++hits;
        }
    }
    return hits;
}

void pfa_fasta_free(pfa_fasta* f) {
    if (!f) return;
    if (f->unpin) f->unpin(f->data);
    free(f->data);
    delete f;
}

int64_t pfa_fasta_nseq(const pfa_fasta* f) { return f ? f->n : 0; }
int64_t pfa_fasta_seqlen(const pfa_fasta* f) { return f ? f->seqlen : -1; }
int64_t pfa_fasta_row_len(const pfa_fasta* f, int64_t row) {
    return (f && row >= 0 && row < f->n) ? f->row_len[(size_t)row] : -1;
}
const char* pfa_fasta_header(const pfa_fasta* f, int64_t row, int64_t* len) {
    if (!f || row < 0 || row >= f->n) return nullptr;
    if (len) *len = f->header_off[(size_t)row + 1] - f->header_off[(size_t)row];
    return f->headers.data() + f->header_off[(size_t)row];
}
int pfa_fasta_copy_row(const pfa_fasta* f, int64_t row, uint8_t* dst, int64_t cap) {
    if (!f || row < 0 || row >= f->n || !dst || cap < f->row_len[(size_t)row]) return PFA_ERR_ARG;
    const unsigned char* src = f->data + f->row_off[(size_t)row];
    for (int64_t i = 0; i < f->row_len[(size_t)row]; ++i) {
        unsigned char c = src[i];
        dst[i] = (c >= 'a' && c <= 'z') ? (unsigned char)(c - 32) : c;  // str.upper() on ASCII
    }
    return PFA_OK;
}

}  // extern "C"
