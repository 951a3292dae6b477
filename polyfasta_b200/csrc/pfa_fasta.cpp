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
#include <stdint.h>
#include <unistd.h>

#include <algorithm>
#include <chrono>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <thread>
#include <unordered_map>
#include <vector>

#include "pfa_host.h"

namespace {
unsigned big_threads() { return std::min<unsigned>(16, std::max(1u, std::thread::hardware_concurrency())); }

// fn(t, nthreads) on nthreads threads
template <typename F>
void parallel_run(unsigned nthreads, F fn) {
    if (nthreads <= 1) {
        fn(0u, 1u);
        return;
    }
    std::vector<std::thread> th;
    for (unsigned t = 0; t < nthreads; ++t) th.emplace_back(fn, t, nthreads);
    for (auto& x : th) x.join();
}

// does the buffer hold byte c?  large buffers are searched by several threads
bool pfa_parallel_memchr(const unsigned char* buf, int c, size_t len) {
    if (len < (64u << 20)) return memchr(buf, c, len) != nullptr;  // small: one call
    const unsigned nt = big_threads();
    std::vector<char> hit(nt, 0);
    parallel_run(nt, [&](unsigned t, unsigned n) {
        const size_t lo = len * t / n, hi = len * (t + 1) / n;
        hit[t] = memchr(buf + lo, c, hi - lo) != nullptr;
    });
    for (char h : hit)
        if (h) return true;
    return false;
}

// str.rstrip() with no argument strips characters for which str.isspace() is true; in ASCII these are
// \t \n \v \f \r, the separators 0x1c-0x1f and the blank.
inline bool py_space(unsigned char c) { return (c >= 0x09 && c <= 0x0d) || (c >= 0x1c && c <= 0x20); }
}  // namespace

// inputs of at least this many bytes are read, scanned and compacted by several threads (PFA_BIG_FILE_MIN overrides: tests)
static size_t big_min() {
    const char* s = getenv("PFA_BIG_FILE_MIN");
    return s ? (size_t)atoll(s) : (size_t)(64u << 20);
}

// end of the line that starts at i (index of its terminator, or len) and the start of the next line
static inline void line_end(const unsigned char* buf, size_t len, size_t i, bool has_cr, size_t* e, size_t* next) {
    if (has_cr) {
        size_t x = i;
        while (x < len && buf[x] != '\n' && buf[x] != '\r') ++x;
        *e = x;
    } else {
        const void* nl = memchr(buf + i, '\n', len - i);
        *e = nl ? (size_t)(static_cast<const unsigned char*>(nl) - buf) : len;
    }
    *next = *e;
    if (*next < len) *next += (buf[*next] == '\r' && *next + 1 < len && buf[*next + 1] == '\n') ? 2 : 1;
}

// one segment of the buffer scanned on its own: the sequence lines in front of its first header (they continue the last
// record of the previous segment) and its records in file order, not yet merged by header
struct PfaSegment {
    PfaLines lead;
    std::vector<PfaRecord> recs;
};

static void scan_segment(const unsigned char* buf, size_t len, size_t lo, size_t hi, bool has_cr, PfaSegment* seg) {
    size_t i = lo;
    bool in_rec = false, keep = false;
    while (i < hi) {
        size_t e, next;
        line_end(buf, len, i, has_cr, &e, &next);
        size_t r = e;
        while (r > i && py_space(buf[r - 1])) --r;
        if (e > i && buf[i] == '>') {
            seg->recs.emplace_back();
            seg->recs.back().header.assign(reinterpret_cast<const char*>(buf + i + 1), r > i + 1 ? r - (i + 1) : 0);
            in_rec = true;
            keep = !seg->recs.back().header.empty();
        } else if (r > i) {
            if (!in_rec) seg->lead.add_line(buf + i, r - i);
            else if (keep) seg->recs.back().lines.add_line(buf + i, r - i);
        }
        i = next;
    }
}

// line scan: records with their line slices (pointing into buf), lengths, and whether all rows have one length.
// The reference's state machine (PolyFastA.py:232-245) runs over the records of the segments in file order: a header seen
// again restarts its record in place, the lines after an empty header are dropped, lines before any header are ignored.
int pfa_parse_lines(const unsigned char* buf, size_t len, PfaParsed* out) {
    std::vector<PfaRecord>& recs = out->recs;
    recs.clear();
    // a '\r' anywhere switches to the byte-wise line scan (a lone '\r' ends a line under universal newlines); otherwise
    // lines end at '\n' only and memchr finds them at memory speed
    const bool has_cr = len && pfa_parallel_memchr(buf, '\r', len);
    unsigned nseg = (len >= big_min() && !has_cr) ? big_threads() : 1;
    if (const char* s = getenv("PFA_PARSE_SEGMENTS")) nseg = has_cr ? 1u : (unsigned)std::max(1, atoi(s));
    // segment boundaries on line starts
    std::vector<size_t> start(nseg + 1, len);
    start[0] = 0;
    for (unsigned t = 1; t < nseg; ++t) {
        const size_t lo = std::max(start[t - 1], len / nseg * t);
        if (lo == 0 || lo >= len) {
            start[t] = std::min(lo, len);
            continue;
        }
        const void* nl = memchr(buf + lo - 1, '\n', len - (lo - 1));
        start[t] = nl ? (size_t)(static_cast<const unsigned char*>(nl) - buf) + 1 : len;
    }
    std::vector<PfaSegment> segs(nseg);
    parallel_run(nseg, [&](unsigned t, unsigned) { scan_segment(buf, len, start[t], start[t + 1], has_cr, &segs[t]); });
    std::unordered_map<std::string, size_t> index;
    bool seen_header = false, head_nonempty = false;
    size_t cur = 0;
    for (PfaSegment& seg : segs) {
        if (seen_header && head_nonempty && seg.lead.nlines) recs[cur].lines.append(seg.lead);
        for (PfaRecord& r : seg.recs) {
            seen_header = true;
            head_nonempty = !r.header.empty();
            auto it = index.find(r.header);
            if (it == index.end()) {
                index.emplace(r.header, recs.size());
                cur = recs.size();
                recs.push_back(std::move(r));
            } else {
                cur = it->second;
                recs[cur].lines = std::move(r.lines);
            }
        }
    }
    if (!seen_header || !head_nonempty) return PFA_ERR_NOT_FASTA;
    out->total = 0;
    out->same = true;
    for (PfaRecord& rec : recs) rec.len = rec.lines.len;
    for (const PfaRecord& rec : recs) {
        out->total += rec.len;
        if (rec.len != recs[0].len) out->same = false;
    }
    out->seqlen = out->same ? recs[0].len : -1;
    return PFA_OK;
}

// OR of all bytes of a slice, eight at a time
static inline unsigned char or_bytes(const unsigned char* p, size_t n) {
    uint64_t a0 = 0, a1 = 0, a2 = 0, a3 = 0;
    size_t i = 0;
    for (; i + 32 <= n; i += 32) {
        uint64_t w0, w1, w2, w3;
        memcpy(&w0, p + i, 8);
        memcpy(&w1, p + i + 8, 8);
        memcpy(&w2, p + i + 16, 8);
        memcpy(&w3, p + i + 24, 8);
        a0 |= w0; a1 |= w1; a2 |= w2; a3 |= w3;
    }
    uint64_t a = a0 | a1 | a2 | a3;
    unsigned char r = 0;
    for (; i < n; ++i) r |= p[i];
    a |= a >> 32;
    a |= a >> 16;
    a |= a >> 8;
    return (unsigned char)(r | (unsigned char)a);
}

// copy the line slices of one record to dst (memmove: dst may be the start of the record's own span in the same buffer);
// returns false when a byte >= 0x80 was seen
bool pfa_copy_record(const PfaRecord& rec, unsigned char* dst) {
    unsigned char acc = 0;
    for (const PfaRun& r : rec.lines.runs) {
        // the bytes between the lines of a run are line ends and stripped blanks, all below 0x80: one pass over the span
        acc |= or_bytes(r.p, (size_t)(r.end() - r.p));
        for (int64_t i = 0; i < r.count; ++i) {
            const unsigned char* src = r.p + i * r.stride;
            if (dst != src) memmove(dst, src, r.w);
            dst += r.w;
        }
    }
    return !(acc & 0x80);
}

static double now_ms() {
    return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

static int parse_impl(const unsigned char* buf, size_t len, pfa_fasta** out) {
    PfaParsed parsed;
    const double t0 = now_ms();
    int rc = pfa_parse_lines(buf, len, &parsed);
    if (rc) return rc;
    const double t1 = now_ms();
    struct Trace {
        double t0, t1;
        size_t len;
        ~Trace() {
            if (getenv("PFA_PARSE_TRACE"))
                fprintf(stderr, "[pfa parse] %zu bytes: line scan %.1f ms, rows -> matrix %.1f ms\n", len, t1 - t0, now_ms() - t1);
        }
    } trace{t0, t1, len};
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
#ifdef MADV_HUGEPAGE
    if (f->data_bytes >= (64u << 20)) madvise(f->data, f->data_bytes, MADV_HUGEPAGE);
#endif
    // rows are independent, so large files are copied by several threads
    unsigned nthreads = total > (64ll << 20) ? big_threads() : 1;
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

// Large files.  The file is mapped privately and NEVER written: rows that sit on one line are used where they lie, rows
// wrapped over lines of one width are described by (first byte, line width, gap) and gathered by whoever reads them (the
// ingest's packer, pfa_fasta_copy_row) -- no copy of the file at all.  Files whose wrapping is irregular (blank lines,
// lines of different widths inside a record), files that cannot be mapped and PFA_PARSE_MMAP=0 take the read path: a
// malloc'ed buffer filled by several threads with pread, every record compacted in place.
static bool regular_wrap(const PfaRecord& rec, int32_t* w, int32_t* gap) {
    // every line but the last has one width W, the first bytes of consecutive lines are one stride apart
    const std::vector<PfaRun>& runs = rec.lines.runs;
    if (rec.lines.nlines < 2) return false;
    const int64_t W = (int64_t)runs[0].w;
    const int64_t stride = runs[0].count > 1 ? runs[0].stride : (int64_t)(runs[1].p - runs[0].p);
    if (W <= 0 || stride <= W || W > 0x7fffffff || stride - W > 0x7fffffff) return false;
    int64_t before = 0;
    for (size_t i = 0; i < runs.size(); ++i) {
        const PfaRun& r = runs[i];
        if (r.p != runs[0].p + before * stride) return false;
        if (r.count > 1 && r.stride != stride) return false;
        const bool last_line = i + 1 == runs.size() && r.count == 1;
        if ((int64_t)r.w != W && !(last_line && (int64_t)r.w < W && r.w > 0)) return false;
        before += r.count;
    }
    *w = (int32_t)W;
    *gap = (int32_t)(stride - W);
    return true;
}

static int parse_big_file(const char* path, size_t size, pfa_fasta** out) {
    const unsigned nt = big_threads();
    const char* mm = getenv("PFA_PARSE_MMAP");
    for (int attempt = (!mm || atoi(mm) != 0) ? 0 : 1; attempt < 2; ++attempt) {
        const bool want_map = attempt == 0;
        int fd = open(path, O_RDONLY);
        if (fd < 0) return PFA_ERR_IO;
        unsigned char* buf = nullptr;
        size_t mapped = 0;
        if (want_map) {
            void* m = mmap(nullptr, size, PROT_READ | PROT_WRITE, MAP_PRIVATE, fd, 0);
            if (m == MAP_FAILED) {
                close(fd);
                continue;
            }
            buf = static_cast<unsigned char*>(m);
            mapped = size;
            madvise(buf, size, MADV_WILLNEED);
        } else {
            buf = static_cast<unsigned char*>(malloc(size));
            if (!buf) {
                close(fd);
                return PFA_ERR_NOMEM;
            }
#ifdef MADV_HUGEPAGE
            madvise(buf, size, MADV_HUGEPAGE);
#endif
            std::vector<char> bad(nt, 0);
            parallel_run(nt, [&](unsigned t, unsigned n) {
                size_t lo = size * t / n;
                const size_t hi = size * (t + 1) / n;
                while (lo < hi) {
                    const ssize_t r = pread(fd, buf + lo, hi - lo, (off_t)lo);
                    if (r <= 0) {
                        bad[t] = 1;
                        return;
                    }
                    lo += (size_t)r;
                }
            });
            for (char b : bad)
                if (b) {
                    close(fd);
                    free(buf);
                    return PFA_ERR_IO;
                }
        }
        close(fd);
        auto release = [&]() {
            if (mapped) munmap(buf, mapped);
            else free(buf);
        };
        const double t0 = now_ms();
        PfaParsed parsed;
        int rc = pfa_parse_lines(buf, size, &parsed);
        if (rc) {
            release();
            return rc;
        }
        const double t1 = now_ms();
        const std::vector<PfaRecord>& recs = parsed.recs;
        // mapped: wrapped rows must be regular, or the file goes through the read path
        std::vector<int32_t> ww, wg;
        if (mapped) {
            bool any_wrapped = false, irregular = false;
            for (const PfaRecord& r : recs) any_wrapped |= r.lines.nlines > 1;
            if (any_wrapped) {
                ww.assign(recs.size(), 0);
                wg.assign(recs.size(), 0);
                for (size_t k = 0; k < recs.size() && !irregular; ++k)
                    if (recs[k].lines.nlines > 1 && !regular_wrap(recs[k], &ww[k], &wg[k])) irregular = true;
            }
            if (irregular) {
                release();
                continue;
            }
        }
        pfa_fasta* f = new pfa_fasta();
        f->n = (int64_t)recs.size();
        f->in_place = true;
        f->data = buf;
        f->data_bytes = size;
        f->mapped_bytes = mapped;
        f->seqlen = parsed.seqlen;
        f->wrap_w = std::move(ww);
        f->wrap_gap = std::move(wg);
        f->row_len.resize(recs.size());
        f->row_off.resize(recs.size() + 1);
        for (size_t k = 0; k < recs.size(); ++k) {
            f->row_len[k] = recs[k].len;
            f->row_off[k] = recs[k].lines.runs.empty() ? 0 : (int64_t)(recs[k].lines.runs[0].p - buf);
            f->header_off.push_back((int64_t)f->headers.size());
            f->headers += recs[k].header;
        }
        f->row_off[recs.size()] = (int64_t)size;
        f->header_off.push_back((int64_t)f->headers.size());
        // bytes >= 0x80 in sequence lines are refused; read path: the same pass compacts every record where its lines were
        // (the lines of a record lie between its header and the next one: disjoint spans, rows move independently)
        std::vector<char> non_ascii(nt, 0);
        parallel_run(nt, [&](unsigned t, unsigned n) {
            for (size_t k = t; k < recs.size(); k += n) {
                if (recs[k].lines.runs.empty()) continue;
                if (mapped) {
                    // one pass over the record's whole span: the bytes between its lines are line ends and stripped blanks,
                    // all below 0x80, so they cannot hide or fake a high bit
                    const unsigned char* first = recs[k].lines.runs.front().p;
                    if (or_bytes(first, (size_t)(recs[k].lines.runs.back().end() - first)) & 0x80) non_ascii[t] = 1;
                } else if (!pfa_copy_record(recs[k], buf + f->row_off[k])) {
                    non_ascii[t] = 1;
                }
            }
        });
        if (getenv("PFA_PARSE_TRACE"))
            fprintf(stderr, "[pfa parse] %zu bytes in place (%s%s): line scan %.1f ms, %s %.1f ms\n", size, mapped ? "mapped" : "read",
                    f->wrap_w.empty() ? "" : ", wrapped rows gathered on access", t1 - t0, mapped ? "byte check" : "compaction", now_ms() - t1);
        for (char b : non_ascii)
            if (b) {
                pfa_fasta_free(f);
                return PFA_ERR_NON_ASCII;
            }
        *out = f;
        return PFA_OK;
    }
    return PFA_ERR_IO;
}

// ---- layout of a file parsed in place, for the other ranks of a column-sharded run ----------------------------------------------
// Every rank of a multi-GPU run needs the rows of the same large file, but only ONE has to find them: the line scan (and the
// byte check) of a multi-GB file reads it once from the page cache, and eight ranks doing it at the same time fight for the
// host's memory bandwidth, which the uploads need.  The rank that parsed exports where the rows are -- a few bytes per row --
// and the others map the file and adopt that layout without reading a byte of it.
static const uint64_t LAYOUT_MAGIC = 0x3159414c41465024ull;  // "$PFALAY1"

int64_t pfa_fasta_layout_bytes(const pfa_fasta* f) {
    if (!f || !f->in_place || !f->mapped_bytes) return 0;  // only files mapped in place can be adopted without a copy
    const size_t n = (size_t)f->n;
    return (int64_t)(8 * 6 + 8 * (n + 1) + 8 * n + 8 * (n + 1) + f->headers.size() + (f->wrap_w.empty() ? 0 : 8 * n));
}

int pfa_fasta_export_layout(const pfa_fasta* f, void* buf, int64_t cap) {
    const int64_t need = pfa_fasta_layout_bytes(f);
    if (!need || !buf || cap < need) return PFA_ERR_ARG;
    unsigned char* p = static_cast<unsigned char*>(buf);
    const size_t n = (size_t)f->n;
    const uint64_t head[6] = {LAYOUT_MAGIC, (uint64_t)f->n, (uint64_t)f->seqlen, (uint64_t)f->data_bytes, (uint64_t)f->headers.size(),
                              (uint64_t)(f->wrap_w.empty() ? 0 : 1)};
    auto put = [&](const void* src, size_t bytes) {
        memcpy(p, src, bytes);
        p += bytes;
    };
    put(head, sizeof head);
    put(f->row_off.data(), 8 * (n + 1));
    put(f->row_len.data(), 8 * n);
    put(f->header_off.data(), 8 * (n + 1));
    put(f->headers.data(), f->headers.size());
    if (!f->wrap_w.empty()) {
        put(f->wrap_w.data(), 4 * n);
        put(f->wrap_gap.data(), 4 * n);
    }
    return PFA_OK;
}

int pfa_fasta_import_layout(const char* path, const void* buf, int64_t bytes, pfa_fasta** out) {
    if (!path || !buf || !out || bytes < 48) return PFA_ERR_ARG;
    *out = nullptr;
    const unsigned char* p = static_cast<const unsigned char*>(buf);
    uint64_t head[6];
    memcpy(head, p, sizeof head);
    p += sizeof head;
    if (head[0] != LAYOUT_MAGIC) return PFA_ERR_ARG;
    const size_t n = (size_t)head[1], hbytes = (size_t)head[4];
    const bool wrapped = head[5] != 0;
    if ((uint64_t)bytes != 48 + 8 * (n + 1) + 8 * n + 8 * (n + 1) + hbytes + (wrapped ? 8 * n : 0)) return PFA_ERR_ARG;
    struct stat st;
    if (stat(path, &st) != 0 || (uint64_t)st.st_size != head[3]) return PFA_ERR_IO;  // not the file the layout describes
    int fd = open(path, O_RDONLY);
    if (fd < 0) return PFA_ERR_IO;
    void* m = mmap(nullptr, (size_t)head[3], PROT_READ | PROT_WRITE, MAP_PRIVATE, fd, 0);
    close(fd);
    if (m == MAP_FAILED) return PFA_ERR_IO;
    pfa_fasta* f = new pfa_fasta();
    f->n = (int64_t)n;
    f->seqlen = (int64_t)head[2];
    f->in_place = true;
    f->data = static_cast<unsigned char*>(m);
    f->data_bytes = f->mapped_bytes = (size_t)head[3];
    auto get = [&](void* dst, size_t b) {
        memcpy(dst, p, b);
        p += b;
    };
    f->row_off.resize(n + 1);
    f->row_len.resize(n);
    f->header_off.resize(n + 1);
    f->headers.resize(hbytes);
    get(f->row_off.data(), 8 * (n + 1));
    get(f->row_len.data(), 8 * n);
    get(f->header_off.data(), 8 * (n + 1));
    get(&f->headers[0], hbytes);
    if (wrapped) {
        f->wrap_w.resize(n);
        f->wrap_gap.resize(n);
        get(f->wrap_w.data(), 4 * n);
        get(f->wrap_gap.data(), 4 * n);
    }
    *out = f;
    return PFA_OK;
}

int pfa_fasta_parse_file(const char* path, pfa_fasta** out) {
    if (!out || !path) return PFA_ERR_ARG;
    *out = nullptr;
    struct stat st;
    if (stat(path, &st) == 0 && S_ISREG(st.st_mode) && (size_t)st.st_size >= big_min() && st.st_size > 0) return parse_big_file(path, (size_t)st.st_size, out);
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
    if (f->mapped_bytes) munmap(f->data, f->mapped_bytes);
    else free(f->data);
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
    const int64_t len = f->row_len[(size_t)row];
    if (!f->wrap_w.empty() && f->wrap_w[(size_t)row] > 0) {
        pfa_gather_wrapped(src, f->wrap_w[(size_t)row], f->wrap_gap[(size_t)row], 0, len, dst);
        src = dst;
    }
    for (int64_t i = 0; i < len; ++i) {
        unsigned char c = src[i];
        dst[i] = (c >= 'a' && c <= 'z') ? (unsigned char)(c - 32) : c;  // str.upper() on ASCII
    }
    return PFA_OK;
}

}  // extern "C"
