// Host-only part of the library (no CUDA headers): the parsed FASTA file.
#pragma once
#include <stdint.h>
#include <string.h>

#include <algorithm>

#include <string>
#include <vector>

#include "../../include/polyfasta_b200.h"

struct pfa_fasta {
    int64_t n = 0;
    int64_t seqlen = -1;             // common row length or -1
    unsigned char* data = nullptr;   // rows back to back in first-seen order; a matrix [n][seqlen] when seqlen >= 0
    size_t data_bytes = 0;
    std::vector<int64_t> row_off;    // n+1 (in_place: n offsets of the rows inside the file buffer, any stride)
    bool in_place = false;           // large files: `data` is the file buffer itself, every row compacted where its lines were
    std::vector<int64_t> row_len;    // n
    std::string headers;             // concatenated header bytes
    std::vector<int64_t> header_off; // n+1
    bool pinned = false;             // data registered with cudaHostRegister by the uploader
    size_t mapped_bytes = 0;         // > 0: `data` is a private file mapping of this size (munmap), not malloc memory
    // rows of a mapped file that are wrapped over lines stay as they are in the file (no write, no copy): row r is lines of
    // wrap_w[r] bytes every wrap_w[r] + wrap_gap[r] bytes from row_off[r]; wrap_w[r] == 0: the row is contiguous.  Empty
    // vectors: every row is contiguous.
    std::vector<int32_t> wrap_w, wrap_gap;
    void (*unpin)(void*) = nullptr;
};

// ---- parser internals shared with the batched path --------------------------------------------------------------------
// The sequence lines of a record as RUNS: `count` lines of `w` bytes whose first bytes are `stride` apart.  A sequence
// on one line is one run of one line; a sequence wrapped at a fixed width is one run (plus one for a shorter last line),
// whatever its length -- 10^8 lines of a 6 GB file take a few dozen bytes, not 16 bytes each.
struct PfaRun {
    const unsigned char* p;
    size_t w;
    int64_t stride;
    int64_t count;
    const unsigned char* end() const { return p + (count - 1) * stride + w; }  // one past the last byte of the last line
};
struct PfaLines {
    std::vector<PfaRun> runs;
    int64_t len = 0;     // bytes of all lines
    int64_t nlines = 0;
    void add_line(const unsigned char* p, size_t n) {
        len += (int64_t)n;
        ++nlines;
        if (!runs.empty()) {
            PfaRun& r = runs.back();
            if (n == r.w) {
                if (r.count == 1) {
                    r.stride = (int64_t)(p - r.p);
                    r.count = 2;
                    return;
                }
                if (p == r.p + r.count * r.stride) {
                    ++r.count;
                    return;
                }
            }
        }
        runs.push_back(PfaRun{p, n, 0, 1});
    }
    void append(const PfaLines& o) {  // the lines of `o` follow these (next segment of the buffer)
        for (const PfaRun& r : o.runs) {
            add_line(r.p, r.w);
            if (r.count == 1) continue;
            const int64_t rest = r.count - 1;
            PfaRun& l = runs.back();
            const bool first_is_tail = l.w == r.w && l.p + (l.count - 1) * l.stride == r.p;
            if (first_is_tail && l.count >= 2 && l.stride == r.stride) l.count += rest;  // the run goes on
            else if (first_is_tail && l.count == 1) {                                     // it was pushed as a single line
                l.stride = r.stride;
                l.count = r.count;
            } else {
                runs.push_back(PfaRun{r.p + r.stride, r.w, r.stride, rest});
            }
            len += (int64_t)r.w * rest;
            nlines += rest;
        }
    }
    void clear() {
        runs.clear();
        len = nlines = 0;
    }
};
struct PfaRecord {
    std::string header;
    PfaLines lines;
    int64_t len = 0;  // == lines.len (kept for the callers that only need the length)
};
struct PfaParsed {
    std::vector<PfaRecord> recs;
    int64_t total = 0;
    int64_t seqlen = -1;
    bool same = true;
};
int pfa_parse_lines(const unsigned char* buf, size_t len, PfaParsed* out);
bool pfa_copy_record(const PfaRecord& rec, unsigned char* dst);
int pfa_read_file(const char* path, std::vector<unsigned char>* buf, size_t* len);

// ---- host packer (pfa_pack.cpp): cols bases of one text row -> ceil(cols/4) bytes, 4 bases per byte, code (byte >> 1) & 3
// (A 0, C 1, T 2, G 3, either case); returns 1 when the row holds any other byte (the chunk then travels as text)
int pfa_pack2_row(const uint8_t* src, int64_t cols, uint8_t* dst);
// the same with a validity bitmap (codes A0 C1 G2 T3 / '-'0 'N'1 '?'2, one validity bit per base): bit 0 of the result = the
// row holds '-', 'N' or '?', bit 1 = it holds any other byte (dirty).  pfa_pack3_fast(): the CPU has AVX-512 VBMI (one
// table-lookup instruction per 64 bases); without it the scalar version is slower than shipping the text.
int pfa_pack3_row(const uint8_t* src, int64_t cols, uint8_t* codes, uint8_t* valid);
bool pfa_pack3_fast();

// columns [c0, c0 + cols) of a row stored as lines of w bytes every (w + gap) bytes from p, copied to dst
static inline void pfa_gather_wrapped(const uint8_t* p, int32_t w, int32_t gap, int64_t c0, int64_t cols, uint8_t* dst) {
    int64_t o = c0 % w;
    const uint8_t* src = p + (c0 / w) * (int64_t)(w + gap) + o;
    while (cols > 0) {
        const int64_t take = std::min<int64_t>(w - o, cols);
        memcpy(dst, src, (size_t)take);
        dst += take;
        cols -= take;
        src += take + gap;  // only used again when `take` reached the end of the line
        o = 0;
    }
}
