// Builds the codon tables of K4 once per process.  See pfa_codon_rules.h.
#include "pfa_codon_rules.h"

#include <algorithm>
#include <map>
#include <set>
#include <vector>

#include "../../include/polyfasta_b200.h"

namespace {

const char kBases[] = "ACGT";
// standard genetic code, TCAG order
const char kAminoTCAG[] = "FFLLSSSSYY**CC*WLLLLPPPPHHQQRRRRIIIMTTTTNNKKSSRRVVVVAAAADDEEGGGG";

char amino(int codon) {
    static const int to_tcag[4] = {2, 1, 3, 0};  // A C G T -> position in TCAG
    return kAminoTCAG[16 * to_tcag[codon >> 4] + 4 * to_tcag[(codon >> 2) & 3] + to_tcag[codon & 3]];
}

std::string codon_str(int c) { return {kBases[c >> 4], kBases[(c >> 2) & 3], kBases[c & 3]}; }

std::string family(int c) {
    const char aa = amino(c);
    int block = 0;
    for (int x = 0; x < 4; ++x) block += amino((c & ~3) | x) == aa;
    const bool pyrimidine = (c & 3) == 1 || (c & 3) == 3;
    const char* tail = block == 4 ? "4N" : block == 3 ? "3H" : block == 2 ? (pyrimidine ? "2Y" : "2R") : "0G";
    return std::string(1, aa) + tail;
}

using Kinds = std::set<std::string>;
bool subset(const Kinds& k, std::initializer_list<const char*> of) {
    for (const auto& x : k)
        if (std::find_if(of.begin(), of.end(), [&](const char* o) { return x == o; }) == of.end()) return false;
    return true;
}
bool has(const Kinds& k, const char* x) { return k.count(x) != 0; }
bool has_tail(const Kinds& k, const char* t) {
    for (const auto& x : k)
        if (x.substr(1) == t) return true;
    return false;
}
bool all_tail(const Kinds& k, const char* t) {
    for (const auto& x : k)
        if (x.substr(1) != t) return false;
    return true;
}

// two distinct sense codons (PolyFastA.py:347-414)
int pair_rule(int a, int b, const std::string& ka, const std::string& kb) {
    int lab[3] = {0, 0, 0};
    std::vector<int> diff;
    for (int i = 0; i < 3; ++i)
        if (((a >> (4 - 2 * i)) & 3) != ((b >> (4 - 2 * i)) & 3)) diff.push_back(i);
    const Kinds kinds = {ka, kb};
    const std::string sa = codon_str(a), sb = codon_str(b);
    auto either = [&](std::initializer_list<const char*> cods) {
        for (const char* c : cods)
            if (sa == c || sb == c) return true;
        return false;
    };
    auto both_in = [&](std::initializer_list<const char*> cods) {
        auto in = [&](const std::string& s) { return std::find_if(cods.begin(), cods.end(), [&](const char* c) { return s == c; }) != cods.end(); };
        return in(sa) && in(sb);
    };
    if (diff.size() == 1) {
        const int i = diff[0];
        const bool syn = ka == kb || (i == 0 && (subset(kinds, {"L4N", "L2R"}) || subset(kinds, {"R4N", "R2R"})));  // :351-355
        lab[i] = syn ? 1 : 2;
    } else {
        if (diff.back() == 2) {  // :363-396
            const bool syn =
                has_tail(kinds, "4N") || all_tail(kinds, "2Y") || all_tail(kinds, "2R") ||
                (has(kinds, "I3H") && (has_tail(kinds, "2Y") ||
                                       (has_tail(kinds, "2R") && !both_in({"AAG", "AGG", "GAG", "TTG", "ATT", "ATC"})))) ||
                (has(kinds, "L2R") && has(kinds, "W0G")) ||
                (has(kinds, "M0G") && (has(kinds, "R2R") || has(kinds, "K2R") || has(kinds, "L2R"))) ||
                subset(kinds, {"R2R", "H2Y"}) || subset(kinds, {"L2R", "H2Y"}) || subset(kinds, {"D2Y", "R2R"}) ||
                subset(kinds, {"F2Y", "Q2R"}) || subset(kinds, {"C2Y", "Q2R"}) || subset(kinds, {"E2R", "S2Y"}) ||
                subset(kinds, {"S2Y", "Q2R"});
            lab[2] = syn ? 1 : 2;
        }
        if (diff.front() == 0) {  // :398-411 (the "T0G" test of :405 can never match)
            const bool from = has(kinds, "L2R") || has(kinds, "R2R");
            const bool to = has(kinds, "P4N") || has(kinds, "L4N") || has(kinds, "R4N") || has(kinds, "H2Y") || has(kinds, "Q2R");
            const bool syn = (from && to) || (has(kinds, "L4N") && either({"TCA", "TCG"})) ||
                             (has(kinds, "R4N") && either({"ATG", "ATA", "ACA", "ACG", "AAA", "AAG"}));
            lab[0] = syn ? 1 : 2;
        }
        if (diff[0] == 1 || diff[1] == 1) lab[1] = 2;  // :413-414
    }
    return lab[0] | (lab[1] << 2) | (lab[2] << 4);
}

PfaCodonTables build() {
    PfaCodonTables t{};
    std::map<std::string, int> ids;
    std::string kind[64];
    t.stop_mask = 0;
    for (int c = 0; c < 64; ++c) {
        if (amino(c) == '*') {
            t.stop_mask |= 1ull << c;
            t.cls[c] = 255;
            t.syn3[c] = 0;
            continue;
        }
        kind[c] = family(c);
        ids.emplace(kind[c], 0);
        int syn = 0;
        for (int i = 0; i < 3; ++i)
            for (int x = 0; x < 4; ++x) {
                const int sh = 4 - 2 * i;
                const int nb = (c & ~(3 << sh)) | (x << sh);
                if (nb != c && amino(nb) == amino(c)) ++syn;
            }
        t.syn3[c] = (uint8_t)syn;
    }
    int next = 0;
    for (auto& kv : ids) {
        kv.second = next;
        t.cls_name[next] = kv.first;
        t.class_mask[next] = 0;
        ++next;
    }
    t.num_classes = next;
    for (int c = 0; c < 64; ++c)
        if (t.cls[c] != 255) {
            t.cls[c] = (uint8_t)ids[kind[c]];
            t.class_mask[t.cls[c]] |= 1ull << c;
        }
    for (int a = 0; a < 64; ++a)
        for (int b = 0; b < 64; ++b)
            t.pair[a][b] = (a == b || t.cls[a] == 255 || t.cls[b] == 255) ? 0 : (uint8_t)pair_rule(a, b, kind[a], kind[b]);
    return t;
}

}  // namespace

const PfaCodonTables& pfa_codon_tables() {
    static const PfaCodonTables t = build();
    return t;
}

// three or more distinct sense codons (PolyFastA.py:415-432)
int pfa_multi_labels_host(uint64_t g) {
    const PfaCodonTables& t = pfa_codon_tables();
    int nb[3] = {0, 0, 0};
    for (int i = 0; i < 3; ++i) {
        unsigned seen = 0;
        for (int c = 0; c < 64; ++c)
            if (g >> c & 1) seen |= 1u << ((c >> (4 - 2 * i)) & 3);
        nb[i] = __builtin_popcount(seen);
    }
    int top = 0;
    for (int k = 0; k < t.num_classes; ++k) top = std::max(top, __builtin_popcountll(g & t.class_mask[k]));
    int last = -1;
    for (int i = 0; i < 3; ++i)
        if (nb[i] > 1) last = i;
    int lab[3] = {0, 0, 0};
    if (top >= 2) {
        for (int i = 0; i < last; ++i)
            if (nb[i] > 1) lab[i] = 2;
        if (top >= nb[last]) lab[last] = 1;
    } else {
        for (int i = 0; i < 3; ++i)
            if (nb[i] > 1) lab[i] = 2;
    }
    return lab[0] | (lab[1] << 2) | (lab[2] << 4);
}

extern "C" {

int pfa_codon_syn3(int codon) { return (codon < 0 || codon > 63) ? -1 : pfa_codon_tables().syn3[codon]; }
int pfa_codon_class(int codon) { return (codon < 0 || codon > 63) ? -1 : pfa_codon_tables().cls[codon]; }
int pfa_codon_pair_labels(int a, int b) {
    if (a < 0 || a > 63 || b < 0 || b > 63) return -1;
    return pfa_codon_tables().pair[a][b];
}
int pfa_codon_set_labels(uint64_t sense_codon_set) {
    const PfaCodonTables& t = pfa_codon_tables();
    const uint64_t g = sense_codon_set & ~t.stop_mask;
    const int n = __builtin_popcountll(g);
    if (n < 2) return 0;
    if (n == 2) return t.pair[__builtin_ctzll(g)][63 - __builtin_clzll(g)];
    return pfa_multi_labels_host(g);
}

}  // extern "C"
