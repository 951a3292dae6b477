// The reference's synonymous / nonsynonymous classifier as tables (host side of K4).
//
// Restates get_syn_nonsyn_cod_sites (PolyFastA.py:319-434) and syncodfreq (:536-557).  Codon index =
// 16*b1 + 4*b2 + b3 with A=0 C=1 G=2 T=3.  Every codon gets the 3-character class of :324-329 -- amino acid,
// size of its synonymous block among the four codons sharing the first two bases, IUPAC letter of the block's
// third bases (4N, 3H, 2Y, 2R, 0G) -- derived here from the standard genetic code instead of being typed in.
// Labels: 0 none, 1 synonymous, 2 nonsynonymous; a label byte packs position 0 in bits 0-1, 1 in 2-3, 2 in 4-5.
#pragma once
#include <stdint.h>

#include <string>

struct PfaCodonTables {
    uint8_t syn3[64];        // 3*syncodfreq: number of synonymous single-base neighbours, 0 for stops
    uint8_t cls[64];         // class id 0..22, 255 for stop codons
    uint8_t pair[64][64];    // label byte of two distinct sense codons (0 when either is a stop or a == b)
    uint64_t class_mask[24]; // codons of each class id
    uint64_t stop_mask;
    int num_classes;
    std::string cls_name[24];
};

const PfaCodonTables& pfa_codon_tables();
int pfa_multi_labels_host(uint64_t sense_set);
