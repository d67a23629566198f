/*
 * oracle/match_oracle.c -- CPU restatement of OpenVSLAM's Hamming matchers.
 * TEST INFRASTRUCTURE ONLY (see orb_oracle.c).  PARITY STATUS: parity unpinned against the
 * real reference (no source under /root/reference); the restatement
 * follows match/base.h and match/robust.cc as recalled in SURVEY.md 8(a) a8, a11.
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include "match_oracle.h"

/* match::compute_descriptor_distance_32 (match/base.h): 8 x 32-bit SWAR popcount. */
unsigned om_hamming(const uint8_t* a, const uint8_t* b) {
    const uint32_t* pa = (const uint32_t*)a;
    const uint32_t* pb = (const uint32_t*)b;
    unsigned dist = 0;
    for (int i = 0; i < 8; ++i, ++pa, ++pb) {
        uint32_t v = *pa ^ *pb;
        v = v - ((v >> 1) & 0x55555555u);
        v = (v & 0x33333333u) + ((v >> 2) & 0x33333333u);
        dist += (((v + (v >> 4)) & 0xF0F0F0Fu) * 0x1010101u) >> 24;
    }
    return dist;
}

/* The sequential nearest / second-nearest scan used by every matcher:
 *   if (d < best) { second = best; best = d; best_idx = i; } else if (d < second) second = d; */
void om_bruteforce(const uint8_t* desc1, int n1, const uint8_t* desc2, int n2,
                   int32_t* best_idx, int32_t* best_dist, int32_t* second_dist) {
    for (int q = 0; q < n1; ++q) {
        unsigned best = OM_MAX_HAMMING_DIST, second = OM_MAX_HAMMING_DIST;
        int bi = -1;
        for (int t = 0; t < n2; ++t) {
            const unsigned d = om_hamming(desc1 + (size_t)q * 32, desc2 + (size_t)t * 32);
            if (d < best) { second = best; best = d; bi = t; }
            else if (d < second) second = d;
        }
        best_idx[q] = bi; best_dist[q] = (int32_t)best; second_dist[q] = (int32_t)second;
    }
}

/* match::robust::brute_force_match (match/robust.cc). */
int om_robust_brute_force_match(const uint8_t* desc_frm, int n1, const uint8_t* desc_keyfrm, int n2,
                                const uint8_t* lm_valid_2, float lowe_ratio, int32_t* pairs_out) {
    uint8_t* already = (uint8_t*)calloc((size_t)n1 + 1, 1);
    int num = 0;
    for (int idx_2 = 0; idx_2 < n2; ++idx_2) {
        if (lm_valid_2 && !lm_valid_2[idx_2]) continue;
        const uint8_t* d2 = desc_keyfrm + (size_t)idx_2 * 32;
        unsigned best = OM_MAX_HAMMING_DIST, second = OM_MAX_HAMMING_DIST;
        int best_idx_1 = -1;
        for (int idx_1 = 0; idx_1 < n1; ++idx_1) {
            if (already[idx_1]) continue;
            const unsigned d = om_hamming(d2, desc_frm + (size_t)idx_1 * 32);
            if (d < best) { second = best; best = d; best_idx_1 = idx_1; }
            else if (d < second) second = d;
        }
        if (OM_HAMMING_DIST_THR_LOW < best) continue;
        if (lowe_ratio * second < (float)best) continue;
        pairs_out[2 * num] = best_idx_1; pairs_out[2 * num + 1] = idx_2;
        already[best_idx_1] = 1;
        ++num;
    }
    free(already);
    return num;
}
