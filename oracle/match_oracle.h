/* oracle/match_oracle.h -- CPU oracle for the Hamming matchers (test infrastructure only). */
#ifndef MATCH_ORACLE_H
#define MATCH_ORACLE_H
#include <stdint.h>
#define OM_HAMMING_DIST_THR_LOW 50
#define OM_HAMMING_DIST_THR_HIGH 100
#define OM_MAX_HAMMING_DIST 256
unsigned om_hamming(const uint8_t* a, const uint8_t* b);
void om_bruteforce(const uint8_t* desc1, int n1, const uint8_t* desc2, int n2,
                   int32_t* best_idx, int32_t* best_dist, int32_t* second_dist);
int om_robust_brute_force_match(const uint8_t* desc_frm, int n1, const uint8_t* desc_keyfrm, int n2,
                                const uint8_t* lm_valid_2, float lowe_ratio, int32_t* pairs_out);
#endif
