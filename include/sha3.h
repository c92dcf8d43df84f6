/*
 * sha3.h -- reference-signature SHA-3 front-end of the B200-native engine (libmlkem_b200.so).
 *
 * Same types and prototypes as the reference's sha3.h (file:line given per item); the Keccak permutations run
 * on the GPU (mlkem_b200_sha3_bits_batch), the conversions between bit / hex / character strings are host-side
 * layout code.  Cells are the reference's 4-byte unions (one bit or one hex digit per cell); every returned
 * array is malloc'ed by the callee and freed by the caller.
 */
#ifndef SHA3_H
#define SHA3_H

#include <stdlib.h>

#ifdef __cplusplus
extern "C" {
#endif
#if defined(__GNUC__)
#pragma GCC visibility push(default)
#endif

union bit { /* reference sha3.h:15-17 */
    unsigned int b : 1;
};
union hex { /* reference sha3.h:20-22 */
    unsigned int d : 4;
};

/* hex string (2m digits) -> SHA-3 bit string truncated to n bits, FIPS 202 B.1; reference sha3.h:28, sha3.c:329 */
union bit *h2b(const union hex *H, unsigned int m, unsigned int n);
/* SHA-3 bit string of n bits -> 2 ceil(n/8) hex digits; reference sha3.h:34, sha3.c:367 */
union hex *b2h(const union bit *S, unsigned int n);
/* n message bits -> d output bits; c = capacity; sfx {0,1,..} for hashes (c = 2d), {1,1,1,1} for XOFs;
 * reference sha3.h:42, sha3.c:408 */
union bit *sha3_b(const union bit *bstr, unsigned int n, unsigned int d, unsigned int c, union bit sfx[4]);
/* 2m hex digits -> 2 (d/8) hex digits; reference sha3.h:52, sha3.c:443 */
union hex *sha3_h(const union hex *hstr, unsigned int m, unsigned int d, unsigned int c, union bit sfx[4]);
/* m characters -> d/8 characters; reference sha3.h:62, sha3.c:465 */
unsigned char *sha3_s(const char *cstr, unsigned int m, unsigned int d, unsigned int c, union bit sfx[4]);

#if defined(__GNUC__)
#pragma GCC visibility pop
#endif
#ifdef __cplusplus
}
#endif
#endif /* SHA3_H */
