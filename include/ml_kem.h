/*
 * ml_kem.h -- the reference-signature API of the B200-native ML-KEM engine.
 *
 * An existing caller of rsjahnige/CRYSTALS-Kyber includes a header called ml_kem.h and links ml_kem.o;
 * it can include this header and link libmlkem_b200.so instead.  The type layouts, the enum values and
 * every prototype below match the reference (reference file:line given with each item), so call sites
 * compile unchanged and the ABI (struct sizes, by-value returns) is the same.  Each call is a batch of
 * one on the GPU; the batched entry points are in mlkem_b200.h.
 *
 * Cell types.  The reference stores ONE byte / coefficient / bit per 4-byte union of unsigned bit-fields
 * (sizeof == 4 for all three; the value sits in the low bits, the rest is unspecified).  Arrays passed to
 * or returned from these functions use that stride-4 layout.
 *
 * Ownership.  Every pointer returned here (PKE.ek, PKE.dk, KEM.c, the result of KEM_Decaps and of all the
 * functions returning union byte* / union integer* / union bit*) is malloc'ed by the callee; the caller
 * frees it.  Inputs are borrowed.
 *
 * Errors.  The public wrappers set the global ml_errno (never reset by the library) and print the
 * reference's message to stderr: -1 invalid parameter set, -2 random-bit generation failed, -3 length
 * ("type") check failed, -5 decapsulation-key hash check failed.  -4 (modulus check) exists in the
 * reference but cannot occur there (its ByteDecode12 never reduces) and does not occur here either.
 * One addition: -10 when no CUDA device is usable (there is no CPU fallback).
 */
#ifndef ML_KEM_H
#define ML_KEM_H

#include "sha3.h" /* union bit and the SHA-3 front-end, as the reference's ml_kem.h:10 does */
#include <stdio.h>
#include <stdlib.h>

#ifdef __cplusplus
extern "C" {
#endif
#if defined(__GNUC__)
#pragma GCC visibility push(default) /* the library itself is built with -fvisibility=hidden */
#endif

#define N 256  /* reference ml_kem.h:22 */
#define Q 3329 /* reference ml_kem.h:23 */

extern int ml_errno; /* reference ml_kem.h:26, ml_kem.c:16 */

union byte { /* reference ml_kem.h:35-38 */
    unsigned int s : 7; /* 7-bit index (BitRev7) */
    unsigned int e : 8; /* the byte value */
};
union integer { /* reference ml_kem.c:20-23 (private there; public here because the functions below use it) */
    unsigned int t : 12; /* coefficient */
    unsigned int l : 24; /* scratch width used by the reference for products */
};

struct PARAMS { /* reference ml_kem.h:42-47 */
    union byte k;
    union byte n1;
    union byte n2;
    union byte du, dv;
};
struct PKE { /* reference ml_kem.h:49-54 */
    union byte *ek;
    union byte *dk;
    unsigned int ek_len;
    unsigned int dk_len;
};
struct KEM { /* reference ml_kem.h:56-60 */
    union byte K[32];
    union byte *c;
    unsigned int c_len;
};
enum ML_KEM { ML_KEM_512 = 512, ML_KEM_768 = 768, ML_KEM_1024 = 1024 }; /* reference ml_kem.h:87-91 */

/* ---- declared in the reference header ------------------------------------------------------------- */
const struct PARAMS init(enum ML_KEM param_set);                                     /* ml_kem.h:94, ml_kem.c:1363 */
struct PKE KEM_KeyGen(const struct PARAMS *params);                                   /* ml_kem.h:68, ml_kem.c:1233 */
struct KEM KEM_Encaps(const struct PARAMS *params, const union byte *ek, unsigned int ek_len); /* ml_kem.h:76, ml_kem.c:1257 */
union byte *KEM_Decaps(const struct PARAMS *params, const union byte *dk, unsigned int dk_len, const union byte *c,
                       unsigned int c_len);                                           /* ml_kem.h:83, ml_kem.c:1310 */

/* ---- extern in ml_kem.c but not declared by its header (the Test_Archive drivers link to them) ------ */
union integer *SampleNTT(union byte *B);                             /* ml_kem.c:189; may bump B[32], B[33] */
union integer *SamplePolyCBD(const union byte *B, unsigned int n);   /* ml_kem.c:253; n = eta in {2,3} */
union integer *NTT(const union integer *f);                          /* ml_kem.c:287; coefficients < q */
union integer *InverseNTT(const union integer *fh);                  /* ml_kem.c:336 */

/* ---- static in the reference's HEAD, called as externs by its archived drivers ---------------------- */
union byte BitRev7(union byte r);                                              /* ml_kem.c:26 */
union byte *BitsToBytes(const union bit *b, unsigned int l);                   /* ml_kem.c:47 */
union bit *BytesToBits(const union byte *B, unsigned int L);                   /* ml_kem.c:62 */
union integer Compress(union integer x, unsigned int d);                       /* ml_kem.c:83 */
union integer Decompress(union integer y, unsigned int d);                     /* ml_kem.c:104 */
union byte *ByteEncode(const union integer *F, unsigned int d);                /* ml_kem.c:125; d in {1,4,5,10,11,12} */
union integer *ByteDecode(const union byte *B, unsigned int d);                /* ml_kem.c:153 */
union integer *BaseCaseMultiply(union integer a0, union integer a1, union integer b0, union integer b1,
                                union integer gamma);                          /* ml_kem.c:395; gamma in .l */
union integer *MultiplyNTTs(const union integer *fh, const union integer *gh); /* ml_kem.c:415 */
struct PKE PKE_KeyGen(const struct PARAMS *params, const union byte *d);       /* ml_kem.c:651 */
union byte *PKE_Encrypt(const struct PARAMS *params, const union byte *ek, const union byte *m,
                        const union byte *r);                                  /* ml_kem.c:776 */
union byte *PKE_Decrypt(const struct PARAMS *params, const union byte *dk, const union byte *c); /* ml_kem.c:942 */
struct PKE KeyGen_internal(const struct PARAMS *params, const union byte *d, const union byte *z); /* ml_kem.c:1034 */
struct KEM Encaps_internal(const struct PARAMS *params, const union byte *ek, const union byte *m); /* ml_kem.c:1093 */
union byte *Decaps_internal(const struct PARAMS *params, const union byte *dk, const union byte *c); /* ml_kem.c:1136 */

#if defined(__GNUC__)
#pragma GCC visibility pop
#endif
#ifdef __cplusplus
}
#endif
#endif /* ML_KEM_H */
