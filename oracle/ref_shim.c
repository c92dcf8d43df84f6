/*
 * oracle/ref_shim.c -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * Makes the UNMODIFIED reference callable with dense byte / uint16 buffers.
 * No reference source is copied: this translation unit textually includes
 * /root/reference/ml_kem.c (found through -I/root/reference, see oracle/Makefile)
 * so that its `static` functions and `union integer` are reachable, and is linked
 * with the reference's own sha3.c.  The result goes to oracle/_ref/ (git-ignored).
 *
 * Every wrapper converts dense inputs to the reference's 4-byte-per-element unions
 * (ml_kem.h:35-38, ml_kem.c:20-23), calls the reference function, reads back only the
 * `.e` / `.t` members (the upper bits are indeterminate) and frees what the reference
 * malloc'ed.
 */
#include "ml_kem.c" /* the reference, from its read-only mount */

#include <pthread.h>
#include <stdint.h>
#include <string.h>
#include <time.h>

#define REF_API __attribute__((visibility("default")))

static union byte *to_ub(const uint8_t *b, size_t n) {
    union byte *o = malloc(sizeof(union byte) * (n ? n : 1));
    for (size_t i = 0; i < n; i++) {
        o[i].e = 0; /* keep valgrind quiet; the reference only ever reads .e */
        o[i].e = b[i];
    }
    return o;
}
static void from_ub(const union byte *u, size_t n, uint8_t *b) {
    for (size_t i = 0; i < n; i++) b[i] = (uint8_t)u[i].e;
}
static union integer *to_ui(const uint16_t *c, size_t n) {
    union integer *o = malloc(sizeof(union integer) * n);
    for (size_t i = 0; i < n; i++) {
        o[i].l = 0;
        o[i].t = c[i];
    }
    return o;
}
static void from_ui(const union integer *u, size_t n, uint16_t *c) {
    for (size_t i = 0; i < n; i++) c[i] = (uint16_t)u[i].t;
}

REF_API void ref_sizes(unsigned out[6]) {
    out[0] = sizeof(union byte);
    out[1] = sizeof(union integer);
    out[2] = sizeof(union bit);
    out[3] = sizeof(struct PARAMS);
    out[4] = sizeof(struct PKE);
    out[5] = sizeof(struct KEM);
}

REF_API int ref_init(int set, unsigned out[5]) {
    int before = ml_errno;
    ml_errno = 0;
    struct PARAMS p = init((enum ML_KEM)set);
    int err = ml_errno;
    ml_errno = before;
    if (err) return err;
    out[0] = p.k.e; out[1] = p.n1.e; out[2] = p.n2.e; out[3] = p.du.e; out[4] = p.dv.e;
    return 0;
}

REF_API uint8_t ref_bitrev7(uint8_t r) {
    union byte b;
    b.e = 0;
    b.s = r;
    return (uint8_t)BitRev7(b).s;
}
REF_API uint16_t ref_compress(uint16_t x, unsigned d) {
    union integer v;
    v.l = 0;
    v.t = x;
    return (uint16_t)Compress(v, d).t;
}
REF_API uint16_t ref_decompress(uint16_t y, unsigned d) {
    union integer v;
    v.l = 0;
    v.t = y;
    return (uint16_t)Decompress(v, d).t;
}
REF_API void ref_byte_encode(const uint16_t F[256], unsigned d, uint8_t *B) {
    union integer *f = to_ui(F, 256);
    union byte *o = ByteEncode(f, d);
    from_ub(o, 32 * d, B);
    free(o);
    free(f);
}
REF_API void ref_byte_decode(const uint8_t *B, unsigned d, uint16_t F[256]) {
    union byte *b = to_ub(B, 32 * d);
    union integer *o = ByteDecode(b, d);
    from_ui(o, 256, F);
    free(o);
    free(b);
}
/* B is in/out: the reference may bump B[32], B[33] (ml_kem.c:237-242). */
REF_API void ref_sample_ntt(uint8_t B[34], uint16_t a[256]) {
    union byte *b = to_ub(B, 34);
    union integer *o = SampleNTT(b);
    from_ui(o, 256, a);
    from_ub(b, 34, B);
    free(o);
    free(b);
}
REF_API void ref_sample_cbd(const uint8_t *B, unsigned eta, uint16_t f[256]) {
    union byte *b = to_ub(B, 64 * eta);
    union integer *o = SamplePolyCBD(b, eta);
    from_ui(o, 256, f);
    free(o);
    free(b);
}
REF_API void ref_ntt(const uint16_t f[256], uint16_t fh[256]) {
    union integer *i = to_ui(f, 256);
    union integer *o = NTT(i);
    from_ui(o, 256, fh);
    free(o);
    free(i);
}
REF_API void ref_intt(const uint16_t fh[256], uint16_t f[256]) {
    union integer *i = to_ui(fh, 256);
    union integer *o = InverseNTT(i);
    from_ui(o, 256, f);
    free(o);
    free(i);
}
REF_API void ref_basecase_multiply(uint16_t a0, uint16_t a1, uint16_t b0, uint16_t b1, uint16_t gamma, uint16_t c[2]) {
    union integer A0, A1, B0, B1, G;
    A0.l = A1.l = B0.l = B1.l = G.l = 0;
    A0.t = a0; A1.t = a1; B0.t = b0; B1.t = b1; G.l = gamma;
    union integer *o = BaseCaseMultiply(A0, A1, B0, B1, G);
    c[0] = (uint16_t)o[0].t;
    c[1] = (uint16_t)o[1].t;
    free(o);
}
REF_API void ref_multiply_ntts(const uint16_t f[256], const uint16_t g[256], uint16_t h[256]) {
    union integer *a = to_ui(f, 256), *b = to_ui(g, 256);
    union integer *o = MultiplyNTTs(a, b);
    from_ui(o, 256, h);
    free(o);
    free(a);
    free(b);
}
REF_API void ref_poly_add(const uint16_t u[256], const uint16_t v[256], uint16_t z[256]) {
    union integer *a = to_ui(u, 256), *b = to_ui(v, 256);
    union integer *o = PolyAddition(a, b);
    from_ui(o, 256, z);
    free(o); free(a); free(b);
}
REF_API void ref_poly_sub(const uint16_t u[256], const uint16_t v[256], uint16_t z[256]) {
    union integer *a = to_ui(u, 256), *b = to_ui(v, 256);
    union integer *o = PolySubtraction(a, b);
    from_ui(o, 256, z);
    free(o); free(a); free(b);
}
/* ml_kem.c:618 VectorMultiply on k polynomials each (u, v: k x 256 contiguous). */
REF_API void ref_vector_multiply(unsigned k, const uint16_t *u, const uint16_t *v, uint16_t w[256]) {
    union integer *up[16], *vp[16];
    if (k > 16) return;
    for (unsigned i = 0; i < k; i++) {
        up[i] = to_ui(u + 256 * i, 256);
        vp[i] = to_ui(v + 256 * i, 256);
    }
    union integer *o = VectorMultiply(up, vp, k);
    from_ui(o, 256, w);
    free(o);
    for (unsigned i = 0; i < k; i++) { free(up[i]); free(vp[i]); }
}
REF_API void ref_PRF(const uint8_t s[32], uint8_t b, unsigned eta, uint8_t *out) {
    union byte *S = to_ub(s, 32), B;
    B.e = 0;
    B.e = b;
    union byte *o = PRF(S, B, eta);
    from_ub(o, 64 * eta, out);
    free(o);
    free(S);
}
REF_API void ref_H(const uint8_t *in, unsigned len, uint8_t out[32]) {
    union byte *I = to_ub(in, len), *o = H(I, len);
    from_ub(o, 32, out);
    free(o); free(I);
}
REF_API void ref_J(const uint8_t *in, unsigned len, uint8_t out[32]) {
    union byte *I = to_ub(in, len), *o = J(I, len);
    from_ub(o, 32, out);
    free(o); free(I);
}
REF_API void ref_G(const uint8_t *in, unsigned len, uint8_t out[64]) {
    union byte *I = to_ub(in, len), *o = G(I, len);
    from_ub(o, 64, out);
    free(o); free(I);
}

/* sha3_b (sha3.c:408) on bit strings packed LSB-first into bytes; out receives ceil(d/8) bytes. */
REF_API void ref_sha3_bits(const uint8_t *msg, unsigned nbits, const uint8_t sfx[4], unsigned c, unsigned d, uint8_t *out) {
    union bit *b = malloc(sizeof(union bit) * (nbits ? nbits : 1));
    union bit sf[4];
    for (unsigned i = 0; i < nbits; i++) b[i].b = (msg[i >> 3] >> (i & 7)) & 1;
    for (int i = 0; i < 4; i++) sf[i].b = sfx[i] & 1;
    union bit *o = sha3_b(b, nbits, d, c, sf);
    memset(out, 0, (d + 7) / 8);
    for (unsigned i = 0; i < d; i++) out[i >> 3] |= (uint8_t)((o[i].b & 1) << (i & 7));
    free(o);
    free(b);
}
/* sha3_s (sha3.c:465) on a character string. */
REF_API void ref_sha3_s(const char *str, unsigned len, const uint8_t sfx[4], unsigned c, unsigned d, uint8_t *out) {
    union bit sf[4];
    for (int i = 0; i < 4; i++) sf[i].b = sfx[i] & 1;
    unsigned char *o = sha3_s(str, len, d, c, sf);
    memcpy(out, o, d / 8);
    free(o);
}

static int params_of(int set, struct PARAMS *p) {
    if (set != 512 && set != 768 && set != 1024) return -1;
    *p = init((enum ML_KEM)set);
    return 0;
}

REF_API int ref_pke_keygen(int set, const uint8_t d[32], uint8_t *ek, uint8_t *dkpke) {
    struct PARAMS p;
    if (params_of(set, &p)) return -1;
    union byte *D = to_ub(d, 32);
    struct PKE keys = PKE_KeyGen(&p, D);
    from_ub(keys.ek, keys.ek_len, ek);
    from_ub(keys.dk, keys.dk_len, dkpke);
    free(keys.ek); free(keys.dk); free(D);
    return 0;
}
REF_API int ref_pke_encrypt(int set, const uint8_t *ek, const uint8_t m[32], const uint8_t r[32], uint8_t *c) {
    struct PARAMS p;
    if (params_of(set, &p)) return -1;
    unsigned k = p.k.e;
    union byte *E = to_ub(ek, 384 * k + 32), *M = to_ub(m, 32), *R = to_ub(r, 32);
    union byte *o = PKE_Encrypt(&p, E, M, R);
    from_ub(o, 32 * (p.du.e * k + p.dv.e), c);
    free(o); free(E); free(M); free(R);
    return 0;
}
REF_API int ref_pke_decrypt(int set, const uint8_t *dkpke, const uint8_t *c, uint8_t m[32]) {
    struct PARAMS p;
    if (params_of(set, &p)) return -1;
    unsigned k = p.k.e;
    union byte *D = to_ub(dkpke, 384 * k), *C = to_ub(c, 32 * (p.du.e * k + p.dv.e));
    union byte *o = PKE_Decrypt(&p, D, C);
    from_ub(o, 32, m);
    free(o); free(D); free(C);
    return 0;
}
REF_API int ref_keygen_internal(int set, const uint8_t d[32], const uint8_t z[32], uint8_t *ek, uint8_t *dk) {
    struct PARAMS p;
    if (params_of(set, &p)) return -1;
    union byte *D = to_ub(d, 32), *Z = to_ub(z, 32);
    struct PKE keys = KeyGen_internal(&p, D, Z);
    from_ub(keys.ek, keys.ek_len, ek);
    from_ub(keys.dk, keys.dk_len, dk);
    free(keys.ek); free(keys.dk); free(D); free(Z);
    return 0;
}
REF_API int ref_encaps_internal(int set, const uint8_t *ek, const uint8_t m[32], uint8_t *c, uint8_t K[32]) {
    struct PARAMS p;
    if (params_of(set, &p)) return -1;
    union byte *E = to_ub(ek, 384 * p.k.e + 32), *M = to_ub(m, 32);
    struct KEM r = Encaps_internal(&p, E, M);
    from_ub(r.K, 32, K);
    from_ub(r.c, r.c_len, c);
    free(r.c); free(E); free(M);
    return 0;
}
REF_API int ref_decaps_internal(int set, const uint8_t *dk, const uint8_t *c, uint8_t K[32]) {
    struct PARAMS p;
    if (params_of(set, &p)) return -1;
    unsigned k = p.k.e;
    union byte *D = to_ub(dk, 768 * k + 96), *C = to_ub(c, 32 * (p.du.e * k + p.dv.e));
    union byte *o = Decaps_internal(&p, D, C);
    from_ub(o, 32, K);
    free(o); free(D); free(C);
    return 0;
}

/*
 * Public wrappers (ml_kem.c:1233-1359) -- random inputs, so only structure and error codes are
 * observable.  Each returns the ml_errno the call produced (ml_errno is reset first because the
 * reference never clears it).  stderr carries the reference's ERR_MSG text.
 */
REF_API int ref_KEM_KeyGen(int set, uint8_t *ek, uint8_t *dk) {
    struct PARAMS p;
    if (params_of(set, &p)) return -1;
    ml_errno = 0;
    struct PKE keys = KEM_KeyGen(&p);
    if (ml_errno) return ml_errno;
    from_ub(keys.ek, keys.ek_len, ek);
    from_ub(keys.dk, keys.dk_len, dk);
    free(keys.ek); free(keys.dk);
    return 0;
}
REF_API int ref_KEM_Encaps(int set, const uint8_t *ek, unsigned ek_len, uint8_t *c, uint8_t K[32]) {
    struct PARAMS p;
    if (params_of(set, &p)) return -1;
    union byte *E = to_ub(ek, ek_len);
    ml_errno = 0;
    struct KEM r = KEM_Encaps(&p, E, ek_len);
    int err = ml_errno;
    if (!err) {
        from_ub(r.K, 32, K);
        from_ub(r.c, r.c_len, c);
        free(r.c);
    }
    free(E);
    return err;
}
REF_API int ref_KEM_Decaps(int set, const uint8_t *dk, unsigned dk_len, const uint8_t *c, unsigned c_len, uint8_t K[32]) {
    struct PARAMS p;
    if (params_of(set, &p)) return -1;
    union byte *D = to_ub(dk, dk_len), *C = to_ub(c, c_len);
    ml_errno = 0;
    union byte *o = KEM_Decaps(&p, D, dk_len, C, c_len);
    int err = ml_errno;
    if (o) {
        from_ub(o, 32, K);
        free(o);
    }
    free(D); free(C);
    return err;
}

/* ------------------------------------------------------------------ */
/* Timing drivers for bench.py (cpu_baseline / --impl reference).       */
/* One pthread per requested thread; the reference is re-entrant apart   */
/* from ml_errno, which the internal functions never touch.              */
/* ------------------------------------------------------------------ */
struct job {
    int set, what; /* what: 0 keygen, 1 encaps, 2 decaps, 3 encaps+decaps */
    size_t begin, end;
    const uint8_t *d, *z, *m, *ek, *dk, *c_in;
    uint8_t *ek_out, *dk_out, *c_out, *K_out;
    unsigned ekl, dkl, cl;
};
static void *job_main(void *arg) {
    struct job *j = arg;
    for (size_t i = j->begin; i < j->end; i++) {
        if (j->what == 0) {
            ref_keygen_internal(j->set, j->d + 32 * i, j->z + 32 * i, j->ek_out + (size_t)j->ekl * i,
                                j->dk_out + (size_t)j->dkl * i);
        } else {
            if (j->what == 1 || j->what == 3)
                ref_encaps_internal(j->set, j->ek + (size_t)j->ekl * i, j->m + 32 * i, j->c_out + (size_t)j->cl * i,
                                    j->K_out + 32 * i);
            if (j->what == 2)
                ref_decaps_internal(j->set, j->dk + (size_t)j->dkl * i, j->c_in + (size_t)j->cl * i, j->K_out + 32 * i);
            if (j->what == 3)
                ref_decaps_internal(j->set, j->dk + (size_t)j->dkl * i, j->c_out + (size_t)j->cl * i,
                                    j->K_out + 32 * i);
        }
    }
    return NULL;
}
static double now_s(void) {
    struct timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    return ts.tv_sec + 1e-9 * ts.tv_nsec;
}
static double run_jobs(struct job proto, size_t n, int threads) {
    if (threads < 1) threads = 1;
    if ((size_t)threads > n) threads = (int)n;
    pthread_t tid[256];
    struct job jobs[256];
    if (threads > 256) threads = 256;
    double t0 = now_s();
    for (int t = 0; t < threads; t++) {
        jobs[t] = proto;
        jobs[t].begin = n * t / threads;
        jobs[t].end = n * (t + 1) / threads;
        pthread_create(&tid[t], NULL, job_main, &jobs[t]);
    }
    for (int t = 0; t < threads; t++) pthread_join(tid[t], NULL);
    return now_s() - t0;
}
/* Each returns wall seconds for n items on `threads` threads. */
REF_API double ref_time_keygen(int set, size_t n, int threads, const uint8_t *d, const uint8_t *z, uint8_t *ek, uint8_t *dk) {
    struct PARAMS p;
    if (params_of(set, &p)) return -1;
    struct job j = {0};
    j.set = set; j.what = 0; j.d = d; j.z = z; j.ek_out = ek; j.dk_out = dk;
    j.ekl = 384 * p.k.e + 32; j.dkl = 768 * p.k.e + 96;
    return run_jobs(j, n, threads);
}
REF_API double ref_time_encaps(int set, size_t n, int threads, const uint8_t *ek, const uint8_t *m, uint8_t *c, uint8_t *K) {
    struct PARAMS p;
    if (params_of(set, &p)) return -1;
    struct job j = {0};
    j.set = set; j.what = 1; j.ek = ek; j.m = m; j.c_out = c; j.K_out = K;
    j.ekl = 384 * p.k.e + 32; j.cl = 32 * (p.du.e * p.k.e + p.dv.e);
    return run_jobs(j, n, threads);
}
REF_API double ref_time_decaps(int set, size_t n, int threads, const uint8_t *dk, const uint8_t *c, uint8_t *K) {
    struct PARAMS p;
    if (params_of(set, &p)) return -1;
    struct job j = {0};
    j.set = set; j.what = 2; j.dk = dk; j.c_in = c; j.K_out = K;
    j.dkl = 768 * p.k.e + 96; j.cl = 32 * (p.du.e * p.k.e + p.dv.e);
    return run_jobs(j, n, threads);
}
/* Encaps immediately followed by Decaps of the fresh ciphertext, per item. */
REF_API double ref_time_pairs(int set, size_t n, int threads, const uint8_t *ek, const uint8_t *dk, const uint8_t *m,
                              uint8_t *c, uint8_t *K) {
    struct PARAMS p;
    if (params_of(set, &p)) return -1;
    struct job j = {0};
    j.set = set; j.what = 3; j.ek = ek; j.dk = dk; j.m = m; j.c_out = c; j.K_out = K;
    j.ekl = 384 * p.k.e + 32; j.dkl = 768 * p.k.e + 96; j.cl = 32 * (p.du.e * p.k.e + p.dv.e);
    return run_jobs(j, n, threads);
}
/* Per-call seconds of the three ring kernels, single thread. */
REF_API void ref_time_ring(size_t reps, const uint16_t f[256], const uint16_t g[256], double out[3]) {
    union integer *a = to_ui(f, 256), *b = to_ui(g, 256);
    double t0 = now_s();
    for (size_t i = 0; i < reps; i++) free(NTT(a));
    out[0] = (now_s() - t0) / reps;
    t0 = now_s();
    for (size_t i = 0; i < reps; i++) free(InverseNTT(a));
    out[1] = (now_s() - t0) / reps;
    t0 = now_s();
    for (size_t i = 0; i < reps; i++) free(MultiplyNTTs(a, b));
    out[2] = (now_s() - t0) / reps;
    free(a); free(b);
}
