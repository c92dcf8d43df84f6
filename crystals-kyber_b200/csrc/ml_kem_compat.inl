// ml_kem_compat.inl -- the reference-signature API (include/ml_kem.h) on top of the batched C ABI.
//
// Every function runs a batch of ONE on the GPU.  KeyGen_internal / Encaps_internal / Decaps_internal (and through them
// KEM_KeyGen / KEM_Encaps / KEM_Decaps) hand the reference's stride-4 union arrays (one value per 4-byte cell,
// ml_kem.h:35-38 / ml_kem.c:20-23 in /root/reference) to the batched cell-layout entry points, which convert on the
// device; the primitives convert on the host.  Only layout conversion, length checks, the random-byte source and the
// error reporting live on the host; every arithmetic result comes from a CUDA kernel.
//
// Included at the end of mlkem_b200.cu (single translation unit).
#include "../../include/ml_kem.h"

#include <fcntl.h>
#include <unistd.h>

extern "C" {

int ml_errno = 0;  // ml_kem.c:16

}  // extern "C"

namespace {

// The reference prints "ERROR: <__FILE__> - <__LINE__>\n\t<msg>\n" (ml_kem.c:11-13); the line numbers are
// those of its ERR_MSG call sites so that stderr reads the same as with the reference's object file.
void err_msg(int ref_line, const char *msg) {
    fprintf(stderr, "ERROR: %s - %d\n", "ml_kem.c", ref_line);
    fprintf(stderr, "\t%s\n", msg);
}
void cuda_failure(const char *where, int rc) {
    fprintf(stderr, "ERROR: mlkem_b200 - %s: rc=%d %s\n", where, rc, mlkem_b200_last_error());
    ml_errno = rc == MLKEM_B200_ERR_CUDA ? -10 : rc;
}

// `union byte` is one 4-byte cell (include/ml_kem.h, D5): the batched cell-layout entry points take it as uint32_t.
static_assert(sizeof(union byte) == 4, "cell size");
const uint32_t *cells_of(const union byte *b) { return reinterpret_cast<const uint32_t *>(b); }
uint32_t *cells_of(union byte *b) { return reinterpret_cast<uint32_t *>(b); }

std::vector<uint8_t> dense_bytes(const union byte *b, size_t n) {
    std::vector<uint8_t> v(n);
    for (size_t i = 0; i < n; i++) v[i] = (uint8_t)b[i].e;
    return v;
}
union byte *wide_bytes(const uint8_t *b, size_t n) {
    union byte *o = (union byte *)calloc(n ? n : 1, sizeof(union byte));
    for (size_t i = 0; i < n; i++) o[i].e = b[i];
    return o;
}
std::vector<uint16_t> dense_coeffs(const union integer *f, size_t n) {
    std::vector<uint16_t> v(n);
    for (size_t i = 0; i < n; i++) v[i] = (uint16_t)f[i].t;
    return v;
}
union integer *wide_coeffs(const uint16_t *f, size_t n) {
    union integer *o = (union integer *)calloc(n ? n : 1, sizeof(union integer));
    for (size_t i = 0; i < n; i++) o[i].t = f[i];
    return o;
}

// The library instantiates the three FIPS 203 parameter sets; PARAMS must describe one of them.
int set_of(const struct PARAMS *p) {
    unsigned k = p->k.e, n1 = p->n1.e, n2 = p->n2.e, du = p->du.e, dv = p->dv.e;
    if (k == 2 && n1 == 3 && n2 == 2 && du == 10 && dv == 4) return 512;
    if (k == 3 && n1 == 2 && n2 == 2 && du == 10 && dv == 4) return 768;
    if (k == 4 && n1 == 2 && n2 == 2 && du == 11 && dv == 5) return 1024;
    return 0;
}

// ml_kem.c:458 getRandomBytes: 32 unsigned ints from /dev/urandom, each reduced mod 256.
bool random_bytes(uint8_t out[32]) {
#ifdef __linux__
    unsigned int rnd[32];
    int fd = open("/dev/urandom", O_RDONLY);
    if (fd < 0) return false;
    ssize_t got = read(fd, rnd, sizeof rnd);
    close(fd);
    if (got != (ssize_t)sizeof rnd) return false;
    for (int i = 0; i < 32; i++) out[i] = (uint8_t)(rnd[i] % N);
    return true;
#else
    printf("ml_kem.c:getRandomBytes() :: Operating system not supported\n");  // ml_kem.c:485
    (void)out;
    return false;
#endif
}

// BaseCaseMultiply takes an arbitrary gamma, which the table-driven kernels do not: a one-thread kernel.
__global__ void k_basecase_single(uint32_t a0, uint32_t a1, uint32_t b0, uint32_t b1, uint32_t gamma, uint32_t *out) {
    uint32_t t = (a1 * b1) % mlkem::kQ;  // ml_kem.c:402-409
    t = (t * (gamma % mlkem::kQ)) % mlkem::kQ;
    t += (a0 * b0) % mlkem::kQ;
    out[0] = t % mlkem::kQ;
    t = (a0 * b1) % mlkem::kQ;
    t += (a1 * b0) % mlkem::kQ;
    out[1] = t % mlkem::kQ;
}

struct PKE empty_pke() {
    struct PKE k;
    memset(&k, 0, sizeof k);
    return k;
}
struct KEM empty_kem() {
    struct KEM k;
    memset(&k, 0, sizeof k);
    return k;
}

struct PKE keygen_common(const struct PARAMS *params, const union byte *d, const union byte *z) {
    struct PKE out = empty_pke();
    int set = set_of(params);
    if (!set) {
        cuda_failure("unsupported PARAMS", MLKEM_B200_ERR_PARAM);
        return out;
    }
    unsigned ekl = mlkem_b200_ek_bytes(set), dkl = z ? mlkem_b200_dk_bytes(set) : mlkem_b200_dkpke_bytes(set);
    if (z) {
        // the caller's cell arrays go to the device as they are; the layout conversion is a kernel (mlkem_b200_*_cells_batch)
        union byte *ekc = (union byte *)calloc(ekl, sizeof(union byte)), *dkc = (union byte *)calloc(dkl, sizeof(union byte));
        int rc = mlkem_b200_keygen_cells_batch(set, 1, cells_of(d), cells_of(z), cells_of(ekc), cells_of(dkc), nullptr);
        if (rc) {
            free(ekc);
            free(dkc);
            cuda_failure("keygen", rc);
            return out;
        }
        out.ek = ekc;
        out.dk = dkc;
    } else {
        std::vector<uint8_t> dd = dense_bytes(d, 32), ek(ekl), dk(dkl);
        int rc = mlkem_b200_pke_keygen_batch(set, 1, dd.data(), ek.data(), dk.data(), nullptr);
        if (rc) {
            cuda_failure("keygen", rc);
            return out;
        }
        out.ek = wide_bytes(ek.data(), ekl);
        out.dk = wide_bytes(dk.data(), dkl);
    }
    out.ek_len = ekl;
    out.dk_len = dkl;
    return out;
}

}  // namespace

extern "C" {

// ---- ml_kem.c:1363 -------------------------------------------------------------------------------------
const struct PARAMS init(enum ML_KEM param_set) {
    struct PARAMS p;
    memset(&p, 0, sizeof p);
    switch ((int)param_set) {
    case 512: p.k.e = 2; p.n1.e = 3; p.n2.e = 2; p.du.e = 10; p.dv.e = 4; break;
    case 768: p.k.e = 3; p.n1.e = 2; p.n2.e = 2; p.du.e = 10; p.dv.e = 4; break;
    case 1024: p.k.e = 4; p.n1.e = 2; p.n2.e = 2; p.du.e = 11; p.dv.e = 5; break;
    default:
        err_msg(1390, "init() :: Invalid paramater set provided\n");
        ml_errno = -1;
    }
    return p;
}

// ---- L1 codec ------------------------------------------------------------------------------------------
union byte BitRev7(union byte r) {  // ml_kem.c:26 (an index permutation, host side)
    union byte o;
    o.e = 0;
    o.s = bitrev7(r.s);
    return o;
}
union byte *BitsToBytes(const union bit *b, unsigned int l) {  // ml_kem.c:47 (layout conversion)
    union byte *B = (union byte *)calloc(l / 8 ? l / 8 : 1, sizeof(union byte));
    for (unsigned i = 0; i < l; i++) B[i / 8].e |= (b[i].b & 1u) << (i % 8);
    return B;
}
union bit *BytesToBits(const union byte *B, unsigned int L) {  // ml_kem.c:62 (layout conversion)
    union bit *b = (union bit *)calloc(L ? 8 * (size_t)L : 1, sizeof(union bit));
    for (unsigned i = 0; i < L; i++)
        for (unsigned j = 0; j < 8; j++) b[8 * i + j].b = (B[i].e >> j) & 1u;
    return b;
}
static union integer compress_one(union integer x, unsigned d, bool inverse) {
    if (d < 1 || d >= 12) return x;  // identity for d = 12, ml_kem.c:86,107
    uint16_t in[8] = {(uint16_t)x.t, 0, 0, 0, 0, 0, 0, 0}, out[8];
    int rc = inverse ? mlkem_b200_decompress_batch((int)d, 8, in, out, nullptr) : mlkem_b200_compress_batch((int)d, 8, in, out, nullptr);
    if (rc) {
        cuda_failure("compress", rc);
        return x;
    }
    x.t = out[0];
    return x;
}
union integer Compress(union integer x, unsigned int d) { return compress_one(x, d, false); }      // ml_kem.c:83
union integer Decompress(union integer y, unsigned int d) { return compress_one(y, d, true); }     // ml_kem.c:104

union byte *ByteEncode(const union integer *F, unsigned int d) {  // ml_kem.c:125
    std::vector<uint16_t> f = dense_coeffs(F, 256);
    std::vector<uint8_t> B(32 * (size_t)d);
    int rc = mlkem_b200_byte_encode_batch((int)d, 1, f.data(), B.data(), nullptr);
    if (rc) {
        cuda_failure("ByteEncode", rc);
        return NULL;
    }
    return wide_bytes(B.data(), B.size());
}
union integer *ByteDecode(const union byte *B, unsigned int d) {  // ml_kem.c:153
    std::vector<uint8_t> b = dense_bytes(B, 32 * (size_t)d);
    std::vector<uint16_t> f(256);
    int rc = mlkem_b200_byte_decode_batch((int)d, 1, b.data(), f.data(), nullptr);
    if (rc) {
        cuda_failure("ByteDecode", rc);
        return NULL;
    }
    return wide_coeffs(f.data(), 256);
}

// ---- L2 samplers ---------------------------------------------------------------------------------------
union integer *SampleNTT(union byte *B) {  // ml_kem.c:189
    std::vector<uint8_t> b = dense_bytes(B, 34), after(34);
    std::vector<uint16_t> a(256);
    int rc = mlkem_b200_sample_ntt_batch(1, b.data(), a.data(), after.data(), nullptr);
    if (rc) {
        cuda_failure("SampleNTT", rc);
        return NULL;
    }
    B[32].e = after[32];  // the reference updates the caller's buffer when it restarts (:237-242)
    B[33].e = after[33];
    return wide_coeffs(a.data(), 256);
}
union integer *SamplePolyCBD(const union byte *B, unsigned int n) {  // ml_kem.c:253
    std::vector<uint8_t> b = dense_bytes(B, 64 * (size_t)n);
    std::vector<uint16_t> f(256);
    int rc = mlkem_b200_cbd_batch((int)n, 1, b.data(), f.data(), nullptr);
    if (rc) {
        cuda_failure("SamplePolyCBD", rc);
        return NULL;
    }
    return wide_coeffs(f.data(), 256);
}

// ---- L3 ring -------------------------------------------------------------------------------------------
static union integer *poly_op(const union integer *a, const union integer *b, int which, const char *name) {
    std::vector<uint16_t> x = dense_coeffs(a, 256), y, out(256);
    int rc;
    if (which == 0) rc = mlkem_b200_ntt_batch(1, x.data(), out.data(), nullptr);
    else if (which == 1) rc = mlkem_b200_intt_batch(1, x.data(), out.data(), nullptr);
    else {
        y = dense_coeffs(b, 256);
        rc = mlkem_b200_multiply_ntts_batch(1, x.data(), y.data(), out.data(), nullptr);
    }
    if (rc) {
        cuda_failure(name, rc);
        return NULL;
    }
    return wide_coeffs(out.data(), 256);
}
union integer *NTT(const union integer *f) { return poly_op(f, NULL, 0, "NTT"); }                     // ml_kem.c:287
union integer *InverseNTT(const union integer *fh) { return poly_op(fh, NULL, 1, "InverseNTT"); }     // ml_kem.c:336
union integer *MultiplyNTTs(const union integer *fh, const union integer *gh) { return poly_op(fh, gh, 2, "MultiplyNTTs"); }  // ml_kem.c:415

union integer *BaseCaseMultiply(union integer a0, union integer a1, union integer b0, union integer b1, union integer gamma) {  // ml_kem.c:395
    int dev;
    DeviceCtx *ctx;
    union integer *C = (union integer *)calloc(2, sizeof(union integer));
    DeviceGuard guard;
    if (int rc = acquire(nullptr, &dev, &ctx, guard)) {
        cuda_failure("BaseCaseMultiply", rc);
        return C;
    }
    uint32_t *d_out = nullptr, h_out[2] = {0, 0};
    bool ok = cudaMalloc(&d_out, 8) == cudaSuccess;
    if (ok) {
        k_basecase_single<<<1, 1, 0, ctx->stream[0]>>>(a0.t, a1.t, b0.t, b1.t, gamma.l, d_out);
        g_launches.fetch_add(1);
        ok = cudaMemcpyAsync(h_out, d_out, 8, cudaMemcpyDeviceToHost, ctx->stream[0]) == cudaSuccess &&
             cudaStreamSynchronize(ctx->stream[0]) == cudaSuccess;
        cudaFree(d_out);
    }
    if (!ok) cuda_failure("BaseCaseMultiply", MLKEM_B200_ERR_CUDA);
    C[0].t = h_out[0];
    C[1].t = h_out[1];
    return C;
}

// ---- L4 K-PKE ------------------------------------------------------------------------------------------
struct PKE PKE_KeyGen(const struct PARAMS *params, const union byte *d) { return keygen_common(params, d, NULL); }  // ml_kem.c:651

union byte *PKE_Encrypt(const struct PARAMS *params, const union byte *ek, const union byte *m, const union byte *r) {  // ml_kem.c:776
    int set = set_of(params);
    if (!set) {
        cuda_failure("unsupported PARAMS", MLKEM_B200_ERR_PARAM);
        return NULL;
    }
    unsigned ekl = mlkem_b200_ek_bytes(set), cl = mlkem_b200_ct_bytes(set);
    std::vector<uint8_t> e = dense_bytes(ek, ekl), mm = dense_bytes(m, 32), rr = dense_bytes(r, 32), c(cl);
    int rc = mlkem_b200_pke_encrypt_batch(set, 1, e.data(), mm.data(), rr.data(), c.data(), nullptr);
    if (rc) {
        cuda_failure("PKE_Encrypt", rc);
        return NULL;
    }
    return wide_bytes(c.data(), cl);
}
union byte *PKE_Decrypt(const struct PARAMS *params, const union byte *dk, const union byte *c) {  // ml_kem.c:942
    int set = set_of(params);
    if (!set) {
        cuda_failure("unsupported PARAMS", MLKEM_B200_ERR_PARAM);
        return NULL;
    }
    unsigned dl = mlkem_b200_dkpke_bytes(set), cl = mlkem_b200_ct_bytes(set);
    std::vector<uint8_t> d = dense_bytes(dk, dl), cc = dense_bytes(c, cl), m(32);
    int rc = mlkem_b200_pke_decrypt_batch(set, 1, d.data(), dl, cc.data(), m.data(), nullptr);
    if (rc) {
        cuda_failure("PKE_Decrypt", rc);
        return NULL;
    }
    return wide_bytes(m.data(), 32);
}

// ---- L5 ML-KEM internal --------------------------------------------------------------------------------
struct PKE KeyGen_internal(const struct PARAMS *params, const union byte *d, const union byte *z) {  // ml_kem.c:1034
    return keygen_common(params, d, z);
}
struct KEM Encaps_internal(const struct PARAMS *params, const union byte *ek, const union byte *m) {  // ml_kem.c:1093
    struct KEM out = empty_kem();
    int set = set_of(params);
    if (!set) {
        cuda_failure("unsupported PARAMS", MLKEM_B200_ERR_PARAM);
        return out;
    }
    unsigned cl = mlkem_b200_ct_bytes(set);
    union byte *cc = (union byte *)calloc(cl, sizeof(union byte));
    int rc = mlkem_b200_encaps_cells_batch(set, 1, cells_of(ek), cells_of(m), cells_of(cc), cells_of(out.K), nullptr);
    if (rc) {
        free(cc);
        cuda_failure("Encaps_internal", rc);
        return out;
    }
    out.c = cc;
    out.c_len = cl;
    return out;
}
union byte *Decaps_internal(const struct PARAMS *params, const union byte *dk, const union byte *c) {  // ml_kem.c:1136
    int set = set_of(params);
    if (!set) {
        cuda_failure("unsupported PARAMS", MLKEM_B200_ERR_PARAM);
        return NULL;
    }
    union byte *K = (union byte *)calloc(32, sizeof(union byte));
    int rc = mlkem_b200_decaps_cells_batch(set, 1, cells_of(dk), cells_of(c), cells_of(K), nullptr);
    if (rc) {
        free(K);
        cuda_failure("Decaps_internal", rc);
        return NULL;
    }
    return K;
}

// ---- L6 public wrappers --------------------------------------------------------------------------------
struct PKE KEM_KeyGen(const struct PARAMS *params) {  // ml_kem.c:1233
    uint8_t d[32], z[32];
    if (!random_bytes(d) || !random_bytes(z)) {
        err_msg(1242, "KEM_KeyGen() :: Random bit generation failed\n");
        ml_errno = -2;
        return empty_pke();  // the reference returns an uninitialised struct here
    }
    union byte *D = wide_bytes(d, 32), *Z = wide_bytes(z, 32);
    struct PKE r = KeyGen_internal(params, D, Z);
    free(D);
    free(Z);
    return r;
}

struct KEM KEM_Encaps(const struct PARAMS *params, const union byte *ek, unsigned int ek_len) {  // ml_kem.c:1257
    unsigned len = 384 * params->k.e;
    if (len + 32 != ek_len) {  // type check, :1267
        err_msg(1268, "KEM_Encaps() :: Type check failed\n");
        ml_errno = -3;
        return empty_kem();
    }
    // modulus check (:1274-1291): ByteEncode12(ByteDecode12(ek)) == ek.  With the reference's ByteDecode12 (no
    // reduction mod q) this holds for every input, so the check is an identity and ml_errno -4 is unreachable.
    uint8_t m[32];
    if (!random_bytes(m)) {
        err_msg(1296, "KEM_Encaps() :: Random bit generation failed\n");
        ml_errno = -2;
        return empty_kem();
    }
    union byte *M = wide_bytes(m, 32);
    struct KEM r = Encaps_internal(params, ek, M);
    free(M);
    return r;
}

union byte *KEM_Decaps(const struct PARAMS *params, const union byte *dk, unsigned int dk_len, const union byte *c,
                       unsigned int c_len) {  // ml_kem.c:1310
    unsigned k = params->k.e;
    if (c_len != 32 * (params->du.e * k + params->dv.e)) {  // :1320
        err_msg(1322, "KEM_Decaps() :: Ciphertext type check failed\n");
        ml_errno = -3;
        return NULL;
    }
    if (dk_len != 768 * k + 96) {  // :1328
        err_msg(1330, "KEM_Decaps() :: Decapsulation key type check failed\n");
        ml_errno = -3;
        return NULL;
    }
    int set = set_of(params);
    if (!set) {
        cuda_failure("unsupported PARAMS", MLKEM_B200_ERR_PARAM);
        return NULL;
    }
    // hash check H(dk[384k : 768k+32]) == dk[768k+32 : 768k+64], :1336-1350 (on the device)
    std::vector<uint8_t> d = dense_bytes(dk, dk_len);
    int32_t status = 0;
    int rc = mlkem_b200_check_dk_batch(set, 1, d.data(), &status, nullptr);
    if (rc) {
        cuda_failure("KEM_Decaps", rc);
        return NULL;
    }
    if (status != 0) {
        err_msg(1346, "KEM_Decaps() :: Hash check failed\n");
        ml_errno = -5;
        return NULL;
    }
    return Decaps_internal(params, dk, c);
}

}  // extern "C"
