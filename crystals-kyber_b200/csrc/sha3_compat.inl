// sha3_compat.inl -- the reference's SHA-3 front-end (include/sha3.h) on top of mlkem_b200_sha3_bits_batch.
// h2b / b2h and the cell conversions are layout code (FIPS 202 B.1); the sponge runs on the GPU.
#include "../../include/sha3.h"

extern "C" {

union bit *h2b(const union hex *H, unsigned int m, unsigned int n) {  // sha3.c:329
    unsigned total = 8 * m, keep = n < total ? n : total;
    union bit *S = (union bit *)calloc(keep ? keep : 1, sizeof(union bit));
    for (unsigned i = 0; i < keep; i++) {
        unsigned h = 16 * H[2 * (i / 8)].d + H[2 * (i / 8) + 1].d;
        S[i].b = (h >> (i % 8)) & 1u;
    }
    return S;
}

union hex *b2h(const union bit *S, unsigned int n) {  // sha3.c:367
    unsigned m = (n + 7) / 8;
    union hex *H = (union hex *)calloc(2 * m ? 2 * m : 1, sizeof(union hex));
    for (unsigned i = 0; i < m; i++) {
        unsigned h = 0;
        for (unsigned j = 0; j < 8; j++)
            if (8 * i + j < n) h += (S[8 * i + j].b & 1u) << j;
        H[2 * i].d = (h >> 4) & 15;
        H[2 * i + 1].d = h & 15;
    }
    return H;
}

union bit *sha3_b(const union bit *bstr, unsigned int n, unsigned int d, unsigned int c, union bit sfx[4]) {  // sha3.c:408
    std::vector<uint8_t> msg((n + 7) / 8 + 1, 0), out((d + 7) / 8 + 1, 0);
    for (unsigned i = 0; i < n; i++) msg[i >> 3] |= (uint8_t)((bstr[i].b & 1u) << (i & 7));
    uint8_t sf[4] = {(uint8_t)sfx[0].b, (uint8_t)sfx[1].b, (uint8_t)sfx[2].b, (uint8_t)sfx[3].b};
    int rc = mlkem_b200_sha3_bits_batch(1, msg.data(), n, sf, c, d, out.data(), nullptr);
    if (rc) {
        cuda_failure("sha3_b", rc);
        return NULL;
    }
    union bit *D = (union bit *)calloc(d ? d : 1, sizeof(union bit));
    for (unsigned i = 0; i < d; i++) D[i].b = (out[i >> 3] >> (i & 7)) & 1u;
    return D;
}

union hex *sha3_h(const union hex *hstr, unsigned int m, unsigned int d, unsigned int c, union bit sfx[4]) {  // sha3.c:443
    union bit *Nb = h2b(hstr, m, 8 * m);
    union bit *M = sha3_b(Nb, 8 * m, d, c, sfx);
    free(Nb);
    if (!M) return NULL;
    union hex *D = b2h(M, d);
    free(M);
    return D;
}

unsigned char *sha3_s(const char *cstr, unsigned int m, unsigned int d, unsigned int c, union bit sfx[4]) {  // sha3.c:465
    union hex *Hx = (union hex *)calloc(2 * m ? 2 * m : 1, sizeof(union hex));
    for (unsigned i = 0; i < m; i++) {
        Hx[2 * i].d = ((unsigned char)cstr[i] >> 4) & 0x0F;
        Hx[2 * i + 1].d = (unsigned char)cstr[i] & 0x0F;
    }
    union hex *Z = sha3_h(Hx, m, d, c, sfx);
    free(Hx);
    if (!Z) return NULL;
    unsigned char *D = (unsigned char *)malloc(d / 8 ? d / 8 : 1);
    for (unsigned i = 0; i < d / 8; i++) D[i] = (unsigned char)((Z[2 * i].d << 4) ^ Z[2 * i + 1].d);
    free(Z);
    return D;
}

}  // extern "C"
