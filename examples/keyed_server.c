/* keyed_server.c -- the batched C ABI from plain C: what a decapsulating server does with libmlkem_b200.so.
 *
 *   make examples && ./build/keyed_server
 *
 * 1. KeyGen_internal for the seeds d = 00..1f, z = 20..3f (the KAT of SURVEY.md 8(c), ml_kem.c:1034),
 * 2. the same key as a resident, expanded key table built from the 64-byte seeds (mlkem_b200_keys_from_seeds),
 * 3. Encaps_internal with m = 40..5f (ml_kem.c:1093), Decaps_internal through the table (ml_kem.c:1136) for the
 *    ciphertext and for a tampered copy (bit 0 of byte 5 flipped: implicit rejection).
 * Prints K and K_rej; they must equal the reference's values quoted in SURVEY.md 8(c):
 *   K     = ca49ed38f11d513390bb0db10b9bf900eb6ce82f1ca0c71acca7947ad0dd2c37
 *   K_rej = 1ff209d0da6ec725d8513af357049d0cb065caa7fd3fd2b038aa4c2487e962b3
 * No CUDA header is needed: the ABI is plain pointers and sizes. */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "mlkem_b200.h"

static void hex(const char *name, const uint8_t *p, size_t n) {
    printf("%s = ", name);
    for (size_t i = 0; i < n; i++) printf("%02x", p[i]);
    printf("\n");
}

int main(void) {
    enum { SET = 768, N = 2 };
    uint8_t d[32], z[32], m[N * 32], K[N * 32], Kd[N * 32];
    unsigned ekb = mlkem_b200_ek_bytes(SET), dkb = mlkem_b200_dk_bytes(SET), cb = mlkem_b200_ct_bytes(SET);
    uint8_t *ek = malloc(ekb), *dk = malloc(dkb), *c = malloc((size_t)N * cb);
    for (int i = 0; i < 32; i++) {
        d[i] = (uint8_t)i;
        z[i] = (uint8_t)(32 + i);
        m[i] = m[32 + i] = (uint8_t)(64 + i);
    }
    int rc = mlkem_b200_keygen_batch(SET, 1, d, z, ek, dk, NULL); /* NULL opts: host memory, current device, blocking */
    if (rc) { fprintf(stderr, "keygen: rc=%d %s\n", rc, mlkem_b200_last_error()); return 1; }

    mlkem_b200_opts o = {-1, MLKEM_B200_MEM_HOST, NULL, 0, 0, MLKEM_B200_FLAG_EXPAND_KEYS};
    mlkem_b200_keys *keys = NULL;
    rc = mlkem_b200_keys_from_seeds(SET, 1, d, z, &o, &keys);
    if (rc) { fprintf(stderr, "keys_from_seeds: rc=%d %s\n", rc, mlkem_b200_last_error()); return 1; }

    uint32_t key_index[N] = {0, 0};
    rc = mlkem_b200_encaps_keyed_batch(keys, N, key_index, m, c, K, NULL);
    if (rc) { fprintf(stderr, "encaps: rc=%d %s\n", rc, mlkem_b200_last_error()); return 1; }
    c[cb + 5] ^= 1; /* the second ciphertext is tampered */
    rc = mlkem_b200_decaps_keyed_batch(keys, N, key_index, c, Kd, NULL);
    if (rc) { fprintf(stderr, "decaps: rc=%d %s\n", rc, mlkem_b200_last_error()); return 1; }

    /* the unkeyed call on the same key must agree */
    uint8_t K2[32];
    rc = mlkem_b200_decaps_batch(SET, 1, dk, c, K2, NULL);
    if (rc || memcmp(K2, Kd, 32) != 0 || memcmp(K, Kd, 32) != 0) { fprintf(stderr, "keyed and unkeyed results differ\n"); return 1; }
    hex("K    ", Kd, 32);
    hex("K_rej", Kd + 32, 32);
    mlkem_b200_keys_free(keys);
    free(ek); free(dk); free(c);
    return 0;
}
