/* Reads one tensor out of a pips_b200 weight pack (pips_b200/pack.py) from plain C: what a C caller of
 * include/pips_b200.h does to fill pips_weights without torch.  File layout (little endian):
 *   [0,16) magic "PIPSB200PACK\0\0\0\1" | [16,24) u64 manifest length M | [24,32) u64 data origin D |
 *   [32,32+M) JSON manifest {"fingerprint":..,"abi":..,"tensors":[{"name":..,"dtype":..,"shape":[..],"offset":..,"nbytes":..},..]}
 *   tensor bytes at D + offset (D, offset multiples of 256: cudaMemcpy straight from the mapping).
 * The manifest is written with fixed key order and no whitespace, so a substring scan is enough here.
 *
 *   gcc -std=c99 -Wall -Wextra -Werror examples/pack_read.c -o /tmp/pack_read
 *   /tmp/pack_read model-000200000.pack mixer.layer3.fc1_w_hi
 */
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

typedef struct {
    const unsigned char* data;   /* first byte of the tensor inside the loaded file */
    uint64_t nbytes;
    char dtype[16];
    char shape[64];              /* "[2048,512]" */
} pack_tensor;

static int pack_find(const unsigned char* file, uint64_t size, const char* name, pack_tensor* out) {
    static const unsigned char magic[16] = {'P', 'I', 'P', 'S', 'B', '2', '0', '0', 'P', 'A', 'C', 'K', 0, 0, 0, 1};
    uint64_t mlen, origin;
    if (size < 32 || memcmp(file, magic, 16) != 0) return -1;
    memcpy(&mlen, file + 16, 8);
    memcpy(&origin, file + 24, 8);
    if (32 + mlen > size || origin > size) return -1;
    char* man = (char*)malloc(mlen + 1);
    if (!man) return -1;
    memcpy(man, file + 32, mlen);
    man[mlen] = 0;
    char key[160];
    snprintf(key, sizeof key, "{\"name\":\"%s\",", name);
    int rc = -2;
    const char* e = strstr(man, key);
    if (e) {
        const char* dt = strstr(e, "\"dtype\":\"");
        const char* sh = strstr(e, "\"shape\":");
        const char* of = strstr(e, "\"offset\":");
        const char* nb = strstr(e, "\"nbytes\":");
        if (dt && sh && of && nb) {
            size_t n = strcspn(dt + 9, "\"");
            if (n >= sizeof out->dtype) n = sizeof out->dtype - 1;
            memcpy(out->dtype, dt + 9, n);
            out->dtype[n] = 0;
            n = strcspn(sh + 8, "]") + 1;
            if (n >= sizeof out->shape) n = sizeof out->shape - 1;
            memcpy(out->shape, sh + 8, n);
            out->shape[n] = 0;
            const uint64_t off = strtoull(of + 9, NULL, 10);
            out->nbytes = strtoull(nb + 9, NULL, 10);
            if (origin + off + out->nbytes <= size && (off % 256) == 0) {
                out->data = file + origin + off;
                rc = 0;
            } else {
                rc = -3;
            }
        }
    }
    free(man);
    return rc;
}

int main(int argc, char** argv) {
    if (argc != 3) {
        fprintf(stderr, "usage: %s file.pack tensor.name\n", argv[0]);
        return 2;
    }
    FILE* f = fopen(argv[1], "rb");
    if (!f) { perror("open"); return 1; }
    fseek(f, 0, SEEK_END);
    const long size = ftell(f);
    fseek(f, 0, SEEK_SET);
    unsigned char* buf = (unsigned char*)malloc((size_t)size);
    if (!buf || fread(buf, 1, (size_t)size, f) != (size_t)size) { fprintf(stderr, "read failed\n"); return 1; }
    fclose(f);
    pack_tensor t;
    const int rc = pack_find(buf, (uint64_t)size, argv[2], &t);
    if (rc != 0) {
        fprintf(stderr, "%s: %s\n", argv[2], rc == -1 ? "not a weight pack" : rc == -2 ? "no such tensor" : "tensor outside the file");
        return 1;
    }
    uint64_t h = 1469598103934665603ull;                      /* FNV-1a over the tensor bytes */
    for (uint64_t i = 0; i < t.nbytes; ++i) h = (h ^ t.data[i]) * 1099511628211ull;
    printf("%s dtype=%s shape=%s nbytes=%llu fnv1a=%016llx\n", argv[2], t.dtype, t.shape, (unsigned long long)t.nbytes,
           (unsigned long long)h);
    free(buf);
    return 0;
}
