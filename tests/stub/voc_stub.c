/* TEST INFRASTRUCTURE: a fake vocoder backend with the C-ABI symbols the native server
 * (csrc/voc_server.cpp) calls, LD_PRELOADed in front of libvoc_b200.so so that the server's
 * protocol / batching logic can be exercised on a box without a GPU.  The "audio" is a pure
 * integer hash of the request's codes and the sample index; lengths come from the real library's
 * voc_plan (host-only).  A code outside [0, 2048) is an error, as in the real backend. */
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#define LC 122325LL
int voc_plan(int max_tokens, long long chunk_samples, int n_tokens, int meta_cap, int* meta, long long* total,
             int* pairwise);                                   /* from libvoc_b200.so (no GPU needed) */

static char g_err[128] = "";

void* voc_create_from_file(const char* path, int device, int wave) {
    (void)device; (void)wave;
    FILE* f = fopen(path, "rb");
    if (!f) { snprintf(g_err, sizeof g_err, "%s: cannot open", path); return NULL; }
    fclose(f);
    return malloc(1);
}
void voc_destroy(void* h) { free(h); }
int voc_max_tokens(void* h) { (void)h; return 64; }
const char* voc_last_error(void* h) { (void)h; return g_err; }

long long voc_out_samples(void* h, int n) {
    (void)h;
    long long total = 0; int pw = 0;
    if (voc_plan(64, LC, n, 0, NULL, &total, &pw) < 0) return -1;
    return total;
}

static uint32_t mix(uint32_t x) { x ^= x >> 16; x *= 0x7feb352dU; x ^= x >> 15; x *= 0x846ca68bU; x ^= x >> 16; return x; }

static int fake(const long long* codes, int n, short* out, long long total) {
    uint32_t seed = 2166136261U;
    for (long long i = 0; i < (long long)n * 16; ++i) {
        if (codes[i] < 0 || codes[i] >= 2048) { snprintf(g_err, sizeof g_err, "audio code outside [0, codebook_size)"); return -1; }
        seed = (seed ^ (uint32_t)codes[i]) * 16777619U;
    }
    for (long long i = 0; i < total; ++i) out[i] = (short)(mix(seed + (uint32_t)i) & 0xffff);
    return 0;
}

int voc_synthesize_pcm16(void* h, const long long* codes, int n, short* out, long long cap, long long* n_out) {
    const long long total = voc_out_samples(h, n);
    if (total < 0 || total > cap) return -1;
    if (fake(codes, n, out, total)) return -1;
    *n_out = total;
    return 0;
}

int voc_synthesize_batch_pcm16(void* h, const long long* codes, const int* n_tokens, int n_requests, short* out,
                               long long cap, long long* out_offsets) {
    long long off = 0, frame = 0;
    for (int u = 0; u < n_requests; ++u) {
        const long long total = voc_out_samples(h, n_tokens[u]);
        if (total < 0 || off + total > cap) return -1;
        out_offsets[u] = off;
        if (fake(codes + frame * 16, n_tokens[u], out + off, total)) return -1;
        off += total; frame += n_tokens[u];
    }
    out_offsets[n_requests] = off;
    return 0;
}
