/*
 * io_oracle.c -- CPU ORACLE file readers (TEST INFRASTRUCTURE ONLY).
 * Radiance RGBE (.hdr) -> linear float RGBA, the behaviour of stb_image's stbi_loadf(.., STBI_rgb_alpha)
 * that the reference calls at VulkanHelper/Source/Utility/AssetImporterImpl.cpp:494-503
 * (stb_image @ f58f558 is not vendored in /root/reference; the published algorithm is restated:
 * new-style RLE scanlines, value = mantissa * 2^(e-136), e==0 -> 0, alpha = 1).
 */
#include "pt_oracle.h"
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <math.h>

static int read_line(FILE *f, char *buf, int n) {
    int i = 0, c;
    while ((c = fgetc(f)) != EOF && c != '\n') { if (i < n - 1) buf[i++] = (char)c; }
    buf[i] = 0;
    return c != EOF || i > 0;
}

float *orc_load_hdr(const char *path, uint32_t *w_out, uint32_t *h_out) {
    FILE *f = fopen(path, "rb");
    if (!f) return NULL;
    char line[1024];
    if (!read_line(f, line, sizeof line) || (strcmp(line, "#?RADIANCE") != 0 && strcmp(line, "#?RGBE") != 0)) { fclose(f); return NULL; }
    int ok = 0;
    for (;;) {
        if (!read_line(f, line, sizeof line)) break;
        if (line[0] == 0) break;
        if (strcmp(line, "FORMAT=32-bit_rle_rgbe") == 0) ok = 1;
    }
    if (!ok) { fclose(f); return NULL; }
    if (!read_line(f, line, sizeof line)) { fclose(f); return NULL; }
    int W = 0, H = 0;
    if (sscanf(line, "-Y %d +X %d", &H, &W) != 2 || W <= 0 || H <= 0) { fclose(f); return NULL; }
    float *out = (float *)malloc(sizeof(float) * 4 * (size_t)W * H);
    unsigned char *scan = (unsigned char *)malloc((size_t)W * 4);
    for (int y = 0; y < H; y++) {
        unsigned char hd[4];
        if (fread(hd, 1, 4, f) != 4) goto fail;
        if (W < 8 || W >= 32768 || hd[0] != 2 || hd[1] != 2 || (hd[2] & 0x80)) {
            /* flat (non-RLE) scanline: first pixel already read */
            memcpy(scan, hd, 4);
            if (fread(scan + 4, 1, (size_t)(W - 1) * 4, f) != (size_t)(W - 1) * 4) goto fail;
        } else {
            if (((hd[2] << 8) | hd[3]) != W) goto fail;
            for (int k = 0; k < 4; k++) {
                int i = 0;
                while (i < W) {
                    int count = fgetc(f);
                    if (count == EOF) goto fail;
                    if (count > 128) { int v = fgetc(f); count -= 128; if (i + count > W) goto fail; for (int z = 0; z < count; z++) scan[(i++) * 4 + k] = (unsigned char)v; }
                    else { if (count == 0 || i + count > W) goto fail; for (int z = 0; z < count; z++) scan[(i++) * 4 + k] = (unsigned char)fgetc(f); }
                }
            }
        }
        for (int x = 0; x < W; x++) {
            const unsigned char *p = scan + x * 4;
            float *o = out + ((size_t)y * W + x) * 4;
            if (p[3] != 0) {
                float f1 = (float)ldexp(1.0, (int)p[3] - (128 + 8));
                o[0] = p[0] * f1; o[1] = p[1] * f1; o[2] = p[2] * f1;
            } else { o[0] = o[1] = o[2] = 0.0f; }
            o[3] = 1.0f;
        }
    }
    free(scan); fclose(f);
    *w_out = (uint32_t)W; *h_out = (uint32_t)H;
    return out;
fail:
    free(scan); free(out); fclose(f);
    return NULL;
}

void orc_free(void *p) { free(p); }
