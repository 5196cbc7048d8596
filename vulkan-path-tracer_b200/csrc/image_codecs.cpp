// image_codecs.cpp -- file codecs of the scene loader and the image-output API (host side, no GPU).
//
// The reference decodes textures with stb_image (stbi_load(.., STBI_rgb_alpha) / stbi_loadf) at
// VulkanHelper/Source/Utility/AssetImporterImpl.cpp:494-545 and writes PNGs with stbi_write_png at
// PathTracer/Editor.cpp:833-840.  stb is a network-fetched dependency (VulkanHelper/Libraries/CMakeLists.txt:58-62)
// that is not vendored, so the formats are decoded here from their public specifications:
//   PNG  (RFC 2083; zlib inflate via libz)      -> RGBA8, lossless: identical to any conforming decoder
//   JPEG (ITU T.81 baseline, Huffman, 8-bit)     -> RGBA8, integer "islow" IDCT + triangle chroma up-sampling + fixed-point
//                                                   YCbCr conversion, the same published algorithms stb_image uses
//   Radiance RGBE .hdr (RLE scanlines)           -> linear float RGBA, value = mantissa * 2^(e-136), alpha 1
#include "host_api.h"
#include <zlib.h>
#include <cstdio>
#include <cstring>
#include <cmath>
#include <string>
#include <vector>

namespace b200pt {

static bool read_file(const std::string &path, std::vector<uint8_t> &out) {
    FILE *f = fopen(path.c_str(), "rb");
    if (!f) return false;
    fseek(f, 0, SEEK_END); long n = ftell(f); fseek(f, 0, SEEK_SET);
    if (n < 0) { fclose(f); return false; }
    out.resize((size_t)n);
    size_t got = n ? fread(out.data(), 1, (size_t)n, f) : 0;
    fclose(f);
    return got == (size_t)n;
}

// =================================================================================== PNG decode
static inline uint32_t be32(const uint8_t *p) { return ((uint32_t)p[0] << 24) | ((uint32_t)p[1] << 16) | ((uint32_t)p[2] << 8) | p[3]; }
static inline int paeth(int a, int b, int c) { int p = a + b - c, pa = abs(p - a), pb = abs(p - b), pc = abs(p - c); return (pa <= pb && pa <= pc) ? a : (pb <= pc ? b : c); }

static bool decode_png(const std::vector<uint8_t> &d, uint32_t &W, uint32_t &H, std::vector<uint8_t> &rgba, std::string &err) {
    static const uint8_t sig[8] = { 0x89, 'P', 'N', 'G', 0x0D, 0x0A, 0x1A, 0x0A };
    if (d.size() < 8 || memcmp(d.data(), sig, 8) != 0) { err = "not a PNG"; return false; }
    size_t pos = 8;
    int depth = 0, ctype = 0, interlace = 0;
    std::vector<uint8_t> idat, plte, trns;
    bool have_ihdr = false;
    while (pos + 8 <= d.size()) {
        uint32_t len = be32(&d[pos]); const uint8_t *type = &d[pos + 4];
        if (pos + 12 + (size_t)len > d.size()) { err = "truncated PNG"; return false; }
        const uint8_t *data = &d[pos + 8];
        if (!memcmp(type, "IHDR", 4)) {
            if (len < 13) { err = "bad IHDR"; return false; }
            W = be32(data); H = be32(data + 4); depth = data[8]; ctype = data[9]; interlace = data[12]; have_ihdr = true;
        } else if (!memcmp(type, "PLTE", 4)) plte.assign(data, data + len);
        else if (!memcmp(type, "tRNS", 4)) trns.assign(data, data + len);
        else if (!memcmp(type, "IDAT", 4)) idat.insert(idat.end(), data, data + len);
        else if (!memcmp(type, "IEND", 4)) break;
        pos += 12 + (size_t)len;
    }
    if (!have_ihdr || W == 0 || H == 0) { err = "missing IHDR"; return false; }
    if (W > 65535u || H > 65535u || (uint64_t)W * H > (1ull << 28)) { err = "PNG dimensions out of range"; return false; }   // texture side limit of the reference's image path
    if (interlace > 1) { err = "unknown PNG interlace method"; return false; }
    int chans = ctype == 0 ? 1 : ctype == 2 ? 3 : ctype == 3 ? 1 : ctype == 4 ? 2 : ctype == 6 ? 4 : 0;
    if (!chans || !(depth == 1 || depth == 2 || depth == 4 || depth == 8 || depth == 16)) { err = "unsupported PNG format"; return false; }
    const size_t bpp_bits = (size_t)chans * depth, bpp = (bpp_bits + 7) / 8;
    auto stride_of = [&](uint32_t w) { return ((size_t)w * bpp_bits + 7) / 8; };
    // Adam7 (PNG spec 8.2): seven reduced images, each filtered on its own; interlace == 0 is the single full-size "pass"
    static const int xs[7] = { 0, 4, 0, 2, 0, 1, 0 }, ys[7] = { 0, 0, 4, 0, 2, 0, 1 }, dxs[7] = { 8, 8, 4, 4, 2, 2, 1 }, dys[7] = { 8, 8, 8, 4, 4, 2, 2 };
    struct Pass { uint32_t w, h; int x0, y0, dx, dy; };
    std::vector<Pass> passes;
    if (!interlace) passes.push_back({ W, H, 0, 0, 1, 1 });
    else for (int p = 0; p < 7; p++) {
        const uint32_t pw = (W > (uint32_t)xs[p]) ? (W - xs[p] + dxs[p] - 1) / dxs[p] : 0, ph = (H > (uint32_t)ys[p]) ? (H - ys[p] + dys[p] - 1) / dys[p] : 0;
        if (pw && ph) passes.push_back({ pw, ph, xs[p], ys[p], dxs[p], dys[p] });
    }
    size_t raw_size = 0; for (const Pass &ps : passes) raw_size += (stride_of(ps.w) + 1) * (size_t)ps.h;
    std::vector<uint8_t> raw(raw_size);
    uLongf rawlen = (uLongf)raw.size();
    int zr = uncompress(raw.data(), &rawlen, idat.data(), (uLong)idat.size());
    if (zr != Z_OK || rawlen != raw.size()) { err = "PNG inflate failed"; return false; }
    rgba.resize((size_t)W * H * 4);
    auto sample = [&](const uint8_t *row, size_t idx) -> int {   // idx = sample index within the row
        if (depth == 8) return row[idx];
        if (depth == 16) return row[idx * 2];                      // high byte (stb's 16->8 conversion)
        size_t bit = idx * depth; int v = (row[bit >> 3] >> (8 - depth - (bit & 7))) & ((1 << depth) - 1);
        return v;
    };
    const int scale = depth < 8 ? (depth == 1 ? 255 : depth == 2 ? 85 : 17) : 1;
    size_t raw_off = 0;
    std::vector<uint8_t> img;
    for (const Pass &ps : passes) {
        const size_t stride = stride_of(ps.w);
        img.assign(stride * (size_t)ps.h, 0);
        for (uint32_t y = 0; y < ps.h; y++) {
            const uint8_t *src = &raw[raw_off + (stride + 1) * (size_t)y]; uint8_t ft = src[0]; src++;
            uint8_t *cur = &img[stride * (size_t)y]; const uint8_t *prev = y ? &img[stride * (size_t)(y - 1)] : nullptr;
            for (size_t i = 0; i < stride; i++) {
                int a = i >= bpp ? cur[i - bpp] : 0, b = prev ? prev[i] : 0, c = (prev && i >= bpp) ? prev[i - bpp] : 0, x = src[i];
                switch (ft) {
                case 0: break;
                case 1: x += a; break;
                case 2: x += b; break;
                case 3: x += (a + b) >> 1; break;
                case 4: x += paeth(a, b, c); break;
                default: err = "bad PNG filter"; return false;
                }
                cur[i] = (uint8_t)x;
            }
        }
        raw_off += (stride + 1) * (size_t)ps.h;
        for (uint32_t y = 0; y < ps.h; y++) {
            const uint8_t *row = &img[stride * (size_t)y];
            for (uint32_t x = 0; x < ps.w; x++) {
                uint8_t *o = &rgba[((size_t)(ps.y0 + (int)y * ps.dy) * W + (size_t)(ps.x0 + (int)x * ps.dx)) * 4];
                switch (ctype) {
                case 0: { int g = sample(row, x); int g8 = depth < 8 ? g * scale : g; o[0] = o[1] = o[2] = (uint8_t)g8; o[3] = 255;
                          if (trns.size() >= 2) { int t = depth == 16 ? trns[0] : ((trns[0] << 8) | trns[1]); if (depth == 16 ? (row[x * 2] == trns[0] && row[x * 2 + 1] == trns[1]) : g == t) o[3] = 0; } } break;
                case 2: { o[0] = (uint8_t)sample(row, x * 3); o[1] = (uint8_t)sample(row, x * 3 + 1); o[2] = (uint8_t)sample(row, x * 3 + 2); o[3] = 255;
                          if (trns.size() >= 6 && depth == 8 && o[0] == trns[1] && o[1] == trns[3] && o[2] == trns[5]) o[3] = 0; } break;
                case 3: { int i = sample(row, x); if ((size_t)i * 3 + 2 >= plte.size()) { err = "PNG palette index out of range"; return false; }
                          o[0] = plte[i * 3]; o[1] = plte[i * 3 + 1]; o[2] = plte[i * 3 + 2]; o[3] = (size_t)i < trns.size() ? trns[i] : 255; } break;
                case 4: { int g = sample(row, x * 2); o[0] = o[1] = o[2] = (uint8_t)g; o[3] = (uint8_t)sample(row, x * 2 + 1); } break;
                case 6: { o[0] = (uint8_t)sample(row, x * 4); o[1] = (uint8_t)sample(row, x * 4 + 1); o[2] = (uint8_t)sample(row, x * 4 + 2); o[3] = (uint8_t)sample(row, x * 4 + 3); } break;
                }
            }
        }
    }
    return true;
}

// =================================================================================== PNG encode (Editor::SaveToFile)
static void put_be32(std::vector<uint8_t> &v, uint32_t x) { v.push_back(x >> 24); v.push_back(x >> 16); v.push_back(x >> 8); v.push_back(x); }
static void put_chunk(std::vector<uint8_t> &out, const char *type, const std::vector<uint8_t> &data) {
    put_be32(out, (uint32_t)data.size());
    size_t s = out.size();
    out.insert(out.end(), type, type + 4); out.insert(out.end(), data.begin(), data.end());
    put_be32(out, (uint32_t)crc32(0L, &out[s], (uInt)(out.size() - s)));
}
bool write_png_rgba8(const std::string &path, uint32_t W, uint32_t H, const uint8_t *rgba) {
    std::vector<uint8_t> out = { 0x89, 'P', 'N', 'G', 0x0D, 0x0A, 0x1A, 0x0A };
    std::vector<uint8_t> ihdr; put_be32(ihdr, W); put_be32(ihdr, H); ihdr.insert(ihdr.end(), { 8, 6, 0, 0, 0 });
    put_chunk(out, "IHDR", ihdr);
    std::vector<uint8_t> raw(((size_t)W * 4 + 1) * H);
    for (uint32_t y = 0; y < H; y++) { raw[((size_t)W * 4 + 1) * y] = 0; memcpy(&raw[((size_t)W * 4 + 1) * y + 1], rgba + (size_t)y * W * 4, (size_t)W * 4); }
    uLongf clen = compressBound((uLong)raw.size());
    std::vector<uint8_t> comp(clen);
    if (compress2(comp.data(), &clen, raw.data(), (uLong)raw.size(), 6) != Z_OK) return false;
    comp.resize(clen);
    put_chunk(out, "IDAT", comp);
    put_chunk(out, "IEND", {});
    FILE *f = fopen(path.c_str(), "wb");
    if (!f) return false;
    bool ok = fwrite(out.data(), 1, out.size(), f) == out.size();
    fclose(f);
    return ok;
}

// =================================================================================== JPEG decode (baseline + progressive, 8-bit Huffman)
namespace {
struct Huff { uint8_t size[257]; uint16_t code[256]; uint8_t values[256]; int maxcode[18]; int delta[17]; uint8_t fast[512]; };
struct Comp { int id, h, v, tq, td, ta, dc_pred; int w2, h2; std::vector<uint8_t> data; int x = 0, y = 0, coeff_w = 0; std::vector<short> coeff; };   // x, y: size in samples; coeff: progressive scans
struct Jpeg {
    const uint8_t *p, *end; uint32_t W = 0, H = 0; Huff dc[4], ac[4]; uint16_t dq[4][64]; Comp c[3]; int ncomp = 0;
    int hmax = 1, vmax = 1, mcux = 0, mcuy = 0, restart = 0;
    uint32_t bits = 0; int nbits = 0; bool marker_hit = false; int todo = 0; std::string err;
    bool progressive = false; int spec_start = 0, spec_end = 63, succ_high = 0, succ_low = 0, eob_run = 0;   // SOS parameters of a progressive scan
    int scan_n = 0, order[3] = { 0, 1, 2 };
};
const uint8_t kZig[64 + 15] = { 0, 1, 8, 16, 9, 2, 3, 10, 17, 24, 32, 25, 18, 11, 4, 5, 12, 19, 26, 33, 40, 48, 41, 34, 27, 20, 13, 6, 7, 14, 21, 28, 35, 42, 49, 56, 57, 50,
                                43, 36, 29, 22, 15, 23, 30, 37, 44, 51, 58, 59, 52, 45, 38, 31, 39, 46, 53, 60, 61, 54, 47, 55, 62, 63, 63, 63, 63, 63, 63, 63, 63, 63, 63, 63, 63, 63, 63, 63, 63 };

bool build_huff(Huff &h, const uint8_t *counts) {
    { int total = 0; for (int i = 0; i < 16; i++) total += counts[i]; if (total > 256) return false; }   // a corrupt DHT must not run past the 256-symbol tables
    int k = 0;
    for (int i = 0; i < 16; i++) for (int j = 0; j < counts[i]; j++) h.size[k++] = (uint8_t)(i + 1);
    h.size[k] = 0;
    int code = 0; k = 0;
    for (int j = 1; j <= 16; j++) {
        h.delta[j] = k - code;
        if (h.size[k] == j) { while (h.size[k] == j) h.code[k++] = (uint16_t)(code++); if (code - 1 >= (1 << j)) return false; }
        h.maxcode[j] = code << (16 - j);
        code <<= 1;
    }
    h.maxcode[17] = 0x7fffffff;
    memset(h.fast, 255, sizeof(h.fast));
    for (int i = 0; i < k; i++) { int s = h.size[i]; if (s <= 9) { int c = h.code[i] << (9 - s), m = 1 << (9 - s); for (int j = 0; j < m; j++) h.fast[c + j] = (uint8_t)i; } }
    return true;
}
void grow_bits(Jpeg &j) {
    while (j.nbits <= 24) {
        int b = 0;
        if (!j.marker_hit && j.p < j.end) {
            b = *j.p++;
            if (b == 0xFF) { int c = j.p < j.end ? *j.p : 0; if (c == 0) j.p++; else { j.marker_hit = true; j.p--; b = 0; } }
        }
        j.bits |= (uint32_t)b << (24 - j.nbits); j.nbits += 8;
    }
}
int huff_decode(Jpeg &j, const Huff &h) {
    if (j.nbits < 16) grow_bits(j);
    int c = (j.bits >> 23) & 511, k = h.fast[c];
    if (k < 255) { int s = h.size[k]; if (s > j.nbits) return -1; j.bits <<= s; j.nbits -= s; return h.values[k]; }
    uint32_t temp = j.bits >> 16; int s;
    for (s = 10;; s++) if ((int)temp < h.maxcode[s]) break;
    if (s == 17 || s > j.nbits) return -1;
    c = (int)((j.bits >> (32 - s)) & ((1u << s) - 1)) + h.delta[s];
    if (c < 0 || c > 255) return -1;
    j.bits <<= s; j.nbits -= s;
    return h.values[c];
}
int extend_receive(Jpeg &j, int n) {
    if (n == 0) return 0;
    if (j.nbits < n) grow_bits(j);
    int sgn = (int)(j.bits >> 31);
    uint32_t k = (j.bits << n) | (j.bits >> (32 - n));   // rotate left
    j.bits = k & ~((1u << n) - 1u);
    k &= (1u << n) - 1u;
    j.nbits -= n;
    return (int)k + ((sgn ? 0 : 1) * (int)((~0u << n) + 1));   // k + (sgn ? 0 : (-1<<n)+1)
}
bool decode_block(Jpeg &j, short data[64], const Huff &hdc, const Huff &hac, int b, const uint16_t *dq) {
    if (j.nbits < 16) grow_bits(j);
    int t = huff_decode(j, hdc);
    if (t < 0 || t > 15) { j.err = "bad huffman code"; return false; }
    memset(data, 0, 64 * sizeof(short));
    int diff = t ? extend_receive(j, t) : 0;
    int dc = j.c[b].dc_pred + diff; j.c[b].dc_pred = dc;
    data[0] = (short)(dc * dq[0]);
    int k = 1;
    do {
        int rs = huff_decode(j, hac);
        if (rs < 0) { j.err = "bad huffman code"; return false; }
        int s = rs & 15, r = rs >> 4;
        if (s == 0) { if (rs != 0xF0) break; k += 16; }
        else { k += r; int zig = kZig[k++]; data[zig] = (short)(extend_receive(j, s) * dq[zig]); }
    } while (k < 64);
    return true;
}
// ---- progressive scans (ITU T.81 annex G): DC first / refinement, AC first / refinement with end-of-band runs
int get_bits(Jpeg &j, int n) {
    if (n == 0) return 0;
    if (j.nbits < n) grow_bits(j);
    uint32_t k = (j.bits << n) | (j.bits >> (32 - n));
    j.bits = k & ~((1u << n) - 1u);
    k &= (1u << n) - 1u;
    j.nbits -= n;
    return (int)k;
}
inline int get_bit(Jpeg &j) { if (j.nbits < 1) grow_bits(j); const uint32_t k = j.bits; j.bits <<= 1; --j.nbits; return (int)(k >> 31); }
bool decode_block_prog_dc(Jpeg &j, short data[64], const Huff &hdc, int b) {
    if (j.spec_end != 0) { j.err = "can't merge dc and ac"; return false; }
    if (j.nbits < 16) grow_bits(j);
    if (j.succ_high == 0) {                                                 // first scan of the DC coefficient
        memset(data, 0, 64 * sizeof(short));
        const int t = huff_decode(j, hdc);
        if (t < 0 || t > 15) { j.err = "bad huffman code"; return false; }
        const int diff = t ? extend_receive(j, t) : 0;
        const int dc = j.c[b].dc_pred + diff; j.c[b].dc_pred = dc;
        data[0] = (short)(dc * (1 << j.succ_low));
    } else if (get_bit(j)) data[0] = (short)(data[0] + (short)(1 << j.succ_low));   // refinement: one more bit
    return true;
}
bool decode_block_prog_ac(Jpeg &j, short data[64], const Huff &hac) {
    if (j.spec_start == 0) { j.err = "can't merge dc and ac"; return false; }
    if (j.succ_high == 0) {                                                 // first scan of this band
        const int shift = j.succ_low;
        if (j.eob_run) { --j.eob_run; return true; }
        int k = j.spec_start;
        do {
            if (j.nbits < 16) grow_bits(j);
            const int rs = huff_decode(j, hac);
            if (rs < 0) { j.err = "bad huffman code"; return false; }
            const int s = rs & 15; int r = rs >> 4;
            if (s == 0) {
                if (r < 15) { j.eob_run = (1 << r); if (r) j.eob_run += get_bits(j, r); --j.eob_run; break; }
                k += 16;
            } else {
                k += r;
                const int zig = kZig[k++];
                data[zig] = (short)(extend_receive(j, s) * (1 << shift));
            }
        } while (k <= j.spec_end);
    } else {                                                                // refinement of a band
        const short bit = (short)(1 << j.succ_low);
        if (j.eob_run) {
            --j.eob_run;
            for (int k = j.spec_start; k <= j.spec_end; ++k) {
                short *p = &data[kZig[k]];
                if (*p != 0 && get_bit(j) && (*p & bit) == 0) *p = (short)(*p > 0 ? *p + bit : *p - bit);
            }
        } else {
            int k = j.spec_start;
            do {
                if (j.nbits < 16) grow_bits(j);
                const int rs = huff_decode(j, hac);
                if (rs < 0) { j.err = "bad huffman code"; return false; }
                int s = rs & 15, r = rs >> 4;
                if (s == 0) {
                    if (r < 15) { j.eob_run = (1 << r) - 1; if (r) j.eob_run += get_bits(j, r); r = 64; }   // end of band: only refinements remain in this block
                } else {
                    if (s != 1) { j.err = "bad huffman code"; return false; }
                    s = get_bit(j) ? bit : -bit;
                }
                while (k <= j.spec_end) {
                    short *p = &data[kZig[k++]];
                    if (*p != 0) { if (get_bit(j) && (*p & bit) == 0) *p = (short)(*p > 0 ? *p + bit : *p - bit); }
                    else { if (r == 0) { *p = (short)s; break; } --r; }
                }
            } while (k <= j.spec_end);
        }
    }
    return true;
}
inline uint8_t clamp8(long long x) { if (x < 0) return 0; if (x > 255) return 255; return (uint8_t)x; }
#define F2F(x) ((int)(((x) * 4096 + 0.5)))
#define FSH(x) ((x) * 4096)
// 64-bit temporaries: identical results for every valid stream (nothing there exceeds 31 bits), no signed overflow on corrupt ones
#define IDCT_1D(s0, s1, s2, s3, s4, s5, s6, s7) \
    long long t0, t1, t2, t3, p1, p2, p3, p4, p5, x0, x1, x2, x3; \
    p2 = s2; p3 = s6; p1 = (p2 + p3) * F2F(0.5411961f); t2 = p1 + p3 * F2F(-1.847759065f); t3 = p1 + p2 * F2F(0.765366865f); \
    p2 = s0; p3 = s4; t0 = FSH(p2 + p3); t1 = FSH(p2 - p3); x0 = t0 + t3; x3 = t0 - t3; x1 = t1 + t2; x2 = t1 - t2; \
    t0 = s7; t1 = s5; t2 = s3; t3 = s1; p3 = t0 + t2; p4 = t1 + t3; p1 = t0 + t3; p2 = t1 + t2; p5 = (p3 + p4) * F2F(1.175875602f); \
    t0 = t0 * F2F(0.298631336f); t1 = t1 * F2F(2.053119869f); t2 = t2 * F2F(3.072711026f); t3 = t3 * F2F(1.501321110f); \
    p1 = p5 + p1 * F2F(-0.899976223f); p2 = p5 + p2 * F2F(-2.562915447f); p3 = p3 * F2F(-1.961570560f); p4 = p4 * F2F(-0.390180644f); \
    t3 += p1 + p4; t2 += p2 + p3; t1 += p2 + p4; t0 += p1 + p3;
void idct_block(uint8_t *out, int stride, short data[64]) {
    int val[64], *v = val; short *d = data;
    for (int i = 0; i < 8; ++i, ++d, ++v) {
        if (d[8] == 0 && d[16] == 0 && d[24] == 0 && d[32] == 0 && d[40] == 0 && d[48] == 0 && d[56] == 0) {
            int dcterm = d[0] * 4; v[0] = v[8] = v[16] = v[24] = v[32] = v[40] = v[48] = v[56] = dcterm;
        } else {
            IDCT_1D(d[0], d[8], d[16], d[24], d[32], d[40], d[48], d[56])
            x0 += 512; x1 += 512; x2 += 512; x3 += 512;
            v[0] = (int)((x0 + t3) >> 10); v[56] = (int)((x0 - t3) >> 10); v[8] = (int)((x1 + t2) >> 10); v[48] = (int)((x1 - t2) >> 10);
            v[16] = (int)((x2 + t1) >> 10); v[40] = (int)((x2 - t1) >> 10); v[24] = (int)((x3 + t0) >> 10); v[32] = (int)((x3 - t0) >> 10);
        }
    }
    v = val; uint8_t *o = out;
    for (int i = 0; i < 8; ++i, v += 8, o += stride) {
        IDCT_1D(v[0], v[1], v[2], v[3], v[4], v[5], v[6], v[7])
        x0 += 65536 + (128 << 17); x1 += 65536 + (128 << 17); x2 += 65536 + (128 << 17); x3 += 65536 + (128 << 17);
        o[0] = clamp8((x0 + t3) >> 17); o[7] = clamp8((x0 - t3) >> 17); o[1] = clamp8((x1 + t2) >> 17); o[6] = clamp8((x1 - t2) >> 17);
        o[2] = clamp8((x2 + t1) >> 17); o[5] = clamp8((x2 - t1) >> 17); o[3] = clamp8((x3 + t0) >> 17); o[4] = clamp8((x3 - t0) >> 17);
    }
}
// chroma up-sampling rows (triangle filters)
void resample_1(uint8_t *out, const uint8_t *near_, const uint8_t *, int w, int) { memcpy(out, near_, (size_t)w); }
void resample_v2(uint8_t *out, const uint8_t *n, const uint8_t *f, int w, int) { for (int i = 0; i < w; i++) out[i] = (uint8_t)((3 * n[i] + f[i] + 2) >> 2); }
void resample_h2(uint8_t *out, const uint8_t *in, const uint8_t *, int w, int) {
    if (w == 1) { out[0] = out[1] = in[0]; return; }
    out[0] = in[0]; out[1] = (uint8_t)((in[0] * 3 + in[1] + 2) >> 2);
    int i;
    for (i = 1; i < w - 1; i++) { int n = 3 * in[i] + 2; out[i * 2] = (uint8_t)((n + in[i - 1]) >> 2); out[i * 2 + 1] = (uint8_t)((n + in[i + 1]) >> 2); }
    out[i * 2] = (uint8_t)((in[w - 2] * 3 + in[w - 1] + 2) >> 2); out[i * 2 + 1] = in[w - 1];
}
void resample_hv2(uint8_t *out, const uint8_t *n, const uint8_t *f, int w, int) {
    if (w == 1) { out[0] = out[1] = (uint8_t)((3 * n[0] + f[0] + 2) >> 2); return; }
    int t0, t1 = 3 * n[0] + f[0];
    out[0] = (uint8_t)((t1 + 2) >> 2);
    int i;
    for (i = 1; i < w; i++) { t0 = t1; t1 = 3 * n[i] + f[i]; out[i * 2 - 1] = (uint8_t)((3 * t0 + t1 + 8) >> 4); out[i * 2] = (uint8_t)((3 * t1 + t0 + 8) >> 4); }
    out[w * 2 - 1] = (uint8_t)((t1 + 2) >> 2);
}
void resample_generic(uint8_t *out, const uint8_t *n, const uint8_t *, int w, int hs) { for (int i = 0; i < w; i++) for (int k = 0; k < hs; k++) out[i * hs + k] = n[i]; }
typedef void (*ResampleFn)(uint8_t *, const uint8_t *, const uint8_t *, int, int);

bool jpeg_decode(Jpeg &j, std::vector<uint8_t> &rgba) {
    auto u16 = [&](const uint8_t *q) { return (q[0] << 8) | q[1]; };
    if (j.end - j.p < 4 || j.p[0] != 0xFF || j.p[1] != 0xD8) { j.err = "not a JPEG"; return false; }
    j.p += 2;
    bool sof = false, sos = false;
    while (j.p + 4 <= j.end && !sos) {
        if (j.p[0] != 0xFF) { j.p++; continue; }
        int m = j.p[1]; j.p += 2;
        if (m == 0xFF || m == 0x00 || (m >= 0xD0 && m <= 0xD9)) continue;
        int L = u16(j.p); const uint8_t *d = j.p + 2, *de = j.p + L;
        if (de > j.end) { j.err = "truncated JPEG"; return false; }
        if (m == 0xDB) { while (d < de) { int q = *d++; int p16 = q >> 4, t = q & 15; if (t > 3) { j.err = "bad DQT"; return false; } for (int i = 0; i < 64; i++) { j.dq[t][kZig[i]] = (uint16_t)(p16 ? u16(d) : *d); d += p16 ? 2 : 1; } } }
        else if (m == 0xC4) { while (d < de) { int q = *d++; int tc = q >> 4, th = q & 15; if (tc > 1 || th > 3) { j.err = "bad DHT"; return false; } if (d + 16 > de) { j.err = "truncated DHT"; return false; } const uint8_t *counts = d; int n = 0; for (int i = 0; i < 16; i++) n += counts[i]; d += 16; if (n > 256 || d + n > de) { j.err = "bad DHT"; return false; } Huff &h = tc ? j.ac[th] : j.dc[th]; if (!build_huff(h, counts)) { j.err = "bad huffman table"; return false; } memcpy(h.values, d, (size_t)n); d += n; } }
        else if (m == 0xC0 || m == 0xC1 || m == 0xC2) {
            if (d[0] != 8) { j.err = "only 8-bit JPEG"; return false; }
            j.progressive = (m == 0xC2);
            j.H = (uint32_t)u16(d + 1); j.W = (uint32_t)u16(d + 3); j.ncomp = d[5];
            if (j.ncomp != 1 && j.ncomp != 3) { j.err = "unsupported JPEG component count"; return false; }
            for (int i = 0; i < j.ncomp; i++) { Comp &c = j.c[i]; c.id = d[6 + i * 3]; c.h = d[7 + i * 3] >> 4; c.v = d[7 + i * 3] & 15; c.tq = d[8 + i * 3]; if (!c.h || !c.v || c.h > 4 || c.v > 4 || c.tq > 3) { j.err = "bad SOF"; return false; } j.hmax = std::max(j.hmax, c.h); j.vmax = std::max(j.vmax, c.v); }
            if (!j.W || !j.H) { j.err = "JPEG without size"; return false; }
            if ((uint64_t)j.W * j.H > (1ull << 28)) { j.err = "JPEG dimensions out of range"; return false; }
            const int mcuw = j.hmax * 8, mcuh = j.vmax * 8;
            j.mcux = ((int)j.W + mcuw - 1) / mcuw; j.mcuy = ((int)j.H + mcuh - 1) / mcuh;
            for (int i = 0; i < j.ncomp; i++) {
                Comp &c = j.c[i]; c.w2 = j.mcux * c.h * 8; c.h2 = j.mcuy * c.v * 8; c.data.assign((size_t)c.w2 * c.h2, 0); c.dc_pred = 0;
                c.x = ((int)j.W * c.h + j.hmax - 1) / j.hmax; c.y = ((int)j.H * c.v + j.vmax - 1) / j.vmax;
                if (j.progressive) { c.coeff_w = c.w2 / 8; c.coeff.assign((size_t)c.w2 * c.h2, 0); }
            }
            sof = true;
        }
        else if (m == 0xDD) j.restart = u16(d);
        else if (m == 0xDA) {
            const int ns = d[0];
            if (!sof || ns < 1 || ns > j.ncomp || (!j.progressive && ns != j.ncomp)) { j.err = "unsupported SOS"; return false; }
            for (int i = 0; i < ns; i++) {
                const int id = d[1 + i * 2], q = d[2 + i * 2]; int which = -1;
                for (int k = 0; k < j.ncomp; k++) if (j.c[k].id == id) which = k;
                if (which < 0 || (!j.progressive && which != i)) { j.err = "bad SOS order"; return false; }
                j.c[which].td = q >> 4; j.c[which].ta = q & 15; j.order[i] = which;
                if (j.c[which].td > 3 || j.c[which].ta > 3) { j.err = "bad SOS tables"; return false; }
            }
            j.scan_n = ns;
            j.spec_start = d[1 + ns * 2]; j.spec_end = d[2 + ns * 2]; j.succ_high = d[3 + ns * 2] >> 4; j.succ_low = d[3 + ns * 2] & 15;
            if (j.progressive) {
                if (j.spec_start > 63 || j.spec_end > 63 || j.spec_start > j.spec_end || j.succ_high > 13 || j.succ_low > 13) { j.err = "bad SOS"; return false; }
                // ---- one progressive scan: entropy-coded data starts right after the header
                j.p = de; j.bits = 0; j.nbits = 0; j.marker_hit = false; j.eob_run = 0;
                for (int n = 0; n < j.ncomp; n++) j.c[n].dc_pred = 0;
                j.todo = j.restart ? j.restart : 0x7fffffff;
                auto restart_check = [&]() {
                    if (--j.todo <= 0) {
                        if (j.nbits < 24) grow_bits(j);
                        if (j.marker_hit && j.p + 1 < j.end && j.p[0] == 0xFF && j.p[1] >= 0xD0 && j.p[1] <= 0xD7) j.p += 2;
                        j.bits = 0; j.nbits = 0; j.marker_hit = false; j.eob_run = 0; for (int n = 0; n < j.ncomp; n++) j.c[n].dc_pred = 0;
                        j.todo = j.restart ? j.restart : 0x7fffffff;
                    }
                };
                if (ns == 1) {                                              // non-interleaved: the component's own block grid
                    const int n = j.order[0]; Comp &c = j.c[n];
                    const int bw = (c.x + 7) >> 3, bh = (c.y + 7) >> 3;
                    for (int by = 0; by < bh; by++) for (int bx = 0; bx < bw; bx++) {
                        short *data = &c.coeff[64 * ((size_t)bx + (size_t)by * c.coeff_w)];
                        if (j.spec_start == 0) { if (!decode_block_prog_dc(j, data, j.dc[c.td], n)) return false; }
                        else if (!decode_block_prog_ac(j, data, j.ac[c.ta])) return false;
                        restart_check();
                    }
                } else {                                                    // interleaved: DC scans only
                    for (int my = 0; my < j.mcuy; my++) for (int mx = 0; mx < j.mcux; mx++) {
                        for (int k = 0; k < ns; k++) { const int n = j.order[k]; Comp &c = j.c[n];
                            for (int y = 0; y < c.v; y++) for (int x = 0; x < c.h; x++) {
                                short *data = &c.coeff[64 * ((size_t)(mx * c.h + x) + (size_t)(my * c.v + y) * c.coeff_w)];
                                if (!decode_block_prog_dc(j, data, j.dc[c.td], n)) return false;
                            } }
                        restart_check();
                    }
                }
                // move to the next marker (the bit reader stops in front of it; otherwise scan forward for 0xFF xx)
                if (!j.marker_hit) { while (j.p + 1 < j.end && !(j.p[0] == 0xFF && j.p[1] != 0x00 && !(j.p[1] >= 0xD0 && j.p[1] <= 0xD7))) j.p++; }
                j.bits = 0; j.nbits = 0; j.marker_hit = false;
                sos = false;                                                // keep reading markers: more scans / tables follow
                if (j.p + 1 < j.end && j.p[0] == 0xFF && j.p[1] == 0xD9) break;
                continue;
            }
            sos = true;
        }
        else if (m == 0xD9) break;
        j.p = de;
    }
    if (!sof || !j.W || !j.H) { j.err = "JPEG without frame"; return false; }
    if (j.progressive) {
        // finish: dequantise and inverse-transform every block of every component
        for (int n = 0; n < j.ncomp; n++) { Comp &c = j.c[n];
            const int bw = (c.x + 7) >> 3, bh = (c.y + 7) >> 3;
            for (int by = 0; by < bh; by++) for (int bx = 0; bx < bw; bx++) {
                short *data = &c.coeff[64 * ((size_t)bx + (size_t)by * c.coeff_w)];
                for (int k = 0; k < 64; k++) data[k] = (short)(data[k] * j.dq[c.tq][k]);
                idct_block(&c.data[(size_t)c.w2 * by * 8 + (size_t)bx * 8], c.w2, data);
            } }
    } else {
        if (!sos) { j.err = "JPEG without scan"; return false; }
        j.todo = j.restart ? j.restart : 0x7fffffff;
        short block[64];
        for (int my = 0; my < j.mcuy; my++) for (int mx = 0; mx < j.mcux; mx++) {
            for (int n = 0; n < j.ncomp; n++) { Comp &c = j.c[n];
                for (int y = 0; y < c.v; y++) for (int x = 0; x < c.h; x++) {
                    int x2 = (mx * c.h + x) * 8, y2 = (my * c.v + y) * 8;
                    if (!decode_block(j, block, j.dc[c.td], j.ac[c.ta], n, j.dq[c.tq])) return false;
                    idct_block(&c.data[(size_t)c.w2 * y2 + x2], c.w2, block);
                } }
            if (--j.todo <= 0) {
                if (j.nbits < 24) grow_bits(j);
                // restart marker
                if (j.marker_hit && j.p + 1 < j.end && j.p[0] == 0xFF && j.p[1] >= 0xD0 && j.p[1] <= 0xD7) j.p += 2;
                j.bits = 0; j.nbits = 0; j.marker_hit = false; for (int n = 0; n < j.ncomp; n++) j.c[n].dc_pred = 0;
                j.todo = j.restart ? j.restart : 0x7fffffff;
            }
        }
    }
    // resample + colour convert
    rgba.resize((size_t)j.W * j.H * 4);
    struct RS { ResampleFn fn; const uint8_t *line0, *line1; int hs, vs, ystep, w_lores, ypos; };
    RS rs[3]; std::vector<uint8_t> linebuf[3];
    for (int k = 0; k < j.ncomp; k++) {
        RS &r = rs[k]; Comp &c = j.c[k];
        linebuf[k].assign((size_t)j.W + 3 + 64, 0);
        r.hs = j.hmax / c.h; r.vs = j.vmax / c.v; r.ystep = r.vs >> 1; r.w_lores = ((int)j.W + r.hs - 1) / r.hs; r.ypos = 0; r.line0 = r.line1 = c.data.data();
        if (r.hs == 1 && r.vs == 1) r.fn = resample_1; else if (r.hs == 1 && r.vs == 2) r.fn = resample_v2; else if (r.hs == 2 && r.vs == 1) r.fn = resample_h2;
        else if (r.hs == 2 && r.vs == 2) r.fn = resample_hv2; else r.fn = resample_generic;
    }
    const int comp_h[3] = { ((int)j.H * j.c[0].v + j.vmax - 1) / j.vmax, j.ncomp == 3 ? ((int)j.H * j.c[1].v + j.vmax - 1) / j.vmax : 0, j.ncomp == 3 ? ((int)j.H * j.c[2].v + j.vmax - 1) / j.vmax : 0 };
    for (uint32_t y = 0; y < j.H; y++) {
        const uint8_t *co[3] = { nullptr, nullptr, nullptr };
        for (int k = 0; k < j.ncomp; k++) {
            RS &r = rs[k]; Comp &c = j.c[k];
            bool y_bot = r.ystep >= (r.vs >> 1);
            const uint8_t *nr = y_bot ? r.line1 : r.line0, *fr = y_bot ? r.line0 : r.line1;
            if (r.fn == resample_1) co[k] = nr;
            else { r.fn(linebuf[k].data(), nr, fr, r.w_lores, r.hs); co[k] = linebuf[k].data(); }
            if (++r.ystep >= r.vs) { r.ystep = 0; r.line0 = r.line1; if (++r.ypos < comp_h[k]) r.line1 += c.w2; }
        }
        uint8_t *o = &rgba[(size_t)y * j.W * 4];
        if (j.ncomp == 3) {
            #define F2X(x) (((int)((x) * 4096.0f + 0.5f)) << 8)
            for (uint32_t i = 0; i < j.W; i++) {
                int yf = (co[0][i] << 20) + (1 << 19), cr = co[2][i] - 128, cb = co[1][i] - 128;
                int r = yf + cr * F2X(1.40200f);
                int g = yf + (cr * -F2X(0.71414f)) + ((cb * -F2X(0.34414f)) & 0xffff0000);
                int b = yf + cb * F2X(1.77200f);
                r >>= 20; g >>= 20; b >>= 20;
                o[i * 4] = clamp8(r); o[i * 4 + 1] = clamp8(g); o[i * 4 + 2] = clamp8(b); o[i * 4 + 3] = 255;
            }
        } else for (uint32_t i = 0; i < j.W; i++) { o[i * 4] = o[i * 4 + 1] = o[i * 4 + 2] = co[0][i]; o[i * 4 + 3] = 255; }
    }
    return true;
}
} // namespace

bool decode_image_rgba8(const std::string &path, uint32_t &W, uint32_t &H, std::vector<uint8_t> &rgba, std::string &err) {
    std::vector<uint8_t> d;
    if (!read_file(path, d)) { err = "cannot read " + path; return false; }
    if (d.size() >= 8 && d[0] == 0x89 && d[1] == 'P') return decode_png(d, W, H, rgba, err);
    if (d.size() >= 4 && d[0] == 0xFF && d[1] == 0xD8) {
        Jpeg *j = new Jpeg(); j->p = d.data(); j->end = d.data() + d.size();
        bool ok = jpeg_decode(*j, rgba); W = j->W; H = j->H; if (!ok) err = j->err; delete j; return ok;
    }
    err = "unsupported image format: " + path;
    return false;
}

// =================================================================================== Radiance RGBE
bool decode_hdr_rgba32f(const std::string &path, uint32_t &W, uint32_t &H, std::vector<float> &rgba, std::string &err) {
    std::vector<uint8_t> d;
    if (!read_file(path, d)) { err = "cannot read " + path; return false; }
    size_t pos = 0;
    auto line = [&](std::string &s) { s.clear(); while (pos < d.size() && d[pos] != '\n') s.push_back((char)d[pos++]); if (pos < d.size()) pos++; return true; };
    std::string s; line(s);
    if (s != "#?RADIANCE" && s != "#?RGBE") { err = "not a Radiance HDR file"; return false; }
    bool fmt = false;
    for (;;) { if (pos >= d.size()) break; line(s); if (s.empty()) break; if (s == "FORMAT=32-bit_rle_rgbe") fmt = true; }
    if (!fmt) { err = "unsupported HDR format"; return false; }
    line(s);
    int w = 0, h = 0;
    if (sscanf(s.c_str(), "-Y %d +X %d", &h, &w) != 2 || w <= 0 || h <= 0) { err = "unsupported HDR orientation"; return false; }
    W = (uint32_t)w; H = (uint32_t)h;
    rgba.resize((size_t)w * h * 4);
    std::vector<uint8_t> scan((size_t)w * 4);
    for (int y = 0; y < h; y++) {
        if (pos + 4 > d.size()) { err = "truncated HDR"; return false; }
        if (w < 8 || w >= 32768 || d[pos] != 2 || d[pos + 1] != 2 || (d[pos + 2] & 0x80)) {
            if (pos + (size_t)w * 4 > d.size()) { err = "truncated HDR"; return false; }
            memcpy(scan.data(), &d[pos], (size_t)w * 4); pos += (size_t)w * 4;
        } else {
            if (((d[pos + 2] << 8) | d[pos + 3]) != w) { err = "corrupt HDR scanline"; return false; }
            pos += 4;
            for (int k = 0; k < 4; k++) {
                int i = 0;
                while (i < w) {
                    if (pos >= d.size()) { err = "truncated HDR"; return false; }
                    int count = d[pos++];
                    if (count > 128) { count -= 128; if (i + count > w || pos >= d.size()) { err = "corrupt HDR"; return false; } uint8_t v = d[pos++]; for (int z = 0; z < count; z++) scan[(size_t)(i++) * 4 + k] = v; }
                    else { if (count == 0 || i + count > w || pos + count > d.size()) { err = "corrupt HDR"; return false; } for (int z = 0; z < count; z++) scan[(size_t)(i++) * 4 + k] = d[pos++]; }
                }
            }
        }
        for (int x = 0; x < w; x++) {
            const uint8_t *p = &scan[(size_t)x * 4]; float *o = &rgba[((size_t)y * w + x) * 4];
            if (p[3] != 0) { float f1 = (float)ldexp(1.0, (int)p[3] - (128 + 8)); o[0] = p[0] * f1; o[1] = p[1] * f1; o[2] = p[2] * f1; }
            else o[0] = o[1] = o[2] = 0.0f;
            o[3] = 1.0f;
        }
    }
    return true;
}

} // namespace b200pt
