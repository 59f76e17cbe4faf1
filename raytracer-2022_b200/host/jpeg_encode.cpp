/*
 * jpeg_encode.cpp -- baseline JPEG writer for the output stage (main.rs:213-221: `image::DynamicImage::ImageRgb8(img)
 * .write_to(&mut file, image::ImageOutputFormat::Jpeg(quality))` with quality = 100).
 *
 * The crate (image 0.23.14, Cargo.lock:211) is not vendored, so this restates the PUBLISHED algorithms its encoder is built
 * from -- parity with the crate's bytes is UNPINNED:
 *   - quantisation tables: ITU-T T.81 Annex K.1 / K.2 scaled with the IJG rule (quality 100 -> every divisor 1);
 *   - JFIF RGB -> YCbCr, three components with 1x1 sampling (4:4:4), one 8x8 block per component per MCU, right and bottom
 *     edges padded by replication;
 *   - forward DCT: the "slow but accurate" integer DCT of the IJG library (jfdctint: 13-bit constants, results scaled by 8);
 *   - Huffman coding with the Annex K.3 tables, byte stuffing, JFIF APP0 header.
 * tests/test_host_and_abi.py checks that the stream decodes (own decoder and libjpeg-turbo through PIL) to within 2 levels
 * of the input at quality 100.  Host code, not on the hot path.
 */
#include "rt_host.h"

#include <cmath>
#include <cstring>

namespace rt {

namespace {

const uint8_t kStdLuma[64] = {16, 11, 10, 16, 24, 40, 51, 61, 12, 12, 14, 19, 26, 58, 60, 55, 14, 13, 16, 24, 40, 57, 69, 56,
                              14, 17, 22, 29, 51, 87, 80, 62, 18, 22, 37, 56, 68, 109, 103, 77, 24, 35, 55, 64, 81, 104, 113, 92,
                              49, 64, 78, 87, 103, 121, 120, 101, 72, 92, 95, 98, 112, 100, 103, 99};
const uint8_t kStdChroma[64] = {17, 18, 24, 47, 99, 99, 99, 99, 18, 21, 26, 66, 99, 99, 99, 99, 24, 26, 56, 99, 99, 99, 99, 99,
                                47, 66, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99,
                                99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99};
const uint8_t kZigzag[64] = {0, 1, 8, 16, 9, 2, 3, 10, 17, 24, 32, 25, 18, 11, 4, 5, 12, 19, 26, 33, 40, 48, 41, 34, 27, 20, 13, 6, 7, 14, 21, 28,
                             35, 42, 49, 56, 57, 50, 43, 36, 29, 22, 15, 23, 30, 37, 44, 51, 58, 59, 52, 45, 38, 31, 39, 46, 53, 60, 61, 54, 47, 55, 62, 63};
/* Annex K.3: BITS (codes per length 1..16) and HUFFVAL */
const uint8_t kDcLumaBits[16] = {0, 1, 5, 1, 1, 1, 1, 1, 1, 0, 0, 0, 0, 0, 0, 0};
const uint8_t kDcChromaBits[16] = {0, 3, 1, 1, 1, 1, 1, 1, 1, 1, 1, 0, 0, 0, 0, 0};
const uint8_t kDcVals[12] = {0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11};
const uint8_t kAcLumaBits[16] = {0, 2, 1, 3, 3, 2, 4, 3, 5, 5, 4, 4, 0, 0, 1, 0x7d};
const uint8_t kAcLumaVals[162] = {
    0x01, 0x02, 0x03, 0x00, 0x04, 0x11, 0x05, 0x12, 0x21, 0x31, 0x41, 0x06, 0x13, 0x51, 0x61, 0x07, 0x22, 0x71, 0x14, 0x32, 0x81, 0x91, 0xa1,
    0x08, 0x23, 0x42, 0xb1, 0xc1, 0x15, 0x52, 0xd1, 0xf0, 0x24, 0x33, 0x62, 0x72, 0x82, 0x09, 0x0a, 0x16, 0x17, 0x18, 0x19, 0x1a, 0x25, 0x26,
    0x27, 0x28, 0x29, 0x2a, 0x34, 0x35, 0x36, 0x37, 0x38, 0x39, 0x3a, 0x43, 0x44, 0x45, 0x46, 0x47, 0x48, 0x49, 0x4a, 0x53, 0x54, 0x55, 0x56,
    0x57, 0x58, 0x59, 0x5a, 0x63, 0x64, 0x65, 0x66, 0x67, 0x68, 0x69, 0x6a, 0x73, 0x74, 0x75, 0x76, 0x77, 0x78, 0x79, 0x7a, 0x83, 0x84, 0x85,
    0x86, 0x87, 0x88, 0x89, 0x8a, 0x92, 0x93, 0x94, 0x95, 0x96, 0x97, 0x98, 0x99, 0x9a, 0xa2, 0xa3, 0xa4, 0xa5, 0xa6, 0xa7, 0xa8, 0xa9, 0xaa,
    0xb2, 0xb3, 0xb4, 0xb5, 0xb6, 0xb7, 0xb8, 0xb9, 0xba, 0xc2, 0xc3, 0xc4, 0xc5, 0xc6, 0xc7, 0xc8, 0xc9, 0xca, 0xd2, 0xd3, 0xd4, 0xd5, 0xd6,
    0xd7, 0xd8, 0xd9, 0xda, 0xe1, 0xe2, 0xe3, 0xe4, 0xe5, 0xe6, 0xe7, 0xe8, 0xe9, 0xea, 0xf1, 0xf2, 0xf3, 0xf4, 0xf5, 0xf6, 0xf7, 0xf8, 0xf9,
    0xfa};
const uint8_t kAcChromaBits[16] = {0, 2, 1, 2, 4, 4, 3, 4, 7, 5, 4, 4, 0, 1, 2, 0x77};
const uint8_t kAcChromaVals[162] = {
    0x00, 0x01, 0x02, 0x03, 0x11, 0x04, 0x05, 0x21, 0x31, 0x06, 0x12, 0x41, 0x51, 0x07, 0x61, 0x71, 0x13, 0x22, 0x32, 0x81, 0x08, 0x14, 0x42,
    0x91, 0xa1, 0xb1, 0xc1, 0x09, 0x23, 0x33, 0x52, 0xf0, 0x15, 0x62, 0x72, 0xd1, 0x0a, 0x16, 0x24, 0x34, 0xe1, 0x25, 0xf1, 0x17, 0x18, 0x19,
    0x1a, 0x26, 0x27, 0x28, 0x29, 0x2a, 0x35, 0x36, 0x37, 0x38, 0x39, 0x3a, 0x43, 0x44, 0x45, 0x46, 0x47, 0x48, 0x49, 0x4a, 0x53, 0x54, 0x55,
    0x56, 0x57, 0x58, 0x59, 0x5a, 0x63, 0x64, 0x65, 0x66, 0x67, 0x68, 0x69, 0x6a, 0x73, 0x74, 0x75, 0x76, 0x77, 0x78, 0x79, 0x7a, 0x82, 0x83,
    0x84, 0x85, 0x86, 0x87, 0x88, 0x89, 0x8a, 0x92, 0x93, 0x94, 0x95, 0x96, 0x97, 0x98, 0x99, 0x9a, 0xa2, 0xa3, 0xa4, 0xa5, 0xa6, 0xa7, 0xa8,
    0xa9, 0xaa, 0xb2, 0xb3, 0xb4, 0xb5, 0xb6, 0xb7, 0xb8, 0xb9, 0xba, 0xc2, 0xc3, 0xc4, 0xc5, 0xc6, 0xc7, 0xc8, 0xc9, 0xca, 0xd2, 0xd3, 0xd4,
    0xd5, 0xd6, 0xd7, 0xd8, 0xd9, 0xda, 0xe2, 0xe3, 0xe4, 0xe5, 0xe6, 0xe7, 0xe8, 0xe9, 0xea, 0xf2, 0xf3, 0xf4, 0xf5, 0xf6, 0xf7, 0xf8, 0xf9,
    0xfa};

struct HuffEnc {
    uint16_t code[256];
    uint8_t size[256];
};
HuffEnc build_enc(const uint8_t bits[16], const uint8_t *vals) { /* T.81 Annex C */
    HuffEnc e;
    std::memset(&e, 0, sizeof e);
    int k = 0;
    uint16_t code = 0;
    for (int len = 1; len <= 16; ++len) {
        for (int i = 0; i < bits[len - 1]; ++i) {
            e.code[vals[k]] = code++;
            e.size[vals[k]] = (uint8_t)len;
            ++k;
        }
        code <<= 1;
    }
    return e;
}

struct BitWriter {
    std::vector<uint8_t> &out;
    uint32_t acc = 0;
    int n = 0;
    explicit BitWriter(std::vector<uint8_t> &o) : out(o) {}
    void put(uint32_t code, int size) {
        acc = (acc << size) | (code & ((1u << size) - 1u));
        n += size;
        while (n >= 8) {
            uint8_t b = (uint8_t)(acc >> (n - 8));
            out.push_back(b);
            if (b == 0xff) out.push_back(0); /* byte stuffing */
            n -= 8;
        }
    }
    void flush() {
        if (n > 0) put(0x7f, 8 - n); /* pad with ones */
    }
};

/* jfdctint (IJG "islow"): in place on 64 level-shifted samples; output scaled up by 8 */
void fdct_islow(int *data) {
    const int CB = 13, P1 = 2;
    const int c0298 = 2446, c0390 = 3196, c0541 = 4433, c0765 = 6270, c0899 = 7373, c1175 = 9633, c1501 = 12299, c1847 = 15137, c1961 = 16069,
              c2053 = 16819, c2562 = 20995, c3072 = 25172;
    auto descale = [](long long x, int n) { return (int)((x + (1ll << (n - 1))) >> n); };
    for (int pass = 0; pass < 2; ++pass) {
        for (int k = 0; k < 8; ++k) {
            int *p = pass == 0 ? data + 8 * k : data + k;
            const int st = pass == 0 ? 1 : 8;
            long long t0 = p[0] + p[7 * st], t7 = p[0] - p[7 * st], t1 = p[st] + p[6 * st], t6 = p[st] - p[6 * st];
            long long t2 = p[2 * st] + p[5 * st], t5 = p[2 * st] - p[5 * st], t3 = p[3 * st] + p[4 * st], t4 = p[3 * st] - p[4 * st];
            long long t10 = t0 + t3, t13 = t0 - t3, t11 = t1 + t2, t12 = t1 - t2;
            if (pass == 0) {
                p[0] = (int)((t10 + t11) << P1);
                p[4 * st] = (int)((t10 - t11) << P1);
            } else {
                p[0] = descale(t10 + t11, P1);
                p[4 * st] = descale(t10 - t11, P1);
            }
            const int sh = pass == 0 ? CB - P1 : CB + P1;
            long long z1 = (t12 + t13) * c0541;
            p[2 * st] = descale(z1 + t13 * c0765, sh);
            p[6 * st] = descale(z1 - t12 * c1847, sh);
            z1 = t4 + t7;
            long long z2 = t5 + t6, z3 = t4 + t6, z4 = t5 + t7, z5 = (z3 + z4) * c1175;
            t4 *= c0298; t5 *= c2053; t6 *= c3072; t7 *= c1501;
            z1 *= -c0899; z2 *= -c2562; z3 *= -c1961; z4 *= -c0390;
            z3 += z5; z4 += z5;
            p[7 * st] = descale(t4 + z1 + z3, sh);
            p[5 * st] = descale(t5 + z2 + z4, sh);
            p[3 * st] = descale(t6 + z2 + z3, sh);
            p[st] = descale(t7 + z1 + z4, sh);
        }
    }
}

void seg(std::vector<uint8_t> &o, uint8_t marker, const std::vector<uint8_t> &payload) {
    o.push_back(0xff);
    o.push_back(marker);
    size_t len = payload.size() + 2;
    o.push_back((uint8_t)(len >> 8));
    o.push_back((uint8_t)len);
    o.insert(o.end(), payload.begin(), payload.end());
}

int bit_size(int v) {
    int a = v < 0 ? -v : v, n = 0;
    while (a) { ++n; a >>= 1; }
    return n;
}

} // namespace

bool encode_jpeg_mem(const uint8_t *rgb, int w, int h, int quality, std::vector<uint8_t> &out, std::string &err) {
    if (!rgb || w <= 0 || h <= 0 || w > 65535 || h > 65535) {
        err = "bad image size for JPEG (1..65535)";
        return false;
    }
    if (quality < 1) quality = 1;
    if (quality > 100) quality = 100;
    const int scale = quality < 50 ? 5000 / quality : 200 - 2 * quality; /* IJG quality scaling */
    uint8_t qt[2][64];
    for (int t = 0; t < 2; ++t)
        for (int i = 0; i < 64; ++i) {
            int v = ((t == 0 ? kStdLuma[i] : kStdChroma[i]) * scale + 50) / 100;
            qt[t][i] = (uint8_t)(v < 1 ? 1 : (v > 255 ? 255 : v));
        }
    out.clear();
    out.push_back(0xff);
    out.push_back(0xd8); /* SOI */
    seg(out, 0xe0, {'J', 'F', 'I', 'F', 0, 1, 2, 0, 0, 1, 0, 1, 0, 0}); /* JFIF 1.02, aspect 1:1, no thumbnail */
    for (int t = 0; t < 2; ++t) {
        std::vector<uint8_t> p(65);
        p[0] = (uint8_t)t; /* 8-bit precision, table id */
        for (int i = 0; i < 64; ++i) p[1 + i] = qt[t][kZigzag[i]];
        seg(out, 0xdb, p);
    }
    seg(out, 0xc0, {8, (uint8_t)(h >> 8), (uint8_t)h, (uint8_t)(w >> 8), (uint8_t)w, 3, 1, 0x11, 0, 2, 0x11, 1, 3, 0x11, 1}); /* SOF0, 4:4:4 */
    auto dht = [&](uint8_t tc_th, const uint8_t bits[16], const uint8_t *vals, int nvals) {
        std::vector<uint8_t> p;
        p.push_back(tc_th);
        p.insert(p.end(), bits, bits + 16);
        p.insert(p.end(), vals, vals + nvals);
        seg(out, 0xc4, p);
    };
    dht(0x00, kDcLumaBits, kDcVals, 12);
    dht(0x10, kAcLumaBits, kAcLumaVals, 162);
    dht(0x01, kDcChromaBits, kDcVals, 12);
    dht(0x11, kAcChromaBits, kAcChromaVals, 162);
    seg(out, 0xda, {3, 1, 0x00, 2, 0x11, 3, 0x11, 0, 63, 0}); /* SOS */
    const HuffEnc dc[2] = {build_enc(kDcLumaBits, kDcVals), build_enc(kDcChromaBits, kDcVals)};
    const HuffEnc ac[2] = {build_enc(kAcLumaBits, kAcLumaVals), build_enc(kAcChromaBits, kAcChromaVals)};
    BitWriter bw(out);
    int pred[3] = {0, 0, 0};
    for (int by = 0; by < h; by += 8)
        for (int bx = 0; bx < w; bx += 8) {
            int blk[3][64];
            for (int y = 0; y < 8; ++y)
                for (int x = 0; x < 8; ++x) {
                    int sy = by + y < h ? by + y : h - 1, sx = bx + x < w ? bx + x : w - 1; /* edge replication */
                    const uint8_t *px = rgb + 3 * ((size_t)sy * (size_t)w + (size_t)sx);
                    float r = px[0], g = px[1], b = px[2];
                    float yy = 0.299f * r + 0.587f * g + 0.114f * b;
                    float cb = -0.1687f * r - 0.3313f * g + 0.5f * b + 128.f;
                    float cr = 0.5f * r - 0.4187f * g - 0.0813f * b + 128.f;
                    auto q8 = [](float v) { int i = (int)(v + 0.5f); return i < 0 ? 0 : (i > 255 ? 255 : i); };
                    blk[0][8 * y + x] = q8(yy) - 128;
                    blk[1][8 * y + x] = q8(cb) - 128;
                    blk[2][8 * y + x] = q8(cr) - 128;
                }
            for (int c = 0; c < 3; ++c) {
                fdct_islow(blk[c]);
                const int t = c == 0 ? 0 : 1;
                int zz[64];
                for (int i = 0; i < 64; ++i) {
                    int v = blk[c][kZigzag[i]], q = qt[t][kZigzag[i]] * 8; /* the DCT output carries a factor 8 */
                    zz[i] = v < 0 ? -((-v + q / 2) / q) : (v + q / 2) / q;
                }
                int diff = zz[0] - pred[c];
                pred[c] = zz[0];
                int s = bit_size(diff);
                bw.put(dc[t].code[s], dc[t].size[s]);
                if (s) bw.put((uint32_t)(diff < 0 ? diff - 1 : diff), s);
                int run = 0;
                for (int i = 1; i < 64; ++i) {
                    if (zz[i] == 0) {
                        ++run;
                        continue;
                    }
                    while (run > 15) {
                        bw.put(ac[t].code[0xf0], ac[t].size[0xf0]); /* ZRL */
                        run -= 16;
                    }
                    int sz = bit_size(zz[i]);
                    int sym = (run << 4) | sz;
                    bw.put(ac[t].code[sym], ac[t].size[sym]);
                    bw.put((uint32_t)(zz[i] < 0 ? zz[i] - 1 : zz[i]), sz);
                    run = 0;
                }
                if (run) bw.put(ac[t].code[0x00], ac[t].size[0x00]); /* EOB */
            }
        }
    bw.flush();
    out.push_back(0xff);
    out.push_back(0xd9); /* EOI */
    return true;
}

bool encode_jpeg(const std::string &path, const uint8_t *rgb, int w, int h, int quality, std::string &err) {
    std::vector<uint8_t> out;
    if (!encode_jpeg_mem(rgb, w, h, quality, out, err)) return false;
    FILE *f = std::fopen(path.c_str(), "wb");
    if (!f) {
        err = "Outputting image fails."; /* main.rs:220 */
        return false;
    }
    size_t n = std::fwrite(out.data(), 1, out.size(), f);
    std::fclose(f);
    if (n != out.size()) {
        err = "short write";
        return false;
    }
    return true;
}

} // namespace rt
