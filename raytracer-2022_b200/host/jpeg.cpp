/*
 * jpeg.cpp -- baseline (SOF0/SOF1, 8-bit, Huffman) JPEG -> RGB8 decoder for
 * ImageTexture::new (texture/mod.rs:89-107), which upstream is
 *     image::open(filename).unwrap()          (image 0.23.14 -> jpeg-decoder 0.1.22)
 * Neither crate is vendored under /root/reference (Cargo.lock:211,242), so this is a
 * restatement of the public algorithm: integer 8x8 IDCT in 12-bit fixed point (the
 * stb_image-style kernel jpeg-decoder ports), JFIF YCbCr -> RGB, and libjpeg-style "fancy"
 * triangle upsampling for 2x chroma (h2v1 / h2v2), replication otherwise.  All assets under
 * source/ are baseline; earthmap.jpg is 4:4:4.  Rounding can differ from the crate by
 * +-1..2 LSB: parity unpinned, bounded by tests/test_host_and_abi.py::test_jpeg_decoder_against_libjpeg / _420_ against libjpeg-turbo (PIL).
 */
#include "rt_host.h"

#include <cstring>
#include <fstream>

namespace rt {
namespace {

struct Huff {
    // canonical decoding tables (ITU T.81 F.2.2.3)
    int mincode[17], maxcode[18], valptr[17];
    uint8_t vals[256];
    uint16_t fast[512]; // 9-bit lookup: (len << 8) | symbol, 0 = miss
    bool present = false;
    void build(const uint8_t counts[16], const uint8_t *symbols, int n) {
        std::memcpy(vals, symbols, (size_t)n);
        int code = 0, k = 0;
        std::memset(fast, 0, sizeof fast);
        for (int len = 1; len <= 16; ++len) {
            valptr[len] = k;
            mincode[len] = code;
            for (int i = 0; i < counts[len - 1]; ++i, ++k, ++code) {
                if (len <= 9) {
                    int lo = code << (9 - len), hi = lo + (1 << (9 - len));
                    for (int f = lo; f < hi; ++f) fast[f] = (uint16_t)((len << 8) | vals[k]);
                }
            }
            maxcode[len] = counts[len - 1] ? code - 1 : -1;
            code <<= 1;
        }
        maxcode[17] = 0x7fffffff;
        present = true;
    }
};

struct BitReader {
    const uint8_t *p, *end;
    uint32_t acc = 0;
    int nbits = 0;
    bool hit_marker = false;
    void fill() {
        while (nbits <= 24) {
            int byte = 0;
            if (!hit_marker && p < end) {
                byte = *p++;
                if (byte == 0xff) {
                    int nxt = p < end ? *p : 0xd9;
                    if (nxt == 0) {
                        ++p; // stuffed zero
                    } else {
                        hit_marker = true; // leave the marker for the caller; feed zeros
                        --p;
                        byte = 0;
                    }
                }
            }
            acc |= (uint32_t)byte << (24 - nbits);
            nbits += 8;
        }
    }
    int peek(int n) {
        if (nbits < n) fill();
        return (int)(acc >> (32 - n));
    }
    void skip(int n) {
        acc <<= n;
        nbits -= n;
    }
    int get(int n) {
        if (n == 0) return 0;
        int v = peek(n);
        skip(n);
        return v;
    }
    void reset() {
        acc = 0;
        nbits = 0;
        hit_marker = false;
    }
};

inline int extend(int v, int n) { return v < (1 << (n - 1)) ? v - (1 << n) + 1 : v; }

int decode_symbol(BitReader &br, const Huff &h) {
    int look = br.peek(9);
    uint16_t f = h.fast[look];
    if (f) {
        br.skip(f >> 8);
        return f & 0xff;
    }
    int code = br.peek(16);
    for (int len = 10; len <= 16; ++len) {
        int c = code >> (16 - len);
        if (h.maxcode[len] >= 0 && c <= h.maxcode[len] && c >= h.mincode[len]) {
            br.skip(len);
            return h.vals[h.valptr[len] + c - h.mincode[len]];
        }
    }
    return -1;
}

const uint8_t ZIGZAG[64] = {0,  1,  8,  16, 9,  2,  3,  10, 17, 24, 32, 25, 18, 11, 4,  5,  12, 19, 26, 33, 40, 48,
                            41, 34, 27, 20, 13, 6,  7,  14, 21, 28, 35, 42, 49, 56, 57, 50, 43, 36, 29, 22, 15, 23,
                            30, 37, 44, 51, 58, 59, 52, 45, 38, 31, 39, 46, 53, 60, 61, 54, 47, 55, 62, 63};

inline int f2f(double x) { return (int)(x * 4096 + 0.5); }
inline uint8_t clamp8(int x) { return (uint8_t)(x < 0 ? 0 : (x > 255 ? 255 : x)); }

// 1-D pass of the 12-bit fixed-point IDCT (Loeffler-style factorisation as in stb_image)
#define IDCT_1D(s0, s1, s2, s3, s4, s5, s6, s7)                                                                    \
    int t0, t1, t2, t3, p1, p2, p3, p4, p5, x0, x1, x2, x3;                                                        \
    p2 = s2;                                                                                                       \
    p3 = s6;                                                                                                       \
    p1 = (p2 + p3) * f2f(0.5411961);                                                                               \
    t2 = p1 + p3 * f2f(-1.847759065);                                                                              \
    t3 = p1 + p2 * f2f(0.765366865);                                                                               \
    p2 = s0;                                                                                                       \
    p3 = s4;                                                                                                       \
    t0 = (p2 + p3) * 4096;                                                                                         \
    t1 = (p2 - p3) * 4096;                                                                                         \
    x0 = t0 + t3;                                                                                                  \
    x3 = t0 - t3;                                                                                                  \
    x1 = t1 + t2;                                                                                                  \
    x2 = t1 - t2;                                                                                                  \
    t0 = s7;                                                                                                       \
    t1 = s5;                                                                                                       \
    t2 = s3;                                                                                                       \
    t3 = s1;                                                                                                       \
    p3 = t0 + t2;                                                                                                  \
    p4 = t1 + t3;                                                                                                  \
    p1 = t0 + t3;                                                                                                  \
    p2 = t1 + t2;                                                                                                  \
    p5 = (p3 + p4) * f2f(1.175875602);                                                                             \
    t0 = t0 * f2f(0.298631336);                                                                                    \
    t1 = t1 * f2f(2.053119869);                                                                                    \
    t2 = t2 * f2f(3.072711026);                                                                                    \
    t3 = t3 * f2f(1.501321110);                                                                                    \
    p1 = p5 + p1 * f2f(-0.899976223);                                                                              \
    p2 = p5 + p2 * f2f(-2.562915447);                                                                              \
    p3 = p3 * f2f(-1.961570560);                                                                                   \
    p4 = p4 * f2f(-0.390180644);                                                                                   \
    t3 += p1 + p4;                                                                                                 \
    t2 += p2 + p3;                                                                                                 \
    t1 += p2 + p4;                                                                                                 \
    t0 += p1 + p3;

void idct_block(uint8_t *out, int out_stride, const int data[64]) {
    int val[64], *v = val;
    const int *d = data;
    for (int i = 0; i < 8; ++i, ++d, ++v) {
        if (d[8] == 0 && d[16] == 0 && d[24] == 0 && d[32] == 0 && d[40] == 0 && d[48] == 0 && d[56] == 0) {
            int dcterm = d[0] * 4;
            v[0] = v[8] = v[16] = v[24] = v[32] = v[40] = v[48] = v[56] = dcterm;
        } else {
            IDCT_1D(d[0], d[8], d[16], d[24], d[32], d[40], d[48], d[56])
            x0 += 512;
            x1 += 512;
            x2 += 512;
            x3 += 512;
            v[0] = (x0 + t3) >> 10;
            v[56] = (x0 - t3) >> 10;
            v[8] = (x1 + t2) >> 10;
            v[48] = (x1 - t2) >> 10;
            v[16] = (x2 + t1) >> 10;
            v[40] = (x2 - t1) >> 10;
            v[24] = (x3 + t0) >> 10;
            v[32] = (x3 - t0) >> 10;
        }
    }
    v = val;
    uint8_t *o = out;
    for (int i = 0; i < 8; ++i, v += 8, o += out_stride) {
        IDCT_1D(v[0], v[1], v[2], v[3], v[4], v[5], v[6], v[7])
        x0 += 65536 + (128 << 17);
        x1 += 65536 + (128 << 17);
        x2 += 65536 + (128 << 17);
        x3 += 65536 + (128 << 17);
        o[0] = clamp8((x0 + t3) >> 17);
        o[7] = clamp8((x0 - t3) >> 17);
        o[1] = clamp8((x1 + t2) >> 17);
        o[6] = clamp8((x1 - t2) >> 17);
        o[2] = clamp8((x2 + t1) >> 17);
        o[5] = clamp8((x2 - t1) >> 17);
        o[3] = clamp8((x3 + t0) >> 17);
        o[4] = clamp8((x3 - t0) >> 17);
    }
}

struct Component {
    int id = 0, h = 1, v = 1, tq = 0, td = 0, ta = 0;
    int pred = 0;
    int stride = 0, rows = 0; // padded plane
    std::vector<uint8_t> plane;
};

uint16_t be16(const uint8_t *p) { return (uint16_t)((p[0] << 8) | p[1]); }

// h2v1 "fancy" upsampling of one row: out[2i] = (3 in[i] + in[i-1] + 2) >> 2 (libjpeg jdsample.c)
void upsample_row_h2(const uint8_t *in, int w_in, uint8_t *out) {
    if (w_in == 1) {
        out[0] = out[1] = in[0];
        return;
    }
    out[0] = in[0];
    out[1] = (uint8_t)((in[0] * 3 + in[1] + 2) >> 2);
    for (int i = 1; i < w_in - 1; ++i) {
        int n = 3 * in[i];
        out[2 * i] = (uint8_t)((n + in[i - 1] + 1) >> 2);
        out[2 * i + 1] = (uint8_t)((n + in[i + 1] + 2) >> 2);
    }
    out[2 * (w_in - 1)] = (uint8_t)((in[w_in - 1] * 3 + in[w_in - 2] + 1) >> 2);
    out[2 * (w_in - 1) + 1] = in[w_in - 1];
}

} // namespace

bool decode_jpeg_mem(const uint8_t *data, size_t n, int &width, int &height, std::vector<uint8_t> &rgb,
                     std::string &err) {
    if (n < 4 || data[0] != 0xff || data[1] != 0xd8) {
        err = "not a JPEG (no SOI)";
        return false;
    }
    uint16_t qt[4][64];
    bool qt_present[4] = {false, false, false, false};
    Huff hdc[4], hac[4];
    std::vector<Component> comps;
    int restart_interval = 0;
    bool have_sof = false;
    int hmax = 1, vmax = 1;
    size_t pos = 2;
    while (pos + 4 <= n) {
        if (data[pos] != 0xff) {
            err = "marker expected";
            return false;
        }
        int m = data[pos + 1];
        if (m == 0xff) { // fill byte
            ++pos;
            continue;
        }
        pos += 2;
        if (m == 0xd8 || m == 0x01 || (m >= 0xd0 && m <= 0xd7)) continue;
        if (m == 0xd9) break;
        if (pos + 2 > n) break;
        size_t len = be16(data + pos);
        if (len < 2 || pos + len > n) {
            err = "truncated segment";
            return false;
        }
        const uint8_t *seg = data + pos + 2;
        size_t seglen = len - 2;
        if (m == 0xdb) { // DQT
            size_t i = 0;
            while (i < seglen) {
                int pq = seg[i] >> 4, tq = seg[i] & 15;
                ++i;
                if (tq > 3 || i + (pq ? 128u : 64u) > seglen) {
                    err = "bad DQT";
                    return false;
                }
                for (int k = 0; k < 64; ++k) {
                    qt[tq][ZIGZAG[k]] = pq ? be16(seg + i + 2 * k) : seg[i + k];
                }
                i += pq ? 128 : 64;
                qt_present[tq] = true;
            }
        } else if (m == 0xc4) { // DHT
            size_t i = 0;
            while (i + 17 <= seglen) {
                int tc = seg[i] >> 4, th = seg[i] & 15;
                const uint8_t *counts = seg + i + 1;
                int total = 0;
                for (int k = 0; k < 16; ++k) total += counts[k];
                if (th > 3 || tc > 1 || total > 256 || i + 17 + (size_t)total > seglen) {
                    err = "bad DHT";
                    return false;
                }
                (tc ? hac[th] : hdc[th]).build(counts, seg + i + 17, total);
                i += 17 + (size_t)total;
            }
        } else if (m == 0xc0 || m == 0xc1) { // SOF0 / SOF1
            if (seglen < 6 || seg[0] != 8) {
                err = "only 8-bit baseline JPEG is supported";
                return false;
            }
            height = be16(seg + 1);
            width = be16(seg + 3);
            int nc = seg[5];
            if ((nc != 1 && nc != 3) || seglen < 6 + 3 * (size_t)nc || width <= 0 || height <= 0) {
                err = "unsupported component count or size";
                return false;
            }
            comps.resize((size_t)nc);
            for (int c = 0; c < nc; ++c) {
                comps[c].id = seg[6 + 3 * c];
                comps[c].h = seg[7 + 3 * c] >> 4;
                comps[c].v = seg[7 + 3 * c] & 15;
                comps[c].tq = seg[8 + 3 * c];
                if (comps[c].h < 1 || comps[c].h > 4 || comps[c].v < 1 || comps[c].v > 4 || comps[c].tq > 3) {
                    err = "bad sampling factors";
                    return false;
                }
                hmax = comps[c].h > hmax ? comps[c].h : hmax;
                vmax = comps[c].v > vmax ? comps[c].v : vmax;
            }
            /* the upsampler replicates by whole factors: a component whose factor does not divide the largest one
             * (e.g. h = 3 beside h = 4) has planes narrower than the reads below assume */
            for (int c = 0; c < nc; ++c)
                if (hmax % comps[c].h != 0 || vmax % comps[c].v != 0) {
                    err = "fractional chroma subsampling is not supported";
                    return false;
                }
            have_sof = true;
        } else if (m == 0xc2 || (m >= 0xc3 && m <= 0xcf && m != 0xc4 && m != 0xc8 && m != 0xcc)) {
            err = "progressive / lossless / arithmetic JPEG is not supported";
            return false;
        } else if (m == 0xdd) { // DRI
            if (seglen >= 2) restart_interval = be16(seg);
        } else if (m == 0xda) { // SOS: one interleaved scan with all components
            if (!have_sof) {
                err = "SOS before SOF";
                return false;
            }
            int ns = seg[0];
            if (ns != (int)comps.size() || seglen < 1 + 2 * (size_t)ns + 3) {
                err = "non-interleaved scans are not supported";
                return false;
            }
            for (int k = 0; k < ns; ++k) {
                int cid = seg[1 + 2 * k];
                Component *cp = nullptr;
                for (Component &c : comps)
                    if (c.id == cid) cp = &c;
                if (!cp) {
                    err = "scan references unknown component";
                    return false;
                }
                cp->td = seg[2 + 2 * k] >> 4;
                cp->ta = seg[2 + 2 * k] & 15;
                if (cp->td > 3 || cp->ta > 3 || !hdc[cp->td].present || !hac[cp->ta].present || !qt_present[cp->tq]) {
                    err = "scan references a missing table";
                    return false;
                }
            }
            pos += len;
            // ---- entropy-coded data ----
            int mcux = (width + 8 * hmax - 1) / (8 * hmax), mcuy = (height + 8 * vmax - 1) / (8 * vmax);
            for (Component &c : comps) {
                c.stride = mcux * c.h * 8;
                c.rows = mcuy * c.v * 8;
                c.plane.assign((size_t)c.stride * (size_t)c.rows, 0);
                c.pred = 0;
            }
            BitReader br;
            br.p = data + pos;
            br.end = data + n;
            int todo = restart_interval ? restart_interval : 0x7fffffff;
            for (int my = 0; my < mcuy; ++my)
                for (int mx = 0; mx < mcux; ++mx) {
                    for (Component &c : comps)
                        for (int by = 0; by < c.v; ++by)
                            for (int bx = 0; bx < c.h; ++bx) {
                                int blk[64];
                                std::memset(blk, 0, sizeof blk);
                                int s = decode_symbol(br, hdc[c.td]);
                                if (s < 0 || s > 15) {
                                    err = "bad DC code";
                                    return false;
                                }
                                int diff = s ? extend(br.get(s), s) : 0;
                                c.pred += diff;
                                blk[0] = c.pred * qt[c.tq][0];
                                for (int k = 1; k < 64;) {
                                    int rs = decode_symbol(br, hac[c.ta]);
                                    if (rs < 0) {
                                        err = "bad AC code";
                                        return false;
                                    }
                                    int r = rs >> 4, sz = rs & 15;
                                    if (sz == 0) {
                                        if (r != 15) break; // EOB
                                        k += 16;
                                        continue;
                                    }
                                    k += r;
                                    if (k > 63) {
                                        err = "AC index overflow";
                                        return false;
                                    }
                                    int z = ZIGZAG[k];
                                    blk[z] = extend(br.get(sz), sz) * qt[c.tq][z];
                                    ++k;
                                }
                                uint8_t *dst = c.plane.data() + (size_t)((my * c.v + by) * 8) * (size_t)c.stride +
                                               (size_t)((mx * c.h + bx) * 8);
                                idct_block(dst, c.stride, blk);
                            }
                    if (--todo <= 0) { // restart marker
                        br.reset();
                        const uint8_t *q = br.p;
                        while (q + 1 < br.end && !(q[0] == 0xff && q[1] >= 0xd0 && q[1] <= 0xd7)) ++q;
                        if (q + 1 < br.end) br.p = q + 2;
                        for (Component &c : comps) c.pred = 0;
                        todo = restart_interval;
                    }
                }
            // ---- upsample + colour-convert ----
            rgb.assign((size_t)width * (size_t)height * 3, 0);
            std::vector<std::vector<uint8_t>> full(comps.size());
            for (size_t ci = 0; ci < comps.size(); ++ci) {
                Component &c = comps[ci];
                int sx = hmax / c.h, sy = vmax / c.v;
                int cw = (width * c.h + hmax - 1) / hmax, ch = (height * c.v + vmax - 1) / vmax;
                std::vector<uint8_t> &f = full[ci];
                f.assign((size_t)width * (size_t)height, 0);
                if (sx == 1 && sy == 1) {
                    for (int y = 0; y < height; ++y)
                        std::memcpy(&f[(size_t)y * width], &c.plane[(size_t)y * c.stride], (size_t)width);
                } else if (sx == 2 && (sy == 1 || sy == 2) && hmax * c.h == hmax && c.h * sx == hmax) {
                    std::vector<uint8_t> row((size_t)cw), wide((size_t)cw * 2 + 2);
                    for (int y = 0; y < height; ++y) {
                        if (sy == 1) {
                            upsample_row_h2(&c.plane[(size_t)y * c.stride], cw, wide.data());
                        } else { // h2v2: vertical 3:1 blend of the nearest two rows, then horizontal
                            int y0 = y >> 1;
                            int y1 = (y & 1) ? (y0 + 1 < ch ? y0 + 1 : y0) : (y0 > 0 ? y0 - 1 : y0);
                            const uint8_t *near = &c.plane[(size_t)y0 * c.stride], *far = &c.plane[(size_t)y1 * c.stride];
                            // libjpeg h2v2 fancy: 16-bit intermediate 3*near+far, then (3*a+b+8)>>4 / (..+7)>>4
                            std::vector<int> col((size_t)cw);
                            for (int i = 0; i < cw; ++i) col[i] = 3 * near[i] + far[i];
                            if (cw == 1) {
                                wide[0] = wide[1] = (uint8_t)((col[0] * 4 + 8) >> 4);
                            } else {
                                wide[0] = (uint8_t)((col[0] * 4 + 8) >> 4);
                                wide[1] = (uint8_t)((col[0] * 3 + col[1] + 7) >> 4);
                                for (int i = 1; i < cw - 1; ++i) {
                                    wide[2 * i] = (uint8_t)((col[i] * 3 + col[i - 1] + 8) >> 4);
                                    wide[2 * i + 1] = (uint8_t)((col[i] * 3 + col[i + 1] + 7) >> 4);
                                }
                                wide[2 * (cw - 1)] = (uint8_t)((col[cw - 1] * 3 + col[cw - 2] + 8) >> 4);
                                wide[2 * (cw - 1) + 1] = (uint8_t)((col[cw - 1] * 4 + 7) >> 4);
                            }
                        }
                        std::memcpy(&f[(size_t)y * width], wide.data(), (size_t)width);
                    }
                } else { // generic: replicate
                    for (int y = 0; y < height; ++y)
                        for (int x = 0; x < width; ++x)
                            f[(size_t)y * width + x] = c.plane[(size_t)(y / sy) * c.stride + (size_t)(x / sx)];
                }
                (void)ch;
            }
            for (size_t i = 0; i < (size_t)width * (size_t)height; ++i) {
                if (comps.size() == 1) {
                    rgb[3 * i] = rgb[3 * i + 1] = rgb[3 * i + 2] = full[0][i];
                } else { // JFIF YCbCr -> RGB
                    float y = (float)full[0][i], cb = (float)full[1][i] - 128.0f, cr = (float)full[2][i] - 128.0f;
                    float r = y + 1.40200f * cr, g = y - 0.34414f * cb - 0.71414f * cr, b = y + 1.77200f * cb;
                    rgb[3 * i] = clamp8((int)(r + 0.5f + 256.0f) - 256);
                    rgb[3 * i + 1] = clamp8((int)(g + 0.5f + 256.0f) - 256);
                    rgb[3 * i + 2] = clamp8((int)(b + 0.5f + 256.0f) - 256);
                }
            }
            return true;
        }
        pos += len;
    }
    err = "no scan found";
    return false;
}

bool decode_jpeg(const std::string &path, int &w, int &h, std::vector<uint8_t> &rgb, std::string &err) {
    std::ifstream in(path, std::ios::binary);
    if (!in) {
        err = "No such file or directory";
        return false;
    }
    std::vector<uint8_t> data((std::istreambuf_iterator<char>(in)), std::istreambuf_iterator<char>());
    return decode_jpeg_mem(data.data(), data.size(), w, h, rgb, err);
}

} // namespace rt
