// Image files -> MIP pyramids for the host-side scene ingest: what `MipMap::new` (texturing/textures/image.rs:218-262,
// 289-335) does with the `image` crate — `image::open`, one `resize_exact(dx, dy, FilterType::Lanczos3)` of the ORIGINAL
// picture per level (level i = max(np2x >> i, 1) x max(np2y >> i, 1), np2 = next power of two), `to_rgb()` / `to_luma()`,
// `convert_in` (u8 -> [0, 1], inverse sRGB gamma if asked, x scale) — and the level-0 mean (`MipMap::mean`).
//
// PARITY UNPINNED: decoder and resampler are third-party code that is not part of the reference tree (Cargo.toml:
// image = "0.12", no lock file).  `resize_lanczos3` restates imageops::resize of that release line as published
// (vertical pass then horizontal pass; kernel support 3 scaled by the down-sampling ratio; window = ceil(x - r) ..
// floor(x + r) clamped to the picture; weights (i - x) / scale, normalised by their sum; clamp to [0, 255], truncation to
// u8), `to_luma` its BT.709 weights.  Nothing in the reference pins these numbers, so no test claims bit-exactness for
// them; the renderer downstream of the pyramid IS pinned (kernels/shade_tex.cuh against oracle/texture.hpp).
//
// Decoders: PNG (8 / 16 bit gray, gray + alpha, RGB, RGBA; palette 1 .. 8 bit; no Adam7 interlace; inflate by zlib), TGA
// (true colour, gray, colour-mapped; raw or run-length encoded) and baseline JPEG (gray / YCbCr, 4:4:4 .. 4:2:0, restart intervals).  Any other file, like a missing one, makes the caller fall back to the MTL constant — `image::open(..)` failing has the
// same effect in load_obj (component/mod.rs:88-95).
#pragma once
#include <zlib.h>
#include <algorithm>
#include <cctype>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>
#include "../../../include/arn.h"

namespace arnhost {

struct Image8 {
    uint32_t w = 0, h = 0, ch = 0;          // ch: 1 gray, 2 gray + alpha, 3 RGB, 4 RGBA
    std::vector<uint8_t> px;                 // row-major, ch bytes per pixel
};

namespace detail {
inline uint32_t be32(const uint8_t* p) { return ((uint32_t)p[0] << 24) | ((uint32_t)p[1] << 16) | ((uint32_t)p[2] << 8) | (uint32_t)p[3]; }
inline int paeth(int a, int b, int c) {
    int p = a + b - c, pa = std::abs(p - a), pb = std::abs(p - b), pc = std::abs(p - c);
    return (pa <= pb && pa <= pc) ? a : (pb <= pc ? b : c);
}
}  // namespace detail

inline bool png_decode(const std::string& path, Image8* out, std::string* err) {
    auto fail = [&](const std::string& why) { if (err) *err = path + ": " + why; return false; };
    FILE* f = std::fopen(path.c_str(), "rb");
    if (!f) return fail("cannot open");
    std::vector<uint8_t> file;
    { uint8_t buf[65536]; size_t n; while ((n = std::fread(buf, 1, sizeof buf, f)) > 0) file.insert(file.end(), buf, buf + n); }
    std::fclose(f);
    static const uint8_t sig[8] = {0x89, 'P', 'N', 'G', 0x0d, 0x0a, 0x1a, 0x0a};
    if (file.size() < 8 + 25 || std::memcmp(file.data(), sig, 8) != 0) return fail("not a PNG file");
    uint32_t w = 0, h = 0; int depth = 0, ctype = -1, interlace = 0;
    std::vector<uint8_t> idat, plte;
    size_t pos = 8; bool end = false;
    while (!end && pos + 12 <= file.size()) {
        uint32_t len = detail::be32(&file[pos]);
        const uint8_t* type = &file[pos + 4];
        if (pos + 12 + (size_t)len > file.size()) return fail("truncated chunk");
        const uint8_t* data = &file[pos + 8];
        if (!std::memcmp(type, "IHDR", 4)) {
            if (len != 13) return fail("bad IHDR");
            w = detail::be32(data); h = detail::be32(data + 4); depth = data[8]; ctype = data[9]; interlace = data[12];
            if (data[10] != 0 || data[11] != 0) return fail("unknown compression / filter method");
        } else if (!std::memcmp(type, "PLTE", 4)) plte.assign(data, data + len);
        else if (!std::memcmp(type, "IDAT", 4)) idat.insert(idat.end(), data, data + len);
        else if (!std::memcmp(type, "IEND", 4)) end = true;
        pos += 12 + (size_t)len;
    }
    if (ctype < 0 || w == 0 || h == 0 || w > 32768 || h > 32768) return fail("missing or implausible IHDR (pictures up to 32768 x 32768)");
    if (interlace != 0) return fail("interlaced PNG is not supported");
    if (idat.empty()) return fail("no image data");
    int samples;                                       // samples per pixel in the file
    switch (ctype) { case 0: samples = 1; break; case 2: samples = 3; break; case 3: samples = 1; break; case 4: samples = 2; break; case 6: samples = 4; break; default: return fail("unknown colour type"); }
    if (ctype == 3) { if (depth != 1 && depth != 2 && depth != 4 && depth != 8) return fail("bad palette bit depth"); if (plte.size() < 3) return fail("palette image without PLTE"); }
    else if (depth != 8 && depth != 16) return fail("only 8 and 16 bit samples are supported");
    const size_t bits_pp = (size_t)samples * (size_t)depth, bpp = (bits_pp + 7) / 8, stride = ((size_t)w * bits_pp + 7) / 8;
    std::vector<uint8_t> raw((stride + 1) * (size_t)h);
    {
        z_stream zs; std::memset(&zs, 0, sizeof zs);
        if (inflateInit(&zs) != Z_OK) return fail("zlib init failed");
        zs.next_in = idat.data(); zs.avail_in = (uInt)idat.size(); zs.next_out = raw.data(); zs.avail_out = (uInt)raw.size();
        int rc = inflate(&zs, Z_FINISH);
        size_t got = raw.size() - zs.avail_out;
        inflateEnd(&zs);
        if ((rc != Z_STREAM_END && rc != Z_OK && rc != Z_BUF_ERROR) || got != raw.size()) return fail("corrupt image data");
    }
    // undo the scan-line filters in place
    for (uint32_t y = 0; y < h; y++) {
        uint8_t* line = &raw[(stride + 1) * (size_t)y];
        const int ft = line[0]; uint8_t* cur = line + 1;
        const uint8_t* up = y ? &raw[(stride + 1) * (size_t)(y - 1) + 1] : nullptr;
        for (size_t i = 0; i < stride; i++) {
            const int a = i >= bpp ? cur[i - bpp] : 0, b = up ? up[i] : 0, c = (up && i >= bpp) ? up[i - bpp] : 0;
            int v = cur[i];
            switch (ft) { case 0: break; case 1: v += a; break; case 2: v += b; break; case 3: v += (a + b) / 2; break; case 4: v += detail::paeth(a, b, c); break; default: return fail("unknown filter type"); }
            cur[i] = (uint8_t)v;
        }
    }
    out->w = w; out->h = h; out->ch = ctype == 3 ? 3u : (uint32_t)samples;
    out->px.assign((size_t)w * h * out->ch, 0);
    for (uint32_t y = 0; y < h; y++) {
        const uint8_t* cur = &raw[(stride + 1) * (size_t)y + 1];
        uint8_t* dst = &out->px[(size_t)y * w * out->ch];
        if (ctype == 3) {
            for (uint32_t x = 0; x < w; x++) {
                const size_t bit = (size_t)x * (size_t)depth;
                const uint32_t idx = (cur[bit >> 3] >> (8 - depth - (int)(bit & 7))) & ((1u << depth) - 1u);
                if ((size_t)idx * 3 + 2 >= plte.size()) return fail("palette index out of range");
                dst[3 * x] = plte[3 * idx]; dst[3 * x + 1] = plte[3 * idx + 1]; dst[3 * x + 2] = plte[3 * idx + 2];
            }
        } else if (depth == 8) std::memcpy(dst, cur, (size_t)w * samples);
        else for (size_t i = 0; i < (size_t)w * samples; i++) dst[i] = cur[2 * i];      // 16 bit: the high byte
    }
    return true;
}

// Truevision TGA: true-colour (24 / 32 bit), gray (8 bit) and colour-mapped (8 bit indices into a 24 / 32 bit map) pictures, raw or
// run-length encoded; rows are returned top to bottom whatever the file's origin flag says.
inline bool tga_decode(const std::string& path, Image8* out, std::string* err) {
    auto fail = [&](const std::string& why) { if (err) *err = path + ": " + why; return false; };
    FILE* f = std::fopen(path.c_str(), "rb");
    if (!f) return fail("cannot open");
    std::vector<uint8_t> file;
    { uint8_t buf[65536]; size_t n; while ((n = std::fread(buf, 1, sizeof buf, f)) > 0) file.insert(file.end(), buf, buf + n); }
    std::fclose(f);
    if (file.size() < 18) return fail("not a TGA file");
    const uint8_t* h = file.data();
    const uint32_t id_len = h[0], cmap_type = h[1], type = h[2], cmap_first = h[3] | (h[4] << 8), cmap_len = h[5] | (h[6] << 8), cmap_bits = h[7];
    const uint32_t w = h[12] | (h[13] << 8), ht = h[14] | (h[15] << 8), bits = h[16], desc = h[17];
    const bool rle = type == 9 || type == 10 || type == 11;
    const uint32_t base = rle ? type - 8 : type;                           // 1 colour-mapped, 2 true colour, 3 gray
    if (w == 0 || ht == 0 || base < 1 || base > 3) return fail("unsupported TGA image type");
    if ((base == 2 && bits != 24 && bits != 32) || (base == 3 && bits != 8) || (base == 1 && (bits != 8 || cmap_type != 1 || (cmap_bits != 24 && cmap_bits != 32))))
        return fail("unsupported TGA pixel format");
    size_t pos = 18 + id_len;
    const uint8_t* cmap = nullptr; const uint32_t cbytes = cmap_bits / 8;
    if (cmap_type == 1) { if (pos + (size_t)cmap_len * cbytes > file.size()) return fail("truncated colour map"); cmap = &file[pos]; pos += (size_t)cmap_len * cbytes; }
    const uint32_t bpp = bits / 8; const size_t npix = (size_t)w * ht;
    std::vector<uint8_t> raw(npix * bpp);
    if (!rle) { if (pos + raw.size() > file.size()) return fail("truncated image data"); std::memcpy(raw.data(), &file[pos], raw.size()); }
    else {
        size_t o = 0;
        while (o < npix) {
            if (pos >= file.size()) return fail("truncated run-length data");
            const uint32_t hdr = file[pos++], cnt = (hdr & 127u) + 1u;
            if (o + cnt > npix) return fail("run past the end of the picture");
            if (hdr & 128u) {
                if (pos + bpp > file.size()) return fail("truncated run-length data");
                for (uint32_t k = 0; k < cnt; k++) std::memcpy(&raw[(o + k) * bpp], &file[pos], bpp);
                pos += bpp;
            } else {
                if (pos + (size_t)cnt * bpp > file.size()) return fail("truncated run-length data");
                std::memcpy(&raw[o * bpp], &file[pos], (size_t)cnt * bpp); pos += (size_t)cnt * bpp;
            }
            o += cnt;
        }
    }
    out->w = w; out->h = ht; out->ch = base == 3 ? 1u : ((base == 2 ? bits : cmap_bits) == 32 ? 4u : 3u);
    out->px.assign(npix * out->ch, 0);
    const bool top_origin = (desc & 0x20u) != 0, right_origin = (desc & 0x10u) != 0;
    for (uint32_t y = 0; y < ht; y++) for (uint32_t x = 0; x < w; x++) {
        const uint32_t sy = top_origin ? y : ht - 1 - y, sx = right_origin ? w - 1 - x : x;
        const uint8_t* s = &raw[((size_t)sy * w + sx) * bpp];
        uint8_t* d = &out->px[((size_t)y * w + x) * out->ch];
        if (base == 3) d[0] = s[0];
        else {
            const uint8_t* c = s;
            if (base == 1) { const uint32_t idx = s[0]; if (idx < cmap_first || idx - cmap_first >= cmap_len) return fail("colour index outside the map"); c = cmap + (size_t)(idx - cmap_first) * cbytes; }
            d[0] = c[2]; d[1] = c[1]; d[2] = c[0]; if (out->ch == 4) d[3] = c[3];             // BGR(A) -> RGB(A)
        }
    }
    return true;
}

// JPEG, baseline / extended sequential Huffman (SOF0 / SOF1, 8 bit; gray or YCbCr, sampling factors 1 .. 2 per axis, restart
// intervals).  Progressive, arithmetic-coded, 12-bit and CMYK files are refused (-> the caller's fall-back, as above).  The
// inverse DCT is the separable double-precision definition, chroma is brought to full resolution with the triangle filter
// libjpeg calls "fancy upsampling" for 2:1 factors (replication otherwise), YCbCr -> RGB by the JFIF equations.
namespace detail {
struct JpegHuff { uint16_t code[256]; uint8_t size[256], val[256]; int n = 0; int maxcode[18]; int valptr[17]; int mincode[17]; bool set = false; };
inline bool jpeg_build_huff(const uint8_t* bits, const uint8_t* vals, int nvals, JpegHuff* h) {
    int k = 0, code = 0;
    for (int l = 1; l <= 16; l++) {
        h->valptr[l] = k; h->mincode[l] = code;
        for (int i = 0; i < bits[l - 1]; i++) { if (k >= nvals || k >= 256) return false; h->val[k] = vals[k]; k++; code++; }
        h->maxcode[l] = bits[l - 1] ? code - 1 : -1;
        if (code > (1 << l)) return false;
        code <<= 1;
    }
    h->n = k; h->set = true;
    return true;
}
struct JpegBits {
    const uint8_t* p; const uint8_t* end; uint32_t acc = 0; int cnt = 0; bool hit_marker = false;
    int bit() {
        if (cnt == 0) {
            uint8_t b = 0;
            if (!hit_marker && p < end) {
                b = *p++;
                if (b == 0xff) {
                    if (p < end && *p == 0x00) p++;                        // stuffed byte
                    else { hit_marker = true; p--; b = 0; }               // a marker: feed zeros until the caller resynchronises
                }
            } else hit_marker = true;
            acc = b; cnt = 8;
        }
        cnt--;
        return (int)((acc >> cnt) & 1u);
    }
    int bits(int n) { int v = 0; for (int i = 0; i < n; i++) v = (v << 1) | bit(); return v; }
    void reset() { cnt = 0; acc = 0; hit_marker = false; }
};
inline int jpeg_decode_sym(JpegBits& br, const JpegHuff& h) {
    int code = 0;
    for (int l = 1; l <= 16; l++) {
        code = (code << 1) | br.bit();
        if (h.maxcode[l] >= 0 && code <= h.maxcode[l] && code >= h.mincode[l]) return h.val[h.valptr[l] + code - h.mincode[l]];
    }
    return -1;
}
inline int jpeg_extend(int v, int t) { return t == 0 ? 0 : (v < (1 << (t - 1)) ? v - (1 << t) + 1 : v); }
inline void jpeg_idct8x8(const int* coef, const uint16_t* q, uint8_t* out, size_t stride) {
    static double c[8][8]; static bool init = false;
    if (!init) { for (int x = 0; x < 8; x++) for (int u = 0; u < 8; u++) c[x][u] = (u == 0 ? std::sqrt(0.125) : 0.5) * std::cos((2 * x + 1) * u * 3.14159265358979323846 / 16.0); init = true; }
    double tmp[64];
    for (int v = 0; v < 8; v++) for (int x = 0; x < 8; x++) { double a = 0; for (int u = 0; u < 8; u++) a += c[x][u] * (double)(coef[v * 8 + u] * (int)q[v * 8 + u]); tmp[v * 8 + x] = a; }
    for (int y = 0; y < 8; y++) for (int x = 0; x < 8; x++) {
        double a = 0; for (int v = 0; v < 8; v++) a += c[y][v] * tmp[v * 8 + x];
        long r = std::lround(a + 128.0);
        out[(size_t)y * stride + x] = (uint8_t)(r < 0 ? 0 : (r > 255 ? 255 : r));
    }
}
}  // namespace detail

inline bool jpeg_decode(const std::string& path, Image8* out, std::string* err) {
    using namespace detail;
    auto fail = [&](const std::string& why) { if (err) *err = path + ": " + why; return false; };
    FILE* f = std::fopen(path.c_str(), "rb");
    if (!f) return fail("cannot open");
    std::vector<uint8_t> file;
    { uint8_t buf[65536]; size_t n; while ((n = std::fread(buf, 1, sizeof buf, f)) > 0) file.insert(file.end(), buf, buf + n); }
    std::fclose(f);
    if (file.size() < 4 || file[0] != 0xff || file[1] != 0xd8) return fail("not a JPEG file");
    static const uint8_t zz[64] = {0, 1, 8, 16, 9, 2, 3, 10, 17, 24, 32, 25, 18, 11, 4, 5, 12, 19, 26, 33, 40, 48, 41, 34, 27, 20, 13, 6, 7, 14, 21, 28,
                                   35, 42, 49, 56, 57, 50, 43, 36, 29, 22, 15, 23, 30, 37, 44, 51, 58, 59, 52, 45, 38, 31, 39, 46, 53, 60, 61, 54, 47, 55, 62, 63};
    uint16_t qt[4][64]; bool qset[4] = {false, false, false, false};
    JpegHuff hdc[4], hac[4];
    struct Comp { int id = 0, h = 1, v = 1, tq = 0, td = 0, ta = 0, pred = 0; uint32_t w = 0, ht = 0; std::vector<uint8_t> plane; } comp[3];
    int ncomp = 0; uint32_t W = 0, H = 0; int hmax = 1, vmax = 1; uint32_t restart = 0; bool have_sof = false, done = false;
    size_t pos = 2;
    while (!done) {
        while (pos < file.size() && file[pos] != 0xff) pos++;                 // to the next marker
        while (pos < file.size() && file[pos] == 0xff) pos++;
        if (pos >= file.size()) return fail("no image data before the end of the file");
        const uint8_t m = file[pos++];
        if (m == 0xd8 || m == 0x01 || (m >= 0xd0 && m <= 0xd7)) continue;      // stand-alone markers
        if (m == 0xd9) return fail("end of image before any scan");
        if (pos + 2 > file.size()) return fail("truncated segment");
        const size_t len = ((size_t)file[pos] << 8) | file[pos + 1];
        if (len < 2 || pos + len > file.size()) return fail("truncated segment");
        const uint8_t* seg = &file[pos + 2]; const size_t n = len - 2;
        if (m == 0xdb) {                                                       // DQT
            size_t o = 0;
            while (o < n) {
                const int pq = seg[o] >> 4, tq = seg[o] & 15; o++;
                if (tq > 3 || pq > 1 || o + (size_t)64 * (pq + 1) > n) return fail("bad quantisation table");
                for (int i = 0; i < 64; i++) { qt[tq][zz[i]] = pq ? (uint16_t)((seg[o] << 8) | seg[o + 1]) : seg[o]; o += pq + 1; }
                qset[tq] = true;
            }
        } else if (m == 0xc4) {                                                // DHT
            size_t o = 0;
            while (o < n) {
                if (o + 17 > n) return fail("bad Huffman table");
                const int tc = seg[o] >> 4, th = seg[o] & 15; int cnt = 0;
                for (int i = 0; i < 16; i++) cnt += seg[o + 1 + i];
                if (tc > 1 || th > 3 || cnt > 256 || o + 17 + (size_t)cnt > n) return fail("bad Huffman table");
                if (!jpeg_build_huff(&seg[o + 1], &seg[o + 17], cnt, tc ? &hac[th] : &hdc[th])) return fail("bad Huffman table");
                o += 17 + (size_t)cnt;
            }
        } else if (m == 0xc0 || m == 0xc1) {                                   // SOF0 / SOF1
            if (n < 6 || seg[0] != 8) return fail("only 8-bit JPEG is supported");
            H = ((uint32_t)seg[1] << 8) | seg[2]; W = ((uint32_t)seg[3] << 8) | seg[4]; ncomp = seg[5];
            if (W == 0 || H == 0 || W > 32768 || H > 32768) return fail("implausible JPEG size");
            if ((ncomp != 1 && ncomp != 3) || n < (size_t)6 + 3 * (size_t)ncomp) return fail("only gray and YCbCr JPEG files are supported");
            for (int i = 0; i < ncomp; i++) {
                comp[i].id = seg[6 + 3 * i]; comp[i].h = seg[7 + 3 * i] >> 4; comp[i].v = seg[7 + 3 * i] & 15; comp[i].tq = seg[8 + 3 * i];
                if (comp[i].h < 1 || comp[i].h > 2 || comp[i].v < 1 || comp[i].v > 2 || comp[i].tq > 3) return fail("unsupported JPEG sampling factors");
                hmax = std::max(hmax, comp[i].h); vmax = std::max(vmax, comp[i].v);
            }
            if (ncomp == 1) { comp[0].h = comp[0].v = 1; hmax = vmax = 1; }
            have_sof = true;
        } else if (m == 0xc2 || (m >= 0xc3 && m <= 0xcf && m != 0xc4 && m != 0xc8 && m != 0xcc)) {
            return fail("progressive, lossless and arithmetic-coded JPEG files are not supported");
        } else if (m == 0xdd) {                                                // DRI
            if (n < 2) return fail("bad restart interval"); restart = ((uint32_t)seg[0] << 8) | seg[1];
        } else if (m == 0xda) {                                                // SOS: the one scan of a sequential file
            if (!have_sof) return fail("scan before the frame header");
            if (n < 1 || seg[0] != ncomp || n < (size_t)1 + 2 * (size_t)ncomp + 3) return fail("multi-scan sequential JPEG files are not supported");
            for (int i = 0; i < ncomp; i++) {
                int ci = -1; for (int k = 0; k < ncomp; k++) if (comp[k].id == seg[1 + 2 * i]) ci = k;
                if (ci != i) return fail("unexpected component order in the scan");
                comp[i].td = seg[2 + 2 * i] >> 4; comp[i].ta = seg[2 + 2 * i] & 15;
                if (comp[i].td > 3 || comp[i].ta > 3 || !hdc[comp[i].td].set || !hac[comp[i].ta].set || !qset[comp[i].tq]) return fail("scan refers to a table that was not defined");
            }
            const uint32_t mcux = (W + 8 * hmax - 1) / (8 * hmax), mcuy = (H + 8 * vmax - 1) / (8 * vmax);
            for (int i = 0; i < ncomp; i++) { comp[i].w = mcux * 8 * comp[i].h; comp[i].ht = mcuy * 8 * comp[i].v; comp[i].plane.assign((size_t)comp[i].w * comp[i].ht, 0); comp[i].pred = 0; }
            JpegBits br; br.p = &file[pos + len]; br.end = file.data() + file.size();
            uint32_t until_restart = restart; int coef[64];
            for (uint32_t my = 0; my < mcuy; my++) for (uint32_t mx = 0; mx < mcux; mx++) {
                if (restart && until_restart == 0) {                           // RSTn: byte-align, skip the marker, reset the predictors
                    br.reset();
                    while (br.p + 1 < br.end && !(br.p[0] == 0xff && br.p[1] >= 0xd0 && br.p[1] <= 0xd7)) br.p++;
                    if (br.p + 1 < br.end) br.p += 2;
                    for (int i = 0; i < ncomp; i++) comp[i].pred = 0;
                    until_restart = restart;
                }
                for (int i = 0; i < ncomp; i++) for (int by = 0; by < comp[i].v; by++) for (int bx = 0; bx < comp[i].h; bx++) {
                    std::memset(coef, 0, sizeof coef);
                    const int t = jpeg_decode_sym(br, hdc[comp[i].td]);
                    if (t < 0 || t > 11) return fail("corrupt entropy-coded data");
                    comp[i].pred += jpeg_extend(br.bits(t), t); coef[0] = comp[i].pred;
                    for (int k = 1; k < 64;) {
                        const int rs = jpeg_decode_sym(br, hac[comp[i].ta]);
                        if (rs < 0) return fail("corrupt entropy-coded data");
                        const int r = rs >> 4, sz = rs & 15;
                        if (sz == 0) { if (r == 15) { k += 16; continue; } break; }
                        k += r; if (k > 63) return fail("corrupt entropy-coded data");
                        coef[zz[k]] = jpeg_extend(br.bits(sz), sz); k++;
                    }
                    if (br.hit_marker && br.p >= br.end) return fail("truncated entropy-coded data");
                    const size_t x0 = ((size_t)mx * comp[i].h + bx) * 8, y0 = ((size_t)my * comp[i].v + by) * 8;
                    jpeg_idct8x8(coef, qt[comp[i].tq], &comp[i].plane[y0 * comp[i].w + x0], comp[i].w);
                }
                if (restart) until_restart--;
            }
            done = true;
        }
        pos += len;
    }
    out->w = W; out->h = H; out->ch = ncomp == 1 ? 1u : 3u; out->px.assign((size_t)W * H * out->ch, 0);
    if (ncomp == 1) { for (uint32_t y = 0; y < H; y++) std::memcpy(&out->px[(size_t)y * W], &comp[0].plane[(size_t)y * comp[0].w], W); return true; }
    // chroma to full resolution: sample centres of a 2:1 plane sit between the luma samples -> weights 3/4, 1/4 per axis
    auto sample_full = [&](const Comp& c, uint32_t x, uint32_t y) -> double {
        const int fx = hmax / c.h, fy = vmax / c.v;
        // the part of the plane that carries picture content (the rest is MCU padding)
        const long cw = (long)((W * (uint32_t)c.h + hmax - 1) / hmax), chh = (long)((H * (uint32_t)c.v + vmax - 1) / vmax);
        auto at = [&](long xx, long yy) { xx = xx < 0 ? 0 : (xx >= cw ? cw - 1 : xx); yy = yy < 0 ? 0 : (yy >= chh ? chh - 1 : yy); return (double)c.plane[(size_t)yy * c.w + (size_t)xx]; };
        long x0 = x, x1 = x; double wx = 0; long y0 = y, y1 = y; double wy = 0;
        if (fx == 2) { x0 = x / 2; x1 = (x & 1) ? x0 + 1 : x0 - 1; wx = 0.25; }
        if (fy == 2) { y0 = y / 2; y1 = (y & 1) ? y0 + 1 : y0 - 1; wy = 0.25; }
        const double a = at(x0, y0) * (1 - wx) + at(x1, y0) * wx, b = at(x0, y1) * (1 - wx) + at(x1, y1) * wx;
        return a * (1 - wy) + b * wy;
    };
    auto clamp8 = [](double v) { long r = std::lround(v); return (uint8_t)(r < 0 ? 0 : (r > 255 ? 255 : r)); };
    for (uint32_t y = 0; y < H; y++) for (uint32_t x = 0; x < W; x++) {
        const double Y = sample_full(comp[0], x, y), cb = sample_full(comp[1], x, y) - 128.0, cr = sample_full(comp[2], x, y) - 128.0;
        uint8_t* d = &out->px[((size_t)y * W + x) * 3];
        d[0] = clamp8(Y + 1.402 * cr); d[1] = clamp8(Y - 0.344136 * cb - 0.714136 * cr); d[2] = clamp8(Y + 1.772 * cb);
    }
    return true;
}
// image::open: by content for PNG and JPEG, by extension for TGA (the format has no signature)
inline bool image_decode(const std::string& path, Image8* out, std::string* err) {
    auto ends_with = [&](const char* e) { const size_t n = std::strlen(e); if (path.size() < n) return false; for (size_t i = 0; i < n; i++) if (std::tolower((unsigned char)path[path.size() - n + i]) != e[i]) return false; return true; };
    if (ends_with(".tga")) return tga_decode(path, out, err);
    { FILE* f = std::fopen(path.c_str(), "rb"); uint8_t sig[2] = {0, 0}; if (f) { size_t got = std::fread(sig, 1, 2, f); std::fclose(f); if (got == 2 && sig[0] == 0xff && sig[1] == 0xd8) return jpeg_decode(path, out, err); } }
    return png_decode(path, out, err);
}

namespace detail {
inline float lanczos3_kernel(float x) {                                   // imageops::sample::lanczos3_kernel
    const float t = 3.f;
    if (std::fabs(x) >= t) return 0.f;
    auto sinc = [](float v) { if (v == 0.f) return 1.f; const float a = v * 3.14159265358979323846f; return std::sin(a) / a; };
    return sinc(x) * sinc(x / t);
}
// one pass of imageops::sample::{horizontal_sample, vertical_sample}: resample `n_in` samples at stride `step` to `n_out`
inline void sample_line(const uint8_t* in, size_t step, uint32_t n_in, uint8_t* outp, size_t out_step, uint32_t n_out, uint32_t ch) {
    const float ratio = (float)n_in / (float)n_out;
    const float filter_scale = ratio > 1.f ? ratio : 1.f;
    const float filter_radius = std::ceil(3.f * filter_scale);
    for (uint32_t o = 0; o < n_out; o++) {
        const float inputx = ((float)o + 0.5f) * ratio;
        long left = (long)std::ceil(inputx - filter_radius), right = (long)std::floor(inputx + filter_radius);
        left = left < 0 ? 0 : (left > (long)n_in - 1 ? (long)n_in - 1 : left);
        right = right < 0 ? 0 : (right > (long)n_in - 1 ? (long)n_in - 1 : right);
        float sum = 0.f, t[4] = {0.f, 0.f, 0.f, 0.f};
        for (long i = left; i <= right; i++) {
            const float w = lanczos3_kernel(((float)i - inputx) / filter_scale);
            sum += w;
            for (uint32_t c = 0; c < ch; c++) t[c] += (float)in[(size_t)i * step + c] * w;
        }
        for (uint32_t c = 0; c < ch; c++) {
            float v = t[c] / sum;
            v = v < 0.f ? 0.f : (v > 255.f ? 255.f : v);
            outp[(size_t)o * out_step + c] = (uint8_t)v;                   // NumCast: truncation
        }
    }
}
}  // namespace detail

// DynamicImage::resize_exact(nw, nh, FilterType::Lanczos3): vertical pass, then horizontal pass
inline Image8 resize_lanczos3(const Image8& src, uint32_t nw, uint32_t nh) {
    Image8 tmp; tmp.w = src.w; tmp.h = nh; tmp.ch = src.ch; tmp.px.assign((size_t)src.w * nh * src.ch, 0);
    for (uint32_t x = 0; x < src.w; x++)
        detail::sample_line(&src.px[(size_t)x * src.ch], (size_t)src.w * src.ch, src.h, &tmp.px[(size_t)x * src.ch], (size_t)src.w * src.ch, nh, src.ch);
    Image8 out; out.w = nw; out.h = nh; out.ch = src.ch; out.px.assign((size_t)nw * nh * src.ch, 0);
    for (uint32_t y = 0; y < nh; y++)
        detail::sample_line(&tmp.px[(size_t)y * src.w * src.ch], src.ch, src.w, &out.px[(size_t)y * nw * src.ch], src.ch, nw, src.ch);
    return out;
}

inline float inverse_gamma_correct(float v) {                             // image.rs:621-627
    if (v <= 0.04045f) return v * (1.0f / 12.92f);
    return std::pow((1.0f / 1.055f) * v, 2.4f);
}

// MipMap::new for an RGB (channels = 3) or Luma (channels = 1) texture.  Fills t->{channels, n_levels, level_*} and `texels`
// (level offsets relative to the start of `texels`), and mean[0 .. channels) = MipMap::mean.
inline bool build_pyramid(const std::string& path, uint32_t channels, bool gamma, float scale, arn_texture* t, std::vector<float>* texels,
                          float* mean, std::string* err) {
    Image8 img;
    if (!image_decode(path, &img, err)) return false;
    auto np2 = [](uint32_t v) { uint32_t p = 1; while (p < v) p <<= 1; return p; };
    const uint32_t np2x = np2(img.w), np2y = np2(img.h);
    uint32_t big = np2x > np2y ? np2x : np2y, levels = 1;
    while ((big >>= 1) != 0) levels++;                                    // trailing_zeros + 1
    if (levels > ARN_TEX_MAX_LEVELS) { if (err) *err = path + ": picture too large for the pyramid table"; return false; }
    t->channels = channels; t->n_levels = levels;
    texels->clear();
    for (uint32_t i = 0; i < levels; i++) {
        const uint32_t dx = (np2x >> i) > 1 ? (np2x >> i) : 1, dy = (np2y >> i) > 1 ? (np2y >> i) : 1;
        const Image8 lv = resize_lanczos3(img, dx, dy);
        t->level_w[i] = dx; t->level_h[i] = dy; t->level_offset[i] = (uint32_t)texels->size();
        for (size_t p = 0; p < (size_t)dx * dy; p++) {
            const uint8_t* s = &lv.px[p * lv.ch];
            uint8_t rgb[3];
            if (lv.ch <= 2) rgb[0] = rgb[1] = rgb[2] = s[0]; else { rgb[0] = s[0]; rgb[1] = s[1]; rgb[2] = s[2]; }      // to_rgb(): alpha dropped
            if (channels == 3) {
                for (int c = 0; c < 3; c++) { const float f = (float)rgb[c] / 255.f; texels->push_back((gamma ? inverse_gamma_correct(f) : f) * scale); }
            } else {
                // to_luma(): BT.709 weights of the `image` crate's rgb -> luma conversion, truncated to u8 (a gray file passes through)
                const uint8_t l = lv.ch <= 2 ? s[0] : (uint8_t)(0.2125f * (float)rgb[0] + 0.7154f * (float)rgb[1] + 0.0721f * (float)rgb[2]);
                const float f = (float)l / 255.f;
                texels->push_back((gamma ? inverse_gamma_correct(f) : f) * scale);
            }
        }
    }
    // MipMap::mean: sum of the level-0 texels in order, times 1 / count
    const size_t n0 = (size_t)t->level_w[0] * t->level_h[0];
    for (uint32_t c = 0; c < channels; c++) {
        float sum = 0.f;
        for (size_t p = 0; p < n0; p++) sum += (*texels)[p * channels + c];
        mean[c] = sum * (1.f / (float)n0);
    }
    return true;
}

}  // namespace arnhost
