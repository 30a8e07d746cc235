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
// Decoders: PNG (8 / 16 bit gray, gray + alpha, RGB, RGBA; palette 1 .. 8 bit; no Adam7 interlace; inflate by zlib) and TGA
// (true colour, gray, colour-mapped; raw or run-length encoded).  Any other file, like a missing one, makes the caller fall back to the MTL constant — `image::open(..)` failing has the
// same effect in load_obj (component/mod.rs:88-95).
#pragma once
#include <zlib.h>
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
// image::open: by content for PNG, by extension for TGA (the format has no signature)
inline bool image_decode(const std::string& path, Image8* out, std::string* err) {
    auto ends_with = [&](const char* e) { const size_t n = std::strlen(e); if (path.size() < n) return false; for (size_t i = 0; i < n; i++) if (std::tolower((unsigned char)path[path.size() - n + i]) != e[i]) return false; return true; };
    if (ends_with(".tga")) return tga_decode(path, out, err);
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
