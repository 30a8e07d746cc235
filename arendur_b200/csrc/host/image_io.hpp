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
// Decoder: PNG only (8 / 16 bit gray, gray + alpha, RGB, RGBA; palette 1 .. 8 bit; no Adam7 interlace), inflate by zlib.
// Any other file, like a missing one, makes the caller fall back to the MTL constant — `image::open(..)` failing has the
// same effect in load_obj (component/mod.rs:88-95).
#pragma once
#include <zlib.h>
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
    if (!png_decode(path, &img, err)) return false;
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
