// PNG -> RGBA8 for the background loader.  The reference reads its backgrounds (backgrounds/*.png, 4096x2048 and
// 4000x2000 RGBA) through sf::Image::loadFromFile (/root/reference/graphics_settings.cpp:214-243, main.cpp:233-234); SFML
// is not part of this build, so this is a small decoder of its own on zlib: non-interlaced PNG, colour types
// 0/2/3/4/6, bit depths 8 and 16 (1/2/4 for grey and palette), tRNS for palettes.  Output is what getPixelsPtr() gives:
// tightly packed RGBA8, top row first.
#include <zlib.h>

#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <stdexcept>
#include <string>
#include <vector>

namespace grt {

static uint32_t be32(const uint8_t* p) { return ((uint32_t)p[0] << 24) | ((uint32_t)p[1] << 16) | ((uint32_t)p[2] << 8) | p[3]; }

static int paeth(int a, int b, int c) {
    int p = a + b - c, pa = abs(p - a), pb = abs(p - b), pc = abs(p - c);
    return (pa <= pb && pa <= pc) ? a : (pb <= pc ? b : c);
}

std::vector<uint8_t> load_png_rgba(const std::string& path, int* width, int* height) {
    FILE* f = fopen(path.c_str(), "rb");
    if (!f) throw std::runtime_error("cannot open " + path);
    std::vector<uint8_t> file;
    uint8_t buf[1 << 16];
    size_t n;
    while ((n = fread(buf, 1, sizeof buf, f)) > 0) file.insert(file.end(), buf, buf + n);
    fclose(f);
    static const uint8_t sig[8] = {0x89, 'P', 'N', 'G', 0x0d, 0x0a, 0x1a, 0x0a};
    if (file.size() < 8 || memcmp(file.data(), sig, 8) != 0) throw std::runtime_error(path + ": not a PNG file");

    uint32_t w = 0, h = 0;
    int depth = 0, ctype = -1, interlace = 0;
    std::vector<uint8_t> idat, palette, trns;
    size_t at = 8;
    bool end = false;
    while (!end && at + 12 <= file.size()) {
        uint32_t len = be32(&file[at]);
        const uint8_t* type = &file[at + 4];
        const uint8_t* data = &file[at + 8];
        if (at + 12 + (size_t)len > file.size()) throw std::runtime_error(path + ": truncated chunk");
        if (!memcmp(type, "IHDR", 4) && len >= 13) {
            w = be32(data); h = be32(data + 4); depth = data[8]; ctype = data[9]; interlace = data[12];
        } else if (!memcmp(type, "PLTE", 4)) palette.assign(data, data + len);
        else if (!memcmp(type, "tRNS", 4)) trns.assign(data, data + len);
        else if (!memcmp(type, "IDAT", 4)) idat.insert(idat.end(), data, data + len);
        else if (!memcmp(type, "IEND", 4)) end = true;
        at += 12 + (size_t)len;
    }
    if (w == 0 || h == 0 || w > 32768 || h > 32768) throw std::runtime_error(path + ": bad dimensions");
    if (interlace) throw std::runtime_error(path + ": interlaced PNG is not supported");
    int channels = ctype == 0 ? 1 : ctype == 2 ? 3 : ctype == 3 ? 1 : ctype == 4 ? 2 : ctype == 6 ? 4 : 0;
    if (!channels) throw std::runtime_error(path + ": unknown colour type");
    bool depth_ok = depth == 8 || (depth == 16 && ctype != 3) || ((depth == 1 || depth == 2 || depth == 4) && (ctype == 0 || ctype == 3));
    if (!depth_ok) throw std::runtime_error(path + ": unsupported bit depth");
    const size_t bits_pp = (size_t)channels * depth, stride = (w * bits_pp + 7) / 8, bpp = bits_pp >= 8 ? bits_pp / 8 : 1;

    std::vector<uint8_t> raw((stride + 1) * (size_t)h);
    uLongf raw_len = (uLongf)raw.size();
    int zr = uncompress(raw.data(), &raw_len, idat.data(), (uLong)idat.size());
    if (zr != Z_OK || raw_len != raw.size()) throw std::runtime_error(path + ": zlib inflate failed");

    // undo the scanline filters in place (PNG spec 9)
    std::vector<uint8_t> prev(stride, 0);
    for (uint32_t y = 0; y < h; y++) {
        uint8_t* line = &raw[(stride + 1) * (size_t)y];
        const int filter = line[0];
        uint8_t* cur = line + 1;
        for (size_t i = 0; i < stride; i++) {
            int a = i >= bpp ? cur[i - bpp] : 0, b = prev[i], c = i >= bpp ? prev[i - bpp] : 0, x = cur[i];
            switch (filter) {
                case 0: break;
                case 1: x += a; break;
                case 2: x += b; break;
                case 3: x += (a + b) / 2; break;
                case 4: x += paeth(a, b, c); break;
                default: throw std::runtime_error(path + ": bad filter type");
            }
            cur[i] = (uint8_t)x;
        }
        memcpy(prev.data(), cur, stride);
    }

    std::vector<uint8_t> out((size_t)w * h * 4);
    for (uint32_t y = 0; y < h; y++) {
        const uint8_t* cur = &raw[(stride + 1) * (size_t)y + 1];
        for (uint32_t x = 0; x < w; x++) {
            uint8_t* px = &out[((size_t)y * w + x) * 4];
            auto sample = [&](int ch) -> int {   // 8-bit value of channel ch of pixel x
                if (depth == 8) return cur[(size_t)x * channels + ch];
                if (depth == 16) return cur[((size_t)x * channels + ch) * 2];   // high byte
                size_t bit = (size_t)x * depth;
                int v = (cur[bit / 8] >> (8 - depth - (bit % 8))) & ((1 << depth) - 1);
                return ctype == 3 ? v : v * 255 / ((1 << depth) - 1);
            };
            switch (ctype) {
                case 0: px[0] = px[1] = px[2] = (uint8_t)sample(0); px[3] = 255; break;
                case 2: px[0] = (uint8_t)sample(0); px[1] = (uint8_t)sample(1); px[2] = (uint8_t)sample(2); px[3] = 255; break;
                case 3: {
                    size_t idx = (size_t)sample(0);
                    if (idx * 3 + 2 >= palette.size()) throw std::runtime_error(path + ": palette index out of range");
                    px[0] = palette[idx * 3]; px[1] = palette[idx * 3 + 1]; px[2] = palette[idx * 3 + 2];
                    px[3] = idx < trns.size() ? trns[idx] : 255;
                    break;
                }
                case 4: px[0] = px[1] = px[2] = (uint8_t)sample(0); px[3] = (uint8_t)sample(1); break;
                case 6: px[0] = (uint8_t)sample(0); px[1] = (uint8_t)sample(1); px[2] = (uint8_t)sample(2); px[3] = (uint8_t)sample(3); break;
            }
        }
    }
    *width = (int)w;
    *height = (int)h;
    return out;
}


// RGBA8 -> PNG (colour type 6, 8 bit, no interlace, filter 0 on every row, one zlib stream): the screenshot the reference writes with
// sf::Image::saveToFile (/root/reference/main.cpp:2783-2812).
void save_png_rgba(const std::string& path, const uint8_t* rgba, int width, int height) {
    if (!rgba || width <= 0 || height <= 0) throw std::runtime_error("bad image");
    std::vector<uint8_t> raw((size_t)height * ((size_t)width * 4 + 1));
    for (int y = 0; y < height; y++) {
        uint8_t* row = &raw[(size_t)y * ((size_t)width * 4 + 1)];
        row[0] = 0;
        memcpy(row + 1, rgba + (size_t)y * width * 4, (size_t)width * 4);
    }
    uLongf zlen = compressBound((uLong)raw.size());
    std::vector<uint8_t> z(zlen);
    if (compress2(z.data(), &zlen, raw.data(), (uLong)raw.size(), 6) != Z_OK) throw std::runtime_error("zlib compress failed");
    z.resize(zlen);
    FILE* f = fopen(path.c_str(), "wb");
    if (!f) throw std::runtime_error("cannot write " + path);
    auto put32 = [](uint8_t* p, uint32_t v) { p[0] = (uint8_t)(v >> 24); p[1] = (uint8_t)(v >> 16); p[2] = (uint8_t)(v >> 8); p[3] = (uint8_t)v; };
    auto chunk = [&](const char* type, const uint8_t* data, size_t len) {
        uint8_t hdr[8];
        put32(hdr, (uint32_t)len);
        memcpy(hdr + 4, type, 4);
        fwrite(hdr, 1, 8, f);
        if (len) fwrite(data, 1, len, f);
        uLong crc = crc32(0L, (const Bytef*)type, 4);
        if (len) crc = crc32(crc, data, (uInt)len);
        uint8_t c[4];
        put32(c, (uint32_t)crc);
        fwrite(c, 1, 4, f);
    };
    static const uint8_t sig[8] = {0x89, 'P', 'N', 'G', 0x0d, 0x0a, 0x1a, 0x0a};
    fwrite(sig, 1, 8, f);
    uint8_t ihdr[13];
    put32(ihdr, (uint32_t)width);
    put32(ihdr + 4, (uint32_t)height);
    ihdr[8] = 8; ihdr[9] = 6; ihdr[10] = 0; ihdr[11] = 0; ihdr[12] = 0;
    chunk("IHDR", ihdr, 13);
    chunk("IDAT", z.data(), z.size());
    chunk("IEND", nullptr, 0);
    if (fclose(f) != 0) throw std::runtime_error("cannot write " + path);
}

}  // namespace grt
