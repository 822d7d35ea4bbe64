// PNG ingest without libpng (only zlib): restates what main.cpp:29-109 obtains from libpng for the image
// classes it configures -- 8/16-bit, gray / gray+alpha / RGB / RGBA / palette, non-interlaced -- and the
// gray conversion of main.cpp:92-99.  16-bit samples are stripped to their high byte (png_set_strip_16),
// palette entries expanded to RGB, alpha ignored.  Gray images are expanded to r=g=b (the reference reads
// them with a 4-byte stride it never configured; that is a bug, not behaviour to reproduce).
#include <zlib.h>

#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "host_internal.h"

namespace pcdh {

static uint32_t be32(const unsigned char *p) { return ((uint32_t)p[0] << 24) | ((uint32_t)p[1] << 16) | ((uint32_t)p[2] << 8) | p[3]; }

static int paeth(int a, int b, int c) {
    int p = a + b - c, pa = std::abs(p - a), pb = std::abs(p - b), pc = std::abs(p - c);
    if (pa <= pb && pa <= pc) return a;
    if (pb <= pc) return b;
    return c;
}

bool load_png_rgb(const std::string &path, int &width, int &height, std::vector<unsigned char> &rgb, std::string &err) {
    FILE *fp = fopen(path.c_str(), "rb");
    if (!fp) { err = "Failed to open PNG file."; return false; }  // main.cpp:32
    std::vector<unsigned char> file;
    unsigned char buf[1 << 16];
    size_t n;
    while ((n = fread(buf, 1, sizeof(buf), fp)) > 0) file.insert(file.end(), buf, buf + n);
    fclose(fp);
    static const unsigned char sig[8] = {0x89, 'P', 'N', 'G', 0x0d, 0x0a, 0x1a, 0x0a};
    if (file.size() < 8 || memcmp(file.data(), sig, 8) != 0) { err = "Error during PNG read initialization."; return false; }
    int bit_depth = 0, color_type = 0, interlace = 0;
    std::vector<unsigned char> idat, palette;
    size_t pos = 8;
    bool have_ihdr = false, have_iend = false;
    while (pos + 12 <= file.size()) {
        const uint32_t len = be32(&file[pos]);
        const char *type = (const char *)&file[pos + 4];
        if (pos + 12 + (size_t)len > file.size()) break;
        const unsigned char *data = &file[pos + 8];
        if (!memcmp(type, "IHDR", 4) && len >= 13) {
            width = (int)be32(data);
            height = (int)be32(data + 4);
            bit_depth = data[8]; color_type = data[9]; interlace = data[12];
            have_ihdr = true;
        } else if (!memcmp(type, "PLTE", 4)) {
            palette.assign(data, data + len);
        } else if (!memcmp(type, "IDAT", 4)) {
            idat.insert(idat.end(), data, data + len);
        } else if (!memcmp(type, "IEND", 4)) {
            have_iend = true;
            break;
        }
        pos += 12 + (size_t)len;
    }
    if (!have_ihdr || !have_iend || width <= 0 || height <= 0) { err = "Error during PNG read initialization."; return false; }
    if (interlace != 0) { err = "interlaced PNG files are not supported"; return false; }
    int channels;
    switch (color_type) {
        case 0: channels = 1; break;
        case 2: channels = 3; break;
        case 3: channels = 1; break;
        case 4: channels = 2; break;
        case 6: channels = 4; break;
        default: err = "unsupported PNG colour type"; return false;
    }
    if (!(bit_depth == 8 || bit_depth == 16 || ((color_type == 0 || color_type == 3) && (bit_depth == 1 || bit_depth == 2 || bit_depth == 4)))) {
        err = "unsupported PNG bit depth";
        return false;
    }
    const size_t bpp_bits = (size_t)channels * bit_depth;
    const size_t stride = ((size_t)width * bpp_bits + 7) / 8;
    const size_t bpp = bpp_bits >= 8 ? bpp_bits / 8 : 1;
    std::vector<unsigned char> raw((stride + 1) * (size_t)height);
    uLongf raw_len = (uLongf)raw.size();
    if (uncompress(raw.data(), &raw_len, idat.data(), (uLong)idat.size()) != Z_OK || raw_len != raw.size()) {
        err = "Error during PNG read initialization.";
        return false;
    }
    // undo the scanline filters in place
    std::vector<unsigned char> prev(stride, 0), cur(stride);
    rgb.assign((size_t)width * height * 3, 0);
    for (int y = 0; y < height; ++y) {
        const unsigned char *line = &raw[(stride + 1) * (size_t)y];
        const int ft = line[0];
        for (size_t i = 0; i < stride; ++i) {
            const int x = line[1 + i];
            const int a = i >= bpp ? cur[i - bpp] : 0, b = prev[i], c = i >= bpp ? prev[i - bpp] : 0;
            int v;
            switch (ft) {
                case 0: v = x; break;
                case 1: v = x + a; break;
                case 2: v = x + b; break;
                case 3: v = x + ((a + b) >> 1); break;
                case 4: v = x + paeth(a, b, c); break;
                default: err = "bad PNG filter"; return false;
            }
            cur[i] = (unsigned char)v;
        }
        for (int x = 0; x < width; ++x) {
            unsigned char r, g, b;
            auto sample = [&](int ch) -> unsigned char {  // 8-bit value of channel ch of pixel x
                if (bit_depth == 8) return cur[(size_t)x * channels + ch];
                if (bit_depth == 16) return cur[((size_t)x * channels + ch) * 2];  // strip_16: high byte
                const int per = 8 / bit_depth, idx = x;                            // packed gray / palette index
                const int shift = (per - 1 - idx % per) * bit_depth;
                return (unsigned char)((cur[idx / per] >> shift) & ((1 << bit_depth) - 1));
            };
            if (color_type == 3) {
                const size_t idx = sample(0);
                if (idx * 3 + 2 < palette.size()) { r = palette[idx * 3]; g = palette[idx * 3 + 1]; b = palette[idx * 3 + 2]; }
                else r = g = b = 0;
            } else if (color_type == 0 || color_type == 4) {
                unsigned char v = sample(0);
                if (bit_depth < 8) v = (unsigned char)(v * 255 / ((1 << bit_depth) - 1));  // expand_gray_1_2_4_to_8
                r = g = b = v;
            } else {
                r = sample(0); g = sample(1); b = sample(2);
            }
            unsigned char *o = &rgb[((size_t)y * width + x) * 3];
            o[0] = r; o[1] = g; o[2] = b;
        }
        prev.swap(cur);
    }
    return true;
}

// main.cpp:92-99
bool image_to_grid(const std::string &path, std::vector<std::vector<double>> &image_grid, std::string &err) {
    int w = 0, h = 0;
    std::vector<unsigned char> rgb;
    if (!load_png_rgb(path, w, h, rgb, err)) return false;
    image_grid.clear();
    for (int i = 0; i < h; ++i) {
        std::vector<double> row;
        row.reserve(w);
        for (int j = 0; j < w; ++j) {
            const unsigned char *px = &rgb[((size_t)i * w + j) * 3];
            const double r = px[0] / 255.0, g = px[1] / 255.0, b = px[2] / 255.0;
            const double gray = (0.299 * r) + (0.587 * g) + (0.114 * b);
            row.push_back(gray);
        }
        image_grid.push_back(row);
    }
    return true;
}

// main.cpp:13-27
void resize_image(const std::vector<std::vector<double>> &input_image, std::vector<std::vector<double>> &output_image,
                  int new_width, int new_height) {
    const int old_height = (int)input_image.size(), old_width = (int)input_image[0].size();
    output_image.assign(new_height, std::vector<double>(new_width));
    for (int y = 0; y < new_height; ++y)
        for (int x = 0; x < new_width; ++x) {
            const int src_x = x * old_width / new_width, src_y = y * old_height / new_height;
            output_image[y][x] = input_image[src_y][src_x];
        }
}

}  // namespace pcdh
