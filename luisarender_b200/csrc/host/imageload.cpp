// Image file readers for the `Image` texture plugin (reference: LoadedImage::load, src/util/imageio.cpp:347-470, which
// delegates to stb_image / tinyexr — neither is available here, so the decoders below are written against the file
// format specifications; JPEG, whose lossy decode only matches the reference's texels if it repeats stb's own integer IDCT and chroma
// filter, has its own file: jpegload.cpp).  What the readers reproduce from the reference is the STORAGE policy, because it decides the
// texel values the sampler sees:
//   * 8-bit files  -> BYTE1/2/4  : texel = x / 255        16-bit files -> SHORT1/2/4 : texel = x / 65535
//   * .hdr         -> HALF4      : RGBE decoded to float, then rounded to binary16 (imageio.cpp:383, 231-245)
//   * .exr         -> HALF or FLOAT by the file's first channel type; 1, 2 or 4 channels
//   * 3-channel sources are widened to 4 with alpha = 1 (stbi "desired channels" = 4)
// Row 0 of the result is the top row of the picture (PFM, stored bottom-up, is flipped).
#include <zlib.h>

#include <algorithm>
#include <cmath>
#include <cstring>
#include <fstream>
#include <sstream>

#include "imageio.h"

namespace lrh {

namespace {

[[noreturn]] void fail(const std::filesystem::path &p, const std::string &why) {
    throw std::runtime_error("Failed to load image '" + p.string() + "': " + why + ".");
}

std::vector<uint8_t> read_file(const std::filesystem::path &p) {
    std::ifstream f{p, std::ios::binary};
    if (!f) fail(p, "cannot open file");
    std::vector<uint8_t> data((std::istreambuf_iterator<char>(f)), std::istreambuf_iterator<char>());
    return data;
}

// binary16 <-> binary32, round-to-nearest-even (what luisa::float_to_half / half_to_float do)
uint16_t float_to_half(float f) {
    uint32_t x;
    std::memcpy(&x, &f, 4);
    uint32_t sign = (x >> 16u) & 0x8000u;
    int32_t exp = static_cast<int32_t>((x >> 23u) & 0xffu) - 127 + 15;
    uint32_t man = x & 0x7fffffu;
    if (((x >> 23u) & 0xffu) == 0xffu) return static_cast<uint16_t>(sign | 0x7c00u | (man ? 0x200u : 0u));
    if (exp >= 31) return static_cast<uint16_t>(sign | 0x7c00u);
    if (exp <= 0) {
        if (exp < -10) return static_cast<uint16_t>(sign);
        man |= 0x800000u;
        uint32_t shift = static_cast<uint32_t>(14 - exp);
        uint32_t half = man >> shift;
        uint32_t rem = man & ((1u << shift) - 1u), mid = 1u << (shift - 1u);
        if (rem > mid || (rem == mid && (half & 1u))) half++;
        return static_cast<uint16_t>(sign | half);
    }
    uint32_t half = (static_cast<uint32_t>(exp) << 10u) | (man >> 13u);
    uint32_t rem = man & 0x1fffu;
    if (rem > 0x1000u || (rem == 0x1000u && (half & 1u))) half++;
    return static_cast<uint16_t>(sign | half);
}
float half_to_float(uint16_t h) {
    uint32_t sign = (h & 0x8000u) << 16u, exp = (h >> 10u) & 0x1fu, man = h & 0x3ffu, x;
    if (exp == 0u) {
        if (man == 0u) {
            x = sign;
        } else {
            int e = -1;
            do {
                e++;
                man <<= 1u;
            } while (!(man & 0x400u));
            x = sign | (static_cast<uint32_t>(127 - 15 - e) << 23u) | ((man & 0x3ffu) << 13u);
        }
    } else if (exp == 31u) {
        x = sign | 0x7f800000u | (man << 13u);
    } else {
        x = sign | ((exp + 127u - 15u) << 23u) | (man << 13u);
    }
    float f;
    std::memcpy(&f, &x, 4);
    return f;
}

// widen `nc`-channel interleaved samples to the storage channel count (1, 2 or 4) as RGBA floats
LoadedImage finish(uint32_t w, uint32_t h, uint32_t nc, const std::vector<float> &samples) {
    LoadedImage img;
    img.width = w;
    img.height = h;
    img.channels = nc >= 3u ? 4u : nc;
    img.rgba.assign(static_cast<size_t>(w) * h * 4u, 0.f);
    for (size_t i = 0; i < static_cast<size_t>(w) * h; i++) {
        const float *s = samples.data() + i * nc;
        float *d = img.rgba.data() + i * 4u;
        if (nc == 1u) {
            d[0] = s[0];// a 1-channel texel reads back as (x, 0, 0, 1) (cpu_texture.h read_pixel), the consumers use .x / .xxx
            d[3] = 1.f;
        } else if (nc == 2u) {
            d[0] = s[0];
            d[1] = s[1];
            d[3] = 1.f;
        } else {
            d[0] = s[0];
            d[1] = s[1];
            d[2] = s[2];
            d[3] = nc == 4u ? s[3] : 1.f;
        }
    }
    return img;
}

// ---- PNG (ISO/IEC 15948): non-interlaced, colour types 0/2/3/4/6, bit depths 1-16 ---------------------------------
uint32_t be32(const uint8_t *p) { return (uint32_t{p[0]} << 24u) | (uint32_t{p[1]} << 16u) | (uint32_t{p[2]} << 8u) | p[3]; }

LoadedImage load_png(const std::filesystem::path &path, const std::vector<uint8_t> &d) {
    static const uint8_t sig[8] = {0x89, 'P', 'N', 'G', 0x0d, 0x0a, 0x1a, 0x0a};
    if (d.size() < 8 || std::memcmp(d.data(), sig, 8) != 0) fail(path, "not a PNG file");
    uint32_t w = 0, h = 0, depth = 0, ctype = 0, interlace = 0;
    std::vector<uint8_t> idat, plte, trns;
    size_t pos = 8;
    bool end = false;
    while (!end && pos + 12 <= d.size()) {
        uint32_t len = be32(&d[pos]);
        if (pos + 12 + len > d.size()) fail(path, "truncated PNG chunk");
        std::string type(reinterpret_cast<const char *>(&d[pos + 4]), 4);
        const uint8_t *body = &d[pos + 8];
        if (type == "IHDR") {
            if (len < 13) fail(path, "bad IHDR");
            w = be32(body);
            h = be32(body + 4);
            depth = body[8];
            ctype = body[9];
            interlace = body[12];
        } else if (type == "PLTE") {
            plte.assign(body, body + len);
        } else if (type == "tRNS") {
            trns.assign(body, body + len);
        } else if (type == "IDAT") {
            idat.insert(idat.end(), body, body + len);
        } else if (type == "IEND") {
            end = true;
        }
        pos += 12 + len;
    }
    if (w == 0 || h == 0 || idat.empty()) fail(path, "missing IHDR / IDAT");
    if (interlace != 0) fail(path, "interlaced PNG files are not supported");
    uint32_t samples_per_pixel = ctype == 0 ? 1u : ctype == 2 ? 3u : ctype == 3 ? 1u : ctype == 4 ? 2u : ctype == 6 ? 4u : 0u;
    if (samples_per_pixel == 0u || !(depth == 1 || depth == 2 || depth == 4 || depth == 8 || depth == 16)) fail(path, "unsupported PNG colour type / depth");
    size_t bits_per_pixel = static_cast<size_t>(samples_per_pixel) * depth;
    size_t stride = (static_cast<size_t>(w) * bits_per_pixel + 7u) / 8u;
    size_t bpp = std::max<size_t>(1u, bits_per_pixel / 8u);
    std::vector<uint8_t> raw((stride + 1u) * h);
    uLongf raw_len = static_cast<uLongf>(raw.size());
    if (uncompress(raw.data(), &raw_len, idat.data(), static_cast<uLong>(idat.size())) != Z_OK || raw_len != raw.size())
        fail(path, "zlib inflate failed");
    // unfilter (PNG spec 9.2)
    std::vector<uint8_t> pix(stride * h);
    for (uint32_t y = 0; y < h; y++) {
        const uint8_t *in = &raw[(stride + 1u) * y];
        uint8_t ft = in[0];
        in++;
        uint8_t *out = &pix[stride * y];
        const uint8_t *up = y ? &pix[stride * (y - 1u)] : nullptr;
        for (size_t x = 0; x < stride; x++) {
            int a = x >= bpp ? out[x - bpp] : 0, b = up ? up[x] : 0, c = (up && x >= bpp) ? up[x - bpp] : 0;
            int v = in[x];
            switch (ft) {
                case 0: break;
                case 1: v += a; break;
                case 2: v += b; break;
                case 3: v += (a + b) / 2; break;
                case 4: {
                    int p = a + b - c, pa = std::abs(p - a), pb = std::abs(p - b), pc = std::abs(p - c);
                    v += (pa <= pb && pa <= pc) ? a : (pb <= pc ? b : c);
                    break;
                }
                default: fail(path, "bad PNG filter type");
            }
            out[x] = static_cast<uint8_t>(v);
        }
    }
    // samples -> floats.  stb_image expands low bit depths and palettes to 8 bits; 16-bit files stay 16-bit.
    const bool is16 = depth == 16;
    const float norm = is16 ? 65535.f : 255.f;
    uint32_t nc = ctype == 3 ? (trns.empty() ? 3u : 4u) : samples_per_pixel;
    std::vector<float> samples(static_cast<size_t>(w) * h * nc);
    auto sample_at = [&](uint32_t y, size_t index) -> uint32_t {// index = x * samples_per_pixel + s
        const uint8_t *row = &pix[stride * y];
        if (depth == 16) return (uint32_t{row[index * 2u]} << 8u) | row[index * 2u + 1u];
        if (depth == 8) return row[index];
        size_t bit = index * depth;
        uint32_t v = (row[bit / 8u] >> (8u - depth - (bit % 8u))) & ((1u << depth) - 1u);
        return v;
    };
    for (uint32_t y = 0; y < h; y++) {
        for (uint32_t x = 0; x < w; x++) {
            float *o = &samples[(static_cast<size_t>(y) * w + x) * nc];
            if (ctype == 3) {
                uint32_t idx = sample_at(y, x);
                if ((idx + 1u) * 3u > plte.size()) fail(path, "palette index out of range");
                for (int k = 0; k < 3; k++) o[k] = static_cast<float>(plte[idx * 3u + k]) / 255.f;
                if (nc == 4u) o[3] = (idx < trns.size() ? static_cast<float>(trns[idx]) : 255.f) / 255.f;
            } else {
                for (uint32_t s = 0; s < samples_per_pixel; s++) {
                    uint32_t v = sample_at(y, static_cast<size_t>(x) * samples_per_pixel + s);
                    if (depth < 8) v = v * 255u / ((1u << depth) - 1u);// stb_image scales 1/2/4-bit grey to 0..255
                    o[s] = static_cast<float>(v) / norm;
                }
            }
        }
    }
    return finish(w, h, nc, samples);
}

// ---- Netpbm P5/P6 (binary grey / RGB, maxval < 65536) and PFM ---------------------------------------------------
std::string next_token(const std::vector<uint8_t> &d, size_t &pos) {
    for (;;) {
        while (pos < d.size() && std::isspace(d[pos])) pos++;
        if (pos < d.size() && d[pos] == '#') {
            while (pos < d.size() && d[pos] != '\n') pos++;
        } else {
            break;
        }
    }
    std::string t;
    while (pos < d.size() && !std::isspace(d[pos])) t.push_back(static_cast<char>(d[pos++]));
    return t;
}

LoadedImage load_pnm(const std::filesystem::path &path, const std::vector<uint8_t> &d) {
    size_t pos = 0;
    std::string magic = next_token(d, pos);
    if (magic == "PF" || magic == "Pf") {
        uint32_t nc = magic == "PF" ? 3u : 1u;
        uint32_t w = static_cast<uint32_t>(std::stoul(next_token(d, pos))), h = static_cast<uint32_t>(std::stoul(next_token(d, pos)));
        float scale = std::stof(next_token(d, pos));
        pos++;// the single whitespace after the scale
        size_t n = static_cast<size_t>(w) * h * nc;
        if (pos + n * 4u > d.size()) fail(path, "truncated PFM");
        std::vector<float> samples(n);
        for (uint32_t y = 0; y < h; y++) {// stored bottom-up
            const uint8_t *row = &d[pos + static_cast<size_t>(h - 1u - y) * w * nc * 4u];
            for (size_t i = 0; i < static_cast<size_t>(w) * nc; i++) {
                uint8_t b[4] = {row[i * 4u], row[i * 4u + 1u], row[i * 4u + 2u], row[i * 4u + 3u]};
                if (scale > 0.f) std::swap(b[0], b[3]), std::swap(b[1], b[2]);// positive scale = big endian
                float f;
                std::memcpy(&f, b, 4);
                samples[static_cast<size_t>(y) * w * nc + i] = f;
            }
        }
        return finish(w, h, nc, samples);
    }
    if (magic != "P5" && magic != "P6") fail(path, "unsupported Netpbm variant '" + magic + "'");
    uint32_t nc = magic == "P6" ? 3u : 1u;
    uint32_t w = static_cast<uint32_t>(std::stoul(next_token(d, pos))), h = static_cast<uint32_t>(std::stoul(next_token(d, pos)));
    uint32_t maxval = static_cast<uint32_t>(std::stoul(next_token(d, pos)));
    pos++;
    size_t n = static_cast<size_t>(w) * h * nc, bytes = maxval > 255u ? 2u : 1u;
    if (maxval == 0u || maxval > 65535u || pos + n * bytes > d.size()) fail(path, "bad / truncated Netpbm data");
    std::vector<float> samples(n);
    for (size_t i = 0; i < n; i++) {
        uint32_t v = bytes == 2u ? (uint32_t{d[pos + i * 2u]} << 8u) | d[pos + i * 2u + 1u] : d[pos + i];
        samples[i] = static_cast<float>(v) / static_cast<float>(maxval);
    }
    return finish(w, h, nc, samples);
}

// ---- Radiance RGBE (.hdr): new-style RLE and flat scanlines, -Y h +X w orientation --------------------------------------
LoadedImage load_hdr(const std::filesystem::path &path, const std::vector<uint8_t> &d) {
    size_t pos = 0;
    auto line = [&]() {
        std::string l;
        while (pos < d.size() && d[pos] != '\n') l.push_back(static_cast<char>(d[pos++]));
        pos++;
        return l;
    };
    std::string first = line();
    if (first.rfind("#?", 0) != 0) fail(path, "not a Radiance HDR file");
    for (;;) {
        if (pos >= d.size()) fail(path, "truncated HDR header");
        std::string l = line();
        if (l.empty()) break;
    }
    std::string res = line();
    uint32_t w = 0, h = 0;
    {
        std::istringstream ss{res};
        std::string ys, xs;
        ss >> ys >> h >> xs >> w;
        if (ys != "-Y" || xs != "+X" || w == 0 || h == 0) fail(path, "unsupported HDR orientation '" + res + "'");
    }
    std::vector<uint8_t> scan(static_cast<size_t>(w) * 4u);
    std::vector<float> samples(static_cast<size_t>(w) * h * 3u);
    for (uint32_t y = 0; y < h; y++) {
        if (pos + 4 > d.size()) fail(path, "truncated HDR data");
        if (w >= 8 && w < 32768 && d[pos] == 2 && d[pos + 1] == 2 && !(d[pos + 2] & 0x80)) {
            if (((uint32_t{d[pos + 2]} << 8u) | d[pos + 3]) != w) fail(path, "bad HDR scanline width");
            pos += 4;
            for (int c = 0; c < 4; c++) {
                uint32_t x = 0;
                while (x < w) {
                    if (pos >= d.size()) fail(path, "truncated HDR RLE data");
                    uint8_t count = d[pos++];
                    if (count > 128) {
                        count -= 128;
                        if (x + count > w || pos >= d.size()) fail(path, "bad HDR RLE run");
                        uint8_t v = d[pos++];
                        for (uint8_t k = 0; k < count; k++) scan[(x++) * 4u + c] = v;
                    } else {
                        if (count == 0 || x + count > w || pos + count > d.size()) fail(path, "bad HDR RLE literal");
                        for (uint8_t k = 0; k < count; k++) scan[(x++) * 4u + c] = d[pos++];
                    }
                }
            }
        } else {
            if (pos + static_cast<size_t>(w) * 4u > d.size()) fail(path, "truncated HDR data");
            std::memcpy(scan.data(), &d[pos], static_cast<size_t>(w) * 4u);
            pos += static_cast<size_t>(w) * 4u;
        }
        for (uint32_t x = 0; x < w; x++) {
            const uint8_t *p = &scan[x * 4u];
            float *o = &samples[(static_cast<size_t>(y) * w + x) * 3u];
            if (p[3] == 0) {
                o[0] = o[1] = o[2] = 0.f;
            } else {
                float f = std::ldexp(1.0f, static_cast<int>(p[3]) - (128 + 8));// stb_image's hdr_convert
                for (int k = 0; k < 3; k++) o[k] = half_to_float(float_to_half(static_cast<float>(p[k]) * f));// stored as HALF4
            }
        }
    }
    return finish(w, h, 3u, samples);
}

// ---- OpenEXR: single-part scan-line files, NO_COMPRESSION / ZIPS / ZIP, HALF / FLOAT channels ---------------------------
LoadedImage load_exr(const std::filesystem::path &path, const std::vector<uint8_t> &d) {
    if (d.size() < 8 || be32(d.data()) != 0x762f3101u) fail(path, "not an OpenEXR file");
    uint32_t version;
    std::memcpy(&version, &d[4], 4);
    if (version & 0x1e00u) fail(path, "tiled / multi-part / deep EXR files are not supported");
    size_t pos = 8;
    auto cstr = [&]() {
        std::string s;
        while (pos < d.size() && d[pos] != 0) s.push_back(static_cast<char>(d[pos++]));
        pos++;
        return s;
    };
    struct Channel {
        std::string name;
        int32_t type;
    };
    std::vector<Channel> channels;
    int32_t compression = -1, dw[4] = {0, 0, -1, -1};
    for (;;) {
        std::string name = cstr();
        if (name.empty()) break;
        std::string type = cstr();
        if (pos + 4 > d.size()) fail(path, "truncated EXR header");
        int32_t size;
        std::memcpy(&size, &d[pos], 4);
        pos += 4;
        size_t start = pos;
        if (size < 0 || start + static_cast<size_t>(size) > d.size()) fail(path, "truncated EXR attribute");
        const size_t end = start + static_cast<size_t>(size);
        if (name == "channels") {
            while (pos < end && d[pos] != 0) {
                Channel c;
                c.name = cstr();
                if (pos + 16 > end) fail(path, "truncated EXR channel list");
                std::memcpy(&c.type, &d[pos], 4);
                pos += 16;// pixel type, pLinear + reserved, xSampling, ySampling
                channels.push_back(c);
            }
        } else if (name == "compression") {
            if (size < 1) fail(path, "truncated EXR compression attribute");
            compression = d[pos];
        } else if (name == "dataWindow") {
            if (size < 16) fail(path, "truncated EXR dataWindow attribute");
            std::memcpy(dw, &d[pos], 16);
        }
        pos = start + static_cast<size_t>(size);
    }
    if (channels.empty() || dw[2] < dw[0] || dw[3] < dw[1]) fail(path, "EXR header without channels / data window");
    if (compression != 0 && compression != 2 && compression != 3) fail(path, "only uncompressed / ZIPS / ZIP EXR files are supported");
    for (auto &c : channels)
        if (c.type != 1 && c.type != 2) fail(path, "only HALF / FLOAT EXR channels are supported");
    // window extents in 64 bits: a hostile header must not wrap the int32 subtraction or the w * h * channels products below
    const int64_t w64 = static_cast<int64_t>(dw[2]) - dw[0] + 1, h64 = static_cast<int64_t>(dw[3]) - dw[1] + 1;
    if (w64 > 65536 || h64 > 65536 || w64 * h64 > (int64_t{1} << 28)) fail(path, "EXR data window too large");
    const uint32_t w = static_cast<uint32_t>(w64), h = static_cast<uint32_t>(h64);
    const uint32_t lines_per_block = compression == 3 ? 16u : 1u;
    const uint32_t blocks = (h + lines_per_block - 1u) / lines_per_block;
    if (pos + static_cast<size_t>(blocks) * 8u > d.size()) fail(path, "truncated EXR offset table");
    std::vector<uint64_t> offsets(blocks);
    std::memcpy(offsets.data(), &d[pos], static_cast<size_t>(blocks) * 8u);
    // channel -> output slot: channels are stored alphabetically (A, B, G, R); luminance-only files have "Y"
    const uint32_t nfile = static_cast<uint32_t>(channels.size());
    uint32_t nc = nfile == 1u ? 1u : nfile == 2u ? 2u : 4u;
    auto slot_of = [&](uint32_t i) -> int {
        const std::string &n = channels[i].name;
        if (nc == 1u) return 0;
        if (nc == 2u) return static_cast<int>(i);
        if (n == "R") return 0;
        if (n == "G") return 1;
        if (n == "B") return 2;
        if (n == "A") return 3;
        return -1;
    };
    size_t line_bytes = 0;
    for (auto &c : channels) line_bytes += static_cast<size_t>(w) * (c.type == 1 ? 2u : 4u);
    const bool as_half = channels[0].type == 1;// storage follows the first channel's type (imageio.cpp:355-376)
    std::vector<float> samples(static_cast<size_t>(w) * h * nc, 0.f);
    if (nc == 4u)
        for (size_t i = 0; i < static_cast<size_t>(w) * h; i++) samples[i * 4u + 3u] = 1.f;
    std::vector<uint8_t> block, tmp;
    for (uint32_t b = 0; b < blocks; b++) {
        size_t p = offsets[b];
        if (p + 8 > d.size()) fail(path, "bad EXR block offset");
        int32_t y0, data_size;
        std::memcpy(&y0, &d[p], 4);
        std::memcpy(&data_size, &d[p + 4], 4);
        p += 8;
        if (data_size < 0 || p + static_cast<size_t>(data_size) > d.size()) fail(path, "truncated EXR block");
        // the block's first scan line comes from the file: it has to lie inside the data window, on a block boundary
        const int64_t row0 = static_cast<int64_t>(y0) - dw[1];
        if (row0 < 0 || row0 >= static_cast<int64_t>(h) || row0 % lines_per_block != 0) fail(path, "EXR block outside the data window");
        uint32_t lines = std::min(lines_per_block, h - static_cast<uint32_t>(row0));
        size_t expect = line_bytes * lines;
        block.resize(expect);
        if (compression == 0 || static_cast<size_t>(data_size) == expect) {
            if (static_cast<size_t>(data_size) < expect) fail(path, "truncated EXR block");
            std::memcpy(block.data(), &d[p], expect);
        } else {
            tmp.resize(expect);
            uLongf out_len = static_cast<uLongf>(expect);
            if (uncompress(tmp.data(), &out_len, &d[p], static_cast<uLong>(data_size)) != Z_OK || out_len != expect) fail(path, "EXR zlib inflate failed");
            for (size_t i = 1; i < expect; i++) tmp[i] = static_cast<uint8_t>(tmp[i - 1] + tmp[i] - 128);// predictor
            size_t half = (expect + 1u) / 2u;// de-interleave
            for (size_t i = 0; i < expect; i++) block[i] = (i & 1u) ? tmp[half + i / 2u] : tmp[i / 2u];
        }
        const uint8_t *q = block.data();
        for (uint32_t l = 0; l < lines; l++) {
            uint32_t y = static_cast<uint32_t>(row0) + l;
            for (uint32_t c = 0; c < nfile; c++) {
                int slot = slot_of(c);
                for (uint32_t x = 0; x < w; x++) {
                    float v;
                    if (channels[c].type == 1) {
                        uint16_t hbits;
                        std::memcpy(&hbits, q, 2);
                        q += 2;
                        v = half_to_float(hbits);
                    } else {
                        std::memcpy(&v, q, 4);
                        q += 4;
                        if (as_half) v = half_to_float(float_to_half(v));
                    }
                    if (slot >= 0) samples[(static_cast<size_t>(y) * w + x) * nc + static_cast<uint32_t>(slot)] = v;
                }
            }
        }
    }
    return finish(w, h, nc, samples);
}


// ---- Windows BMP: uncompressed (BI_RGB) and BI_BITFIELDS pictures, 1 / 4 / 8-bit palettes, 16 / 24 / 32 bits per pixel, core (12)
// and info (40 / 56 / 108 / 124 byte) headers, bottom-up or top-down.  What the reference sees for such a file is what stb_image
// hands it with four channels requested (imageio.cpp:522-537: an RGB picture has >= 3 channels, so BYTE4): 8-bit RGBA, alpha 255
// unless the picture has an alpha mask - the implicit 0xff000000 of a 32-bit BI_RGB picture counts, but an alpha channel that is
// zero everywhere is taken to be absent - and narrow bit fields widened by bit replication.  RLE-compressed pictures are refused
// (as there).
struct ByteReader {
    const std::vector<uint8_t> &d;
    const std::filesystem::path &path;
    size_t pos{0};
    uint8_t u8() {
        if (pos >= d.size()) fail(path, "truncated file");
        return d[pos++];
    }
    uint32_t u16() { uint32_t a = u8(); return a | (uint32_t{u8()} << 8u); }
    uint32_t u32() { uint32_t a = u16(); return a | (u16() << 16u); }
    void skip(size_t n) {
        if (pos + n > d.size()) fail(path, "truncated file");
        pos += n;
    }
    void skip_padding(size_t n) { pos = std::min(pos + n, d.size()); }// row padding: the last row's may be missing
};

// an n-bit field value (n in 1..8) widened to 8 bits by repeating its bit pattern: 5 bits abcde -> abcdeabc
uint32_t replicate_bits(uint32_t v, uint32_t bits) {
    uint32_t out = 0u;
    for (int shift = 8 - static_cast<int>(bits); shift > -static_cast<int>(bits); shift -= static_cast<int>(bits))
        out |= shift >= 0 ? v << shift : v >> -shift;
    return out & 0xffu;
}
struct BitField {
    uint32_t mask{0u}, low{0u}, bits{0u};
    explicit BitField(uint32_t m) : mask{m} {
        if (m == 0u) return;
        while (!((m >> low) & 1u)) low++;
        for (uint32_t b = low; b < 32u && ((m >> b) & 1u); b++) bits++;
    }
    bool contiguous() const { return mask == 0u || (bits < 32u && mask == (((1u << bits) - 1u) << low)); }
    uint32_t extract(uint32_t v) const { return replicate_bits((v & mask) >> low, bits); }
};

LoadedImage load_bmp(const std::filesystem::path &path, const std::vector<uint8_t> &d) {
    ByteReader r{d, path};
    if (r.u8() != 'B' || r.u8() != 'M') fail(path, "not a BMP file");
    r.skip(8);// file size, reserved
    const uint32_t data_offset = r.u32(), header_size = r.u32();
    if (header_size != 12u && header_size != 40u && header_size != 56u && header_size != 108u && header_size != 124u)
        fail(path, "unsupported BMP header");
    int64_t w, h;
    if (header_size == 12u) {
        w = r.u16();
        h = r.u16();
    } else {
        w = static_cast<int32_t>(r.u32());
        h = static_cast<int32_t>(r.u32());
    }
    if (r.u16() != 1u) fail(path, "bad BMP plane count");
    const uint32_t bpp = r.u16();
    uint32_t compression = 0u, mr = 0u, mg = 0u, mb = 0u, ma = 0u;
    size_t header_end = 14u + header_size;
    bool implicit_alpha = false;
    if (header_size != 12u) {
        compression = r.u32();
        if (compression == 1u || compression == 2u) fail(path, "RLE-compressed BMP files are not supported");
        if (compression > 3u) fail(path, "unsupported BMP compression");
        if (compression == 3u && bpp != 16u && bpp != 32u) fail(path, "BMP bit fields need 16 or 32 bits per pixel");
        r.skip(20);// image size, resolution, colour counts
        auto defaults = [&] {// BI_RGB: 5-5-5 for 16 bits, 8-8-8-8 for 32 bits (whose alpha may turn out to be unused)
            mr = mg = mb = ma = 0u;
            if (bpp == 16u) { mr = 31u << 10u; mg = 31u << 5u; mb = 31u; }
            if (bpp == 32u) { mr = 0xffu << 16u; mg = 0xffu << 8u; mb = 0xffu; ma = 0xffu << 24u; implicit_alpha = true; }
        };
        if (header_size == 40u || header_size == 56u) {
            if (header_size == 56u) r.skip(16);
            if (bpp == 16u || bpp == 32u) {
                if (compression == 0u) {
                    defaults();
                } else {// the three masks follow a 40-byte header
                    mr = r.u32(); mg = r.u32(); mb = r.u32();
                    header_end += 12u;
                    if (mr == mg && mg == mb) fail(path, "bad BMP bit masks");
                }
            }
        } else {
            mr = r.u32(); mg = r.u32(); mb = r.u32(); ma = r.u32();
            if (compression != 3u) defaults();
            r.skip(header_size == 124u ? 68u : 52u);// colour space, and the profile fields of the V5 header
        }
    }
    const bool flip = h > 0;
    if (h < 0) h = -h;
    if (w <= 0 || h == 0 || w > (1 << 24) || h > (1 << 24)) fail(path, "bad BMP size");
    const uint32_t width = static_cast<uint32_t>(w), height = static_cast<uint32_t>(h);
    std::vector<float> samples(static_cast<size_t>(width) * height * 4u);
    auto put = [&](uint32_t x, uint32_t y, uint32_t red, uint32_t green, uint32_t blue, uint32_t alpha) {
        float *px = &samples[(static_cast<size_t>(flip ? height - 1u - y : y) * width + x) * 4u];
        px[0] = static_cast<float>(red) / 255.f;
        px[1] = static_cast<float>(green) / 255.f;
        px[2] = static_cast<float>(blue) / 255.f;
        px[3] = static_cast<float>(alpha) / 255.f;
    };
    if (bpp < 16u) {
        if (bpp != 1u && bpp != 4u && bpp != 8u) fail(path, "unsupported BMP bit depth");
        const size_t entry = header_size == 12u ? 3u : 4u;
        if (data_offset < header_end) fail(path, "bad BMP data offset");
        // (Core-header files: stb_image sizes the palette as (offset - 14 - 24) / 3, four entries short, and looks the missing ones
        // up in an uninitialised table - the reference's texels for those indices change from run to run, observed as the real colours
        // in one render and black in another.  This reader takes the whole palette; the fixtures stay below stb's count.)
        const size_t entries = (data_offset - header_end) / entry;
        if (entries == 0u || entries > 256u) fail(path, "bad BMP palette");
        uint8_t palette[256][3] = {};
        for (size_t i = 0; i < entries; i++) {
            palette[i][2] = r.u8(); palette[i][1] = r.u8(); palette[i][0] = r.u8();
            if (entry == 4u) r.u8();
        }
        r.pos = data_offset;
        const size_t row_bytes = (static_cast<size_t>(width) * bpp + 7u) / 8u, padding = (4u - row_bytes % 4u) % 4u;
        for (uint32_t y = 0; y < height; y++) {
            const size_t row = r.pos;
            r.skip(row_bytes);
            r.skip_padding(padding);
            for (uint32_t x = 0; x < width; x++) {
                const uint8_t byte = d[row + static_cast<size_t>(x) * bpp / 8u];
                const uint32_t index = bpp == 8u ? byte : bpp == 4u ? (x & 1u ? byte & 15u : byte >> 4u) : (byte >> (7u - (x & 7u))) & 1u;
                put(x, y, palette[index][0], palette[index][1], palette[index][2], 255u);
            }
        }
        return finish(width, height, 4u, samples);
    }
    if (bpp != 16u && bpp != 24u && bpp != 32u) fail(path, "unsupported BMP bit depth");
    if (data_offset < header_end || data_offset - header_end > 1024u) fail(path, "bad BMP data offset");
    r.pos = data_offset;
    const BitField fr{mr}, fg{mg}, fb{mb}, fa{ma};
    if (bpp != 24u) {
        if (!mr || !mg || !mb || fr.bits > 8u || fg.bits > 8u || fb.bits > 8u || fa.bits > 8u ||
            !fr.contiguous() || !fg.contiguous() || !fb.contiguous() || !fa.contiguous())
            fail(path, "bad BMP bit masks");
    }
    const size_t row_bytes = static_cast<size_t>(width) * (bpp / 8u), padding = (4u - row_bytes % 4u) % 4u;
    uint32_t alpha_seen = implicit_alpha ? 0u : 255u;
    for (uint32_t y = 0; y < height; y++) {
        for (uint32_t x = 0; x < width; x++) {
            if (bpp == 24u) {
                const uint32_t blue = r.u8(), green = r.u8(), red = r.u8();
                put(x, y, red, green, blue, 255u);
            } else {
                const uint32_t v = bpp == 16u ? r.u16() : r.u32();
                const uint32_t alpha = ma ? fa.extract(v) : 255u;
                alpha_seen |= alpha;
                put(x, y, fr.extract(v), fg.extract(v), fb.extract(v), alpha);
            }
        }
        r.skip_padding(padding);
    }
    if (alpha_seen == 0u)// a 32-bit BI_RGB picture whose fourth byte is zero everywhere has no alpha channel
        for (size_t i = 3; i < samples.size(); i += 4u) samples[i] = 1.f;
    return finish(width, height, 4u, samples);
}

// ---- Truevision TGA: colour-mapped (1 / 9), true-colour (2 / 10) and grey (3 / 11) pictures, raw or run-length encoded, 8 / 15 /
// 16 / 24 / 32 bits; the origin bit of the descriptor decides the row order.  Channel count as stb_image reports it - which picks
// the reference's storage (imageio.cpp:522-530): 8-bit grey -> 1 channel, 16-bit grey -> grey + alpha, 15 / 16-bit colour -> RGB with
// each 5-bit field scaled as v * 255 / 31, 24 / 32 bits -> RGB(A) stored blue first; for colour-mapped files the palette's entry size
// decides.  Run-length packets may run across rows.
LoadedImage load_tga(const std::filesystem::path &path, const std::vector<uint8_t> &d) {
    ByteReader r{d, path};
    const uint32_t id_length = r.u8(), map_type = r.u8(), image_type = r.u8();
    const uint32_t map_start = r.u16(), map_length = r.u16(), map_bits = r.u8();
    r.skip(4);// x / y origin
    const uint32_t width = r.u16(), height = r.u16(), pixel_bits = r.u8(), descriptor = r.u8();
    const bool rle = image_type >= 8u, grey = (image_type & 7u) == 3u, mapped = map_type == 1u;
    if (map_type > 1u || (mapped ? (image_type & 7u) != 1u : ((image_type & 7u) != 2u && !grey)) || (image_type & ~15u) != 0u)
        fail(path, "unsupported TGA image type");
    if (width == 0u || height == 0u) fail(path, "bad TGA size");
    if (mapped && pixel_bits != 8u && pixel_bits != 16u) fail(path, "unsupported TGA index size");
    const uint32_t colour_bits = mapped ? map_bits : pixel_bits;
    uint32_t nc = 0u;
    bool packed16 = false;
    switch (colour_bits) {
        case 8: nc = 1u; break;
        case 16: if (grey) { nc = 2u; break; } [[fallthrough]];
        case 15: nc = 3u; packed16 = true; break;
        case 24: nc = 3u; break;
        case 32: nc = 4u; break;
        default: fail(path, "unsupported TGA bit depth");
    }
    r.skip(id_length);
    // one colour as it is stored -> nc 8-bit channels (red first)
    auto read_colour = [&](uint8_t *out) {
        if (packed16) {
            const uint32_t v = r.u16();
            out[0] = static_cast<uint8_t>(((v >> 10u) & 31u) * 255u / 31u);
            out[1] = static_cast<uint8_t>(((v >> 5u) & 31u) * 255u / 31u);
            out[2] = static_cast<uint8_t>((v & 31u) * 255u / 31u);
            return;
        }
        for (uint32_t c = 0; c < nc; c++) out[c] = r.u8();
        if (nc >= 3u) std::swap(out[0], out[2]);
    };
    std::vector<uint8_t> palette;
    if (mapped) {
        if (map_length == 0u) fail(path, "bad TGA palette");
        r.skip(map_start);
        palette.resize(static_cast<size_t>(map_length) * nc);
        for (uint32_t i = 0; i < map_length; i++) read_colour(&palette[static_cast<size_t>(i) * nc]);
    } else if (map_type == 0u && map_length != 0u) {
        fail(path, "TGA file with a colour map but no colour-mapped picture");
    }
    auto read_pixel = [&](uint8_t *out) {
        if (!mapped) return read_colour(out);
        uint32_t index = pixel_bits == 8u ? r.u8() : r.u16();
        if (index >= map_length) index = 0u;
        std::memcpy(out, &palette[static_cast<size_t>(index) * nc], nc);
    };
    const size_t count = static_cast<size_t>(width) * height;
    std::vector<uint8_t> pixels(count * nc);
    uint8_t px[4] = {};
    for (size_t i = 0; i < count;) {
        size_t run = 1u;
        bool repeat = false;
        if (rle) {
            const uint32_t packet = r.u8();
            run = (packet & 127u) + 1u;
            repeat = (packet & 128u) != 0u;
        }
        if (repeat) read_pixel(px);
        for (size_t k = 0; k < run && i < count; k++, i++) {
            if (!repeat) read_pixel(px);
            std::memcpy(&pixels[i * nc], px, nc);
        }
    }
    const bool bottom_up = !((descriptor >> 5u) & 1u);
    std::vector<float> samples(count * nc);
    for (uint32_t y = 0; y < height; y++) {
        const uint8_t *row = &pixels[static_cast<size_t>(bottom_up ? height - 1u - y : y) * width * nc];
        for (size_t i = 0; i < static_cast<size_t>(width) * nc; i++) samples[static_cast<size_t>(y) * width * nc + i] = static_cast<float>(row[i]) / 255.f;
    }
    return finish(width, height, nc, samples);
}

// ---- JPEG: decoded by jpegload.cpp (stb_image's arithmetic), stored like every other 8-bit file: grey -> BYTE1, colour -> BYTE4 ----
LoadedImage load_jpeg(const std::filesystem::path &path, const std::vector<uint8_t> &d) {
    uint32_t w = 0, h = 0, nc = 0;
    std::vector<uint8_t> pixels;
    decode_jpeg(path, d, w, h, nc, pixels);
    std::vector<float> samples(pixels.size());
    for (size_t i = 0; i < pixels.size(); i++) samples[i] = static_cast<float>(pixels[i]) / 255.f;
    return finish(w, h, nc, samples);
}

}// namespace

LoadedImage load_image(const std::filesystem::path &path) {
    auto ext = path.extension().string();
    for (auto &c : ext) c = static_cast<char>(std::tolower(c));
    auto data = read_file(path);
    if (ext == ".png") return load_png(path, data);
    if (ext == ".ppm" || ext == ".pgm" || ext == ".pnm" || ext == ".pfm") return load_pnm(path, data);
    if (ext == ".hdr") return load_hdr(path, data);
    if (ext == ".exr") return load_exr(path, data);
    if (ext == ".bmp") return load_bmp(path, data);
    if (ext == ".tga") return load_tga(path, data);
    if (ext == ".jpg" || ext == ".jpeg") return load_jpeg(path, data);
    fail(path, "unsupported image format '" + ext + "' (supported: .png .jpg .jpeg .bmp .tga .ppm .pgm .pfm .hdr .exr)");
}

}// namespace lrh
