// JPEG reader for the `Image` texture plugin (ITU-T T.81: Huffman-coded baseline / extended-sequential / progressive DCT,
// 8-bit samples, one (grey), three (YCbCr or RGB) or four (CMYK / YCCK) components, restart intervals, interleaved and single-component scans).
//
// The reference reads .jpg / .jpeg through stb_image (LoadedImage::load, src/util/imageio.cpp:347-470 -> stbi_load).  Entropy
// decoding is exact by the standard; everything after it is lossy, so the texels only equal the reference's if the arithmetic
// behind the coefficients is stb_image's own.  This file restates that arithmetic (src/compute/src/ext/stb/stb/stb_image.h):
//   * dequantised coefficients wrap to 16 bits                                       (:2251, :3343-3346)
//   * the inverse DCT is the 12-bit fixed-point "islow" one with two extra bits kept between the passes, +128 and a clamp (:2425-2516)
//   * chroma planes are upsampled row by row: the 3:1 "tent" for a factor of two along x and / or y centred like JFIF, sample
//     repetition for every other factor, with stb's row bookkeeping at the top and bottom edge   (:3449-3527, :3645-3656, :3895-3939)
//   * YCbCr -> RGB in 20-bit fixed point with the green chroma term truncated to 16 bits        (:3658-3684)
//   * a three-component file is RGB as it stands when its component ids are 'R','G','B' or when an Adobe marker says
//     "no transform" and there is no JFIF marker                                                   (:3877)
// The pins: tests/golden/jpeg_texels.npz holds what the reference's own stb_image (compiled into oracle/_ref/bin/libluisa-ref.so)
// decodes for every fixture file (tools/gen_jpeg_pins.py); tests/test_textures_meshes.py compares this reader with it byte for byte.
// Four-component files (CMYK / YCCK by the Adobe marker) come out as RGB like stb's (:3858-3862, :3950-3975).
// Not read: arithmetic coding, lossless and hierarchical processes, 12-bit samples, DNL.
#include <array>
#include <cstdint>
#include <cstring>
#include <filesystem>
#include <stdexcept>
#include <string>
#include <vector>

namespace lrh {

namespace {

[[noreturn]] void jfail(const std::filesystem::path &p, const std::string &why) {
    throw std::runtime_error("Failed to load image '" + p.string() + "': " + why + ".");
}

constexpr uint8_t kZigzag[64] = {0, 1, 8, 16, 9, 2, 3, 10, 17, 24, 32, 25, 18, 11, 4, 5, 12, 19, 26, 33, 40, 48,
                                 41, 34, 27, 20, 13, 6, 7, 14, 21, 28, 35, 42, 49, 56, 57, 50, 43, 36, 29, 22,
                                 15, 23, 30, 37, 44, 51, 58, 59, 52, 45, 38, 31, 39, 46, 53, 60, 61, 54, 47, 55, 62, 63};

// canonical Huffman code (T.81 annex C / F.2.2.3): codes of each length are consecutive, decoded one length at a time
struct HuffmanTable {
    bool defined{false};
    uint8_t values[256]{};
    int32_t max_code[17]{}, first_code[17]{}, first_index[17]{};
    void build(const uint8_t counts[16], const uint8_t *symbols) {
        int32_t code = 0, index = 0;
        for (int len = 1; len <= 16; len++) {
            first_code[len] = code;
            first_index[len] = index;
            code += counts[len - 1];
            index += counts[len - 1];
            max_code[len] = counts[len - 1] ? code - 1 : -1;
            code <<= 1;
        }
        std::memcpy(values, symbols, static_cast<size_t>(index));
        defined = true;
    }
};

struct Component {
    uint32_t id{}, h{}, v{}, tq{}, dc_table{}, ac_table{};
    uint32_t width{}, height{};          // samples this component really has: ceil(image * h / h_max)
    uint32_t blocks_x{}, blocks_y{};     // blocks stored (whole MCUs)
    std::vector<int16_t> coefficients;   // blocks_x * blocks_y * 64, natural (row-major) order inside a block
    std::vector<uint8_t> plane;          // (blocks_x * 8) x (blocks_y * 8) samples after the inverse DCT
    int32_t dc_prediction{};
};

inline uint8_t clamp8(int32_t x) { return static_cast<uint8_t>(x < 0 ? 0 : x > 255 ? 255 : x); }

// stb_image.h:2425-2516 (itself after libjpeg's jidctint): 12-bit constants, the column pass keeps two extra bits
inline int32_t fx(double x) { return static_cast<int32_t>(x * 4096 + 0.5); }

struct Idct1D {
    int32_t x0, x1, x2, x3, t0, t1, t2, t3;
    Idct1D(int32_t s0, int32_t s1, int32_t s2, int32_t s3, int32_t s4, int32_t s5, int32_t s6, int32_t s7) {
        int32_t p2 = s2, p3 = s6;
        int32_t p1 = (p2 + p3) * fx(0.5411961f);
        int32_t e2 = p1 + p3 * fx(-1.847759065f);
        int32_t e3 = p1 + p2 * fx(0.765366865f);
        int32_t e0 = (s0 + s4) * 4096, e1 = (s0 - s4) * 4096;
        x0 = e0 + e3;
        x3 = e0 - e3;
        x1 = e1 + e2;
        x2 = e1 - e2;
        t0 = s7;
        t1 = s5;
        t2 = s3;
        t3 = s1;
        p3 = t0 + t2;
        int32_t p4 = t1 + t3;
        p1 = t0 + t3;
        p2 = t1 + t2;
        int32_t p5 = (p3 + p4) * fx(1.175875602f);
        t0 = t0 * fx(0.298631336f);
        t1 = t1 * fx(2.053119869f);
        t2 = t2 * fx(3.072711026f);
        t3 = t3 * fx(1.501321110f);
        p1 = p5 + p1 * fx(-0.899976223f);
        p2 = p5 + p2 * fx(-2.562915447f);
        p3 = p3 * fx(-1.961570560f);
        p4 = p4 * fx(-0.390180644f);
        t3 += p1 + p4;
        t2 += p2 + p3;
        t1 += p2 + p4;
        t0 += p1 + p3;
    }
};

void inverse_dct(uint8_t *out, size_t stride, const int16_t d[64]) {
    int32_t v[64];
    for (int i = 0; i < 8; i++) {
        Idct1D c{d[i], d[i + 8], d[i + 16], d[i + 24], d[i + 32], d[i + 40], d[i + 48], d[i + 56]};
        const int32_t x0 = c.x0 + 512, x1 = c.x1 + 512, x2 = c.x2 + 512, x3 = c.x3 + 512;
        v[i] = (x0 + c.t3) >> 10;
        v[i + 56] = (x0 - c.t3) >> 10;
        v[i + 8] = (x1 + c.t2) >> 10;
        v[i + 48] = (x1 - c.t2) >> 10;
        v[i + 16] = (x2 + c.t1) >> 10;
        v[i + 40] = (x2 - c.t1) >> 10;
        v[i + 24] = (x3 + c.t0) >> 10;
        v[i + 32] = (x3 - c.t0) >> 10;
    }
    for (int i = 0; i < 8; i++) {
        const int32_t *r = v + i * 8;
        uint8_t *o = out + static_cast<size_t>(i) * stride;
        Idct1D c{r[0], r[1], r[2], r[3], r[4], r[5], r[6], r[7]};
        constexpr int32_t bias = 65536 + (128 << 17);// rounding of the 17 fraction bits, and the level shift
        const int32_t x0 = c.x0 + bias, x1 = c.x1 + bias, x2 = c.x2 + bias, x3 = c.x3 + bias;
        o[0] = clamp8((x0 + c.t3) >> 17);
        o[7] = clamp8((x0 - c.t3) >> 17);
        o[1] = clamp8((x1 + c.t2) >> 17);
        o[6] = clamp8((x1 - c.t2) >> 17);
        o[2] = clamp8((x2 + c.t1) >> 17);
        o[5] = clamp8((x2 - c.t1) >> 17);
        o[3] = clamp8((x3 + c.t0) >> 17);
        o[4] = clamp8((x3 - c.t0) >> 17);
    }
}

class JpegReader {
    const std::filesystem::path &path;
    const std::vector<uint8_t> &d;
    size_t pos{0};
    // entropy-coded segment reader
    uint32_t bit_buffer{0};
    int bit_count{0};
    bool at_marker{false};

    std::array<std::array<uint16_t, 64>, 4> quant{};// natural order
    std::array<bool, 4> quant_defined{};
    std::array<HuffmanTable, 4> dc_tables{}, ac_tables{};
    std::vector<Component> components;
    uint32_t image_w{0}, image_h{0}, h_max{1}, v_max{1}, mcus_x{0}, mcus_y{0};
    bool progressive{false}, frame_seen{false}, jfif{false};
    int adobe_transform{-1};
    uint32_t restart_interval{0};
    uint32_t eob_run{0};

    [[noreturn]] void fail(const std::string &why) const { jfail(path, why); }
    uint8_t byte() {
        if (pos >= d.size()) fail("truncated JPEG file");
        return d[pos++];
    }
    uint32_t be16() {
        uint32_t hi = byte();
        return (hi << 8u) | byte();
    }

    // ---- bits of an entropy-coded segment: a 0xff data byte is followed by a stuffed zero; any other 0xff xx is a marker, which
    // ends the segment (the decoder then sees zero bits, T.81 F.2.2.5) ----
    int bit() {
        if (bit_count == 0) {
            uint8_t b = 0;
            if (!at_marker && pos < d.size()) {
                b = d[pos];
                if (b == 0xffu) {
                    size_t q = pos + 1u;
                    while (q < d.size() && d[q] == 0xffu) q++;// fill bytes
                    if (q < d.size() && d[q] == 0u) {
                        pos = q + 1u;
                    } else {
                        at_marker = true;
                        b = 0;
                    }
                } else {
                    pos++;
                }
            }
            bit_buffer = b;
            bit_count = 8;
        }
        bit_count--;
        return static_cast<int>((bit_buffer >> bit_count) & 1u);
    }
    int32_t bits(int n) {
        int32_t v = 0;
        for (int i = 0; i < n; i++) v = (v << 1) | bit();
        return v;
    }
    // a value of the size category s (T.81 F.2.2.1): the first half of the codes are the negative numbers
    int32_t receive_extend(int s) {
        if (s == 0) return 0;
        int32_t v = bits(s);
        return v < (1 << (s - 1)) ? v - (1 << s) + 1 : v;
    }
    uint8_t symbol(const HuffmanTable &t) {
        int32_t code = 0;
        for (int len = 1; len <= 16; len++) {
            code = (code << 1) | bit();
            if (t.max_code[len] >= 0 && code <= t.max_code[len] && code >= t.first_code[len]) return t.values[t.first_index[len] + code - t.first_code[len]];
        }
        fail("bad Huffman code in JPEG data");
    }
    void reset_segment() {
        bit_count = 0;
        bit_buffer = 0;
        at_marker = false;
        eob_run = 0;
        for (auto &c : components) c.dc_prediction = 0;
    }
    // the marker that follows the current position (fill bytes skipped); `pos` moves past it
    uint8_t next_marker() {
        while (pos < d.size() && d[pos] != 0xffu) pos++;// (stb skips junk before a marker too)
        while (pos < d.size() && d[pos] == 0xffu) pos++;
        if (pos >= d.size()) fail("truncated JPEG file");
        return d[pos++];
    }

    // ---- marker segments ----
    void read_quantisation_tables() {
        int32_t left = static_cast<int32_t>(be16()) - 2;
        while (left > 0) {
            uint8_t pq_tq = byte();
            uint32_t precision = pq_tq >> 4u, index = pq_tq & 15u;
            if (precision > 1u || index > 3u) fail("bad JPEG quantisation table");
            for (int i = 0; i < 64; i++) quant[index][kZigzag[i]] = static_cast<uint16_t>(precision ? be16() : byte());
            quant_defined[index] = true;
            left -= precision ? 129 : 65;
        }
        if (left != 0) fail("bad JPEG quantisation table length");
    }
    void read_huffman_tables() {
        int32_t left = static_cast<int32_t>(be16()) - 2;
        while (left > 0) {
            uint8_t tc_th = byte();
            uint32_t cls = tc_th >> 4u, index = tc_th & 15u;
            if (cls > 1u || index > 3u) fail("bad JPEG Huffman table");
            uint8_t counts[16], symbols[256];
            uint32_t total = 0;
            for (auto &c : counts) total += (c = byte());
            if (total > 256u) fail("bad JPEG Huffman table");
            for (uint32_t i = 0; i < total; i++) symbols[i] = byte();
            (cls ? ac_tables : dc_tables)[index].build(counts, symbols);
            left -= static_cast<int32_t>(17u + total);
        }
        if (left != 0) fail("bad JPEG Huffman table length");
    }
    void read_frame_header(uint8_t marker) {
        if (frame_seen) fail("JPEG file with more than one frame");
        frame_seen = true;
        progressive = marker == 0xc2u;
        uint32_t length = be16();
        if (byte() != 8u) fail("only 8-bit JPEG files are supported");
        image_h = be16();
        image_w = be16();
        uint32_t n = byte();
        if (image_w == 0u || image_h == 0u) fail("JPEG file without a size in its frame header");
        if (static_cast<uint64_t>(image_w) * image_h > (1ull << 28u)) fail("JPEG picture larger than 2^28 pixels");// a header must not buy 25 GB
        if (n != 1u && n != 3u && n != 4u) fail("bad JPEG component count");
        if (length != 8u + 3u * n) fail("bad JPEG frame header length");
        components.resize(n);
        for (auto &c : components) {
            c.id = byte();
            uint8_t hv = byte();
            c.h = hv >> 4u;
            c.v = hv & 15u;
            c.tq = byte();
            if (c.h == 0u || c.h > 4u || c.v == 0u || c.v > 4u || c.tq > 3u) fail("bad JPEG component header");
            h_max = std::max(h_max, c.h);
            v_max = std::max(v_max, c.v);
        }
        for (auto &c : components)
            if (h_max % c.h != 0u || v_max % c.v != 0u) fail("JPEG sampling factors that do not divide the largest one are not supported");
        mcus_x = (image_w + 8u * h_max - 1u) / (8u * h_max);
        mcus_y = (image_h + 8u * v_max - 1u) / (8u * v_max);
        for (auto &c : components) {
            c.width = (image_w * c.h + h_max - 1u) / h_max;
            c.height = (image_h * c.v + v_max - 1u) / v_max;
            c.blocks_x = mcus_x * c.h;
            c.blocks_y = mcus_y * c.v;
            c.coefficients.assign(static_cast<size_t>(c.blocks_x) * c.blocks_y * 64u, 0);
        }
    }

    // ---- one block of a scan ----
    void decode_sequential_block(Component &c, int16_t *block) {
        const auto &dc = dc_tables[c.dc_table], &ac = ac_tables[c.ac_table];
        c.dc_prediction += receive_extend(symbol(dc));
        block[0] = static_cast<int16_t>(c.dc_prediction);
        for (int k = 1; k < 64;) {
            uint8_t rs = symbol(ac);
            int r = rs >> 4, s = rs & 15;
            if (s == 0) {
                if (r != 15) break;// end of block
                k += 16;
            } else {
                k += r;
                if (k > 63) fail("bad JPEG coefficient run");
                block[kZigzag[k++]] = static_cast<int16_t>(receive_extend(s));
            }
        }
    }
    void decode_progressive_dc(Component &c, int16_t *block, int ah, int al) {
        if (ah == 0) {
            c.dc_prediction += receive_extend(symbol(dc_tables[c.dc_table]));
            block[0] = static_cast<int16_t>(c.dc_prediction * (1 << al));
        } else if (bit()) {
            block[0] = static_cast<int16_t>(block[0] + (1 << al));
        }
    }
    void decode_progressive_ac_first(Component &c, int16_t *block, int ss, int se, int al) {
        if (eob_run > 0u) {
            eob_run--;
            return;
        }
        const auto &ac = ac_tables[c.ac_table];
        for (int k = ss; k <= se;) {
            uint8_t rs = symbol(ac);
            int r = rs >> 4, s = rs & 15;
            if (s == 0) {
                if (r < 15) {// end of band for this and the next eob_run blocks
                    eob_run = (1u << r) - 1u;
                    if (r) eob_run += static_cast<uint32_t>(bits(r));
                    break;
                }
                k += 16;
            } else {
                k += r;
                if (k > se) fail("bad JPEG coefficient run");
                block[kZigzag[k++]] = static_cast<int16_t>(receive_extend(s) * (1 << al));
            }
        }
    }
    // T.81 G.1.2.3: one more bit for every coefficient that is already nonzero, new +-1 coefficients in the zero gaps between them
    void refine(int16_t &coefficient, int32_t plus) {
        if (bit() && (coefficient & plus) == 0) coefficient = static_cast<int16_t>(coefficient + (coefficient > 0 ? plus : -plus));
    }
    void decode_progressive_ac_refine(Component &c, int16_t *block, int ss, int se, int al) {
        const int32_t plus = 1 << al;
        const auto &ac = ac_tables[c.ac_table];
        int k = ss;
        if (eob_run == 0u) {
            while (k <= se) {
                uint8_t rs = symbol(ac);
                int r = rs >> 4, s = rs & 15;
                int32_t value = 0;
                if (s == 0) {
                    if (r < 15) {
                        eob_run = 1u << r;
                        if (r) eob_run += static_cast<uint32_t>(bits(r));
                        break;
                    }
                } else {
                    if (s != 1) fail("bad JPEG refinement scan");
                    value = bit() ? plus : -plus;
                }
                // pass r zero-valued coefficients (refining the nonzero ones on the way); the next zero one takes the new value
                while (k <= se) {
                    int16_t &coefficient = block[kZigzag[k++]];
                    if (coefficient != 0) {
                        refine(coefficient, plus);
                    } else {
                        if (r == 0) {
                            coefficient = static_cast<int16_t>(value);
                            break;
                        }
                        r--;
                    }
                }
            }
        }
        if (eob_run > 0u) {
            for (; k <= se; k++)
                if (int16_t &coefficient = block[kZigzag[k]]; coefficient != 0) refine(coefficient, plus);
            eob_run--;
        }
    }

    void restart() {
        // the segment ends on a byte boundary with RSTn
        bit_count = 0;
        uint8_t m = next_marker();
        if (m < 0xd0u || m > 0xd7u) fail("missing JPEG restart marker");
        reset_segment();
    }

    void read_scan() {
        if (!frame_seen) fail("JPEG scan before the frame header");
        uint32_t length = be16();
        uint32_t n = byte();
        if (n < 1u || n > components.size() || length != 6u + 2u * n) fail("bad JPEG scan header");
        std::vector<Component *> scan;
        for (uint32_t i = 0; i < n; i++) {
            uint32_t id = byte(), tables = byte();
            Component *found = nullptr;
            for (auto &c : components)
                if (c.id == id) found = &c;
            if (found == nullptr) fail("JPEG scan of an unknown component");
            found->dc_table = tables >> 4u;
            found->ac_table = tables & 15u;
            if (found->dc_table > 3u || found->ac_table > 3u) fail("bad JPEG scan header");
            scan.push_back(found);
        }
        int ss = byte(), se = byte();
        uint8_t a = byte();
        int ah = a >> 4, al = a & 15;
        if (progressive) {
            if (ss > 63 || se > 63 || ss > se || ah > 13 || al > 13 || (ss == 0 && se != 0) || (ss != 0 && n != 1u)) fail("bad JPEG progressive scan header");
        } else {
            if (ss != 0 || ah != 0 || al != 0) fail("bad JPEG scan header");
        }
        for (auto *c : scan) {
            bool need_dc = !progressive || ss == 0, need_ac = !progressive || ss != 0;
            if (need_dc && !(progressive && ah != 0) && !dc_tables[c->dc_table].defined) fail("JPEG scan uses an undefined Huffman table");
            if (need_ac && !ac_tables[c->ac_table].defined) fail("JPEG scan uses an undefined Huffman table");
        }
        reset_segment();
        auto decode = [&](Component &c, uint32_t bx, uint32_t by) {
            int16_t *block = c.coefficients.data() + (static_cast<size_t>(by) * c.blocks_x + bx) * 64u;
            if (!progressive) decode_sequential_block(c, block);
            else if (ss == 0) decode_progressive_dc(c, block, ah, al);
            else if (ah == 0) decode_progressive_ac_first(c, block, ss, se, al);
            else decode_progressive_ac_refine(c, block, ss, se, al);
        };
        uint32_t until_restart = restart_interval;
        auto unit_done = [&](bool last) {
            if (restart_interval != 0u && --until_restart == 0u && !last) {
                restart();
                until_restart = restart_interval;
            }
        };
        if (n == 1u) {
            // a scan of one component covers only the blocks that hold picture samples (T.81 A.2.2)
            Component &c = *scan[0];
            uint32_t bw = (c.width + 7u) / 8u, bh = (c.height + 7u) / 8u;
            for (uint32_t by = 0; by < bh; by++)
                for (uint32_t bx = 0; bx < bw; bx++) {
                    decode(c, bx, by);
                    unit_done(by + 1u == bh && bx + 1u == bw);
                }
        } else {
            for (uint32_t my = 0; my < mcus_y; my++)
                for (uint32_t mx = 0; mx < mcus_x; mx++) {
                    for (auto *c : scan)
                        for (uint32_t y = 0; y < c->v; y++)
                            for (uint32_t x = 0; x < c->h; x++) decode(*c, mx * c->h + x, my * c->v + y);
                    unit_done(my + 1u == mcus_y && mx + 1u == mcus_x);
                }
        }
        bit_count = 0;// the next marker starts on a byte boundary
    }

    void reconstruct_planes() {
        for (auto &c : components) {
            if (!quant_defined[c.tq]) fail("JPEG component uses an undefined quantisation table");
            const auto &q = quant[c.tq];
            const size_t stride = static_cast<size_t>(c.blocks_x) * 8u;
            c.plane.assign(stride * c.blocks_y * 8u, 0);
            int16_t block[64];
            for (uint32_t by = 0; by < c.blocks_y; by++)
                for (uint32_t bx = 0; bx < c.blocks_x; bx++) {
                    const int16_t *src = c.coefficients.data() + (static_cast<size_t>(by) * c.blocks_x + bx) * 64u;
                    for (int i = 0; i < 64; i++) block[i] = static_cast<int16_t>(src[i] * static_cast<int32_t>(q[i]));// wraps like stb's short
                    inverse_dct(c.plane.data() + static_cast<size_t>(by) * 8u * stride + static_cast<size_t>(bx) * 8u, stride, block);
                }
        }
    }

    // ---- upsampling (stb_image.h:3449-3527, 3645-3656; the row bookkeeping of load_jpeg_image :3895-3939) ----
    static void upsample_row(uint8_t *out, const uint8_t *near, const uint8_t *far, uint32_t w, uint32_t hs, uint32_t vs) {
        if (hs == 1u && vs == 2u) {
            for (uint32_t i = 0; i < w; i++) out[i] = static_cast<uint8_t>((3 * near[i] + far[i] + 2) >> 2);
        } else if (hs == 2u && vs == 1u) {
            if (w == 1u) {
                out[0] = out[1] = near[0];
                return;
            }
            out[0] = near[0];
            out[1] = static_cast<uint8_t>((near[0] * 3 + near[1] + 2) >> 2);
            uint32_t i = 1;
            for (; i + 1u < w; i++) {
                int n = 3 * near[i] + 2;
                out[i * 2u] = static_cast<uint8_t>((n + near[i - 1u]) >> 2);
                out[i * 2u + 1u] = static_cast<uint8_t>((n + near[i + 1u]) >> 2);
            }
            out[i * 2u] = static_cast<uint8_t>((near[w - 2u] * 3 + near[w - 1u] + 2) >> 2);
            out[i * 2u + 1u] = near[w - 1u];
        } else if (hs == 2u && vs == 2u) {
            if (w == 1u) {
                out[0] = out[1] = static_cast<uint8_t>((3 * near[0] + far[0] + 2) >> 2);
                return;
            }
            int t1 = 3 * near[0] + far[0];
            out[0] = static_cast<uint8_t>((t1 + 2) >> 2);
            for (uint32_t i = 1; i < w; i++) {
                int t0 = t1;
                t1 = 3 * near[i] + far[i];
                out[i * 2u - 1u] = static_cast<uint8_t>((3 * t0 + t1 + 8) >> 4);
                out[i * 2u] = static_cast<uint8_t>((3 * t1 + t0 + 8) >> 4);
            }
            out[w * 2u - 1u] = static_cast<uint8_t>((t1 + 2) >> 2);
        } else {
            for (uint32_t i = 0; i < w; i++)
                for (uint32_t j = 0; j < hs; j++) out[i * hs + j] = near[i];
        }
    }

public:
    JpegReader(const std::filesystem::path &p, const std::vector<uint8_t> &data) : path{p}, d{data} {}

    void decode(uint32_t &w, uint32_t &h, uint32_t &nc, std::vector<uint8_t> &pixels) {
        if (d.size() < 4u || d[0] != 0xffu || d[1] != 0xd8u) fail("not a JPEG file");
        pos = 2u;
        for (bool done = false; !done;) {
            uint8_t m = next_marker();
            switch (m) {
                case 0xd9u: done = true; break;
                case 0xdbu: read_quantisation_tables(); break;
                case 0xc4u: read_huffman_tables(); break;
                case 0xc0u: case 0xc1u: case 0xc2u: read_frame_header(m); break;
                case 0xdau: read_scan(); break;
                case 0xddu:
                    if (be16() != 4u) fail("bad JPEG restart interval");
                    restart_interval = be16();
                    break;
                case 0xc3u: case 0xc5u: case 0xc6u: case 0xc7u: case 0xc9u: case 0xcau: case 0xcbu: case 0xcdu: case 0xceu: case 0xcfu: case 0xccu:
                    fail("unsupported JPEG process (lossless, hierarchical or arithmetic-coded)");
                case 0xdcu: fail("JPEG files that define their height after the first scan (DNL) are not supported");
                default: {
                    if (m == 0x01u || (m >= 0xd0u && m <= 0xd7u)) break;// parameterless
                    uint32_t length = be16();
                    if (length < 2u || pos + (length - 2u) > d.size()) fail("truncated JPEG file");
                    const uint8_t *body = d.data() + pos;
                    if (m == 0xe0u && length >= 7u && std::memcmp(body, "JFIF\0", 5u) == 0) jfif = true;
                    if (m == 0xeeu && length >= 14u && std::memcmp(body, "Adobe\0", 6u) == 0) adobe_transform = body[11];
                    pos += length - 2u;
                }
            }
        }
        if (!frame_seen) fail("JPEG file without a frame");
        reconstruct_planes();

        const uint32_t n = static_cast<uint32_t>(components.size());
        const bool as_rgb = n == 3u && ((components[0].id == 'R' && components[1].id == 'G' && components[2].id == 'B') || (adobe_transform == 0 && !jfif));
        w = image_w;
        h = image_h;
        nc = n == 1u ? 1u : 3u;// four components (CMYK / YCCK, told apart by the Adobe marker) come out as RGB (:3950-3975)
        pixels.assign(static_cast<size_t>(w) * h * nc, 0);
        // stb_image.h:3858-3862: x * y / 255, rounded
        auto scale8 = [](uint32_t x, uint32_t y) { uint32_t t = x * y + 128u; return static_cast<uint8_t>((t + (t >> 8u)) >> 8u); };
        struct Rows {
            uint32_t hs, vs, step, row, width;
            const uint8_t *line0, *line1;
            std::vector<uint8_t> buffer;
        };
        std::vector<Rows> rows(n);
        for (uint32_t k = 0; k < n; k++) {
            auto &c = components[k];
            auto &r = rows[k];
            r.hs = h_max / c.h;
            r.vs = v_max / c.v;
            r.step = r.vs >> 1u;
            r.row = 0u;
            r.width = (w + r.hs - 1u) / r.hs;
            r.line0 = r.line1 = c.plane.data();
            r.buffer.assign(static_cast<size_t>(w) + 2u * r.hs + 3u, 0);
        }
        std::vector<const uint8_t *> line(n);
        for (uint32_t y = 0; y < h; y++) {
            for (uint32_t k = 0; k < n; k++) {
                auto &c = components[k];
                auto &r = rows[k];
                const bool below = r.step >= (r.vs >> 1u);// this output row lies in the lower half of its source row
                const uint8_t *near = below ? r.line1 : r.line0, *far = below ? r.line0 : r.line1;
                if (r.hs == 1u && r.vs == 1u) {
                    line[k] = near;
                } else {
                    upsample_row(r.buffer.data(), near, far, r.width, r.hs, r.vs);
                    line[k] = r.buffer.data();
                }
                if (++r.step >= r.vs) {
                    r.step = 0u;
                    r.line0 = r.line1;
                    if (++r.row < c.height) r.line1 += static_cast<size_t>(c.blocks_x) * 8u;
                }
            }
            uint8_t *out = pixels.data() + static_cast<size_t>(y) * w * nc;
            if (n == 1u) {
                std::memcpy(out, line[0], w);
            } else if (n == 4u && adobe_transform == 0) {// CMYK, stored inverted: every colour scaled by the black channel
                for (uint32_t x = 0; x < w; x++, out += 3)
                    for (uint32_t k = 0; k < 3u; k++) out[k] = scale8(line[k][x], line[3][x]);
            } else if (as_rgb) {
                for (uint32_t x = 0; x < w; x++, out += 3)
                    for (uint32_t k = 0; k < 3u; k++) out[k] = line[k][x];
            } else {
                // stb_image.h:3658-3684: 20-bit fixed point, the constants rounded to 12 bits first
                auto fixed = [](float f) { return static_cast<int32_t>(f * 4096.0f + 0.5f) << 8; };
                const int32_t cr_r = fixed(1.40200f), cr_g = -fixed(0.71414f), cb_g = -fixed(0.34414f), cb_b = fixed(1.77200f);
                for (uint32_t x = 0; x < w; x++, out += 3) {
                    const int32_t luma = (static_cast<int32_t>(line[0][x]) << 20) + (1 << 19);
                    const int32_t cb = line[1][x] - 128, cr = line[2][x] - 128;
                    const int32_t red = luma + cr * cr_r;
                    const int32_t green = luma + cr * cr_g + static_cast<int32_t>(static_cast<uint32_t>(cb * cb_g) & 0xffff0000u);
                    const int32_t blue = luma + cb * cb_b;
                    out[0] = clamp8(red >> 20);
                    out[1] = clamp8(green >> 20);
                    out[2] = clamp8(blue >> 20);
                    if (n == 4u && adobe_transform == 2)// YCCK: the converted colours are the inverted CMY
                        for (uint32_t k = 0; k < 3u; k++) out[k] = scale8(255u - out[k], line[3][x]);
                }
            }
        }
    }
};

}// namespace

// nc = 1 (grey) or 3 (RGB); row 0 is the top row
void decode_jpeg(const std::filesystem::path &path, const std::vector<uint8_t> &data, uint32_t &w, uint32_t &h, uint32_t &nc,
                 std::vector<uint8_t> &pixels) {
    JpegReader{path, data}.decode(w, h, nc, pixels);
}

}// namespace lrh
