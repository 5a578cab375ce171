// rtcore_io.cpp — host-side formats either side of the path (include/rtcore_io.h): Wavefront .obj -> rt_geometry
// (the reference's next assignment, vulkan-raytracing-basic/README.md:225-226) and RGBA8 framebuffer -> PPM (the
// headless stand-in for the copy into the B8G8R8A8_SRGB swapchain, main.cpp:50,1371-1375). No GPU code.
#include <errno.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <string>
#include <vector>

#include "../../include/rtcore_io.h"

struct rt_obj_mesh {
    std::vector<float> vertices;            // x y z
    std::vector<uint32_t> indices;          // 3 per triangle, 0-based
    struct Group { std::string name; uint32_t first_tri, tri_count; };
    std::vector<Group> groups;
};

namespace {

thread_local std::string g_error;

int fail(int code, const char* fmt, long line, const char* what) {
    char buf[256];
    snprintf(buf, sizeof buf, fmt, line, what);
    g_error = buf;
    return code;
}

inline bool is_space(char c) { return c == ' ' || c == '\t' || c == '\r' || c == '\v' || c == '\f'; }

// one logical line = physical lines joined at a trailing backslash
struct LineReader {
    const char* p; const char* end; long line_no = 0;
    std::string buf;
    bool next() {
        if (p >= end) return false;
        buf.clear();
        for (;;) {
            const char* e = (const char*)memchr(p, '\n', (size_t)(end - p));
            const char* stop = e ? e : end;
            ++line_no;
            const char* q = stop;
            while (q > p && is_space(q[-1])) --q;                    // strips the CR of CRLF too
            const bool cont = q > p && q[-1] == '\\';
            buf.append(p, (size_t)((cont ? q - 1 : q) - p));
            p = e ? e + 1 : end;
            if (!cont || p >= end) break;
            buf.push_back(' ');
        }
        return true;
    }
};

// parses the leading integer of a face reference "v", "v/vt", "v//vn", "v/vt/vn"; advances s past the whole token
bool face_ref(const char*& s, long& v) {
    char* e;
    errno = 0;
    v = strtol(s, &e, 10);
    if (e == s || errno) return false;
    s = e;
    while (*s && !is_space(*s)) {                                    // "/vt/vn" tail: digits, '-', '/'
        if (!(*s == '/' || *s == '-' || *s == '+' || (*s >= '0' && *s <= '9'))) return false;
        ++s;
    }
    return true;
}

}  // namespace

extern "C" {

const char* rt_obj_last_error(void) { return g_error.c_str(); }

int rt_obj_parse(const char* text, size_t length, rt_obj_mesh** out) {
    if (!text || !out) return RT_ERROR_INVALID_ARG;
    *out = nullptr;
    g_error.clear();
    rt_obj_mesh* m = new rt_obj_mesh();
    auto open_group = [&](const std::string& name) {
        if (!m->groups.empty() && m->groups.back().tri_count == 0) m->groups.back().name = name;   // empty group: replace
        else m->groups.push_back({name, (uint32_t)(m->indices.size() / 3), 0u});
    };
    LineReader r{text, text + length};
    std::vector<long> poly;
    while (r.next()) {
        const char* s = r.buf.c_str();
        while (is_space(*s)) ++s;
        if (!*s || *s == '#') continue;
        const char* kw = s;
        while (*s && !is_space(*s)) ++s;
        const size_t kl = (size_t)(s - kw);
        if (kl == 1 && kw[0] == 'v') {
            double c[3];
            for (int k = 0; k < 3; ++k) {
                char* e;
                c[k] = strtod(s, &e);
                if (e == s) { delete m; return fail(RT_ERROR_PARSE, "line %ld: %s", r.line_no, "vertex needs three coordinates"); }
                s = e;
            }
            for (int k = 0; k < 3; ++k) m->vertices.push_back((float)c[k]);
        } else if (kl == 1 && kw[0] == 'f') {
            poly.clear();
            const long nv = (long)(m->vertices.size() / 3);
            for (;;) {
                while (is_space(*s)) ++s;
                if (!*s || *s == '#') break;
                long v;
                if (!face_ref(s, v)) { delete m; return fail(RT_ERROR_PARSE, "line %ld: %s", r.line_no, "bad face reference"); }
                const long idx = v > 0 ? v - 1 : nv + v;            // 1-based, or relative to the vertices read so far
                if (v == 0 || idx < 0 || idx >= nv) { delete m; return fail(RT_ERROR_PARSE, "line %ld: %s", r.line_no, "face index out of range"); }
                poly.push_back(idx);
            }
            if (poly.size() < 3) { delete m; return fail(RT_ERROR_PARSE, "line %ld: %s", r.line_no, "face needs at least three vertices"); }
            if (m->groups.empty()) open_group("");
            for (size_t i = 1; i + 1 < poly.size(); ++i) {
                m->indices.push_back((uint32_t)poly[0]); m->indices.push_back((uint32_t)poly[i]); m->indices.push_back((uint32_t)poly[i + 1]);
                ++m->groups.back().tri_count;
            }
        } else if (kl == 1 && (kw[0] == 'o' || kw[0] == 'g')) {
            while (is_space(*s)) ++s;
            open_group(std::string(s));
        }
        // everything else (vt, vn, vp, usemtl, mtllib, s, l, p, ...) is not part of the geometry the build consumes
        if (m->vertices.size() / 3 > 0xFFFFFFF0ull || m->indices.size() / 3 > 0xFFFFFFF0ull) {
            delete m; return fail(RT_ERROR_PARSE, "line %ld: %s", r.line_no, "mesh exceeds 32-bit counts");
        }
    }
    if (!m->groups.empty() && m->groups.back().tri_count == 0) m->groups.pop_back();
    *out = m;
    return RT_SUCCESS;
}

int rt_obj_load(const char* path, rt_obj_mesh** out) {
    if (!path || !out) return RT_ERROR_INVALID_ARG;
    *out = nullptr;
    FILE* f = fopen(path, "rb");
    if (!f) return fail(RT_ERROR_IO, "cannot open (errno %ld): %s", (long)errno, path);
    std::string text;
    char buf[1 << 16];
    size_t n;
    while ((n = fread(buf, 1, sizeof buf, f)) > 0) text.append(buf, n);
    const bool bad = ferror(f) != 0;
    fclose(f);
    if (bad) return fail(RT_ERROR_IO, "read error (errno %ld): %s", (long)errno, path);
    return rt_obj_parse(text.data(), text.size(), out);
}

void rt_obj_free(rt_obj_mesh* mesh) { delete mesh; }

uint32_t rt_obj_vertex_count(const rt_obj_mesh* m) { return m ? (uint32_t)(m->vertices.size() / 3) : 0u; }
uint32_t rt_obj_triangle_count(const rt_obj_mesh* m) { return m ? (uint32_t)(m->indices.size() / 3) : 0u; }
uint32_t rt_obj_group_count(const rt_obj_mesh* m) { return m ? (uint32_t)m->groups.size() : 0u; }
const float* rt_obj_vertices(const rt_obj_mesh* m) { return m && !m->vertices.empty() ? m->vertices.data() : nullptr; }
const uint32_t* rt_obj_indices(const rt_obj_mesh* m) { return m && !m->indices.empty() ? m->indices.data() : nullptr; }
const char* rt_obj_group_name(const rt_obj_mesh* m, uint32_t g) { return m && g < m->groups.size() ? m->groups[g].name.c_str() : ""; }
uint32_t rt_obj_group_first_triangle(const rt_obj_mesh* m, uint32_t g) { return m && g < m->groups.size() ? m->groups[g].first_tri : 0u; }
uint32_t rt_obj_group_triangle_count(const rt_obj_mesh* m, uint32_t g) { return m && g < m->groups.size() ? m->groups[g].tri_count : 0u; }

int rt_obj_geometry(const rt_obj_mesh* m, uint32_t g, rt_geometry* out) {
    if (!m || !out || g >= m->groups.size()) return RT_ERROR_INVALID_ARG;
    out->vertices = m->vertices.data();
    out->vertex_count = (uint32_t)(m->vertices.size() / 3);
    out->vertex_stride_bytes = 12;
    out->indices = m->indices.data() + 3 * (size_t)m->groups[g].first_tri;
    out->triangle_count = m->groups[g].tri_count;
    out->transform3x4 = nullptr;
    out->flags = RT_GEOMETRY_OPAQUE;
    return RT_SUCCESS;
}

// out[i] = round(255 * oetf(i/255)), oetf = the sRGB transfer function (IEC 61966-2-1), evaluated in double
void rt_srgb8_table(uint8_t out[256]) {
    for (int i = 0; i < 256; ++i) {
        const double l = i / 255.0;
        const double v = l <= 0.0031308 ? 12.92 * l : 1.055 * pow(l, 1.0 / 2.4) - 0.055;
        out[i] = (uint8_t)floor(255.0 * v + 0.5);
    }
}

int rt_write_ppm(const char* path, const uint8_t* rgba, uint32_t width, uint32_t height, uint32_t flags) {
    if (!path || !rgba || !width || !height) return RT_ERROR_INVALID_ARG;
    FILE* f = fopen(path, "wb");
    if (!f) return fail(RT_ERROR_IO, "cannot create (errno %ld): %s", (long)errno, path);
    uint8_t lut[256];
    if (flags & RT_IMAGE_SRGB_ENCODE) rt_srgb8_table(lut);
    else for (int i = 0; i < 256; ++i) lut[i] = (uint8_t)i;
    fprintf(f, "P6\n%u %u\n255\n", width, height);
    std::vector<uint8_t> row(3 * (size_t)width);
    for (uint32_t y = 0; y < height; ++y) {
        const uint8_t* src = rgba + 4 * (size_t)width * ((flags & RT_IMAGE_FLIP_Y) ? height - 1 - y : y);
        for (uint32_t x = 0; x < width; ++x) { row[3 * x] = lut[src[4 * x]]; row[3 * x + 1] = lut[src[4 * x + 1]]; row[3 * x + 2] = lut[src[4 * x + 2]]; }
        if (fwrite(row.data(), 1, row.size(), f) != row.size()) { fclose(f); return fail(RT_ERROR_IO, "write error (errno %ld): %s", (long)errno, path); }
    }
    if (fclose(f) != 0) return fail(RT_ERROR_IO, "write error (errno %ld): %s", (long)errno, path);
    return RT_SUCCESS;
}

namespace {
uint32_t crc32_update(uint32_t crc, const uint8_t* p, size_t n) {
    static uint32_t table[256];
    static bool init = false;
    if (!init) {
        for (uint32_t i = 0; i < 256; ++i) { uint32_t c = i; for (int k = 0; k < 8; ++k) c = (c & 1u) ? 0xEDB88320u ^ (c >> 1) : c >> 1; table[i] = c; }
        init = true;
    }
    for (size_t i = 0; i < n; ++i) crc = table[(crc ^ p[i]) & 255u] ^ (crc >> 8);
    return crc;
}
void put_be32(std::vector<uint8_t>& v, uint32_t x) { v.push_back((uint8_t)(x >> 24)); v.push_back((uint8_t)(x >> 16)); v.push_back((uint8_t)(x >> 8)); v.push_back((uint8_t)x); }
bool write_chunk(FILE* f, const char type[4], const std::vector<uint8_t>& data) {
    std::vector<uint8_t> head;
    put_be32(head, (uint32_t)data.size());
    head.insert(head.end(), type, type + 4);
    uint32_t crc = crc32_update(0xFFFFFFFFu, (const uint8_t*)type, 4);
    if (!data.empty()) crc = crc32_update(crc, data.data(), data.size());
    std::vector<uint8_t> tail;
    put_be32(tail, crc ^ 0xFFFFFFFFu);
    return fwrite(head.data(), 1, head.size(), f) == head.size() && (data.empty() || fwrite(data.data(), 1, data.size(), f) == data.size()) &&
           fwrite(tail.data(), 1, 4, f) == 4;
}
}  // namespace

int rt_write_png(const char* path, const uint8_t* rgba, uint32_t width, uint32_t height, uint32_t flags) {
    if (!path || !rgba || !width || !height) return RT_ERROR_INVALID_ARG;
    uint8_t lut[256];
    if (flags & RT_IMAGE_SRGB_ENCODE) rt_srgb8_table(lut);
    else for (int i = 0; i < 256; ++i) lut[i] = (uint8_t)i;
    // raw scanlines: filter byte 0 + RGB
    const size_t stride = 1 + 3 * (size_t)width;
    std::vector<uint8_t> raw(stride * height);
    for (uint32_t y = 0; y < height; ++y) {
        const uint8_t* src = rgba + 4 * (size_t)width * ((flags & RT_IMAGE_FLIP_Y) ? height - 1 - y : y);
        uint8_t* dst = raw.data() + stride * y;
        dst[0] = 0;
        for (uint32_t x = 0; x < width; ++x) { dst[1 + 3 * x] = lut[src[4 * x]]; dst[2 + 3 * x] = lut[src[4 * x + 1]]; dst[3 + 3 * x] = lut[src[4 * x + 2]]; }
    }
    // zlib stream: header, stored blocks of <= 65535 bytes, Adler-32
    std::vector<uint8_t> z;
    z.reserve(raw.size() + raw.size() / 65535 * 5 + 16);
    z.push_back(0x78); z.push_back(0x01);
    uint32_t a = 1, b = 0;
    for (size_t off = 0; off < raw.size(); off += 65535) {
        const size_t n = raw.size() - off < 65535 ? raw.size() - off : 65535;
        z.push_back(off + n == raw.size() ? 1 : 0);
        z.push_back((uint8_t)(n & 255)); z.push_back((uint8_t)(n >> 8)); z.push_back((uint8_t)(~n & 255)); z.push_back((uint8_t)((~n >> 8) & 255));
        z.insert(z.end(), raw.begin() + off, raw.begin() + off + n);
        for (size_t i = 0; i < n; ++i) { a = (a + raw[off + i]) % 65521u; b = (b + a) % 65521u; }
    }
    put_be32(z, (b << 16) | a);
    FILE* f = fopen(path, "wb");
    if (!f) return fail(RT_ERROR_IO, "cannot create (errno %ld): %s", (long)errno, path);
    static const uint8_t sig[8] = {0x89, 'P', 'N', 'G', 0x0D, 0x0A, 0x1A, 0x0A};
    std::vector<uint8_t> ihdr;
    put_be32(ihdr, width); put_be32(ihdr, height);
    ihdr.push_back(8); ihdr.push_back(2); ihdr.push_back(0); ihdr.push_back(0); ihdr.push_back(0);     // 8-bit, RGB, deflate, adaptive filtering, no interlace
    const bool ok = fwrite(sig, 1, 8, f) == 8 && write_chunk(f, "IHDR", ihdr) && write_chunk(f, "IDAT", z) && write_chunk(f, "IEND", {});
    if (fclose(f) != 0 || !ok) return fail(RT_ERROR_IO, "write error (errno %ld): %s", (long)errno, path);
    return RT_SUCCESS;
}

}  // extern "C"
