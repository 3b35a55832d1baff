// rtc_oracle.hpp — CPU ORACLE. TEST INFRASTRUCTURE ONLY: nothing in the product path may include, link
// or execute this file (only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
// reference legs do).
//
// A from-scratch C++17 restatement of the algorithm behind `Camera::render` in
// garfieldnate/ray_tracer_challenge (pure, single-threaded Rust, f32).  The Rust toolchain is absent
// from this image, so the reference itself cannot be run; parity is pinned instead on the reference's
// own golden vectors, which tests/test_oracle_golden_*.py replay against this
// file.  Every function cites the reference file:line (paths relative to lib/src/) it follows.
//
// Numeric contract: all arithmetic is IEEE f32 evaluated strictly left to right exactly as the Rust
// source spells it (Rust never contracts a*b+c into an FMA); build with
// `-O2 -ffp-contract=off -fno-fast-math`.  libm calls (powf, cosf, sinf, tanf, acosf, atan2f, fmodf) are
// glibc's, which is what Rust's std uses on Linux.
//
// What is NOT pinned: the `jitter_fn = None` path of RectangleLight (rectangle_light.rs:46) draws from
// an OS-seeded thread_rng and cannot be reproduced by anyone, including the reference.  The counter
// based generator below stands in for it: "parity unpinned" for that mode only.
#pragma once

#include <algorithm>
#include <cctype>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <limits>
#include <memory>
#include <stdexcept>
#include <string>
#include <unordered_map>
#include <vector>

namespace orc {

constexpr float kInf = std::numeric_limits<float>::infinity();
constexpr float kF32Epsilon = 1.1920929e-7f;  // f32::EPSILON

// Rust's f32::min / f32::max return the non-NaN operand (cube.rs:90-129 relies on it).
inline float rmin(float a, float b) { return fminf(a, b); }
inline float rmax(float a, float b) { return fmaxf(a, b); }

// Rust `as i32` on a float saturates and maps NaN to 0.
inline int32_t sat_i32(float f) {
    if (std::isnan(f)) return 0;
    if (f >= 2147483648.0f) return INT32_MAX;
    if (f <= -2147483648.0f) return INT32_MIN;
    return (int32_t)f;
}
inline size_t sat_usize(float f) {
    if (std::isnan(f) || f <= 0.0f) return 0;
    if (f >= 18446744073709551616.0f) return SIZE_MAX;
    return (size_t)f;
}
// Rust rounds half away from zero, like roundf.

// ---------------------------------------------------------------- tuple.rs
struct Tuple {
    float x, y, z, w;
};
inline Tuple point(float x, float y, float z) { return {x, y, z, 1.0f}; }    // tuple.rs:73-78
inline Tuple vector(float x, float y, float z) { return {x, y, z, 0.0f}; }   // tuple.rs:80-85
inline Tuple operator+(Tuple a, Tuple b) { return {a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w}; }  // :87-97
inline Tuple operator-(Tuple a, Tuple b) { return {a.x - b.x, a.y - b.y, a.z - b.z, a.w - b.w}; }  // :99-109
inline Tuple operator*(Tuple a, float s) { return {a.x * s, a.y * s, a.z * s, a.w * s}; }          // :111-122
inline Tuple operator/(Tuple a, float s) { return {a.x / s, a.y / s, a.z / s, a.w / s}; }          // :131-142
inline Tuple operator-(Tuple a) { return {-a.x, -a.y, -a.z, -a.w}; }                               // :144-155
// tuple.rs:29-33 — the w lane is part of the magnitude.
inline float magnitude(Tuple a) { return sqrtf(a.x * a.x + a.y * a.y + a.z * a.z + a.w * a.w); }
// tuple.rs:34-43 — xyz divided by the 4-lane magnitude, w kept.
inline Tuple norm(Tuple a) {
    float m = magnitude(a);
    return {a.x / m, a.y / m, a.z / m, a.w};
}
// tuple.rs:44-46 — 4-lane dot.
inline float dot(Tuple a, Tuple b) { return a.x * b.x + a.y * b.y + a.z * b.z + (a.w * b.w); }
// tuple.rs:47-55
inline Tuple cross(Tuple a, Tuple b) {
    return {a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x, 0.0f};
}

// ---------------------------------------------------------------- color.rs
struct Color {
    float r, g, b;
};
inline Color operator+(Color a, Color b) { return {a.r + b.r, a.g + b.g, a.b + b.b}; }   // color.rs:33-39
inline Color operator-(Color a, Color b) { return {a.r - b.r, a.g - b.g, a.b - b.b}; }   // :41-47
inline Color operator*(Color a, float s) { return {a.r * s, a.g * s, a.b * s}; }         // :50-56
inline Color operator/(Color a, float s) { return {a.r / s, a.g / s, a.b / s}; }         // :61-67
inline Color operator*(Color a, Color b) { return {a.r * b.r, a.g * b.g, a.b * b.b}; }   // :70-76
inline Color black() { return {0, 0, 0}; }
inline Color white() { return {1, 1, 1}; }

// ---------------------------------------------------------------- matrix.rs
struct Matrix {
    int n = 4;
    float d[4][4] = {{1, 0, 0, 0}, {0, 1, 0, 0}, {0, 0, 1, 0}, {0, 0, 0, 1}};

    static Matrix zeros(int n) {
        Matrix m;
        m.n = n;
        for (auto& r : m.d)
            for (float& v : r) v = 0.0f;
        return m;
    }
    static Matrix from16(const float* p) {
        Matrix m;
        for (int r = 0; r < 4; r++)
            for (int c = 0; c < 4; c++) m.d[r][c] = p[r * 4 + c];
        return m;
    }
    void to16(float* p) const {
        for (int r = 0; r < 4; r++)
            for (int c = 0; c < 4; c++) p[r * 4 + c] = d[r][c];
    }
    // matrix.rs:134-143
    Matrix transpose() const {
        Matrix m = zeros(n);
        for (int r = 0; r < n; r++)
            for (int c = 0; c < n; c++) m.d[c][r] = d[r][c];
        return m;
    }
    // matrix.rs:162-180
    Matrix submatrix(int rr, int rc) const {
        Matrix m = zeros(n - 1);
        int nr = 0;
        for (int r = 0; r < n; r++) {
            if (r == rr) continue;
            int nc = 0;
            for (int c = 0; c < n; c++) {
                if (c == rc) continue;
                m.d[nr][nc] = d[r][c];
                nc++;
            }
            nr++;
        }
        return m;
    }
    // matrix.rs:145-159 — cofactor expansion along row 0, accumulated from 0.0 left to right.
    float determinant() const {
        if (n == 2) return d[0][0] * d[1][1] - d[0][1] * d[1][0];
        float det = 0.0f;
        for (int c = 0; c < n; c++) {
            float cf = cofactor(0, c);
            det += cf * d[0][c];
        }
        return det;
    }
    float minor(int r, int c) const { return submatrix(r, c).determinant(); }  // matrix.rs:193-195
    // matrix.rs:182-191
    float cofactor(int r, int c) const {
        float m = minor(r, c);
        return ((r + c) % 2 == 0) ? m : -m;
    }
    // matrix.rs:201-212
    Matrix inverse() const {
        float det = determinant();
        Matrix inv = zeros(n);
        for (int r = 0; r < n; r++)
            for (int c = 0; c < n; c++) inv.d[c][r] = cofactor(r, c) / det;
        return inv;
    }
};
// matrix.rs:73-84 — all four rows, left-to-right sums.
inline Tuple operator*(const Matrix& a, Tuple b) {
    float x = a.d[0][0] * b.x + a.d[0][1] * b.y + a.d[0][2] * b.z + a.d[0][3] * b.w;
    float y = a.d[1][0] * b.x + a.d[1][1] * b.y + a.d[1][2] * b.z + a.d[1][3] * b.w;
    float z = a.d[2][0] * b.x + a.d[2][1] * b.y + a.d[2][2] * b.z + a.d[2][3] * b.w;
    float w = a.d[3][0] * b.x + a.d[3][1] * b.y + a.d[3][2] * b.z + a.d[3][3] * b.w;
    return {x, y, z, w};
}
// matrix.rs:86-103
inline Matrix operator*(const Matrix& a, const Matrix& b) {
    Matrix m = Matrix::zeros(4);
    for (int r = 0; r < 4; r++)
        for (int c = 0; c < 4; c++)
            m.d[r][c] = a.d[r][0] * b.d[0][c] + a.d[r][1] * b.d[1][c] + a.d[r][2] * b.d[2][c] + a.d[r][3] * b.d[3][c];
    return m;
}

// ---------------------------------------------------------------- transformations.rs
inline Matrix translation(float x, float y, float z) {  // :4-6
    Matrix m;
    m.d[0][3] = x;
    m.d[1][3] = y;
    m.d[2][3] = z;
    return m;
}
inline Matrix scaling(float x, float y, float z) {  // :8-10
    Matrix m;
    m.d[0][0] = x;
    m.d[1][1] = y;
    m.d[2][2] = z;
    return m;
}
inline Matrix rotation_x(float rad) {  // :12-21
    float c = cosf(rad), s = sinf(rad);
    Matrix m;
    m.d[1][1] = c;
    m.d[1][2] = -s;
    m.d[2][1] = s;
    m.d[2][2] = c;
    return m;
}
inline Matrix rotation_y(float rad) {  // :23-32
    float c = cosf(rad), s = sinf(rad);
    Matrix m;
    m.d[0][0] = c;
    m.d[0][2] = s;
    m.d[2][0] = -s;
    m.d[2][2] = c;
    return m;
}
inline Matrix rotation_z(float rad) {  // :34-43
    float c = cosf(rad), s = sinf(rad);
    Matrix m;
    m.d[0][0] = c;
    m.d[0][1] = -s;
    m.d[1][0] = s;
    m.d[1][1] = c;
    return m;
}
inline Matrix shearing(float xy, float xz, float yx, float yz, float zx, float zy) {  // :46-53
    Matrix m;
    m.d[0][1] = xy;
    m.d[0][2] = xz;
    m.d[1][0] = yx;
    m.d[1][2] = yz;
    m.d[2][0] = zx;
    m.d[2][1] = zy;
    return m;
}
inline Matrix view_transform(Tuple from, Tuple to, Tuple approximate_up) {  // :57-68
    Tuple forward = norm(to - from);
    Tuple left = cross(forward, norm(approximate_up));
    Tuple true_up = cross(left, forward);
    Matrix o;
    o.d[0][0] = left.x, o.d[0][1] = left.y, o.d[0][2] = left.z;
    o.d[1][0] = true_up.x, o.d[1][1] = true_up.y, o.d[1][2] = true_up.z;
    o.d[2][0] = -forward.x, o.d[2][1] = -forward.y, o.d[2][2] = -forward.z;
    return o * translation(-from.x, -from.y, -from.z);
}

// ---------------------------------------------------------------- ray.rs
struct Ray {
    Tuple origin, direction, direction_inverses;
    Ray() = default;
    // ray.rs:13-22 — three divides on every construction.
    Ray(Tuple o, Tuple d) : origin(o), direction(d), direction_inverses(vector(1.f / d.x, 1.f / d.y, 1.f / d.z)) {}
    Tuple position(float t) const { return origin + direction * t; }                           // :23-25
    Ray transform(const Matrix& m) const { return Ray(m * origin, m * direction); }            // :26-31
};
// ray.rs:42-44
inline Tuple reflect(Tuple in, Tuple normal) { return -(normal * 2.0f * dot(in, normal) - in); }

// ---------------------------------------------------------------- bounding_box.rs
struct BoundingBox {
    Tuple min = point(kInf, kInf, kInf);      // :14-21
    Tuple max = point(-kInf, -kInf, -kInf);
    void add_point(Tuple p) {  // :42-50
        min.x = rmin(min.x, p.x);
        min.y = rmin(min.y, p.y);
        min.z = rmin(min.z, p.z);
        max.x = rmax(max.x, p.x);
        max.y = rmax(max.y, p.y);
        max.z = rmax(max.z, p.z);
    }
    void add_bounding_box(const BoundingBox& o) {  // :52-55
        add_point(o.min);
        add_point(o.max);
    }
    bool contains_point(Tuple p) const {  // :57-61
        return p.x >= min.x && p.x <= max.x && p.y >= min.y && p.y <= max.y && p.z >= min.z && p.z <= max.z;
    }
    bool contains_bounding_box(const BoundingBox& o) const {  // :63-65
        return contains_point(o.min) && contains_point(o.max);
    }
    BoundingBox transform(const Matrix& m) const {  // :67-84
        BoundingBox nb;
        Tuple c[8] = {min,
                      point(min.x, min.y, max.z),
                      point(min.x, max.y, min.z),
                      point(min.x, max.y, max.z),
                      point(max.x, min.y, min.z),
                      point(max.x, min.y, max.z),
                      point(max.x, max.y, min.z),
                      max};
        for (const Tuple& p : c) nb.add_point(m * p);
        return nb;
    }
    bool intersects(const Ray& r) const;  // :86-88
    void split(BoundingBox& left, BoundingBox& right) const {  // :90-125
        float dx = max.x - min.x, dy = max.y - min.y, dz = max.z - min.z;
        float greatest = rmax(rmax(dx, dy), dz);
        float x0 = min.x, y0 = min.y, z0 = min.z, x1 = max.x, y1 = max.y, z1 = max.z;
        if (greatest == dx) {
            x0 = x0 + dx / 2.f;
            x1 = x0;
        } else if (greatest == dy) {
            y0 = y0 + dy / 2.f;
            y1 = y0;
        } else {
            z0 = z0 + dz / 2.f;
            z1 = z0;
        }
        left.min = min;
        left.max = point(x1, y1, z1);
        right.min = point(x0, y0, z0);
        right.max = max;
    }
};

// cube.rs:90-129 — branchless slab test on the cached direction inverses; used by Cube and every bbox.
inline bool aabb_intersection(const Ray& r, Tuple mn, Tuple mx, float& tmin, float& tmax) {
    float min_x = (mn.x - r.origin.x) * r.direction_inverses.x;
    float max_x = (mx.x - r.origin.x) * r.direction_inverses.x;
    float lo = rmin(min_x, max_x);
    float hi = rmax(min_x, max_x);
    float min_y = (mn.y - r.origin.y) * r.direction_inverses.y;
    float max_y = (mx.y - r.origin.y) * r.direction_inverses.y;
    lo = rmax(lo, rmin(min_y, max_y));
    hi = rmin(hi, rmax(min_y, max_y));
    float min_z = (mn.z - r.origin.z) * r.direction_inverses.z;
    float max_z = (mx.z - r.origin.z) * r.direction_inverses.z;
    lo = rmax(lo, rmin(min_z, max_z));
    hi = rmin(hi, rmax(min_z, max_z));
    if (hi >= rmax(0.0f, lo)) {
        tmin = lo;
        tmax = hi;
        return true;
    }
    return false;
}
inline bool BoundingBox::intersects(const Ray& r) const {
    float a, b;
    return aabb_intersection(r, min, max, a, b);
}

// ---------------------------------------------------------------- canvas.rs
// Canvas (canvas.rs:5-43), to_ppm (canvas.rs:45-96) and canvas_from_ppm (canvas.rs:119-200), restated line by line:
// P3 text, lines of at most 70 columns, 8-bit values through the truncating scale_color.
struct Canvas {
    size_t width = 0, height = 0;
    std::vector<Color> data;  // [y][x], black (canvas.rs:19-25)
    Canvas() = default;
    Canvas(size_t w, size_t h) : width(w), height(h), data(w * h, Color{0, 0, 0}) {}
    void write_pixel(size_t x, size_t y, Color c) {  // canvas.rs:26-32 (`<=`: an index equal to the size panics there)
        if (x < width && y < height) data[y * width + x] = c;
    }
    Color pixel_at(size_t x, size_t y) const { return data.at(y * width + x); }  // canvas.rs:34-36 (out of range: panic)
    static uint8_t scale_color(float rgb) {  // canvas.rs:39-43: f32::min / max return the non-NaN operand; `as u8` saturates
        float v = rgb * 255.0f;
        v = (v != v) ? 255.0f : (v < 255.0f ? v : 255.0f);
        v = v > 0.0f ? v : 0.0f;
        return (uint8_t)v;
    }
    static void write_rgb_separator(std::string& line, std::string& ppm) {  // canvas.rs:47-55
        if (line.size() < 70 - 3) {
            line.push_back(' ');
        } else {
            ppm += line;
            ppm.push_back('\n');
            line.clear();
        }
    }
    std::string to_ppm() const {  // canvas.rs:58-96
        std::string ppm = "P3\n" + std::to_string(width) + " " + std::to_string(height) + "\n255\n";
        std::string line;
        for (size_t row = 0; row < height; row++) {
            line.clear();
            for (size_t i = 0; i < width; i++) {
                Color c = pixel_at(i, row);
                line += std::to_string((unsigned)scale_color(c.r));
                write_rgb_separator(line, ppm);
                line += std::to_string((unsigned)scale_color(c.g));
                write_rgb_separator(line, ppm);
                line += std::to_string((unsigned)scale_color(c.b));
                if (i != width - 1) write_rgb_separator(line, ppm);
            }
            if (!line.empty()) {
                ppm += line;
                ppm.push_back('\n');
            }
        }
        return ppm;
    }
};

struct PpmError : std::runtime_error {
    using std::runtime_error::runtime_error;
};
inline std::string trim_ws(const std::string& s) {
    size_t a = 0, b = s.size();
    while (a < b && std::isspace((unsigned char)s[a])) a++;
    while (b > a && std::isspace((unsigned char)s[b - 1])) b--;
    return s.substr(a, b - a);
}
inline std::vector<std::string> split_ws(const std::string& s) {
    std::vector<std::string> out;
    size_t i = 0;
    while (i < s.size()) {
        while (i < s.size() && std::isspace((unsigned char)s[i])) i++;
        size_t j = i;
        while (j < s.size() && !std::isspace((unsigned char)s[j])) j++;
        if (j > i) out.push_back(s.substr(i, j - i));
        i = j;
    }
    return out;
}
inline uint64_t parse_unsigned(const std::string& s, uint64_t max_value) {  // str::parse::<u32 / usize>: digits, optional '+'
    size_t i = 0;
    if (!s.empty() && s[0] == '+') i = 1;
    if (i >= s.size()) throw PpmError("ParseIntError: cannot parse integer from '" + s + "'");
    uint64_t v = 0;
    for (; i < s.size(); i++) {
        if (s[i] < '0' || s[i] > '9') throw PpmError("ParseIntError: invalid digit found in '" + s + "'");
        v = v * 10 + (uint64_t)(s[i] - '0');
        if (v > max_value) throw PpmError("ParseIntError: number too large in '" + s + "'");
    }
    return v;
}
inline Canvas canvas_from_ppm(const std::string& text) {  // canvas.rs:119-182
    // BufRead::lines + clean_line (canvas.rs:184-200): trimmed, comment and empty lines dropped
    std::vector<std::string> lines;
    size_t pos = 0;
    while (pos <= text.size()) {
        size_t nl = text.find('\n', pos);
        std::string line = text.substr(pos, nl == std::string::npos ? std::string::npos : nl - pos);
        pos = nl == std::string::npos ? text.size() + 1 : nl + 1;
        std::string t = trim_ws(line);
        if (t.empty() || t[0] == '#') continue;
        lines.push_back(t);
    }
    if (lines.size() < 3) throw PpmError("unexpected end of file in the PPM header");  // the reference unwrap()s: a panic
    if (lines[0] != "P3") throw PpmError("IncorrectFormat: Incorrect magic number at line 1: expected P3, found " + lines[0]);
    std::vector<std::string> dims = split_ws(lines[1]);
    if (dims.size() != 2) throw PpmError("MalformedDimensionHeader: Expected width and height at line 2; found " + lines[1]);
    size_t width = (size_t)parse_unsigned(dims[0], UINT64_MAX / 16), height = (size_t)parse_unsigned(dims[1], UINT64_MAX / 16);
    float scale = (float)(uint32_t)parse_unsigned(lines[2], 0xffffffffull);
    Canvas canvas(width, height);
    std::vector<uint32_t> raw;
    size_t head = 0, x = 0, y = 0;
    for (size_t li = 3; li < lines.size(); li++) {
        for (const std::string& tok : split_ws(lines[li])) raw.push_back((uint32_t)parse_unsigned(tok, 0xffffffffull));
        while (raw.size() - head >= 3) {
            float r = (float)raw[head] / scale, g = (float)raw[head + 1] / scale, b = (float)raw[head + 2] / scale;
            head += 3;
            if (y >= height && x < width) throw PpmError("more pixel data than width x height");  // data[y][x] panics there
            canvas.write_pixel(x, y, Color{r, g, b});
            x += 1;
            if (x >= width) {
                x = 0;
                y += 1;
            }
        }
    }
    return canvas;
}

// ---------------------------------------------------------------- pattern/*.rs
struct Shape;

struct UVPattern {  // uv.rs:14-16
    virtual ~UVPattern() = default;
    virtual Color color_at(float u, float v) const = 0;
};
struct UVCheckers : UVPattern {  // uv.rs:20-56
    float width, height;
    Color a, b;
    Color color_at(float u, float v) const override {
        int32_t u2 = sat_i32(floorf(u * width));
        int32_t v2 = sat_i32(floorf(v * height));
        // release-mode i32 add wraps; `%` keeps the dividend's sign
        int32_t s = (int32_t)((uint32_t)u2 + (uint32_t)v2);
        return (s % 2 == 0) ? a : b;
    }
};
struct AlignCheck : UVPattern {  // uv.rs:136-176
    Color main, ul, ur, bl, br;
    Color color_at(float u, float v) const override {
        if (v > 0.8f) {
            if (u < 0.2f) return ul;
            if (u > 0.8f) return ur;
        } else if (v < 0.2f) {
            if (u < 0.2f) return bl;
            if (u > 0.8f) return br;
        }
        return main;
    }
};

struct UVImage : UVPattern {  // uv.rs:346-377
    Canvas canvas;
    Color color_at(float u, float v) const override {
        v = 1.0f - v;  // flip v over so it matches the image layout, with y at the top
        float x = u * (float)(canvas.width - 1);
        float y = v * (float)(canvas.height - 1);
        // `x.round() as usize`: half away from zero, saturating (NaN and negatives give 0); an index beyond the canvas
        // panics in the reference (canvas.rs:35) — here it throws
        auto index = [](float f) -> size_t {
            float r = roundf(f);
            if (!(r > 0.0f)) return 0;
            if (r >= 1.8446744e19f) return SIZE_MAX;
            return (size_t)r;
        };
        return canvas.pixel_at(index(x), index(y));
    }
};

constexpr float kPi = 3.14159265358979323846f;         // std::f32::consts::PI
constexpr float kFrac1Pi = 0.318309886183790671538f;   // FRAC_1_PI
constexpr float kFrac12Pi = 1.0f / (2.0f * kPi);       // uv.rs:12

inline float rem_euclid(float a, float b) {  // Rust f32::rem_euclid
    float r = fmodf(a, b);
    return (r < 0.0f) ? r + fabsf(b) : r;
}
inline float u_from_azimuth(Tuple p) {  // uv.rs:117-132
    float theta = atan2f(p.x, p.z);
    float raw_u = theta * kFrac12Pi;
    return 1.f - (raw_u + 0.5f);
}
inline void map_spherical(Tuple p, float& u, float& v) {  // uv.rs:97-115
    u = u_from_azimuth(p);
    float radius = magnitude(vector(p.x, p.y, p.z));
    float phi = acosf(p.y / radius);
    v = 1.f - phi * kFrac1Pi;
}
inline void map_planar(Tuple p, float& u, float& v) {  // uv.rs:196-200
    u = rem_euclid(p.x, 1.f);
    v = rem_euclid(p.z, 1.f);
}
inline void map_cylindrical(Tuple p, float& u, float& v) {  // uv.rs:204-212
    u = u_from_azimuth(p);
    v = rem_euclid(p.y, 2.f * kPi) * kFrac12Pi;
}
enum Face { FRONT = 0, BACK = 1, LEFT = 2, RIGHT = 3, UP = 4, DOWN = 5 };  // uv.rs:186-193
inline Face face_from_point(Tuple p) {                                     // uv.rs:271-290
    float coord = rmax(rmax(fabsf(p.x), fabsf(p.y)), fabsf(p.z));
    if (coord == p.x) return RIGHT;
    if (coord == -p.x) return LEFT;
    if (coord == p.y) return UP;
    if (coord == -p.y) return DOWN;
    if (coord == p.z) return FRONT;
    return BACK;
}
// uv.rs:292-326 — Rust `%` on floats is fmod (sign of the dividend).
inline void cube_uv(Face f, Tuple p, float& u, float& v) {
    switch (f) {
        case FRONT: u = fmodf(p.x + 1.f, 2.f) / 2.f; v = fmodf(p.y + 1.f, 2.f) / 2.f; break;
        case BACK:  u = fmodf(1.f - p.x, 2.f) / 2.f; v = fmodf(p.y + 1.f, 2.f) / 2.f; break;
        case LEFT:  u = fmodf(p.z + 1.f, 2.f) / 2.f; v = fmodf(p.y + 1.f, 2.f) / 2.f; break;
        case RIGHT: u = fmodf(1.f - p.z, 2.f) / 2.f; v = fmodf(p.y + 1.f, 2.f) / 2.f; break;
        case UP:    u = fmodf(p.x + 1.f, 2.f) / 2.f; v = fmodf(1.f - p.z, 2.f) / 2.f; break;
        case DOWN:  u = fmodf(p.x + 1.f, 2.f) / 2.f; v = fmodf(p.z + 1.f, 2.f) / 2.f; break;
    }
}

struct Pattern {  // pattern.rs:8-26
    int kind = 0;
    Matrix t_inverse;  // BasePattern, pattern.rs:31-34
    Color a = white(), b = black();
    Color distance = black();  // gradient.rs:18, sine_2d.rs:17 — (b - a), precomputed at construction
    std::shared_ptr<UVPattern> uv;       // TextureMap
    int mapping = 0;
    std::shared_ptr<UVPattern> faces[6];  // CubicMap
    void set_transformation(const Matrix& t) { t_inverse = t.inverse(); }  // pattern.rs:53-55
    Color color_at_world(Tuple p) const;
    Color color_at_object(Tuple world_point, const Shape& object) const;  // pattern.rs:15-19
};

// ---------------------------------------------------------------- material.rs:19-51
struct Material {
    Color color = white();
    float ambient = 0.1f, diffuse = 0.9f, specular = 0.9f, shininess = 200.0f;
    float reflective = 0.0f, transparency = 0.0f, refractive_index = 1.0f;
    std::shared_ptr<Pattern> pattern;
};

// ---------------------------------------------------------------- intersection.rs
struct Intersection {  // :5-10
    float distance;
    const Shape* object;
    float u = 0.f, v = 0.f;
};
inline bool same_intersection(const Intersection& a, const Intersection& b);  // derive(PartialEq), :4

// ---------------------------------------------------------------- shape/*.rs
struct Counters {
    uint64_t primary = 0, secondary = 0, shadow = 0, shades = 0;
    uint64_t xforms = 0, aabb = 0, prim[7] = {0, 0, 0, 0, 0, 0, 0}, patterns = 0, cells = 0, schlick = 0,
             refr_dirs = 0, combines = 0;
};

extern uint64_t g_next_id;  // object_id.rs:5-16 — process-wide counter

struct Shape {
    // BaseShape, base_shape.rs:13-20
    bool casts_shadow_ = true;
    uint64_t id;
    Matrix t, t_inverse, t_inverse_transpose;
    Material m;
    int kind;

    explicit Shape(int k) : id(g_next_id++), kind(k) {}
    virtual ~Shape() = default;
    virtual std::unique_ptr<Shape> clone() const = 0;  // fresh id on clone, object_id.rs:27-31

    virtual void local_intersect(const Ray& object_ray, std::vector<Intersection>& out, Counters* k) const = 0;
    virtual Tuple local_norm_at(Tuple object_point, const Intersection& hit) const = 0;
    virtual BoundingBox bounding_box() const = 0;

    // base_shape.rs:56-60
    virtual void set_transformation(const Matrix& nt) {
        t = nt;
        t_inverse = t.inverse();
        t_inverse_transpose = t.inverse().transpose();
    }
    virtual void set_material(const Material& nm) { m = nm; }
    virtual void set_casts_shadow(bool c) { casts_shadow_ = c; }
    // shape.rs:60-70
    virtual void intersect(const Ray& world_ray, std::vector<Intersection>& out, Counters* k) const {
        if (k) k->xforms++;
        Ray object_ray = world_ray.transform(t_inverse);
        local_intersect(object_ray, out, k);
    }
    // shape.rs:72-146
    Tuple normal_to_world(Tuple object_normal) const {
        Tuple wn = t_inverse_transpose * object_normal;
        wn.w = 0.0f;
        return norm(wn);
    }
    // shape.rs:148-154
    Tuple normal_at(Tuple world_point, const Intersection& hit) const {
        Tuple object_point = t_inverse * world_point;
        Tuple object_normal = local_norm_at(object_point, hit);
        return normal_to_world(object_normal);
    }
    virtual bool includes(const Shape* other) const { return id == other->id; }                   // shape.rs:157-160
    virtual BoundingBox parent_space_bounding_box() const { return bounding_box().transform(t); }  // shape.rs:162-164
    virtual void divide(size_t) {}                                                                 // shape.rs:167
    virtual void prefill_caches() const {}
};
inline bool same_intersection(const Intersection& a, const Intersection& b) {
    return a.distance == b.distance && a.object->id == b.object->id && a.u == b.u && a.v == b.v;
}

struct Sphere : Shape {
    Tuple center = point(0, 0, 0);
    Sphere() : Shape(0) {}
    std::unique_ptr<Shape> clone() const override {
        auto s = std::make_unique<Sphere>(*this);
        s->id = g_next_id++;
        return s;
    }
    void local_intersect(const Ray& r, std::vector<Intersection>& out, Counters* k) const override {  // sphere.rs:47-70
        if (k) k->prim[0]++;
        Tuple sphere_to_ray = r.origin - center;
        float a = dot(r.direction, r.direction);
        float b = 2.0f * dot(r.direction, sphere_to_ray);
        float c = dot(sphere_to_ray, sphere_to_ray) - 1.0f;
        float discriminant = b * b - 4.0f * a * c;
        if (discriminant < 0.0f) return;
        float two_a = 2.0f * a;
        float ds = sqrtf(discriminant);
        out.push_back({(-b - ds) / two_a, this});
        out.push_back({(-b + ds) / two_a, this});
    }
    Tuple local_norm_at(Tuple p, const Intersection&) const override { return p - center; }  // :71-73
    BoundingBox bounding_box() const override {  // :75-80
        BoundingBox b;
        b.min = point(-1, -1, -1);
        b.max = point(1, 1, 1);
        return b;
    }
};

struct Plane : Shape {
    Plane() : Shape(1) {}
    std::unique_ptr<Shape> clone() const override {
        auto s = std::make_unique<Plane>(*this);
        s->id = g_next_id++;
        return s;
    }
    void local_intersect(const Ray& r, std::vector<Intersection>& out, Counters* k) const override {  // plane.rs:45-56
        if (k) k->prim[1]++;
        if (fabsf(r.direction.y) < kF32Epsilon * 10000.0f) return;
        out.push_back({-r.origin.y / r.direction.y, this});
    }
    Tuple local_norm_at(Tuple, const Intersection&) const override { return vector(0, 1, 0); }  // :57-59
    BoundingBox bounding_box() const override {  // :61-66
        BoundingBox b;
        b.min = point(-kInf, 0, -kInf);
        b.max = point(kInf, 0, kInf);
        return b;
    }
};

struct Cube : Shape {
    Cube() : Shape(2) {}
    std::unique_ptr<Shape> clone() const override {
        auto s = std::make_unique<Cube>(*this);
        s->id = g_next_id++;
        return s;
    }
    void local_intersect(const Ray& r, std::vector<Intersection>& out, Counters* k) const override {  // cube.rs:55-63
        if (k) k->prim[2]++;
        float lo, hi;
        if (aabb_intersection(r, point(-1, -1, -1), point(1, 1, 1), lo, hi)) {
            out.push_back({lo, this});
            out.push_back({hi, this});
        }
    }
    Tuple local_norm_at(Tuple p, const Intersection&) const override {  // :66-80
        float xa = fabsf(p.x), ya = fabsf(p.y), za = fabsf(p.z);
        float max_c = rmax(xa, rmax(ya, za));
        if (xa == max_c) return vector(p.x, 0, 0);
        if (ya == max_c) return vector(0, p.y, 0);
        return vector(0, 0, p.z);
    }
    BoundingBox bounding_box() const override {  // :82-87
        BoundingBox b;
        b.min = point(-1, -1, -1);
        b.max = point(1, 1, 1);
        return b;
    }
};

// shape/test_shape.rs:11-57 — the reference's own test double (records the object-space ray; normal is
// (2x, 3y, 4z)); only the transcribed unit tests use it.
struct TestShape : Shape {
    mutable Ray saved_ray;
    mutable bool has_saved_ray = false;
    TestShape() : Shape(9) {}
    std::unique_ptr<Shape> clone() const override {
        auto s = std::make_unique<TestShape>(*this);
        s->id = g_next_id++;
        return s;
    }
    void local_intersect(const Ray& r, std::vector<Intersection>&, Counters*) const override {
        saved_ray = r;
        has_saved_ray = true;
    }
    Tuple local_norm_at(Tuple p, const Intersection&) const override { return vector(2.0f * p.x, 3.0f * p.y, 4.0f * p.z); }
    BoundingBox bounding_box() const override {
        BoundingBox b;
        b.min = point(-1, -1, -1);
        b.max = point(1, 1, 1);
        return b;
    }
};

constexpr float kCloseToZero = 0.000001f;  // cylinder.rs:82, cone.rs:87

struct Cylinder : Shape {
    float minimum_y = -kInf, maximum_y = kInf;
    bool closed = false;
    Cylinder() : Shape(3) {}
    std::unique_ptr<Shape> clone() const override {
        auto s = std::make_unique<Cylinder>(*this);
        s->id = g_next_id++;
        return s;
    }
    void local_intersect(const Ray& r, std::vector<Intersection>& out, Counters* k) const override {  // cylinder.rs:52-59
        if (k) k->prim[3]++;
        size_t before = out.size();
        intersect_sides(r, out);
        if (out.size() - before < 2) intersect_caps(r, out);
    }
    void intersect_sides(const Ray& r, std::vector<Intersection>& out) const {  // :84-122
        float two_a = 2.0f * (r.direction.x * r.direction.x + r.direction.z * r.direction.z);
        if (fabsf(two_a) < kCloseToZero) return;
        float b = 2.0f * (r.origin.x * r.direction.x + r.origin.z * r.direction.z);
        float c = r.origin.x * r.origin.x + r.origin.z * r.origin.z - 1.0f;
        float discriminant = b * b - 2.0f * two_a * c;
        if (discriminant < 0.0f) return;
        float ds = sqrtf(discriminant);
        float d1 = (-b - ds) / two_a;
        float d2 = (-b + ds) / two_a;
        if (d1 > d2) std::swap(d1, d2);
        float y1 = r.origin.y + d1 * r.direction.y;
        if (minimum_y < y1 && y1 < maximum_y) out.push_back({d1, this});
        float y2 = r.origin.y + d2 * r.direction.y;
        if (minimum_y < y2 && y2 < maximum_y) out.push_back({d2, this});
    }
    static bool check_cap(const Ray& r, float t) {  // :125-130
        float x = r.origin.x + t * r.direction.x;
        float z = r.origin.z + t * r.direction.z;
        return (x * x + z * z) <= 1.0f + kCloseToZero;
    }
    void intersect_caps(const Ray& r, std::vector<Intersection>& out) const {  // :133-151
        if (!closed) return;
        float t = (minimum_y - r.origin.y) / r.direction.y;
        if (check_cap(r, t)) out.push_back({t, this});
        t = (maximum_y - r.origin.y) / r.direction.y;
        if (check_cap(r, t)) out.push_back({t, this});
    }
    Tuple local_norm_at(Tuple p, const Intersection&) const override {  // :62-72
        float dist_square = p.x * p.x + p.z * p.z;
        if (dist_square < 1.0f) {
            if (p.y >= maximum_y - kCloseToZero) return vector(0, 1, 0);
            if (p.y <= minimum_y + kCloseToZero) return vector(0, -1, 0);
        }
        return vector(p.x, 0, p.z);
    }
    BoundingBox bounding_box() const override {  // :74-79
        BoundingBox b;
        b.min = point(-1, minimum_y, -1);
        b.max = point(1, maximum_y, 1);
        return b;
    }
};

struct Cone : Shape {
    float minimum_y = -kInf, maximum_y = kInf;
    bool closed = false;
    Cone() : Shape(4) {}
    std::unique_ptr<Shape> clone() const override {
        auto s = std::make_unique<Cone>(*this);
        s->id = g_next_id++;
        return s;
    }
    void local_intersect(const Ray& r, std::vector<Intersection>& out, Counters* k) const override {  // cone.rs:52-57
        if (k) k->prim[4]++;
        intersect_sides(r, out);
        intersect_caps(r, out);
    }
    static float calc_c(const Ray& r) {  // :143-145
        return r.origin.x * r.origin.x - r.origin.y * r.origin.y + r.origin.z * r.origin.z;
    }
    void intersect_sides(const Ray& r, std::vector<Intersection>& out) const {  // :89-139
        float two_a = 2.0f * (r.direction.x * r.direction.x - r.direction.y * r.direction.y + r.direction.z * r.direction.z);
        float b = 2.0f * (r.origin.x * r.direction.x - r.origin.y * r.direction.y + r.origin.z * r.direction.z);
        if (fabsf(two_a) < kCloseToZero) {
            if (fabsf(b) < kCloseToZero) return;
            float c = calc_c(r);
            out.push_back({-c / (2.0f * b), this});
            return;
        }
        float c = calc_c(r);
        float discriminant = b * b - 2.0f * two_a * c;
        if (discriminant < 0.0f) return;
        float ds = sqrtf(discriminant);
        float d1 = (-b - ds) / two_a;
        float d2 = (-b + ds) / two_a;
        if (d1 > d2) std::swap(d1, d2);
        float y1 = r.origin.y + d1 * r.direction.y;
        if (minimum_y < y1 && y1 < maximum_y) out.push_back({d1, this});
        float y2 = r.origin.y + d2 * r.direction.y;
        if (minimum_y < y2 && y2 < maximum_y) out.push_back({d2, this});
    }
    static bool check_cap(float radius, const Ray& r, float t) {  // :148-153 — radius is |y|, not y^2
        float x = r.origin.x + t * r.direction.x;
        float z = r.origin.z + t * r.direction.z;
        return (x * x + z * z) <= radius + kCloseToZero;
    }
    void intersect_caps(const Ray& r, std::vector<Intersection>& out) const {  // :156-174
        if (!closed) return;
        float t = (minimum_y - r.origin.y) / r.direction.y;
        if (check_cap(fabsf(minimum_y), r, t)) out.push_back({t, this});
        t = (maximum_y - r.origin.y) / r.direction.y;
        if (check_cap(fabsf(maximum_y), r, t)) out.push_back({t, this});
    }
    Tuple local_norm_at(Tuple p, const Intersection&) const override {  // :60-73
        float dist_square = p.x * p.x + p.z * p.z;
        if (dist_square < 1.0f) {
            if (p.y >= maximum_y - kCloseToZero) return vector(0, 1, 0);
            if (p.y <= minimum_y + kCloseToZero) return vector(0, -1, 0);
        }
        float y = sqrtf(p.x * p.x + p.z * p.z);
        y = (p.y > 0.0f) ? -y : y;
        return vector(p.x, y, p.z);
    }
    BoundingBox bounding_box() const override {  // :75-84
        float limit = rmax(fabsf(minimum_y), fabsf(maximum_y));
        BoundingBox b;
        b.min = point(-limit, minimum_y, -limit);
        b.max = point(limit, maximum_y, limit);
        return b;
    }
};

struct Triangle : Shape {
    Tuple p1, p2, p3, e1, e2, normal;
    Triangle(Tuple a, Tuple b, Tuple c) : Shape(5), p1(a), p2(b), p3(c) {  // triangle.rs:20-33
        e1 = p2 - p1;
        e2 = p3 - p1;
        normal = norm(cross(e2, e1));
    }
    std::unique_ptr<Shape> clone() const override {
        auto s = std::make_unique<Triangle>(*this);
        s->id = g_next_id++;
        return s;
    }
    void local_intersect(const Ray& r, std::vector<Intersection>& out, Counters* k) const override {  // :45-76
        if (k) k->prim[5]++;
        Tuple dir_cross_e2 = cross(r.direction, e2);
        float determinant = dot(e1, dir_cross_e2);
        if (fabsf(determinant) < 0.0000001f) return;
        float f = 1.0f / determinant;
        Tuple p1_to_origin = r.origin - p1;
        float u = f * dot(p1_to_origin, dir_cross_e2);
        if (u < 0.0f || u > 1.0f) return;
        Tuple origin_cross_e1 = cross(p1_to_origin, e1);
        float v = f * dot(r.direction, origin_cross_e1);
        if (v < 0.0f || (u + v) > 1.0f) return;
        float t = f * dot(e2, origin_cross_e1);
        out.push_back({t, this, u, v});
    }
    Tuple local_norm_at(Tuple, const Intersection&) const override { return normal; }  // :78-81
    BoundingBox bounding_box() const override {  // :83-89
        BoundingBox b;
        b.add_point(p1);
        b.add_point(p2);
        b.add_point(p3);
        return b;
    }
};

// smooth_triangle.rs — local_intersect delegates to the INNER flat triangle (:39-41), so the hit's
// object is the inner Triangle and render-path shading uses the flat normal.
struct SmoothTriangle : Shape {
    Triangle base;
    Tuple n1, n2, n3;
    SmoothTriangle(Tuple a, Tuple b, Tuple c, Tuple na, Tuple nb, Tuple nc)
        : Shape(6), base(a, b, c), n1(na), n2(nb), n3(nc) {
        // get_base() forwards to the inner triangle's BaseShape (:30-36): one identity, one transform.
        id = base.id;
    }
    std::unique_ptr<Shape> clone() const override {
        auto s = std::make_unique<SmoothTriangle>(*this);
        s->base.id = g_next_id++;
        s->id = s->base.id;
        return s;
    }
    void set_transformation(const Matrix& nt) override {
        Shape::set_transformation(nt);
        base.set_transformation(nt);
    }
    void set_material(const Material& nm) override {
        Shape::set_material(nm);
        base.set_material(nm);
    }
    void set_casts_shadow(bool c) override {
        Shape::set_casts_shadow(c);
        base.set_casts_shadow(c);
    }
    void local_intersect(const Ray& r, std::vector<Intersection>& out, Counters* k) const override {
        base.local_intersect(r, out, k);
    }
    Tuple local_norm_at(Tuple, const Intersection& hit) const override {  // :43-46 (direct calls only)
        return n2 * hit.u + n3 * hit.v + n1 * (1.f - hit.u - hit.v);
    }
    BoundingBox bounding_box() const override { return base.bounding_box(); }  // :48-51
};

struct GroupShape : Shape {
    std::vector<std::unique_ptr<Shape>> children;
    mutable bool has_cached_bbox = false;  // group.rs:19 — cached on first use, never invalidated
    mutable BoundingBox cached_bbox;
    GroupShape() : Shape(7) {}
    std::unique_ptr<Shape> clone() const override {  // group.rs:175-183
        auto g = std::make_unique<GroupShape>();
        g->casts_shadow_ = casts_shadow_;
        g->t = t;
        g->t_inverse = t_inverse;
        g->t_inverse_transpose = t_inverse_transpose;
        g->m = m;
        for (auto& c : children) g->children.push_back(c->clone());
        return g;
    }
    void add_child(std::unique_ptr<Shape> child) {  // :39-44
        Matrix old = child->t;
        child->set_transformation(t * old);
        children.push_back(std::move(child));
    }
    bool includes(const Shape* other) const override {  // :87-93
        if (id == other->id) return true;
        for (auto& c : children)
            if (c->includes(other)) return true;
        return false;
    }
    void set_material(const Material& nm) override {  // :96-100
        for (auto& c : children) c->set_material(nm);
    }
    void set_transformation(const Matrix& nt) override {  // :101-114
        if (!children.empty()) {
            Matrix child_transformer = nt * t_inverse;
            for (auto& c : children) {
                Matrix old = c->t;
                c->set_transformation(child_transformer * old);
            }
        }
        Shape::set_transformation(nt);
    }
    void intersect(const Ray& world_ray, std::vector<Intersection>& out, Counters* k) const override {  // :115-118
        local_intersect(world_ray, out, k);
    }
    void local_intersect(const Ray& r, std::vector<Intersection>& out, Counters* k) const override {  // :119-133
        if (k) k->aabb++;
        BoundingBox b = bounding_box();
        if (!b.intersects(r)) return;
        for (auto& c : children) c->intersect(r, out, k);
    }
    Tuple local_norm_at(Tuple, const Intersection&) const override { return vector(0, 0, 0); }  // unreachable, :134-136
    BoundingBox bounding_box() const override {  // :138-150
        if (!has_cached_bbox) {
            BoundingBox b;
            for (auto& c : children) b.add_bounding_box(c->parent_space_bounding_box());
            cached_bbox = b;
            has_cached_bbox = true;
        }
        return cached_bbox;
    }
    BoundingBox parent_space_bounding_box() const override { return bounding_box(); }  // :152-156
    void partition_children(std::vector<std::unique_ptr<Shape>>& left, std::vector<std::unique_ptr<Shape>>& right) {  // :48-66
        BoundingBox lb, rb;
        bounding_box().split(lb, rb);
        std::vector<std::unique_ptr<Shape>> keep;
        for (auto& c : children) {
            BoundingBox cb = c->parent_space_bounding_box();
            if (lb.contains_bounding_box(cb))
                left.push_back(std::move(c));
            else if (rb.contains_bounding_box(cb))
                right.push_back(std::move(c));
            else
                keep.push_back(std::move(c));
        }
        children = std::move(keep);
    }
    void make_subgroup(std::vector<std::unique_ptr<Shape>>& kids) {  // :70-77
        if (kids.size() == 1) {
            children.push_back(std::move(kids[0]));
        } else {
            auto g = std::make_unique<GroupShape>();
            g->children = std::move(kids);
            children.push_back(std::move(g));
        }
    }
    void divide(size_t threshold) override {  // :158-172
        if (threshold <= children.size()) {
            std::vector<std::unique_ptr<Shape>> left, right;
            partition_children(left, right);
            if (!left.empty()) make_subgroup(left);
            if (!right.empty()) make_subgroup(right);
        }
        for (auto& c : children) c->divide(threshold);
    }
    void prefill_caches() const override {
        bounding_box();
        for (auto& c : children) c->prefill_caches();
    }
};

struct CSG : Shape {
    int op;
    std::unique_ptr<Shape> s1, s2;
    mutable bool has_cached_bbox = false;
    mutable BoundingBox cached_bbox;
    CSG(int o, std::unique_ptr<Shape> a, std::unique_ptr<Shape> b) : Shape(8), op(o), s1(std::move(a)), s2(std::move(b)) {}
    std::unique_ptr<Shape> clone() const override {  // csg.rs:137-147
        auto c = std::make_unique<CSG>(op, s1->clone(), s2->clone());
        c->casts_shadow_ = casts_shadow_;
        c->t = t;
        c->t_inverse = t_inverse;
        c->t_inverse_transpose = t_inverse_transpose;
        c->m = m;
        return c;
    }
    static bool intersection_allowed(int op, bool hit_s1, bool in_s1, bool in_s2) {  // :63-74
        switch (op) {
            case 0: return (hit_s1 && !in_s2) || (!hit_s1 && !in_s1);
            case 1: return (hit_s1 && in_s2) || (!hit_s1 && in_s1);
            default: return (hit_s1 && !in_s2) || (!hit_s1 && in_s1);
        }
    }
    void filter_intersections(const std::vector<Intersection>& xs, std::vector<Intersection>& out) const {  // :37-58
        bool in_s1 = false, in_s2 = false;
        for (const Intersection& i : xs) {
            bool hit_s1 = s1->includes(i.object);
            if (intersection_allowed(op, hit_s1, in_s1, in_s2)) out.push_back(i);
            if (hit_s1)
                in_s1 = !in_s1;
            else
                in_s2 = !in_s2;
        }
    }
    void local_intersect(const Ray& r, std::vector<Intersection>& out, Counters* k) const override {  // :87-104
        if (k) k->aabb++;
        if (!bounding_box().intersects(r)) return;
        std::vector<Intersection> xs;
        s1->intersect(r, xs, k);
        s2->intersect(r, xs, k);
        std::stable_sort(xs.begin(), xs.end(), [](const Intersection& a, const Intersection& b) { return a.distance < b.distance; });
        filter_intersections(xs, out);
    }
    Tuple local_norm_at(Tuple, const Intersection&) const override { return vector(0, 0, 0); }  // unimplemented!, :106-109
    bool includes(const Shape* other) const override {  // :111-117
        if (id == other->id) return true;
        return s1->includes(other) || s2->includes(other);
    }
    BoundingBox bounding_box() const override {  // :119-130
        if (!has_cached_bbox) {
            BoundingBox b;
            b.add_bounding_box(s1->parent_space_bounding_box());
            b.add_bounding_box(s2->parent_space_bounding_box());
            cached_bbox = b;
            has_cached_bbox = true;
        }
        return cached_bbox;
    }
    void divide(size_t threshold) override {  // :132-135
        s1->divide(threshold);
        s2->divide(threshold);
    }
    void prefill_caches() const override {
        bounding_box();
        s1->prefill_caches();
        s2->prefill_caches();
    }
};

// ---------------------------------------------------------------- pattern formulas
inline Color Pattern::color_at_world(Tuple p) const {
    switch (kind) {
        case 0:  // stripes.rs:39-45
            return (sat_i32(floorf(p.x)) % 2 == 0) ? a : b;
        case 1: {  // gradient.rs:33-36
            float fraction = p.x - floorf(p.x);
            return a + (distance * fraction);
        }
        case 2:  // rings.rs:38-50
            return (sat_i32(floorf(sqrtf(p.x * p.x + p.z * p.z))) % 2 == 0) ? a : b;
        case 3:  // checkers.rs:38-46 — |x|+|y|+|z|, not the book's sum of floors
            return (sat_i32(floorf(fabsf(p.x) + fabsf(p.y) + fabsf(p.z))) % 2 == 0) ? a : b;
        case 4: {  // sine_2d.rs:39-44
            float cosine = cosf(p.x + p.z);
            float fraction = (-cosine + 1.0f) / 2.0f;
            return a + (distance * fraction);
        }
        case 5:  // pattern.rs:84-86 (TestPattern)
            return {p.x, p.y, p.z};
        case 6: {  // uv.rs:89-92 (TextureMap)
            float u, v;
            if (mapping == 0)
                map_spherical(p, u, v);
            else if (mapping == 1)
                map_planar(p, u, v);
            else
                map_cylindrical(p, u, v);
            return uv->color_at(u, v);
        }
        default: {  // uv.rs:256-268 (CubicMap)
            Face f = face_from_point(p);
            float u, v;
            cube_uv(f, p, u, v);
            return faces[f]->color_at(u, v);
        }
    }
}
inline Color Pattern::color_at_object(Tuple world_point, const Shape& object) const {
    Tuple object_point = object.t_inverse * world_point;
    Tuple pattern_point = t_inverse * object_point;
    return color_at_world(pattern_point);
}

// ---------------------------------------------------------------- light/*.rs, world.rs
struct World;

// Counter-based stand-in for thread_rng().sample(OpenClosed01) (rectangle_light.rs:46): 24 random bits
// mapped to (k+1)*2^-24 in (0,1].  Shared bit-for-bit with the device path (dev_patterns.cuh: jitter_key / jitter_value).
// 32-bit arithmetic throughout (two rounds of a multiply-xorshift finaliser per draw): a shade of a 4x4 light draws 32
// values, and 64-bit multiplies cost a GPU six instructions each.  One key per (seed, pixel, path) — i.e. per
// intensity_at call — then value i = mix(key + i * golden ratio).
inline uint32_t jitter_mix32(uint32_t x) {
    x ^= x >> 16;
    x *= 0x7feb352du;
    x ^= x >> 15;
    x *= 0x846ca68bu;
    x ^= x >> 16;
    return x;
}
inline uint32_t jitter_key(uint64_t seed, uint32_t pixel, uint32_t path) {
    uint32_t k = jitter_mix32((uint32_t)seed ^ pixel);
    return jitter_mix32(k ^ (uint32_t)(seed >> 32) ^ (path * 0x9E3779B9u));
}
inline uint32_t jitter_hash(uint64_t seed, uint32_t pixel, uint32_t path, uint32_t index) {
    return jitter_mix32(jitter_key(seed, pixel, path) + index * 0x9E3779B9u);
}
inline float jitter_open_closed01(uint32_t bits) { return (float)((bits >> 8) + 1u) * 5.9604644775390625e-08f; }

// Per-ray context for the counter generator: which pixel, and which branch of the reflect/refract tree.
struct PathCtx {
    uint32_t pixel = 0;
    uint32_t path = 1;  // 1 = primary; child = path*3+1 (reflect) / path*3+2 (refract)
};

struct Light {  // light.rs:5-11
    bool is_rect = false;
    Color intensity = white();
    Tuple position = point(0, 0, 0);
    // RectangleLight, rectangle_light.rs:12-31
    Tuple corner, u_vec, v_vec;  // u_vec / v_vec are per-cell after construction (:52-53)
    int u_steps = 1, v_steps = 1, cells = 1;
    std::vector<float> jitter;  // cyclic table; empty = counter generator
    uint64_t seed = 0;

    static Light point_light(Tuple pos, Color i) {  // point_light.rs:12-18
        Light l;
        l.position = pos;
        l.intensity = i;
        return l;
    }
    static Light rectangle(Color i, Tuple corner, Tuple full_u, int us, Tuple full_v, int vs) {  // rectangle_light.rs:34-59
        Light l;
        l.is_rect = true;
        l.intensity = i;
        l.corner = corner;
        l.u_vec = full_u / (float)us;
        l.v_vec = full_v / (float)vs;
        l.u_steps = us;
        l.v_steps = vs;
        l.cells = us * vs;
        l.position = corner + (full_u / 2.f) + (full_v / 2.f);
        return l;
    }
    // rectangle_light.rs:60-66 with the two jitter values made explicit
    Tuple point_on_light(int u, int v, float j1, float j2) const {
        return corner + u_vec * ((float)u + j1) + v_vec * ((float)v + j2);
    }
    float intensity_at(Tuple p, const World& w, const PathCtx& ctx, Counters* k) const;
};

// world.rs:165-182
struct Comps {
    float distance;
    const Shape* object;
    Tuple point, eye_vector, reflection_vector, surface_normal;
    bool inside;
    Tuple over_point, under_point;
    float n1, n2;
};

constexpr float kAcneEpsilon = kF32Epsilon * 10000.0f;  // world.rs:210

// intersection.rs:30-35 — first minimum among t >= 0
inline const Intersection* hit(const std::vector<Intersection>& xs) {
    const Intersection* best = nullptr;
    for (const Intersection& i : xs) {
        if (!(i.distance >= 0.0f)) continue;
        if (!best || i.distance < best->distance) best = &i;
    }
    return best;
}

// world.rs:212-283
inline Comps precompute_values(const Ray& r, const Intersection& h, const std::vector<Intersection>& xs) {
    Comps c;
    c.point = r.position(h.distance);
    Tuple n = h.object->normal_at(c.point, h);
    c.eye_vector = -r.direction;
    c.reflection_vector = reflect(r.direction, n);
    if (dot(n, c.eye_vector) < 0.0f) {
        c.inside = true;
        n = -n;
    } else {
        c.inside = false;
    }
    c.surface_normal = n;
    c.over_point = c.point + n * kAcneEpsilon;
    c.under_point = c.point - n * kAcneEpsilon;
    c.distance = h.distance;
    c.object = h.object;
    // n1 / n2 — insertion-ordered set of containing shapes (linked_hash_set 0.1.3: remove, else insert
    // at the back; back() is the most recently inserted), world.rs:235-263
    float n1 = NAN, n2 = NAN;
    std::vector<const Shape*> containers;
    for (const Intersection& i : xs) {
        bool is_hit = same_intersection(i, h);
        if (is_hit) n1 = containers.empty() ? 1.0f : containers.back()->m.refractive_index;
        auto it = std::find_if(containers.begin(), containers.end(), [&](const Shape* s) { return s->id == i.object->id; });
        if (it != containers.end())
            containers.erase(it);
        else
            containers.push_back(i.object);
        if (is_hit) {
            n2 = containers.empty() ? 1.0f : containers.back()->m.refractive_index;
            break;
        }
    }
    c.n1 = n1;
    c.n2 = n2;
    return c;
}

inline float powi5(float x) {  // llvm powi expansion for a constant 5: x * (x^2)^2
    float x2 = x * x;
    return x * (x2 * x2);
}
// world.rs:285-303
inline float schlick_reflectance(const Comps& c) {
    float cosine = dot(c.eye_vector, c.surface_normal);
    if (c.n1 > c.n2) {
        float n = c.n1 / c.n2;
        float sin2_refracted = (n * n) * (1.0f - cosine * cosine);
        if (sin2_refracted > 1.0f) return 1.0f;
        cosine = sqrtf(1.0f - sin2_refracted);
    }
    float q = (c.n1 - c.n2) / (c.n1 + c.n2);
    float r0 = q * q;
    return r0 + (1.0f - r0) * powi5(1.0f - cosine);
}

// phong_lighting.rs:12-63
inline Color phong_lighting(const Shape& object, const Material& material, const Light& light, Tuple point,
                            Tuple eye_vector, Tuple surface_normal, float light_intensity, Counters* k = nullptr) {
    Color material_color = material.color;
    if (material.pattern) {
        if (k) k->patterns++;
        material_color = material.pattern->color_at_object(point, object);
    }
    Color effective_color = material_color * light.intensity;
    Color ambient = effective_color * material.ambient;
    if (light_intensity == 0.f) return ambient;
    Tuple to_light = norm(light.position - point);
    float light_normal_cosine = dot(to_light, surface_normal);
    Color diffuse, specular;
    if (light_normal_cosine < 0.0f) {
        diffuse = black();
        specular = black();
    } else {
        diffuse = effective_color * material.diffuse * light_normal_cosine;
        Tuple surface_reflection = reflect(-to_light, surface_normal);
        float reflection_eye_cosine = dot(surface_reflection, eye_vector);
        if (reflection_eye_cosine <= 0.0f) {
            specular = black();
        } else {
            float factor = powf(reflection_eye_cosine, material.shininess);
            specular = light.intensity * material.specular * factor;
        }
    }
    return ambient + (diffuse + specular) * light_intensity;
}

struct World {  // world.rs:18-21
    std::vector<Shape*> objects;
    bool has_light = false;
    Light light;

    // world.rs:52-60
    void intersect(const Ray& r, std::vector<Intersection>& xs, Counters* k) const {
        for (const Shape* o : objects) o->intersect(r, xs, k);
        std::stable_sort(xs.begin(), xs.end(), [](const Intersection& a, const Intersection& b) { return a.distance < b.distance; });
    }
    // world.rs:104-119
    bool is_shadowed(Tuple light_position, Tuple p, Counters* k) const {
        if (k) k->shadow++;
        Tuple v = light_position - p;
        float distance = magnitude(v);
        Tuple direction = norm(v);
        Ray r(p, direction);
        std::vector<Intersection> xs;
        intersect(r, xs, k);
        const Intersection* h = hit(xs);
        return h ? (h->object->casts_shadow_ && h->distance < distance) : false;
    }
    // world.rs:121-133
    Color reflected_color(const Comps& c, int remaining, PathCtx ctx, Counters* k) const {
        if (c.object->m.reflective == 0.0f || remaining < 1) return black();
        Ray rr(c.over_point, c.reflection_vector);
        if (k) k->secondary++;
        ctx.path = ctx.path * 3u + 1u;
        Color col = color_at(rr, remaining - 1, ctx, k);
        return col * c.object->m.reflective;
    }
    // world.rs:135-162 (+ refracted_angle_values :196-207)
    Color refracted_color(const Comps& c, int remaining, PathCtx ctx, Counters* k) const {
        if (c.object->m.transparency == 0.0f || remaining == 0) return black();
        float n_ratio = c.n1 / c.n2;
        float cos_incoming = dot(c.eye_vector, c.surface_normal);
        float sin2 = (n_ratio * n_ratio) * (1.0f - cos_incoming * cos_incoming);
        if (sin2 > 1.0f) return black();
        if (k) k->refr_dirs++;
        float cos_refracted = sqrtf(1.0f - sin2);
        Tuple direction = c.surface_normal * (n_ratio * cos_incoming - cos_refracted) - (c.eye_vector * n_ratio);
        Ray rr(c.under_point, direction);
        if (k) k->secondary++;
        ctx.path = ctx.path * 3u + 2u;
        return color_at(rr, remaining - 1, ctx, k) * c.object->m.transparency;
    }
    // world.rs:62-86
    Color shade_hit(const Comps& c, int remaining, PathCtx ctx, Counters* k) const {
        if (k) k->shades++;
        const Material& mat = c.object->m;
        float li = light.intensity_at(c.over_point, *this, ctx, k);
        Color surface = phong_lighting(*c.object, mat, light, c.over_point, c.eye_vector, c.surface_normal, li, k);
        Color reflected = reflected_color(c, remaining, ctx, k);
        Color refracted = refracted_color(c, remaining, ctx, k);
        if (k) k->combines++;
        if (mat.reflective > 0.0f && mat.transparency > 0.0f) {
            if (k) k->schlick++;
            float reflectance = schlick_reflectance(c);
            return surface + reflected * reflectance + refracted * (1.0f - reflectance);
        }
        return surface + reflected + refracted;
    }
    // world.rs:88-101
    Color color_at(const Ray& r, int remaining, PathCtx ctx, Counters* k) const {
        std::vector<Intersection> xs;
        intersect(r, xs, k);
        if (xs.empty()) return black();
        const Intersection* h = hit(xs);
        if (!h) return black();
        Comps c = precompute_values(r, *h, xs);
        return shade_hit(c, remaining, ctx, k);
    }
};

inline float Light::intensity_at(Tuple p, const World& w, const PathCtx& ctx, Counters* k) const {
    if (!is_rect) return w.is_shadowed(position, p, k) ? 0.f : 1.f;  // point_light.rs:28-34
    // rectangle_light.rs:76-88 — v outer, u inner, two jitter draws per cell (u first)
    float total = 0.f;
    size_t cursor = 0;
    for (int v = 0; v < v_steps; v++) {
        for (int u = 0; u < u_steps; u++) {
            float j1, j2;
            if (!jitter.empty()) {
                j1 = jitter[cursor % jitter.size()];
                j2 = jitter[(cursor + 1) % jitter.size()];
            } else {
                j1 = jitter_open_closed01(jitter_hash(seed, ctx.pixel, ctx.path, (uint32_t)cursor));
                j2 = jitter_open_closed01(jitter_hash(seed, ctx.pixel, ctx.path, (uint32_t)cursor + 1u));
            }
            cursor += 2;
            if (k) k->cells++;
            Tuple lp = point_on_light(u, v, j1, j2);
            if (!w.is_shadowed(lp, p, k)) total += 1.0f;
        }
    }
    return total / (float)cells;
}

// ---------------------------------------------------------------- camera.rs
struct Camera {
    uint32_t width, height;
    float field_of_view, half_width, half_height, pixel_size;
    Matrix transform_inverse;
    Camera(uint32_t w, uint32_t h, float fov, const Matrix& transform) : width(w), height(h), field_of_view(fov) {  // :23-56
        float half_view = tanf(fov / 2.0f);
        float aspect = (float)w / (float)h;
        if (aspect >= 1.0f) {
            half_width = half_view;
            half_height = half_view / aspect;
        } else {
            half_width = half_view * aspect;
            half_height = half_view;
        }
        pixel_size = (half_width * 2.0f) / (float)w;
        transform_inverse = transform.inverse();
    }
    Ray ray_for_pixel(uint32_t x, uint32_t y) const {  // :60-74
        float x_offset = ((float)x + 0.5f) * pixel_size;
        float y_offset = ((float)y + 0.5f) * pixel_size;
        float world_x = half_width - x_offset;
        float world_y = half_height - y_offset;
        Tuple pixel = transform_inverse * point(world_x, world_y, -1);
        Tuple origin = transform_inverse * point(0, 0, 0);
        Tuple direction = norm(pixel - origin);
        return Ray(origin, direction);
    }
};

// canvas.rs:39-43 — clamp then TRUNCATE; NaN.min(255) = 255
inline uint8_t scale_color(float c) {
    float s = rmax(rmin(c * 255.0f, 255.0f), 0.0f);
    return (uint8_t)s;
}

}  // namespace orc
