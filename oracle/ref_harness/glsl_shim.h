// TEST INFRASTRUCTURE (oracle side). A GLSL-450 vocabulary for g++, just large enough to compile the reference's own
// compute shaders (pathfinder/shaders/d3d11/fill.comp, tile.comp) as C++ and run them on the CPU, one invocation at a
// time: vectors with the swizzles those two shaders use, the built-in functions they call, sampler2D / image2D objects
// over host memory. oracle/ref_harness/make_shader_cpp.py turns a shader's text -- read where it lies under
// /root/reference -- into a translation unit that includes this header; nothing of the shader is stored in this repo.
//
// Arithmetic is IEEE fp32, one rounding per GLSL operation (compile with -ffp-contract=off); mix / mod / clamp follow
// the GLSL 4.50 specification's formulas; texture() filters with exact fp32 weights (GL 4.5 section 8.14: texel
// centres at i + 0.5, CLAMP_TO_EDGE or REPEAT) except within 2^-9 texel of a texel centre (see texture()); imageStore on an rgba8 image converts like the specification's
// float-to-unorm rule, round(clamp(c, 0, 1) * 255).
#pragma once
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstring>

namespace glsl {

typedef unsigned int uint;

struct vec2;
struct vec3;
struct vec4;

// ---- swizzle proxies: they live in a union with the vector's components
template <int A, int B>
struct Swz2 {
    float d[4];
    operator vec2() const;
    Swz2 &operator=(const vec2 &v);
    Swz2 &operator*=(const vec2 &v);
    Swz2 &operator+=(const vec2 &v);
    Swz2 &operator*=(float s);
};
template <int A, int B, int C>
struct Swz3 {
    float d[4];
    operator vec3() const;
    Swz3 &operator=(const vec3 &v);
    Swz3 &operator*=(float s);
    Swz3 &operator*=(const vec3 &v);
};
template <int A, int B, int C, int D>
struct Swz4 {
    float d[4];
    operator vec4() const;
};

struct vec2 {
    union {
        struct { float x, y; };
        struct { float r, g; };
        float d[2];
        Swz2<0, 1> xy;
        Swz2<1, 0> yx;
        Swz4<0, 1, 0, 1> xyxy;
    };
    vec2() : x(0), y(0) {}
    explicit vec2(float s) : x(s), y(s) {}
    vec2(float a, float b) : x(a), y(b) {}
    explicit vec2(const struct ivec2 &v);
    explicit vec2(const struct uvec2 &v);
    float &operator[](int i) { return d[i]; }
    float operator[](int i) const { return d[i]; }
};

struct vec3 {
    union {
        struct { float x, y, z; };
        struct { float r, g, b; };
        float d[3];
        Swz2<0, 1> xy; Swz2<1, 2> yz; Swz2<0, 1> rg; Swz2<1, 2> gb; Swz2<2, 0> br;
        Swz3<0, 1, 2> xyz; Swz3<0, 1, 2> rgb; Swz3<2, 1, 0> zyx; Swz3<2, 2, 2> zzz; Swz3<0, 0, 0> rrr;
    };
    vec3() : x(0), y(0), z(0) {}
    explicit vec3(float s) : x(s), y(s), z(s) {}
    vec3(float a, float b_, float c) : x(a), y(b_), z(c) {}
    vec3(const vec2 &v, float c) : x(v.x), y(v.y), z(c) {}
    vec3(float a, const vec2 &v) : x(a), y(v.x), z(v.y) {}
    float &operator[](int i) { return d[i]; }
    float operator[](int i) const { return d[i]; }
};

struct vec4 {
    union {
        struct { float x, y, z, w; };
        struct { float r, g, b, a; };
        float d[4];
        Swz2<0, 1> xy; Swz2<2, 3> zw; Swz2<1, 2> yz; Swz2<0, 1> rg;
        Swz3<0, 1, 2> xyz; Swz3<0, 1, 2> rgb; Swz3<1, 2, 3> yzw; Swz3<2, 1, 0> zyx; Swz3<0, 0, 0> rrr;
        Swz4<2, 3, 0, 1> zwxy;
    };
    vec4() : x(0), y(0), z(0), w(0) {}
    explicit vec4(float s) : x(s), y(s), z(s), w(s) {}
    vec4(float a_, float b_, float c, float d_) : x(a_), y(b_), z(c), w(d_) {}
    vec4(const vec3 &v, float d_) : x(v.x), y(v.y), z(v.z), w(d_) {}
    vec4(const vec2 &p, const vec2 &q) : x(p.x), y(p.y), z(q.x), w(q.y) {}
    vec4(const vec2 &p, float c, float d_) : x(p.x), y(p.y), z(c), w(d_) {}
    // GLSL constructors convert every scalar argument on their own (uint / int arguments in fill.comp)
    explicit vec4(const struct ivec4 &v);
    template <class A, class B, class C, class D>
    vec4(A a_, B b_, C c, D d_) : x((float)a_), y((float)b_), z((float)c), w((float)d_) {}
    float &operator[](int i) { return d[i]; }
    float operator[](int i) const { return d[i]; }
};

template <int A, int B> Swz2<A, B>::operator vec2() const { return vec2(d[A], d[B]); }
template <int A, int B> Swz2<A, B> &Swz2<A, B>::operator=(const vec2 &v) { d[A] = v.x; d[B] = v.y; return *this; }
template <int A, int B> Swz2<A, B> &Swz2<A, B>::operator*=(const vec2 &v) { const float a = d[A] * v.x, b = d[B] * v.y; d[A] = a; d[B] = b; return *this; }
template <int A, int B> Swz2<A, B> &Swz2<A, B>::operator+=(const vec2 &v) { const float a = d[A] + v.x, b = d[B] + v.y; d[A] = a; d[B] = b; return *this; }
template <int A, int B> Swz2<A, B> &Swz2<A, B>::operator*=(float s) { d[A] *= s; d[B] *= s; return *this; }
template <int A, int B, int C> Swz3<A, B, C>::operator vec3() const { return vec3(d[A], d[B], d[C]); }
template <int A, int B, int C> Swz3<A, B, C> &Swz3<A, B, C>::operator=(const vec3 &v) { d[A] = v.x; d[B] = v.y; d[C] = v.z; return *this; }
template <int A, int B, int C> Swz3<A, B, C> &Swz3<A, B, C>::operator*=(float s) { d[A] *= s; d[B] *= s; d[C] *= s; return *this; }
template <int A, int B, int C> Swz3<A, B, C> &Swz3<A, B, C>::operator*=(const vec3 &v) { d[A] *= v.x; d[B] *= v.y; d[C] *= v.z; return *this; }
template <int A, int B, int C, int D> Swz4<A, B, C, D>::operator vec4() const { return vec4(d[A], d[B], d[C], d[D]); }

struct ivec4;
template <int A, int B, int C_, int D>
struct ISwz4 {
    int d[4];
    operator ivec4() const;
};
struct ivec2 {
    union {
        struct { int x, y; };
        ISwz4<0, 1, 0, 1> xyxy;
    };
    ivec2() : x(0), y(0) {}
    explicit ivec2(int s) : x(s), y(s) {}
    ivec2(int a, int b) : x(a), y(b) {}
    ivec2(uint a, uint b) : x((int)a), y((int)b) {}
    ivec2(int a, uint b) : x(a), y((int)b) {}
    ivec2(uint a, int b) : x((int)a), y(b) {}
    explicit ivec2(const vec2 &v) : x((int)v.x), y((int)v.y) {}
    explicit ivec2(const struct uvec2 &v);
};
struct uvec2 {
    uint x, y;
    uvec2() : x(0), y(0) {}
    explicit uvec2(uint s) : x(s), y(s) {}
    explicit uvec2(int s) : x((uint)s), y((uint)s) {}
    uvec2(uint a, uint b) : x(a), y(b) {}
    uvec2(int a, int b) : x((uint)a), y((uint)b) {}
    uvec2(int a, uint b) : x((uint)a), y(b) {}
    uvec2(uint a, int b) : x(a), y((uint)b) {}
    explicit uvec2(const vec2 &v) : x((uint)v.x), y((uint)v.y) {}
};
inline uvec2 operator*(uvec2 a, uvec2 b) { return uvec2(a.x * b.x, a.y * b.y); }
inline uvec2 operator+(uvec2 a, uvec2 b) { return uvec2(a.x + b.x, a.y + b.y); }
template <int A, int B>
struct USwz2 {
    uint d[4];
    operator uvec2() const { return uvec2(d[A], d[B]); }
};
struct uvec4 {
    union {
        struct { uint x, y, z, w; };
        USwz2<0, 1> xy;
        USwz2<2, 3> zw;
    };
    uvec4() : x(0), y(0), z(0), w(0) {}
    explicit uvec4(uint s) : x(s), y(s), z(s), w(s) {}
    uvec4(uint a, uint b, uint c, uint d_) : x(a), y(b), z(c), w(d_) {}
    explicit uvec4(const struct vec4 &v);
};
inline uvec2 operator-(uvec2 a, uvec2 b) { return uvec2(a.x - b.x, a.y - b.y); }
template <int A, int B>
struct ISwz2 {
    int d[4];
    operator ivec2() const { return ivec2(d[A], d[B]); }
};
struct ivec4 {
    union {
        struct { int x, y, z, w; };
        ISwz2<0, 1> xy;
        ISwz2<2, 3> zw;
    };
    ivec4() : x(0), y(0), z(0), w(0) {}
    explicit ivec4(int s) : x(s), y(s), z(s), w(s) {}
    ivec4(int a, int b, int c, int d_) : x(a), y(b), z(c), w(d_) {}
    explicit ivec4(const uvec4 &v) : x((int)v.x), y((int)v.y), z((int)v.z), w((int)v.w) {}
    explicit ivec4(const vec4 &v);
};
template <int A, int B, int C_, int D> ISwz4<A, B, C_, D>::operator ivec4() const { return ivec4(d[A], d[B], d[C_], d[D]); }
inline ivec4 operator-(ivec4 a, ivec4 b) { return ivec4(a.x - b.x, a.y - b.y, a.z - b.z, a.w - b.w); }
inline ivec4 operator+(ivec4 a, ivec4 b) { return ivec4(a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w); }
inline ivec4 operator*(ivec4 a, ivec4 b) { return ivec4(a.x * b.x, a.y * b.y, a.z * b.z, a.w * b.w); }
inline ivec4 operator*(ivec4 a, int s) { return ivec4(a.x * s, a.y * s, a.z * s, a.w * s); }
struct uvec3 {
    uint x, y, z;
    struct XY { uint x, y; } ;
    uvec2 xy_() const { return uvec2(x, y); }
};
inline ivec2::ivec2(const uvec2 &v) : x((int)v.x), y((int)v.y) {}
inline vec2::vec2(const ivec2 &v) : x((float)v.x), y((float)v.y) {}
inline vec2::vec2(const uvec2 &v) : x((float)v.x), y((float)v.y) {}
struct bvec3 {
    bool x, y, z;
    bvec3(bool a, bool b_, bool c) : x(a), y(b_), z(c) {}
};
struct bvec2 {
    bool x, y;
    bvec2(bool a, bool b_) : x(a), y(b_) {}
};

inline bool operator==(ivec2 a, ivec2 b) { return a.x == b.x && a.y == b.y; }
inline bool operator!=(ivec2 a, ivec2 b) { return !(a == b); }
inline ivec2 operator+(ivec2 a, ivec2 b) { return ivec2(a.x + b.x, a.y + b.y); }
inline ivec2 operator-(ivec2 a, ivec2 b) { return ivec2(a.x - b.x, a.y - b.y); }
inline ivec2 operator*(ivec2 a, ivec2 b) { return ivec2(a.x * b.x, a.y * b.y); }
inline ivec2 operator*(ivec2 a, int s) { return ivec2(a.x * s, a.y * s); }
inline ivec2 operator/(ivec2 a, int s) { return ivec2(a.x / s, a.y / s); }
inline ivec2 operator/(ivec2 a, ivec2 b) { return ivec2(a.x / b.x, a.y / b.y); }
inline ivec2 operator%(ivec2 a, ivec2 b) { return ivec2(a.x % b.x, a.y % b.y); }

// ---- componentwise arithmetic (non-template on purpose: swizzle proxies convert implicitly)
#define GLSL_BIN(V, OP)                                                                        \
    inline V operator OP(const V &a, const V &b) { V r; for (int i = 0; i < (int)(sizeof(a.d) / 4); i++) r.d[i] = a.d[i] OP b.d[i]; return r; } \
    inline V operator OP(const V &a, float s) { V r; for (int i = 0; i < (int)(sizeof(a.d) / 4); i++) r.d[i] = a.d[i] OP s; return r; }          \
    inline V operator OP(float s, const V &a) { V r; for (int i = 0; i < (int)(sizeof(a.d) / 4); i++) r.d[i] = s OP a.d[i]; return r; }          \
    inline V &operator OP##=(V &a, const V &b) { for (int i = 0; i < (int)(sizeof(a.d) / 4); i++) a.d[i] = a.d[i] OP b.d[i]; return a; }          \
    inline V &operator OP##=(V &a, float s) { for (int i = 0; i < (int)(sizeof(a.d) / 4); i++) a.d[i] = a.d[i] OP s; return a; }
#define GLSL_VEC(V) GLSL_BIN(V, +) GLSL_BIN(V, -) GLSL_BIN(V, *) GLSL_BIN(V, /)                \
    inline V operator-(const V &a) { V r; for (int i = 0; i < (int)(sizeof(a.d) / 4); i++) r.d[i] = -a.d[i]; return r; }
GLSL_VEC(vec2)
GLSL_VEC(vec3)
GLSL_VEC(vec4)
inline vec4::vec4(const ivec4 &v) : x((float)v.x), y((float)v.y), z((float)v.z), w((float)v.w) {}
inline ivec4::ivec4(const vec4 &v) : x((int)v.x), y((int)v.y), z((int)v.z), w((int)v.w) {}
inline uvec4::uvec4(const vec4 &v) : x((uint)v.x), y((uint)v.y), z((uint)v.z), w((uint)v.w) {}

// ---- built-in functions, scalar
inline float abs(float a) { return std::fabs(a); }
inline int abs(int a) { return a < 0 ? -a : a; }
inline float sign(float a) { return a > 0.0f ? 1.0f : (a < 0.0f ? -1.0f : 0.0f); }
inline float floor(float a) { return std::floor(a); }
inline float ceil(float a) { return std::ceil(a); }
inline float fract(float a) { return a - std::floor(a); }
inline float round(float a) { return std::nearbyint(a); }
inline float sqrt(float a) { return std::sqrt(a); }
inline float inversesqrt(float a) { return 1.0f / std::sqrt(a); }
inline float exp(float a) { return std::exp(a); }
inline float pow(float a, float b) { return std::pow(a, b); }
inline float min(float a, float b) { return b < a ? b : a; }
inline float max(float a, float b) { return a < b ? b : a; }
inline int min(int a, int b) { return b < a ? b : a; }
inline int max(int a, int b) { return a < b ? b : a; }
inline uint min(uint a, uint b) { return b < a ? b : a; }
inline uint max(uint a, uint b) { return a < b ? b : a; }
inline float clamp(float v, float lo, float hi) { return min(max(v, lo), hi); }
inline int clamp(int v, int lo, int hi) { return min(max(v, lo), hi); }
inline float mix(float a, float b, float t) { return a * (1.0f - t) + b * t; }
inline float mod(float a, float b) { return a - b * std::floor(a / b); }
inline float step(float edge, float v) { return v < edge ? 0.0f : 1.0f; }

// ---- built-in functions, vectors
#define GLSL_MAP1(V, F) inline V F(const V &a) { V r; for (int i = 0; i < (int)(sizeof(a.d) / 4); i++) r.d[i] = F(a.d[i]); return r; }
#define GLSL_MAP2(V, F)                                                                                                     \
    inline V F(const V &a, const V &b) { V r; for (int i = 0; i < (int)(sizeof(a.d) / 4); i++) r.d[i] = F(a.d[i], b.d[i]); return r; } \
    inline V F(const V &a, float s) { V r; for (int i = 0; i < (int)(sizeof(a.d) / 4); i++) r.d[i] = F(a.d[i], s); return r; }
#define GLSL_FUNCS(V)                                                                                                     \
    GLSL_MAP1(V, round) GLSL_MAP1(V, abs) GLSL_MAP1(V, sign) GLSL_MAP1(V, floor) GLSL_MAP1(V, ceil) GLSL_MAP1(V, fract) GLSL_MAP1(V, sqrt)    \
    GLSL_MAP1(V, exp) GLSL_MAP2(V, min) GLSL_MAP2(V, max) GLSL_MAP2(V, mod) GLSL_MAP2(V, pow)                                    \
    inline V clamp(const V &v, float lo, float hi) { V r; for (int i = 0; i < (int)(sizeof(v.d) / 4); i++) r.d[i] = clamp(v.d[i], lo, hi); return r; } \
    inline V clamp(const V &v, const V &lo, const V &hi) { V r; for (int i = 0; i < (int)(sizeof(v.d) / 4); i++) r.d[i] = clamp(v.d[i], lo.d[i], hi.d[i]); return r; } \
    inline V mix(const V &a, const V &b, float t) { V r; for (int i = 0; i < (int)(sizeof(a.d) / 4); i++) r.d[i] = mix(a.d[i], b.d[i], t); return r; } \
    inline V mix(const V &a, const V &b, const V &t) { V r; for (int i = 0; i < (int)(sizeof(a.d) / 4); i++) r.d[i] = mix(a.d[i], b.d[i], t.d[i]); return r; } \
    inline V step(const V &e, const V &v) { V r; for (int i = 0; i < (int)(sizeof(v.d) / 4); i++) r.d[i] = step(e.d[i], v.d[i]); return r; } \
    inline float dot(const V &a, const V &b) { float s = a.d[0] * b.d[0]; for (int i = 1; i < (int)(sizeof(a.d) / 4); i++) s = s + a.d[i] * b.d[i]; return s; } \
    inline float length(const V &a) { return std::sqrt(dot(a, a)); }                                                      \
    inline V normalize(const V &a) { return a / length(a); }
GLSL_FUNCS(vec2)
GLSL_FUNCS(vec3)
GLSL_FUNCS(vec4)

inline bvec3 lessThanEqual(const vec3 &a, const vec3 &b) { return bvec3(a.x <= b.x, a.y <= b.y, a.z <= b.z); }
inline bvec3 lessThan(const vec3 &a, const vec3 &b) { return bvec3(a.x < b.x, a.y < b.y, a.z < b.z); }
inline bvec3 equal(const vec3 &a, const vec3 &b) { return bvec3(a.x == b.x, a.y == b.y, a.z == b.z); }
inline bvec2 lessThanEqual(const vec2 &a, const vec2 &b) { return bvec2(a.x <= b.x, a.y <= b.y); }
struct bvec4 {
    bool x, y, z, w;
    bvec4(const bvec2 &a, const bvec2 &b) : x(a.x), y(a.y), z(b.x), w(b.y) {}
};
inline bvec2 greaterThanEqual(uvec2 a, uvec2 b) { return bvec2(a.x >= b.x, a.y >= b.y); }
inline bvec2 lessThan(uvec2 a, uvec2 b) { return bvec2(a.x < b.x, a.y < b.y); }
inline bool all(const bvec4 &v) { return v.x && v.y && v.z && v.w; }
inline bool any(const bvec4 &v) { return v.x || v.y || v.z || v.w; }
inline bvec2 lessThan(ivec2 a, ivec2 b) { return bvec2(a.x < b.x, a.y < b.y); }
inline bvec2 greaterThanEqual(ivec2 a, ivec2 b) { return bvec2(a.x >= b.x, a.y >= b.y); }
inline bool all(const bvec2 &v) { return v.x && v.y; }
// atomics on buffer elements: invocations run one after the other, so plain read-modify-write
inline int atomicAdd(int &mem, int v) { const int old = mem; mem = old + v; return old; }
inline uint atomicAdd(uint &mem, uint v) { const uint old = mem; mem = old + v; return old; }
inline int atomicMax(int &mem, int v) { const int old = mem; if (v > old) mem = v; return old; }
inline int atomicExchange(int &mem, int v) { const int old = mem; mem = v; return old; }
inline uint atomicExchange(uint &mem, uint v) { const uint old = mem; mem = v; return old; }
inline bool all(const bvec3 &v) { return v.x && v.y && v.z; }
inline bool any(const bvec3 &v) { return v.x || v.y || v.z; }

struct mat2 {
    vec2 c0, c1;  // columns
    mat2() {}
    mat2(const vec2 &a, const vec2 &b) : c0(a), c1(b) {}
    explicit mat2(const vec4 &v) : c0(v.x, v.y), c1(v.z, v.w) {}
    mat2(float a, float b, float c, float d) : c0(a, b), c1(c, d) {}
};
inline vec2 operator*(const mat2 &m, const vec2 &v) { return vec2(m.c0.x * v.x + m.c1.x * v.y, m.c0.y * v.x + m.c1.y * v.y); }
struct mat4 {
    vec4 c[4];
    mat4() {}
    mat4(const vec4 &a, const vec4 &b, const vec4 &c_, const vec4 &d) { c[0] = a; c[1] = b; c[2] = c_; c[3] = d; }
    vec4 &operator[](int i) { return c[i]; }
    const vec4 &operator[](int i) const { return c[i]; }
};
inline vec4 operator*(const mat4 &m, const vec4 &v) {
    vec4 r;
    for (int i = 0; i < 4; i++) r.d[i] = m.c[0].d[i] * v.x + m.c[1].d[i] * v.y + m.c[2].d[i] * v.z + m.c[3].d[i] * v.w;
    return r;
}

// ---- textures and images over host memory
struct sampler2D {
    const float *texels = nullptr;  // width * height RGBA fp32 (an RGBA8 or RGBA16F texture already converted)
    int width = 0, height = 0;
    bool linear = true, repeat_u = false, repeat_v = false;
    vec4 fetch(int x, int y) const {
        x = repeat_u ? ((x % width) + width) % width : (x < 0 ? 0 : (x >= width ? width - 1 : x));
        y = repeat_v ? ((y % height) + height) % height : (y < 0 ? 0 : (y >= height ? height - 1 : y));
        const float *t = texels + ((size_t)y * width + x) * 4;
        return vec4(t[0], t[1], t[2], t[3]);
    }
};
inline ivec2 textureSize(const sampler2D &s, int) { return ivec2(s.width, s.height); }
inline vec4 texelFetch(const sampler2D &s, ivec2 p, int) { return s.fetch(p.x, p.y); }
inline bool &debug_trace() { static thread_local bool on = false; return on; }
inline int &subtexel_bits() { static int bits = 0; return bits; }  // 0: exact fp32 weights (default)
inline vec4 texture(const sampler2D &s, const vec2 &uv) {
    const float x = uv.x * (float)s.width, y = uv.y * (float)s.height;
    if (debug_trace()) fprintf(stderr, "texture %dx%d uv %.9g %.9g -> texel %.6f %.6f\n", s.width, s.height, uv.x, uv.y, x - 0.5f, y - 0.5f);
    if (!s.linear) return s.fetch((int)std::floor(x), (int)std::floor(y));
    const float fx = x - 0.5f, fy = y - 0.5f;
    const float x0 = std::floor(fx), y0 = std::floor(fy);
    float ax = fx - x0, ay = fy - y0;
    // Sub-texel precision: a GPU computes the filter weights in fixed point with 8 fraction bits (Vulkan's
    // subTexelPrecisionBits, GL's equivalent), so a sample within 2^-9 texel of a texel centre returns exactly that texel.
    // The reference's shaders rely on it: they fetch the paint metadata and the mask texture through LINEAR samplers at
    // (i + 0.5) * (1 / size), which in fp32 lands ~1e-6 texel off centre (tile.comp:694-726, :586-607). Everywhere else
    // the weights stay exact fp32.
    const float snap = 1.0f / 512.0f;
    if (ax < snap) ax = 0.0f; else if (ax > 1.0f - snap) ax = 1.0f;
    if (ay < snap) ay = 0.0f; else if (ay > 1.0f - snap) ay = 1.0f;
    if (subtexel_bits() > 0) {  // diagnostic: ALL weights in fixed point, as a GPU's texture unit computes them
        const float q = (float)(1 << subtexel_bits());
        ax = std::nearbyint(ax * q) / q;
        ay = std::nearbyint(ay * q) / q;
    }
    const vec4 t00 = s.fetch((int)x0, (int)y0), t10 = s.fetch((int)x0 + 1, (int)y0);
    const vec4 t01 = s.fetch((int)x0, (int)y0 + 1), t11 = s.fetch((int)x0 + 1, (int)y0 + 1);
    return mix(mix(t00, t10, ax), mix(t01, t11, ax), ay);
}

struct image2D {  // rgba8
    uint8_t *texels = nullptr;
    int width = 0, height = 0;
};
inline vec4 imageLoad(const image2D &im, ivec2 p) {
    if (p.x < 0 || p.y < 0 || p.x >= im.width || p.y >= im.height) return vec4(0.0f);
    const uint8_t *t = im.texels + ((size_t)p.y * im.width + p.x) * 4;
    return vec4(t[0] / 255.0f, t[1] / 255.0f, t[2] / 255.0f, t[3] / 255.0f);
}
inline void imageStore(image2D &im, ivec2 p, const vec4 &v) {
    if (p.x < 0 || p.y < 0 || p.x >= im.width || p.y >= im.height) return;
    uint8_t *t = im.texels + ((size_t)p.y * im.width + p.x) * 4;
    for (int i = 0; i < 4; i++) t[i] = (uint8_t)std::nearbyint(clamp(v.d[i], 0.0f, 1.0f) * 255.0f);
}

// GLSL lets gl_LocalInvocationID.xy be a uvec2 expression: the generator rewrites `.xy` on the built-ins to `.xy_()`.
inline void barrier() {}
inline void memoryBarrierShared() {}

}  // namespace glsl
