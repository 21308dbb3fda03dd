// b200sim: common definitions for the warp-per-environment kernels.
//
// Execution model: ONE WARP = ONE ENVIRONMENT.  Lanes map to dofs / bodies / constraint rows; all per-env
// working data lives in shared memory (Work<D>), component-major so that lane-strided accesses are
// bank-conflict free.  Code is written as a sequence of lane-parallel phases:
//
//     LANE_FOR(i, n) { ... independent work on item i ... }   WSYNC();
//
// Rule: inside one phase no lane reads a location another lane writes in the same phase.  The same source
// compiles for the host with B2S_EMU defined, where a phase degenerates to a plain loop; tests use that
// build only to debug the warp code on GPU-less machines (tests/emu) -- it is never part of the product.
#pragma once

#include <math.h>
#include <stdint.h>

#include "../../include/b200sim_codes.h"
#include "../../include/b200sim_model.h"

#if defined(__CUDACC__) && !defined(B2S_EMU)
#define B2S_DEVICE 1
#define DEV __device__ __forceinline__
#define DEVNI static __device__ __noinline__
#define LANE_FOR(i, n) for (int i = lane; i < (n); i += 32)
#define WSYNC() __syncwarp()
// Stage barriers keep a CTA's warps on the same code (shared instruction-cache fills); they carry no data dependence.
// B2S_BAR_GROUP = g > 0 (tuning experiments) synchronises groups of g consecutive warps on named barriers instead.
#if defined(B2S_BAR_GROUP) && B2S_BAR_GROUP > 0
#define CTA_SYNC() asm volatile("bar.sync %0, %1;" ::"r"((int)(threadIdx.x >> 5) / B2S_BAR_GROUP + 1), "r"(B2S_BAR_GROUP * 32) : "memory")
#else
#define CTA_SYNC() __syncthreads()
#endif
#define CTA_OR(p) __syncthreads_or(p)
#define IF_LANE0 if (lane == 0)
#define LANE_ARG const int lane
#define LANE_PASS lane
#else
#define B2S_DEVICE 0
#define DEV static inline
#define DEVNI static
#define LANE_FOR(i, n) for (int i = 0; i < (n); i++)
#define WSYNC() ((void)0)
#define CTA_SYNC() ((void)0)
#define CTA_OR(p) (p)
#define IF_LANE0 if (true)
#define LANE_ARG const int lane
#define LANE_PASS 0
#endif

// Optional per-stage cycle accounting (-DB2S_PROFILE_STAGES; tools/stage_profile.py): lane 0 of every warp adds the
// clock64() delta of each stage to Work::prof, the kernel sums them into g_stage_cycles.  Off in the product build.
// -DB2S_DEBUG_STOP (fault bisection; compute-sanitizer is closed on the pool): threads exit at stop point k when bits
// 16..23 of B2S_SYNC_MASK equal k.  Run with B2S_WPB=1 so that no CTA barrier waits for the exited warp.
#if defined(B2S_DEBUG_STOP) && B2S_DEVICE
#define DBG_STOP(m, k) do { if ((((m).sync_mask >> 16) & 0xff) == (k)) asm volatile("exit;"); } while (0)
#else
#define DBG_STOP(m, k) do { } while (0)
#endif
#if defined(B2S_PROFILE_STAGES) && B2S_DEVICE
#define B2S_NSTAGE 32
#define STAGE_BEGIN() long long stage_t0_ = clock64()
#define STAGE_MARK(w, k) do { const long long t1_ = clock64(); if (lane == 0) (w).prof[k] += (unsigned)(t1_ - stage_t0_); stage_t0_ = t1_; } while (0)
// nested accounting inside a stage (its own clock): adds to slot k without disturbing the enclosing stage's clock
#define SUB_BEGIN() long long sub_t0_ = clock64()
#define SUB_MARK(w, k) do { const long long t1_ = clock64(); if (lane == 0) (w).prof[k] += (unsigned)(t1_ - sub_t0_); sub_t0_ = t1_; } while (0)
#define PROF_COUNT(w, k) do { if (lane == 0) (w).prof[k] += 1u; } while (0)
#else
#define STAGE_BEGIN() ((void)0)
#define STAGE_MARK(w, k) ((void)0)
#define SUB_BEGIN() ((void)0)
#define SUB_MARK(w, k) ((void)0)
#define PROF_COUNT(w, k) ((void)0)
#endif

namespace b2s {

constexpr float MINVAL = 1e-15f;
constexpr float MINIMP = 0.0001f;
constexpr float MAXIMP = 0.9999f;
constexpr float PI_F = 3.14159265358979323846f;

enum { JNT_FREE = 0, JNT_HINGE = 3 };
enum { GEOM_PLANE = 0, GEOM_SPHERE = 2, GEOM_CAPSULE = 3 };
enum {
  SENS_TOUCH = 0, SENS_ACCEL = 1, SENS_VELOCIMETER = 2, SENS_GYRO = 3, SENS_FORCE = 4, SENS_TORQUE = 5, SENS_MAGNETOMETER = 6,
  SENS_FRAMEPOS = 10, SENS_FRAMEQUAT = 11, SENS_FRAMEXAXIS = 12, SENS_FRAMEYAXIS = 13, SENS_FRAMEZAXIS = 14,
  SENS_FRAMELINVEL = 15, SENS_FRAMEANGVEL = 16, SENS_FRAMELINACC = 17, SENS_FRAMEANGACC = 18
};
enum { OBJ_BODY = 1, OBJ_XBODY = 2, OBJ_SITE = 6 };

// ---- warp reductions -------------------------------------------------------------------------------
#if B2S_DEVICE
DEV float wsum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
DEV float wmax(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
DEV int wany(int p) { return __any_sync(0xffffffffu, p); }
DEV int wsum_i(int v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
#else
DEV float wsum(float v) { return v; }
DEV float wmax(float v) { return v; }
DEV int wany(int p) { return p; }
DEV int wsum_i(int v) { return v; }
#endif

// ---- model view --------------------------------------------------------------------------------------
struct Mdl {
  int h[MH_SIZE];
  const int* mi;
  const float* mf;
  const int* ti;
  const float* tf;
  int sync_mask;  // which per-sub-step CTA barriers run (bit 0..5: forward's stages, bit 6: sub-step start)
  int stage[4];   // words of mi / mf / ti / tf the kernels copy into shared memory (0: read from global memory)
#ifdef B2S_POISON_SMEM
  int poison_lo, poison_hi;  // debug build: bytes [lo, hi) of every workspace start as 0xFF (tools/debug_poison.py)
#endif
};
#define MI(m, name) ((m).mi + (m).h[name])
#define MF(m, name) ((m).mf + (m).h[name])

// ---- small vector math -------------------------------------------------------------------------------
struct V3 { float x, y, z; };
struct Q4 { float w, x, y, z; };
struct M9 { float m[9]; };
struct V6 { float v[6]; };

DEV V3 v3(float x, float y, float z) { V3 r; r.x = x; r.y = y; r.z = z; return r; }
DEV V3 operator+(V3 a, V3 b) { return v3(a.x + b.x, a.y + b.y, a.z + b.z); }
DEV V3 operator-(V3 a, V3 b) { return v3(a.x - b.x, a.y - b.y, a.z - b.z); }
DEV V3 operator*(V3 a, float s) { return v3(a.x * s, a.y * s, a.z * s); }
DEV float dot(V3 a, V3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
DEV V3 cross(V3 a, V3 b) { return v3(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x); }

DEV Q4 qmul(Q4 a, Q4 b) {
  Q4 r;
  r.w = a.w * b.w - a.x * b.x - a.y * b.y - a.z * b.z;
  r.x = a.w * b.x + a.x * b.w + a.y * b.z - a.z * b.y;
  r.y = a.w * b.y - a.x * b.z + a.y * b.w + a.z * b.x;
  r.z = a.w * b.z + a.x * b.y - a.y * b.x + a.z * b.w;
  return r;
}
DEV Q4 qnormalize(Q4 q) {
  float n = sqrtf(q.w * q.w + q.x * q.x + q.y * q.y + q.z * q.z);
  if (n < MINVAL) { q.w = 1; q.x = q.y = q.z = 0; return q; }
  float inv = 1.0f / n;
  q.w *= inv; q.x *= inv; q.y *= inv; q.z *= inv;
  return q;
}
DEV M9 q2mat(Q4 q) {
  M9 r;
  float w = q.w, x = q.x, y = q.y, z = q.z;
  r.m[0] = w * w + x * x - y * y - z * z; r.m[1] = 2 * (x * y - w * z); r.m[2] = 2 * (x * z + w * y);
  r.m[3] = 2 * (x * y + w * z); r.m[4] = w * w - x * x + y * y - z * z; r.m[5] = 2 * (y * z - w * x);
  r.m[6] = 2 * (x * z - w * y); r.m[7] = 2 * (y * z + w * x); r.m[8] = w * w - x * x - y * y + z * z;
  return r;
}
DEV V3 mulv(const M9& m, V3 v) {
  return v3(m.m[0] * v.x + m.m[1] * v.y + m.m[2] * v.z, m.m[3] * v.x + m.m[4] * v.y + m.m[5] * v.z,
            m.m[6] * v.x + m.m[7] * v.y + m.m[8] * v.z);
}
DEV V3 mulTv(const M9& m, V3 v) {
  return v3(m.m[0] * v.x + m.m[3] * v.y + m.m[6] * v.z, m.m[1] * v.x + m.m[4] * v.y + m.m[7] * v.z,
            m.m[2] * v.x + m.m[5] * v.y + m.m[8] * v.z);
}
DEV V3 qrot(Q4 q, V3 v) { return mulv(q2mat(q), v); }
DEVNI Q4 axis_angle(V3 axis, float angle) {
  float s, c;
#if B2S_DEVICE
  sincosf(angle * 0.5f, &s, &c);  // one shared range reduction
#else
  s = sinf(angle * 0.5f); c = cosf(angle * 0.5f);
#endif
  Q4 q; q.w = c; q.x = axis.x * s; q.y = axis.y * s; q.z = axis.z * s;
  return q;
}
// xax.rotate_vector_by_quat (normalises with eps 1e-6); same expression order as the oracle (oracle/physics.c)
DEVNI V3 rot_vec_quat(V3 v, Q4 q, bool inverse) {
  float n = sqrtf(q.w * q.w + q.x * q.x + q.y * q.y + q.z * q.z) + 1e-6f;
  float w = q.w / n, x = q.x / n, y = q.y / n, z = q.z / n;
  if (inverse) { x = -x; y = -y; z = -z; }
  float vx = v.x, vy = v.y, vz = v.z;
  float xx = w * w * vx + 2 * y * w * vz - 2 * z * w * vy + x * x * vx + 2 * y * x * vy + 2 * z * x * vz - z * z * vx -
             y * y * vx;
  float yy = 2 * x * y * vx + y * y * vy + 2 * z * y * vz + 2 * w * z * vx - z * z * vy + w * w * vy - 2 * w * x * vz -
             x * x * vy;
  float zz = 2 * x * z * vx + 2 * y * z * vy + z * z * vz - 2 * w * y * vx + w * w * vz + 2 * w * x * vy - y * y * vz -
             x * x * vz;
  return v3(xx, yy, zz);
}
DEVNI Q4 euler_to_quat(float roll, float pitch, float yaw) {
  float cr = cosf(roll * 0.5f), sr = sinf(roll * 0.5f);
  float cp = cosf(pitch * 0.5f), sp = sinf(pitch * 0.5f);
  float cy = cosf(yaw * 0.5f), sy = sinf(yaw * 0.5f);
  Q4 q;
  q.w = cr * cp * cy + sr * sp * sy;
  q.x = sr * cp * cy - cr * sp * sy;
  q.y = cr * sp * cy + sr * cp * sy;
  q.z = cr * cp * sy - sr * sp * cy;
  return q;
}

// xax.quat_to_euler(wxyz) -> (roll, pitch, yaw), normalising with eps 1e-6; same expressions as the oracle (oracle/engine.c)
DEVNI V3 quat_to_euler(Q4 q) {
  const float n = sqrtf(q.w * q.w + q.x * q.x + q.y * q.y + q.z * q.z) + 1e-6f;
  const float w = q.w / n, x = q.x / n, y = q.y / n, z = q.z / n;
  const float roll = atan2f(2.0f * (w * x + y * z), 1.0f - 2.0f * (x * x + y * y));
  const float sinp = 2.0f * (w * y - z * x);
  const float pitch = fabsf(sinp) >= 1.0f ? (sinp > 0 ? 1.57079632679489661923f : -1.57079632679489661923f) : asinf(sinp);
  const float yaw = atan2f(2.0f * (w * z + x * y), 1.0f - 2.0f * (y * y + z * z));
  return v3(roll, pitch, yaw);
}
DEV float qdot(Q4 a, Q4 b) { return a.w * b.w + a.x * b.x + a.y * b.y + a.z * b.z; }

// spatial 6-vectors are [rotational(3), translational(3)]; inertia is the 10-vector MuJoCo uses for cinert
DEV V6 mul_inert_vec(const float* i, const V6& v) {
  V6 r;
  r.v[0] = i[0] * v.v[0] + i[3] * v.v[1] + i[4] * v.v[2] - i[8] * v.v[4] + i[7] * v.v[5];
  r.v[1] = i[3] * v.v[0] + i[1] * v.v[1] + i[5] * v.v[2] + i[8] * v.v[3] - i[6] * v.v[5];
  r.v[2] = i[4] * v.v[0] + i[5] * v.v[1] + i[2] * v.v[2] - i[7] * v.v[3] + i[6] * v.v[4];
  r.v[3] = i[8] * v.v[1] - i[7] * v.v[2] + i[9] * v.v[3];
  r.v[4] = i[6] * v.v[2] - i[8] * v.v[0] + i[9] * v.v[4];
  r.v[5] = i[7] * v.v[0] - i[6] * v.v[1] + i[9] * v.v[5];
  return r;
}
DEV V6 cross_motion(const V6& vel, const V6& v) {
  V3 a = cross(v3(vel.v[0], vel.v[1], vel.v[2]), v3(v.v[0], v.v[1], v.v[2]));
  V3 b = cross(v3(vel.v[0], vel.v[1], vel.v[2]), v3(v.v[3], v.v[4], v.v[5]));
  V3 c = cross(v3(vel.v[3], vel.v[4], vel.v[5]), v3(v.v[0], v.v[1], v.v[2]));
  V6 r;
  r.v[0] = a.x; r.v[1] = a.y; r.v[2] = a.z; r.v[3] = b.x + c.x; r.v[4] = b.y + c.y; r.v[5] = b.z + c.z;
  return r;
}
DEV V6 cross_force(const V6& vel, const V6& f) {
  V3 a = cross(v3(vel.v[0], vel.v[1], vel.v[2]), v3(f.v[0], f.v[1], f.v[2]));
  V3 b = cross(v3(vel.v[3], vel.v[4], vel.v[5]), v3(f.v[3], f.v[4], f.v[5]));
  V3 c = cross(v3(vel.v[0], vel.v[1], vel.v[2]), v3(f.v[3], f.v[4], f.v[5]));
  V6 r;
  r.v[0] = a.x + b.x; r.v[1] = a.y + b.y; r.v[2] = a.z + b.z; r.v[3] = c.x; r.v[4] = c.y; r.v[5] = c.z;
  return r;
}

// ---- threefry2x32 / jax.random restatement (same as oracle/engine.c and ksim_b200/rng.py) --------------
struct Key { uint32_t k0, k1; };

DEV uint32_t rotl32(uint32_t x, int r) { return (x << r) | (x >> (32 - r)); }

struct U2 { uint32_t a, b; };

// returns by value: out-pointers into a noinline function would go through local memory
DEVNI U2 threefry2x32(uint32_t k0, uint32_t k1, uint32_t x0, uint32_t x1) {
  const uint32_t ks0 = k0, ks1 = k1, ks2 = k0 ^ k1 ^ 0x1BD11BDAu;
  x0 += ks0; x1 += ks1;
#define B2S_R4A x0 += x1; x1 = rotl32(x1, 13); x1 ^= x0; x0 += x1; x1 = rotl32(x1, 15); x1 ^= x0; \
                x0 += x1; x1 = rotl32(x1, 26); x1 ^= x0; x0 += x1; x1 = rotl32(x1, 6); x1 ^= x0;
#define B2S_R4B x0 += x1; x1 = rotl32(x1, 17); x1 ^= x0; x0 += x1; x1 = rotl32(x1, 29); x1 ^= x0; \
                x0 += x1; x1 = rotl32(x1, 16); x1 ^= x0; x0 += x1; x1 = rotl32(x1, 24); x1 ^= x0;
  B2S_R4A x0 += ks1; x1 += ks2 + 1u;
  B2S_R4B x0 += ks2; x1 += ks0 + 2u;
  B2S_R4A x0 += ks0; x1 += ks1 + 3u;
  B2S_R4B x0 += ks1; x1 += ks2 + 4u;
  B2S_R4A x0 += ks2; x1 += ks0 + 5u;
#undef B2S_R4A
#undef B2S_R4B
  U2 r; r.a = x0; r.b = x1;
  return r;
}
DEV Key key_split(Key k, int i) {
  const U2 r = threefry2x32(k.k0, k.k1, 0u, (uint32_t)i);
  Key o; o.k0 = r.a; o.k1 = r.b;
  return o;
}
// All children of a key that a warp needs, from ONE threefry evaluation: lane l computes split(k, l) and a shuffle
// hands child i to every lane (jax.random.split(k, n)[i] is threefry(k, (0, i)) for every n).  Must be called by the
// whole warp.  The host build evaluates the requested child directly.
struct KeyFan { Key parent; U2 mine; };
DEV KeyFan key_fan(Key k, LANE_ARG) {
  KeyFan f;
  f.parent = k;
#if B2S_DEVICE
  f.mine = threefry2x32(k.k0, k.k1, 0u, (uint32_t)lane);
#else
  (void)lane;
  f.mine.a = f.mine.b = 0u;
#endif
  return f;
}
DEV Key key_child(const KeyFan& f, int i) {
#if B2S_DEVICE
  Key o;
  o.k0 = __shfl_sync(0xffffffffu, f.mine.a, i);
  o.k1 = __shfl_sync(0xffffffffu, f.mine.b, i);
  return o;
#else
  return key_split(f.parent, i);
#endif
}
DEV uint32_t key_bits(Key k, int i) {
  const U2 r = threefry2x32(k.k0, k.k1, 0u, (uint32_t)i);
  return r.a ^ r.b;
}
DEV float bits_to_unit(uint32_t bits) {
  union { uint32_t u; float f; } c;
  c.u = (bits >> 9) | 0x3F800000u;
  return c.f - 1.0f;
}
// the two-step form (mul, add -- no fma) keeps the draw bit-identical to XLA's and the oracle's
DEV float key_uniform(Key k, int i, float lo, float hi) {
  float u = bits_to_unit(key_bits(k, i));
#if B2S_DEVICE
  float v = __fadd_rn(__fmul_rn(u, __fsub_rn(hi, lo)), lo);
#else
  volatile float t = u * (hi - lo);
  float v = t + lo;
#endif
  return v > lo ? v : lo;
}
DEV float erfinv_f32poly(float x) {
#if B2S_DEVICE
  float w = -logf(__fmul_rn(__fsub_rn(1.0f, x), __fadd_rn(1.0f, x))), p;
#define B2S_PA(c) p = __fadd_rn((c), __fmul_rn(p, w))
#else
  volatile float t0 = (1.0f - x) * (1.0f + x);
  float w = -logf(t0), p;
#define B2S_PA(c) { volatile float t_ = p * w; p = (c) + t_; }
#endif
  if (w < 5.0f) {
    w = w - 2.5f;
    p = 2.81022636e-08f;
    B2S_PA(3.43273939e-07f); B2S_PA(-3.5233877e-06f); B2S_PA(-4.39150654e-06f); B2S_PA(0.00021858087f);
    B2S_PA(-0.00125372503f); B2S_PA(-0.00417768164f); B2S_PA(0.246640727f); B2S_PA(1.50140941f);
  } else {
    w = sqrtf(w) - 3.0f;
    p = -0.000200214257f;
    B2S_PA(0.000100950558f); B2S_PA(0.00134934322f); B2S_PA(-0.00367342844f); B2S_PA(0.00573950773f);
    B2S_PA(-0.0076224613f); B2S_PA(0.00943887047f); B2S_PA(1.00167406f); B2S_PA(2.83297682f);
  }
#undef B2S_PA
  return p * x;
}
DEV float key_normal(Key k, int i) {
  float u = key_uniform(k, i, -0.99999994f, 1.0f);
  return 1.41421356237309504880f * erfinv_f32poly(u);
}
DEV int key_bernoulli(Key k, int i, float p) { return bits_to_unit(key_bits(k, i)) < p; }

// ---- scales / random variables / noise ------------------------------------------------------------------
DEV float scale_get(int kind, const float* f, float level) {
  switch (kind) {
    case SCALE_CONST: return f[0];
    case SCALE_LINEAR: return f[0] * level + f[1];
    case SCALE_QUADRATIC: return f[0] * level * level + f[1];
    case SCALE_SQRT: return f[0] * sqrtf(level) + f[1];
    case SCALE_NONZERO: return level > 0 ? f[0] : f[1];
  }
  return f[0];
}
DEVNI float rv_sample(int kind, const float* f, Key key, int i, float level) {
  float mean = f[0], std = f[1], mag = f[2];
  switch (kind) {
    case RV_UNIFORM: return key_uniform(key, i, mean - mag * level, mean + mag * level);
    case RV_GAUSSIAN: return key_normal(key, i) * (std * level) + mean;
    case RV_UNIFORM_GAUSSIAN: {
      float n = key_normal(key, i) * (std * level);
      float u = key_uniform(key, i, -mag * level, mag * level);
      return n + u + mean;
    }
  }
  return 0.0f;
}
DEVNI float noise_apply(const int* ni, const float* nf, float x, Key key, int i, float level) {
  int n = ni[0];
  for (int l = 0; l < n; l++) {
    int mul = ni[1 + 2 * l], kind = ni[2 + 2 * l];
    float r = rv_sample(kind, nf + 3 * l, key, i, level);
    x = mul ? x * r : x + r;
  }
  return x;
}
// 1/sqrt(x) for x >= MINVAL: hardware approximation + one Newton step on the device (no slow-path call)
DEV float inv_sqrt(float x) {
#if B2S_DEVICE
  float r = rsqrtf(x);
  return r * fmaf(-0.5f * x, r * r, 1.5f);
#else
  return 1.0f / sqrtf(x);
#endif
}
// a / b without the IEEE slow path on the device (MUFU reciprocal, ~2 ulp; |b| must stay below 2^126)
DEV float fast_div(float a, float b) {
#if B2S_DEVICE
  return __fdividef(a, b);
#else
  return a / b;
#endif
}
DEV float clipf(float x, float lo, float hi) { return x < lo ? lo : (x > hi ? hi : x); }

}  // namespace b2s
