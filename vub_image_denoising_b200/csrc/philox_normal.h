/*
 * philox_normal.h — the noise generator SPEC of this repo, shared verbatim by the CUDA kernel
 * (csrc/noise.cu) and by the CPU oracle (oracle/noise_oracle.c), so that both sides execute the same
 * sequence of correctly-rounded IEEE-754 binary32 operations and agree bit for bit.
 *
 * The reference synthesises noise with numpy's global, never-seeded MT19937
 * (dataset_creation/custom_dataset.py:83-87), so there is no reference bit stream to match; this
 * header defines ours:
 *
 *   Philox4x32-10 (Salmon et al., SC'11), key = (seed_lo, seed_hi),
 *   counter = (q_lo, q_hi, stream_id, 0) for element quad q = floor(elem_index / 4);
 *   the four 32-bit outputs r0..r3 give two Box-Muller pairs:
 *       (z[4q+0], z[4q+1]) from (r0, r1),   (z[4q+2], z[4q+3]) from (r2, r3)
 *   u1 = ((r >> 9) + 0.5) * 2^-23 in (0,1),  u2 likewise;
 *   rad = sqrt(-2 ln u1),  z_even = rad * cos(2 pi u2),  z_odd = rad * sin(2 pi u2).
 *
 * ln / sin / cos are fixed polynomial evaluations written only with fmaf / + / * / sqrtf, every one
 * of which is correctly rounded on both the host (compile with -ffp-contract=off) and the device
 * (explicit __f*_rn intrinsics), hence reproducible everywhere.
 *
 * Only macros B2N_FMA/B2N_MUL/B2N_ADD/B2N_SUB/B2N_SQRT and B2N_FN differ between the two builds.
 */
#ifndef B200DN_PHILOX_NORMAL_H_
#define B200DN_PHILOX_NORMAL_H_

#include <stdint.h>

#ifdef __CUDA_ARCH__
#define B2N_FN __device__ __forceinline__
#define B2N_FMA(a, b, c) __fmaf_rn((a), (b), (c))
#define B2N_MUL(a, b) __fmul_rn((a), (b))
#define B2N_ADD(a, b) __fadd_rn((a), (b))
#define B2N_SUB(a, b) __fsub_rn((a), (b))
#define B2N_SQRT(a) __fsqrt_rn((a))
#define B2N_MULHI(a, b) __umulhi((a), (b))
#define B2N_F2U(f) __float_as_uint((f))
#define B2N_U2F(u) __uint_as_float((u))
#else
#include <math.h>
#include <string.h>
#define B2N_FN static inline
#define B2N_FMA(a, b, c) fmaf((a), (b), (c))
#define B2N_MUL(a, b) ((float)((a) * (b)))
#define B2N_ADD(a, b) ((float)((a) + (b)))
#define B2N_SUB(a, b) ((float)((a) - (b)))
#define B2N_SQRT(a) sqrtf((a))
#define B2N_MULHI(a, b) ((uint32_t)(((uint64_t)(a) * (uint64_t)(b)) >> 32))
static inline uint32_t b2n_f2u(float f) { uint32_t u; memcpy(&u, &f, 4); return u; }
static inline float b2n_u2f(uint32_t u) { float f; memcpy(&f, &u, 4); return f; }
#define B2N_F2U(f) b2n_f2u((f))
#define B2N_U2F(u) b2n_u2f((u))
#endif

/* Philox4x32-10. ctr/out are 4 words, key 2 words. */
B2N_FN void b2n_philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]) {
  uint32_t c0 = ctr[0], c1 = ctr[1], c2 = ctr[2], c3 = ctr[3];
  uint32_t k0 = key[0], k1 = key[1];
  for (int r = 0; r < 10; ++r) {
    const uint32_t hi0 = B2N_MULHI(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
    const uint32_t hi1 = B2N_MULHI(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
    const uint32_t n0 = hi1 ^ c1 ^ k0;
    const uint32_t n2 = hi0 ^ c3 ^ k1;
    c0 = n0;
    c1 = lo1;
    c2 = n2;
    c3 = lo0;
    k0 += 0x9E3779B9u;
    k1 += 0xBB67AE85u;
  }
  out[0] = c0;
  out[1] = c1;
  out[2] = c2;
  out[3] = c3;
}

/* uniform in (0,1), exactly representable: ((r >> 9) + 0.5) * 2^-23 */
B2N_FN float b2n_uniform(uint32_t r) {
  const float a = (float)(r >> 9);                 /* < 2^23, exact */
  return B2N_MUL(B2N_ADD(a, 0.5f), 1.1920928955078125e-07f); /* exact: 24 significant bits */
}

/* natural log for x in (0,1], Cephes-style: x = m * 2^e, m in [sqrt(1/2), sqrt(2)) */
B2N_FN float b2n_logf(float x) {
  uint32_t ix = B2N_F2U(x);
  int e = (int)(ix >> 23) - 127;
  ix = (ix & 0x007fffffu) | 0x3f800000u; /* m in [1,2) */
  float m = B2N_U2F(ix);
  if (m > 1.41421356237f) {
    m = B2N_MUL(m, 0.5f);
    e += 1;
  }
  const float f = B2N_SUB(m, 1.0f);
  const float z = B2N_MUL(f, f);
  float p = 7.0376836292e-2f;
  p = B2N_FMA(p, f, -1.1514610310e-1f);
  p = B2N_FMA(p, f, 1.1676998740e-1f);
  p = B2N_FMA(p, f, -1.2420140846e-1f);
  p = B2N_FMA(p, f, 1.4249322787e-1f);
  p = B2N_FMA(p, f, -1.6668057665e-1f);
  p = B2N_FMA(p, f, 2.0000714765e-1f);
  p = B2N_FMA(p, f, -2.4999993993e-1f);
  p = B2N_FMA(p, f, 3.3333331174e-1f);
  float y = B2N_MUL(B2N_MUL(p, f), z); /* f^3 * P(f) */
  const float fe = (float)e;
  y = B2N_FMA(fe, -2.12194440e-4f, y);
  y = B2N_FMA(-0.5f, z, y);
  float r = B2N_ADD(f, y);
  r = B2N_FMA(fe, 0.693359375f, r);
  return r;
}

/* sin and cos of 2*pi*u for u in (0,1) (u a multiple of 2^-24): exact octant reduction, then
   Cephes single-precision kernels on [-pi/4, pi/4]. */
B2N_FN void b2n_sincos2pi(float u, float* s_out, float* c_out) {
  const float x4 = B2N_MUL(u, 4.0f);          /* exact; angle = x4 * pi/2 */
  const float kf = (float)(int)B2N_ADD(x4, 0.5f); /* nearest quadrant 0..4 (x4 + .5 exact: < 2^3 with 2^-22 grid) */
  const int k = (int)kf;
  const float r = B2N_SUB(x4, kf);            /* exact, in [-0.5, 0.5] */
  const float t = B2N_MUL(r, 1.57079632679489661923f); /* one rounding */
  const float z = B2N_MUL(t, t);
  /* sin(t) */
  float ps = -1.9515295891e-4f;
  ps = B2N_FMA(ps, z, 8.3321608736e-3f);
  ps = B2N_FMA(ps, z, -1.6666654611e-1f);
  const float sn = B2N_FMA(B2N_MUL(ps, z), t, t);
  /* cos(t) */
  float pc = 2.443315711809948e-5f;
  pc = B2N_FMA(pc, z, -1.388731625493765e-3f);
  pc = B2N_FMA(pc, z, 4.166664568298827e-2f);
  float cs = B2N_MUL(B2N_MUL(pc, z), z);
  cs = B2N_FMA(-0.5f, z, cs);
  cs = B2N_ADD(cs, 1.0f);
  float s, c;
  switch (k & 3) {
    case 0: s = sn; c = cs; break;
    case 1: s = cs; c = -sn; break;
    case 2: s = -sn; c = -cs; break;
    default: s = -cs; c = sn; break;
  }
  *s_out = s;
  *c_out = c;
}

/* four standard normals for element quad q */
B2N_FN void b2n_normal4(uint64_t q, uint64_t seed, uint32_t stream_id, float z[4]) {
  uint32_t ctr[4], key[2], r[4];
  ctr[0] = (uint32_t)q;
  ctr[1] = (uint32_t)(q >> 32);
  ctr[2] = stream_id;
  ctr[3] = 0u;
  key[0] = (uint32_t)seed;
  key[1] = (uint32_t)(seed >> 32);
  b2n_philox4x32_10(ctr, key, r);
  for (int h = 0; h < 2; ++h) {
    const float u1 = b2n_uniform(r[2 * h]);
    const float u2 = b2n_uniform(r[2 * h + 1]);
    const float rad = B2N_SQRT(B2N_MUL(-2.0f, b2n_logf(u1)));
    float s, c;
    b2n_sincos2pi(u2, &s, &c);
    z[2 * h] = B2N_MUL(rad, c);
    z[2 * h + 1] = B2N_MUL(rad, s);
  }
}

/* the reference's degradation + normalisation pipeline for one sample
   (dataset_creation/custom_dataset.py:84-86, dataset_creation/data_loader.py:35-38):
   n = float32(u8) + sigma*z ; clip to [0,255] ; truncate to uint8 ; /255 ; (v-0.5)/0.5           */
B2N_FN uint8_t b2n_degrade_u8(uint8_t clean, float sigma, float z) {
  float n = B2N_ADD((float)clean, B2N_MUL(sigma, z));
  n = n < 0.0f ? 0.0f : (n > 255.0f ? 255.0f : n);
  return (uint8_t)(int)n;
}

#endif /* B200DN_PHILOX_NORMAL_H_ */
