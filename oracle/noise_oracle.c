/*
 * ORACLE (test infrastructure, not product): CPU restatement of the repo's Gaussian-noise synthesis
 * spec and of the reference's degradation / normalisation arithmetic.
 *
 * Reference arithmetic restated here:
 *   noisy = clip(float32(patch) + noise, 0, 255).astype(uint8)      dataset_creation/custom_dataset.py:84-86
 *   ToTensor (/255) and Normalize(0.5, 0.5)                        dataset_creation/data_loader.py:35-38
 *   (y+1)/2 -> clip(y*255, 0, 255).astype(uint8)                   evaluate_SIDD/benchmark.py:42-44
 *
 * The reference's noise comes from numpy's global MT19937, never seeded (custom_dataset.py:85): there is
 * no reference bit stream.  The generator SPEC is this repo's own ("Noise synthesis"):
 *   Philox4x32-10, key = seed (lo, hi), counter = (q lo, q hi, stream_id, 0), q = element_index / 4;
 *   outputs r0..r3 -> Box-Muller pairs (r0,r1) -> z[4q], z[4q+1] and (r2,r3) -> z[4q+2], z[4q+3];
 *   u = ((r >> 9) + 0.5) * 2^-23;  rad = sqrt(-2 ln u1);  z_even = rad cos(2 pi u2), z_odd = rad sin(2 pi u2);
 *   ln / sin / cos = the fixed Cephes single-precision polynomials below, evaluated with fused
 *   multiply-adds exactly where the spec says so.  Every operation is a correctly rounded binary32 op,
 *   so a conforming implementation reproduces these outputs bit for bit.
 * Philox itself is pinned by the Random123 known-answer vectors (tests/test_noise_oracle.py).
 *
 * Build: gcc -O2 -ffp-contract=off -fPIC -shared (oracle/build.py).  -ffp-contract=off matters: the
 * compiler must not fuse a*b+c on its own; fusion happens only through the explicit fmaf() calls.
 */
#include <math.h>
#include <stdint.h>
#include <string.h>

static const uint32_t PHILOX_M0 = 0xD2511F53u, PHILOX_M1 = 0xCD9E8D57u;
static const uint32_t PHILOX_W0 = 0x9E3779B9u, PHILOX_W1 = 0xBB67AE85u;

void b200dn_oracle_philox4x32_10(const uint32_t counter[4], const uint32_t key[2], uint32_t out[4]) {
  uint32_t x[4] = {counter[0], counter[1], counter[2], counter[3]};
  uint32_t k[2] = {key[0], key[1]};
  for (int round = 0; round < 10; ++round) {
    if (round > 0) {
      k[0] += PHILOX_W0;
      k[1] += PHILOX_W1;
    }
    const uint64_t p0 = (uint64_t)PHILOX_M0 * x[0];
    const uint64_t p1 = (uint64_t)PHILOX_M1 * x[2];
    const uint32_t y0 = (uint32_t)(p1 >> 32) ^ x[1] ^ k[0];
    const uint32_t y1 = (uint32_t)p1;
    const uint32_t y2 = (uint32_t)(p0 >> 32) ^ x[3] ^ k[1];
    const uint32_t y3 = (uint32_t)p0;
    x[0] = y0; x[1] = y1; x[2] = y2; x[3] = y3;
  }
  memcpy(out, x, sizeof(x));
}

static float bits_to_float(uint32_t u) { float f; memcpy(&f, &u, 4); return f; }
static uint32_t float_to_bits(float f) { uint32_t u; memcpy(&u, &f, 4); return u; }

static float unit_open(uint32_t r) { return ((float)(r >> 9) + 0.5f) * 1.1920928955078125e-07f; }

static const float LOG_P[9] = {7.0376836292e-2f, -1.1514610310e-1f, 1.1676998740e-1f, -1.2420140846e-1f,
                               1.4249322787e-1f, -1.6668057665e-1f, 2.0000714765e-1f, -2.4999993993e-1f,
                               3.3333331174e-1f};

static float spec_log(float x) {
  uint32_t bits = float_to_bits(x);
  int expo = (int)(bits >> 23) - 127;
  float mant = bits_to_float((bits & 0x007fffffu) | 0x3f800000u);
  if (mant > 1.41421356237f) { mant = mant * 0.5f; expo += 1; }
  const float f = mant - 1.0f;
  const float f2 = f * f;
  float poly = LOG_P[0];
  for (int i = 1; i < 9; ++i) poly = fmaf(poly, f, LOG_P[i]);
  float y = (poly * f) * f2;
  const float fe = (float)expo;
  y = fmaf(fe, -2.12194440e-4f, y);
  y = fmaf(-0.5f, f2, y);
  float r = f + y;
  return fmaf(fe, 0.693359375f, r);
}

static void spec_sincos_2pi(float u, float* s, float* c) {
  const float quarter_turns = u * 4.0f;
  const int k = (int)(quarter_turns + 0.5f);
  const float frac = quarter_turns - (float)k;
  const float t = frac * 1.57079632679489661923f;
  const float t2 = t * t;
  float ps = fmaf(-1.9515295891e-4f, t2, 8.3321608736e-3f);
  ps = fmaf(ps, t2, -1.6666654611e-1f);
  const float sn = fmaf(ps * t2, t, t);
  float pc = fmaf(2.443315711809948e-5f, t2, -1.388731625493765e-3f);
  pc = fmaf(pc, t2, 4.166664568298827e-2f);
  float cs = (pc * t2) * t2;
  cs = fmaf(-0.5f, t2, cs);
  cs = cs + 1.0f;
  switch (k & 3) {
    case 0: *s = sn;  *c = cs;  break;
    case 1: *s = cs;  *c = -sn; break;
    case 2: *s = -sn; *c = -cs; break;
    default: *s = -cs; *c = sn; break;
  }
}

/* z[0..n): standard normals of elements 0..n-1 */
void b200dn_oracle_normals(float* z, int64_t n, uint64_t seed, uint32_t stream_id) {
  const uint32_t key[2] = {(uint32_t)seed, (uint32_t)(seed >> 32)};
  for (int64_t q = 0; q * 4 < n; ++q) {
    const uint32_t ctr[4] = {(uint32_t)q, (uint32_t)((uint64_t)q >> 32), stream_id, 0u};
    uint32_t r[4];
    b200dn_oracle_philox4x32_10(ctr, key, r);
    float quad[4];
    for (int pair = 0; pair < 2; ++pair) {
      const float u1 = unit_open(r[2 * pair]), u2 = unit_open(r[2 * pair + 1]);
      const float rad = sqrtf(-2.0f * spec_log(u1));
      float s, c;
      spec_sincos_2pi(u2, &s, &c);
      quad[2 * pair] = rad * c;
      quad[2 * pair + 1] = rad * s;
    }
    for (int j = 0; j < 4 && q * 4 + j < n; ++j) z[q * 4 + j] = quad[j];
  }
}

static float normalise_u8(uint8_t v) { return (((float)v / 255.0f) - 0.5f) / 0.5f; }

/* clean: u8 [B,H,W,C]; sigma[B]; outputs optional (NULL to skip): noisy_u8 [B,H,W,C],
   noisy_norm / clean_norm fp32 [B,C,H,W]. Element index of the generator = linear HWC index. */
void b200dn_oracle_degrade(const uint8_t* clean, int B, int H, int W, int C, const float* sigma, uint64_t seed,
                           uint32_t stream_id, const float* z_all, uint8_t* noisy_u8, float* noisy_norm,
                           float* clean_norm) {
  (void)seed; (void)stream_id;
  const int64_t hw = (int64_t)H * W;
  for (int b = 0; b < B; ++b)
    for (int64_t sp = 0; sp < hw; ++sp)
      for (int c = 0; c < C; ++c) {
        const int64_t e = ((int64_t)b * hw + sp) * C + c;
        float v = (float)clean[e] + sigma[b] * z_all[e];
        if (v < 0.0f) v = 0.0f;
        if (v > 255.0f) v = 255.0f;
        const uint8_t q = (uint8_t)(int)v; /* astype(uint8): truncation toward zero */
        if (noisy_u8) noisy_u8[e] = q;
        if (noisy_norm) noisy_norm[((int64_t)b * C + c) * hw + sp] = normalise_u8(q);
        if (clean_norm) clean_norm[((int64_t)b * C + c) * hw + sp] = normalise_u8(clean[e]);
      }
}

/* fp32 [B,C,H,W] in [-1,1] -> u8 [B,H,W,C]  (benchmark.py:42-44) */
void b200dn_oracle_norm_to_u8(const float* in, int B, int H, int W, int C, uint8_t* out) {
  const int64_t hw = (int64_t)H * W;
  for (int b = 0; b < B; ++b)
    for (int64_t sp = 0; sp < hw; ++sp)
      for (int c = 0; c < C; ++c) {
        float v = in[((int64_t)b * C + c) * hw + sp];
        v = (v + 1.0f) / 2.0f;
        v = v * 255.0f;
        if (v < 0.0f) v = 0.0f;
        if (v > 255.0f) v = 255.0f;
        out[((int64_t)b * hw + sp) * C + c] = (uint8_t)(int)v;
      }
}
