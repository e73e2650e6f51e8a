// elementwise.cu — HBM-bound elementwise kernels of the sampler and the data-format boundary.
//   sampler_step : diffusion_denoising/diffusion_RDUnet.py:45,48,49 (one fused launch per timestep instead of 7)
//   lerp         : diffusion_denoising/diffusion_RDUnet.py:33-36 (forward_diffusion)
//   u8_to_norm   : ToTensor + Normalize(0.5,0.5)  evaluate_SIDD/evaluate_SIDD.py:23-26, benchmark.py:35-36
//   norm_to_u8   : (y+1)/2 -> clip(*255,0,255) -> uint8   evaluate_SIDD/benchmark.py:42-44
// All fp32 arithmetic uses explicit round-to-nearest intrinsics in the reference's operation order, so
// the results are bit-identical to PyTorch/numpy fp32 elementwise math (no FMA contraction).
#include "common.cuh"

namespace b200dn {

namespace {

constexpr int EW_THREADS = 256;

__device__ __forceinline__ float step1(float x, float u1, float u2, float y, float c1, float at, float c2, float ap) {
  // x_tilde      = (1 - alpha_t)      * unet(x_t, t)      + alpha_t      * noisy
  // x_tilde_prev = (1 - alpha_t_prev) * unet(x_t, t_prev) + alpha_t_prev * noisy
  // x_{t-1}      = x_t - x_tilde + x_tilde_prev
  const float xt = __fadd_rn(__fmul_rn(c1, u1), __fmul_rn(at, y));
  const float xp = __fadd_rn(__fmul_rn(c2, u2), __fmul_rn(ap, y));
  return __fadd_rn(__fsub_rn(x, xt), xp);
}

__global__ void __launch_bounds__(EW_THREADS) sampler_step_kernel(const float* __restrict__ x,
                                                                  const float* __restrict__ u1,
                                                                  const float* __restrict__ u2,
                                                                  const float* __restrict__ y, float c1, float at,
                                                                  float c2, float ap, float* __restrict__ xn,
                                                                  int64_t n4, int64_t n) {
  const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n4; i += stride) {
    const float4 a = __ldg(reinterpret_cast<const float4*>(x) + i);
    const float4 b = __ldg(reinterpret_cast<const float4*>(u1) + i);
    const float4 c = __ldg(reinterpret_cast<const float4*>(u2) + i);
    const float4 d = __ldg(reinterpret_cast<const float4*>(y) + i);
    float4 o;
    o.x = step1(a.x, b.x, c.x, d.x, c1, at, c2, ap);
    o.y = step1(a.y, b.y, c.y, d.y, c1, at, c2, ap);
    o.z = step1(a.z, b.z, c.z, d.z, c1, at, c2, ap);
    o.w = step1(a.w, b.w, c.w, d.w, c1, at, c2, ap);
    reinterpret_cast<float4*>(xn)[i] = o;
  }
  // tail (n not a multiple of 4)
  for (int64_t i = n4 * 4 + static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += stride)
    xn[i] = step1(x[i], u1[i], u2[i], y[i], c1, at, c2, ap);
}

__global__ void __launch_bounds__(EW_THREADS) lerp_kernel(const float* __restrict__ clean,
                                                          const float* __restrict__ noisy, float alpha, float oma,
                                                          float* __restrict__ out, int64_t n) {
  const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += stride)
    out[i] = __fadd_rn(__fmul_rn(alpha, noisy[i]), __fmul_rn(oma, clean[i]));
}

// (k/255 - 0.5)/0.5 for k = 0..255 with the reference's two IEEE divisions (ToTensor's /255, Normalize's /0.5):
// only 256 inputs exist, so each block tabulates them once and the per-sample work is one shared-memory lookup.
__device__ __forceinline__ void build_norm_lut(float* lut) {
  for (int k = threadIdx.x; k < 256; k += blockDim.x)
    lut[k] = __fdiv_rn(__fsub_rn(__fdiv_rn(static_cast<float>(k), 255.f), 0.5f), 0.5f);
  __syncthreads();
}

// in: u8 [B,H,W,C]; out: fp32 [B,C,H,W].  One thread = 4 consecutive pixels of a row (W % 4 == 0 fast path).
template <int C>
__global__ void __launch_bounds__(EW_THREADS) u8_to_norm_vec_kernel(const uint8_t* __restrict__ in, int64_t hw,
                                                                    float* __restrict__ out, int64_t n_groups) {
  __shared__ float lut[256];
  build_norm_lut(lut);
  const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
  for (int64_t g = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; g < n_groups; g += stride) {
    const int64_t pix0 = g * 4;
    const int64_t b = pix0 / hw, sp0 = pix0 - b * hw;
    uint8_t u[4 * C];
    const uint32_t* src = reinterpret_cast<const uint32_t*>(in + pix0 * C);
#pragma unroll
    for (int i = 0; i < C; ++i) {
      const uint32_t w = __ldg(src + i);
      u[4 * i + 0] = w & 0xff, u[4 * i + 1] = (w >> 8) & 0xff, u[4 * i + 2] = (w >> 16) & 0xff, u[4 * i + 3] = (w >> 24) & 0xff;
    }
#pragma unroll
    for (int c = 0; c < C; ++c) {
      float4 o;
      o.x = lut[u[0 * C + c]], o.y = lut[u[1 * C + c]], o.z = lut[u[2 * C + c]], o.w = lut[u[3 * C + c]];
      *reinterpret_cast<float4*>(out + (b * C + c) * hw + sp0) = o;
    }
  }
}

// generic fallback (any C, any W)
__global__ void __launch_bounds__(EW_THREADS) u8_to_norm_kernel(const uint8_t* __restrict__ in, int H, int W, int C,
                                                                float* __restrict__ out, int64_t n_pix_total) {
  __shared__ float lut[256];
  build_norm_lut(lut);
  const int64_t hw = static_cast<int64_t>(H) * W;
  const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
  for (int64_t p = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; p < n_pix_total; p += stride) {
    const int64_t b = p / hw, sp = p - b * hw;
    for (int c = 0; c < C; ++c) out[(b * C + c) * hw + sp] = lut[in[p * C + c]];
  }
}

__device__ __forceinline__ uint32_t quantise_u8(float v) {
  // (v + 1) / 2 * 255, clip, truncate (benchmark.py:42-44); the division by 2 is exact, so a multiply matches it
  v = __fmul_rn(__fmul_rn(__fadd_rn(v, 1.f), 0.5f), 255.f);
  v = fminf(fmaxf(v, 0.f), 255.f);  // NaN -> 0 like a saturating cast would; reference is UB there
  return static_cast<uint32_t>(static_cast<int>(v));  // truncation, as numpy astype(uint8)
}

// in: fp32 [B,C,H,W]; out: u8 [B,H,W,C].  One thread = 4 consecutive pixels (W % 4 == 0 fast path).
template <int C>
__global__ void __launch_bounds__(EW_THREADS) norm_to_u8_vec_kernel(const float* __restrict__ in, int64_t hw,
                                                                    uint8_t* __restrict__ out, int64_t n_groups) {
  const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
  for (int64_t g = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; g < n_groups; g += stride) {
    const int64_t pix0 = g * 4;
    const int64_t b = pix0 / hw, sp0 = pix0 - b * hw;
    uint32_t q[4 * C];
#pragma unroll
    for (int c = 0; c < C; ++c) {
      const float4 v = __ldg(reinterpret_cast<const float4*>(in + (b * C + c) * hw + sp0));
      q[0 * C + c] = quantise_u8(v.x), q[1 * C + c] = quantise_u8(v.y), q[2 * C + c] = quantise_u8(v.z),
                q[3 * C + c] = quantise_u8(v.w);
    }
    uint32_t* dst = reinterpret_cast<uint32_t*>(out + pix0 * C);
#pragma unroll
    for (int i = 0; i < C; ++i) dst[i] = q[4 * i] | (q[4 * i + 1] << 8) | (q[4 * i + 2] << 16) | (q[4 * i + 3] << 24);
  }
}

__global__ void __launch_bounds__(EW_THREADS) norm_to_u8_kernel(const float* __restrict__ in, int H, int W, int C,
                                                                uint8_t* __restrict__ out, int64_t n_pix_total) {
  const int64_t hw = static_cast<int64_t>(H) * W;
  const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
  for (int64_t p = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; p < n_pix_total; p += stride) {
    const int64_t b = p / hw, sp = p - b * hw;
    for (int c = 0; c < C; ++c) out[p * C + c] = static_cast<uint8_t>(quantise_u8(in[(b * C + c) * hw + sp]));
  }
}

int ew_grid(int64_t work_items) {
  int sms = device_sm_count();
  if (sms <= 0) return sms;
  int64_t blocks = cdiv64(work_items, EW_THREADS);
  const int64_t cap = static_cast<int64_t>(sms) * 16;
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  return static_cast<int>(blocks);
}

}  // namespace
}  // namespace b200dn

extern "C" int b200dn_sampler_step(const float* x, const float* u1, const float* u2, const float* y, float one_m_at,
                                   float at, float one_m_ap, float ap, float* x_next, int64_t n, void* stream) {
  using namespace b200dn;
  B200DN_CHECK_ARG(x && u1 && u2 && y && x_next && n > 0, "sampler_step: bad arguments");
  if (int rc = require_sm100()) return rc;
  const bool aligned = ((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(u1) |
                         reinterpret_cast<uintptr_t>(u2) | reinterpret_cast<uintptr_t>(y) |
                         reinterpret_cast<uintptr_t>(x_next)) & 15) == 0;
  const int64_t n4 = aligned ? n / 4 : 0;
  const int grid = ew_grid(n4 > 0 ? n4 : n);
  if (grid <= 0) return B200DN_E_CUDA;
  sampler_step_kernel<<<grid, EW_THREADS, 0, static_cast<cudaStream_t>(stream)>>>(x, u1, u2, y, one_m_at, at, one_m_ap,
                                                                                  ap, x_next, n4, n);
  B200DN_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int b200dn_lerp(const float* clean, const float* noisy, float alpha, float one_m_alpha, float* out,
                           int64_t n, void* stream) {
  using namespace b200dn;
  B200DN_CHECK_ARG(clean && noisy && out && n > 0, "lerp: bad arguments");
  if (int rc = require_sm100()) return rc;
  const int grid = ew_grid(n);
  if (grid <= 0) return B200DN_E_CUDA;
  lerp_kernel<<<grid, EW_THREADS, 0, static_cast<cudaStream_t>(stream)>>>(clean, noisy, alpha, one_m_alpha, out, n);
  B200DN_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int b200dn_u8_to_norm(const uint8_t* in, int B, int H, int W, int C, float* out, void* stream) {
  using namespace b200dn;
  B200DN_CHECK_ARG(in && out && B > 0 && H > 0 && W > 0 && C > 0, "u8_to_norm: bad arguments");
  if (int rc = require_sm100()) return rc;
  const int64_t npix = static_cast<int64_t>(B) * H * W;
  const bool vec = (W % 4 == 0) && (C == 1 || C == 3) && ((reinterpret_cast<uintptr_t>(in) & 3) == 0) &&
                   ((reinterpret_cast<uintptr_t>(out) & 15) == 0);
  const int grid = ew_grid(vec ? npix / 4 : npix);
  if (grid <= 0) return B200DN_E_CUDA;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int64_t hw = static_cast<int64_t>(H) * W;
  if (vec && C == 3)
    u8_to_norm_vec_kernel<3><<<grid, EW_THREADS, 0, st>>>(in, hw, out, npix / 4);
  else if (vec)
    u8_to_norm_vec_kernel<1><<<grid, EW_THREADS, 0, st>>>(in, hw, out, npix / 4);
  else
    u8_to_norm_kernel<<<grid, EW_THREADS, 0, st>>>(in, H, W, C, out, npix);
  B200DN_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int b200dn_norm_to_u8(const float* in, int B, int H, int W, int C, uint8_t* out, void* stream) {
  using namespace b200dn;
  B200DN_CHECK_ARG(in && out && B > 0 && H > 0 && W > 0 && C > 0, "norm_to_u8: bad arguments");
  if (int rc = require_sm100()) return rc;
  const int64_t npix = static_cast<int64_t>(B) * H * W;
  const bool vec = (W % 4 == 0) && (C == 1 || C == 3) && ((reinterpret_cast<uintptr_t>(out) & 3) == 0) &&
                   ((reinterpret_cast<uintptr_t>(in) & 15) == 0);
  const int grid = ew_grid(vec ? npix / 4 : npix);
  if (grid <= 0) return B200DN_E_CUDA;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int64_t hw = static_cast<int64_t>(H) * W;
  if (vec && C == 3)
    norm_to_u8_vec_kernel<3><<<grid, EW_THREADS, 0, st>>>(in, hw, out, npix / 4);
  else if (vec)
    norm_to_u8_vec_kernel<1><<<grid, EW_THREADS, 0, st>>>(in, hw, out, npix / 4);
  else
    norm_to_u8_kernel<<<grid, EW_THREADS, 0, st>>>(in, H, W, C, out, npix);
  B200DN_CUDA(cudaGetLastError());
  return 0;
}
