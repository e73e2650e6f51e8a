// noise.cu — on-device Gaussian-noise synthesis for the sigma = 10..50 noisy-patch stage
// (dataset_creation/custom_dataset.py:83-87 + dataset_creation/data_loader.py:35-38): Philox4x32-10 +
// Box-Muller, then clip / truncate-to-uint8 / normalise, fused in one pass.  HBM-bound: reads 1 byte and
// writes up to 1 + 4 + 4 bytes per sample.  The generator spec lives in philox_normal.h.
#include "common.cuh"
#include "philox_normal.h"

namespace b200dn {

namespace {

constexpr int NZ_THREADS = 256;

// One thread = 4 consecutive pixels of one row = 4*C samples = C Philox quads.
template <int C>
__global__ void __launch_bounds__(NZ_THREADS) gauss_noise_kernel(const uint8_t* __restrict__ clean, int H, int W,
                                                                 const float* __restrict__ sigma, uint64_t seed,
                                                                 uint32_t stream_id, uint8_t* __restrict__ noisy_u8,
                                                                 float* __restrict__ noisy_norm,
                                                                 float* __restrict__ clean_norm, int64_t n_groups) {
  // (k/255 - 0.5)/0.5 tabulated with the reference's IEEE divisions (256 possible inputs)
  __shared__ float lut[256];
  for (int k = threadIdx.x; k < 256; k += blockDim.x)
    lut[k] = __fdiv_rn(__fsub_rn(__fdiv_rn(static_cast<float>(k), 255.f), 0.5f), 0.5f);
  __syncthreads();
  const int64_t hw = static_cast<int64_t>(H) * W;
  const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
  for (int64_t g = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; g < n_groups; g += stride) {
    const int64_t pix0 = g * 4;           // first pixel (linear over B*H*W)
    const int64_t b = pix0 / hw;
    const int64_t sp0 = pix0 - b * hw;    // W % 4 == 0 -> the 4 pixels share a row and an image
    const float sg = __ldg(sigma + b);
    uint8_t cu[4 * C], nu[4 * C];
    const uint32_t* src = reinterpret_cast<const uint32_t*>(clean + pix0 * C);
#pragma unroll
    for (int i = 0; i < C; ++i) {
      const uint32_t w = __ldg(src + i);
      cu[4 * i + 0] = w & 0xff;
      cu[4 * i + 1] = (w >> 8) & 0xff;
      cu[4 * i + 2] = (w >> 16) & 0xff;
      cu[4 * i + 3] = (w >> 24) & 0xff;
    }
#pragma unroll
    for (int i = 0; i < C; ++i) {
      float z[4];
      b2n_normal4(static_cast<uint64_t>(g) * C + i, seed, stream_id, z);
#pragma unroll
      for (int j = 0; j < 4; ++j) nu[4 * i + j] = b2n_degrade_u8(cu[4 * i + j], sg, z[j]);
    }
    if (noisy_u8 != nullptr) {
      uint32_t* dst = reinterpret_cast<uint32_t*>(noisy_u8 + pix0 * C);
#pragma unroll
      for (int i = 0; i < C; ++i)
        dst[i] = nu[4 * i] | (nu[4 * i + 1] << 8) | (nu[4 * i + 2] << 16) | (static_cast<uint32_t>(nu[4 * i + 3]) << 24);
    }
#pragma unroll
    for (int c = 0; c < C; ++c) {
      if (noisy_norm != nullptr) {
        float4 o;
        o.x = lut[nu[0 * C + c]];
        o.y = lut[nu[1 * C + c]];
        o.z = lut[nu[2 * C + c]];
        o.w = lut[nu[3 * C + c]];
        *reinterpret_cast<float4*>(noisy_norm + (b * C + c) * hw + sp0) = o;
      }
      if (clean_norm != nullptr) {
        float4 o;
        o.x = lut[cu[0 * C + c]];
        o.y = lut[cu[1 * C + c]];
        o.z = lut[cu[2 * C + c]];
        o.w = lut[cu[3 * C + c]];
        *reinterpret_cast<float4*>(clean_norm + (b * C + c) * hw + sp0) = o;
      }
    }
  }
}

__global__ void __launch_bounds__(NZ_THREADS) philox_normal_kernel(float* __restrict__ z, int64_t n, uint64_t seed,
                                                                   uint32_t stream_id) {
  const int64_t nq = (n + 3) / 4;
  const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
  for (int64_t q = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; q < nq; q += stride) {
    float v[4];
    b2n_normal4(static_cast<uint64_t>(q), seed, stream_id, v);
#pragma unroll
    for (int j = 0; j < 4; ++j)
      if (q * 4 + j < n) z[q * 4 + j] = v[j];
  }
}

int nz_grid(int64_t items) {
  int sms = device_sm_count();
  if (sms <= 0) return sms;
  int64_t blocks = cdiv64(items, NZ_THREADS);
  const int64_t cap = static_cast<int64_t>(sms) * 16;
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  return static_cast<int>(blocks);
}

}  // namespace
}  // namespace b200dn

extern "C" int b200dn_gauss_noise_u8(const uint8_t* clean_u8, int B, int H, int W, int C, const float* sigma,
                                     uint64_t seed, uint32_t stream_id, uint8_t* noisy_u8, float* noisy_norm,
                                     float* clean_norm, void* stream) {
  using namespace b200dn;
  B200DN_CHECK_ARG(clean_u8 && sigma && B > 0 && H > 0 && W > 0, "gauss_noise_u8: bad arguments");
  B200DN_CHECK_ARG(C == 1 || C == 3, "gauss_noise_u8: C must be 1 or 3 (got %d)", C);
  B200DN_CHECK_ARG(W % 4 == 0, "gauss_noise_u8: W %d must be a multiple of 4", W);
  B200DN_CHECK_ARG((reinterpret_cast<uintptr_t>(clean_u8) & 3) == 0 && (reinterpret_cast<uintptr_t>(noisy_u8) & 3) == 0,
                   "gauss_noise_u8: u8 buffers must be 4-byte aligned");
  B200DN_CHECK_ARG((reinterpret_cast<uintptr_t>(noisy_norm) & 15) == 0 && (reinterpret_cast<uintptr_t>(clean_norm) & 15) == 0,
                   "gauss_noise_u8: fp32 outputs must be 16-byte aligned");
  if (int rc = require_sm100()) return rc;
  const int64_t n_groups = static_cast<int64_t>(B) * H * W / 4;
  const int grid = nz_grid(n_groups);
  if (grid <= 0) return B200DN_E_CUDA;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  if (C == 3)
    gauss_noise_kernel<3><<<grid, NZ_THREADS, 0, s>>>(clean_u8, H, W, sigma, seed, stream_id, noisy_u8, noisy_norm,
                                                      clean_norm, n_groups);
  else
    gauss_noise_kernel<1><<<grid, NZ_THREADS, 0, s>>>(clean_u8, H, W, sigma, seed, stream_id, noisy_u8, noisy_norm,
                                                      clean_norm, n_groups);
  B200DN_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int b200dn_philox_normal(float* z, int64_t n, uint64_t seed, uint32_t stream_id, void* stream) {
  using namespace b200dn;
  B200DN_CHECK_ARG(z && n > 0, "philox_normal: bad arguments");
  if (int rc = require_sm100()) return rc;
  const int grid = nz_grid((n + 3) / 4);
  if (grid <= 0) return B200DN_E_CUDA;
  philox_normal_kernel<<<grid, NZ_THREADS, 0, static_cast<cudaStream_t>(stream)>>>(z, n, seed, stream_id);
  B200DN_CUDA(cudaGetLastError());
  return 0;
}
