// metrics.cu — on-device PSNR / SSIM reductions (warp-shuffle + 128-bit loads).
//   PSNR: mse = mean((X-Y)^2); 10*log10(R^2/mse)          evaluate_Unet_diffusion/evaluate_model.py:36-41,
//                                                         evaluate_SIDD/evaluate_SIDD.py:63 (skimage PSNR)
//   SSIM: skimage.metrics.structural_similarity defaults   evaluate_Unet_diffusion/evaluate_model.py:30-34,
//         (scikit-image==0.22.0, requirements.txt:98)      evaluate_SIDD/evaluate_SIDD.py:64
// The kernels return raw sums (fp64); the host-side mirror turns them into dB / means, so a sharded run
// can all-reduce the sums.
#include "common.cuh"

namespace b200dn {

namespace {

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

template <int THREADS>
__device__ __forceinline__ double block_sum(double v, double* red) {
  v = warp_sum(v);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (lane == 0) red[warp] = v;
  __syncthreads();
  double t = 0.0;
  if (warp == 0) {
    t = (lane < THREADS / 32) ? red[lane] : 0.0;
    t = warp_sum(t);
  }
  return t;  // valid in warp 0
}

constexpr int SSE_THREADS = 256;

// grid = (chunks, n_images); each block reduces a contiguous chunk of one image, one atomicAdd(double) per block.
__global__ void __launch_bounds__(SSE_THREADS) sse_kernel(const float* __restrict__ a, const float* __restrict__ b,
                                                          int64_t n_per_image, int vec_ok, double* __restrict__ sse) {
  __shared__ double red[SSE_THREADS / 32];
  const int64_t img = blockIdx.y;
  const float* pa = a + img * n_per_image;
  const float* pb = b + img * n_per_image;
  float acc0 = 0.f, acc1 = 0.f, acc2 = 0.f, acc3 = 0.f;
  double dacc = 0.0;
  const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
  const int64_t tid = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (vec_ok) {
    const int64_t n4 = n_per_image / 4;
    int it = 0;
    for (int64_t i = tid; i < n4; i += stride) {
      const float4 x = __ldg(reinterpret_cast<const float4*>(pa) + i);
      const float4 y = __ldg(reinterpret_cast<const float4*>(pb) + i);
      const float d0 = __fsub_rn(x.x, y.x), d1 = __fsub_rn(x.y, y.y), d2 = __fsub_rn(x.z, y.z), d3 = __fsub_rn(x.w, y.w);
      acc0 += __fmul_rn(d0, d0);
      acc1 += __fmul_rn(d1, d1);
      acc2 += __fmul_rn(d2, d2);
      acc3 += __fmul_rn(d3, d3);
      if (++it == 16) {  // spill the short fp32 partial sums into fp64 regularly
        dacc += static_cast<double>(acc0) + static_cast<double>(acc1) + static_cast<double>(acc2) + static_cast<double>(acc3);
        acc0 = acc1 = acc2 = acc3 = 0.f;
        it = 0;
      }
    }
    for (int64_t i = n4 * 4 + tid; i < n_per_image; i += stride) {
      const float d = __fsub_rn(pa[i], pb[i]);
      dacc += static_cast<double>(__fmul_rn(d, d));
    }
  } else {
    for (int64_t i = tid; i < n_per_image; i += stride) {
      const float d = __fsub_rn(pa[i], pb[i]);
      dacc += static_cast<double>(__fmul_rn(d, d));
    }
  }
  dacc += static_cast<double>(acc0) + static_cast<double>(acc1) + static_cast<double>(acc2) + static_cast<double>(acc3);
  const double t = block_sum<SSE_THREADS>(dacc, red);
  if (threadIdx.x == 0) atomicAdd(sse + img, t);
}

// ------------------------------------------------------------------ SSIM
// Tile of TS x TS interior outputs per block; needs a (TS+6)^2 input halo.  Follows skimage 0.22:
// uniform_filter (7 taps, axis 0 then axis 1, each pass accumulated in double and rounded to float32),
// cov_norm = 49/48, float32 elementwise math in skimage's operation order, crop 3, float64 sum.
constexpr int TS = 32;
constexpr int HALO = 6;
constexpr int IN_T = TS + HALO;  // 38
constexpr int SSIM_THREADS = 160;   // 152 vertical / 128 horizontal strip tasks per 32x32 tile
constexpr double kInv7 = 1.0 / 7.0;

// Each pass is done in strips of 8 outputs with a running window sum (add the entering sample, subtract the
// leaving one; all in double, where sums of seven float32 values are exact), which cuts the double adds and the
// shared-memory reads per output by ~2.7x / 4x against recomputing every 7-tap sum.
__global__ void __launch_bounds__(SSIM_THREADS) ssim_kernel(const float* __restrict__ a, const float* __restrict__ b,
                                                            int H, int W, float C1, float C2, float cov_norm,
                                                            double* __restrict__ ssim_sum) {
  __shared__ float sx[IN_T][IN_T + 1];
  __shared__ float sy[IN_T][IN_T + 1];
  __shared__ float v[5][TS][IN_T + 1];  // vertical (axis-0) pass of x, y, xx, yy, xy
  __shared__ double red[SSIM_THREADS / 32];

  const int64_t plane = blockIdx.z;
  const float* pa = a + plane * static_cast<int64_t>(H) * W;
  const float* pb = b + plane * static_cast<int64_t>(H) * W;
  const int oy0 = blockIdx.y * TS, ox0 = blockIdx.x * TS;  // origin in the cropped (H-6)x(W-6) output
  const int OH = H - HALO, OW = W - HALO;

  for (int i = threadIdx.x; i < IN_T * IN_T; i += SSIM_THREADS) {
    const int r = i / IN_T, c = i - r * IN_T;
    const int gy = oy0 + r, gx = ox0 + c;
    const bool in = gy < H && gx < W;
    sx[r][c] = in ? __ldg(pa + static_cast<int64_t>(gy) * W + gx) : 0.f;
    sy[r][c] = in ? __ldg(pb + static_cast<int64_t>(gy) * W + gx) : 0.f;
  }
  __syncthreads();

  // vertical pass: task = (column c, strip of 8 output rows)
  for (int task = threadIdx.x; task < IN_T * (TS / 8); task += SSIM_THREADS) {
    const int c = task % IN_T, r0 = (task / IN_T) * 8;
    double s0 = 0, s1 = 0, s2 = 0, s3 = 0, s4 = 0;
#pragma unroll
    for (int k = 0; k < 6; ++k) {
      const float x = sx[r0 + k][c], y = sy[r0 + k][c];
      s0 += static_cast<double>(x);
      s1 += static_cast<double>(y);
      s2 += static_cast<double>(__fmul_rn(x, x));
      s3 += static_cast<double>(__fmul_rn(y, y));
      s4 += static_cast<double>(__fmul_rn(x, y));
    }
#pragma unroll
    for (int o = 0; o < 8; ++o) {
      const float x = sx[r0 + o + 6][c], y = sy[r0 + o + 6][c];
      s0 += static_cast<double>(x);
      s1 += static_cast<double>(y);
      s2 += static_cast<double>(__fmul_rn(x, x));
      s3 += static_cast<double>(__fmul_rn(y, y));
      s4 += static_cast<double>(__fmul_rn(x, y));
      // scipy's uniform_filter1d keeps a running mean in double and casts to float32; sum * (1/7) differs from
      // sum / 7 by at most one double ulp, i.e. changes the float32 result in ~1e-9 of cases
      v[0][r0 + o][c] = static_cast<float>(s0 * kInv7);
      v[1][r0 + o][c] = static_cast<float>(s1 * kInv7);
      v[2][r0 + o][c] = static_cast<float>(s2 * kInv7);
      v[3][r0 + o][c] = static_cast<float>(s3 * kInv7);
      v[4][r0 + o][c] = static_cast<float>(s4 * kInv7);
      const float xo = sx[r0 + o][c], yo = sy[r0 + o][c];
      s0 -= static_cast<double>(xo);
      s1 -= static_cast<double>(yo);
      s2 -= static_cast<double>(__fmul_rn(xo, xo));
      s3 -= static_cast<double>(__fmul_rn(yo, yo));
      s4 -= static_cast<double>(__fmul_rn(xo, yo));
    }
  }
  __syncthreads();

  // horizontal pass + SSIM map: task = (row r, strip of 8 output columns)
  double local = 0.0;
  for (int task = threadIdx.x; task < TS * (TS / 8); task += SSIM_THREADS) {
    const int r = task / (TS / 8), c0 = (task % (TS / 8)) * 8;
    if (oy0 + r >= OH) continue;
    double s0 = 0, s1 = 0, s2 = 0, s3 = 0, s4 = 0;
#pragma unroll
    for (int k = 0; k < 6; ++k) {
      s0 += static_cast<double>(v[0][r][c0 + k]);
      s1 += static_cast<double>(v[1][r][c0 + k]);
      s2 += static_cast<double>(v[2][r][c0 + k]);
      s3 += static_cast<double>(v[3][r][c0 + k]);
      s4 += static_cast<double>(v[4][r][c0 + k]);
    }
#pragma unroll
    for (int o = 0; o < 8; ++o) {
      s0 += static_cast<double>(v[0][r][c0 + o + 6]);
      s1 += static_cast<double>(v[1][r][c0 + o + 6]);
      s2 += static_cast<double>(v[2][r][c0 + o + 6]);
      s3 += static_cast<double>(v[3][r][c0 + o + 6]);
      s4 += static_cast<double>(v[4][r][c0 + o + 6]);
      if (ox0 + c0 + o < OW) {
        const float ux = static_cast<float>(s0 * kInv7), uy = static_cast<float>(s1 * kInv7);
        const float uxx = static_cast<float>(s2 * kInv7), uyy = static_cast<float>(s3 * kInv7);
        const float uxy = static_cast<float>(s4 * kInv7);
        const float vx = __fmul_rn(cov_norm, __fsub_rn(uxx, __fmul_rn(ux, ux)));
        const float vy = __fmul_rn(cov_norm, __fsub_rn(uyy, __fmul_rn(uy, uy)));
        const float vxy = __fmul_rn(cov_norm, __fsub_rn(uxy, __fmul_rn(ux, uy)));
        const float A1 = __fadd_rn(__fmul_rn(__fmul_rn(2.f, ux), uy), C1);
        const float A2 = __fadd_rn(__fmul_rn(2.f, vxy), C2);
        const float B1 = __fadd_rn(__fadd_rn(__fmul_rn(ux, ux), __fmul_rn(uy, uy)), C1);
        const float B2 = __fadd_rn(__fadd_rn(vx, vy), C2);
        const float D = __fmul_rn(B1, B2);
        const float S = __fdiv_rn(__fmul_rn(A1, A2), D);
        local += static_cast<double>(S);
      }
      s0 -= static_cast<double>(v[0][r][c0 + o]);
      s1 -= static_cast<double>(v[1][r][c0 + o]);
      s2 -= static_cast<double>(v[2][r][c0 + o]);
      s3 -= static_cast<double>(v[3][r][c0 + o]);
      s4 -= static_cast<double>(v[4][r][c0 + o]);
    }
  }
  const double t = block_sum<SSIM_THREADS>(local, red);
  if (threadIdx.x == 0) atomicAdd(ssim_sum + plane, t);
}

}  // namespace
}  // namespace b200dn

extern "C" int b200dn_psnr_sse(const float* a, const float* b, int64_t n_images, int64_t n_per_image, double* sse,
                               void* stream) {
  using namespace b200dn;
  B200DN_CHECK_ARG(a && b && sse && n_images > 0 && n_per_image > 0, "psnr_sse: bad arguments");
  B200DN_CHECK_ARG(n_images <= 65535, "psnr_sse: at most 65535 images per call");
  if (int rc = require_sm100()) return rc;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  B200DN_CUDA(cudaMemsetAsync(sse, 0, sizeof(double) * n_images, s));
  const int vec_ok = ((reinterpret_cast<uintptr_t>(a) | reinterpret_cast<uintptr_t>(b)) & 15) == 0 && (n_per_image % 4 == 0);
  int sms = device_sm_count();
  if (sms <= 0) return B200DN_E_CUDA;
  // enough blocks per image to cover the machine ~4x, at least 4 float4 per thread
  int64_t chunks = cdiv64(n_per_image / 4, static_cast<int64_t>(SSE_THREADS) * 4);
  const int64_t want = cdiv64(static_cast<int64_t>(sms) * 16, n_images);
  if (chunks > want) chunks = want;
  if (chunks < 1) chunks = 1;
  dim3 grid(static_cast<unsigned>(chunks), static_cast<unsigned>(n_images));
  sse_kernel<<<grid, SSE_THREADS, 0, s>>>(a, b, n_per_image, vec_ok, sse);
  B200DN_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int b200dn_ssim(const float* a, const float* b, int64_t n_planes, int H, int W, float data_range,
                           double* ssim_sum, void* stream) {
  using namespace b200dn;
  B200DN_CHECK_ARG(a && b && ssim_sum && n_planes > 0, "ssim: bad arguments");
  B200DN_CHECK_ARG(H >= 7 && W >= 7, "ssim: win_size 7 exceeds image extent %d x %d", H, W);
  B200DN_CHECK_ARG(n_planes <= 65535, "ssim: at most 65535 planes per call");
  if (int rc = require_sm100()) return rc;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  B200DN_CUDA(cudaMemsetAsync(ssim_sum, 0, sizeof(double) * n_planes, s));
  // python: C1 = (K1*R)**2 in double, then used against float32 arrays (NEP 50 weak scalar -> float32)
  const double R = static_cast<double>(data_range);
  const float C1 = static_cast<float>((0.01 * R) * (0.01 * R));
  const float C2 = static_cast<float>((0.03 * R) * (0.03 * R));
  const float cov_norm = static_cast<float>(49.0 / 48.0);
  dim3 grid(cdiv(W - HALO, TS), cdiv(H - HALO, TS), static_cast<unsigned>(n_planes));
  ssim_kernel<<<grid, SSIM_THREADS, 0, s>>>(a, b, H, W, C1, C2, cov_norm, ssim_sum);
  B200DN_CUDA(cudaGetLastError());
  return 0;
}
