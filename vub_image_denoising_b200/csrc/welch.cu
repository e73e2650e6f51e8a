// welch.cu — Welch power-spectral-density estimate of flattened images on the device.
//
// Reference call: `welch(image.flatten(), nperseg=256)` (scipy.signal), evaluate_Unet_diffusion/plot.py:155-157,
// 233-235, 287-290 — the remaining per-image host cost of evaluate_model.py once PSNR/SSIM are on the device
// (SURVEY.md §8 row f4).  scipy defaults reproduced here: fs = 1, periodic Hann window, noverlap = 128, nfft = 256,
// detrend = 'constant' (each segment's mean removed), one-sided density scaling (x2 except DC / Nyquist,
// 1 / (fs * sum(w^2)) = 1/96), mean over segments; float32 input -> complex64 transform -> float32 spectrum.
//
// One block = 128 threads = one 256-point complex radix-2 Stockham FFT in shared memory per step, carrying TWO real
// segments (z = a + i b; A[k] = (Z[k] + conj Z[N-k]) / 2, B[k] = (Z[k] - conj Z[N-k]) / 2i).  A block walks a strided
// share of the segment pairs of one signal and writes one partial spectrum; a second kernel adds the partials in a
// fixed order (bit-reproducible, no atomics).  Bytes: each sample is read twice (50 % overlap, the second time from
// L1/L2) and 129 floats are written per signal, so the kernel is bound by the FFT's shared-memory traffic, not by HBM.
#include "common.cuh"

namespace b200dn {

namespace {

constexpr int NSEG = 256;            // nperseg = nfft
constexpr int HOP = 128;             // nperseg - noverlap
constexpr int NBINS = NSEG / 2 + 1;  // 129
constexpr int WT = 128;              // threads per block
constexpr int MAX_BLOCKS_PER_SIGNAL = 64;

__device__ __forceinline__ float2 cmul(float2 a, float2 b) { return make_float2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x); }

__device__ __forceinline__ float block_sum128(float v, float* red, int slot) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  if ((threadIdx.x & 31) == 0) red[slot * 4 + (threadIdx.x >> 5)] = v;
  __syncthreads();
  return (red[slot * 4] + red[slot * 4 + 1]) + (red[slot * 4 + 2] + red[slot * 4 + 3]);
}

__global__ void __launch_bounds__(WT) welch_kernel(const float* __restrict__ x, int64_t n, int64_t nseg,
                                                   float* __restrict__ partial) {
  __shared__ float2 buf[2][NSEG];
  __shared__ float2 tw[NSEG / 2];
  __shared__ float win[NSEG];
  __shared__ float red[16];
  const int tid = threadIdx.x;
  const float* sig = x + static_cast<int64_t>(blockIdx.y) * n;
  {
    float s, c;
    sincospif(-static_cast<float>(tid) / 128.0f, &s, &c);   // exp(-2 pi i tid / 256)
    tw[tid] = make_float2(c, s);
    win[tid] = 0.5f - 0.5f * cospif(static_cast<float>(tid) / 128.0f);
    win[tid + 128] = 0.5f - 0.5f * cospif(static_cast<float>(tid + 128) / 128.0f);
  }
  __syncthreads();
  float acc = 0.f, acc_nyq = 0.f;   // bin tid; thread 0 also carries bin 128
  const int64_t npairs = (nseg + 1) / 2;
  int flip = 0;
  for (int64_t pair = blockIdx.x; pair < npairs; pair += gridDim.x) {
    const int64_t s0 = 2 * pair;
    const bool has_b = s0 + 1 < nseg;
    const float* pa = sig + s0 * HOP;
    // segment b = samples [HOP, HOP + 256) of the same window: the pair spans 384 consecutive samples
    const float a0 = pa[tid], a1 = pa[tid + 128];
    const float b0 = a1, b1 = has_b ? pa[tid + 256] : 0.f;
    const float mean_a = block_sum128(a0 + a1, red, flip) * (1.0f / NSEG);
    const float mean_b = block_sum128(has_b ? b0 + b1 : 0.f, red, flip + 2) * (1.0f / NSEG);
    float2* in = buf[0];
    float2* out = buf[1];
    in[tid] = make_float2((a0 - mean_a) * win[tid], has_b ? (b0 - mean_b) * win[tid] : 0.f);
    in[tid + 128] = make_float2((a1 - mean_a) * win[tid + 128], has_b ? (b1 - mean_b) * win[tid + 128] : 0.f);
    __syncthreads();
#pragma unroll
    for (int s = 0; s < 8; ++s) {
      const int ns = 1 << s;
      const int k = tid & (ns - 1);
      const float2 u = in[tid];
      const float2 v = cmul(in[tid + 128], tw[k * (128 >> s)]);
      const int o = ((tid - k) << 1) + k;
      out[o] = make_float2(u.x + v.x, u.y + v.y);
      out[o + ns] = make_float2(u.x - v.x, u.y - v.y);
      __syncthreads();
      float2* t = in;
      in = out;
      out = t;
    }
    // 8 stages: the result is back in buf[0] (= in)
    {
      const float2 zk = in[tid];
      const float2 zn = in[(NSEG - tid) & (NSEG - 1)];
      // A = (zk + conj zn) / 2, B = (zk - conj zn) / (2i)
      const float ar = 0.5f * (zk.x + zn.x), ai = 0.5f * (zk.y - zn.y);
      const float br = 0.5f * (zk.y + zn.y), bi = -0.5f * (zk.x - zn.x);
      acc += (ar * ar + ai * ai) + (br * br + bi * bi);
      if (tid == 0) {
        const float2 zq = in[128];   // Nyquist: A = Re, B = Im
        acc_nyq += zq.x * zq.x + zq.y * zq.y;
      }
    }
    flip ^= 1;   // alternate the reduction scratch so the next pair's writes cannot race this pair's reads
    __syncthreads();
  }
  float* p = partial + (static_cast<int64_t>(blockIdx.y) * gridDim.x + blockIdx.x) * NBINS;
  p[tid] = acc;
  if (tid == 0) p[128] = acc_nyq;
}

// pxx[sig][k] = scale_k / nseg * sum over blocks of partial[sig][block][k], fixed order
__global__ void welch_finish_kernel(const float* __restrict__ partial, int n_blocks, int64_t nseg, float* __restrict__ pxx) {
  const int k = threadIdx.x;
  if (k >= NBINS) return;
  const float* p = partial + static_cast<int64_t>(blockIdx.x) * n_blocks * NBINS + k;
  double s = 0.0;
  for (int b = 0; b < n_blocks; ++b) s += static_cast<double>(p[static_cast<int64_t>(b) * NBINS]);
  const double scale = ((k == 0 || k == NBINS - 1) ? 1.0 : 2.0) / 96.0;   // one-sided density, sum(hann^2) = 96
  pxx[static_cast<int64_t>(blockIdx.x) * NBINS + k] = static_cast<float>(s * scale / static_cast<double>(nseg));
}

int blocks_per_signal(int64_t n_signals, int64_t nseg, int sms) {
  const int64_t npairs = (nseg + 1) / 2;
  int64_t want = cdiv64(static_cast<int64_t>(sms) * 16, n_signals);   // ~16 resident blocks (of 128 threads) per SM over all signals
  if (want > npairs) want = npairs;
  if (want > MAX_BLOCKS_PER_SIGNAL) want = MAX_BLOCKS_PER_SIGNAL;
  if (want < 1) want = 1;
  return static_cast<int>(want);
}

}  // namespace
}  // namespace b200dn

extern "C" int64_t b200dn_welch_psd_workspace_bytes(int64_t n_signals, int64_t n) {
  using namespace b200dn;
  if (n_signals <= 0 || n < NSEG) return 0;
  return n_signals * MAX_BLOCKS_PER_SIGNAL * NBINS * static_cast<int64_t>(sizeof(float));
}

extern "C" int b200dn_welch_psd(const float* x, int64_t n_signals, int64_t n, float* pxx, void* workspace,
                                int64_t workspace_bytes, void* stream) {
  using namespace b200dn;
  B200DN_CHECK_ARG(x && pxx && n_signals > 0, "welch_psd: bad arguments");
  B200DN_CHECK_ARG(n >= NSEG, "welch_psd: signals shorter than nperseg = 256 are not supported (got %lld)", (long long)n);
  B200DN_CHECK_ARG(n_signals <= 65535, "welch_psd: at most 65535 signals per call");
  B200DN_CHECK_ARG(workspace && (reinterpret_cast<uintptr_t>(workspace) & 3) == 0 &&
                       workspace_bytes >= b200dn_welch_psd_workspace_bytes(n_signals, n),
                   "welch_psd: workspace of b200dn_welch_psd_workspace_bytes() bytes required");
  if (int rc = require_sm100()) return rc;
  int sms = device_sm_count();
  if (sms <= 0) return B200DN_E_CUDA;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const int64_t nseg = (n - (NSEG - HOP)) / HOP;
  const int nb = blocks_per_signal(n_signals, nseg, sms);
  float* partial = static_cast<float*>(workspace);
  dim3 grid(static_cast<unsigned>(nb), static_cast<unsigned>(n_signals));
  welch_kernel<<<grid, WT, 0, s>>>(x, n, nseg, partial);
  B200DN_CUDA(cudaGetLastError());
  welch_finish_kernel<<<static_cast<unsigned>(n_signals), 160, 0, s>>>(partial, nb, nseg, pxx);
  B200DN_CUDA(cudaGetLastError());
  return 0;
}
