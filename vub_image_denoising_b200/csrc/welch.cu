// welch.cu — Welch power-spectral-density estimate of flattened images on the device.
//
// Reference call: `welch(image.flatten(), nperseg=256)` (scipy.signal), evaluate_Unet_diffusion/plot.py:155-157,
// 233-235, 287-290 — the remaining per-image host cost of evaluate_model.py once PSNR/SSIM are on the device
// (SURVEY.md §8 row f4).  scipy defaults reproduced here: fs = 1, periodic Hann window, noverlap = 128, nfft = 256,
// detrend = 'constant' (each segment's mean removed), one-sided density scaling (x2 except DC / Nyquist,
// 1 / (fs * sum(w^2)) = 1/96), mean over segments; float32 input -> complex64 transform -> float32 spectrum.
//
// One HALF-WARP = one 256-point complex FFT carrying TWO real segments (z = a + i b; A[k] = (Z[k] + conj Z[N-k]) / 2,
// B[k] = (Z[k] - conj Z[N-k]) / 2i), as 16 x 16 Cooley-Tukey with the data in registers: thread t holds z[16 n1 + t],
// does a 16-point radix-2 DIF over n1 (compile-time twiddles), multiplies by W256^(t k1), transposes through a
// per-half-warp shared-memory tile (__syncwarp only) and does the second 16-point FFT over n2; thread k1 then holds the
// bins k1 + 16 k2 and fetches the mirrored bins from lane (16 - k1) & 15 with shuffles.  A block (8 half-warps) walks a
// strided share of one signal's segment pairs, adds its half-warps' spectra in a fixed order and writes one partial; a
// second kernel adds the partials in a fixed order (bit-reproducible, no atomics).
// The first version ran one block-wide radix-2 Stockham FFT per pair (eight __syncthreads per transform): 0.38 TB/s.
#include "common.cuh"

namespace b200dn {

namespace {

constexpr int NSEG = 256;            // nperseg = nfft
constexpr int HOP = 128;             // nperseg - noverlap
constexpr int NBINS = NSEG / 2 + 1;  // 129
constexpr int WT = 128;              // threads per block = 8 half-warps
constexpr int HW_PER_BLOCK = WT / 16;
constexpr int MAX_BLOCKS_PER_SIGNAL = 64;
constexpr int TP = 17;               // transpose tile pitch (float2), conflict-free for 8-byte accesses of a half-warp

__device__ __forceinline__ float2 cmul(float2 a, float2 b) { return make_float2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x); }
__device__ __forceinline__ float2 cadd(float2 a, float2 b) { return make_float2(a.x + b.x, a.y + b.y); }
__device__ __forceinline__ float2 csub(float2 a, float2 b) { return make_float2(a.x - b.x, a.y - b.y); }

// W16^j = exp(-2 pi i j / 16), j = 0..7
__device__ __forceinline__ float2 w16(int j) {
  constexpr float c1 = 0.92387953251128674f, s1 = 0.38268343236508977f, r = 0.70710678118654752f;
  switch (j) {
    case 0: return make_float2(1.f, 0.f);
    case 1: return make_float2(c1, -s1);
    case 2: return make_float2(r, -r);
    case 3: return make_float2(s1, -c1);
    case 4: return make_float2(0.f, -1.f);
    case 5: return make_float2(-s1, -c1);
    case 6: return make_float2(-r, -r);
    default: return make_float2(-c1, -s1);
  }
}
__device__ __forceinline__ constexpr int bitrev4(int i) { return ((i & 1) << 3) | ((i & 2) << 1) | ((i & 4) >> 1) | ((i & 8) >> 3); }

// in-register 16-point radix-2 DIF: afterwards a[i] = X[bitrev4(i)]
__device__ __forceinline__ void fft16(float2 (&a)[16]) {
#pragma unroll
  for (int h = 8; h >= 1; h >>= 1) {
#pragma unroll
    for (int b = 0; b < 16; b += 2 * h) {
#pragma unroll
      for (int j = 0; j < h; ++j) {
        const float2 u = a[b + j], v = a[b + j + h];
        a[b + j] = cadd(u, v);
        const float2 d = csub(u, v);
        a[b + j + h] = (j == 0) ? d : cmul(d, w16(j * (8 / h)));
      }
    }
  }
}

__global__ void __launch_bounds__(WT) welch_kernel(const float* __restrict__ x, int64_t n, int64_t nseg,
                                                   float* __restrict__ partial) {
  __shared__ float2 tile[HW_PER_BLOCK][16 * TP];
  __shared__ float2 tw[NSEG];
  __shared__ float win[NSEG];
  __shared__ float accs[HW_PER_BLOCK][NBINS + 3];
  const int tid = threadIdx.x;
  const float* sig = x + static_cast<int64_t>(blockIdx.y) * n;
  for (int i = tid; i < NSEG; i += WT) {
    float s, c;
    sincospif(-static_cast<float>(i) / 128.0f, &s, &c);   // exp(-2 pi i k / 256)
    tw[i] = make_float2(c, s);
    win[i] = 0.5f - 0.5f * cospif(static_cast<float>(i) / 128.0f);   // periodic Hann
  }
  __syncthreads();
  const int hw = tid >> 4, t = tid & 15;
  const unsigned hmask = 0xffffu << (tid & 16);           // this half-warp's lanes
  float2* my = tile[hw];
  float acc[8];
#pragma unroll
  for (int k2 = 0; k2 < 8; ++k2) acc[k2] = 0.f;
  float acc_nyq = 0.f;
  const int64_t npairs = (nseg + 1) / 2;
  const int64_t stride = static_cast<int64_t>(gridDim.x) * HW_PER_BLOCK;
  for (int64_t pair = static_cast<int64_t>(blockIdx.x) * HW_PER_BLOCK + hw; pair < npairs; pair += stride) {
    const int64_t s0 = 2 * pair;
    const bool has_b = s0 + 1 < nseg;
    const float* pa = sig + s0 * HOP;
    // segment a = samples [0, 256), segment b = [128, 384) of the same window; thread t takes samples 16 n1 + t
    float va[16], vb[16];
    float sa = 0.f, sb = 0.f;
#pragma unroll
    for (int n1 = 0; n1 < 16; ++n1) {
      va[n1] = pa[16 * n1 + t];
      vb[n1] = has_b ? pa[HOP + 16 * n1 + t] : 0.f;
      sa += va[n1];
      sb += vb[n1];
    }
#pragma unroll
    for (int o = 8; o > 0; o >>= 1) {
      sa += __shfl_xor_sync(hmask, sa, o);
      sb += __shfl_xor_sync(hmask, sb, o);
    }
    const float mean_a = sa * (1.0f / NSEG), mean_b = sb * (1.0f / NSEG);
    float2 a[16];
#pragma unroll
    for (int n1 = 0; n1 < 16; ++n1) {
      const float w = win[16 * n1 + t];
      a[n1] = make_float2((va[n1] - mean_a) * w, has_b ? (vb[n1] - mean_b) * w : 0.f);
    }
    fft16(a);                                   // over n1: a[bitrev4(k1)] = Y_t[k1]
    __syncwarp(hmask);                          // the previous pair's reads of the tile are done
#pragma unroll
    for (int k1 = 0; k1 < 16; ++k1) my[t * TP + k1] = cmul(a[bitrev4(k1)], tw[(t * k1) & (NSEG - 1)]);
    __syncwarp(hmask);
#pragma unroll
    for (int n2 = 0; n2 < 16; ++n2) a[n2] = my[n2 * TP + t];       // thread t = k1 now
    fft16(a);                                   // over n2: a[bitrev4(k2)] = Z[k1 + 16 k2]
    // power of both real segments at k = k1 + 16 k2, k2 = 0..7; mirrored bin 256 - k lives in lane (16 - k1) & 15
    const int partner = ((16 - t) & 15) | (tid & 16);
#pragma unroll
    for (int k2 = 0; k2 < 8; ++k2) {
      const float2 zk = a[bitrev4(k2)];
      const float2 own = a[bitrev4((16 - k2) & 15)];                // lane 0: 256 - 16 k2 = 16 (16 - k2)
      const float2 give = a[bitrev4(15 - k2)];                      // what this lane holds for its partner
      float2 zn;
      zn.x = __shfl_sync(hmask, give.x, partner);
      zn.y = __shfl_sync(hmask, give.y, partner);
      if (t == 0) zn = own;
      // A = (zk + conj zn) / 2, B = (zk - conj zn) / (2i)
      const float ar = 0.5f * (zk.x + zn.x), ai = 0.5f * (zk.y - zn.y);
      const float br = 0.5f * (zk.y + zn.y), bi = -0.5f * (zk.x - zn.x);
      acc[k2] += (ar * ar + ai * ai) + (br * br + bi * bi);
    }
    if (t == 0) {
      const float2 zq = a[bitrev4(8)];   // Nyquist bin 128: A = Re, B = Im
      acc_nyq += zq.x * zq.x + zq.y * zq.y;
    }
  }
  // fixed-order sum of the block's eight half-warp spectra -> one partial per block
#pragma unroll
  for (int k2 = 0; k2 < 8; ++k2) accs[hw][t + 16 * k2] = acc[k2];
  if (t == 0) accs[hw][128] = acc_nyq;
  __syncthreads();
  float* p = partial + (static_cast<int64_t>(blockIdx.y) * gridDim.x + blockIdx.x) * NBINS;
  for (int k = tid; k < NBINS; k += WT) {
    float s = 0.f;
#pragma unroll
    for (int h = 0; h < HW_PER_BLOCK; ++h) s += accs[h][k];
    p[k] = s;
  }
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
  const int64_t max_useful = cdiv64(npairs, HW_PER_BLOCK);     // a block runs eight transforms at a time
  if (want > max_useful) want = max_useful;
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
