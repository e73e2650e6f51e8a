// conv_in.cu — InputBlock.conv_1 + actv_1 (UNet/RDUNet_model.py:71-81) fused with the module-boundary
// ingest: reads the caller's fp32 NCHW image (and, for RDUNet_T, the broadcast timestep plane that the
// reference concatenates at diffusion_denoising/Unet/Unet_model.py:135-136), writes NHWC 16-bit planes.
// K = 27 or 36 is far too small for the tensor pipe and the layer is bandwidth-bound (AI ~ 26 FLOP/B),
// so this runs on CUDA cores in fp32.
#include "common.cuh"

namespace b200dn {

namespace {

constexpr int PX = 32;     // pixels along W per block (one warp = 32 consecutive pixels: coalesced NCHW reads)
constexpr int ROWS = 8;    // warps per block; each thread owns TWO vertically adjacent pixels per pass
constexpr int PASSES = 4;  // a block covers PX x (2 * ROWS * PASSES) pixels, so the weights are staged once per 2048 px

__device__ __forceinline__ uint32_t pack_f16x2_sat(float lo, float hi) {   // saturates at +-65504 like the igemm epilogue
  uint32_t r;
  asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
  return r;
}
__device__ __forceinline__ uint32_t sat_track(uint32_t m, uint32_t h) {
  uint32_t r;
  asm("{\n\t.reg .b32 a;\n\tabs.f16x2 a, %2;\n\tmax.NaN.f16x2 %0, %1, a;\n\t}" : "=r"(r) : "r"(m), "r"(h));
  return r;
}

template <bool kBf16>
__device__ __forceinline__ void store_group(const float (&acc)[8], const float* b_s, const float* s_s, int g,
                                            uint16_t* o0, uint16_t* o1, uint32_t& satm) {
  uint32_t hi[4], lo[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    float a = acc[2 * j] + b_s[g * 8 + 2 * j];
    float c = acc[2 * j + 1] + b_s[g * 8 + 2 * j + 1];
    a = a > 0.f ? a : a * s_s[g * 8 + 2 * j];
    c = c > 0.f ? c : c * s_s[g * 8 + 2 * j + 1];
    if (kBf16) {
      hi[j] = pack_bf16x2(a, c);
      lo[j] = pack_bf16x2(a - bf16_lo(hi[j]), c - bf16_hi(hi[j]));
    } else {
      hi[j] = pack_f16x2_sat(a, c);
      lo[j] = pack_f16x2_sat(a - f16_lo(hi[j]), c - f16_hi(hi[j]));
      satm = sat_track(satm, hi[j]);
    }
  }
  *reinterpret_cast<uint4*>(o0 + g * 8) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
  if (o1 != nullptr) *reinterpret_cast<uint4*>(o1 + g * 8) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
}

// One thread = two vertically adjacent pixels, all output channels in groups of 8.  The weights are broadcast from
// shared memory (2 x LDS.128 per k and group); sharing them between two pixels halves the shared-memory reads per FMA
// (the one-pixel version issued one LDS.128 per 4 FFMA and ran at a third of the FP32 rate), and the two pixels share
// 8 of their 12 input rows x columns.
// CIMG = channels of the image (3 RGB, 1 grayscale); HAS_T = the RDUNet_T timestep plane as an extra input channel
template <int CIMG, bool HAS_T>
__global__ void __launch_bounds__(PX * ROWS) conv_in_kernel(const float* __restrict__ x, int Bx,
                                                            const float* __restrict__ t, int64_t t_sb, int64_t t_sh,
                                                            int64_t t_sw, int H, int W, int cout,
                                                            const float* __restrict__ w, const float* __restrict__ bias,
                                                            const float* __restrict__ slope, int prec,
                                                            uint16_t* __restrict__ out0, uint16_t* __restrict__ out1,
                                                            int out_ctot, int* __restrict__ sat_flag) {
  constexpr int CIN = CIMG + (HAS_T ? 1 : 0);
  extern __shared__ float w_s[];  // [CIN*9][cout], then bias[cout], slope[cout]
  constexpr int K = CIN * 9;
  float* b_s = w_s + K * cout;
  float* s_s = b_s + cout;
  const int tid = threadIdx.y * PX + threadIdx.x;
  for (int i = tid; i < K * cout; i += PX * ROWS) {
    const int o = i / K, k = i - o * K;  // w is [o][ci][ky][kx] = [o][k]
    w_s[k * cout + o] = w[i];
  }
  for (int i = tid; i < cout; i += PX * ROWS) {
    b_s[i] = bias[i];
    s_s[i] = slope[i];
  }
  __syncthreads();

  const int b = blockIdx.z;
  const int xx = blockIdx.x * PX + threadIdx.x;
  const int64_t hw = static_cast<int64_t>(H) * W;
  const float* xb = x + static_cast<int64_t>(b % Bx) * CIMG * hw;
  const bool is_bf16 = (prec != B200DN_PREC_FP16) && (prec != B200DN_PREC_FP16X2);
  if (xx >= W) return;
  uint32_t satm = 0;

  for (int pass = 0; pass < PASSES; ++pass) {
    const int y = ((blockIdx.y * PASSES + pass) * ROWS + threadIdx.y) * 2;   // rows y and y + 1
    if (y >= H) break;
    const bool two = (y + 1) < H;
    float v[CIN][4][3];   // input rows y-1 .. y+2, columns xx-1 .. xx+1
#pragma unroll
    for (int r = 0; r < 4; ++r) {
#pragma unroll
      for (int kx = 0; kx < 3; ++kx) {
        const int yy = y + r - 1, xc = xx + kx - 1;
        const bool in = (yy >= 0) && (yy < H) && (xc >= 0) && (xc < W);
        const int64_t sp = static_cast<int64_t>(yy) * W + xc;
#pragma unroll
        for (int ci = 0; ci < CIMG; ++ci) v[ci][r][kx] = in ? __ldg(xb + ci * hw + sp) : 0.f;
        if (HAS_T) v[CIMG][r][kx] = in ? __ldg(t + b * t_sb + yy * t_sh + xc * t_sw) : 0.f;
      }
    }
    const int64_t pix = (static_cast<int64_t>(b) * H + y) * W + xx;
    uint16_t* o0 = out0 + pix * out_ctot;
    uint16_t* o1 = out1 ? out1 + pix * out_ctot : nullptr;
    const int64_t row = static_cast<int64_t>(W) * out_ctot;
    for (int g = 0; g * 8 < cout; ++g) {
      // packed fp32 FMAs (FFMA2: two IEEE fmas per instruction, same bits as fmaf) — the kernel is FP32-issue-bound
      float2 p0[4], p1[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) p0[j] = p1[j] = make_float2(0.f, 0.f);
#pragma unroll
      for (int ci = 0; ci < CIN; ++ci) {
#pragma unroll
        for (int ky = 0; ky < 3; ++ky) {
#pragma unroll
          for (int kx = 0; kx < 3; ++kx) {
            const int k = ci * 9 + ky * 3 + kx;
            const float4 w0 = *reinterpret_cast<const float4*>(w_s + k * cout + g * 8);
            const float4 w1 = *reinterpret_cast<const float4*>(w_s + k * cout + g * 8 + 4);
            const float2 a = make_float2(v[ci][ky][kx], v[ci][ky][kx]);
            const float2 c = make_float2(v[ci][ky + 1][kx], v[ci][ky + 1][kx]);
            const float2 wa = make_float2(w0.x, w0.y), wb = make_float2(w0.z, w0.w);
            const float2 wc = make_float2(w1.x, w1.y), wd = make_float2(w1.z, w1.w);
            p0[0] = __ffma2_rn(a, wa, p0[0]);
            p0[1] = __ffma2_rn(a, wb, p0[1]);
            p0[2] = __ffma2_rn(a, wc, p0[2]);
            p0[3] = __ffma2_rn(a, wd, p0[3]);
            p1[0] = __ffma2_rn(c, wa, p1[0]);
            p1[1] = __ffma2_rn(c, wb, p1[1]);
            p1[2] = __ffma2_rn(c, wc, p1[2]);
            p1[3] = __ffma2_rn(c, wd, p1[3]);
          }
        }
      }
      float acc0[8], acc1[8];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        acc0[2 * j] = p0[j].x, acc0[2 * j + 1] = p0[j].y;
        acc1[2 * j] = p1[j].x, acc1[2 * j + 1] = p1[j].y;
      }
      if (is_bf16) {
        store_group<true>(acc0, b_s, s_s, g, o0, o1, satm);
        if (two) store_group<true>(acc1, b_s, s_s, g, o0 + row, o1 ? o1 + row : nullptr, satm);
      } else {
        store_group<false>(acc0, b_s, s_s, g, o0, o1, satm);
        if (two) store_group<false>(acc1, b_s, s_s, g, o0 + row, o1 ? o1 + row : nullptr, satm);
      }
    }
  }
  if (sat_flag != nullptr && (((satm & 0x7fffu) >= 0x7bffu) || ((satm >> 16) >= 0x7bffu))) atomicOr(sat_flag, 1);
}

}  // namespace
}  // namespace b200dn

namespace b200dn {
namespace {
template <int CIMG, bool HAS_T>
int launch_conv_in(dim3 grid, dim3 block, size_t smem, cudaStream_t s, const float* x, int Bx, const float* t, int64_t t_sb,
                   int64_t t_sh, int64_t t_sw, int H, int W, int cout, const float* w, const float* bias, const float* slope,
                   int prec, uint16_t* o0, uint16_t* o1, int out_ctot, int* sat_flag) {
  if (smem > 48 * 1024)
    B200DN_CUDA(cudaFuncSetAttribute(conv_in_kernel<CIMG, HAS_T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  conv_in_kernel<CIMG, HAS_T><<<grid, block, smem, s>>>(x, Bx, t, t_sb, t_sh, t_sw, H, W, cout, w, bias, slope, prec, o0, o1,
                                                         out_ctot, sat_flag);
  B200DN_CUDA(cudaGetLastError());
  return 0;
}
}  // namespace
}  // namespace b200dn

extern "C" int b200dn_conv_in(const float* x, int Bx, int img_channels, const float* t, int64_t t_sb, int64_t t_sh,
                              int64_t t_sw, int B, int H, int W, int cout, const float* w, const float* bias,
                              const float* slope, int prec, void* out0, void* out1, int out_ctot, int32_t* sat_flag,
                              void* stream) {
  using namespace b200dn;
  B200DN_CHECK_ARG(x && w && bias && slope && out0, "conv_in: null pointer");
  B200DN_CHECK_ARG(B > 0 && Bx > 0 && H > 0 && W > 0, "conv_in: non-positive dims");
  B200DN_CHECK_ARG(img_channels == 3 || img_channels == 1, "conv_in: image channels must be 3 (RGB) or 1 (grayscale), got %d",
                   img_channels);
  B200DN_CHECK_ARG(cout > 0 && cout % 8 == 0, "conv_in: cout %d must be a multiple of 8", cout);
  B200DN_CHECK_ARG(out_ctot % 8 == 0 && out_ctot >= cout, "conv_in: out_ctot %d invalid", out_ctot);
  B200DN_CHECK_ARG(prec >= 0 && prec <= 4, "conv_in: bad prec %d", prec);
  const bool two = (prec == B200DN_PREC_BF16X2 || prec == B200DN_PREC_BF16X3 || prec == B200DN_PREC_FP16X2);
  B200DN_CHECK_ARG(!two || out1, "conv_in: prec %d needs the lo output plane", prec);
  B200DN_CHECK_ARG(B <= 65535 && H <= 65535 * 2 * ROWS * PASSES, "conv_in: B/H exceed the grid limit");
  if (int rc = require_sm100()) return rc;
  const int cin = img_channels + (t ? 1 : 0);
  const size_t smem = (static_cast<size_t>(cin) * 9 + 2) * cout * sizeof(float);
  B200DN_CHECK_ARG(smem <= 160 * 1024, "conv_in: cout %d too large", cout);
  dim3 block(PX, ROWS), grid(cdiv(W, PX), cdiv(H, 2 * ROWS * PASSES), B);
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  uint16_t* o0 = static_cast<uint16_t*>(out0);
  uint16_t* o1 = two ? static_cast<uint16_t*>(out1) : nullptr;
  if (img_channels == 3)
    return t ? launch_conv_in<3, true>(grid, block, smem, s, x, Bx, t, t_sb, t_sh, t_sw, H, W, cout, w, bias, slope, prec, o0, o1, out_ctot, sat_flag)
             : launch_conv_in<3, false>(grid, block, smem, s, x, Bx, nullptr, 0, 0, 0, H, W, cout, w, bias, slope, prec, o0, o1, out_ctot, sat_flag);
  return t ? launch_conv_in<1, true>(grid, block, smem, s, x, Bx, t, t_sb, t_sh, t_sw, H, W, cout, w, bias, slope, prec, o0, o1, out_ctot, sat_flag)
           : launch_conv_in<1, false>(grid, block, smem, s, x, Bx, nullptr, 0, 0, 0, H, W, cout, w, bias, slope, prec, o0, o1, out_ctot, sat_flag);
}
