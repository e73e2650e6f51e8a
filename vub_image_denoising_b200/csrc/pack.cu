// pack.cu — state_dict weight layout -> tensor-core operand layout.
// Conv2d weights are OIHW fp32 [Cout,Cin,kh,kw] (UNet/RDUNet_model.py:52,61,74-75,86-87,98-101);
// ConvTranspose2d weights are IOHW fp32 [Cin,Cout,2,2] (UNet/RDUNet_model.py:62).
// Packed: [wplane][group][cout_pad][cin_pad] 16-bit, cin contiguous (K-major B operand of the UMMA),
// zero padded so that padded K columns / N rows contribute exact zeros.
#include "common.cuh"

namespace b200dn {

namespace {

__device__ __forceinline__ uint16_t to16(float v, int prec) {
  if (prec == B200DN_PREC_FP16 || prec == B200DN_PREC_FP16X2) return __half_as_ushort(__float2half_rn(v));
  return __bfloat16_as_ushort(__float2bfloat16_rn(v));
}

// transposed = 0: src[o][i][g] with g = ky*kw+kx        (Conv2d)
// transposed = 1: src[i][o][g] with g = ky*2+kx         (ConvTranspose2d)
__global__ void pack_kernel(const float* __restrict__ src, int cout, int cin, int groups, int cout_pad, int cin_pad,
                            int prec, int transposed, uint16_t* __restrict__ dst) {
  const int64_t per_plane = static_cast<int64_t>(groups) * cout_pad * cin_pad;
  const int64_t idx = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (idx >= per_plane) return;
  const int i = static_cast<int>(idx % cin_pad);
  const int o = static_cast<int>((idx / cin_pad) % cout_pad);
  const int g = static_cast<int>(idx / (static_cast<int64_t>(cin_pad) * cout_pad));
  float w = 0.f;
  if (i < cin && o < cout) {
    const int64_t s = transposed ? (static_cast<int64_t>(i) * cout + o) * groups + g
                                 : (static_cast<int64_t>(o) * cin + i) * groups + g;
    w = src[s];
  }
  const uint16_t hi = to16(w, prec);
  dst[idx] = hi;
  if (prec == B200DN_PREC_BF16X3) {
    const float r = w - __bfloat162float(__ushort_as_bfloat16(hi));
    dst[per_plane + idx] = __bfloat16_as_ushort(__float2bfloat16_rn(r));
  }
}

int pack(const float* src, int cout, int cin, int groups, int prec, int transposed, void* packed, cudaStream_t stream) {
  B200DN_CHECK_ARG(src && packed, "pack: null pointer");
  B200DN_CHECK_ARG(cout > 0 && cin > 0 && groups > 0, "pack: non-positive dims");
  B200DN_CHECK_ARG(prec >= 0 && prec <= 4, "pack: bad prec %d", prec);
  const int cin_pad = round_up(cin, 64), cout_pad = round_up(cout, 16);
  const int64_t n = static_cast<int64_t>(groups) * cout_pad * cin_pad;
  const int threads = 256;
  const int64_t blocks = cdiv64(n, threads);
  pack_kernel<<<static_cast<unsigned>(blocks), threads, 0, stream>>>(src, cout, cin, groups, cout_pad, cin_pad, prec,
                                                                     transposed, static_cast<uint16_t*>(packed));
  B200DN_CUDA(cudaGetLastError());
  return 0;
}

}  // namespace
}  // namespace b200dn

extern "C" int64_t b200dn_packed_weight_bytes(int cout, int cin, int groups, int prec) {
  if (cout <= 0 || cin <= 0 || groups <= 0) return B200DN_E_ARG;
  const int64_t planes = (prec == B200DN_PREC_BF16X3) ? 2 : 1;
  return planes * groups * static_cast<int64_t>(b200dn::round_up(cout, 16)) * b200dn::round_up(cin, 64) * 2;
}

extern "C" int b200dn_pack_conv_weight(const float* w, int cout, int cin, int kh, int kw, int prec, void* packed,
                                       void* stream) {
  return b200dn::pack(w, cout, cin, kh * kw, prec, 0, packed, static_cast<cudaStream_t>(stream));
}

extern "C" int b200dn_pack_convt_weight(const float* w, int cin, int cout, int prec, void* packed, void* stream) {
  return b200dn::pack(w, cout, cin, 4, prec, 1, packed, static_cast<cudaStream_t>(stream));
}
