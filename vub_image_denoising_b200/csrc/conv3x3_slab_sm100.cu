// conv3x3_slab_sm100.cu — 3x3 / stride 1 / pad 1 convolution as a tcgen05 implicit GEMM that loads every input
// pixel into shared memory ONCE per 64-channel block instead of once per filter tap.
//
// Reference op: nn.Conv2d(.,.,3,padding=1) + nn.PReLU (+ dense-block residual / output-block residual),
// UNet/RDUNet_model.py:61,74-75,86-87,98-115,186 — 63 of the network's 69 convolutions, 97.5 % of its FLOPs.
//
// Why: the per-tap kernel (igemm_sm100.cu) is bound by shared-memory bandwidth — every pipeline stage writes
// an A tile (TMA) that is read by only four UMMAs, and TMA writes + UMMA operand reads share ~128 B/clk/SM.
// Here one TMA box brings a haloed slab [ (16*MT+2) rows x 16 px x 64 ch ] (zero-filled outside the image =
// the conv's padding); the 9 taps are 9 UMMA descriptors into that slab, shifted by (dy*16 + dx) rows of 128 B.
// The slab pitch is 16 pixels, so the 8-row core-matrix groups of a tap sit 2048 B apart (SBO).  Measured on
// B200 (tools/slab_probe.py): the 128B-swizzle XOR of both TMA and UMMA is a function of the ABSOLUTE shared
// memory address bits [7,10), so a descriptor whose start address is shifted by whole 128 B rows reads the
// TMA-written slab correctly with the descriptor's base-offset field left at 0 (setting it to
// (addr >> 7) & 7 produces garbage).
// A-side smem write traffic drops 4x, A TMA instructions 9x, L2->SM traffic ~4x.
//
// Warp roles: 0 = slab TMA producer, 2 = TMEM allocator then W-tile TMA producer (one [N x 64] tile per tap),
// 1 / 3 = MMA issuers of sub-tile 0 / 1 (MT = 2 stacks two 8x16-pixel tiles vertically), 4..7 = epilogue.
#include "igemm_common.cuh"


namespace b200dn {
namespace igemm {

namespace {

constexpr int TW = SLAB_TILE_W;   // 8
constexpr int TH = SLAB_TILE_H;   // 16
constexpr int MAX_SLABS = SLAB_MAX_SLABS;
constexpr int DATA_BYTES = SLAB_DATA_BYTES + EPI_STAGING_BYTES;   // [slabs | W ring or resident W | epilogue staging]
constexpr int SMEM_BYTES_SLAB = 1024 + DATA_BYTES + SLAB_CTRL_BYTES + 2 * 2 * MAX_N * 4;

// barrier map (byte offsets from `bars`)
// barrier map (byte offsets from `bars`): up to 8 slab slots and 8 W stages
constexpr uint32_t B_SLAB_FULL = 0, B_SLAB_EMPTY = 64, B_W_FULL = 128, B_W_EMPTY = 192, B_TFULL = 256, B_TEMPTY = 272,
                   B_TMEM_PTR = 288;

// WRES: the layer's whole packed weight set ([w plane][64-ch block][tap][N x 64]) is loaded into shared memory once
// per CTA and stays resident; the main loop then only streams slabs.  Used for the small layers (F = 32 levels 0/1),
// which are latency-bound: per tile it removes 9 W-tile round trips and keeps the issue loop small enough for the
// instruction cache (ncu: the unrolled 9-tap path showed mostly `no_inst` stalls there).
// Epilogue warp groups: the resident-weight layers (N <= 48, K <= 9 * 128) do so little tensor work per output that
// the epilogue's CUDA-core instruction stream is what paces a tile (measured: time proportional to pixels, independent
// of K, of MT and of DRAM vs L2 residency; 4 epilogue warps run at ~0.27 IPC each).  With MT = 2 a second group of four
// warps (hardware warps 8..11) drains sub-tile 1 while the first drains sub-tile 0.
template <int MT, bool WRES>
constexpr int slab_epi_groups() { return (WRES && MT == 2) ? 2 : 1; }

template <int MT, bool WRES>
__global__ void __launch_bounds__(NUM_THREADS + EPI_THREADS * (slab_epi_groups<MT, WRES>() - 1), 1)
conv3x3_slab_kernel(const __grid_constant__ KParams p) {
  constexpr int EG = slab_epi_groups<MT, WRES>();
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_u32 = smem_u32(smem_raw);
  const uint32_t smem_base = (raw_u32 + 1023u) & ~1023u;
  if (threadIdx.x == 0) TL_MARK(p, TL_ENTRY);
  uint8_t* smem_gen = smem_raw + (smem_base - raw_u32);
  const uint32_t bars = smem_base + DATA_BYTES;
  uint32_t* tmem_ptr_s = reinterpret_cast<uint32_t*>(smem_gen + DATA_BYTES + B_TMEM_PTR);
  float* epi_bias = reinterpret_cast<float*>(smem_gen + DATA_BYTES + SLAB_CTRL_BYTES);  // [2][MAX_N]
  float* epi_slope = epi_bias + 2 * MAX_N;

  // Role index: hardware warps 4..7 run the single-thread producer / MMA-issue loops, warps 0..3 the epilogue.
  // The SMSP arbiter favours the higher warp id, and each issuer shares its SMSP with one epilogue warp, so this
  // order keeps the ALU-heavy epilogue from starving the issue loops that feed the tensor pipe.
  const int warp = (threadIdx.x >> 5) ^ 4;
  const int lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&p.tmA0);
    tma_prefetch_desc(&p.tmW);
    if (p.n_pairs > 1) tma_prefetch_desc(&p.tmA1);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < MAX_SLABS; ++s) {
      mbar_init(bars + B_SLAB_FULL + s * 8, 1);
      mbar_init(bars + B_SLAB_EMPTY + s * 8, MT);
    }
    for (int s = 0; s < MAX_STAGES; ++s) {
      mbar_init(bars + B_W_FULL + s * 8, 1);
      mbar_init(bars + B_W_EMPTY + s * 8, MT);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(bars + B_TFULL + a * 8, MT);
      mbar_init(bars + B_TEMPTY + a * 8, (EPI_THREADS / 32) * EG);   // one arrive per epilogue warp
    }
    mbar_fence_init();
  }
  if (warp == 2) {
    tmem_alloc(smem_u32(tmem_ptr_s), static_cast<uint32_t>(p.tmem_cols));
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_s;
  if (threadIdx.x == 0) TL_MARK(p, TL_SETUP);
  // PDL: everything above (barrier init, TMEM allocation, descriptor prefetch) may overlap the previous kernel of
  // the stream; let the next kernel start its own prologue on idle SMs, then wait for our producer to finish before
  // any activation memory is touched (packed weights / bias are static and need no wait).
  griddep_launch_dependents();
  if (warp != 2) griddep_wait();   // the W producer (resident or streamed) only reads static data
  if (threadIdx.x == 0) TL_MARK(p, TL_DEP);

  const int num_tiles = p.num_tiles, grid = gridDim.x;
  const int num_n_tiles = p.num_n_tiles, n_tiles_per_group = p.n_tiles_per_group, block_n = p.block_n;
  const int tiles_x = p.tiles_x, tiles_y = p.tiles_y;
  const int n_cblk = p.n_cblk, n_pairs = p.n_pairs;
  const int num_slabs = p.num_slabs, slab_bytes = p.slab_bytes;
  const int num_stages = p.num_stages, stage_bytes = p.stage_bytes;
  const uint32_t wring = smem_base + static_cast<uint32_t>(num_slabs * slab_bytes);
  constexpr int STH = TH * MT;

  if (warp == 0) {
    // ===================================================== slab TMA producer: one box per (pair, 64-channel block)
    int s = 0;
    uint32_t sph = 0;
    const int pa0 = p.pair_a[0], pa1 = p.pair_a[1], pa2 = p.pair_a[2];
    TileWalker tw_;
    tw_.init(blockIdx.x, grid, num_n_tiles, tiles_x, tiles_y);
    for (int tile = blockIdx.x; tile < num_tiles; tile += grid, tw_.next()) {
      const TileCoord t = tw_.coord(block_n, TW, STH);
      for (int pair = 0; pair < n_pairs; ++pair) {
        const int pa = pair == 0 ? pa0 : pair == 1 ? pa1 : pa2;
        const CUtensorMap* tmA = pa ? &p.tmA1 : &p.tmA0;
        for (int cb = 0; cb < n_cblk; ++cb) {
          mbar_wait(bars + B_SLAB_EMPTY + s * 8, sph ^ 1u);
          if (elect_one()) {
            const uint32_t full = bars + B_SLAB_FULL + s * 8;
            mbar_arrive_expect_tx(full, static_cast<uint32_t>(p.slab_tx));
            tma_load_4d(smem_base + s * slab_bytes, tmA, full, cb * BLOCK_K, t.x0 - 1, t.y0 - 1, t.b);
          }
          __syncwarp();
          if (++s == num_slabs) {
            s = 0;
            sph ^= 1u;
          }
        }
      }
    }
  } else if (warp == 2) {
    if (WRES) {
      // ===================================================== resident weights: 9 taps per box, loaded once
      if (elect_one()) {
        const int n_boxes = p.n_wplanes * n_cblk;
        const uint32_t box_bytes = static_cast<uint32_t>(9 * block_n * 128);
        const uint32_t full = bars + B_W_FULL;
        mbar_arrive_expect_tx(full, static_cast<uint32_t>(n_boxes) * box_bytes);
        for (int wp = 0; wp < p.n_wplanes; ++wp)
          for (int cb = 0; cb < n_cblk; ++cb)
            tma_load_3d(wring + static_cast<uint32_t>(wp * n_cblk + cb) * box_bytes, &p.tmW, full, cb * BLOCK_K, 0,
                        wp * 9);
      }
      __syncwarp();
      griddep_wait();
    } else {
    // ===================================================== W-tile TMA producer: one [N x 64] tile per tap
    int ws = 0;
    uint32_t wph = 0;
    const int w_taps = p.w_taps;
    const uint32_t w_bytes = static_cast<uint32_t>(block_n * 128 * w_taps);   // one box = w_taps tap tiles
    const int pw0 = p.pair_w[0] * 9, pw1 = p.pair_w[1] * 9, pw2 = p.pair_w[2] * 9;
    TileWalker tw_;
    tw_.init(blockIdx.x, grid, num_n_tiles, tiles_x, tiles_y);
    for (int tile = blockIdx.x; tile < num_tiles; tile += grid, tw_.next()) {
      const TileCoord t = tw_.coord(block_n, TW, STH);
      for (int pair = 0; pair < n_pairs; ++pair) {
        const int wbase = pair == 0 ? pw0 : pair == 1 ? pw1 : pw2;
        for (int cb = 0; cb < n_cblk; ++cb) {
#pragma unroll 1
          for (int tap = 0; tap < 9; tap += w_taps) {
            mbar_wait(bars + B_W_EMPTY + ws * 8, wph ^ 1u);
            if (elect_one()) {
              const uint32_t full = bars + B_W_FULL + ws * 8;
              mbar_arrive_expect_tx(full, w_bytes);
              tma_load_3d(wring + ws * stage_bytes, &p.tmW, full, cb * BLOCK_K, t.n0, wbase + tap);
            }
            __syncwarp();
            if (++ws == num_stages) {
              ws = 0;
              wph ^= 1u;
            }
          }
        }
      }
    }
    }
  } else if (warp == 1 || (warp == 3 && MT == 2)) {
    // ===================================================== MMA issuer of sub-tile j
    // The loop is deliberately small: dy rolled, descriptors and barrier addresses advanced incrementally, the k16
    // count of a tap a compile-time constant per code path.  ncu (profiles/r01_n64_issuer_warp_issue_bound_ncu.txt)
    // showed this warp issue-bound — ~130 SASS instructions per tap, never waiting for data — which capped every
    // layer with N <= 64.
    const int j = warp == 1 ? 0 : 1;
    const uint32_t idesc = make_idesc_f16(static_cast<uint32_t>(p.fmt), static_cast<uint32_t>(block_n));
    const int kc_iters = n_pairs * n_cblk;
    const int last_k16 = p.last_k16;
    const uint32_t pitch = static_cast<uint32_t>(p.slab_w * 128);   // bytes per slab pixel row = SBO
    const uint64_t row_step = static_cast<uint64_t>(pitch >> 4);    // descriptor address units (16 B)
    const int pw_0 = p.pair_w[0], pw_1 = p.pair_w[1], pw_2 = p.pair_w[2];
    const uint32_t w_tile_bytes = static_cast<uint32_t>(block_n * 128);
    // streamed-W ring cursor (unused when the weights are resident)
    WRing wr;
    wr.full0 = bars + B_W_FULL, wr.empty0 = bars + B_W_EMPTY;
    wr.full_end = wr.full0 + static_cast<uint32_t>(num_stages) * 8;
    wr.desc0 = make_sw128_desc(wring, 1024);
    wr.step = static_cast<uint64_t>(stage_bytes >> 4);
    wr.full = wr.full0, wr.empty = wr.empty0, wr.phase = 0, wr.desc = wr.desc0;
    const uint64_t bstep = static_cast<uint64_t>(w_tile_bytes >> 4);   // one [N x 64] tap tile
    const int w_taps = p.w_taps;
    if (WRES) mbar_wait(bars + B_W_FULL, 0);   // resident weights have landed
    int s = 0;
    uint32_t sph = 0;
    int local_tile = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += grid, ++local_tile) {
      const int acc = local_tile & 1;
      const uint32_t acc_phase = (local_tile >> 1) & 1;
      mbar_wait(bars + B_TEMPTY + acc * 8, acc_phase ^ 1u);
      tc_fence_after();
      const uint32_t d = tmem_base + static_cast<uint32_t>((acc * MT + j) * block_n);
      uint32_t accumulate = 0;
      int cb = 0;
      for (int kc = 0; kc < kc_iters; ++kc) {
        const int nk = (cb != n_cblk - 1) ? BLOCK_K / 16 : last_k16;
        mbar_wait(bars + B_SLAB_FULL + s * 8, sph);
        tc_fence_after();
        const uint32_t slab = smem_base + static_cast<uint32_t>(s * slab_bytes) + static_cast<uint32_t>(j * TH) * pitch;
        uint64_t arow = make_sw128_desc(slab, pitch);
        if (WRES) {
          // resident weights: [w plane][cb][tap][N x 64]; pair index = kc / n_cblk
          const int pair = kc / n_cblk;
          const int wpl = pair == 0 ? pw_0 : pair == 1 ? pw_1 : pw_2;
          const uint64_t bdesc = make_sw128_desc(wring + static_cast<uint32_t>((wpl * n_cblk + cb) * 9) * w_tile_bytes, 1024);
          if (elect_one()) {
            switch (nk) {   // one dispatch per 9 taps; inside, the k16 count is a compile-time constant
              case 4: issue_slab_block_resident<4>(d, arow, row_step, bdesc, bstep, idesc, accumulate); break;
              case 3: issue_slab_block_resident<3>(d, arow, row_step, bdesc, bstep, idesc, accumulate); break;
              case 2: issue_slab_block_resident<2>(d, arow, row_step, bdesc, bstep, idesc, accumulate); break;
              default: issue_slab_block_resident<1>(d, arow, row_step, bdesc, bstep, idesc, accumulate); break;
            }
            umma_commit(bars + B_SLAB_EMPTY + s * 8);
            if (kc == kc_iters - 1) umma_commit(bars + B_TFULL + acc * 8);
          }
          __syncwarp();
          accumulate = 1;
        } else {
          issue_slab_block_streamed_n<false>(nk, w_taps, d, arow, row_step, wr, bstep, idesc, accumulate);
          if (elect_one()) {
            umma_commit(bars + B_SLAB_EMPTY + s * 8);                         // slab consumed by all 9 taps
            if (kc == kc_iters - 1) umma_commit(bars + B_TFULL + acc * 8);    // accumulator complete
          }
          __syncwarp();
        }
        if (++s == num_slabs) {
          s = 0;
          sph ^= 1u;
        }
        if (++cb == n_cblk) cb = 0;
      }
    }
  } else if (warp >= 4) {
    // ======================================================= epilogue
    const int we = warp & 3;
    const int row = we * 32 + lane;
    const int th = row / TW, tw = row - th * TW;
    const int eg = (EG == 2 && warp >= 8) ? 1 : 0;    // epilogue group (hardware warps 0..3 / 8..11)
    const int et = (threadIdx.x & (EPI_THREADS - 1)) + eg * EPI_THREADS;
    const EpiArgs ea = make_epi_args(p);
    const int H = p.H, W = p.W, cout = p.cout;
    const float* bias = p.bias;
    const float* slope = p.slope;
    const bool staged = !WRES && p.epi_staged;
    const bool nchw_small = p.out_kind == B200DN_OUT_NCHW32 && cout <= 4 && block_n == 16;
    const int res_bmod = p.res_bmod > 0 ? p.res_bmod : 1;
    const int64_t hw = static_cast<int64_t>(H) * W;
    uint8_t* stg = smem_gen + SLAB_DATA_BYTES + we * 4096;
    const bool one_n_tile = p.num_n_tiles == 1;
    if (one_n_tile) stage_bias_slope(epi_bias, epi_slope, bias, slope, 0, block_n, cout, et, EPI_THREADS * EG);
    int local_tile = 0;
    uint32_t satm = 0;
    TileWalker tw_;
    tw_.init(blockIdx.x, grid, num_n_tiles, tiles_x, tiles_y);
    for (int tile = blockIdx.x; tile < num_tiles; tile += grid, ++local_tile, tw_.next()) {
      const TileCoord t = tw_.coord(block_n, TW, STH);
      const int acc = local_tile & 1;
      const uint32_t acc_phase = (local_tile >> 1) & 1;
      // bias / PReLU slopes of the tile's N range: staged once when the layer has a single N tile (per-tile global
      // loads + a named barrier are a large share of a tile of the small-K layers), else per tile, double-buffered
      float* bs = epi_bias + (one_n_tile ? 0 : acc * MAX_N);
      float* ss = epi_slope + (one_n_tile ? 0 : acc * MAX_N);
      if (!one_n_tile) stage_bias_slope(bs, ss, bias, slope, t.n0, block_n, cout, et, EPI_THREADS * EG);

      // Output block (fp32 NCHW, cout <= 4): fetch the fp32 residual (the network input) BEFORE waiting for the
      // accumulator.  ncu showed the direct path serialising one DRAM round trip per channel behind the TMEM read
      // (long_sb on each FADD), 10x above the layer's HBM time.
      float pre[MT][4];
      if (nchw_small) {
        const int rb = t.b % res_bmod;
#pragma unroll
        for (int j = 0; j < MT; ++j) {
          if (EG == 2 && j != eg) continue;
          const int y = t.y0 + j * TH + th, x = t.x0 + tw;
          const bool valid = (y < H) && (x < W);
          const int64_t sp = static_cast<int64_t>(y) * W + x;
#pragma unroll
          for (int c = 0; c < 4; ++c)
            pre[j][c] = (valid && c < cout && p.res_nchw != nullptr)
                            ? __ldg(p.res_nchw + (static_cast<int64_t>(rb) * cout + c) * hw + sp) : 0.f;
        }
      }

      mbar_wait(bars + B_TFULL + acc * 8, acc_phase);
      tc_fence_after();

#pragma unroll
      for (int j = 0; j < MT; ++j) {
        if (EG == 2 && j != eg) continue;   // two epilogue groups: one sub-tile each
        const int y = t.y0 + j * TH + th, x = t.x0 + tw;
        const bool valid = (y < H) && (x < W);
        const int64_t pix = (static_cast<int64_t>(t.b) * H + y) * W + x;
        const uint32_t taddr =
            tmem_base + (static_cast<uint32_t>(we * 32) << 16) + static_cast<uint32_t>((acc * MT + j) * block_n);
        const uint32_t rel = (EG == 2 || j == MT - 1) ? bars + B_TEMPTY + acc * 8 : 0u;
        if (nchw_small) {
          uint32_t r[16];
          tmem_ld16(taddr, r);
          tmem_ld_wait();
          if (rel != 0) {
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(rel);
          }
          if (valid) {
            const int64_t sp = static_cast<int64_t>(y) * W + x;
#pragma unroll
            for (int c = 0; c < 4; ++c) {
              if (c < cout) {
                const float a = __uint_as_float(r[c]) + bs[c];
                p.out_nchw[(static_cast<int64_t>(t.b) * cout + c) * hw + sp] = (a > 0.f ? a : a * ss[c]) + pre[j][c];
              }
            }
          }
        } else if (staged) {
          RowMap rm;
          rm.b = t.b, rm.y0 = t.y0 + j * TH, rm.x0 = t.x0, rm.tw_shift = 3, rm.H = H, rm.W = W, rm.up = 0, rm.ky = 0, rm.kx = 0;
          epilogue_subtile_staged(ea, taddr, block_n, bs, ss, rm, we * 32, lane, t.n0, rel, stg, satm);
        } else {
          epilogue_subtile(ea, taddr, block_n, bs, ss, valid, t.b, y, x, pix, pix, t.n0, rel, satm);
        }
      }
    }
    sat_report(ea.sat_flag, satm);
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc(tmem_base, static_cast<uint32_t>(p.tmem_cols));
  }
  if (threadIdx.x == 0) TL_MARK(p, TL_EXIT);
}

SmemOptIn g_smem_opt_in;

}  // namespace

int resolve_conv3x3_slab(LaunchCfg* cfg, int grid) {
  const KParams& p = cfg->p;
  B200DN_CHECK_ARG(p.wres || p.num_stages >= 2, "conv3x3 slab: W ring too small for block_n %d", p.block_n);
  B200DN_CHECK_ARG(p.num_slabs <= MAX_SLABS, "conv3x3 slab: too many slabs");
  static const void* const kernels[4] = {reinterpret_cast<const void*>(conv3x3_slab_kernel<1, false>),
                                         reinterpret_cast<const void*>(conv3x3_slab_kernel<1, true>),
                                         reinterpret_cast<const void*>(conv3x3_slab_kernel<2, false>),
                                         reinterpret_cast<const void*>(conv3x3_slab_kernel<2, true>)};
  if (int rc = ensure_max_dyn_smem(g_smem_opt_in, kernels, 4, SMEM_BYTES_SLAB,
                                   "cudaFuncSetAttribute(conv3x3_slab_kernel, smem)"))
    return rc;
  cfg->kernel = kernels[(p.mt - 1) * 2 + (p.wres ? 1 : 0)];
  cfg->grid = grid;
  cfg->threads = NUM_THREADS + ((p.mt == 2 && p.wres) ? EPI_THREADS : 0);   // second epilogue group
  cfg->smem = SMEM_BYTES_SLAB;
  cfg->cluster = 1;
  return 0;
}

}  // namespace igemm
}  // namespace b200dn
