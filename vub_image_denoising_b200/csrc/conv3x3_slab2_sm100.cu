// conv3x3_slab2_sm100.cu — the haloed-slab 3x3 convolution (conv3x3_slab_sm100.cu) on CTA PAIRS:
// `tcgen05.mma.cta_group::2`, M = 256 per instruction (128 pixels per SM), thread-block cluster of 2.
//
// Reference op: nn.Conv2d(.,.,3,padding=1) + nn.PReLU (+ residual), UNet/RDUNet_model.py:61,74-75,98-115.
//
// Why: with cta_group::1 every UMMA re-reads its whole B operand ([N x 16] weights) from the issuing SM's shared
// memory, and the N = 64 / 128 layers are bound by shared-memory bandwidth (A 4 KB + B N*32 B per UMMA against
// ~128 B/clk/SM).  In a CTA pair each SM keeps only HALF of the W tile (rows [r*N/2, (r+1)*N/2))
// and the hardware shares the halves across the pair, so per SM the W smem reads, the W TMA writes and the
// L2->SM weight traffic all halve, while each CTA still owns its pixels (slab), its accumulators (TMEM) and its
// epilogue exactly as in the single-CTA kernel.
//
// Protocol (rank 0 = leader):
//   * both CTAs run the slab producer (warp 0) and the W producer (warp 2) into their OWN shared memory, but every
//     TMA credits its bytes to the LEADER's full barrier (`.cta_group::2` TMA + mapa address);
//   * only the leader's issuer warps (1 / 3) issue MMAs; `tcgen05.commit ... multicast::cluster` with mask 0b11
//     arrives on the empty / accumulator-full barriers at the same offset in BOTH CTAs;
//   * both epilogues (warps 4..7) drain their own TMEM and release the accumulator stage by a remote arrive on the
//     leader's barrier (one relaxed arrive per epilogue warp, count = 2 x 4).
// Tile schedule: pair tile = (two consecutive spatial super tiles, one N tile); clusters take pair tiles round-robin.
#include "igemm_common.cuh"


namespace b200dn {
namespace igemm {

namespace {

constexpr int TW = SLAB_TILE_W;   // 8
constexpr int TH = SLAB_TILE_H;   // 16
constexpr int MAX_SLABS = SLAB_MAX_SLABS;
constexpr int DATA_BYTES = SLAB_DATA_BYTES + EPI_STAGING_BYTES;   // [slabs | W ring (half tiles) | epilogue staging]
constexpr int SMEM_BYTES_SLAB2 = 1024 + DATA_BYTES + SLAB_CTRL_BYTES + 2 * 2 * MAX_N * 4;

// barrier map (byte offsets from `bars`): up to 8 slab slots and 8 W stages
constexpr uint32_t B_SLAB_FULL = 0, B_SLAB_EMPTY = 64, B_W_FULL = 128, B_W_EMPTY = 192, B_TFULL = 256, B_TEMPTY = 272,
                   B_TMEM_PTR = 288;

// Division-free walk over the pair tiles cluster_id, cluster_id + num_clusters, ... : pair tile = (m_pair, N tile), this
// CTA's spatial super tile is m_tile = 2 * m_pair + rank (an all-out-of-bounds tile past the end for the odd CTA of the
// last pair).  The strides are decomposed once into (tile column, tile row, image) digits and added with carries.
struct PairWalker {
  int nt, tx, ty, b;
  int s_nt, d_tx, d_ty, d_b, e_tx, e_ty, e_b;   // stride digits: N tile; 2 * pair stride; the extra 2 on an N-tile carry
  int num_n_tiles, tiles_x, tiles_y, images;
  __device__ __forceinline__ static void split(int m, int tiles_x, int tiles_y, int& tx, int& ty, int& b) {
    const int q = m / tiles_x;
    tx = m - q * tiles_x;
    b = q / tiles_y;
    ty = q - b * tiles_y;
  }
  __device__ __forceinline__ void init(int pair_tile, int stride, int rank, const KParams& p) {
    num_n_tiles = p.num_n_tiles, tiles_x = p.tiles_x, tiles_y = p.tiles_y, images = p.B;
    const int m_pair = pair_tile / num_n_tiles;
    nt = pair_tile - m_pair * num_n_tiles;
    split(2 * m_pair + rank, tiles_x, tiles_y, tx, ty, b);
    const int s_pair = stride / num_n_tiles;
    s_nt = stride - s_pair * num_n_tiles;
    split(2 * s_pair, tiles_x, tiles_y, d_tx, d_ty, d_b);
    split(2, tiles_x, tiles_y, e_tx, e_ty, e_b);
  }
  __device__ __forceinline__ void add(int ax, int ay, int ab) {
    tx += ax;
    int c = tx >= tiles_x ? 1 : 0;
    tx -= c ? tiles_x : 0;
    ty += ay + c;
    c = ty >= tiles_y ? 1 : 0;
    ty -= c ? tiles_y : 0;
    b += ab + c;
  }
  __device__ __forceinline__ void next() {
    nt += s_nt;
    const bool carry = nt >= num_n_tiles;
    nt -= carry ? num_n_tiles : 0;
    add(d_tx, d_ty, d_b);
    if (carry) add(e_tx, e_ty, e_b);
  }
  __device__ __forceinline__ TileCoord coord(int block_n, int tile_w, int tile_h) const {
    TileCoord t;
    t.grp = 0, t.n0 = nt * block_n;
    if (b >= images) {
      t.b = 0, t.x0 = 0, t.y0 = tiles_y * tile_h;   // >= H: TMA zero-fills, the epilogue masks every pixel
    } else {
      t.b = b, t.y0 = ty * tile_h, t.x0 = tx * tile_w;
    }
    return t;
  }
};

template <int MT>
__global__ void __launch_bounds__(NUM_THREADS, 1) conv3x3_slab2_kernel(const __grid_constant__ KParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_u32 = smem_u32(smem_raw);
  const uint32_t smem_base = (raw_u32 + 1023u) & ~1023u;
  uint8_t* smem_gen = smem_raw + (smem_base - raw_u32);
  const uint32_t bars = smem_base + DATA_BYTES;
  uint32_t* tmem_ptr_s = reinterpret_cast<uint32_t*>(smem_gen + DATA_BYTES + B_TMEM_PTR);
  float* epi_bias = reinterpret_cast<float*>(smem_gen + DATA_BYTES + SLAB_CTRL_BYTES);  // [2][MAX_N]
  float* epi_slope = epi_bias + 2 * MAX_N;

  const int warp = (threadIdx.x >> 5) ^ 4;   // hardware warps 4..7 = producers / issuers, 0..3 = epilogue
  const int lane = threadIdx.x & 31;
  const int rank = static_cast<int>(cluster_ctarank());
  const bool leader = rank == 0;
  const int cluster_id = blockIdx.x >> 1, num_clusters = gridDim.x >> 1;
  if (threadIdx.x == 0) TL_MARK(p, TL_ENTRY);

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&p.tmA0);
    tma_prefetch_desc(&p.tmW);
    if (p.n_pairs > 1) tma_prefetch_desc(&p.tmA1);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < MAX_SLABS; ++s) {
      mbar_init(bars + B_SLAB_FULL + s * 8, 1);      // leader's: one arrive.expect_tx for both CTAs' bytes
      mbar_init(bars + B_SLAB_EMPTY + s * 8, MT);    // multicast commits from the leader's issuers
    }
    for (int s = 0; s < MAX_STAGES; ++s) {
      mbar_init(bars + B_W_FULL + s * 8, 1);
      mbar_init(bars + B_W_EMPTY + s * 8, MT);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(bars + B_TFULL + a * 8, MT);
      mbar_init(bars + B_TEMPTY + a * 8, 2 * (EPI_THREADS / 32));   // leader's: one arrive per epilogue warp of both CTAs
    }
    mbar_fence_init();
  }
  if (warp == 2) {
    tmem_alloc_2cta(smem_u32(tmem_ptr_s), static_cast<uint32_t>(p.tmem_cols));
    tmem_relinquish_2cta();
  }
  tc_fence_before();
  cluster_sync_all();   // barrier inits and the TMEM allocation of BOTH CTAs are visible before any remote arrive
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_s;
  if (threadIdx.x == 0) TL_MARK(p, TL_SETUP);
  griddep_launch_dependents();
  // the W producer only reads the packed weights (static data): it fills its ring while the previous kernel of the
  // stream is still running; every other role touches activations and waits for it
  if (warp != 2) griddep_wait();
  if (threadIdx.x == 0) TL_MARK(p, TL_DEP);
  [[maybe_unused]] bool tl_first = false;

  const int num_pair_tiles = p.num_tiles;   // pair tiles
  const int block_n = p.block_n, half_n = p.block_n >> 1;
  const int n_cblk = p.n_cblk, n_pairs = p.n_pairs;
  const int num_slabs = p.num_slabs, slab_bytes = p.slab_bytes;
  const int num_stages = p.num_stages, stage_bytes = p.stage_bytes;
  const uint32_t wring = smem_base + static_cast<uint32_t>(num_slabs * slab_bytes);
  constexpr int STH = TH * MT;

  if (warp == 0) {
    // ===================================================== slab TMA producer (both CTAs, own pixels)
    int s = 0;
    uint32_t sph = 0;
    const int pa0 = p.pair_a[0], pa1 = p.pair_a[1], pa2 = p.pair_a[2];
    PairWalker pw_;
    pw_.init(cluster_id, num_clusters, rank, p);
    for (int tile = cluster_id; tile < num_pair_tiles; tile += num_clusters, pw_.next()) {
      const TileCoord t = pw_.coord(block_n, TW, STH);
      for (int pair = 0; pair < n_pairs; ++pair) {
        const int pa = pair == 0 ? pa0 : pair == 1 ? pa1 : pa2;
        const CUtensorMap* tmA = pa ? &p.tmA1 : &p.tmA0;
        for (int cb = 0; cb < n_cblk; ++cb) {
          mbar_wait(bars + B_SLAB_EMPTY + s * 8, sph ^ 1u);
          if (elect_one()) {
            const uint32_t full = bars + B_SLAB_FULL + s * 8;
            if (leader) mbar_arrive_expect_tx(full, static_cast<uint32_t>(2 * p.slab_tx));
            tma_load_4d_2sm(smem_base + s * slab_bytes, tmA, mapa_shared(full, 0), cb * BLOCK_K, t.x0 - 1, t.y0 - 1, t.b);
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
    // ===================================================== W producer (both CTAs): HALF of each [N x 64] tap tile
    int ws = 0;
    uint32_t wph = 0;
    const int w_taps = p.w_taps;
    const uint32_t w_bytes = static_cast<uint32_t>(block_n * 128 * w_taps);   // both halves of w_taps tap tiles
    const int pw0 = p.pair_w[0] * 9, pw1 = p.pair_w[1] * 9, pw2 = p.pair_w[2] * 9;
    PairWalker pw_;
    pw_.init(cluster_id, num_clusters, rank, p);
    for (int tile = cluster_id; tile < num_pair_tiles; tile += num_clusters, pw_.next()) {
      const TileCoord t = pw_.coord(block_n, TW, STH);
      const int n_row = t.n0 + rank * half_n;
      for (int pair = 0; pair < n_pairs; ++pair) {
        const int wbase = pair == 0 ? pw0 : pair == 1 ? pw1 : pw2;
        for (int cb = 0; cb < n_cblk; ++cb) {
#pragma unroll 1
          for (int tap = 0; tap < 9; tap += w_taps) {
            mbar_wait(bars + B_W_EMPTY + ws * 8, wph ^ 1u);
            if (elect_one()) {
              const uint32_t full = bars + B_W_FULL + ws * 8;
              if (leader) mbar_arrive_expect_tx(full, w_bytes);
              tma_load_3d_2sm(wring + ws * stage_bytes, &p.tmW, mapa_shared(full, 0), cb * BLOCK_K, n_row, wbase + tap);
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
  } else if (leader && (warp == 1 || (warp == 3 && MT == 2))) {
    // ===================================================== MMA issuer of sub-tile j (leader CTA, drives both SMs)
    // The loop is kept small (dy rolled, descriptors and barrier addresses advanced incrementally, k16 count a
    // compile-time constant per 9-tap block): ncu showed this warp issue-bound (~130 SASS instructions per tap with
    // `no_inst` / `wait` stalls, never waiting for data), which capped the N = 64 layers — 4 UMMAs of 32 clk per tap.
    {
      const int j = warp == 1 ? 0 : 1;
      const uint32_t idesc = make_idesc_f16(static_cast<uint32_t>(p.fmt), static_cast<uint32_t>(block_n), 256u);
      const int kc_iters = n_pairs * n_cblk;
      const int last_k16 = p.last_k16;
      const uint32_t pitch = static_cast<uint32_t>(p.slab_w * 128);   // bytes per slab pixel row = SBO
      const uint64_t row_step = static_cast<uint64_t>(pitch >> 4);    // descriptor address units (16 B)
      const int w_taps = p.w_taps;
      const uint64_t tap_step = static_cast<uint64_t>((half_n * 128) >> 4);   // this CTA's half of one tap tile
      WRing wr;
      wr.full0 = bars + B_W_FULL, wr.empty0 = bars + B_W_EMPTY;
      wr.full_end = wr.full0 + static_cast<uint32_t>(num_stages) * 8;
      wr.desc0 = make_sw128_desc(wring, 1024);
      wr.step = static_cast<uint64_t>(stage_bytes >> 4);
      wr.full = wr.full0, wr.empty = wr.empty0, wr.phase = 0, wr.desc = wr.desc0;
      int s = 0;
      uint32_t sph = 0;
      int local_tile = 0;
      for (int tile = cluster_id; tile < num_pair_tiles; tile += num_clusters, ++local_tile) {
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
          if (j == 0 && lane == 0) TL_MARK_ONCE(p, TL_MMA0, tl_first);
          const uint32_t slab = smem_base + static_cast<uint32_t>(s * slab_bytes) + static_cast<uint32_t>(j * TH) * pitch;
          uint64_t arow = make_sw128_desc(slab, pitch);
          issue_slab_block_streamed_n<true>(nk, w_taps, d, arow, row_step, wr, tap_step, idesc, accumulate);
          if (elect_one()) {
            umma_commit_2cta(bars + B_SLAB_EMPTY + s * 8, 3);                       // both CTAs' slabs consumed
            if (kc == kc_iters - 1) umma_commit_2cta(bars + B_TFULL + acc * 8, 3);  // both accumulators complete
          }
          __syncwarp();
          if (++s == num_slabs) {
            s = 0;
            sph ^= 1u;
          }
          if (++cb == n_cblk) cb = 0;
        }
      }
      if (j == 0 && lane == 0) TL_MARK(p, TL_MMA_END);
    }
  } else if (warp >= 4) {
    // ======================================================= epilogue (both CTAs, own accumulators)
    const int we = warp & 3;
    const int row = we * 32 + lane;
    const int th = row / TW, tw = row - th * TW;
    const int et = threadIdx.x & (EPI_THREADS - 1);
    const EpiArgs ea = make_epi_args(p);
    const int H = p.H, W = p.W, cout = p.cout;
    const float* bias = p.bias;
    const float* slope = p.slope;
    const bool staged = p.epi_staged;
    uint8_t* stg = smem_gen + SLAB_DATA_BYTES + we * 4096;
    const uint32_t tempty_leader = mapa_shared(bars + B_TEMPTY, 0);
    const bool one_n_tile = p.num_n_tiles == 1;
    if (one_n_tile) stage_bias_slope(epi_bias, epi_slope, bias, slope, 0, block_n, cout, et);
    int local_tile = 0;
    uint32_t satm = 0;
    PairWalker pw_;
    pw_.init(cluster_id, num_clusters, rank, p);
    for (int tile = cluster_id; tile < num_pair_tiles; tile += num_clusters, ++local_tile, pw_.next()) {
      const TileCoord t = pw_.coord(block_n, TW, STH);
      const int acc = local_tile & 1;
      const uint32_t acc_phase = (local_tile >> 1) & 1;
      // bias / PReLU slopes of the tile's N range: staged once when the layer has a single N tile (per-tile global
      // loads + a named barrier are a large share of a tile of the small-K layers), else per tile, double-buffered
      float* bs = epi_bias + (one_n_tile ? 0 : acc * MAX_N);
      float* ss = epi_slope + (one_n_tile ? 0 : acc * MAX_N);
      if (!one_n_tile) stage_bias_slope(bs, ss, bias, slope, t.n0, block_n, cout, et);

      mbar_wait(bars + B_TFULL + acc * 8, acc_phase);
      tc_fence_after();
      if (we == 0 && lane == 0) TL_MARK_ONCE(p, TL_ACC0, tl_first);

#pragma unroll 1
      for (int j = 0; j < MT; ++j) {
        const int y = t.y0 + j * TH + th, x = t.x0 + tw;
        const bool valid = (y < H) && (x < W);
        const int64_t pix = (static_cast<int64_t>(t.b) * H + y) * W + x;
        const uint32_t taddr =
            tmem_base + (static_cast<uint32_t>(we * 32) << 16) + static_cast<uint32_t>((acc * MT + j) * block_n);
        const uint32_t rel = j == MT - 1 ? tempty_leader + acc * 8 : 0u;
        if (staged) {
          RowMap rm;
          rm.b = t.b, rm.y0 = t.y0 + j * TH, rm.x0 = t.x0, rm.tw_shift = 3, rm.H = H, rm.W = W, rm.up = 0, rm.ky = 0, rm.kx = 0;
          epilogue_subtile_staged(ea, taddr, block_n, bs, ss, rm, we * 32, lane, t.n0, rel, stg, satm, true);
        } else {
          epilogue_subtile(ea, taddr, block_n, bs, ss, valid, t.b, y, x, pix, pix, t.n0, rel, satm, true);
        }
      }
    }
    sat_report(ea.sat_flag, satm);
    if (we == 0 && lane == 0) TL_MARK(p, TL_EPI_END);
  }

  // Neither CTA may leave (or free its TMEM) while the other can still read its shared memory through an MMA or
  // signal one of its barriers.
  tc_fence_before();
  cluster_sync_all();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc_2cta(tmem_base, static_cast<uint32_t>(p.tmem_cols));
  }
#ifdef B200DN_TIMELINE
  if (threadIdx.x == 0 && p.tl != nullptr && blockIdx.x == 0) {
    uint32_t smid;
    asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
    p.tl[TL_SM] = smid;
    TL_MARK(p, TL_EXIT);
  }
#endif
}

SmemOptIn g_smem_opt_in;

}  // namespace

int resolve_conv3x3_slab2(LaunchCfg* cfg, int grid) {
  const KParams& p = cfg->p;
  B200DN_CHECK_ARG(p.cta2 && !p.wres, "conv3x3 slab2: not a CTA-pair configuration");
  B200DN_CHECK_ARG(p.num_stages >= 2, "conv3x3 slab2: W ring too small for block_n %d", p.block_n);
  B200DN_CHECK_ARG(p.num_slabs <= MAX_SLABS, "conv3x3 slab2: too many slabs");
  B200DN_CHECK_ARG(grid >= 2 && grid % 2 == 0, "conv3x3 slab2: grid %d must be a positive multiple of 2", grid);
  static const void* const kernels[2] = {reinterpret_cast<const void*>(conv3x3_slab2_kernel<1>),
                                         reinterpret_cast<const void*>(conv3x3_slab2_kernel<2>)};
  if (int rc = ensure_max_dyn_smem(g_smem_opt_in, kernels, 2, SMEM_BYTES_SLAB2,
                                   "cudaFuncSetAttribute(conv3x3_slab2_kernel, smem)"))
    return rc;
  cfg->kernel = kernels[p.mt - 1];
  cfg->grid = grid;
  cfg->threads = NUM_THREADS;
  cfg->smem = SMEM_BYTES_SLAB2;
  cfg->cluster = 2;
  return 0;
}

}  // namespace igemm
}  // namespace b200dn
