// conv3x3_chain_sm100.cu — several DEPENDENT 3x3 convolutions of one resolution in ONE persistent launch.
//
// Reference op: the four chained convolutions of a DenoisingBlock (UNet/RDUNet_model.py:98-115: conv_0..conv_3, every
// conv reads the block input plus all earlier outputs, `torch.cat` = channel slices of one NHWC buffer).
//
// Why: launched one by one (conv3x3_slab2_sm100.cu) every layer pays, on top of its MMAs, a drain (the last tile's
// epilogue, 1.5-12 us), the hand-off to the dependent launch (1-3 us after the last CTA exits) and a cold pipeline fill
// (first TMA 1-2.5 us): ~7 us per layer at 32 images (10 % of the RDUNet_T(32) forward), ~4 us of every ~10 us layer
// at 2 images (profiles/r02_launch_timeline_*.txt).  Here the layers of a block are ONE work list, layer-major:
// item = (layer, pair tile), cluster c takes items c, c + G, c + 2G, ...; a tile of layer k only needs the 3 x 3
// neighbourhood of spatial tiles of layer k - 1 (its 1-pixel halo), so it waits on per-tile arrival counters in global
// memory instead of a grid-wide boundary.  The TMA / MMA / epilogue pipelines of a CTA never drain between layers, the
// weights of the next layer stream in while the current one computes, and tiles of layer k + 1 start while other
// SMs still finish layer k.
//
// Protocol per item (everything else is conv3x3_slab2_sm100.cu: CTA pairs, cta_group::2, streamed half-W tiles):
//   * slab producer, layer k > 0: nine lanes poll the counters of layer k - 1 for the tile's 3 x 3 neighbourhood
//     (ld.acquire.gpu) until each holds epoch x (N tiles x 4 epilogue warps) arrivals, then fence.proxy.async and the
//     TMA.  By induction the neighbourhood is then complete for every earlier layer too.
//   * epilogue warp, layer k < last: after its stores, __threadfence + one red.add on its tile's counter.
//   * counters are monotonic; `epoch` (launches completed + 1) is read from the workspace at kernel start and advanced
//     by the last CTA to leave, so nothing is reset between launches and a graph replay needs no memset node.
// Deadlock freedom: a cluster runs its items in increasing order and an item only waits on lower-numbered items, so the
// lowest unfinished item can always run — provided every cluster becomes resident: the grid is clamped to
// cudaOccupancyMaxActiveClusters, and clusters that find their SMs taken by the previous launch of the stream start
// when it drains (it never waits on this one).  What this does NOT cover is a second chain launch on ANOTHER stream
// of the same device holding SMs while it waits for its own clusters: chain launches of one device must not overlap
// (one plan = one stream here).  B200DN_CHAIN_COOP=1 launches cooperatively (all-or-nothing residency), which lifts
// that restriction but cannot be profiled: ncu 2025.2 fails cooperative cluster launches with LaunchFailed.
// Polls are bounded by the same watchdog as the mbarrier waits (trap, not hang).
// Arithmetic is identical to the per-layer launches (same tiles, same MMA order): outputs are bit-equal.
#include "igemm_common.cuh"

namespace b200dn {
namespace igemm {

namespace {

constexpr int TW = SLAB_TILE_W;   // 8
constexpr int TH = SLAB_TILE_H;   // 16
constexpr int MAX_SLABS = SLAB_MAX_SLABS;
constexpr int DATA_BYTES = SLAB_DATA_BYTES + EPI_STAGING_BYTES;   // [slabs | W ring (half tiles) | epilogue staging]
// bias / PReLU slopes: [CHAIN_MAX_LAYERS][MAX_N] each, staged once per launch for layers of <= MAX_N outputs; layers
// with more outputs restage their N tile per item into slot [acc] of the same arrays' tail (not needed by the RDUNet
// widths that chain: their wide layers have MT = 1 != MT of the block's other layers)
constexpr int SMEM_BYTES_CHAIN = 1024 + DATA_BYTES + SLAB_CTRL_BYTES + 2 * CHAIN_MAX_LAYERS * MAX_N * 4;
constexpr uint32_t B_SLAB_FULL = 0, B_SLAB_EMPTY = 64, B_W_FULL = 128, B_W_EMPTY = 192, B_TFULL = 256, B_TEMPTY = 272,
                   B_TMEM_PTR = 288;

__device__ __forceinline__ uint32_t ld_acquire_gpu(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void red_add_gpu(uint32_t* p, uint32_t v) {
  asm volatile("red.relaxed.gpu.global.add.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

// one work item: which layer, which N tile, which spatial super tile (of this CTA of the pair)
struct Item {
  int layer, n0, m, b, ty, tx;   // m = spatial tile index (b, ty, tx); b >= images: the odd CTA of the last pair
};
__device__ __forceinline__ Item decode_item(const ChainParams& c, int item, int rank) {
  Item it;
  it.layer = (item >= c.item_base[1] ? 1 : 0) + (item >= c.item_base[2] ? 1 : 0) + (item >= c.item_base[3] ? 1 : 0);
  const KParams& p = c.L[it.layer];
  const int pt = item - c.item_base[it.layer];
  const int m_pair = pt / p.num_n_tiles;
  it.n0 = (pt - m_pair * p.num_n_tiles) * p.block_n;
  it.m = 2 * m_pair + rank;
  const int q = it.m / p.tiles_x;
  it.tx = it.m - q * p.tiles_x;
  it.b = q / p.tiles_y;
  it.ty = q - it.b * p.tiles_y;
  return it;
}

// Wait until layer `layer`'s outputs exist on the 3 x 3 spatial tiles around (b, ty, tx): lanes 0..8 poll one counter each.
__device__ __forceinline__ void wait_neighbourhood(const ChainParams& c, int layer, const Item& it, uint32_t epoch, int lane) {
  const KParams& p = c.L[layer];
  if (lane < 9) {
    const int dy = lane / 3 - 1, dx = lane - (lane / 3) * 3 - 1;
    const int ny = it.ty + dy, nx = it.tx + dx;
    if (ny >= 0 && ny < p.tiles_y && nx >= 0 && nx < p.tiles_x) {
      const uint32_t* f = c.flags + static_cast<size_t>(layer) * p.num_m_tiles + ((it.b * p.tiles_y + ny) * p.tiles_x + nx);
      const uint32_t need = epoch * static_cast<uint32_t>(c.flag_need[layer]);
      if (static_cast<int32_t>(ld_acquire_gpu(f) - need) < 0) {
        const long long t0 = clock64();
        uint32_t spins = 0;
        while (static_cast<int32_t>(ld_acquire_gpu(f) - need) < 0) {
          if ((++spins & 0xff) == 0 && (clock64() - t0) > B200DN_WATCHDOG_CYCLES) {
            printf("b200dn: conv chain watchdog (block %d layer %d tile %d,%d,%d)\n", (int)blockIdx.x, layer, it.b, ny, nx);
            __trap();
          }
        }
      }
    }
  }
  __syncwarp();
  asm volatile("fence.proxy.async;" ::: "memory");   // generic-proxy writes of other SMs -> this SM's TMA reads
}

template <int MT>
__global__ void __launch_bounds__(NUM_THREADS, 1) conv3x3_chain_kernel(const __grid_constant__ ChainParams c) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_u32 = smem_u32(smem_raw);
  const uint32_t smem_base = (raw_u32 + 1023u) & ~1023u;
  uint8_t* smem_gen = smem_raw + (smem_base - raw_u32);
  const uint32_t bars = smem_base + DATA_BYTES;
  uint32_t* tmem_ptr_s = reinterpret_cast<uint32_t*>(smem_gen + DATA_BYTES + B_TMEM_PTR);
  float* epi_bias = reinterpret_cast<float*>(smem_gen + DATA_BYTES + SLAB_CTRL_BYTES);  // [layers][MAX_N]
  float* epi_slope = epi_bias + CHAIN_MAX_LAYERS * MAX_N;

  const int warp = (threadIdx.x >> 5) ^ 4;   // hardware warps 4..7 = producers / issuers, 0..3 = epilogue
  const int lane = threadIdx.x & 31;
  const int rank = static_cast<int>(cluster_ctarank());
  const bool leader = rank == 0;
  const int cluster_id = blockIdx.x >> 1, num_clusters = gridDim.x >> 1;
  const int n_layers = c.n_layers;
  const int total_items = c.item_base[n_layers];

  if (warp == 0 && lane == 0) {
    for (int l = 0; l < n_layers; ++l) {
      tma_prefetch_desc(&c.L[l].tmA0);
      tma_prefetch_desc(&c.L[l].tmW);
    }
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
      mbar_init(bars + B_TEMPTY + a * 8, 2 * (EPI_THREADS / 32));
    }
    mbar_fence_init();
  }
  if (warp == 2) {
    tmem_alloc_2cta(smem_u32(tmem_ptr_s), static_cast<uint32_t>(c.tmem_cols));
    tmem_relinquish_2cta();
  }
  tc_fence_before();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_s;
  griddep_launch_dependents();
  if (warp != 2) griddep_wait();   // the W producer only reads the packed weights (static data)

  // ring geometry is common to the layers (the host unified it); everything else is looked up per item
  const int num_slabs = c.L[0].num_slabs, slab_bytes = c.L[0].slab_bytes;
  const int num_stages = c.L[0].num_stages, stage_bytes = c.L[0].stage_bytes;
  const uint32_t wring = smem_base + static_cast<uint32_t>(num_slabs * slab_bytes);
  const uint32_t acc_stride = static_cast<uint32_t>(c.nmax);   // TMEM columns between accumulators
  constexpr int STH = TH * MT;
  const int images = c.L[0].B;

  if (warp == 0) {
    // ===================================================== slab TMA producer (both CTAs, own pixels)
    const uint32_t epoch = *reinterpret_cast<const volatile uint32_t*>(c.sync) + 1u;
    int s = 0;
    uint32_t sph = 0;
    for (int item = cluster_id; item < total_items; item += num_clusters) {
      const Item it = decode_item(c, item, rank);
      const KParams& p = c.L[it.layer];
      const bool live = it.b < images;
      if (it.layer > 0 && live) wait_neighbourhood(c, it.layer - 1, it, epoch, lane);
      const int y0 = live ? it.ty * STH : p.tiles_y * STH, x0 = live ? it.tx * TW : 0, b = live ? it.b : 0;
      const int n_cblk = p.n_cblk;
      for (int cb = 0; cb < n_cblk; ++cb) {
        mbar_wait(bars + B_SLAB_EMPTY + s * 8, sph ^ 1u);
        if (elect_one()) {
          const uint32_t full = bars + B_SLAB_FULL + s * 8;
          if (leader) mbar_arrive_expect_tx(full, static_cast<uint32_t>(2 * p.slab_tx));
          tma_load_4d_2sm(smem_base + s * slab_bytes, &p.tmA0, mapa_shared(full, 0), cb * BLOCK_K, x0 - 1, y0 - 1, b);
        }
        __syncwarp();
        if (++s == num_slabs) {
          s = 0;
          sph ^= 1u;
        }
      }
    }
  } else if (warp == 2) {
    // ===================================================== W producer (both CTAs): HALF of each [N x 64] tap tile
    int ws = 0;
    uint32_t wph = 0;
    for (int item = cluster_id; item < total_items; item += num_clusters) {
      const Item it = decode_item(c, item, rank);
      const KParams& p = c.L[it.layer];
      const int w_taps = p.w_taps, n_cblk = p.n_cblk;
      const uint32_t w_bytes = static_cast<uint32_t>(p.block_n * 128 * w_taps);
      const int n_row = it.n0 + rank * (p.block_n >> 1);
      for (int cb = 0; cb < n_cblk; ++cb) {
#pragma unroll 1
        for (int tap = 0; tap < 9; tap += w_taps) {
          mbar_wait(bars + B_W_EMPTY + ws * 8, wph ^ 1u);
          if (elect_one()) {
            const uint32_t full = bars + B_W_FULL + ws * 8;
            if (leader) mbar_arrive_expect_tx(full, w_bytes);
            tma_load_3d_2sm(wring + ws * stage_bytes, &p.tmW, mapa_shared(full, 0), cb * BLOCK_K, n_row, tap);
          }
          __syncwarp();
          if (++ws == num_stages) {
            ws = 0;
            wph ^= 1u;
          }
        }
      }
    }
  } else if (leader && (warp == 1 || (warp == 3 && MT == 2))) {
    // ===================================================== MMA issuer of sub-tile j (leader CTA, drives both SMs)
    const int j = warp == 1 ? 0 : 1;
    const uint32_t pitch = static_cast<uint32_t>(c.L[0].slab_w * 128);   // bytes per slab pixel row = SBO
    const uint64_t row_step = static_cast<uint64_t>(pitch >> 4);
    WRing wr;
    wr.full0 = bars + B_W_FULL, wr.empty0 = bars + B_W_EMPTY;
    wr.full_end = wr.full0 + static_cast<uint32_t>(num_stages) * 8;
    wr.desc0 = make_sw128_desc(wring, 1024);
    wr.step = static_cast<uint64_t>(stage_bytes >> 4);
    wr.full = wr.full0, wr.empty = wr.empty0, wr.phase = 0, wr.desc = wr.desc0;
    int s = 0;
    uint32_t sph = 0;
    int local = 0;
    for (int item = cluster_id; item < total_items; item += num_clusters, ++local) {
      const int layer = (item >= c.item_base[1] ? 1 : 0) + (item >= c.item_base[2] ? 1 : 0) + (item >= c.item_base[3] ? 1 : 0);
      const KParams& p = c.L[layer];
      const uint32_t idesc = make_idesc_f16(static_cast<uint32_t>(p.fmt), static_cast<uint32_t>(p.block_n), 256u);
      const int n_cblk = p.n_cblk, last_k16 = p.last_k16, w_taps = p.w_taps;
      const uint64_t tap_step = static_cast<uint64_t>(((p.block_n >> 1) * 128) >> 4);
      const int acc = local & 1;
      const uint32_t acc_phase = (local >> 1) & 1;
      mbar_wait(bars + B_TEMPTY + acc * 8, acc_phase ^ 1u);
      tc_fence_after();
      const uint32_t d = tmem_base + static_cast<uint32_t>(acc * MT + j) * acc_stride;
      uint32_t accumulate = 0;
      for (int cb = 0; cb < n_cblk; ++cb) {
        const int nk = (cb != n_cblk - 1) ? BLOCK_K / 16 : last_k16;
        mbar_wait(bars + B_SLAB_FULL + s * 8, sph);
        const uint32_t slab = smem_base + static_cast<uint32_t>(s * slab_bytes) + static_cast<uint32_t>(j * TH) * pitch;
        uint64_t arow = make_sw128_desc(slab, pitch);
        issue_slab_block_streamed_n<true>(nk, w_taps, d, arow, row_step, wr, tap_step, idesc, accumulate);
        if (elect_one()) {
          umma_commit_2cta(bars + B_SLAB_EMPTY + s * 8, 3);
          if (cb == n_cblk - 1) umma_commit_2cta(bars + B_TFULL + acc * 8, 3);
        }
        __syncwarp();
        if (++s == num_slabs) {
          s = 0;
          sph ^= 1u;
        }
      }
    }
  } else if (warp >= 4) {
    // ======================================================= epilogue (both CTAs, own accumulators)
    const int we = warp & 3;
    const int row = we * 32 + lane;
    const int th = row / TW, tw = row - th * TW;
    const int et = threadIdx.x & (EPI_THREADS - 1);
    uint8_t* stg = smem_gen + SLAB_DATA_BYTES + we * 4096;
    const uint32_t tempty_leader = mapa_shared(bars + B_TEMPTY, 0);
    int local = 0;
    uint32_t satm = 0;
    int* sat_flag = nullptr;
    // bias / slopes of every layer once (static data: the global loads overlap the first accumulator's MMAs)
    for (int l = 0; l < n_layers; ++l) {
      const KParams& pl = c.L[l];
      for (int i = et; i < MAX_N; i += EPI_THREADS) {
        epi_bias[l * MAX_N + i] = i < pl.cout ? __ldg(pl.bias + i) : 0.f;
        epi_slope[l * MAX_N + i] = (pl.slope != nullptr && i < pl.cout) ? __ldg(pl.slope + i) : 1.f;
      }
    }
    asm volatile("bar.sync 1, %0;" ::"r"(EPI_THREADS) : "memory");
    for (int item = cluster_id; item < total_items; item += num_clusters, ++local) {
      const Item it = decode_item(c, item, rank);
      const KParams& p = c.L[it.layer];
      const EpiArgs ea = make_epi_args(p);
      sat_flag = ea.sat_flag;
      const int H = p.H, W = p.W, block_n = p.block_n;
      const bool live = it.b < images;
      const int y0 = live ? it.ty * STH : p.tiles_y * STH, x0 = live ? it.tx * TW : 0, b = live ? it.b : 0;
      const int acc = local & 1;
      const uint32_t acc_phase = (local >> 1) & 1;
      const float* bs = epi_bias + it.layer * MAX_N + it.n0;
      const float* ss = epi_slope + it.layer * MAX_N + it.n0;

      mbar_wait(bars + B_TFULL + acc * 8, acc_phase);
      tc_fence_after();
#pragma unroll 1
      for (int j = 0; j < MT; ++j) {
        const int y = y0 + j * TH + th, x = x0 + tw;
        const bool valid = (y < H) && (x < W);
        const int64_t pix = (static_cast<int64_t>(b) * H + y) * W + x;
        const uint32_t taddr = tmem_base + (static_cast<uint32_t>(we * 32) << 16) + static_cast<uint32_t>(acc * MT + j) * acc_stride;
        const uint32_t rel = j == MT - 1 ? tempty_leader + acc * 8 : 0u;
        if (p.epi_staged) {
          RowMap rm;
          rm.b = b, rm.y0 = y0 + j * TH, rm.x0 = x0, rm.tw_shift = 3, rm.H = H, rm.W = W, rm.up = 0, rm.ky = 0, rm.kx = 0;
          epilogue_subtile_staged(ea, taddr, block_n, bs, ss, rm, we * 32, lane, it.n0, rel, stg, satm, true);
        } else {
          epilogue_subtile(ea, taddr, block_n, bs, ss, valid, b, y, x, pix, pix, it.n0, rel, satm, true);
        }
      }
      if (it.layer + 1 < n_layers && live) {
        // publish: this warp's share of the tile is in global memory
        // (posting is never deferred past a wait: an arrival held back until the next accumulator is complete can
        // close a cycle of clusters waiting on each other's held-back arrivals in the last round of a layer)
        __threadfence();
        __syncwarp();
        if (lane == 0) red_add_gpu(c.flags + static_cast<size_t>(it.layer) * p.num_m_tiles + it.m, 1u);
      }
    }
    sat_report(sat_flag, satm);
  }

  // Neither CTA may leave (or free its TMEM) while the other can still read its shared memory through an MMA or
  // signal one of its barriers.
  tc_fence_before();
  cluster_sync_all();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc_2cta(tmem_base, static_cast<uint32_t>(c.tmem_cols));
  }
  if (threadIdx.x == 0) {
    // the last CTA to leave closes the epoch: every poll of this launch has returned by then
    __threadfence();
    const uint32_t old = atomicAdd(c.sync + 1, 1u);
    if (old == gridDim.x - 1) {
      c.sync[1] = 0;
      __threadfence();
      atomicAdd(c.sync, 1u);
    }
  }
}

SmemOptIn g_chain_opt_in;

bool chain_cooperative() {
  static const bool v = [] {
    const char* e = getenv("B200DN_CHAIN_COOP");
    return e ? atoi(e) != 0 : false;
  }();
  return v;
}

}  // namespace

size_t conv_chain_workspace_bytes(const b200dn_igemm_args* a, int n) {
  // upper bound: one counter per 8 x 16 tile (MT = 1) and layer, + {epoch, exit counter}
  const size_t tiles = static_cast<size_t>(a[0].B) * cdiv(a[0].W, TW) * cdiv(a[0].H, TH);
  return (tiles * static_cast<size_t>(n > 1 ? n - 1 : 1) + 16) * sizeof(uint32_t);
}

int configure_conv_chain(const b200dn_igemm_args* a, int n, void* workspace, int flags, LaunchCfg* cfg) {
  B200DN_CHECK_ARG(a != nullptr && cfg != nullptr, "conv_chain: null argument");
  B200DN_CHECK_ARG(n >= 2 && n <= CHAIN_MAX_LAYERS, "conv_chain: %d layers (2..%d supported)", n, CHAIN_MAX_LAYERS);
  B200DN_CHECK_ARG(workspace != nullptr && (reinterpret_cast<uintptr_t>(workspace) & 15) == 0,
                   "conv_chain: workspace must be a 16-byte aligned device buffer, zero-filled once");
  ChainParams& c = cfg->c;
  memset(&c, 0, sizeof(c));
  int grid_clusters = 0;
  for (int l = 0; l < n; ++l) {
    const b200dn_igemm_args& al = a[l];
    if (al.mode != B200DN_MODE_CONV3X3 || al.out_kind != B200DN_OUT_NHWC16 || al.impl != 0 || al.block_n != 0 ||
        al.m_tiles != 0 || al.max_ctas != 0 ||
        !(al.prec == B200DN_PREC_BF16 || al.prec == B200DN_PREC_FP16)) {
      set_error("conv_chain: layer %d is not a default-configured single-plane 3x3 convolution with NHWC output", l);
      return B200DN_E_UNSUP;
    }
    B200DN_CHECK_ARG(al.B == a[0].B && al.H == a[0].H && al.W == a[0].W && al.prec == a[0].prec,
                     "conv_chain: layer %d differs from layer 0 in batch / size / precision", l);
    // a layer may overwrite nothing that it or an EARLIER layer of the chain reads as input (their tiles may still be
    // in flight): its outputs must lie beyond those input prefixes [0, cin) of the same buffer.  Later layers reading
    // them is the dependency the counters order.
    for (int j = 0; j <= l; ++j)
      B200DN_CHECK_ARG(al.out[0] != a[j].in[0] || al.out_coff >= a[j].cin,
                       "conv_chain: layer %d writes channels [%d, %d) that layer %d reads", l, al.out_coff,
                       al.out_coff + al.cout, j);
    LaunchCfg one;
    if (int rc = igemm_configure(al, &one)) return rc;
    const KParams& p = one.p;
    if (al.cout > MAX_N) {
      set_error("conv_chain: layer %d has %d outputs (bias / slopes of at most %d are staged)", l, al.cout, MAX_N);
      return B200DN_E_UNSUP;
    }
    if (!(p.cta2 && one.cluster == 2 && p.n_pairs == 1 && !p.wres)) {
      set_error("conv_chain: layer %d does not run on the CTA-pair slab kernel", l);
      return B200DN_E_UNSUP;
    }
    c.L[l] = p;
    if (l > 0 && (p.mt != c.L[0].mt || p.tiles_x != c.L[0].tiles_x || p.tiles_y != c.L[0].tiles_y ||
                  p.num_m_tiles != c.L[0].num_m_tiles || p.slab_w != c.L[0].slab_w || p.slab_bytes != c.L[0].slab_bytes)) {
      set_error("conv_chain: layers 0 and %d are tiled differently", l);
      return B200DN_E_UNSUP;
    }
    c.item_base[l + 1] = c.item_base[l] + p.num_tiles;
    c.flag_need[l] = p.num_n_tiles * (EPI_THREADS / 32);
    if (p.block_n > c.nmax) c.nmax = p.block_n;
    if (one.grid / 2 > grid_clusters) grid_clusters = one.grid / 2;
  }
  // Few tiles per layer (batch 1-2, the deep levels): every tile of layer k + 1 waits for tiles of layer k that are still
  // in flight, and a store -> fence -> counter -> poll -> TMA hand-off is no shorter than a launch boundary: measured
  // 3 % SLOWER than four launches at 2 images, 2 % faster at 32 (profiles/r02_conv_chain.txt).
  if (c.item_base[1] < 2 * grid_clusters && !(flags & B200DN_CHAIN_FORCE)) {
    set_error("conv_chain: %d pair tiles per layer on %d clusters: the layers would not overlap", c.item_base[1], grid_clusters);
    return B200DN_E_UNSUP;
  }
  for (int l = n; l < CHAIN_MAX_LAYERS; ++l) c.item_base[l + 1] = c.item_base[n];
  c.n_layers = n;
  // one ring geometry for all layers: the shallowest slab ring, the widest W stage
  int num_slabs = SLAB_MAX_SLABS, stage = 0;
  for (int l = 0; l < n; ++l) {
    if (c.L[l].num_slabs < num_slabs) num_slabs = c.L[l].num_slabs;
    if (c.L[l].stage_bytes > stage) stage = c.L[l].stage_bytes;
  }
  int num_stages = (SLAB_DATA_BYTES - num_slabs * c.L[0].slab_bytes) / stage;
  if (num_stages > MAX_STAGES) num_stages = MAX_STAGES;
  if (num_stages < 2 || num_slabs < 2) {
    set_error("conv_chain: no common ring geometry (%d slabs, %d W stages)", num_slabs, num_stages);
    return B200DN_E_UNSUP;
  }
  for (int l = 0; l < n; ++l) c.L[l].num_slabs = num_slabs, c.L[l].stage_bytes = stage, c.L[l].num_stages = num_stages;
  const int mt = c.L[0].mt;
  int cols = 32;
  while (cols < 2 * mt * c.nmax) cols <<= 1;
  if (cols > 512) {
    set_error("conv_chain: %d accumulators of %d columns do not fit TMEM", 2 * mt, c.nmax);
    return B200DN_E_UNSUP;
  }
  c.tmem_cols = cols;
  uint32_t* ws = static_cast<uint32_t*>(workspace);
  c.sync = ws;             // [0] epoch, [1] exit counter
  c.flags = ws + 16;
  static const void* const kernels[2] = {reinterpret_cast<const void*>(conv3x3_chain_kernel<1>),
                                         reinterpret_cast<const void*>(conv3x3_chain_kernel<2>)};
  if (int rc = ensure_max_dyn_smem(g_chain_opt_in, kernels, 2, SMEM_BYTES_CHAIN, "cudaFuncSetAttribute(conv3x3_chain_kernel, smem)"))
    return rc;
  // every cluster must be resident (tiles wait on tiles of other clusters)
  {
    cudaLaunchConfig_t lc;
    memset(&lc, 0, sizeof(lc));
    lc.gridDim = dim3(static_cast<unsigned>(2 * grid_clusters));
    lc.blockDim = dim3(NUM_THREADS);
    lc.dynamicSmemBytes = SMEM_BYTES_CHAIN;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = 2, at[0].val.clusterDim.y = 1, at[0].val.clusterDim.z = 1;
    lc.attrs = at, lc.numAttrs = 1;
    int max_clusters = 0;
    B200DN_CUDA(cudaOccupancyMaxActiveClusters(&max_clusters, kernels[mt - 1], &lc));
    if (max_clusters < 1) {
      set_error("conv_chain: no resident cluster");
      return B200DN_E_UNSUP;
    }
    if (grid_clusters > max_clusters) grid_clusters = max_clusters;
  }
#ifdef B200DN_TIMELINE
  for (int l = 0; l < n; ++l) c.L[l].tl = nullptr;
#endif
  cfg->kind = 2;
  cfg->kernel = kernels[mt - 1];
  cfg->grid = 2 * grid_clusters;
  cfg->threads = NUM_THREADS;
  cfg->smem = SMEM_BYTES_CHAIN;
  cfg->cluster = 2;
  cfg->cooperative = chain_cooperative() ? 1 : 0;
  return 0;
}

}  // namespace igemm
}  // namespace b200dn
