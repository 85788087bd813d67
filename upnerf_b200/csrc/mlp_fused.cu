// Fused NeRF trunk on tcgen05 -- upnerf_mlp_trunk_fwd_bf16 / upnerf_mlp_trunk_bwd_bf16.
//
// One persistent kernel evaluates, per 128-sample tile, the whole xyz trunk of NeRF.forward
// (reference models/nerf.py:84-93): PE -> 8 x (Linear 256 + ReLU) with the skip concat at
// layer 5 -> xyz_encoding_final, plus the share_sigma row-dot + Softplus (:89); a second one runs the
// data-gradient chain of the same layers.  Activations never leave the SM between layers: the epilogue
// of layer l writes bf16 straight into the 128-byte-swizzled K-major shared-memory boxes that layer
// l+1's MMAs read as their A operand.  Every layer output also goes to HBM once (write-only) because
// the weight gradients need it.
//
//   warp 0      TMA producer: streams the layer weights (N-halves of 64-column K-chunks, 16 KB stages)
//               from L2 -- CTA pairs share the stream by multicast -- and the input tiles of the two slots.
//   warp 1      TMEM allocator + warp-uniform tcgen05.mma issuer (UMMA 128 x 128 x 16, one elected lane).
//   warps 2..17 epilogue, two sets of eight warps (set s owns output boxes s and s+2):
//               thread = one output row x 16 of the 32 columns of a chunk.
//   warps 18,19 copy-out: wait for "box written", read their 64 rows of the swizzled box and stream
//               them to HBM with coalesced 16-byte st.global.
//
// Two tiles are in flight per CTA (slots A and B, strictly alternating MMA A(l) B(l) A(l+1) ... /
// epilogue A(l) B(l) ...), so one tile's hand-offs and epilogue run under the other tile's MMAs; inside a
// tile the epilogue releases its output box by box (64 columns, one mbarrier each), so the MMAs of layer
// l+1 over K-box b can start as soon as box b of layer l has been written.
//
// History and measurements (one tile per CTA with double-buffered accumulators, TMA-store and
// epilogue-copy store paths, cta_group::2 pair MMAs, the deletion profile that showed the ~3600-cycle
// MMA -> epilogue -> MMA round trip, and why none of it moves the wall time: the kernels run under the
// board's power cap): DESIGN.md section 5, profiles/r2_trunk_study.md, `git log -- upnerf_b200/csrc/mlp_fused.cu`.
#include <cuda_bf16.h>
#include <stdlib.h>
#include <string.h>

#include <type_traits>

#include "common.h"
#include "ptx_sm100.cuh"

namespace upnerf {
namespace {

using namespace ptx;

constexpr int kNL = UPNERF_TRUNK_LAYERS;  // 8 trunk layers + xyz_encoding_final
constexpr int kTileM = 128;
constexpr int kBoxBytes = kTileM * 128;   // 128 rows x 64 bf16 columns
constexpr int kActBytes = 4 * kBoxBytes;  // 128 x 256 activation tile
constexpr int kEpiThreads = 512;          // two epilogue sets of eight warps
constexpr int kCopyWarps = 2;             // dedicated copy-out warps
constexpr int kThreadsF = 64 + kEpiThreads + 32 * kCopyWarps;

struct LayerDesc {
  int w_act;    // first K column (in Wcat) of the weights multiplying the activation tile; -1: none
  int w_pe;     // K column of the weights multiplying the PE tile; -1: none
  int pe_last;  // last user of the PE tile within a sample tile
  int store;    // keep this layer's output in HBM (backward needs it; inference does not)
};

struct TrunkMaps {
  CUtensorMap w, pe;
};

struct TrunkArgs {
  int64_t M;
  int num_tiles;
  LayerDesc layer[kNL];
  const float* bias[kNL];
  const float* head_w;
  const float* head_b;
  float* head_out;
  uint32_t* relu_mask;  // [tiles][8 layers][8 chunks][128 rows] or nullptr
  __nv_bfloat16* out[kNL];
  int64_t ld_out[kNL];
};

__device__ __forceinline__ float softplus_ref(float x) {
  // torch.nn.Softplus(beta=1, threshold=20)
  return x > 20.f ? x : log1pf(expf(x));
}

// bf16x2 packing of two fp32 values (lo = first argument), plain and with ReLU folded into the conversion
__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {
  uint32_t d;
  asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(d) : "f"(hi), "f"(lo));
  return d;
}
__device__ __forceinline__ uint32_t pack_bf16x2_relu(float lo, float hi) {
  uint32_t d;
  asm("cvt.rn.relu.bf16x2.f32 %0, %1, %2;" : "=r"(d) : "f"(hi), "f"(lo));
  return d;
}
// ReLU bit masks (relu_mask): one 32-bit word per row and 32-column chunk, two 16-column groups of 16 bits; inside a
// group the even columns sit in the low byte and the odd columns in the high byte (the order in which the forward
// epilogue reads them off the packed bf16x2 registers).  Bit of column e (0..15) of a group:
__host__ __device__ constexpr int relu_mask_bit(int e) { return (e >> 1) + 8 * (e & 1); }

// Copy-out of one finished 128 x 64 bf16 box (the dedicated copy-out warps).  A warp owns
// kRowsPer = 128 / kCopyWarps rows: lane -> rows r0 + 4 i (i = 0 .. kRowsPer/4 - 1), 16-byte chunk lane % 8;
// row r keeps chunk c at slot c ^ (r & 7), and (r & 7) = lane / 8 for even i, lane / 8 + 4 for odd i.
// The loads of the box are issued eight at a time (immediate offsets off two base registers), the box is
// released as soon as the last ones are queued (shared-memory accesses of an SM retire in order, so the epilogue's
// next write cannot overtake them), then the rows stream out as coalesced 16-byte stores.  The common case --
// a full tile with the standard leading dimension of 256 -- addresses every store as base + immediate: the
// generic loop (a 64-bit multiply-add and a row predicate per store, ~13 instructions per 512-byte store)
// kept the two copy-out warps busy for a whole layer period and the epilogue waiting for released boxes
// (ncu source view: 13 % / 23 % of the forward / backward epilogue's time in the `box released' wait).
template <int kRows>
__device__ __forceinline__ void copy_box_out(uint32_t s_even, uint32_t s_odd, __nv_bfloat16* g, int64_t ld,
                                             bool fast, int64_t rows_avail, uint64_t* bar_free, int lane) {
  constexpr int kN = kRows / 4;          // 16-byte stores per lane and box
  constexpr int kP = kN < 8 ? kN : 8;    // loads in flight per pass (register budget of the epilogue's kernel)
#pragma unroll
  for (int p0 = 0; p0 < kN; p0 += kP) {
    float4 vv[kP];
#pragma unroll
    for (int i = 0; i < kP; ++i) vv[i] = lds128((((p0 + i) & 1) ? s_odd : s_even) + (p0 + i) * 512);
    if (p0 + kP >= kN) {
      __syncwarp();
      if (lane == 0) mbar_arrive(bar_free);   // all my reads of this box are queued: the epilogue may overwrite it
    }
    if (fast && ld == 256) {
      char* gb = reinterpret_cast<char*>(g);
#pragma unroll
      for (int i = 0; i < kP; ++i) __stcs(reinterpret_cast<float4*>(gb + (p0 + i) * (4 * 256 * 2)), vv[i]);
    } else if (fast) {
      __nv_bfloat16* gp = g + static_cast<int64_t>(p0 * 4) * ld;
      const int64_t step = 4 * ld;
#pragma unroll
      for (int i = 0; i < kP; ++i, gp += step) __stcs(reinterpret_cast<float4*>(gp), vv[i]);
    } else {
#pragma unroll
      for (int i = 0; i < kP; ++i)
        if ((p0 + i) * 4 < rows_avail)
          __stcs(reinterpret_cast<float4*>(g + static_cast<int64_t>((p0 + i) * 4) * ld), vv[i]);
    }
  }
}

// =====================================================================================
// Forward trunk, two tiles in flight per CTA ("ping-pong").  kCluster = 2: the two CTAs of a cluster walk their
// units in lockstep and share the weight stream -- each loads one half of every weight stage and TMA-multicasts it
// into both CTAs, which halves the L2 reads (1.1 MB of weights per 128-sample tile).
//
// Why two tiles: with one tile per CTA the MMAs of layer l+1 cannot start before the epilogue of layer l has handed
// the activation boxes back, and the epilogue cannot start before the MMAs are complete.  A deletion
// experiment (tools/trunk_variants.py: an epilogue that only waits, reads TMEM and arrives; copy-out warps that
// only release) measured the bare MMA -> epilogue -> MMA round trip at ~3600 cycles per layer and tile against
// 2048 cycles of MMA -- ~800 cycles of hand-off latency in each direction that nothing covers -- and 5300 with
// the real epilogue and stores.  Here a CTA owns tiles A and B (2 x 64 KB activations, one 256-column
// accumulator each: all 512 TMEM columns, so the accumulators are single-buffered and a `drained' barrier
// replaces the ping-pong of the single-tile kernel) and strictly alternates
//     MMA      : A(l)  B(l)  A(l+1)  B(l+1) ...
//     epilogue :       A(l)  B(l)    A(l+1) ...
// so one tile's hand-offs and epilogue run under the other tile's MMAs.  Shared memory: 128 KB activations +
// 2 x 16 KB encodings leave 48 KB for the weight ring, so a stage is one N-half of a K-chunk (128 output
// features x 64 K columns, 16 KB, UMMA 128 x 128 x 16); CTA pairs still share the weight stream by multicast.
namespace pp {

constexpr int kStagesP = 3;
constexpr int kWHalfBytes = 128 * 128;                    // 128 output features x 64 K columns
constexpr int kOffActP = 0;                               // [2 tiles][4 boxes]
constexpr int kOffPEP = kOffActP + 2 * kActBytes;         // [2 tiles]
constexpr int kOffWP = kOffPEP + 2 * kBoxBytes;
constexpr int kOffBiasP = kOffWP + kStagesP * kWHalfBytes;
constexpr int kOffHeadWP = kOffBiasP + kNL * 256 * 4;
constexpr int kOffHeadP = kOffHeadWP + 256 * 4;
constexpr int kOffBarP = kOffHeadP + 4 * kTileM * 4;
// wfull[3] wempty[3] pefull[2] peempty[2] act[2][4] tfull[2] tempty[2] st[2][4] stfree[2][4]
constexpr int kNumBarsP = 2 * kStagesP + 4 + 8 + 4 + 16;
constexpr int kOffTmemP = kOffBarP + kNumBarsP * 8;
constexpr int kSmemBytesP = kOffTmemP + 16 + 1024;
static_assert(kSmemBytesP <= 232448, "shared memory budget exceeded");

template <int kCluster>
__global__ void __launch_bounds__(kThreadsF, 1)
mlp_trunk_fwd_pp_kernel(const __grid_constant__ TrunkMaps maps, const __grid_constant__ TrunkArgs args) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                             ~static_cast<uintptr_t>(1023));
  uint8_t* sAct = smem + kOffActP;
  uint8_t* sPE = smem + kOffPEP;
  uint8_t* sW = smem + kOffWP;
  float* sBias = reinterpret_cast<float*>(smem + kOffBiasP);
  float* sHeadW = reinterpret_cast<float*>(smem + kOffHeadWP);
  float* sHead = reinterpret_cast<float*>(smem + kOffHeadP);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kOffBarP);
  uint64_t* bar_wfull = bars;
  uint64_t* bar_wempty = bar_wfull + kStagesP;
  uint64_t* bar_pefull = bar_wempty + kStagesP;   // [tile]
  uint64_t* bar_peempty = bar_pefull + 2;         // [tile]
  uint64_t* bar_act = bar_peempty + 2;            // [tile * 4 + box] box written: the 8 warps of its set
  uint64_t* bar_tfull = bar_act + 8;              // [tile] accumulator complete
  uint64_t* bar_tempty = bar_tfull + 2;           // [tile] accumulator read out: 16 epilogue warps
  uint64_t* bar_st = bar_tempty + 2;              // [tile * 4 + box] box written -> copy-out warps
  uint64_t* bar_stfree = bar_st + 8;              // [tile * 4 + box] box copied out
  uint32_t* tmem_holder = reinterpret_cast<uint32_t*>(smem + kOffTmemP);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int cta_rank = kCluster > 1 ? static_cast<int>(cluster_ctarank()) : 0;
  // a unit = 2 * kCluster tiles; CTA r of the cluster takes tiles (2 * kCluster) * unit + kCluster * x + r, x = 0, 1
  const int unit0 = blockIdx.x / kCluster;
  const int unit_step = gridDim.x / kCluster;
  const int num_units = (args.num_tiles + 2 * kCluster - 1) / (2 * kCluster);
  constexpr uint16_t kMask = static_cast<uint16_t>((1u << kCluster) - 1);
  auto tile_of = [&](int unit, int x) { return (unit * 2 + x) * kCluster + cta_rank; };

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&maps.w);
    prefetch_tmap(&maps.pe);
    for (int i = 0; i < kStagesP; ++i) {
      mbar_init(&bar_wfull[i], 1);
      mbar_init(&bar_wempty[i], kCluster);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&bar_pefull[i], 1);
      mbar_init(&bar_peempty[i], 1);
      mbar_init(&bar_tfull[i], 1);
      mbar_init(&bar_tempty[i], 16);
    }
    for (int i = 0; i < 8; ++i) {
      mbar_init(&bar_act[i], 8);
      mbar_init(&bar_st[i], 8);
      mbar_init(&bar_stfree[i], kCopyWarps);
    }
    fence_mbar_init();
  }
  if (warp == 1) tmem_alloc<512>(tmem_holder);
  if (warp >= 2 && warp < 18) {
    const int t = threadIdx.x - 64;
    for (int i = t; i < kNL * 256; i += kEpiThreads) {
      const float* b = args.bias[i >> 8];
      sBias[i] = b ? b[i & 255] : 0.f;
    }
    for (int i = t; i < 256; i += kEpiThreads) sHeadW[i] = args.head_w ? args.head_w[i] : 0.f;
  }
  tc_fence_before_sync();
  __syncthreads();
  if (kCluster > 1) cluster_sync_all();
  tc_fence_after_sync();
  const uint32_t tmem_base = *tmem_holder;

  if (warp == 0) {
    // ------------------------------------------------------------ TMA producer
    if (lane == 0) {
      int ws = 0;
      uint32_t wph = 0;
      auto load_w = [&](int kcol) {      // both N-halves of a K-chunk, one stage each
        for (int nh = 0; nh < 2; ++nh) {
          mbar_wait(&bar_wempty[ws], wph ^ 1);
          mbar_arrive_expect_tx(&bar_wfull[ws], kWHalfBytes);
          if (kCluster == 1) {
            tma_load_2d(sW + ws * kWHalfBytes, &maps.w, &bar_wfull[ws], kcol, nh * 128);
          } else {
            constexpr int kPart = kWHalfBytes / kCluster;   // my share of the stage: 128 / kCluster features
            tma_load_2d_mc(sW + ws * kWHalfBytes + cta_rank * kPart, &maps.w, &bar_wfull[ws], kcol,
                           nh * 128 + cta_rank * (128 / kCluster), kMask);
          }
          if (++ws == kStagesP) {
            ws = 0;
            wph ^= 1;
          }
        }
      };
      auto load_pe = [&](int t, int x, int tile) {   // t-th load into slot x
        mbar_wait(&bar_peempty[x], (t & 1) ^ 1);
        mbar_arrive_expect_tx(&bar_pefull[x], kBoxBytes);
        tma_load_2d(sPE + x * kBoxBytes, &maps.pe, &bar_pefull[x], 0, tile * kTileM);
      };
      int t = 0;
      if (unit0 < num_units) {
        load_pe(0, 0, tile_of(unit0, 0));
        load_pe(0, 1, tile_of(unit0, 1));
      }
      for (int unit = unit0; unit < num_units; unit += unit_step, ++t) {
        for (int l = 0; l < kNL; ++l) {
          const LayerDesc& L = args.layer[l];
          for (int x = 0; x < 2; ++x) {
            if (L.w_pe >= 0) load_w(L.w_pe);
            if (L.w_act >= 0)
              for (int b = 0; b < 4; ++b) load_w(L.w_act + b * 64);
            // the next unit's encoding tile: slot x was released by the skip layer's MMAs (layer 4)
            if (l == 6 && unit + unit_step < num_units) load_pe(t + 1, x, tile_of(unit + unit_step, x));
          }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------ MMA issuer (warp-uniform, one elected lane issues)
    const uint32_t idesc = umma_idesc_bf16(kTileM, 128, 0, 0);
    int ws = 0;
    uint32_t wph = 0;
    uint32_t act_ph[2] = {0, 0};
    int t = 0;
    auto free_stage = [&](uint64_t* bar) {
      if (elect_one()) {
        if (kCluster == 1) mma_commit(bar);
        else mma_commit_mc(bar, kMask);
      }
      __syncwarp();
    };
    auto commit_local = [&](uint64_t* bar) {
      if (elect_one()) mma_commit(bar);
      __syncwarp();
    };
    // one K-chunk (64 columns of A at a_addr) against both N-halves (two weight stages)
    auto chunk = [&](uint32_t d_tmem, uint32_t a_addr, uint32_t accum) {
      for (int nh = 0; nh < 2; ++nh) {
        mbar_wait(&bar_wfull[ws], wph);
        tc_fence_after_sync();
        const uint32_t b_addr = smem_u32(sW + ws * kWHalfBytes);
        if (elect_one()) {
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            const uint64_t da = umma_desc(a_addr + k * 32, 16, 1024, kLayoutSw128);
            const uint64_t db = umma_desc(b_addr + k * 32, 16, 1024, kLayoutSw128);
            mma_bf16_ss(d_tmem + nh * 128, da, db, idesc, (k > 0) ? 1u : accum);
          }
        }
        __syncwarp();
        free_stage(&bar_wempty[ws]);
        if (++ws == kStagesP) {
          ws = 0;
          wph ^= 1;
        }
      }
    };
    for (int unit = unit0; unit < num_units; unit += unit_step, ++t) {
      for (int l = 0; l < kNL; ++l) {
        const LayerDesc& L = args.layer[l];
        const int lay = t * kNL + l;          // layers issued so far per tile slot
        for (int x = 0; x < 2; ++x) {
          const uint32_t d_tmem = tmem_base + x * 256;
          // the accumulator of slot x is single-buffered: the epilogue of its previous layer must have read it out
          if (lay > 0) mbar_wait(&bar_tempty[x], (lay - 1) & 1);
          tc_fence_after_sync();
          uint32_t accum = 0;
          if (L.w_pe >= 0) {
            mbar_wait(&bar_pefull[x], t & 1);
            chunk(d_tmem, smem_u32(sPE + x * kBoxBytes), accum);
            accum = 1;
            if (L.pe_last) commit_local(&bar_peempty[x]);
          }
          if (L.w_act >= 0) {
#pragma unroll 1
            for (int b = 0; b < 4; ++b) {
              mbar_wait(&bar_act[x * 4 + b], act_ph[x]);
              chunk(d_tmem, smem_u32(sAct + x * kActBytes + b * kBoxBytes), accum);
              accum = 1;
            }
            act_ph[x] ^= 1;
          }
          commit_local(&bar_tfull[x]);
        }
      }
    }
  } else if (warp >= 18) {
    // ------------------------------------------------------------ copy-out warps
    const int cw = warp - 18;
    constexpr int kRowsPer = kTileM / kCopyWarps;
    const int r0 = cw * kRowsPer + (lane >> 3);
    const uint32_t x0 = static_cast<uint32_t>((lane & 7) ^ (lane >> 3)) << 4;
    const uint32_t so_even = smem_u32(sAct) + r0 * 128 + x0, so_odd = smem_u32(sAct) + r0 * 128 + (x0 ^ 64u);
    uint32_t ph = 0;
    for (int unit = unit0; unit < num_units; unit += unit_step) {
      for (int l = 0; l < kNL; ++l) {
        if (!args.layer[l].store) continue;
        const int64_t ldo = args.ld_out[l];
        for (int x = 0; x < 2; ++x) {
          const int tile = tile_of(unit, x);
          const int64_t rows_avail = args.M - static_cast<int64_t>(tile) * kTileM - r0;
          const bool fast = rows_avail >= kRowsPer;
          __nv_bfloat16* o0 = args.out[l] + (static_cast<int64_t>(tile) * kTileM + r0) * ldo + (lane & 7) * 8;
#pragma unroll 1
          for (int b = 0; b < 4; ++b) {
            mbar_wait(&bar_st[x * 4 + b], ph);
            const uint32_t off = x * kActBytes + b * kBoxBytes;
            copy_box_out<kRowsPer>(so_even + off, so_odd + off, o0 + b * 64, ldo, fast, rows_avail,
                                   &bar_stfree[x * 4 + b], lane);
          }
        }
        ph ^= 1;
      }
    }
  } else {
    // ------------------------------------------------------------ epilogue warps
    // Two sets of eight warps; set s owns the 64-column boxes s and s+2 of every layer, warp (quad, half) handles
    // rows quad*32.. and the 32-column chunk `half` of a box as two 16-column steps.  With two tiles in flight these
    // warps are the kernel's bottleneck (ncu: 81 % busy, MMA pipe 40 %), so the per-element path is kept short:
    //  * layer kinds are compile-time (plain ReLU layer / ReLU + sigma head / final linear): no flag tests inside;
    //  * ReLU rides on the bf16 conversion (cvt.rn.relu.bf16x2.f32) and the ReLU bit mask is read off the PACKED
    //    result -- x + 0x7fff7fff carries a non-zero 16-bit half into its top bit -- 3 instructions per column
    //    pair instead of 2 x (max, negate, funnel shift).  Bit order inside a 16-column group: even columns in the
    //    low byte, odd columns in the high byte (relu_mask_bit(); the backward kernels decode the same way);
    //  * the bias of the NEXT 16 columns is requested right after the adds that consumed the current one;
    //  * the mask words leave after the last hand-off (a global store in flight stalls the next proxy fence).
    const int ew = warp - 2;
    const int grp = ew >> 2;
    const int half = grp & 1;
    const int set = grp >> 1;
    const int quad = warp & 3;
    const int row = quad * 32 + lane;
    const bool lane0 = lane_id() == 0;
    const uint32_t sbias = smem_u32(sBias);
    const uint32_t sheadw = smem_u32(sHeadW);
    const uint32_t swz = row & 7;
    const bool have_mask = args.relu_mask != nullptr;
    uint32_t nst = 0;   // stored layers so far (= releases seen per box and slot)
    int t = 0;
    // one layer of one slot; kind: 0 = Linear + ReLU, 1 = Linear + ReLU + share_sigma head, 2 = Linear (final)
    auto epi = [&](auto kind_c, int l, int x, int unit, uint32_t lay, int store) {
      constexpr int kKind = decltype(kind_c)::value;
      constexpr bool kRelu = kKind != 2, kHead = kKind == 1, kFeeds = kKind != 2;
      const int tile = tile_of(unit, x);
      const uint32_t sact_row = smem_u32(sAct) + x * kActBytes + row * 128;
      const uint32_t bias_l = sbias + l * 1024;
      mbar_wait(&bar_tfull[x], lay & 1);
      tc_fence_after_sync();
      const uint32_t taddr = tmem_base + (static_cast<uint32_t>(quad * 32) << 16) + x * 256;
      uint32_t mw[2] = {0, 0};
      float hacc = 0.f;
      uint32_t r[2][16];
      tmem_ld_32x16(taddr + (2 * set + half) * 32, r[0]);
      float4 b4[4];
#pragma unroll
      for (int k = 0; k < 4; ++k) b4[k] = lds128(bias_l + (set * 64 + half * 32 + k * 4) * 4);
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const int box = set + (q & 2);
        const int col0 = box * 64 + half * 32 + (q & 1) * 16;
        // the previous contents of this box must have been copied out before my first write
        if ((q & 1) == 0 && nst) mbar_wait(&bar_stfree[x * 4 + box], (nst - 1) & 1);
        tmem_ld_wait_dep(r[q & 1]);
        const uint32_t* rr = r[q & 1];
        float v[16];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          v[k * 4 + 0] = __uint_as_float(rr[k * 4 + 0]) + b4[k].x;
          v[k * 4 + 1] = __uint_as_float(rr[k * 4 + 1]) + b4[k].y;
          v[k * 4 + 2] = __uint_as_float(rr[k * 4 + 2]) + b4[k].z;
          v[k * 4 + 3] = __uint_as_float(rr[k * 4 + 3]) + b4[k].w;
        }
        if (q < 3) {
          const int ncol0 = (set + ((q + 1) & 2)) * 64 + half * 32 + ((q + 1) & 1) * 16;
          tmem_ld_32x16(taddr + ncol0, r[(q + 1) & 1]);
#pragma unroll
          for (int k = 0; k < 4; ++k) b4[k] = lds128(bias_l + (ncol0 + k * 4) * 4);
        } else {
          // my last read of this accumulator is in registers: the MMAs of this slot's next layer may overwrite it
          tc_fence_before_sync();
          __syncwarp();
          if (lane0) mbar_arrive(&bar_tempty[x]);
        }
        uint32_t o[8];
        if (kHead) {
#pragma unroll
          for (int e = 0; e < 16; ++e) v[e] = fmaxf(v[e], 0.f);
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            const float4 w4 = lds128(sheadw + (col0 + k * 4) * 4);
            hacc += v[k * 4] * w4.x + v[k * 4 + 1] * w4.y + v[k * 4 + 2] * w4.z + v[k * 4 + 3] * w4.w;
          }
        }
#pragma unroll
        for (int e = 0; e < 8; ++e) o[e] = (kRelu && !kHead) ? pack_bf16x2_relu(v[2 * e], v[2 * e + 1]) : pack_bf16x2(v[2 * e], v[2 * e + 1]);
        const uint32_t box_row = sact_row + box * kBoxBytes;
        const uint32_t s0 = half * 4 + (q & 1) * 2;
        sts128(box_row + ((s0 ^ swz) << 4), make_uint4(o[0], o[1], o[2], o[3]));
        sts128(box_row + (((s0 + 1) ^ swz) << 4), make_uint4(o[4], o[5], o[6], o[7]));
        if (kRelu && have_mask) {
          uint32_t acc = 0;
#pragma unroll
          for (int e = 0; e < 8; ++e) acc = (acc >> 1) | ((o[e] + 0x7fff7fffu) & 0x80008000u);
          const uint32_t m16 = ((acc >> 16) & 0xff00u) | ((acc >> 8) & 0xffu);
          mw[q >> 1] = (q & 1) ? (mw[q >> 1] | (m16 << 16)) : m16;
        }
        if (q & 1) {
          fence_proxy_async_smem();   // my writes -> visible to the async proxy (the next layer's MMAs)
          if (kFeeds) {
            tc_fence_before_sync();
            __syncwarp();
            if (lane0) mbar_arrive(&bar_act[x * 4 + box]);
          }
          if (store) {
            __syncwarp();
            if (lane0) mbar_arrive(&bar_st[x * 4 + box]);
          }
        }
      }
      if (kRelu && have_mask) {
        uint32_t* mrow = args.relu_mask + ((static_cast<int64_t>(tile) * 8 + l) * 8 + half) * kTileM + row;
        mrow[(set * 2) * kTileM] = mw[0];
        mrow[((set + 2) * 2) * kTileM] = mw[1];
      }
      if (kHead) {
        const int64_t grow = static_cast<int64_t>(tile) * kTileM + row;
        sHead[grp * kTileM + row] = hacc;
        named_bar_sync(4, kEpiThreads);
        if (grp == 0 && grow < args.M)
          args.head_out[grow] = softplus_ref(sHead[row] + sHead[kTileM + row] + sHead[2 * kTileM + row] +
                                             sHead[3 * kTileM + row] + args.head_b[0]);
        named_bar_sync(4, kEpiThreads);   // the other slot's partial sums follow in the same buffer
      }
    };
    for (int unit = unit0; unit < num_units; unit += unit_step, ++t) {
      for (int l = 0; l < kNL; ++l) {
        const int store = args.layer[l].store;
        const uint32_t lay = static_cast<uint32_t>(t * kNL + l);
        for (int x = 0; x < 2; ++x) {
          if (l < 7) epi(std::integral_constant<int, 0>{}, l, x, unit, lay, store);
          else if (l == 7) epi(std::integral_constant<int, 1>{}, l, x, unit, lay, store);
          else epi(std::integral_constant<int, 2>{}, l, x, unit, lay, store);
        }
        if (store) ++nst;
      }
    }
  }

  tc_fence_before_sync();
  __syncthreads();
  if (kCluster > 1) cluster_sync_all();
  if (warp == 1) {
    tc_fence_after_sync();
    tmem_dealloc<512>(tmem_base);
  }
}

}  // namespace pp

// =====================================================================================
// Backward data-gradient chain of the trunk -- upnerf_mlp_trunk_bwd_bf16.
//
//   dY8 = (dHF . W_final + d_ssig (x) w_sigma) * [H8 > 0]          (layer index j = 0)
//   dYl = (dY(l+1) . W(l+1)[:, h part]) * [Hl > 0],  l = 7..1      (j = 1..7)
//
// i.e. the autograd backward of models/nerf.py:84-93 with respect to the activations; the
// weight gradients are separate launches that read the dY tensors this kernel stores.  The ReLU
// masks are the bit masks the forward kernel wrote (32 B per sample and layer instead of re-reading
// 512 B of activations).
namespace bwd {

constexpr int kNLb = UPNERF_TRUNK_BWD_LAYERS;  // 8

struct BwdMaps {
  CUtensorMap w, in;
};
struct BwdArgs {
  int64_t M;
  int num_tiles;
  const float* d_ssig;
  const float* sig_w;
  const uint32_t* relu_mask;
  __nv_bfloat16* out[UPNERF_TRUNK_BWD_LAYERS];
  int64_t ld_out[UPNERF_TRUNK_BWD_LAYERS];
};

}  // namespace bwd

// =====================================================================================
// Backward chain, two tiles in flight per CTA (see pp::mlp_trunk_fwd_pp_kernel) -- the default backward kernel.
// The incoming dHF tile of a slot lands directly in the slot's activation buffer (it is layer 0's A operand and
// layer 0's epilogue overwrites it in place); the buffer is refilled box by box as the copy-out warps release
// the last layer's boxes.  No separate input buffer, so the weight ring has six 16 KB stages.
namespace pp {

constexpr int kStagesB = 6;
constexpr int kOffActBP = 0;                               // [2 slots][4 boxes]
constexpr int kOffWBP = kOffActBP + 2 * kActBytes;
constexpr int kOffSigWP = kOffWBP + kStagesB * kWHalfBytes;
constexpr int kOffBarBP = kOffSigWP + 256 * 4;
// wfull[6] wempty[6] infull[2] infree[2] act[2][4] tfull[2] tempty[2] st[2][4] stfree[2][4]
constexpr int kNumBarsBP = 2 * kStagesB + 4 + 8 + 4 + 16;
constexpr int kOffTmemBP = kOffBarBP + kNumBarsBP * 8;
constexpr int kSmemBytesBP = kOffTmemBP + 16 + 1024;
static_assert(kSmemBytesBP <= 232448, "shared memory budget exceeded");

template <int kCluster>
__global__ void __launch_bounds__(kThreadsF, 1)
mlp_trunk_bwd_pp_kernel(const __grid_constant__ bwd::BwdMaps maps, const __grid_constant__ bwd::BwdArgs args) {
  constexpr int kNLb = bwd::kNLb;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                             ~static_cast<uintptr_t>(1023));
  uint8_t* sAct = smem + kOffActBP;
  uint8_t* sW = smem + kOffWBP;
  float* sSigW = reinterpret_cast<float*>(smem + kOffSigWP);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kOffBarBP);
  uint64_t* bar_wfull = bars;
  uint64_t* bar_wempty = bar_wfull + kStagesB;
  uint64_t* bar_infull = bar_wempty + kStagesB;   // [slot] dHF tile landed
  uint64_t* bar_infree = bar_infull + 2;          // [slot] last layer of the unit copied out: the buffer may be refilled
  uint64_t* bar_act = bar_infree + 2;             // [slot * 4 + box]
  uint64_t* bar_tfull = bar_act + 8;              // [slot]
  uint64_t* bar_tempty = bar_tfull + 2;           // [slot]
  uint64_t* bar_st = bar_tempty + 2;              // [slot * 4 + box]
  uint64_t* bar_stfree = bar_st + 8;              // [slot * 4 + box]
  uint32_t* tmem_holder = reinterpret_cast<uint32_t*>(smem + kOffTmemBP);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int cta_rank = kCluster > 1 ? static_cast<int>(cluster_ctarank()) : 0;
  const int unit0 = blockIdx.x / kCluster;
  const int unit_step = gridDim.x / kCluster;
  const int num_units = (args.num_tiles + 2 * kCluster - 1) / (2 * kCluster);
  constexpr uint16_t kMask = static_cast<uint16_t>((1u << kCluster) - 1);
  auto tile_of = [&](int unit, int x) { return (unit * 2 + x) * kCluster + cta_rank; };

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&maps.w);
    prefetch_tmap(&maps.in);
    for (int i = 0; i < kStagesB; ++i) {
      mbar_init(&bar_wfull[i], 1);
      mbar_init(&bar_wempty[i], kCluster);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&bar_infull[i], 1);
      mbar_init(&bar_infree[i], kCopyWarps);
      mbar_init(&bar_tfull[i], 1);
      mbar_init(&bar_tempty[i], 16);
    }
    for (int i = 0; i < 8; ++i) {
      mbar_init(&bar_act[i], 8);
      mbar_init(&bar_st[i], 8);
      mbar_init(&bar_stfree[i], kCopyWarps);
    }
    fence_mbar_init();
  }
  if (warp == 1) tmem_alloc<512>(tmem_holder);
  if (warp >= 2 && warp < 18)
    for (int i = threadIdx.x - 64; i < 256; i += kEpiThreads) sSigW[i] = args.sig_w ? args.sig_w[i] : 0.f;
  tc_fence_before_sync();
  __syncthreads();
  if (kCluster > 1) cluster_sync_all();
  tc_fence_after_sync();
  const uint32_t tmem_base = *tmem_holder;

  if (warp == 0) {
    // ------------------------------------------------------------ TMA producer
    if (lane == 0) {
      int ws = 0;
      uint32_t wph = 0;
      auto load_w = [&](int kcol) {
        for (int nh = 0; nh < 2; ++nh) {
          mbar_wait(&bar_wempty[ws], wph ^ 1);
          mbar_arrive_expect_tx(&bar_wfull[ws], kWHalfBytes);
          if (kCluster == 1) {
            tma_load_2d(sW + ws * kWHalfBytes, &maps.w, &bar_wfull[ws], kcol, nh * 128);
          } else {
            constexpr int kPart = kWHalfBytes / kCluster;
            tma_load_2d_mc(sW + ws * kWHalfBytes + cta_rank * kPart, &maps.w, &bar_wfull[ws], kcol,
                           nh * 128 + cta_rank * (128 / kCluster), kMask);
          }
          if (++ws == kStagesB) {
            ws = 0;
            wph ^= 1;
          }
        }
      };
      int t = 0;
      for (int unit = unit0; unit < num_units; unit += unit_step, ++t) {
        for (int j = 0; j < kNLb; ++j) {
          for (int x = 0; x < 2; ++x) {
            if (j == 0) {
              // dHF tile of this slot: the buffer may be refilled once the copy-out warps have read the previous
              // unit's last layer out of it (a barrier of its own, one phase per unit: this thread runs several
              // stages ahead of the copy-out warps, and a parity wait must never be more than one phase early)
              if (t > 0) mbar_wait(&bar_infree[x], (t - 1) & 1);
              mbar_arrive_expect_tx(&bar_infull[x], kActBytes);
              for (int b = 0; b < 4; ++b) {
                tma_load_2d(sAct + x * kActBytes + b * kBoxBytes, &maps.in, &bar_infull[x], b * 64,
                            tile_of(unit, x) * kTileM);
              }
            }
            for (int b = 0; b < 4; ++b) load_w(j * 256 + b * 64);
          }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------ MMA issuer (warp-uniform, one elected lane issues)
    const uint32_t idesc = umma_idesc_bf16(kTileM, 128, 0, 0);
    int ws = 0;
    uint32_t wph = 0;
    uint32_t act_ph[2] = {0, 0};
    int t = 0;
    auto free_stage = [&](uint64_t* bar) {
      if (elect_one()) {
        if (kCluster == 1) mma_commit(bar);
        else mma_commit_mc(bar, kMask);
      }
      __syncwarp();
    };
    for (int unit = unit0; unit < num_units; unit += unit_step, ++t) {
      for (int j = 0; j < kNLb; ++j) {
        const int lay = t * kNLb + j;
        for (int x = 0; x < 2; ++x) {
          const uint32_t d_tmem = tmem_base + x * 256;
          if (lay > 0) mbar_wait(&bar_tempty[x], (lay - 1) & 1);
          if (j == 0) mbar_wait(&bar_infull[x], t & 1);
#pragma unroll 1
          for (int b = 0; b < 4; ++b) {
            if (j > 0) mbar_wait(&bar_act[x * 4 + b], act_ph[x]);
            const uint32_t a_addr = smem_u32(sAct + x * kActBytes + b * kBoxBytes);
            for (int nh = 0; nh < 2; ++nh) {
              mbar_wait(&bar_wfull[ws], wph);
              tc_fence_after_sync();
              const uint32_t b_addr = smem_u32(sW + ws * kWHalfBytes);
              if (elect_one()) {
#pragma unroll
                for (int k = 0; k < 4; ++k)
                  mma_bf16_ss(d_tmem + nh * 128, umma_desc(a_addr + k * 32, 16, 1024, kLayoutSw128),
                              umma_desc(b_addr + k * 32, 16, 1024, kLayoutSw128), idesc, (b > 0 || k > 0) ? 1u : 0u);
              }
              __syncwarp();
              free_stage(&bar_wempty[ws]);
              if (++ws == kStagesB) {
                ws = 0;
                wph ^= 1;
              }
            }
          }
          if (j > 0) act_ph[x] ^= 1;
          if (elect_one()) mma_commit(&bar_tfull[x]);
          __syncwarp();
        }
      }
    }
  } else if (warp >= 18) {
    // ------------------------------------------------------------ copy-out warps
    const int cw = warp - 18;
    constexpr int kRowsPer = kTileM / kCopyWarps;
    const int r0 = cw * kRowsPer + (lane >> 3);
    const uint32_t x0 = static_cast<uint32_t>((lane & 7) ^ (lane >> 3)) << 4;
    const uint32_t so_even = smem_u32(sAct) + r0 * 128 + x0, so_odd = smem_u32(sAct) + r0 * 128 + (x0 ^ 64u);
    uint32_t ph = 0;
    for (int unit = unit0; unit < num_units; unit += unit_step) {
      for (int j = 0; j < kNLb; ++j) {
        const int64_t ldo = args.ld_out[j];
        for (int x = 0; x < 2; ++x) {
          const int tile = tile_of(unit, x);
          const int64_t rows_avail = args.M - static_cast<int64_t>(tile) * kTileM - r0;
          const bool fast = rows_avail >= kRowsPer;
          __nv_bfloat16* o0 = args.out[j] + (static_cast<int64_t>(tile) * kTileM + r0) * ldo + (lane & 7) * 8;
#pragma unroll 1
          for (int b = 0; b < 4; ++b) {
            mbar_wait(&bar_st[x * 4 + b], ph);
            const uint32_t off = x * kActBytes + b * kBoxBytes;
            copy_box_out<kRowsPer>(so_even + off, so_odd + off, o0 + b * 64, ldo, fast, rows_avail,
                                   &bar_stfree[x * 4 + b], lane);
          }
          if (j == kNLb - 1) {
            __syncwarp();
            if (lane == 0) mbar_arrive(&bar_infree[x]);
          }
        }
        ph ^= 1;
      }
    }
  } else {
    // ------------------------------------------------------------ epilogue warps
    const int ew = warp - 2;
    const int grp = ew >> 2;
    const int half = grp & 1;
    const int set = grp >> 1;
    const int quad = warp & 3;
    const int row = quad * 32 + lane;
    const bool lane0 = lane_id() == 0;
    const uint32_t ssigw = smem_u32(sSigW);
    const uint32_t swz = row & 7;
    uint32_t nst = 0;
    int t = 0;
    for (int unit = unit0; unit < num_units; unit += unit_step, ++t) {
      for (int j = 0; j < kNLb; ++j) {
        const uint32_t lay = static_cast<uint32_t>(t * kNLb + j);
        for (int x = 0; x < 2; ++x) {
          const int tile = tile_of(unit, x);
          const int64_t grow = static_cast<int64_t>(tile) * kTileM + row;
          const bool in_range = tile < args.num_tiles;
          const float ds = (j == 0 && args.d_ssig && grow < args.M) ? args.d_ssig[grow] : 0.f;
          const uint32_t sact_row = smem_u32(sAct) + x * kActBytes + row * 128;
          // ReLU mask words of H(8-j) for my two chunks, requested before the accumulator wait
          const uint32_t* mrow =
              args.relu_mask + ((static_cast<int64_t>(in_range ? tile : 0) * 8 + (7 - j)) * 8) * kTileM + row;
          const uint32_t mA = __ldg(mrow + (set * 2 + half) * kTileM);
          const uint32_t mB = __ldg(mrow + ((set + 2) * 2 + half) * kTileM);
          mbar_wait(&bar_tfull[x], lay & 1);
          tc_fence_after_sync();
          const uint32_t taddr = tmem_base + (static_cast<uint32_t>(quad * 32) << 16) + x * 256;
          uint32_t r[2][16];
          tmem_ld_32x16(taddr + (2 * set + half) * 32, r[0]);
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            const int box = set + (q & 2);
            const int col0 = box * 64 + half * 32 + (q & 1) * 16;
            if ((q & 1) == 0 && nst) mbar_wait(&bar_stfree[x * 4 + box], (nst - 1) & 1);
            tmem_ld_wait_dep(r[q & 1]);
            const uint32_t* rr = r[q & 1];
            const uint32_t m16 = ((q & 2) ? mB : mA) >> ((q & 1) * 16);
            float v[16];
#pragma unroll
            for (int e = 0; e < 16; ++e) v[e] = __uint_as_float(rr[e]);
            if (q < 3) {
              const int nbox = set + ((q + 1) & 2);
              tmem_ld_32x16(taddr + nbox * 64 + half * 32 + ((q + 1) & 1) * 16, r[(q + 1) & 1]);
            } else {
              tc_fence_before_sync();
              __syncwarp();
              if (lane0) mbar_arrive(&bar_tempty[x]);   // the accumulator of this slot is in registers
            }
            if (j == 0) {
              // + d_ssig (x) w_sigma : the share_sigma head hangs off H8 (models/nerf.py:89)
#pragma unroll
              for (int k = 0; k < 4; ++k) {
                const float4 w4 = lds128(ssigw + (col0 + k * 4) * 4);
                v[k * 4 + 0] += ds * w4.x;
                v[k * 4 + 1] += ds * w4.y;
                v[k * 4 + 2] += ds * w4.z;
                v[k * 4 + 3] += ds * w4.w;
              }
            }
#pragma unroll
            for (int e = 0; e < 16; ++e) v[e] = ((m16 >> relu_mask_bit(e)) & 1u) ? v[e] : 0.f;
            uint4 o[2];
            __nv_bfloat162* ob = reinterpret_cast<__nv_bfloat162*>(o);
#pragma unroll
            for (int e = 0; e < 8; ++e) ob[e] = __floats2bfloat162_rn(v[2 * e], v[2 * e + 1]);
            const uint32_t box_row = sact_row + box * kBoxBytes;
            const uint32_t s0 = half * 4 + (q & 1) * 2;
            sts128(box_row + ((s0 ^ swz) << 4), o[0]);
            sts128(box_row + (((s0 + 1) ^ swz) << 4), o[1]);
            if (q & 1) {
              fence_proxy_async_smem();
              if (j < kNLb - 1) {
                tc_fence_before_sync();
                __syncwarp();
                if (lane0) mbar_arrive(&bar_act[x * 4 + box]);
              }
              __syncwarp();
              if (lane0) mbar_arrive(&bar_st[x * 4 + box]);
            }
          }
        }
        ++nst;
      }
    }
  }

  tc_fence_before_sync();
  __syncthreads();
  if (kCluster > 1) cluster_sync_all();
  if (warp == 1) {
    tc_fence_after_sync();
    tmem_dealloc<512>(tmem_base);
  }
}

}  // namespace pp

// 2 (default): CTA pairs share the weight stream by TMA multicast; UPNERF_TRUNK_CLUSTER=1
// selects independent CTAs (for A/B measurements and the tests; read on every call).
int trunk_cluster_size() {
  const char* e = getenv("UPNERF_TRUNK_CLUSTER");
  return (e && e[0] == '1') ? 1 : 2;
}
}  // namespace
}  // namespace upnerf

// Launch of either trunk kernel: one cluster of `cluster` CTAs per unit of 2 * cluster tiles, at most one CTA per SM
template <typename K1, typename K2, typename Maps, typename Args>
static int launch_trunk(K1 k1, K2 k2, int cluster, int64_t tiles, int smem_bytes, cudaStream_t st, const Maps& maps,
                        const Args& args) {
  using namespace upnerf;
  const int64_t units = ceil_div64(tiles, 2 * cluster);
  if (cluster == 1) {
    const int grid = static_cast<int>(units < sm_count() ? units : sm_count());
    k1<<<grid, kThreadsF, smem_bytes, st>>>(maps, args);
    return UPNERF_OK;
  }
  const int max_clusters = sm_count() / 2;
  const int clusters = static_cast<int>(units < max_clusters ? units : max_clusters);
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = dim3(2 * clusters);
  cfg.blockDim = dim3(kThreadsF);
  cfg.dynamicSmemBytes = smem_bytes;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 2;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  UPNERF_CHECK_CUDA(cudaLaunchKernelEx(&cfg, k2, maps, args));
  return UPNERF_OK;
}

extern "C" int upnerf_mlp_trunk_fwd_bf16(const upnerf_trunk_args* a, void* stream) {
  using namespace upnerf;
  UPNERF_REQUIRE(a && a->M > 0, UPNERF_ERR_BAD_SHAPE, "mlp_trunk_fwd: M=%lld", a ? (long long)a->M : -1ll);
  UPNERF_REQUIRE(a->pe && a->wcat && a->s_sigma && a->sigma_w && a->sigma_b && a->out[kNL - 1],
                 UPNERF_ERR_BAD_SHAPE, "mlp_trunk_fwd: missing operand");
  UPNERF_REQUIRE(a->ld_w >= UPNERF_TRUNK_WCAT_COLS, UPNERF_ERR_BAD_SHAPE, "mlp_trunk_fwd: ld_w=%lld < %d",
                 (long long)a->ld_w, UPNERF_TRUNK_WCAT_COLS);
  const int64_t tiles = ceil_div64(a->M, kTileM);
  UPNERF_REQUIRE(tiles < (1ll << 24), UPNERF_ERR_BAD_SHAPE, "mlp_trunk_fwd: M too large");

  TrunkMaps maps;  // rebuilt on every call (host-side encode only)
  TrunkArgs args;
  memset(&args, 0, sizeof(args));
  args.M = a->M;
  args.num_tiles = static_cast<int>(tiles);
  const int cluster = trunk_cluster_size();
  // a weight stage is one N-half of a K-chunk: 128 / cluster output features per TMA box (multicast across the pair)
  UPNERF_TRY(make_tmap_bf16_2d(&maps.w, a->wcat, 256, UPNERF_TRUNK_WCAT_COLS, a->ld_w, 128 / cluster, 64));
  UPNERF_TRY(make_tmap_bf16_2d(&maps.pe, a->pe, a->M, 64, a->ld_pe, kTileM, 64));
  // Wcat columns: [W1 (64) | W2 | W3 | W4 | W5 = [h (256) | PE (64)] | W6 | W7 | W8 | WF]
  const int w_act[kNL] = {-1, 64, 320, 576, 832, 1152, 1408, 1664, 1920};
  const int w_pe[kNL] = {0, -1, -1, -1, 1088, -1, -1, -1, -1};
  int n_store = 0;
  for (int l = 0; l < kNL; ++l) {
    LayerDesc& L = args.layer[l];
    L.store = a->out[l] != nullptr;
    n_store += L.store;
    L.w_act = w_act[l];
    L.w_pe = w_pe[l];
    // (the epilogue's layer kinds are compile-time: Linear + ReLU for l = 0..6, + share_sigma head for l = 7, linear l = 8)
    L.pe_last = l == 4;
    args.bias[l] = a->bias[l];
    args.out[l] = static_cast<__nv_bfloat16*>(a->out[l]);
    args.ld_out[l] = a->ld_out[l];
    // the copy-out warps store 16-byte vectors: every row must start 16-byte aligned
    UPNERF_REQUIRE(!a->out[l] || ((reinterpret_cast<uintptr_t>(a->out[l]) & 15) == 0 && (a->ld_out[l] & 7) == 0),
                   UPNERF_ERR_BAD_SHAPE, "mlp_trunk_fwd: out[%d] must be 16-byte aligned with ld %% 8 == 0", l);
  }
  args.head_w = a->sigma_w;
  args.head_b = a->sigma_b;
  args.head_out = a->s_sigma;
  args.relu_mask = a->relu_mask;
  static bool attr_set = false;
  if (!attr_set) {
    UPNERF_CHECK_CUDA(cudaFuncSetAttribute(pp::mlp_trunk_fwd_pp_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                           pp::kSmemBytesP));
    UPNERF_CHECK_CUDA(cudaFuncSetAttribute(pp::mlp_trunk_fwd_pp_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                           pp::kSmemBytesP));
    attr_set = true;
  }
  const double flop = 2.0 * a->M * 256.0 * (64 + 7 * 256 + 320);
  // algorithmic traffic: read PE, write the stored activation tensors + the ReLU bit masks
  const double bytes = 2.0 * a->M * (64 + n_store * 256) + (a->relu_mask ? 8.0 * 32 * a->M : 0.0);
  LaunchScope scope(kCatTrunkFwd, as_stream(stream), flop, bytes);
  UPNERF_TRY(launch_trunk(pp::mlp_trunk_fwd_pp_kernel<1>, pp::mlp_trunk_fwd_pp_kernel<2>, cluster, tiles,
                          pp::kSmemBytesP, as_stream(stream), maps, args));
  UPNERF_CHECK_LAUNCH("mlp_trunk_fwd_pp_kernel");
  return UPNERF_OK;
}

extern "C" int64_t upnerf_trunk_mask_words(int64_t M) {
  // tiles are processed in units of CTA pairs: a CTA whose tile index is past the end writes (all-zero) masks
  // for a phantom tile, so the buffer covers a whole number of units (rounded to four tiles: the ABI's historic size)
  const int64_t tiles = upnerf::ceil_div64(M, upnerf::kTileM);
  return ((tiles + 3) / 4) * 4 * 8 * 8 * upnerf::kTileM;
}

extern "C" int upnerf_mlp_trunk_bwd_bf16(const upnerf_trunk_bwd_args* a, void* stream) {
  using namespace upnerf;
  using namespace upnerf::bwd;
  UPNERF_REQUIRE(a && a->M > 0, UPNERF_ERR_BAD_SHAPE, "mlp_trunk_bwd: M=%lld", a ? (long long)a->M : -1ll);
  UPNERF_REQUIRE(a->d_hf && a->wcat_t && a->relu_mask && a->sigma_w, UPNERF_ERR_BAD_SHAPE,
                 "mlp_trunk_bwd: missing operand");
  UPNERF_REQUIRE(a->ld_w >= UPNERF_TRUNK_WCATT_COLS, UPNERF_ERR_BAD_SHAPE, "mlp_trunk_bwd: ld_w=%lld < %d",
                 (long long)a->ld_w, UPNERF_TRUNK_WCATT_COLS);
  const int64_t tiles = ceil_div64(a->M, kTileM);
  UPNERF_REQUIRE(tiles < (1ll << 24), UPNERF_ERR_BAD_SHAPE, "mlp_trunk_bwd: M too large");
  BwdMaps maps;
  BwdArgs args;
  memset(&args, 0, sizeof(args));
  args.M = a->M;
  args.num_tiles = static_cast<int>(tiles);
  args.d_ssig = a->d_ssig;
  args.sig_w = a->sigma_w;
  args.relu_mask = a->relu_mask;
  const int cluster = trunk_cluster_size();
  UPNERF_TRY(make_tmap_bf16_2d(&maps.w, a->wcat_t, 256, UPNERF_TRUNK_WCATT_COLS, a->ld_w, 128 / cluster, 64));
  UPNERF_TRY(make_tmap_bf16_2d(&maps.in, a->d_hf, a->M, 256, a->ld_dhf, kTileM, 64));
  for (int j = 0; j < kNLb; ++j) {
    UPNERF_REQUIRE(a->d_out[j], UPNERF_ERR_BAD_SHAPE, "mlp_trunk_bwd: d_out[%d] missing", j);
    UPNERF_REQUIRE((reinterpret_cast<uintptr_t>(a->d_out[j]) & 15) == 0 && (a->ld_dout[j] & 7) == 0,
                   UPNERF_ERR_BAD_SHAPE, "mlp_trunk_bwd: d_out[%d] must be 16-byte aligned with ld %% 8 == 0", j);
    args.out[j] = static_cast<__nv_bfloat16*>(a->d_out[j]);
    args.ld_out[j] = a->ld_dout[j];
  }
  static bool attr_set = false;
  if (!attr_set) {
    UPNERF_CHECK_CUDA(cudaFuncSetAttribute(pp::mlp_trunk_bwd_pp_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                           pp::kSmemBytesBP));
    UPNERF_CHECK_CUDA(cudaFuncSetAttribute(pp::mlp_trunk_bwd_pp_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                           pp::kSmemBytesBP));
    attr_set = true;
  }
  const double flop = 2.0 * a->M * 256.0 * 256.0 * kNLb;
  // algorithmic traffic: read dHF + bit masks + d_ssig, write eight gradient tensors
  const double bytes = 2.0 * a->M * (256 + 8 * 256) + 8.0 * 32 * a->M + 4.0 * a->M;
  LaunchScope scope(kCatTrunkBwd, as_stream(stream), flop, bytes);
  UPNERF_TRY(launch_trunk(pp::mlp_trunk_bwd_pp_kernel<1>, pp::mlp_trunk_bwd_pp_kernel<2>, cluster, tiles,
                          pp::kSmemBytesBP, as_stream(stream), maps, args));
  UPNERF_CHECK_LAUNCH("mlp_trunk_bwd_pp_kernel");
  return UPNERF_OK;
}
