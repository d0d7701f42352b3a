// Pipelined, warp-specialised fused attention for sm_100a (tcgen05.mma + tensor memory + TMA), head_dim 64, S <= 384, with or without
// probability dropout: the ViLT stack's masked-softmax attention (HF:models/vilt/modeling_vilt.py:306-365) forward and backward at every shape the
// reference pipeline produces (T <= 128 text tokens + CLS + <= 240 patches = 369 keys).
//
// One skeleton, three instantiations.  A CTA is persistent over (sample, head) items; per item two [S x 64] operands stay resident
// in shared memory (128-row blocks, each with its own full / empty mbarrier so the next item's blocks stream in as soon as the last
// MMA that reads the old one has retired) and the item is cut in 128-row tiles:
//
//   MODE_F   tile = 128 queries.  S_b = Q_t K_b^T per 128-key block -> softmax -> P_b (bf16, smem) -> O_t += P_b V_b.
//            Exact two-pass softmax without an online rescale: the score blocks of a tile are streamed twice, first for the row
//            maximum, then (last block kept, the others recomputed -- the tensor pipe is the idle unit here, MUFU the busy one) for
//            exp2 / row sum / P.  Block stream of a tile: max(b0) .. max(b_{n-2}), max+exp(b_{n-1}), exp(b_{n-2}) .. exp(b0).
//   MODE_BQ  tile = 128 queries.  S_b, dP_b = dO_t V_b^T -> dS_b = P_b o (dP_b - delta) -> dQ_t += dS_b K_b.
//   MODE_BKV tile = 128 keys, blocks = 128 queries.  S_b = Q_b K_t^T, dP_b = dO_b V_t^T (lanes = queries) -> P_b, dS_b ->
//            dV_t += P_b^T dO_b, dK_t += dS_b^T Q_b (the P / dS buffers are read as MN-major A operands).
//
// Roles (384 threads, 1 CTA / SM): warp 0 = TMA producer, warp 1 = tcgen05.mma issuer (one thread, runs the score MMAs up to two
// blocks ahead of the output MMAs), warp 2 = TMEM allocator, warps 4-11 = two softmax warpgroups that share every tile: thread =
// row = TMEM lane, warpgroup g owns columns [64g, 64g+64) of each 128-column score block (row max / row sum / delta-free exchange
// through shared memory once per tile).  Score blocks live in a TMEM ring (3 x 128 columns forward, 2 x 128 + 128 for dP
// backward), accumulators in the remaining 128 columns (double-buffered per tile in F / BQ so a tile's epilogue runs under the next
// tile's first block).  All synchronisation between roles is mbarrier-based; the only named barrier is the 256-thread exchange.
//
// Lanes are always queries, columns keys: lse / delta are per-thread scalars and key validity is a per-column bit mask.  The LSE
// convention (natural log of the sum of exp of the scaled, masked scores) is the one of attention.cu / attention_tc.cu.
#include "common.cuh"

namespace vb {

int encode_tmap_2d(CUtensorMap* tm, CUtensorMapDataType dt, int elem_bytes, const void* base, uint64_t inner, uint64_t outer, uint64_t ld_elems,
                   uint32_t box_inner, uint32_t box_outer);

namespace {

constexpr int kThreadsA = 384;
constexpr int MODE_F = 0, MODE_BQ = 1, MODE_BKV = 2;
constexpr float kLog2eA = 1.4426950408889634f;
constexpr float kLn2A = 0.6931471805599453f;

struct AParams {
  const uint8_t* key_mask;  // [B,S]
  bf16* ctx;                // F: out [B*S, H]
  float* lse;               // F: out / B: in [B,heads,S]
  const float* delta;       // B: in [B,heads,S]
  bf16* dqkv;               // B: out [B*S, 3H]
  int B, S, heads, nblk, n_items;
  float scale_log2, scale;
  float dropout_p;                       // probability dropout (HF:models/bert/modeling_bert.py eager attention: dropout(softmax(...)) before @V)
  unsigned long long seed;
  const unsigned long long* seed_dev;    // device counter added to the seed (advanced once per training step), may be NULL
  unsigned site;
  long long* trace;  // VAULT_B200_ATTN_TRACE=1 (debug): (clock64, event id) pairs of CTA 0's MMA thread [0..1023] and first softmax warp [1024..2047]
};

#define VB_TR(who, id)                                                                                     \
  do {                                                                                                     \
    if (p.trace != nullptr && blockIdx.x == 0 && tr_n < 1023) {                                            \
      p.trace[((who) * 1024 + tr_n) * 2] = clock64();                                                      \
      p.trace[((who) * 1024 + tr_n) * 2 + 1] = (id);                                                       \
      ++tr_n;                                                                                              \
    }                                                                                                      \
  } while (0)

// shared memory map (bytes from the 1024-aligned base)
constexpr uint32_t kBlk = 16384;                 // one [128 x 64] bf16 operand block (two 64-row TMA boxes)
constexpr uint32_t oR1 = 0, oR2 = 3 * kBlk;      // resident operands, up to 3 blocks each
constexpr uint32_t oT = 6 * kBlk;                // per-tile operands: F: Q_t x 2 (double buffer); B: T1, T2
constexpr uint32_t oOP = 8 * kBlk;               // P / dS staging: 4 x 16 KB
constexpr uint32_t oStage = 12 * kBlk;           // epilogue staging: 2 KB per softmax warp (32 rows x 32 bf16), coalesced global stores
constexpr uint32_t oMisc = 13 * kBlk;            // barriers, exchange arrays
constexpr uint32_t kSmemA = 13 * kBlk + 6144 + 1024;  // misc: 512 B barriers + 2 KB xm + 2 KB xl + mask words

// barrier indices
constexpr int bR1F = 0, bR1E = 3, bR2F = 6, bR2E = 9, bTF = 12, bTE = 14, bSF = 16, bSE = 19, bDPF = 22, bDPE = 23, bOPF = 24, bOPE = 28, bOF = 32,
              bOE = 34, kNumBars = 36;

__device__ __forceinline__ void fence_async_smem_a() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void st_shared_v4a(uint32_t addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}
// 32 consecutive bf16 (64 bytes) of row `row`, element columns [32c, 32c+32) of a [128 x 64] 128B-swizzled chunk
__device__ __forceinline__ void store_row32a(uint32_t buf, int row, int c, const uint32_t (&pk)[16]) {
  const uint32_t rowaddr = buf + (uint32_t)row * 128u;
  const uint32_t seg0 = (uint32_t)c * 4u;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const uint32_t seg = (seg0 + i) ^ ((uint32_t)row & 7u);
    st_shared_v4a(rowaddr + (seg << 4), pk[4 * i], pk[4 * i + 1], pk[4 * i + 2], pk[4 * i + 3]);
  }
}
__device__ __forceinline__ uint4 ld_shared_v4a(uint32_t addr) {
  uint4 v;
  asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr));
  return v;
}
// A warp owns 32 accumulator rows (lane = row) x 32 fp32 columns in registers and writes them, scaled and rounded to bf16, to 32 rows
// (64 bytes each) of a row-major global matrix.  A lane storing its own row costs 32 L1 wavefronts per instruction, so the rows go
// through the warp's 2 KB of shared memory (16-byte segments XOR-swizzled) and leave 8 rows per instruction.
__device__ __forceinline__ void store_rows32_staged(const uint32_t (&r)[32], float mul, uint32_t stage, bf16* gptr0, long long row_stride, int n_valid,
                                                    int lane) {
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const uint32_t seg = (uint32_t)i ^ (((uint32_t)lane >> 1) & 3u);
    st_shared_v4a(stage + (uint32_t)lane * 64u + (seg << 4),
                  pack_bf16x2(__uint_as_float(r[8 * i]) * mul, __uint_as_float(r[8 * i + 1]) * mul),
                  pack_bf16x2(__uint_as_float(r[8 * i + 2]) * mul, __uint_as_float(r[8 * i + 3]) * mul),
                  pack_bf16x2(__uint_as_float(r[8 * i + 4]) * mul, __uint_as_float(r[8 * i + 5]) * mul),
                  pack_bf16x2(__uint_as_float(r[8 * i + 6]) * mul, __uint_as_float(r[8 * i + 7]) * mul));
  }
  __syncwarp();
#pragma unroll
  for (int it = 0; it < 4; ++it) {
    const int rr = 8 * it + (lane >> 2), sg = lane & 3;
    const uint4 v = ld_shared_v4a(stage + (uint32_t)rr * 64u + (uint32_t)((sg ^ ((rr >> 1) & 3)) << 4));
    if (rr < n_valid) *reinterpret_cast<uint4*>(gptr0 + (long long)rr * row_stride + sg * 8) = v;
  }
  __syncwarp();
}
__device__ __forceinline__ void bar_sync_256() { asm volatile("bar.sync 1, 256;" ::: "memory"); }

// Probability dropout (DROP): one Philox4x32 call yields eight 16-bit draws = the keep decisions of 8 consecutive keys of one query row,
// counter = ((sample * heads + head) * S + query) * 64 + key / 8, stream = dropout site.  Lanes are queries and a thread walks keys in
// aligned groups of 32 in all three kernels, so forward, dQ and dK/dV regenerate the same mask without exchanging a bit.
//   forward:  O = (P o keep / (1 - p)) V, row sums / LSE over the undropped P
//   backward: dP = (dO V^T) o keep / (1 - p);  dS = P o (dP - delta);  dV = (P o keep / (1 - p))^T dO;  delta = sum_d dO o O
template <int MODE, bool DROP>
__global__ void __launch_bounds__(kThreadsA, 1)
attn_sm100_kernel(const __grid_constant__ CUtensorMap tmQKV, const __grid_constant__ CUtensorMap tmDO, const AParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  uint8_t* gen = smem_raw + (base - raw);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int S = p.S, H = p.heads * 64, nblk = p.nblk;
  const int ntiles = nblk;                             // 128-row tiles per item (queries in F / BQ, keys in BKV)
  const int L = MODE == MODE_F ? 2 * nblk - 1 : nblk;  // score blocks per tile
  // TMEM map.  F: score ring 3 x 128 | O 2 x 64.  BQ / BKV: S 128 | dP 128 | accumulators 2 x 128 (BQ uses 64 of each): in the backward
  // a softmax warp has its piece of S and dP in registers half-way through the block, so a single buffer each already lets the
  // next block's score MMAs run under the second half, and the spare columns double-buffer the accumulators (epilogues drain
  // under the next tile).
  constexpr int kRing = MODE == MODE_F ? 3 : 1;
  constexpr uint32_t colDP = 128, colAcc = MODE == MODE_F ? 384u : 256u, accStride = MODE == MODE_F ? 64u : 128u;

  const uint32_t bars = base + oMisc;
  auto bar = [&](int i) { return bars + 8u * (uint32_t)i; };
  const uint32_t tmem_slot = bars + 8u * kNumBars;
  volatile uint32_t* tmem_slot_gen = reinterpret_cast<volatile uint32_t*>(gen + oMisc + 8 * kNumBars);
  float* xm = reinterpret_cast<float*>(gen + oMisc + 512);           // [2 parity][2 wg][128] row-max exchange
  float* xl = xm + 512;                                              // [2 parity][2 wg][128] row-sum exchange
  uint32_t* mws = reinterpret_cast<uint32_t*>(xl + 512);             // [2 parity][12] key-validity words of the item

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmQKV);
    if (MODE != MODE_F) tma_prefetch_desc(&tmDO);
  }
  if (warp == 1 && lane == 0) {
    // resident-block "empty" barriers: one arrival per issuer thread that reads the block (score MMAs and / or output MMAs)
    const uint32_t r1e = MODE == MODE_F ? 1u : 2u;                       // F: K_b scores only; BQ: K_b scores + dQ; BKV: Q_b scores + dK
    const uint32_t r2e = MODE == MODE_BKV ? 2u : 1u;                     // F: V_b output only; BQ: V_b dP only; BKV: dO_b dP + dV
    for (int i = 0; i < 3; ++i) {
      mbar_init(bar(bR1F + i), 1); mbar_init(bar(bR1E + i), r1e); mbar_init(bar(bR2F + i), 1); mbar_init(bar(bR2E + i), r2e);
      mbar_init(bar(bSF + i), 1); mbar_init(bar(bSE + i), 8);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(bar(bTF + i), 1); mbar_init(bar(bTE + i), 1); mbar_init(bar(bOF + i), 1); mbar_init(bar(bOE + i), 8);
    }
    mbar_init(bar(bDPF), 1); mbar_init(bar(bDPE), 8);
    for (int i = 0; i < 4; ++i) {
      mbar_init(bar(bOPF + i), 8);  // P / dS of one score block: all eight softmax warps
      mbar_init(bar(bOPE + i), 1);
    }
    mbar_fence_init();
  }
  if (warp == 2) tmem_alloc(tmem_slot, 512u);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot_gen;
  pdl_enter();

  // block index / kind of stream position j of tile nt (F: kind 0 = max only, 1 = max + exp, 2 = exp; B modes: always 2).  In F the
  // direction of the walk alternates from tile to tile (b0 .. b_last .. b0, then b_last .. b0 .. b_last): a tile ends on the block the
  // next one does NOT start with, so across an item boundary the blocks the next item needs first were released first.
  auto blk_of = [&](int j, int nt) {
    if (MODE != MODE_F) return j;
    const int i = j < nblk ? j : 2 * nblk - 2 - j;
    return (nt & 1) ? nblk - 1 - i : i;
  };
  auto kind_of = [&](int j) { return MODE == MODE_F ? (j < nblk - 1 ? 0 : (j == nblk - 1 ? 1 : 2)) : 2; };
  // resident blocks go through a 3-slot ring indexed by the global block number g = it * nblk + b (identity for nblk = 3; for shorter
  // sequences the next item's blocks land in free slots while the current item is still being worked on)
  auto rslot = [&](int it, int b) { return (it * nblk + b) % 3; };
  auto rpar = [&](int it, int b) { return (uint32_t)(((it * nblk + b) / 3) & 1); };
  // Work of this CTA: a contiguous range [t0, t1) of the launch's n_items x ntiles tiles, balanced to one tile (S = 369: 1,152 tiles over 148
  // CTAs = 7 or 8 each, where whole items gave 6 or 9).  An item may be shared by two CTAs; each then loads the item's resident blocks itself.
  const long long Tt = (long long)p.n_items * ntiles;
  const int t0 = (int)(Tt * blockIdx.x / gridDim.x), t1 = (int)(Tt * (blockIdx.x + 1) / gridDim.x);
  const int item0 = t0 / ntiles;
  const int my_items = t1 > t0 ? (t1 - 1) / ntiles - item0 + 1 : 0;
  auto tile_lo = [&](int it) { return it == 0 ? t0 - item0 * ntiles : 0; };
  auto tile_hi = [&](int it) { return it == my_items - 1 ? t1 - (item0 + it) * ntiles : ntiles; };

  if (warp == 0) {
    // =========================================== TMA producer ===========================================
    if (lane == 0) {
      int nt = 0;
      for (int it = 0; it < my_items; ++it) {
        const int item = item0 + it;
        const int b_ = item / p.heads, h_ = item % p.heads;
        const int row0 = b_ * S;
        for (int tile = tile_lo(it), tile_end = tile_hi(it); tile < tile_end; ++tile, ++nt) {
          const int ts = MODE == MODE_F ? (nt & 1) : 0;
          const uint32_t tpar = (uint32_t)((MODE == MODE_F ? (nt >> 1) : nt) & 1);
          mbar_wait(bar(bTE + ts), tpar ^ 1u);
          const uint32_t fb = bar(bTF + ts);
          const int r = row0 + tile * 128;
          if (MODE == MODE_F) {
            mbar_expect_tx(fb, kBlk);
            const uint32_t d = base + oT + (uint32_t)ts * kBlk;
            tma_load_2d(d, &tmQKV, fb, h_ * 64, r);
            tma_load_2d(d + 8192u, &tmQKV, fb, h_ * 64, r + 64);
          } else {
            mbar_expect_tx(fb, 2 * kBlk);
            const uint32_t d = base + oT;
            const int c1 = MODE == MODE_BQ ? h_ * 64 : H + h_ * 64;  // BQ: Q_t ; BKV: K_t
            tma_load_2d(d, &tmQKV, fb, c1, r);
            tma_load_2d(d + 8192u, &tmQKV, fb, c1, r + 64);
            if (MODE == MODE_BQ) {  // dO_t
              tma_load_2d(d + kBlk, &tmDO, fb, h_ * 64, r);
              tma_load_2d(d + kBlk + 8192u, &tmDO, fb, h_ * 64, r + 64);
            } else {  // V_t
              tma_load_2d(d + kBlk, &tmQKV, fb, 2 * H + h_ * 64, r);
              tma_load_2d(d + kBlk + 8192u, &tmQKV, fb, 2 * H + h_ * 64, r + 64);
            }
          }
          if (tile == tile_lo(it)) {
            // resident blocks in the order the item's first tile needs them (= the order the previous item released them)
            const int dir0 = MODE == MODE_F ? (nt & 1) : 0;
            auto load_r = [&](int which, int bb) {  // which 0: R1 (F / BQ: K_b, BKV: Q_b); 1: R2 (F / BQ: V_b, BKV: dO_b)
              const int sl = rslot(it, bb);
              mbar_wait(bar((which ? bR2E : bR1E) + sl), rpar(it, bb) ^ 1u);
              const uint32_t f = bar((which ? bR2F : bR1F) + sl);
              mbar_expect_tx(f, kBlk);
              const uint32_t d = base + (which ? oR2 : oR1) + (uint32_t)sl * kBlk;
              const int rr = row0 + bb * 128;
              if (MODE == MODE_BKV && which == 1) {
                tma_load_2d(d, &tmDO, f, h_ * 64, rr);
                tma_load_2d(d + 8192u, &tmDO, f, h_ * 64, rr + 64);
              } else {
                const int c = MODE == MODE_BKV ? h_ * 64 : (which ? 2 * H + h_ * 64 : H + h_ * 64);
                tma_load_2d(d, &tmQKV, f, c, rr);
                tma_load_2d(d + 8192u, &tmQKV, f, c, rr + 64);
              }
            };
            if (MODE == MODE_F) {
              for (int i = 0; i < nblk; ++i) load_r(0, dir0 ? nblk - 1 - i : i);  // K: walk order of the max pass
              for (int i = 0; i < nblk; ++i) load_r(1, dir0 ? i : nblk - 1 - i);  // V: first needed by the max+exp block, then back
            } else {
              for (int i = 0; i < nblk; ++i) { load_r(0, i); load_r(1, i); }
            }
          }
        }
      }
    }
  } else if (warp == 1) {
    // =========================================== score-MMA issuer (one thread) ===========================================
    // Runs as far ahead of the softmax warps as the TMEM score ring allows; every wait is an in-order blocking one.
    if (lane == 0) {
      const uint32_t id_s = umma_idesc(1u, 128u, 128u, 0u, 0u);  // both operands K-major (dh contiguous)
      int n = 0, nt = 0, tr_n = 0;
      for (int it = 0; it < my_items; ++it) {
        uint32_t seen = 0;  // bit b: R1_b of this item already waited for, bit 4 + b: R2_b
        for (int tile = tile_lo(it), tile_end = tile_hi(it); tile < tile_end; ++tile, ++nt) {
          const int ts = MODE == MODE_F ? (nt & 1) : 0;
          const bool last_tile = tile == tile_end - 1;
          for (int j = 0; j < L; ++j, ++n) {
            const int b = blk_of(j, nt), kind = kind_of(j);
            const int sl = n % kRing, rs = rslot(it, b);
            VB_TR(0, 100 + n);
            if (j == 0) mbar_wait(bar(bTF + ts), (uint32_t)((MODE == MODE_F ? (nt >> 1) : nt) & 1));
            if (!(seen & (1u << b))) { mbar_wait(bar(bR1F + rs), rpar(it, b)); seen |= 1u << b; }
            mbar_wait(bar(bSE + sl), (uint32_t)((n / kRing) & 1) ^ 1u);
            tc_fence_after();
            const uint32_t sT1 = base + oT + (MODE == MODE_F ? (uint32_t)ts * kBlk : 0u), sT2 = base + oT + kBlk;
            const uint32_t sR1b = base + oR1 + (uint32_t)rs * kBlk, sR2b = base + oR2 + (uint32_t)rs * kBlk;
            const uint32_t aS = MODE == MODE_BKV ? sR1b : sT1, bS = MODE == MODE_BKV ? sT1 : sR1b;
#pragma unroll
            for (int k = 0; k < 4; ++k)
              tc_mma_f16(tmem + (uint32_t)(sl * 128), umma_desc_sw128(aS + k * 32u, 16u, 1024u), umma_desc_sw128(bS + k * 32u, 16u, 1024u), id_s,
                         k > 0 ? 1u : 0u);
            tc_commit(bar(bSF + sl));
            if (MODE != MODE_F) {
              if (!(seen & (16u << b))) { mbar_wait(bar(bR2F + rs), rpar(it, b)); seen |= 16u << b; }
              mbar_wait(bar(bDPE), (uint32_t)(n & 1) ^ 1u);
              tc_fence_after();
              const uint32_t aD = MODE == MODE_BKV ? sR2b : sT2, bD = MODE == MODE_BKV ? sT2 : sR2b;
#pragma unroll
              for (int k = 0; k < 4; ++k)
                tc_mma_f16(tmem + colDP, umma_desc_sw128(aD + k * 32u, 16u, 1024u), umma_desc_sw128(bD + k * 32u, 16u, 1024u), id_s, k > 0 ? 1u : 0u);
              tc_commit(bar(bDPF));
            }
            if (j == L - 1) tc_commit(bar(bTE + ts));  // the tile's operands are dead
            if (last_tile) {
              // last score-side use of the resident blocks in this item (F: K_b at its exp-kind visit; B modes: every block once)
              if (MODE == MODE_F) { if (kind != 0) tc_commit(bar(bR1E + rs)); }
              else { tc_commit(bar(bR1E + rs)); if (MODE == MODE_BQ || MODE == MODE_BKV) tc_commit(bar(bR2E + rs)); }
            }
            VB_TR(0, 200 + n);
          }
        }
      }
    }
  } else if (warp == 3) {
    // =========================================== output-MMA issuer (one thread) ===========================================
    if (lane == 0) {
      const uint32_t id_o = umma_idesc(1u, 128u, 64u, 0u, 1u);  // F / BQ: A = P / dS chunk K-major, B = V / K rows MN-major
      const uint32_t id_t = umma_idesc(1u, 128u, 64u, 1u, 1u);  // BKV: A = P / dS read MN-major (M = keys), B = dO / Q rows MN-major
      int opn = 0, nt = 0, tr_n = 1 << 20;
      (void)tr_n;
      for (int it = 0; it < my_items; ++it) {
        uint32_t seen = 0;
        for (int tile = tile_lo(it), tile_end = tile_hi(it); tile < tile_end; ++tile, ++nt) {
          const bool last_tile = tile == tile_end - 1;
          const int ob = nt & 1;
          for (int j = 0; j < L; ++j) {
            const int b = blk_of(j, nt), kind = kind_of(j);
            if (kind == 0) continue;
            const bool first = MODE == MODE_F ? kind == 1 : j == 0;
            const int slot = opn & 1, rs = rslot(it, b);
            if (first) mbar_wait(bar(bOE + ob), (uint32_t)((nt >> 1) & 1) ^ 1u);
            if (!(seen & (1u << b))) {  // the MN-major operand rows of this block (loaded by TMA: observe its barrier in this thread too)
              if (MODE != MODE_F) mbar_wait(bar(bR1F + rs), rpar(it, b));
              if (MODE != MODE_BQ) mbar_wait(bar(bR2F + rs), rpar(it, b));
              seen |= 1u << b;
            }
            if (MODE != MODE_BKV) {
              mbar_wait(bar(bOPF + slot), (uint32_t)((opn >> 1) & 1));
              tc_fence_after();
              const uint32_t acc = tmem + colAcc + (uint32_t)ob * accStride;
              const uint32_t sRb = base + (MODE == MODE_F ? oR2 : oR1) + (uint32_t)rs * kBlk;
#pragma unroll
              for (int g = 0; g < 2; ++g) {
                const uint32_t sP = base + oOP + (uint32_t)(g * 2 + slot) * kBlk;
#pragma unroll
                for (int ks = 0; ks < 4; ++ks)
                  tc_mma_f16(acc, umma_desc_sw128(sP + ks * 32u, 16u, 1024u), umma_desc_sw128(sRb + (uint32_t)g * 8192u + ks * 2048u, 8192u, 1024u), id_o,
                             (first && g == 0 && ks == 0) ? 0u : 1u);
              }
              tc_commit(bar(bOPE + slot));
              if (last_tile) tc_commit(bar((MODE == MODE_F ? bR2E : bR1E) + rs));  // F: V_b, BQ: K_b -- last output-side use in this item
            } else {
              // dV_t += P_b^T dO_b ; dK_t += dS_b^T Q_b  (A read MN-major: M = keys, two 64-key chunks 16 KB apart; K = the block's 128 queries)
              mbar_wait(bar(bOPF + 0), (uint32_t)(opn & 1));
              tc_fence_after();
              const uint32_t sP = base + oOP, sDS = base + oOP + 2 * kBlk;
              const uint32_t sQb = base + oR1 + (uint32_t)rs * kBlk, sDOb = base + oR2 + (uint32_t)rs * kBlk;
#pragma unroll
              for (int ks = 0; ks < 8; ++ks)
                tc_mma_f16(tmem + colAcc + (uint32_t)ob * accStride, umma_desc_sw128(sP + ks * 2048u, 16384u, 1024u), umma_desc_sw128(sDOb + ks * 2048u, 8192u, 1024u), id_t,
                           (first && ks == 0) ? 0u : 1u);
#pragma unroll
              for (int ks = 0; ks < 8; ++ks)
                tc_mma_f16(tmem + colAcc + (uint32_t)ob * accStride + 64u, umma_desc_sw128(sDS + ks * 2048u, 16384u, 1024u), umma_desc_sw128(sQb + ks * 2048u, 8192u, 1024u), id_t,
                           (first && ks == 0) ? 0u : 1u);
              tc_commit(bar(bOPE + 0));
              if (last_tile) { tc_commit(bar(bR1E + rs)); tc_commit(bar(bR2E + rs)); }
            }
            if (j == L - 1) tc_commit(bar(bOF + ob));
            ++opn;
          }
        }
      }
    }
  } else if (warp >= 4) {
    // =========================================== softmax / epilogue warps ===========================================
    const int ew = warp - 4, g = ew >> 2, q = warp & 3;
    const int row = q * 32 + lane;
    const uint32_t lanef = (uint32_t)(q * 32) << 16;
    int opn = 0;
    int tr_n = (warp == 4 && lane == 0) ? 0 : 1 << 20;
    const unsigned long long dseed = DROP ? p.seed + (p.seed_dev != nullptr ? *p.seed_dev : 0ull) : 0ull;
    const uint32_t thr16 = DROP ? (uint32_t)fminf(p.dropout_p * 65536.0f, 65535.0f) : 0u;
    const float inv_keep = DROP ? 1.0f / (1.0f - p.dropout_p) : 1.0f;
    // keep decisions of elements 2i, 2i+1 (i = 0..3) of an 8-key group from one Philox call
    auto keep2 = [&](const uint4& rb, int i, bool& k0, bool& k1) {
      const uint32_t w = i == 0 ? rb.x : (i == 1 ? rb.y : (i == 2 ? rb.z : rb.w));
      k0 = (w & 0xffffu) >= thr16;
      k1 = (w >> 16) >= thr16;
    };
    // deferred epilogue state (F / BQ): the accumulator of tile `pend_nt` is drained while the next tile is in flight
    bool pend = false;
    int pend_nt = 0, pend_item = 0, pend_tile = 0;
    float pend_msc = 0.f;

    // Drain the accumulator of tile `nt` (F: the caller has passed a 256-thread barrier since both warpgroups wrote the tile's partial
    // row sums).  F / BQ: warpgroup g owns output columns [32g, 32g+32); BKV: warpgroup 0 = dV_t, warpgroup 1 = dK_t (x softmax scale).
    const uint32_t stage = base + oStage + (uint32_t)ew * 2048u;
    auto epilogue = [&](int nt, int item, int tile, float msc) {
      const int b_ = item / p.heads, h_ = item % p.heads;
      const int ob = nt & 1;
      VB_TR(1, 8000 + nt);
      mbar_wait(bar(bOF + ob), (uint32_t)((nt >> 1) & 1));
      tc_fence_after();
      const int r0g = tile * 128 + q * 32;  // first row (query, or key in BKV) of this warp inside the sample
      if (MODE != MODE_BKV) {
        uint32_t r[32];
        tmem_ld32(tmem + lanef + colAcc + (uint32_t)ob * accStride + (uint32_t)(g * 32), r);
        float mul = p.scale;
        float l_tot = 0.f;
        if (MODE == MODE_F) {
          l_tot = xl[(nt & 1) * 256 + row] + xl[(nt & 1) * 256 + 128 + row];
          mul = l_tot > 0.f ? 1.f / l_tot : 0.f;
        }
        tmem_ld_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(bar(bOE + ob));
        bf16* g0 = MODE == MODE_F ? p.ctx + ((long long)b_ * S + r0g) * H + h_ * 64 + g * 32 : p.dqkv + ((long long)b_ * S + r0g) * 3LL * H + h_ * 64 + g * 32;
        store_rows32_staged(r, mul, stage, g0, MODE == MODE_F ? (long long)H : 3LL * H, S - r0g, lane);
        if (MODE == MODE_F && g == 0 && r0g + lane < S && p.lse != nullptr)
          p.lse[((long long)b_ * p.heads + h_) * S + r0g + lane] = msc * kLn2A + __logf(fmaxf(l_tot, 1e-30f));
      } else {
        uint32_t ra[32], rb[32];
        tmem_ld32(tmem + lanef + colAcc + (uint32_t)ob * accStride + (uint32_t)(g * 64), ra);
        tmem_ld32(tmem + lanef + colAcc + (uint32_t)ob * accStride + (uint32_t)(g * 64 + 32), rb);
        tmem_ld_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(bar(bOE + ob));
        bf16* g0 = p.dqkv + ((long long)b_ * S + r0g) * 3LL * H + (g == 0 ? 2 * H : H) + h_ * 64;
        const float mul = g == 0 ? 1.f : p.scale;
        store_rows32_staged(ra, mul, stage, g0, 3LL * H, S - r0g, lane);
        store_rows32_staged(rb, mul, stage, g0 + 32, 3LL * H, S - r0g, lane);
      }
      VB_TR(1, 8500 + nt);
    };

    // key-validity bytes of the NEXT item are fetched one item ahead (two keys per lane per warp: words ew and ew + 8)
    auto mask_bytes = [&](int item, int w) -> uint32_t {
      if (item >= p.n_items || w >= nblk * 4) return 0u;
      const int key = 32 * w + lane;
      return (key < S && p.key_mask[(long long)(item / p.heads) * S + key] != 0) ? 1u : 0u;
    };
    const int first_item = my_items > 0 ? item0 : p.n_items;
    uint32_t nx0 = mask_bytes(first_item, ew), nx1 = mask_bytes(first_item, ew + 8);

    int n = 0, nt = 0;
    for (int it = 0; it < my_items; ++it) {
      const int item = item0 + it;
      const int b_ = item / p.heads, h_ = item % p.heads;
      const long long bh = (long long)b_ * p.heads + h_;
      {
        const uint32_t w0 = __ballot_sync(0xffffffffu, nx0 != 0u), w1 = __ballot_sync(0xffffffffu, nx1 != 0u);
        if (lane == 0) {
          mws[(it & 1) * 12 + ew] = w0;
          if (ew + 8 < 12) mws[(it & 1) * 12 + ew + 8] = w1;
        }
        const int nitem = it + 1 < my_items ? item + 1 : p.n_items;
        nx0 = mask_bytes(nitem, ew);
        nx1 = mask_bytes(nitem, ew + 8);
        bar_sync_256();
      }
      const uint32_t* mw_item = mws + (it & 1) * 12;

      for (int tile = tile_lo(it), tile_end = tile_hi(it); tile < tile_end; ++tile, ++nt) {
        const bool warp_active = MODE == MODE_BKV ? true : (tile * 128 + q * 32 < S);  // warp-uniform
        float m_run = -INFINITY, msc = 0.f, l = 0.f;
        float lse2 = INFINITY, dl = 0.f;
        if (MODE == MODE_BQ) {
          const int qg = tile * 128 + row;
          if (qg < S) { lse2 = p.lse[bh * S + qg] * kLog2eA; dl = p.delta[bh * S + qg]; }
        }
        float lse2_n = INFINITY, dl_n = 0.f;
        if (MODE == MODE_BKV) {
          if (row < S) { lse2_n = p.lse[bh * S + row] * kLog2eA; dl_n = p.delta[bh * S + row]; }
        }
        for (int j = 0; j < L; ++j, ++n) {
          const int b = blk_of(j, nt), kind = kind_of(j);
          const int sl = n % kRing;
          // validity of my 64 columns: keys of block b (F / BQ) or of the tile (BKV)
          const int kb = MODE == MODE_BKV ? tile : b;
          const uint32_t mw0 = mw_item[kb * 4 + g * 2], mw1 = mw_item[kb * 4 + g * 2 + 1];
          if (MODE == MODE_BKV) {
            lse2 = lse2_n; dl = dl_n;
            lse2_n = INFINITY; dl_n = 0.f;
            const int qn = (b + 1) * 128 + row;
            if (b + 1 < nblk && qn < S) { lse2_n = p.lse[bh * S + qn] * kLog2eA; dl_n = p.delta[bh * S + qn]; }
          }
          const int slot = opn & 1;
          VB_TR(1, 1000 + n);
          mbar_wait(bar(bSF + sl), (uint32_t)((n / kRing) & 1));
          tc_fence_after();
          VB_TR(1, 2000 + n);
          const uint32_t tS = tmem + lanef + (uint32_t)(sl * 128 + g * 64);
          if (MODE == MODE_F) {
            // ---- my 32 x 64 piece of the block goes to registers in one go; the TMEM buffer is free again right away ----
            uint32_t r0[32], r1[32];
            tmem_ld32(tS, r0);
            tmem_ld32(tS + 32u, r1);
            const uint32_t sP = base + oOP + (uint32_t)(g * 2 + slot) * kBlk;
            if (kind != 0) mbar_wait(bar(bOPE + slot), (uint32_t)((opn >> 1) & 1) ^ 1u);  // under the load latency
            tmem_ld_wait();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(bar(bSE + sl));
            if (kind != 2 && warp_active) {
              // ---- row maximum over my 64 columns ----
              auto rmax = [&](const uint32_t (&r)[32], uint32_t mw) {
                if (mw == 0xffffffffu) {
                  float m0 = __uint_as_float(r[0]), m1 = __uint_as_float(r[1]), m2 = __uint_as_float(r[2]), m3 = __uint_as_float(r[3]);
#pragma unroll
                  for (int i = 4; i < 32; i += 4) {
                    m0 = fmaxf(m0, __uint_as_float(r[i])); m1 = fmaxf(m1, __uint_as_float(r[i + 1]));
                    m2 = fmaxf(m2, __uint_as_float(r[i + 2])); m3 = fmaxf(m3, __uint_as_float(r[i + 3]));
                  }
                  m_run = fmaxf(m_run, fmaxf(fmaxf(m0, m1), fmaxf(m2, m3)));
                } else if (mw != 0u) {
                  float m0 = -INFINITY, m1 = -INFINITY;
#pragma unroll
                  for (int i = 0; i < 32; i += 2) {
                    m0 = fmaxf(m0, ((mw >> i) & 1u) ? __uint_as_float(r[i]) : -INFINITY);
                    m1 = fmaxf(m1, ((mw >> (i + 1)) & 1u) ? __uint_as_float(r[i + 1]) : -INFINITY);
                  }
                  m_run = fmaxf(m_run, fmaxf(m0, m1));
                }
              };
              rmax(r0, mw0);
              rmax(r1, mw1);
            }
            if (kind == 0) continue;
            if (kind == 1) {
              // ---- the two warpgroups exchange their partial maxima (the one 256-thread barrier of the tile) ----
              VB_TR(1, 3000 + n);
              xm[(nt & 1) * 256 + g * 128 + row] = m_run;
              bar_sync_256();
              m_run = fmaxf(m_run, xm[(nt & 1) * 256 + (g ^ 1) * 128 + row]);
              msc = m_run == -INFINITY ? 0.f : m_run * p.scale_log2;
              VB_TR(1, 4000 + n);
            }
            // ---- exp pass: P for my 64 columns ----
            const unsigned long long rowctr = DROP ? ((unsigned long long)(bh * S + tile * 128 + row) << 6) + (unsigned)((b * 128 + g * 64) >> 3) : 0ull;
            auto pexp = [&](const uint32_t (&r)[32], uint32_t mw, int c) {
              uint32_t pk[16];
              if (DROP) {
                if (warp_active && mw != 0u) {
#pragma unroll
                  for (int j = 0; j < 4; ++j) {
                    const uint4 rb = Philox(dseed)(rowctr + (unsigned)(c * 4 + j), p.site);
#pragma unroll
                    for (int e = 0; e < 4; ++e) {
                      const int i = 4 * j + e;
                      const float p0 = ((mw >> (2 * i)) & 1u) ? fast_exp2(fmaf(__uint_as_float(r[2 * i]), p.scale_log2, -msc)) : 0.f;
                      const float p1 = ((mw >> (2 * i + 1)) & 1u) ? fast_exp2(fmaf(__uint_as_float(r[2 * i + 1]), p.scale_log2, -msc)) : 0.f;
                      l += p0 + p1;
                      bool k0, k1;
                      keep2(rb, e, k0, k1);
                      pk[i] = pack_bf16x2(k0 ? p0 * inv_keep : 0.f, k1 ? p1 * inv_keep : 0.f);
                    }
                  }
                } else {
#pragma unroll
                  for (int i = 0; i < 16; ++i) pk[i] = 0u;
                }
              } else if (warp_active && mw == 0xffffffffu) {
                float l0 = 0.f, l1 = 0.f;
#pragma unroll
                for (int i = 0; i < 16; ++i) {
                  const float p0 = fast_exp2(fmaf(__uint_as_float(r[2 * i]), p.scale_log2, -msc));
                  const float p1 = fast_exp2(fmaf(__uint_as_float(r[2 * i + 1]), p.scale_log2, -msc));
                  l0 += p0; l1 += p1;
                  pk[i] = pack_bf16x2(p0, p1);
                }
                l += l0 + l1;
              } else if (warp_active && mw != 0u) {
#pragma unroll
                for (int i = 0; i < 16; ++i) {
                  const float p0 = ((mw >> (2 * i)) & 1u) ? fast_exp2(fmaf(__uint_as_float(r[2 * i]), p.scale_log2, -msc)) : 0.f;
                  const float p1 = ((mw >> (2 * i + 1)) & 1u) ? fast_exp2(fmaf(__uint_as_float(r[2 * i + 1]), p.scale_log2, -msc)) : 0.f;
                  l += p0 + p1;
                  pk[i] = pack_bf16x2(p0, p1);
                }
              } else {
#pragma unroll
                for (int i = 0; i < 16; ++i) pk[i] = 0u;
              }
              store_row32a(sP, row, c, pk);
            };
            VB_TR(1, 5000 + n);
            pexp(r0, mw0, 0);
            pexp(r1, mw1, 1);
            VB_TR(1, 6000 + n);
            fence_async_smem_a();
            __syncwarp();
            if (lane == 0) mbar_arrive(bar(bOPF + slot));
            ++opn;
            if (j == L - 1) xl[(nt & 1) * 256 + g * 128 + row] = l;
            VB_TR(1, 7000 + n);
            if (kind == 1 && pend) {
              // previous tile's accumulator (its row sums were published before this tile's exchange barrier): drained while the
              // tensor pipe works on this tile
              epilogue(pend_nt, pend_item, pend_tile, pend_msc);
              pend = false;
            }
          } else {
            // ---- BQ / BKV: P and dS for my 64 columns ----
            mbar_wait(bar(bDPF), (uint32_t)(n & 1));
            tc_fence_after();
            uint32_t sP, sDS = 0;
            if (MODE == MODE_BKV) {
              sP = base + oOP + (uint32_t)g * kBlk;
              sDS = base + oOP + 2 * kBlk + (uint32_t)g * kBlk;
            } else {
              sP = base + oOP + (uint32_t)(g * 2 + slot) * kBlk;
            }
#pragma unroll
            for (int c = 0; c < 2; ++c) {
              uint32_t r[32], d[32];
              uint32_t pk[16];
              const uint32_t mw = c == 0 ? mw0 : mw1;
              tmem_ld32(tS + (uint32_t)(c * 32), r);
              tmem_ld32(tmem + lanef + colDP + (uint32_t)(g * 64 + c * 32), d);
              if (c == 0) {  // the staging buffers of this block must have been read by the output MMAs of the previous use
                if (MODE == MODE_BKV) mbar_wait(bar(bOPE + 0), (uint32_t)(opn & 1) ^ 1u);
                else mbar_wait(bar(bOPE + slot), (uint32_t)((opn >> 1) & 1) ^ 1u);
              }
              tmem_ld_wait();
              if (c == 1) {
                tc_fence_before();
                __syncwarp();
                if (lane == 0) { mbar_arrive(bar(bSE + sl)); mbar_arrive(bar(bDPE)); }
              }
              uint32_t pp[16];
              if (DROP && warp_active && mw != 0u) {
                const int qg = MODE == MODE_BKV ? b * 128 + row : tile * 128 + row;
                const unsigned long long rowctr = ((unsigned long long)(bh * S + qg) << 6) + (unsigned)((kb * 128 + g * 64 + c * 32) >> 3);
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                  const uint4 rb = Philox(dseed)(rowctr + (unsigned)j, p.site);
#pragma unroll
                  for (int e = 0; e < 4; ++e) {
                    const int i = 4 * j + e;
                    const float p0 = ((mw >> (2 * i)) & 1u) ? fast_exp2(fmaf(__uint_as_float(r[2 * i]), p.scale_log2, -lse2)) : 0.f;
                    const float p1 = ((mw >> (2 * i + 1)) & 1u) ? fast_exp2(fmaf(__uint_as_float(r[2 * i + 1]), p.scale_log2, -lse2)) : 0.f;
                    bool k0, k1;
                    keep2(rb, e, k0, k1);
                    const float d0 = k0 ? __uint_as_float(d[2 * i]) * inv_keep : 0.f, d1 = k1 ? __uint_as_float(d[2 * i + 1]) * inv_keep : 0.f;
                    pk[i] = pack_bf16x2(p0 * (d0 - dl), p1 * (d1 - dl));
                    if (MODE == MODE_BKV) pp[i] = pack_bf16x2(k0 ? p0 * inv_keep : 0.f, k1 ? p1 * inv_keep : 0.f);
                  }
                }
              } else if (warp_active && mw != 0u) {
#pragma unroll
                for (int i = 0; i < 16; ++i) {
                  float p0 = fast_exp2(fmaf(__uint_as_float(r[2 * i]), p.scale_log2, -lse2));
                  float p1 = fast_exp2(fmaf(__uint_as_float(r[2 * i + 1]), p.scale_log2, -lse2));
                  if (mw != 0xffffffffu) {
                    p0 = ((mw >> (2 * i)) & 1u) ? p0 : 0.f;
                    p1 = ((mw >> (2 * i + 1)) & 1u) ? p1 : 0.f;
                  }
                  pk[i] = pack_bf16x2(p0 * (__uint_as_float(d[2 * i]) - dl), p1 * (__uint_as_float(d[2 * i + 1]) - dl));
                  if (MODE == MODE_BKV) pp[i] = pack_bf16x2(p0, p1);
                }
              } else {
#pragma unroll
                for (int i = 0; i < 16; ++i) { pk[i] = 0u; pp[i] = 0u; }
              }
              if (MODE == MODE_BKV) {
                store_row32a(sP, row, c, pp);
                store_row32a(sDS, row, c, pk);
              } else {
                store_row32a(sP, row, c, pk);
              }
            }
            fence_async_smem_a();
            __syncwarp();
            if (lane == 0) mbar_arrive(bar(bOPF + (MODE == MODE_BKV ? 0 : slot)));
            ++opn;
            if (pend && j == 0) {  // previous tile's accumulators: drained under this tile's next score block
              epilogue(pend_nt, pend_item, pend_tile, pend_msc);
              pend = false;
            }
          }
        }
        if (pend) {  // not reached: every tile drains its predecessor at its max+exp block (F) / first block (BQ, BKV)
          if (MODE == MODE_F) bar_sync_256();
          epilogue(pend_nt, pend_item, pend_tile, pend_msc);
        }
        pend = true; pend_nt = nt; pend_item = item; pend_tile = tile; pend_msc = msc;
      }
    }
    if (pend) {
      if (MODE == MODE_F) bar_sync_256();  // the last tile's row sums
      epilogue(pend_nt, pend_item, pend_tile, pend_msc);
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc(tmem, 512u);
  }
}

// ============================================ forward, one tile per warpgroup ("ping-pong") ============================================
// Same resident-block ring, block stream and two-pass softmax as MODE_F above, but the two softmax warpgroups do not share a tile: the
// CTA's tile stream alternates between them (tile nt -> warpgroup nt & 1), each with its own Q buffer, TMEM columns (S 128 + O 2 x 64),
// P staging (two 64-key halves) and issuer thread (warp 1 for warpgroup 0, warp 3 for warpgroup 1: score MMA of block n+1, then the two
// half-block output MMAs of block n).  A thread owns a whole query row: no cross-warpgroup exchange and no CTA-wide barrier inside an
// item, and -- the point -- the warpgroups drift apart, so one's exp pass (MUFU) runs under the other's TMEM loads, row-max passes,
// mbarrier waits and epilogue instead of all eight warps queueing for the same unit at the same time.
constexpr int pR1F = 0, pR1E = 3, pR2F = 6, pR2E = 9, pTF = 12, pTE = 14, pSF = 16, pSE = 18, pOPF = 20, pOPE = 24, pOF = 28, pOE = 32, kNumBarsPP = 36;

template <bool DROP>
__global__ void __launch_bounds__(kThreadsA, 1) attn_fwd_pp_kernel(const __grid_constant__ CUtensorMap tmQKV, const AParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  uint8_t* gen = smem_raw + (base - raw);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int S = p.S, H = p.heads * 64, nblk = p.nblk;
  const int ntiles = nblk, L = 2 * nblk - 1;
  const uint32_t bars = base + oMisc;
  auto bar = [&](int i) { return bars + 8u * (uint32_t)i; };
  const uint32_t tmem_slot = bars + 8u * kNumBarsPP;
  volatile uint32_t* tmem_slot_gen = reinterpret_cast<volatile uint32_t*>(gen + oMisc + 8 * kNumBarsPP);
  uint32_t* mws = reinterpret_cast<uint32_t*>(gen + oMisc + 512);  // [2 wg][2 parity][12] key-validity words

  if (warp == 0 && lane == 0) tma_prefetch_desc(&tmQKV);
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < 3; ++i) {
      mbar_init(bar(pR1F + i), 1); mbar_init(bar(pR1E + i), 2);  // K_b / V_b: released by both issuer threads
      mbar_init(bar(pR2F + i), 1); mbar_init(bar(pR2E + i), 2);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(bar(pTF + i), 1); mbar_init(bar(pTE + i), 1);
      mbar_init(bar(pSF + i), 1); mbar_init(bar(pSE + i), 4);
    }
    for (int i = 0; i < 4; ++i) {
      mbar_init(bar(pOPF + i), 4); mbar_init(bar(pOPE + i), 1);
      mbar_init(bar(pOF + i), 1); mbar_init(bar(pOE + i), 4);
    }
    mbar_fence_init();
  }
  if (warp == 2) tmem_alloc(tmem_slot, 512u);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot_gen;
  pdl_enter();

  auto blk_of = [&](int j, int nt) {
    const int i = j < nblk ? j : 2 * nblk - 2 - j;
    return (nt & 1) ? nblk - 1 - i : i;
  };
  auto kind_of = [&](int j) { return j < nblk - 1 ? 0 : (j == nblk - 1 ? 1 : 2); };
  auto rslot = [&](int it, int b) { return (it * nblk + b) % 3; };
  auto rpar = [&](int it, int b) { return (uint32_t)(((it * nblk + b) / 3) & 1); };
  const int my_items = ((int)p.n_items - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;
  const int n_tiles_total = my_items * ntiles;

  if (warp == 0) {
    // =========================================== TMA producer ===========================================
    if (lane == 0) {
      // Issue order inside an item: Q tiles 0 and 1 (one per warpgroup), the K blocks, the V blocks, then the remaining Q tiles.  An
      // issuer thread runs its score MMAs one block AHEAD of its output MMAs, and the V slots of the previous item are only released by
      // those output MMAs: everything a look-ahead score block needs (its Q tile, the K blocks) must therefore sit in front of the V
      // loads in this in-order queue.
      for (int it = 0; it < my_items; ++it) {
        const int item = (int)blockIdx.x + it * (int)gridDim.x;
        const int b_ = item / p.heads, h_ = item % p.heads;
        const int row0 = b_ * S;
        auto load_t = [&](int tile) {
          const int nt = it * ntiles + tile, w = nt & 1;
          mbar_wait(bar(pTE + w), (uint32_t)((nt >> 1) & 1) ^ 1u);
          const uint32_t fb = bar(pTF + w);
          mbar_expect_tx(fb, kBlk);
          const uint32_t d = base + oT + (uint32_t)w * kBlk;
          tma_load_2d(d, &tmQKV, fb, h_ * 64, row0 + tile * 128);
          tma_load_2d(d + 8192u, &tmQKV, fb, h_ * 64, row0 + tile * 128 + 64);
        };
        auto load_r = [&](int which, int bb) {
          const int sl = rslot(it, bb);
          mbar_wait(bar((which ? pR2E : pR1E) + sl), rpar(it, bb) ^ 1u);
          const uint32_t f = bar((which ? pR2F : pR1F) + sl);
          mbar_expect_tx(f, kBlk);
          const uint32_t dd = base + (which ? oR2 : oR1) + (uint32_t)sl * kBlk;
          const int c = which ? 2 * H + h_ * 64 : H + h_ * 64;
          tma_load_2d(dd, &tmQKV, f, c, row0 + bb * 128);
          tma_load_2d(dd + 8192u, &tmQKV, f, c, row0 + bb * 128 + 64);
        };
        load_t(0);
        if (ntiles > 1) load_t(1);
        const int dir0 = (it * ntiles) & 1;
        for (int i = 0; i < nblk; ++i) load_r(0, dir0 ? nblk - 1 - i : i);
        for (int i = 0; i < nblk; ++i) load_r(1, dir0 ? i : nblk - 1 - i);
        for (int tile = 2; tile < ntiles; ++tile) load_t(tile);
      }
    }
  } else if (warp == 1 || warp == 3) {
    // =========================================== MMA issuer of warpgroup w (one thread) ===========================================
    if (lane == 0) {
      const int w = warp == 1 ? 0 : 1;
      const uint32_t id_s = umma_idesc(1u, 128u, 128u, 0u, 0u), id_o = umma_idesc(1u, 128u, 64u, 0u, 1u);
      const uint32_t tS = tmem + (uint32_t)(w * 256);
      const uint32_t sQ = base + oT + (uint32_t)w * kBlk;
      int nb = 0, ne = 0;  // score blocks / exp-kind blocks of this warpgroup so far
      int tr_n = w == 0 ? 0 : 1 << 20;
      auto issue_score = [&](int nt, int j, uint32_t& seen) {
        const int it = nt / ntiles, k = nt >> 1;
        const int b = blk_of(j, nt), kind = kind_of(j), rs = rslot(it, b);
        VB_TR(0, 100 + nb);
        if (j == 0) mbar_wait(bar(pTF + w), (uint32_t)(k & 1));
        if (!(seen & (1u << b))) { mbar_wait(bar(pR1F + rs), rpar(it, b)); seen |= 1u << b; }
        mbar_wait(bar(pSE + w), (uint32_t)(nb & 1) ^ 1u);
        tc_fence_after();
        const uint32_t sK = base + oR1 + (uint32_t)rs * kBlk;
#pragma unroll
        for (int kk = 0; kk < 4; ++kk)
          tc_mma_f16(tS, umma_desc_sw128(sQ + kk * 32u, 16u, 1024u), umma_desc_sw128(sK + kk * 32u, 16u, 1024u), id_s, kk > 0 ? 1u : 0u);
        tc_commit(bar(pSF + w));
        VB_TR(0, 200 + nb);
        if (j == L - 1) tc_commit(bar(pTE + w));
        // this warpgroup's last score-side use of K_b in the item: the exp-kind visit in its last tile of the item
        const bool my_last_tile = (nt + 2) / ntiles != it;
        if (my_last_tile && kind != 0) tc_commit(bar(pR1E + rs));
        ++nb;
      };
      int cur_it = -1;
      uint32_t seen_s = 0, seen_o = 0;
      int seen_s_it = -1;
      // cursor of the score block issued ahead
      int s_nt = w, s_j = 0;
      auto score_next = [&]() {
        if (s_nt >= n_tiles_total) return;
        const int it = s_nt / ntiles;
        if (it != seen_s_it) {
          // items this warpgroup never touches (single-tile items go to the warpgroups alternately): it still owes their K / V slots
          // one arrival each, or the producer would wait for it forever.  Through tcgen05.commit, not a plain arrive: the arrival
          // must stay ordered behind this thread's earlier commits on the same slot (a plain arrive would overtake a commit whose
          // MMAs are still running and complete the slot's PREVIOUS phase too early)
          for (int skipped = seen_s_it + 1; skipped < it; ++skipped)
            for (int b = 0; b < nblk; ++b) { tc_commit(bar(pR1E + rslot(skipped, b))); tc_commit(bar(pR2E + rslot(skipped, b))); }
          seen_s_it = it; seen_s = 0;
        }
        issue_score(s_nt, s_j, seen_s);
        if (++s_j == L) { s_j = 0; s_nt += 2; }
      };
      score_next();
      for (int nt = w; nt < n_tiles_total; nt += 2) {
        const int it = nt / ntiles, k = nt >> 1;
        if (it != cur_it) { cur_it = it; seen_o = 0; }
        const bool my_last_tile = (nt + 2) / ntiles != it;
        const uint32_t acc = tS + 128u + (uint32_t)((k & 1) * 64);
        for (int j = 0; j < L; ++j) {
          score_next();  // block n+1 (its TMEM buffer is free once the warps hold block n in registers)
          const int b = blk_of(j, nt), kind = kind_of(j), rs = rslot(it, b);
          if (kind == 0) continue;
          if (kind == 1) mbar_wait(bar(pOE + w * 2 + (k & 1)), (uint32_t)((k >> 1) & 1) ^ 1u);
          if (!(seen_o & (1u << b))) { mbar_wait(bar(pR2F + rs), rpar(it, b)); seen_o |= 1u << b; }
          const uint32_t sV = base + oR2 + (uint32_t)rs * kBlk;
          // the max+exp block delivers its second half first (still in registers after the max pass), the exp blocks the first
#pragma unroll
          for (int hh = 0; hh < 2; ++hh) {
            const int half = kind == 1 ? 1 - hh : hh;
            VB_TR(0, 500 + ne * 2 + hh);
            mbar_wait(bar(pOPF + w * 2 + half), (uint32_t)(ne & 1));
            tc_fence_after();
            VB_TR(0, 600 + ne * 2 + hh);
            const uint32_t sP = base + oOP + (uint32_t)(w * 2 + half) * kBlk;
#pragma unroll
            for (int ks = 0; ks < 4; ++ks)
              tc_mma_f16(acc, umma_desc_sw128(sP + ks * 32u, 16u, 1024u), umma_desc_sw128(sV + (uint32_t)half * 8192u + ks * 2048u, 8192u, 1024u), id_o,
                         (kind == 1 && hh == 0 && ks == 0) ? 0u : 1u);
            tc_commit(bar(pOPE + w * 2 + half));
          }
          if (my_last_tile) tc_commit(bar(pR2E + rs));
          if (j == L - 1) tc_commit(bar(pOF + w * 2 + (k & 1)));
          ++ne;
        }
      }
      // trailing items of the CTA this warpgroup never touched
      for (int skipped = (seen_s_it < 0 ? 0 : seen_s_it + 1); skipped < my_items; ++skipped)
        for (int b = 0; b < nblk; ++b) { tc_commit(bar(pR1E + rslot(skipped, b))); tc_commit(bar(pR2E + rslot(skipped, b))); }
    }
  } else if (warp >= 4) {
    // =========================================== softmax / epilogue warps ===========================================
    const int ew = warp - 4, w = ew >> 2, q = warp & 3;
    const int row = q * 32 + lane;
    const uint32_t lanef = (uint32_t)(q * 32) << 16;
    const uint32_t tS = tmem + lanef + (uint32_t)(w * 256);
    const uint32_t stage = base + oStage + (uint32_t)ew * 2048u;
    const unsigned long long dseed = DROP ? p.seed + (p.seed_dev != nullptr ? *p.seed_dev : 0ull) : 0ull;
    const uint32_t thr16 = DROP ? (uint32_t)fminf(p.dropout_p * 65536.0f, 65535.0f) : 0u;
    const float inv_keep = DROP ? 1.0f / (1.0f - p.dropout_p) : 1.0f;
    uint32_t* mw_wg = mws + w * 24;
    int nb = 0, ne = 0;
    int tr_n = (warp == 4 && lane == 0) ? 0 : 1 << 20;
    bool pend = false;
    int pend_k = 0, pend_item = 0, pend_tile = 0;
    float pend_msc = 0.f, pend_l = 0.f;

    auto epilogue = [&](int k, int item, int tile, float msc, float l) {
      const int b_ = item / p.heads, h_ = item % p.heads;
      VB_TR(1, 8000 + k);
      mbar_wait(bar(pOF + w * 2 + (k & 1)), (uint32_t)((k >> 1) & 1));
      tc_fence_after();
      VB_TR(1, 8500 + k);
      uint32_t ra[32], rb[32];
      tmem_ld32(tS + 128u + (uint32_t)((k & 1) * 64), ra);
      tmem_ld32(tS + 128u + (uint32_t)((k & 1) * 64 + 32), rb);
      tmem_ld_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(bar(pOE + w * 2 + (k & 1)));
      const float mul = l > 0.f ? 1.f / l : 0.f;
      const int r0g = tile * 128 + q * 32;
      bf16* g0 = p.ctx + ((long long)b_ * S + r0g) * H + h_ * 64;
      store_rows32_staged(ra, mul, stage, g0, (long long)H, S - r0g, lane);
      store_rows32_staged(rb, mul, stage, g0 + 32, (long long)H, S - r0g, lane);
      if (r0g + lane < S && p.lse != nullptr) p.lse[((long long)b_ * p.heads + h_) * S + r0g + lane] = msc * kLn2A + __logf(fmaxf(l, 1e-30f));
      VB_TR(1, 9000 + k);
    };
    auto mask_bit = [&](int item, int wd) -> bool {
      if (item >= p.n_items || wd >= nblk * 4) return false;
      const int key = 32 * wd + lane;
      return key < S && p.key_mask[(long long)(item / p.heads) * S + key] != 0;
    };

    int cur_it = -1, ci = 0;  // ci: items this warpgroup has started (double-buffers its mask words)
    for (int nt = w, k = 0; nt < n_tiles_total; nt += 2, ++k) {
      const int it = nt / ntiles, tile = nt % ntiles;
      const int item = (int)blockIdx.x + it * (int)gridDim.x;
      const long long bh = item;  // = b * heads + h
      if (it != cur_it) {
        // key-validity words of the item, per warpgroup: warp q computes words q, q+4, q+8 (a 128-thread barrier publishes them; a warp is
        // at most one such barrier ahead of its neighbours, hence two buffers)
        cur_it = it;
        ++ci;
#pragma unroll
        for (int i = 0; i < 3; ++i) {
          const uint32_t wv = __ballot_sync(0xffffffffu, mask_bit(item, q + 4 * i));
          if (lane == 0) mw_wg[(ci & 1) * 12 + q + 4 * i] = wv;
        }
        asm volatile("bar.sync %0, 128;" ::"r"(2 + w) : "memory");
      }
      const uint32_t* mw_item = mw_wg + (ci & 1) * 12;
      const bool warp_active = tile * 128 + q * 32 < S;
      float m_run = -INFINITY, msc = 0.f, l = 0.f;
      const unsigned long long rowbase = DROP ? ((unsigned long long)(bh * S + tile * 128 + row) << 6) : 0ull;

      auto rmax = [&](const uint32_t (&r)[32], uint32_t mw) {
        if (mw == 0xffffffffu) {
          float m0 = __uint_as_float(r[0]), m1 = __uint_as_float(r[1]), m2 = __uint_as_float(r[2]), m3 = __uint_as_float(r[3]);
#pragma unroll
          for (int i = 4; i < 32; i += 4) {
            m0 = fmaxf(m0, __uint_as_float(r[i])); m1 = fmaxf(m1, __uint_as_float(r[i + 1]));
            m2 = fmaxf(m2, __uint_as_float(r[i + 2])); m3 = fmaxf(m3, __uint_as_float(r[i + 3]));
          }
          m_run = fmaxf(m_run, fmaxf(fmaxf(m0, m1), fmaxf(m2, m3)));
        } else if (mw != 0u) {
          float m0 = -INFINITY, m1 = -INFINITY;
#pragma unroll
          for (int i = 0; i < 32; i += 2) {
            m0 = fmaxf(m0, ((mw >> i) & 1u) ? __uint_as_float(r[i]) : -INFINITY);
            m1 = fmaxf(m1, ((mw >> (i + 1)) & 1u) ? __uint_as_float(r[i + 1]) : -INFINITY);
          }
          m_run = fmaxf(m_run, fmaxf(m0, m1));
        }
      };
      // exp of 32 columns (chunk c of half `half` of block b) -> bf16 P in the half's staging chunk
      auto pexp = [&](const uint32_t (&r)[32], uint32_t mw, int b, int half, int c) {
        uint32_t pk[16];
        const uint32_t sP = base + oOP + (uint32_t)(w * 2 + half) * kBlk;
        if (DROP) {
          if (warp_active && mw != 0u) {
            const unsigned long long ctr = rowbase + (unsigned)((b * 128 + half * 64 + c * 32) >> 3);
#pragma unroll
            for (int jj = 0; jj < 4; ++jj) {
              const uint4 rb = Philox(dseed)(ctr + (unsigned)jj, p.site);
#pragma unroll
              for (int e = 0; e < 4; ++e) {
                const int i = 4 * jj + e;
                const float p0 = ((mw >> (2 * i)) & 1u) ? fast_exp2(fmaf(__uint_as_float(r[2 * i]), p.scale_log2, -msc)) : 0.f;
                const float p1 = ((mw >> (2 * i + 1)) & 1u) ? fast_exp2(fmaf(__uint_as_float(r[2 * i + 1]), p.scale_log2, -msc)) : 0.f;
                l += p0 + p1;
                const uint32_t wd = e == 0 ? rb.x : (e == 1 ? rb.y : (e == 2 ? rb.z : rb.w));
                pk[i] = pack_bf16x2((wd & 0xffffu) >= thr16 ? p0 * inv_keep : 0.f, (wd >> 16) >= thr16 ? p1 * inv_keep : 0.f);
              }
            }
          } else {
#pragma unroll
            for (int i = 0; i < 16; ++i) pk[i] = 0u;
          }
        } else if (warp_active && mw == 0xffffffffu) {
          float l0 = 0.f, l1 = 0.f;
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            const float p0 = fast_exp2(fmaf(__uint_as_float(r[2 * i]), p.scale_log2, -msc));
            const float p1 = fast_exp2(fmaf(__uint_as_float(r[2 * i + 1]), p.scale_log2, -msc));
            l0 += p0; l1 += p1;
            pk[i] = pack_bf16x2(p0, p1);
          }
          l += l0 + l1;
        } else if (warp_active && mw != 0u) {
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            const float p0 = ((mw >> (2 * i)) & 1u) ? fast_exp2(fmaf(__uint_as_float(r[2 * i]), p.scale_log2, -msc)) : 0.f;
            const float p1 = ((mw >> (2 * i + 1)) & 1u) ? fast_exp2(fmaf(__uint_as_float(r[2 * i + 1]), p.scale_log2, -msc)) : 0.f;
            l += p0 + p1;
            pk[i] = pack_bf16x2(p0, p1);
          }
        } else {
#pragma unroll
          for (int i = 0; i < 16; ++i) pk[i] = 0u;
        }
        store_row32a(sP, row, c, pk);
      };
      auto release_s = [&]() {
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(bar(pSE + w));
      };
      auto publish_half = [&](int half) {
        fence_async_smem_a();
        __syncwarp();
        if (lane == 0) mbar_arrive(bar(pOPF + w * 2 + half));
      };

      for (int j = 0; j < L; ++j, ++nb) {
        const int b = blk_of(j, nt), kind = kind_of(j);
        VB_TR(1, 1000 + nb);
        mbar_wait(bar(pSF + w), (uint32_t)(nb & 1));
        tc_fence_after();
        VB_TR(1, 2000 + nb);
        uint32_t r0[32], r1[32];
        if (kind == 0) {
          // ---- row maximum only: both halves through the same 64 registers ----
#pragma unroll
          for (int half = 0; half < 2; ++half) {
            tmem_ld32(tS + (uint32_t)(half * 64), r0);
            tmem_ld32(tS + (uint32_t)(half * 64 + 32), r1);
            tmem_ld_wait();
            if (half == 1) release_s();
            if (warp_active) { rmax(r0, mw_item[b * 4 + half * 2]); rmax(r1, mw_item[b * 4 + half * 2 + 1]); }
          }
        } else if (kind == 1) {
          // ---- max over the whole row, then exp: second half straight from the registers of the max pass, first half re-read ----
          tmem_ld32(tS, r0);
          tmem_ld32(tS + 32u, r1);
          tmem_ld_wait();
          if (warp_active) { rmax(r0, mw_item[b * 4]); rmax(r1, mw_item[b * 4 + 1]); }
          tmem_ld32(tS + 64u, r0);
          tmem_ld32(tS + 96u, r1);
          mbar_wait(bar(pOPE + w * 2 + 1), (uint32_t)(ne & 1) ^ 1u);
          tmem_ld_wait();
          if (warp_active) { rmax(r0, mw_item[b * 4 + 2]); rmax(r1, mw_item[b * 4 + 3]); }
          msc = m_run == -INFINITY ? 0.f : m_run * p.scale_log2;
          VB_TR(1, 4000 + nb);
          pexp(r0, mw_item[b * 4 + 2], b, 1, 0);
          pexp(r1, mw_item[b * 4 + 3], b, 1, 1);
          publish_half(1);
          VB_TR(1, 5000 + nb);
          tmem_ld32(tS, r0);
          tmem_ld32(tS + 32u, r1);
          mbar_wait(bar(pOPE + w * 2 + 0), (uint32_t)(ne & 1) ^ 1u);
          tmem_ld_wait();
          release_s();
          VB_TR(1, 6000 + nb);
          pexp(r0, mw_item[b * 4], b, 0, 0);
          pexp(r1, mw_item[b * 4 + 1], b, 0, 1);
          publish_half(0);
          ++ne;
        } else {
#pragma unroll
          for (int half = 0; half < 2; ++half) {
            tmem_ld32(tS + (uint32_t)(half * 64), r0);
            tmem_ld32(tS + (uint32_t)(half * 64 + 32), r1);
            mbar_wait(bar(pOPE + w * 2 + half), (uint32_t)(ne & 1) ^ 1u);
            tmem_ld_wait();
            if (half == 1) release_s();
            VB_TR(1, 6000 + nb);
            pexp(r0, mw_item[b * 4 + half * 2], b, half, 0);
            pexp(r1, mw_item[b * 4 + half * 2 + 1], b, half, 1);
            publish_half(half);
          }
          ++ne;
        }
        VB_TR(1, 7000 + nb);
        if (pend && j == 0) {  // this warpgroup's previous tile: drained while the tensor pipe works on the next score block
          epilogue(pend_k, pend_item, pend_tile, pend_msc, pend_l);
          pend = false;
        }
      }
      if (pend) epilogue(pend_k, pend_item, pend_tile, pend_msc, pend_l);  // (not reached: every tile drains its predecessor at j == 0)
      pend = true; pend_k = k; pend_item = item; pend_tile = tile; pend_msc = msc; pend_l = l;
    }
    if (pend) epilogue(pend_k, pend_item, pend_tile, pend_msc, pend_l);
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc(tmem, 512u);
  }
}

// delta[b,h,q] = sum_d dO[q,d] O[q,d]  (one 16-byte segment per lane, 8 lanes per (row, head))
__global__ void __launch_bounds__(256) attn_delta_kernel(const bf16* __restrict__ ctx, const bf16* __restrict__ dctx, float* __restrict__ delta, int B, int S,
                                                         int heads) {
  pdl_enter();
  const long long idx = (long long)blockIdx.x * 256 + threadIdx.x;  // (row, head, seg)
  const long long total = (long long)B * S * heads * 8;
  float d = 0.f;
  long long rh = idx >> 3;
  if (idx < total) {
    const uint4 x = *reinterpret_cast<const uint4*>(dctx + idx * 8), y = *reinterpret_cast<const uint4*>(ctx + idx * 8);
    float2 u, v;
    u = unpack_bf16x2(x.x); v = unpack_bf16x2(y.x); d += u.x * v.x + u.y * v.y;
    u = unpack_bf16x2(x.y); v = unpack_bf16x2(y.y); d += u.x * v.x + u.y * v.y;
    u = unpack_bf16x2(x.z); v = unpack_bf16x2(y.z); d += u.x * v.x + u.y * v.y;
    u = unpack_bf16x2(x.w); v = unpack_bf16x2(y.w); d += u.x * v.x + u.y * v.y;
  }
  d += __shfl_xor_sync(0xffffffffu, d, 1);
  d += __shfl_xor_sync(0xffffffffu, d, 2);
  d += __shfl_xor_sync(0xffffffffu, d, 4);
  if (idx < total && (idx & 7) == 0) {
    const long long row = rh / heads;
    const int h = (int)(rh % heads);
    const long long b = row / S, q = row % S;
    delta[(b * heads + h) * S + q] = d;
  }
}

template <typename K>
int set_smem_a(K kernel) {
  cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemA);
  if (e != cudaSuccess) return fail(VAULT_ERR_LAUNCH, "attention_sm100: cudaFuncSetAttribute(%d): %s", (int)kSmemA, cudaGetErrorString(e));
  return VAULT_OK;
}

long long* trace_buf() {
  static long long* buf = nullptr;
  const char* e = getenv("VAULT_B200_ATTN_TRACE");
  if (e == nullptr || e[0] != '1') return nullptr;
  if (!buf) cudaMalloc(&buf, 2 * 1024 * 2 * sizeof(long long));
  cudaMemset(buf, 0, 2 * 1024 * 2 * sizeof(long long));
  return buf;
}
void trace_dump(const char* what, long long* dbuf) {
  if (!dbuf) return;
  static long long h[2 * 1024 * 2];
  cudaDeviceSynchronize();
  cudaMemcpy(h, dbuf, sizeof(h), cudaMemcpyDeviceToHost);
  long long t0 = h[0] ? h[0] : h[2048];
  for (int who = 0; who < 2; ++who) {
    printf("TRACE %s %s:", what, who ? "softmax" : "mma");
    for (int i = 0; i < 1024 && h[(who * 1024 + i) * 2]; ++i) printf(" %lld@%lld", h[(who * 1024 + i) * 2 + 1], h[(who * 1024 + i) * 2] - t0);
    printf("\n");
  }
  fflush(stdout);
}

int fill(AParams& p, int B, int S, int heads) {
  p.B = B; p.S = S; p.heads = heads;
  p.nblk = (S + 127) / 128;
  p.n_items = B * heads;
  p.scale = 0.125f;
  p.scale_log2 = 0.125f * kLog2eA;
  return VAULT_OK;
}

}  // namespace

int g_attn_sm100 = 1;  // 0: disabled (A/B switch through vault_attn_set_impl)
int g_attn_fwd_pp = 0;

// Shapes served: no dropout, 193..384 keys (measured on B200, B = 32, 12 heads: S = 369 forward 56 us / backward 125 us against 75 / 313 us
// for the mma.sync kernels; at S <= 192 the whole-row kernels of attention_tc.cu -- 2 CTAs / SM, no key-block loop -- are still ahead:
// 22 / 49 us against 28 / 56 us at S = 185).  g_attn_sm100 = 2 forces these kernels for every S <= 384 (tests).
bool attn_sm100_ok(int S, float dropout_p) {
  if (g_attn_sm100 == 0 || S < 1 || S > 384) return false;
  if (g_attn_sm100 == 2) return true;
  // with probability dropout (the LM stack in training) the alternative is the mma.sync family: ahead only up to 64 keys
  return dropout_p > 0.f ? S > 64 : S > 192;
}
void attn_sm100_enable(int on) { g_attn_sm100 = on; }
void attn_sm100_fwd_pp(int on) { g_attn_fwd_pp = on; }

int attn_fwd_sm100(const void* qkv, const uint8_t* key_mask, void* ctx, float* lse, int B, int S, int heads, float dropout_p, unsigned long long seed,
                   const unsigned long long* seed_dev, unsigned site, cudaStream_t st) {
  AParams p{};
  fill(p, B, S, heads);
  p.dropout_p = dropout_p; p.seed = seed; p.seed_dev = seed_dev; p.site = site;
  p.key_mask = key_mask; p.ctx = reinterpret_cast<bf16*>(ctx); p.lse = lse;
  const int H = heads * 64;
  CUtensorMap tm;
  int rc = encode_tmap_2d(&tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, qkv, (uint64_t)3 * H, (uint64_t)B * S, (uint64_t)3 * H, 64, 64);
  if (rc) return rc;
  const int grid_pp = p.n_items < device_sm_count() ? p.n_items : device_sm_count();
  const long long tiles = (long long)p.n_items * p.nblk;  // the shared-tile kernels split the launch's tiles, not its items, over the CTAs
  const int grid = (int)(tiles < device_sm_count() ? tiles : device_sm_count());
  p.trace = trace_buf();
  // g_attn_fwd_pp (vault_attn_set_impl(4)): the forward with one tile per warpgroup instead of the shared-tile MODE_F.  Measured on B200
  // (B = 32, 12 heads): S = 369 60.5 us against 56.0 us, S = 209 (B = 64) 49.1 against 53.4 us -- the warpgroups do drift apart as intended,
  // but each now walks whole 128-column blocks alone behind a single score buffer (a row-max-only block waits ~300 cycles for the next
  // score MMA), and because they finish an item at different times the next item's K / V blocks, released only when BOTH have retired
  // their last MMA on a slot, arrive late: ~8 k idle cycles per item.  Kept as a tested alternative, not routed by default.
  if (g_attn_fwd_pp) {
    if (dropout_p > 0.f) {
      if ((rc = set_smem_a(attn_fwd_pp_kernel<true>))) return rc;
      launch(attn_fwd_pp_kernel<true>, dim3(grid_pp), dim3(kThreadsA), (size_t)kSmemA, st, tm, p);
    } else {
      if ((rc = set_smem_a(attn_fwd_pp_kernel<false>))) return rc;
      launch(attn_fwd_pp_kernel<false>, dim3(grid_pp), dim3(kThreadsA), (size_t)kSmemA, st, tm, p);
    }
  } else if (dropout_p > 0.f) {
    if ((rc = set_smem_a(attn_sm100_kernel<MODE_F, true>))) return rc;
    launch(attn_sm100_kernel<MODE_F, true>, dim3(grid), dim3(kThreadsA), (size_t)kSmemA, st, tm, tm, p);
  } else {
    if ((rc = set_smem_a(attn_sm100_kernel<MODE_F, false>))) return rc;
    launch(attn_sm100_kernel<MODE_F, false>, dim3(grid), dim3(kThreadsA), (size_t)kSmemA, st, tm, tm, p);
  }
  trace_dump("F", p.trace);
  return check_launch("attn_sm100_kernel<F>");
}

int attn_bwd_sm100(const void* qkv, const uint8_t* key_mask, const void* ctx, const void* dctx, const float* lse, float* delta, void* dqkv, int B, int S,
                   int heads, float dropout_p, unsigned long long seed, const unsigned long long* seed_dev, unsigned site, cudaStream_t st) {
  AParams p{};
  fill(p, B, S, heads);
  p.dropout_p = dropout_p; p.seed = seed; p.seed_dev = seed_dev; p.site = site;
  p.key_mask = key_mask; p.lse = const_cast<float*>(lse); p.delta = delta; p.dqkv = reinterpret_cast<bf16*>(dqkv);
  const int H = heads * 64;
  CUtensorMap tmQ, tmD;
  int rc = encode_tmap_2d(&tmQ, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, qkv, (uint64_t)3 * H, (uint64_t)B * S, (uint64_t)3 * H, 64, 64);
  if (rc) return rc;
  rc = encode_tmap_2d(&tmD, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, dctx, (uint64_t)H, (uint64_t)B * S, (uint64_t)H, 64, 64);
  if (rc) return rc;
  const bool drop = dropout_p > 0.f;
  if ((rc = drop ? set_smem_a(attn_sm100_kernel<MODE_BQ, true>) : set_smem_a(attn_sm100_kernel<MODE_BQ, false>))) return rc;
  if ((rc = drop ? set_smem_a(attn_sm100_kernel<MODE_BKV, true>) : set_smem_a(attn_sm100_kernel<MODE_BKV, false>))) return rc;
  const long long total = (long long)B * S * heads * 8;
  launch(attn_delta_kernel, dim3((unsigned)((total + 255) / 256)), dim3(256), 0, st, reinterpret_cast<const bf16*>(ctx), reinterpret_cast<const bf16*>(dctx), delta,
         B, S, heads);
  if ((rc = check_launch("attn_delta_kernel"))) return rc;
  const long long tiles = (long long)p.n_items * p.nblk;
  const int grid = (int)(tiles < device_sm_count() ? tiles : device_sm_count());
  p.trace = trace_buf();
  if (drop) launch(attn_sm100_kernel<MODE_BQ, true>, dim3(grid), dim3(kThreadsA), (size_t)kSmemA, st, tmQ, tmD, p);
  else launch(attn_sm100_kernel<MODE_BQ, false>, dim3(grid), dim3(kThreadsA), (size_t)kSmemA, st, tmQ, tmD, p);
  trace_dump("BQ", p.trace);
  if ((rc = check_launch("attn_sm100_kernel<BQ>"))) return rc;
  p.trace = trace_buf();
  if (drop) launch(attn_sm100_kernel<MODE_BKV, true>, dim3(grid), dim3(kThreadsA), (size_t)kSmemA, st, tmQ, tmD, p);
  else launch(attn_sm100_kernel<MODE_BKV, false>, dim3(grid), dim3(kThreadsA), (size_t)kSmemA, st, tmQ, tmD, p);
  trace_dump("BKV", p.trace);
  return check_launch("attn_sm100_kernel<BKV>");
}

}  // namespace vb
