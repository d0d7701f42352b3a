// Pipelined, warp-specialised fused attention for sm_100a (tcgen05.mma + tensor memory + TMA), head_dim 64, S <= 384, no dropout:
// the ViLT stack's masked-softmax attention (HF:models/vilt/modeling_vilt.py:306-365) forward and backward at every shape the
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
};

// shared memory map (bytes from the 1024-aligned base)
constexpr uint32_t kBlk = 16384;                 // one [128 x 64] bf16 operand block (two 64-row TMA boxes)
constexpr uint32_t oR1 = 0, oR2 = 3 * kBlk;      // resident operands, up to 3 blocks each
constexpr uint32_t oT = 6 * kBlk;                // per-tile operands: F: Q_t x 2 (double buffer); B: T1, T2
constexpr uint32_t oOP = 8 * kBlk;               // P / dS staging: 4 x 16 KB
constexpr uint32_t oMisc = 12 * kBlk;            // barriers, exchange arrays
constexpr uint32_t kSmemA = 12 * kBlk + 6144 + 1024;  // misc: 512 B barriers + 2 KB xm + 2 KB xl + mask words

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
__device__ __forceinline__ void bar_sync_256() { asm volatile("bar.sync 1, 256;" ::: "memory"); }

// position in this CTA's stream of score blocks
struct Cursor {
  int it, tile, j, n, nt;
};

template <int MODE>
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
  constexpr int kRing = MODE == MODE_F ? 3 : 2;        // TMEM score ring
  constexpr int kLook = MODE == MODE_F ? 2 : 1;        // score MMAs issued ahead of the output MMAs
  constexpr bool kTwoAcc = MODE != MODE_BKV;           // double-buffered accumulators
  constexpr uint32_t colDP = 256, colAcc = 384;

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
    for (int i = 0; i < 3; ++i) {
      mbar_init(bar(bR1F + i), 1); mbar_init(bar(bR1E + i), 1); mbar_init(bar(bR2F + i), 1); mbar_init(bar(bR2E + i), 1);
      mbar_init(bar(bSF + i), 1); mbar_init(bar(bSE + i), 8);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(bar(bTF + i), 1); mbar_init(bar(bTE + i), 1); mbar_init(bar(bOF + i), 1); mbar_init(bar(bOE + i), 8);
    }
    mbar_init(bar(bDPF), 1); mbar_init(bar(bDPE), 8);
    for (int i = 0; i < 4; ++i) {
      mbar_init(bar(bOPF + i), MODE == MODE_BKV ? 8 : 4);
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

  auto valid = [&](const Cursor& c) { return (int)blockIdx.x + c.it * (int)gridDim.x < p.n_items; };
  auto advance = [&](Cursor& c) {
    ++c.n;
    if (++c.j == L) {
      c.j = 0; ++c.nt;
      if (++c.tile == ntiles) { c.tile = 0; ++c.it; }
    }
  };
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

  if (warp == 0) {
    // =========================================== TMA producer ===========================================
    if (lane == 0) {
      int nt = 0;
      for (int it = 0; (int)blockIdx.x + it * (int)gridDim.x < p.n_items; ++it) {
        const int item = (int)blockIdx.x + it * (int)gridDim.x;
        const int b_ = item / p.heads, h_ = item % p.heads;
        const int row0 = b_ * S;
        for (int tile = 0; tile < ntiles; ++tile, ++nt) {
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
          if (tile == 0) {
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
    // =========================================== MMA issuer (one thread) ===========================================
    if (lane == 0) {
      const uint32_t id_s = umma_idesc(1u, 128u, 128u, 0u, 0u);    // score blocks: both operands K-major (dh contiguous)
      const uint32_t id_o = umma_idesc(1u, 128u, 64u, 0u, 1u);     // F / BQ output: A = P / dS chunk K-major, B = V / K rows MN-major
      const uint32_t id_t = umma_idesc(1u, 128u, 64u, 1u, 1u);     // BKV output: A = P / dS read MN-major (M = keys), B = dO / Q rows MN-major
      int opn = 0;                                                 // exp-kind blocks whose output MMAs have been issued (operand ring position)
      // Every wait below is on a barrier the MMA thread really needs; `probe` turns the same sequence into a non-blocking readiness test,
      // used for the score blocks issued AHEAD of the output MMAs (a look-ahead that blocked could wait for a resident block of the
      // next item whose slot is only released by an output MMA this thread has not issued yet).
      auto need = [&](uint32_t barrier, uint32_t parity, bool probe) -> bool {
        if (probe) return mbar_test_wait(barrier, parity);
        mbar_wait(barrier, parity);
        return true;
      };
      auto scores_waits = [&](const Cursor& c, bool probe) -> bool {
        const int b = blk_of(c.j, c.nt);
        const int ts = MODE == MODE_F ? (c.nt & 1) : 0;
        if (c.j == 0 && !need(bar(bTF + ts), (uint32_t)((MODE == MODE_F ? (c.nt >> 1) : c.nt) & 1), probe)) return false;
        if (!need(bar(bR1F + rslot(c.it, b)), rpar(c.it, b), probe)) return false;
        if (!need(bar(bSE + c.n % kRing), (uint32_t)((c.n / kRing) & 1) ^ 1u, probe)) return false;
        if (MODE != MODE_F) {
          if (!need(bar(bR2F + rslot(c.it, b)), rpar(c.it, b), probe)) return false;
          if (!need(bar(bDPE), (uint32_t)(c.n & 1) ^ 1u, probe)) return false;
        }
        return true;
      };
      auto issue_scores = [&](const Cursor& c) {
        const int b = blk_of(c.j, c.nt), kind = kind_of(c.j);
        scores_waits(c, false);
        tc_fence_after();
        const int ts = MODE == MODE_F ? (c.nt & 1) : 0;
        const int sl = c.n % kRing, rs = rslot(c.it, b);
        const uint32_t sT1 = base + oT + (MODE == MODE_F ? (uint32_t)ts * kBlk : 0u), sT2 = base + oT + kBlk;
        const uint32_t sR1b = base + oR1 + (uint32_t)rs * kBlk, sR2b = base + oR2 + (uint32_t)rs * kBlk;
        const uint32_t aS = MODE == MODE_BKV ? sR1b : sT1, bS = MODE == MODE_BKV ? sT1 : sR1b;
#pragma unroll
        for (int k = 0; k < 4; ++k)
          tc_mma_f16(tmem + (uint32_t)(sl * 128), umma_desc_sw128(aS + k * 32u, 16u, 1024u), umma_desc_sw128(bS + k * 32u, 16u, 1024u), id_s, k > 0 ? 1u : 0u);
        tc_commit(bar(bSF + sl));
        if (MODE != MODE_F) {
          const uint32_t aD = MODE == MODE_BKV ? sR2b : sT2, bD = MODE == MODE_BKV ? sT2 : sR2b;
#pragma unroll
          for (int k = 0; k < 4; ++k)
            tc_mma_f16(tmem + colDP, umma_desc_sw128(aD + k * 32u, 16u, 1024u), umma_desc_sw128(bD + k * 32u, 16u, 1024u), id_s, k > 0 ? 1u : 0u);
          tc_commit(bar(bDPF));
        }
        const bool last_tile = c.tile == ntiles - 1;
        if (c.j == L - 1) tc_commit(bar(bTE + ts));                                    // the tile's operands are dead
        // K_b's last use in this item: its exp-kind visit in the last tile (every block is visited exactly once with kind != 0 there)
        if (MODE == MODE_F && last_tile && kind != 0) tc_commit(bar(bR1E + rs));
        if (MODE == MODE_BQ && last_tile) tc_commit(bar(bR2E + rs));                   // V_b's last use (dP)
      };
      auto out_waits = [&](const Cursor& c, bool probe) -> bool {
        const int b = blk_of(c.j, c.nt), kind = kind_of(c.j);
        if (kind == 0) return true;
        const bool first = MODE == MODE_F ? kind == 1 : c.j == 0;
        const int ob = kTwoAcc ? (c.nt & 1) : 0;
        if (first && !need(bar(bOE + ob), (uint32_t)((kTwoAcc ? (c.nt >> 1) : c.nt) & 1) ^ 1u, probe)) return false;
        if (MODE == MODE_BKV) {
          if (!need(bar(bOPF + 0), (uint32_t)(opn & 1), probe)) return false;
          if (!need(bar(bOPF + 1), (uint32_t)(opn & 1), probe)) return false;
        } else {
          if (MODE == MODE_F && !need(bar(bR2F + rslot(c.it, b)), rpar(c.it, b), probe)) return false;
          if (!need(bar(bOPF + 0 * 2 + (opn & 1)), (uint32_t)((opn >> 1) & 1), probe)) return false;
          if (!need(bar(bOPF + 1 * 2 + (opn & 1)), (uint32_t)((opn >> 1) & 1), probe)) return false;
        }
        return true;
      };
      auto issue_out = [&](const Cursor& c) {
        const int b = blk_of(c.j, c.nt), kind = kind_of(c.j);
        if (kind == 0) return;
        out_waits(c, false);
        tc_fence_after();
        const bool first = MODE == MODE_F ? kind == 1 : c.j == 0;
        const bool last = c.j == L - 1;
        const bool last_tile = c.tile == ntiles - 1;
        const int ob = kTwoAcc ? (c.nt & 1) : 0;
        const int slot = opn & 1, rs = rslot(c.it, b);
        if (MODE != MODE_BKV) {
          const uint32_t acc = tmem + colAcc + (uint32_t)(ob * 64);
          const uint32_t sRb = base + (MODE == MODE_F ? oR2 : oR1) + (uint32_t)rs * kBlk;
#pragma unroll
          for (int g = 0; g < 2; ++g) {
            const uint32_t sP = base + oOP + (uint32_t)(g * 2 + slot) * kBlk;
#pragma unroll
            for (int ks = 0; ks < 4; ++ks)
              tc_mma_f16(acc, umma_desc_sw128(sP + ks * 32u, 16u, 1024u), umma_desc_sw128(sRb + (uint32_t)g * 8192u + ks * 2048u, 8192u, 1024u), id_o,
                         (first && g == 0 && ks == 0) ? 0u : 1u);
            tc_commit(bar(bOPE + g * 2 + slot));
          }
          if (last_tile) tc_commit(bar((MODE == MODE_F ? bR2E : bR1E) + rs));  // F: V_b, BQ: K_b -- last use in this item
        } else {
          // dV_t += P_b^T dO_b ; dK_t += dS_b^T Q_b  (A read MN-major: M = keys, two 64-key chunks 16 KB apart; K = the block's 128 queries)
          const uint32_t sP = base + oOP, sDS = base + oOP + 2 * kBlk;
          const uint32_t sQb = base + oR1 + (uint32_t)rs * kBlk, sDOb = base + oR2 + (uint32_t)rs * kBlk;
#pragma unroll
          for (int ks = 0; ks < 8; ++ks)
            tc_mma_f16(tmem + colAcc, umma_desc_sw128(sP + ks * 2048u, 16384u, 1024u), umma_desc_sw128(sDOb + ks * 2048u, 8192u, 1024u), id_t,
                       (first && ks == 0) ? 0u : 1u);
          tc_commit(bar(bOPE + 0));
#pragma unroll
          for (int ks = 0; ks < 8; ++ks)
            tc_mma_f16(tmem + colAcc + 64u, umma_desc_sw128(sDS + ks * 2048u, 16384u, 1024u), umma_desc_sw128(sQb + ks * 2048u, 8192u, 1024u), id_t,
                       (first && ks == 0) ? 0u : 1u);
          tc_commit(bar(bOPE + 1));
          if (last_tile) { tc_commit(bar(bR1E + rs)); tc_commit(bar(bR2E + rs)); }
        }
        if (last) tc_commit(bar(bOF + ob));
        ++opn;
      };
      Cursor cs{0, 0, 0, 0, 0}, co{0, 0, 0, 0, 0};
      while (valid(co)) {
        while (valid(cs) && cs.n <= co.n) { issue_scores(cs); advance(cs); }  // the block the output MMAs below consume
        // until the output operands of block co are ready, keep issuing score blocks ahead whenever one can go without waiting
        for (;;) {
          if (valid(cs) && cs.n <= co.n + kLook && scores_waits(cs, true)) { issue_scores(cs); advance(cs); continue; }
          if (out_waits(co, true)) break;
        }
        issue_out(co);
        advance(co);
      }
    }
  } else if (warp >= 4) {
    // =========================================== softmax / epilogue warps ===========================================
    const int ew = warp - 4, g = ew >> 2, q = warp & 3;
    const int row = q * 32 + lane;
    const uint32_t lanef = (uint32_t)(q * 32) << 16;
    int opn = 0;
    // deferred epilogue state (F / BQ): the accumulator of tile `pend_nt` is drained under the next tile's first block
    bool pend = false;
    int pend_nt = 0, pend_item = 0, pend_tile = 0;
    float pend_msc = 0.f;

    auto epilogue_fq = [&](int nt, int item, int tile, float msc) {
      const int b_ = item / p.heads, h_ = item % p.heads;
      const int ob = nt & 1;
      mbar_wait(bar(bOF + ob), (uint32_t)((nt >> 1) & 1));
      tc_fence_after();
      float mul = p.scale;
      float l_tot = 0.f;
      if (MODE == MODE_F) {
        bar_sync_256();
        l_tot = xl[(nt & 1) * 256 + row] + xl[(nt & 1) * 256 + 128 + row];
        mul = l_tot > 0.f ? 1.f / l_tot : 0.f;
      }
      uint32_t r[32];
      tmem_ld32(tmem + lanef + colAcc + (uint32_t)(ob * 64 + g * 32), r);
      tmem_ld_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(bar(bOE + ob));
      const int qg = tile * 128 + row;
      if (qg < S) {
        uint4* dst;
        if (MODE == MODE_F) dst = reinterpret_cast<uint4*>(p.ctx + ((long long)b_ * S + qg) * H + h_ * 64 + g * 32);
        else dst = reinterpret_cast<uint4*>(p.dqkv + ((long long)b_ * S + qg) * 3LL * H + h_ * 64 + g * 32);
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          uint4 v;
          v.x = pack_bf16x2(__uint_as_float(r[8 * i]) * mul, __uint_as_float(r[8 * i + 1]) * mul);
          v.y = pack_bf16x2(__uint_as_float(r[8 * i + 2]) * mul, __uint_as_float(r[8 * i + 3]) * mul);
          v.z = pack_bf16x2(__uint_as_float(r[8 * i + 4]) * mul, __uint_as_float(r[8 * i + 5]) * mul);
          v.w = pack_bf16x2(__uint_as_float(r[8 * i + 6]) * mul, __uint_as_float(r[8 * i + 7]) * mul);
          dst[i] = v;
        }
        if (MODE == MODE_F && g == 0 && p.lse != nullptr)
          p.lse[((long long)b_ * p.heads + h_) * S + qg] = msc * kLn2A + __logf(fmaxf(l_tot, 1e-30f));
      }
    };

    // key-validity bytes of the NEXT item are fetched one item ahead (two keys per lane per warp: words ew and ew + 8)
    auto mask_bytes = [&](int item, int w) -> uint32_t {
      if (item >= p.n_items || w >= nblk * 4) return 0u;
      const int key = 32 * w + lane;
      return (key < S && p.key_mask[(long long)(item / p.heads) * S + key] != 0) ? 1u : 0u;
    };
    uint32_t nx0 = mask_bytes((int)blockIdx.x, ew), nx1 = mask_bytes((int)blockIdx.x, ew + 8);

    int n = 0, nt = 0;
    for (int it = 0; (int)blockIdx.x + it * (int)gridDim.x < p.n_items; ++it) {
      const int item = (int)blockIdx.x + it * (int)gridDim.x;
      const int b_ = item / p.heads, h_ = item % p.heads;
      const long long bh = (long long)b_ * p.heads + h_;
      {
        const uint32_t w0 = __ballot_sync(0xffffffffu, nx0 != 0u), w1 = __ballot_sync(0xffffffffu, nx1 != 0u);
        if (lane == 0) {
          mws[(it & 1) * 12 + ew] = w0;
          if (ew + 8 < 12) mws[(it & 1) * 12 + ew + 8] = w1;
        }
        const int nitem = item + (int)gridDim.x;
        nx0 = mask_bytes(nitem, ew);
        nx1 = mask_bytes(nitem, ew + 8);
        bar_sync_256();
      }
      const uint32_t* mw_item = mws + (it & 1) * 12;

      for (int tile = 0; tile < ntiles; ++tile, ++nt) {
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
          mbar_wait(bar(bSF + sl), (uint32_t)((n / kRing) & 1));
          tc_fence_after();
          const uint32_t tS = tmem + lanef + (uint32_t)(sl * 128 + g * 64);
          if (MODE == MODE_F && kind != 2) {
            // ---- row maximum over my 64 columns ----
#pragma unroll
            for (int c = 0; c < 2; ++c) {
              uint32_t r[32];
              tmem_ld32(tS + (uint32_t)(c * 32), r);
              tmem_ld_wait();
              if (kind == 0 && c == 1) {  // max-only block: the buffer is free as soon as it is in registers
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(bar(bSE + sl));
              }
              const uint32_t mw = c == 0 ? mw0 : mw1;
              if (warp_active) {
                if (mw == 0xffffffffu) {
                  float m0 = __uint_as_float(r[0]), m1 = __uint_as_float(r[1]), m2 = __uint_as_float(r[2]), m3 = __uint_as_float(r[3]);
#pragma unroll
                  for (int i = 4; i < 32; i += 4) {
                    m0 = fmaxf(m0, __uint_as_float(r[i])); m1 = fmaxf(m1, __uint_as_float(r[i + 1]));
                    m2 = fmaxf(m2, __uint_as_float(r[i + 2])); m3 = fmaxf(m3, __uint_as_float(r[i + 3]));
                  }
                  m_run = fmaxf(m_run, fmaxf(fmaxf(m0, m1), fmaxf(m2, m3)));
                } else if (mw != 0u) {
#pragma unroll
                  for (int i = 0; i < 32; ++i) m_run = fmaxf(m_run, ((mw >> i) & 1u) ? __uint_as_float(r[i]) : -INFINITY);
                }
              }
            }
            if (kind == 0) {
              if (pend && j == 0) {  // previous tile's accumulator: drained under this tile's next score block
                epilogue_fq(pend_nt, pend_item, pend_tile, pend_msc);
                pend = false;
              }
              continue;
            }
            // ---- the two warpgroups exchange their partial maxima ----
            xm[(nt & 1) * 256 + g * 128 + row] = m_run;
            bar_sync_256();
            m_run = fmaxf(m_run, xm[(nt & 1) * 256 + (g ^ 1) * 128 + row]);
            msc = m_run == -INFINITY ? 0.f : m_run * p.scale_log2;
          }
          // ---- exp pass: P (F) / dS (BQ) / P and dS (BKV) for my 64 columns ----
          if (MODE != MODE_F) {
            mbar_wait(bar(bDPF), (uint32_t)(n & 1));
            tc_fence_after();
          }
          const int slot = opn & 1;
          uint32_t sP, sDS = 0;
          if (MODE == MODE_BKV) {
            sP = base + oOP + (uint32_t)g * kBlk;
            sDS = base + oOP + 2 * kBlk + (uint32_t)g * kBlk;
            mbar_wait(bar(bOPE + 0), (uint32_t)(opn & 1) ^ 1u);
            mbar_wait(bar(bOPE + 1), (uint32_t)(opn & 1) ^ 1u);
          } else {
            sP = base + oOP + (uint32_t)(g * 2 + slot) * kBlk;
            mbar_wait(bar(bOPE + g * 2 + slot), (uint32_t)((opn >> 1) & 1) ^ 1u);
          }
#pragma unroll
          for (int c = 0; c < 2; ++c) {
            uint32_t r[32];
            uint32_t pk[16];
            const uint32_t mw = c == 0 ? mw0 : mw1;
            tmem_ld32(tS + (uint32_t)(c * 32), r);
            if (MODE == MODE_F) {
              tmem_ld_wait();
              if (c == 1) {
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(bar(bSE + sl));
              }
              if (warp_active && mw == 0xffffffffu) {
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
            } else {
              uint32_t d[32];
              tmem_ld32(tmem + lanef + colDP + (uint32_t)(g * 64 + c * 32), d);
              tmem_ld_wait();
              if (c == 1) {
                tc_fence_before();
                __syncwarp();
                if (lane == 0) { mbar_arrive(bar(bSE + sl)); mbar_arrive(bar(bDPE)); }
              }
              uint32_t pp[16];
              if (warp_active && mw != 0u) {
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
          }
          fence_async_smem_a();
          __syncwarp();
          if (lane == 0) {
            if (MODE == MODE_BKV) { mbar_arrive(bar(bOPF + 0)); mbar_arrive(bar(bOPF + 1)); }
            else mbar_arrive(bar(bOPF + g * 2 + slot));
          }
          ++opn;
          if (MODE != MODE_BKV && pend && j == 0) {  // previous tile's accumulator: drained under this tile's next score block
            epilogue_fq(pend_nt, pend_item, pend_tile, pend_msc);
            pend = false;
          }
        }
        if (MODE == MODE_F) xl[(nt & 1) * 256 + g * 128 + row] = l;
        if (MODE != MODE_BKV) {
          if (pend) {  // (single-block tiles whose first block was max-only never get here with pend set; safety for L == 1 streams)
            epilogue_fq(pend_nt, pend_item, pend_tile, pend_msc);
          }
          pend = true; pend_nt = nt; pend_item = item; pend_tile = tile; pend_msc = msc;
        } else {
          // dV_t (warpgroup 0) / dK_t (warpgroup 1, x softmax scale): rows = keys of the tile
          mbar_wait(bar(bOF + 0), (uint32_t)(nt & 1));
          tc_fence_after();
          const int key = tile * 128 + row;
          const float mul = g == 0 ? 1.f : p.scale;
          bf16* dst0 = p.dqkv + ((long long)b_ * S + key) * 3LL * H + (g == 0 ? 2 * H : H) + h_ * 64;
#pragma unroll
          for (int c = 0; c < 2; ++c) {
            uint32_t r[32];
            tmem_ld32(tmem + lanef + colAcc + (uint32_t)(g * 64 + c * 32), r);
            tmem_ld_wait();
            if (c == 1) {
              tc_fence_before();
              __syncwarp();
              if (lane == 0) mbar_arrive(bar(bOE + 0));
            }
            if (key < S) {
              uint4* dst = reinterpret_cast<uint4*>(dst0 + c * 32);
#pragma unroll
              for (int i = 0; i < 4; ++i) {
                uint4 v;
                v.x = pack_bf16x2(__uint_as_float(r[8 * i]) * mul, __uint_as_float(r[8 * i + 1]) * mul);
                v.y = pack_bf16x2(__uint_as_float(r[8 * i + 2]) * mul, __uint_as_float(r[8 * i + 3]) * mul);
                v.z = pack_bf16x2(__uint_as_float(r[8 * i + 4]) * mul, __uint_as_float(r[8 * i + 5]) * mul);
                v.w = pack_bf16x2(__uint_as_float(r[8 * i + 6]) * mul, __uint_as_float(r[8 * i + 7]) * mul);
                dst[i] = v;
              }
            }
          }
        }
      }
    }
    if (MODE != MODE_BKV && pend) epilogue_fq(pend_nt, pend_item, pend_tile, pend_msc);
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

bool attn_sm100_ok(int S, float dropout_p) { return g_attn_sm100 != 0 && dropout_p == 0.f && S >= 1 && S <= 384; }
void attn_sm100_enable(int on) { g_attn_sm100 = on; }

int attn_fwd_sm100(const void* qkv, const uint8_t* key_mask, void* ctx, float* lse, int B, int S, int heads, cudaStream_t st) {
  AParams p{};
  fill(p, B, S, heads);
  p.key_mask = key_mask; p.ctx = reinterpret_cast<bf16*>(ctx); p.lse = lse;
  const int H = heads * 64;
  CUtensorMap tm;
  int rc = encode_tmap_2d(&tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, qkv, (uint64_t)3 * H, (uint64_t)B * S, (uint64_t)3 * H, 64, 64);
  if (rc) return rc;
  if ((rc = set_smem_a(attn_sm100_kernel<MODE_F>))) return rc;
  const int grid = p.n_items < device_sm_count() ? p.n_items : device_sm_count();
  launch(attn_sm100_kernel<MODE_F>, dim3(grid), dim3(kThreadsA), (size_t)kSmemA, st, tm, tm, p);
  return check_launch("attn_sm100_kernel<F>");
}

int attn_bwd_sm100(const void* qkv, const uint8_t* key_mask, const void* ctx, const void* dctx, const float* lse, float* delta, void* dqkv, int B, int S,
                   int heads, cudaStream_t st) {
  AParams p{};
  fill(p, B, S, heads);
  p.key_mask = key_mask; p.lse = const_cast<float*>(lse); p.delta = delta; p.dqkv = reinterpret_cast<bf16*>(dqkv);
  const int H = heads * 64;
  CUtensorMap tmQ, tmD;
  int rc = encode_tmap_2d(&tmQ, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, qkv, (uint64_t)3 * H, (uint64_t)B * S, (uint64_t)3 * H, 64, 64);
  if (rc) return rc;
  rc = encode_tmap_2d(&tmD, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, dctx, (uint64_t)H, (uint64_t)B * S, (uint64_t)H, 64, 64);
  if (rc) return rc;
  if ((rc = set_smem_a(attn_sm100_kernel<MODE_BQ>))) return rc;
  if ((rc = set_smem_a(attn_sm100_kernel<MODE_BKV>))) return rc;
  const long long total = (long long)B * S * heads * 8;
  launch(attn_delta_kernel, dim3((unsigned)((total + 255) / 256)), dim3(256), 0, st, reinterpret_cast<const bf16*>(ctx), reinterpret_cast<const bf16*>(dctx), delta,
         B, S, heads);
  if ((rc = check_launch("attn_delta_kernel"))) return rc;
  const int grid = p.n_items < device_sm_count() ? p.n_items : device_sm_count();
  launch(attn_sm100_kernel<MODE_BQ>, dim3(grid), dim3(kThreadsA), (size_t)kSmemA, st, tmQ, tmD, p);
  if ((rc = check_launch("attn_sm100_kernel<BQ>"))) return rc;
  launch(attn_sm100_kernel<MODE_BKV>, dim3(grid), dim3(kThreadsA), (size_t)kSmemA, st, tmQ, tmD, p);
  return check_launch("attn_sm100_kernel<BKV>");
}

}  // namespace vb
