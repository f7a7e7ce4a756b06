// One WaveNet layer (custom_AE_layers.py:305-335) as ONE persistent tcgen05 kernel for sm_100a.
//
//   z   = sum_tap h[r + (tap - 1) d, :] @ W1[tap]                 dilated conv, K = k C, N = 2 C        ("gate tiles")
//   act = tanh(z_t + b + cond) * sigmoid(z_s + b + cond)          conditioning interpolated in the epilogue
//   rs  = act @ [R_res | R_skip W_end]                            1x1, K = C, N = C + c_out              ("res tiles")
//   h' = h + rs[:, :C];  wn_out (+)= rs[:, C:]
//
// The two-launch version (k_wavenet_tc.cu) writes `act` (4 bytes per element) to HBM and reads it back, and reads the
// residual stream twice: 3.5 GB per layer at 64 x 5 s against ~1.5 GB that have to move.  Here a CTA pair (cluster of 2,
// cta_group::2, 256 rows) runs the res tiles of M tile m - 1 between the gate tiles of M tile m:
//
//   accumulator tiles of step j:   G0(m_j) | R0(m_j-1) R1(m_j-1) | G1(m_j) G2(m_j)      (C = 320: 256 | 176 + 176 | 192 + 192)
//
//   ("tc_interleave", default; 0 = G0 G1 G2 | R0 R1 with 224 + 224 + 192).  The tile sequence of a step is a small table
//   (LayerParams::seq) that every role walks; G0 ends on a 64-channel block so that the staging blocks complete in order.
// * `act` of an M tile goes through a per-pair, double-buffered scratch in global memory (2 x 256 rows x 4 cpad bytes per
//   pair, 48 MB for 74 pairs): TMA store from the gate epilogue's staging tiles, TMA load as the A operand of the res
//   tiles.  The skew by one M tile takes the store -> load round trip and the last gate epilogue off the tensor pipe's
//   critical path; with the res tiles right behind G0 of the next M tile the activations (and the residual rows the res
//   epilogue reads back) are still in the L2 when they are needed -- one whole step of all 74 pairs later they are not
//   (profiles/README.md, round 2).  Once the res tiles have read them the scratch rows are dropped from the L2 with
//   discard.global.L2 instead of being written back ("tc_discard").
// * the residual stream ping-pongs between two buffers (the dilated taps of the neighbouring M tiles must keep reading
//   the layer's *input*): old block TMA-loaded from h_in into a staging tile, updated in place by the epilogue warps,
//   TMA-stored to h_out.  Guard rows pass through as the zeros they are.
// * operand feed.  The trace of the first version of this kernel (profiles/r02a_trace.txt, "tc_trace") showed the MMA warp
//   waiting for operands 65 % of the time with the producer waiting for free stages 78 % of the time: the L2 -> SM path
//   delivers ~52 bytes per clock and SM (7.7 KB / clock for the chip, also with the MMAs switched off), and a 128 x 256 x 64
//   block per CTA wants 32 KB per 512 cycles = 62.5.  So the bytes per MMA cycle are what had to shrink:
//     - the three dilated taps of a 64-channel block read ONE slab of pad_l + 128 + pad_r rows (18 KB instead of 3 x 16 KB);
//       tap t is the MMA descriptor view that starts (pad_l + shift_t) rows into the slab -- SWIZZLE_128B keeps its
//       phase through the descriptor's base-offset field ((start >> 7) & 7);
//     - N tiles are cut evenly in 16-channel chunks (C = 320: 224 + 224 + 192 gate columns, 176 + 176 res columns, instead of
//       256 + 256 + 128 and 256 + 96) and every B tile is loaded with a box of exactly its rows: a 128-wide tile cost the
//       tensor pipe as many cycles as a 256-wide one (its smem operand reads bound it) and its box moved 16 KB for 8;
//   W1 rows are packed [16 tanh channels | their 16 sigmoid partners] per chunk (tc_pack.py), so any split in whole chunks
//   keeps the gate local to a tile and a chunk is 32 adjacent accumulator columns.
// * two operand rings: A slabs (3 x 18 KB) and B tiles (5 x 16 KB; 6 x 14 KB without the interleaving), each slot with a full /
//   empty mbarrier pair.  The kernel's speed is the operand bytes in flight over the ~3000-cycle turn-around of a ring slot.
// * the loops of the control warps are the most sensitive code of the kernel: run-time options inside them (polling waits, L2
//   cache-hint forms of the TMA instructions, cluster-of-4 bookkeeping) cost 3 - 5 % each and are compiled out (template
//   parameters QUAD, MODE, GTU) or gone; the epilogue's code size matters too (one copy of the gate / res tile code).
// * the epilogue warps pull ALL their accumulator columns of a tile into registers first and hand the TMEM buffer back
//   before any math (tmem_empty right after tcgen05.wait::ld): the tensor pipe never waits for gate / residual math.
// * staging tiles (2 x 32 KB) are handed around with mbarriers only -- no CTA-wide named barrier in the epilogue: the
//   16 epilogue warps arrive on `stg_ready`, one manager thread (warp 2) issues every TMA store, recycles the buffer
//   (`stg_avail`: plain arrive for a gate block, expect_tx + TMA load of the old residual block for a res block) one block
//   ahead of the epilogue, and signals `act_ready` to the producer once the scratch writes of a step have completed.
//
// Warp roles (640 threads): 0 TMA producer of the B tiles, 1 MMA issuer in the leader CTA / TMA producer of BOTH CTAs' A slabs
// in the other one (it has no MMAs to issue; the leader's slab goes into the leader's shared memory as a multicast to that one
// CTA), 2 TMEM allocator + staging manager, 3 conditioning stager, 4-19 epilogue (TMEM lane quarter = warp % 4, 16-column chunk
// of every 64-column block = (warp - 4) / 4).
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cuda_fp8.h>

#include <string>

#include "kernels.cuh"
#include "tc_common.cuh"
#include "wn_tc.cuh"

namespace mbx {

namespace {

using namespace tcx;

constexpr int TILE_M = 128, TILE_K = 64, ACC_COLS = 256;
constexpr int NA = 3;                                    // A slab ring: slots of (pad_l + 128 + pad_r rows) x 128 bytes
constexpr int SLAB_ROWS_MAX = 160;
constexpr int MAX_NB = 8;                                // B tile ring: as many slots of (widest tile / 2 rows) x 128 bytes as fit
constexpr int NSTG = 2;                                  // staging blocks: [16-bit plane tile 16 KB | lo plane tile 16 KB]
constexpr int STG_BYTES = 2 * TILE_M * 128;
constexpr int COND_ROWS = 15;                            // conditioning rows a 128-row tile can touch at lin_up = 10
// dynamic shared memory: [barriers 512 B | K tables 1.5 KB | A ring | B ring | staging | 2 conditioning stages], offsets in LayerParams
constexpr int OFF_A = 3072;
constexpr int SMEM_LIMIT = 232448;
constexpr int EW = 16, L_THREADS = 128 + 32 * EW;
constexpr int MAX_TILES = 4, MAX_BLK = 24, MAX_LIN = 32;
constexpr int TRACE_SLOTS = 384;                          // tile records per role and CTA
// barrier slots (uint64) at the start of the dynamic shared memory
constexpr int BAR_FULL_A = 0, BAR_EMPTY_A = BAR_FULL_A + NA, BAR_FULL_B = BAR_EMPTY_A + NA, BAR_EMPTY_B = BAR_FULL_B + MAX_NB;
constexpr int BAR_TMEM_FULL = BAR_EMPTY_B + MAX_NB, BAR_TMEM_EMPTY = BAR_TMEM_FULL + 2, BAR_STG_AVAIL = BAR_TMEM_EMPTY + 2;
constexpr int BAR_STG_READY = BAR_STG_AVAIL + NSTG, BAR_COND_FULL = BAR_STG_READY + NSTG, BAR_COND_EMPTY = BAR_COND_FULL + 2;
constexpr int BAR_ACT_READY = BAR_COND_EMPTY + 2, BAR_END = BAR_ACT_READY + 2;
static_assert(BAR_END * 8 + 16 <= 512, "barrier block");

enum : int { KF_NEW_SLAB = 1, KF_LAST_OF_SLAB = 2, KF_F8 = 4, KF_SCALE_D = 8, KF_OVERWRITE = 16 };

// One B tile (64 K) of an accumulator tile's K loop and the A slab view it multiplies.  The tables are built on the host and
// copied to shared memory at kernel start: the producer and MMA warps walk them once per tile, and indexed reads of the
// kernel parameters cost them hundreds of cycles per entry (trace of the first attempt, profiles/README.md).
struct KEnt {
    int a_col;      // KF_NEW_SLAB: first column of the slab in the A tensor
    short a_row;    // KF_NEW_SLAB: first row of the slab relative to the tile's first row (-pad_l, or the tap shift)
    short a_view;   // rows between the slab's first row and the MMA operand's first row (pad_l + shift, or 0)
    int b_col;      // first column (K) of the B tile
    int flags;
};
constexpr int MAX_K1 = 64, MAX_K2 = 24;
constexpr int OFF_KTAB = 512;                            // K tables behind the 512-byte barrier block, in front of the A ring
static_assert(sizeof(KEnt) == 16 && OFF_KTAB + (MAX_K1 + MAX_K2) * 16 + (MAX_K1 + MAX_K2 + 2) * 8 + MAX_K1 <= 3072, "K tables");

struct TileDesc {
    int n0, w;      // first packed column and width (MMA N) of the tile
    int bmap;       // which of the two B tensor maps has a box of w / 2 rows
    int c0, nch;    // first chunk and number of chunks (gate: 16 channels = 32 columns; res: 16 columns)
};

struct alignas(64) LayerParams {
    CUtensorMap tm_hs;       // layer input (rows, 2 cpad), box 64 x slab rows: gate A operand
    CUtensorMap tm_h;        // the same tensor, box 64 x 128: old residual blocks
    CUtensorMap tm_scr;      // act scratch (n_groups * 512, 2 cpad), box 64 x 128: gate stores, res A operand
    CUtensorMap tm_hout;     // layer output (rows, 2 cpad), box 64 x 128
    CUtensorMap tm_w1[2];    // (n1, k1) dilated-conv weights, boxes 64 x (w / 2) for the two tile widths
    CUtensorMap tm_w2[2];    // (n2, k2) res / skip weights
    KEnt k1[MAX_K1];         // K loop of a gate tile
    KEnt k2[MAX_K2];         // K loop of a res tile
    int n_k1, n_k2, n8_1, n8_2;   // entries of a gate / res tile; the first n8 are e4m3 blocks
    int n_slab1, n_slab2;         // A slabs one gate / res tile starts (KEnt::flags >> 8 numbers them)
    TileDesc t1[MAX_TILES], t2[MAX_TILES];
    int n_t1, n_t2;
    int slab_bytes;          // bytes one gate slab load brings (slab rows x 128)
    // shared memory carve-up (bytes from the 1024-aligned base)
    int slab_slot, n_a, off_b, n_b, b_slot, off_stg, off_cond, cond_ld;
    int cluster;             // CTAs per cluster: 2 (one pair) or 4 (two pairs sharing the B tiles by multicast)
    int f16;                 // 16-bit operands are fp16 (MBEXWN_PREC_F16F8), else bf16
    long long rows;
    int tiles_mg;            // 256-row M tiles
    int n_gate_blk, n_res_blk;   // staged 64-channel blocks of a step: gate outputs and residual read-modify-writes
    // Order of a step: gate tiles [0, g_first) of M tile j, the res tiles of M tile j - 1, gate tiles [g_first, n_t1) of M tile j.
    // g_first = n_t1 is "all gate tiles, then the res tiles"; g_first = 1 (the first gate tile ends on a 64-channel block, nb0
    // blocks) lets the res tiles follow the gated activations and the residual rows they read back after ONE tile instead
    // of a whole step, while those are still in the L2.
    int g_first, nb0;
    int n_seq, seq[2 * MAX_TILES];   // the tiles of a step in issue order: t = gate tile t, 16 + t = res tile t
    uint8_t* scr_discard;    // the activation scratch when its dead rows are to be discarded from the L2 (option "tc_discard"), else nullptr
    // gate epilogue
    const float* bias1;
    const float* cond;
    long long cond_total;
    int cond_rows;
    int c, cpad, lin_up, gate, write_lo, steps_per_frame, out_f16f8;
    float act_lo_scale, h_lo_scale, h_lo_inv;
    float lin_w0[MAX_LIN], lin_w1[MAX_LIN];
    // res epilogue
    const float* bias2;
    float* skip;
    int skip_ld, skip_c, res_cols, first;
    FrameGrid grid;
    int* range_flag;         // f16f8 range guard (tc_common.cuh: range_check8), or nullptr
    uint32_t* trace;         // TRACE builds: [cta][role 0..2][TRACE_SLOTS][4]
    int debug;               // TRACE builds only, timing experiments (results are wrong): 1 = no B loads, 2 = no A loads,
                             // 4 = no MMAs are issued, 8 = the epilogue skips its math, 64 = no loads of the old residual blocks,
                             // 128 = no stores (scratch, layer output)
};

__device__ __forceinline__ uint32_t clk32() {
    uint32_t c;
    asm volatile("mov.u32 %0, %%clock;" : "=r"(c));
    return c;
}

// conditioning rows (+ bias) of HALF a gate tile (columns 128 half .. + 127: the chunks the epilogue warps work on at the same
// time) -> smem stage, by the 32 lanes of warp 3.  Stage column j is packed column n0 + 128 half + j: chunk j / 32,
// [16 tanh | 16 sigmoid] channels inside a chunk.
constexpr int COND_COLS = 128, COND_LD = COND_COLS + 4;
constexpr int NCOND = 2;                                 // half-tile conditioning stages (1 would give the B ring 8 KB more, but the stager can then not run ahead: measured slower)
__device__ __forceinline__ void cond_stage_fill(const LayerParams& p, float* buf, int m0, const TileDesc& td, int lane, int half) {
    const int col0 = COND_COLS * half;
    const int w = td.w - col0 < COND_COLS ? td.w - col0 : COND_COLS;     // columns of this half (<= 0: the tile ends before it)
    if (w <= 0) return;
    const int w4 = w >> 2;
    const int rc0 = m0 / p.lin_up;
    const int total = p.cond_rows * w4;
    for (int f0 = lane; f0 < total; f0 += 32 * 4) {
        float4 v[4], b[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int f = f0 + 32 * i;
            const int r = f / w4, j = (f - r * w4) * 4;
            const int jt = col0 + j;                                     // column inside the tile
            const int within = jt & 31;
            const int ch = 16 * (td.c0 + (jt >> 5)) + (within & 15);
            const int src_col = (within >= 16 ? p.c : 0) + ch;
            v[i] = make_float4(0.f, 0.f, 0.f, 0.f);
            b[i] = make_float4(0.f, 0.f, 0.f, 0.f);
            if (f < total && ch < p.c && rc0 + r < p.cond_total) {           // channel padding: C is a multiple of 4
                b[i] = __ldg(reinterpret_cast<const float4*>(p.bias1 + td.n0 + jt));
                v[i] = __ldg(reinterpret_cast<const float4*>(p.cond + (long long)(rc0 + r) * 2 * p.c + src_col));
            }
        }
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int f = f0 + 32 * i;
            const int r = f / w4, j = (f - r * w4) * 4;
            if (f < total)
                *reinterpret_cast<float4*>(buf + r * COND_LD + j) =
                    make_float4(v[i].x + b[i].x, v[i].y + b[i].y, v[i].z + b[i].z, v[i].w + b[i].w);
        }
    }
}

// tanh(t) * sigmoid(s) with three MUFU operations: u = e^-2t, v = e^-s, (1 - u) / ((1 + u) (1 + v)).  u and v are capped so
// that the denominator stays finite (tanh is +-1 to fp32 precision long before).
__device__ __forceinline__ float min_nan(float a, float b) {     // NaN-propagating minimum (fminf would swallow a NaN accumulator)
    float r;
    asm("min.NaN.f32 %0, %1, %2;" : "=f"(r) : "f"(a), "f"(b));
    return r;
}
__device__ __forceinline__ float gate_gtu(float t, float s) {
    const float u = min_nan(ex2_approx(t * -2.885390081777927f), 1e18f);
    const float v = min_nan(ex2_approx(s * -1.4426950408889634f), 1e18f);
    return (1.f - u) * rcp_approx((1.f + u) * (1.f + v));
}

struct EpiState {
    uint8_t* smem;           // 1024-byte aligned base of the dynamic shared memory: barriers at its start
    uint8_t* stg;            // staging tiles
    uint32_t blk;            // staged blocks of this CTA so far
    uint32_t wait_cyc;       // TRACE: cycles spent waiting for staging tiles
};
__device__ __forceinline__ uint64_t* epi_bar(const EpiState& es, int idx) { return reinterpret_cast<uint64_t*>(es.smem) + idx; }

// Eight gate outputs of one row: zt / zs = tanh / sigmoid pre-activations of channels 8 half .. + 7 of a 16-channel chunk,
// cs0 / cs1 = the two conditioning rows (+ bias) of this row in the smem stage at the chunk's first column
// ([16 tanh | 16 sigmoid]).  Result: 16 bytes of the fp16 / bf16 plane and the matching bytes of the lo plane in the staging tile.
// MODE: 0 = F16F8 planes (fp16 + e4m3 lo8 / hi8), 1 = bf16 (hi plane only), 2 = bf16x3 (hi + lo planes); GTU: the gate is
// tanh * sigmoid (compile-time: the other gate types keep the run-time switch)
template <int MODE, bool GTU>
__device__ __forceinline__ void gate_half(const LayerParams& p, const float (&zt)[8], const float (&zs)[8], const float* cs0, const float* cs1,
                                          float w0, float w1, bool valid, uint8_t* t_hi, int sw, int part, int half) {
    float a[8];
    if (valid) {
#pragma unroll
        for (int v4 = 0; v4 < 2; ++v4) {
            const int col = 8 * half + 4 * v4;
            const float4 x0 = *reinterpret_cast<const float4*>(cs0 + col);
            const float4 x1 = *reinterpret_cast<const float4*>(cs1 + col);
            const float4 y0 = *reinterpret_cast<const float4*>(cs0 + 16 + col);
            const float4 y1 = *reinterpret_cast<const float4*>(cs1 + 16 + col);
            const float xa[4] = {x0.x, x0.y, x0.z, x0.w}, xb[4] = {x1.x, x1.y, x1.z, x1.w};
            const float ya[4] = {y0.x, y0.y, y0.z, y0.w}, yb[4] = {y1.x, y1.y, y1.z, y1.w};
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                float t = fmaf(xb[e], w1, fmaf(xa[e], w0, zt[4 * v4 + e]));
                const float sg = fmaf(yb[e], w1, fmaf(ya[e], w0, zs[4 * v4 + e]));
                if (GTU) {
                    a[4 * v4 + e] = gate_gtu(t, sg);
                } else {
                    if (p.gate == GATE_GFU) t = t * rcp_approx(1.f + fabsf(t));
                    else if (p.gate == GATE_GSU) t = t * rcp_approx(1.f + sqrtf(fabsf(t)));
                    a[4 * v4 + e] = t * fast_sigmoid(sg);
                }
            }
        }
    } else {
#pragma unroll
        for (int e = 0; e < 8; ++e) a[e] = 0.f;                  // guard rows stay zero
    }
    // staging tiles, SWIZZLE_128B: 16-byte chunk c of row r sits at r * 128 + ((c ^ (r & 7)) << 4)
    uint8_t* t_lo = t_hi + TILE_M * 128;
    if (MODE == 0) {
        uint4 h16;
        uint2 l8, h8;
        split_f16f8(a, p.act_lo_scale, h16, l8, h8);
        *reinterpret_cast<uint4*>(t_hi + (((2 * part + half) ^ sw) << 4)) = h16;
        *reinterpret_cast<uint2*>(t_lo + ((part ^ sw) << 4) + 8 * half) = l8;                 // lo8: bytes 0 .. 63 of the row
        *reinterpret_cast<uint2*>(t_lo + (((4 + part) ^ sw) << 4) + 8 * half) = h8;           // hi8: bytes 64 .. 127
    } else {
        uint32_t hw4[4], lw4[4];
#pragma unroll
        for (int e = 0; e < 8; e += 2) {
            __nv_bfloat16 h0, l0, h1, l1;
            split_bf16(a[e], h0, l0);
            split_bf16(a[e + 1], h1, l1);
            hw4[e / 2] = pack2(h0, h1);
            lw4[e / 2] = pack2(l0, l1);
        }
        *reinterpret_cast<uint4*>(t_hi + (((2 * part + half) ^ sw) << 4)) = make_uint4(hw4[0], hw4[1], hw4[2], hw4[3]);
        if (MODE == 2) *reinterpret_cast<uint4*>(t_lo + (((2 * part + half) ^ sw) << 4)) = make_uint4(lw4[0], lw4[1], lw4[2], lw4[3]);
    }
}

// One residual chunk: v = 16 accumulator values of one row, residual channels n .. n + 15; the old values sit in the staging
// tile the manager loaded (updated in place, guard rows untouched).
template <bool TRACE, int MODE>
__device__ __forceinline__ void res_chunk(const LayerParams& p, EpiState& es, const float (&v)[16], int n, bool valid, int r, int part,
                                          int lane) {
    const uint32_t buf = es.blk % NSTG;
    {
        const uint32_t t0 = TRACE ? clk32() : 0u;
        mbar_wait(epi_bar(es, BAR_STG_AVAIL + buf), (es.blk / NSTG) & 1);
        if (TRACE) es.wait_cyc += clk32() - t0;
    }
    if (valid) {
        const int sw = r & 7;
        uint8_t* t_hi = es.stg + buf * STG_BYTES + r * 128;
        uint8_t* t_lo = t_hi + TILE_M * 128;
#pragma unroll
        for (int i = 0; i < 16; i += 8) {
            const float4 b0 = __ldg(reinterpret_cast<const float4*>(p.bias2 + n + i)), b1 = __ldg(reinterpret_cast<const float4*>(p.bias2 + n + i) + 1);
            const float bv[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
            uint4* ph16 = reinterpret_cast<uint4*>(t_hi + (((2 * part + (i >> 3)) ^ sw) << 4));
            float prev[8], o[8];
            if (MODE == 0) {
                uint2* pl8 = reinterpret_cast<uint2*>(t_lo + ((part ^ sw) << 4) + i);
                join_f16f8(*ph16, *pl8, p.h_lo_inv, prev);
#pragma unroll
                for (int e = 0; e < 8; ++e) o[e] = (n + i + e < p.c) ? prev[e] + (v[i + e] + bv[e]) : 0.f;
                range_check8(o, p.range_flag);
                uint4 h16;
                uint2 l8, h8;
                split_f16f8(o, p.h_lo_scale, h16, l8, h8);
                *ph16 = h16;
                *pl8 = l8;
                *reinterpret_cast<uint2*>(t_lo + (((4 + part) ^ sw) << 4) + i) = h8;
            } else {
                uint4* plo = reinterpret_cast<uint4*>(t_lo + (((2 * part + (i >> 3)) ^ sw) << 4));
                const uint4 oh = *ph16, ol = *plo;
                const uint32_t hw[4] = {oh.x, oh.y, oh.z, oh.w}, lw[4] = {ol.x, ol.y, ol.z, ol.w};
                uint32_t nh[4], nl[4];
#pragma unroll
                for (int w = 0; w < 4; ++w) {
                    float x[2];
#pragma unroll
                    for (int e = 0; e < 2; ++e) {
                        const int idx = 2 * w + e;
                        const float pv = __bfloat162float(__ushort_as_bfloat16((unsigned short)(hw[w] >> (16 * e)))) +
                                         __bfloat162float(__ushort_as_bfloat16((unsigned short)(lw[w] >> (16 * e))));
                        x[e] = (n + i + idx < p.c) ? pv + (v[i + idx] + bv[idx]) : 0.f;
                    }
                    __nv_bfloat16 h0, l0, h1, l1;
                    split_bf16(x[0], h0, l0);
                    split_bf16(x[1], h1, l1);
                    nh[w] = pack2(h0, h1);
                    nl[w] = pack2(l0, l1);
                }
                *ph16 = make_uint4(nh[0], nh[1], nh[2], nh[3]);
                *plo = make_uint4(nl[0], nl[1], nl[2], nl[3]);
            }
        }
    }
    fence_proxy_async();
    __syncwarp();
    if (lane == 0) mbar_arrive(epi_bar(es, BAR_STG_READY + buf));
    ++es.blk;
}

// 16 WaveNet-output columns of one row: accumulated in global memory (fp32)
__device__ __forceinline__ void skip_chunk(const LayerParams& p, const float (&v)[16], int n, long long row, bool valid) {
    const int sc = n - p.res_cols;
    if (!valid || sc >= p.skip_c) return;
    float* dst = p.skip + row * p.skip_ld + sc;
#pragma unroll
    for (int i = 0; i < 16; i += 4) {
        const float4 b4 = __ldg(reinterpret_cast<const float4*>(p.bias2 + n + i));
        float4 nv = make_float4(v[i] + b4.x, v[i + 1] + b4.y, v[i + 2] + b4.z, v[i + 3] + b4.w);
        if (!p.first) {
            const float4 ov = *reinterpret_cast<const float4*>(dst + i);
            nv.x += ov.x; nv.y += ov.y; nv.z += ov.z; nv.w += ov.w;
        }
        *reinterpret_cast<float4*>(dst + i) = nv;
    }
}

template <bool TRACE, bool QUAD, int MODE, bool GTU>
__global__ void __launch_bounds__(L_THREADS, 1) wn_layer_kernel(const __grid_constant__ LayerParams p) {
    // K-major SWIZZLE_128B smem matrix descriptor without the address field (see k_wavenet_tc.cu)
    constexpr uint64_t DESC_HI = ((uint64_t)1 << 16) | ((uint64_t)(1024 >> 4) << 32) | ((uint64_t)1 << 46) | ((uint64_t)2 << 61);
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem);
    uint64_t* full_a = bars + BAR_FULL_A;
    uint64_t* empty_a = bars + BAR_EMPTY_A;
    uint64_t* full_b = bars + BAR_FULL_B;
    uint64_t* empty_b = bars + BAR_EMPTY_B;
    uint64_t* tmem_full = bars + BAR_TMEM_FULL;
    uint64_t* tmem_empty = bars + BAR_TMEM_EMPTY;
    uint64_t* stg_avail = bars + BAR_STG_AVAIL;
    uint64_t* stg_ready = bars + BAR_STG_READY;
    uint64_t* cond_full = bars + BAR_COND_FULL;
    uint64_t* cond_empty = bars + BAR_COND_EMPTY;
    uint64_t* act_ready = bars + BAR_ACT_READY;
    uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bars + BAR_END);
    uint8_t* ring_a = smem + OFF_A;
    uint8_t* ring_b = smem + p.off_b;
    uint8_t* stg = smem + p.off_stg;
    float* cond_stage = reinterpret_cast<float*>(smem + p.off_cond);
    constexpr int cond_buf = COND_ROWS * COND_LD;                  // floats per conditioning stage (half a gate tile)
    KEnt* k1 = reinterpret_cast<KEnt*>(smem + OFF_KTAB);
    KEnt* k2 = k1 + MAX_K1;
    // the MMA warp's view of the same loops: {A view offset in 16-byte units, flags} (one spare entry behind each table)
    int2* km1 = reinterpret_cast<int2*>(smem + OFF_KTAB + (MAX_K1 + MAX_K2) * 16);
    int2* km2 = km1 + MAX_K1 + 1;
    // slab_ent[k] = entry of a gate tile that starts its A slab k; slab_ent[MAX_K1 / 2 + k] the same for a res tile
    uint8_t* slab_ent = reinterpret_cast<uint8_t*>(km2 + MAX_K2 + 1);
    for (int i = threadIdx.x; i < MAX_K1 + MAX_K2; i += L_THREADS) {
        const KEnt& e = i < MAX_K1 ? p.k1[i] : p.k2[i - MAX_K1];
        (i < MAX_K1 ? km1[i] : km2[i - MAX_K1]) = make_int2((int)e.a_view * 8, e.flags & 0xff);
        const bool live = i < MAX_K1 ? i < p.n_k1 : i - MAX_K1 < p.n_k2;
        if (live && (e.flags & KF_NEW_SLAB)) slab_ent[(i < MAX_K1 ? 0 : MAX_K1 / 2) + (e.flags >> 8)] = (uint8_t)(i < MAX_K1 ? i : i - MAX_K1);
    }
    for (int i = threadIdx.x; i < (MAX_K1 + MAX_K2) * 4; i += L_THREADS) {
        const int e = i >> 2, w = i & 3;
        const int* src = e < MAX_K1 ? reinterpret_cast<const int*>(&p.k1[e]) : reinterpret_cast<const int*>(&p.k2[e - MAX_K1]);
        reinterpret_cast<int*>(k1)[i] = src[w];
    }

    const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);
    const int lane = threadIdx.x & 31;
    // Cluster of 2 (one CTA pair) or 4 (two pairs that walk the same B tile sequence and share every B load by multicast:
    // the weights are re-streamed from the L2 for every M tile, and at one pair per cluster the L2 -> SM traffic of the 148
    // CTAs, ~39 B/clk/SM, is what bounds the kernel -- the L2 delivers ~42 B/clk/SM).
    const int crank = (int)cluster_ctarank();
    const int rank = crank & 1;                                     // rank inside the CTA pair
    const int pair_in_cluster = crank >> 1;
    const uint32_t lead_rank = (uint32_t)(crank & ~1);              // cluster rank of this pair's leader CTA
    const bool leader = rank == 0;
    constexpr bool quad = QUAD;                                     // clusters of 4 are a separate instantiation: none of it in the pair kernel
    const int group = blockIdx.x >> 1, n_groups = gridDim.x >> 1;
    // M tiles of this pair; the pairs of a cluster run the same number of steps (the second one may end on a tile past the
    // last row: loads are zero-filled, stores clipped, the epilogue's row checks fail)
    const int group0 = quad ? (group & ~1) : group;
    const int n_j = group0 < p.tiles_mg ? (p.tiles_mg - group0 + n_groups - 1) / n_groups : 0;
    // rows of this CTA in M tile j of the pair / in the act scratch
    auto m0_of = [&](int j) { return ((group + n_groups * j) * 2 + rank) * TILE_M; };
    auto scr_of = [&](int j) { return ((group * 2 + (j & 1)) * 2 + rank) * TILE_M; };

    if (warp == 0 && elect_one()) {
        asm volatile("prefetch.tensormap [%0];" ::"l"((uint64_t)&p.tm_hs) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"((uint64_t)&p.tm_h) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"((uint64_t)&p.tm_w1[0]) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"((uint64_t)&p.tm_w1[1]) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"((uint64_t)&p.tm_w2[0]) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"((uint64_t)&p.tm_w2[1]) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"((uint64_t)&p.tm_scr) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"((uint64_t)&p.tm_hout) : "memory");
    }
    if (warp == 1 && elect_one()) {
        for (int s = 0; s < NA; ++s) { mbar_init(&full_a[s], 1); mbar_init(&empty_a[s], 1); }
        for (int s = 0; s < p.n_b; ++s) { mbar_init(&full_b[s], 1); mbar_init(&empty_b[s], quad ? 2 : 1); }
        for (int s = 0; s < 2; ++s) {
            mbar_init(&tmem_full[s], 1);
            mbar_init(&tmem_empty[s], 2 * EW);
            mbar_init(&cond_full[s], 32);
            mbar_init(&cond_empty[s], EW);
            mbar_init(&act_ready[s], 2);                        // both CTAs' managers (the A producer of the pair runs in the non-leader CTA)
        }
        for (int s = 0; s < NSTG; ++s) { mbar_init(&stg_avail[s], 1); mbar_init(&stg_ready[s], EW); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 2) {
        asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_ptr)), "r"(512) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    cluster_sync_all();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_ptr;
    uint32_t* trc = TRACE ? p.trace + (size_t)blockIdx.x * 3 * TRACE_SLOTS * 4 : nullptr;
    const int dbg = TRACE ? p.debug : 0;

    if (warp == 0) {
        asm volatile("setmaxnreg.dec.sync.aligned.u32 64;");
        // ===== TMA producer of the B tiles (weights): the whole warp walks the K loops, one elected lane issues.  One
        // cp.async.bulk.tensor costs its issuing thread ~270 cycles whatever the box size; a K block of a res tile (352 MMA cycles)
        // needs an A tile and a B tile, so the A operand has a producer of its own: warp 1 of the NON-leader CTA, which has no
        // MMAs to issue, loads the A slabs / tiles of both CTAs of the pair =====
        // In a cluster of 4 the two pairs take turns with the B tiles: the CTAs of pair (entry & 1) load their half of the
        // tile and multicast it to the CTA of the same rank in the other pair.  A slot is free when BOTH pairs have consumed
        // it (empty_b counts two commits, each multicast to the four CTAs); every leader arms its own full barrier, whoever
        // loads -- the bytes of the other pair's load may be counted before that (the tx count goes negative meanwhile).
        uint32_t sb = 0, phb = 0, tile_it = 0, ent = 0;            // B ring slot and phase parity, B tiles so far
        const uint32_t lfb0 = map_to_cta(smem_u32(&full_b[0]), lead_rank);
        const uint32_t fb0_mc = smem_u32(&full_b[0]) & 0xFEFFFFFFu;  // multicast form: this offset in the even CTA of each destination's pair
        const uint16_t mc_mask = (uint16_t)(0x5u << rank);
        uint32_t wait_cyc = 0;
        auto load_b = [&](const CUtensorMap* tmb, uint32_t b_bytes, int col, int row) {
            const bool mine = !quad || (int)(ent & 1u) == pair_in_cluster;
            ++ent;
            if (mine || leader) {
                const uint32_t t0 = TRACE ? clk32() : 0u;
                mbar_wait(&empty_b[sb], phb ^ 1);
                if (TRACE) wait_cyc += clk32() - t0;
                if (elect_one()) {
                    if (dbg & 1) {
                        if (leader) mbar_arrive(&full_b[sb]);
                    } else {
                        if (leader) mbar_expect_tx(&full_b[sb], 2 * b_bytes);
                        if (!quad) tma_load_2d_2sm(tmb, lfb0 + sb * 8, ring_b + sb * p.b_slot, col, row);
                        else if (mine) tma_load_2d_2sm_mc(tmb, fb0_mc + sb * 8, ring_b + sb * p.b_slot, col, row, mc_mask, L2_EVICT_NORMAL);
                    }
                }
                __syncwarp();
            }
            if (++sb == (uint32_t)p.n_b) { sb = 0; phb ^= 1; }
        };
        auto load_tile = [&](const CUtensorMap* tmb, uint32_t b_bytes, const KEnt* ke, int n, int b_row) {
            wait_cyc = 0;
            int b_col = ke[0].b_col;
            for (int i = 0; i < n; ++i) {
                const int col = b_col;
                if (i + 1 < n) b_col = ke[i + 1].b_col;
                load_b(tmb, b_bytes, col, b_row);
            }
            if (TRACE && lane == 0 && tile_it < TRACE_SLOTS) {
                uint32_t* t = trc + (0 * TRACE_SLOTS + tile_it) * 4;
                t[0] = clk32(); t[1] = wait_cyc; t[2] = (uint32_t)n; t[3] = 0;
            }
            ++tile_it;
        };
        for (int j = 0; j <= n_j; ++j)
            for (int e = 0; e < p.n_seq; ++e) {
                const int code = p.seq[e], t = code & 15;
                if (code < 16) {
                    if (j < n_j) load_tile(&p.tm_w1[p.t1[t].bmap], (uint32_t)(p.t1[t].w >> 1) * 128u, k1, p.n_k1, p.t1[t].n0 + rank * (p.t1[t].w >> 1));
                } else if (j > 0) {
                    load_tile(&p.tm_w2[p.t2[t].bmap], (uint32_t)(p.t2[t].w >> 1) * 128u, k2, p.n_k2, p.t2[t].n0 + rank * (p.t2[t].w >> 1));
                }
            }
    } else if (warp == 1) {
        asm volatile("setmaxnreg.dec.sync.aligned.u32 64;");
        // ===== MMA issuer (leader CTA): the whole warp walks the K loops, one elected lane issues =====
        if (leader) {
            uint32_t sa = 0, pha = 0, sb = 0, phb = 0, tile_it = 0;
            const uint32_t a_base = smem_u32(ring_a), b_base = smem_u32(ring_b);
            // running descriptor words of the current ring slots (bits 0..13 = shared address >> 4; the upper descriptor word is constant)
            const uint32_t a_slot16 = (uint32_t)p.slab_slot >> 4, b_slot16 = (uint32_t)p.b_slot >> 4;
            const uint32_t a_desc0 = a_base >> 4, b_desc0 = b_base >> 4;
            uint32_t a_desc = a_desc0, b_desc = b_desc0;            // descriptor word of slot sa / sb
            uint32_t cur_a_desc = a_desc0, cur_a = 0, wait_a = 0, wait_b = 0;   // TRACE: cycles waited for A slabs / B tiles
            const uint16_t mask_pair = (uint16_t)(3u << lead_rank), mask_b = quad ? (uint16_t)0xF : mask_pair;
            // one K block: 4 MMAs on (A view, B slot sb), then the slot(s) go back to the producer
            auto block = [&](const uint32_t tacc, const uint32_t idesc, const int2 e, const int mode, const bool last) {
                // e.x = A view offset in 16-byte units, e.y = flags; mode 0: e4m3, 1: 16-bit
                const int flags = e.y;
                if (flags & KF_NEW_SLAB) {
                    const uint32_t w0 = TRACE ? clk32() : 0u;
                    mbar_wait(&full_a[sa], pha);
                    if (TRACE) wait_a += clk32() - w0;
                    cur_a = sa;
                    cur_a_desc = a_desc;
                    a_desc += a_slot16;
                    if (++sa == (uint32_t)p.n_a) { sa = 0; pha ^= 1; a_desc = a_desc0; }
                }
                const uint32_t w1 = TRACE ? clk32() : 0u;
                mbar_wait(&full_b[sb], phb);
                if (TRACE) wait_b += clk32() - w1;
                tc_fence_after();
                if (elect_one()) {
                    const uint64_t da = DESC_HI | (uint64_t)(cur_a_desc + (uint32_t)e.x);
                    const uint64_t db = DESC_HI | (uint64_t)b_desc;
                    if (!(dbg & 4)) {
                        if (mode == 0) {
#pragma unroll
                            for (int k = 0; k < 4; ++k) tc_mma_f8_2sm(tacc, da + 2 * k, db + 2 * k, idesc, !((flags & KF_OVERWRITE) && k == 0));
                        } else {
                            if (flags & KF_SCALE_D) tc_mma_f16_sd_2sm(tacc, da, db, idesc);      // rescales the e4m3 products by 2^-15
                            else tc_mma_bf16_2sm(tacc, da, db, idesc, (flags & KF_OVERWRITE) ? 0u : 1u);
#pragma unroll
                            for (int k = 1; k < 4; ++k) tc_mma_bf16_2sm(tacc, da + 2 * k, db + 2 * k, idesc, 1u);
                        }
                    }
                    if (quad) {
                        tc_commit_mc(&empty_b[sb], mask_b);
                        if (flags & KF_LAST_OF_SLAB) tc_commit_mc(&empty_a[cur_a], mask_pair);
                        if (last) tc_commit_mc(&tmem_full[(tile_it & 1)], mask_pair);
                    } else {
                        tc_commit_2sm(&empty_b[sb]);
                        if (flags & KF_LAST_OF_SLAB) tc_commit_2sm(&empty_a[cur_a]);
                        if (last) tc_commit_2sm(&tmem_full[(tile_it & 1)]);
                    }
                }
                __syncwarp();
                b_desc += b_slot16;
                if (++sb == (uint32_t)p.n_b) { sb = 0; phb ^= 1; b_desc = b_desc0; }
            };
            auto mma_tile = [&](int width, const int2* km, int n8, int n) {
                const uint32_t idesc = MODE == 0 ? make_idesc_fmt0(2 * TILE_M, width) : make_idesc(2 * TILE_M, width);
                const uint32_t as = tile_it & 1, aph = (tile_it >> 1) & 1;
                const uint32_t t0 = TRACE ? clk32() : 0u;
                mbar_wait(&tmem_empty[as], aph ^ 1);
                tc_fence_after();
                const uint32_t t1 = TRACE ? clk32() : 0u;
                const uint32_t tacc = tmem_base + as * ACC_COLS;
                int2 e = km[0];
                int i = 0;
                for (; i < n8; ++i) {                               // e4m3 correction blocks
                    const int2 cur = e;
                    e = km[i + 1];                                  // the table has one spare entry behind the last
                    block(tacc, idesc, cur, 0, false);
                }
                for (; i < n - 1; ++i) {                            // 16-bit blocks
                    const int2 cur = e;
                    e = km[i + 1];
                    block(tacc, idesc, cur, 1, false);
                }
                block(tacc, idesc, e, 1, true);
                if (TRACE && lane == 0 && tile_it < TRACE_SLOTS) {
                    uint32_t* t = trc + (1 * TRACE_SLOTS + tile_it) * 4;
                    t[0] = t0; t[1] = t1; t[2] = clk32(); t[3] = ((wait_a >> 3) << 16) | ((wait_b >> 3) & 0xffffu);
                }
                wait_a = wait_b = 0;
                ++tile_it;
            };
            for (int j = 0; j <= n_j; ++j)
                for (int e = 0; e < p.n_seq; ++e) {
                    const int code = p.seq[e], t = code & 15;
                    const bool gate = code < 16;
                    if (gate ? j < n_j : j > 0) mma_tile(gate ? p.t1[t].w : p.t2[t].w, gate ? km1 : km2, gate ? p.n8_1 : p.n8_2, gate ? p.n_k1 : p.n_k2);
                }
        } else {
            // ===== A producer of the pair (non-leader CTA): lane 0 loads this CTA's slab, lane 1 the leader's -- into the leader's
            // shared memory, as a multicast to that one CTA -- both reporting to the leader's full barrier.  This CTA's empty
            // barriers see the leader's commits (multicast to the pair), so one wait covers both slots. =====
            uint32_t sa = 0, pha = 0;
            const uint32_t lfa0 = map_to_cta(smem_u32(&full_a[0]), lead_rank);
            const uint32_t fa0_mc = smem_u32(&full_a[0]) & 0xFEFFFFFFu;
            const uint16_t lead_mask = (uint16_t)(1u << lead_rank);
            auto load_slabs = [&](const CUtensorMap* tma, uint32_t a_bytes, const KEnt* ke, const uint8_t* first, int n_slabs, int a_row0) {
                for (int k = 0; k < n_slabs; ++k) {
                    const int4 ent = *reinterpret_cast<const int4*>(ke + first[k]);   // {a_col, a_row | a_view << 16, b_col, flags}
                    const int row = a_row0 + (int)(short)(ent.y & 0xffff);
                    mbar_wait(&empty_a[sa], pha ^ 1);
                    uint8_t* dst = ring_a + sa * p.slab_slot;
                    if (dbg & 2) {
                        if (lane == 0) { mbar_expect_tx_cluster(lfa0 + sa * 8, 0); }
                    } else if (lane == 0) {
                        mbar_expect_tx_cluster(lfa0 + sa * 8, 2 * a_bytes);
                        tma_load_2d_2sm(tma, lfa0 + sa * 8, dst, ent.x, row);
                    } else if (lane == 1) {
                        tma_load_2d_2sm_mc(tma, fa0_mc + sa * 8, dst, ent.x, row - TILE_M, lead_mask, L2_EVICT_NORMAL);
                    }
                    __syncwarp();
                    if (++sa == (uint32_t)p.n_a) { sa = 0; pha ^= 1; }
                }
            };
            for (int j = 0; j <= n_j; ++j) {
                bool act_seen = false;
                for (int e = 0; e < p.n_seq; ++e) {
                    const int code = p.seq[e];
                    if (code < 16) {
                        if (j < n_j) load_slabs(&p.tm_hs, (uint32_t)p.slab_bytes, k1, slab_ent, p.n_slab1, m0_of(j));
                    } else if (j > 0) {
                        if (!act_seen) {
                            // the act of M tile j - 1 must have landed in the scratch: both CTAs' managers have completed their
                            // TMA stores and arrived here
                            mbar_wait(&act_ready[(j - 1) & 1], ((j - 1) >> 1) & 1);
                            asm volatile("fence.acq_rel.cluster;" ::: "memory");
                            asm volatile("fence.proxy.async;" ::: "memory");
                            act_seen = true;
                        }
                        load_slabs(&p.tm_scr, (uint32_t)(TILE_M * 128), k2, slab_ent + MAX_K1 / 2, p.n_slab2, scr_of(j - 1));
                    }
                }
            }
        }
    } else if (warp == 2) {
        asm volatile("setmaxnreg.dec.sync.aligned.u32 64;");
        // ===== staging manager: one thread recycles the staging tiles ahead of the epilogue and issues every store =====
        if (lane == 0) {
            const int n_blk = p.n_gate_blk + p.n_res_blk;
            // `act_ready` lives in the non-leader CTA (the pair's A producer): the leader's manager arrives there remotely
            const uint32_t act_remote = map_to_cta(smem_u32(&act_ready[0]), lead_rank + 1);
            auto act_arrive = [&](int b) {
                if (leader) mbar_arrive_cluster_release(act_remote + b * 8);
                else mbar_arrive(&act_ready[b]);
            };
            int pj = 0, pe = -1, sj = 0, se = -1;                   // prepare / store cursors over (step, block-of-step)
            // position e of a step: gate blocks 0 .. nb0 - 1, the res blocks, the other gate blocks
            auto is_gate = [&](int e) { return e < p.nb0 || e >= p.nb0 + p.n_res_blk; };
            auto gate_idx = [&](int e) { return e < p.nb0 ? e : e - p.n_res_blk; };
            auto advance = [&](int& j, int& e) {
                for (;;) {
                    if (++e >= n_blk) { e = 0; ++j; }
                    if (j > n_j) return false;
                    if (is_gate(e) ? j < n_j : j > 0) return true;
                }
            };
            uint32_t pb = 0, sb = 0;                                // blocks prepared / stored
            int pending_act = -1;                                   // step whose gate stores have been issued but not completed
            bool more_p = advance(pj, pe), more_s = advance(sj, se);
            while (more_s) {
                // prepare as far ahead as the buffers allow: block pb reuses the tile of block pb - NSTG, whose store has been
                // issued (pb - NSTG < sb) and must have finished reading the tile
                while (more_p && pb < sb + NSTG) {
                    tma_store_wait_read();
                    const uint32_t buf = pb % NSTG;
                    if (is_gate(pe)) {
                        mbar_arrive(&stg_avail[buf]);
                    } else {
                        const int col = 64 * (pe - p.nb0), row0 = m0_of(pj - 1);
                        if (dbg & 64) { mbar_arrive(&stg_avail[buf]); ++pb; more_p = advance(pj, pe); continue; }
                        mbar_expect_tx(&stg_avail[buf], STG_BYTES);
                        tma_load_2d(&p.tm_h, &stg_avail[buf], stg + buf * STG_BYTES, col, row0);
                        tma_load_2d(&p.tm_h, &stg_avail[buf], stg + buf * STG_BYTES + TILE_M * 128, p.cpad + col, row0);
                    }
                    ++pb;
                    more_p = advance(pj, pe);
                }
                if (pending_act >= 0) {
                    // scratch writes of a finished step: complete them (not just the smem reads) and tell the producer
                    tma_store_wait_all();
                    asm volatile("fence.proxy.async;" ::: "memory");
                    act_arrive(pending_act & 1);
                    pending_act = -1;
                }
                const uint32_t buf = sb % NSTG;
                mbar_wait(&stg_ready[buf], (sb / NSTG) & 1);
                const uint8_t* t_hi = stg + buf * STG_BYTES;
                const bool last_gate = is_gate(se) && gate_idx(se) == p.n_gate_blk - 1;
                if (dbg & 128) {
                    if (last_gate) pending_act = sj;
                } else if (is_gate(se)) {
                    const int col = 64 * gate_idx(se), row0 = scr_of(sj);
                    tma_store_2d(&p.tm_scr, t_hi, col, row0);
                    if (MODE != 1) tma_store_2d(&p.tm_scr, t_hi + TILE_M * 128, p.cpad + col, row0);
                    if (last_gate) pending_act = sj;
                } else {
                    const int col = 64 * (se - p.nb0), row0 = m0_of(sj - 1);
                    tma_store_2d(&p.tm_hout, t_hi, col, row0);
                    tma_store_2d(&p.tm_hout, t_hi + TILE_M * 128, p.cpad + col, row0);
                }
                tma_store_commit();
                ++sb;
                more_s = advance(sj, se);
            }
            tma_store_wait_all();                                   // staging tiles and global writes outlive the loop
            if (pending_act >= 0) {
                asm volatile("fence.proxy.async;" ::: "memory");
                act_arrive(pending_act & 1);
            }
        }
    } else if (warp == 3) {
        asm volatile("setmaxnreg.dec.sync.aligned.u32 64;");
        // ===== conditioning stager: half a gate tile (the chunks the epilogue warps work on at the same time) ahead of them;
        // two half-tile stages instead of two tile stages give the B ring one more tile =====
        uint32_t hs = 0;
        for (int j = 0; j < n_j; ++j)
            for (int t = 0; t < p.n_t1; ++t)
                for (int half = 0; half < 2; ++half, ++hs) {
                    const uint32_t b = hs % NCOND;
                    mbar_wait(&cond_empty[b], ((hs / NCOND) & 1) ^ 1);
                    cond_stage_fill(p, cond_stage + b * cond_buf, m0_of(j), p.t1[t], lane, half);
                    mbar_arrive(&cond_full[b]);                     // every lane: its own writes are released
                }
    } else {
        asm volatile("setmaxnreg.inc.sync.aligned.u32 104;");
        // ===== epilogue warps =====
        const int q4 = warp & 3, part = (warp - 4) >> 2;
        EpiState es{smem, stg, 0u, 0u};
        uint32_t tile_it = 0, gt = 0;
        const uint32_t lempty0 = map_to_cta(smem_u32(&tmem_empty[0]), lead_rank);
        const int rl = q4 * 32 + lane;                              // row of this thread inside the CTA's 128 rows
        auto wait_tile = [&](uint32_t& t0, uint32_t& t1) -> uint32_t {
            const uint32_t as = tile_it & 1, aph = (tile_it >> 1) & 1;
            t0 = TRACE ? clk32() : 0u;
            mbar_wait(&tmem_full[as], aph);
            tc_fence_after();
            t1 = TRACE ? clk32() : 0u;
            return tmem_base + ((uint32_t)(q4 * 32) << 16) + as * ACC_COLS;
        };
        // the accumulator values of this warp are in registers: hand the TMEM buffer back to the MMA warp
        auto release_tile = [&]() {
            tmem_ld_wait();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) {
                const uint32_t as = tile_it & 1;
                if (leader) mbar_arrive(&tmem_empty[as]);
                else mbar_arrive_cluster(lempty0 + as * 8);
            }
        };
        auto end_tile = [&](uint32_t t0, uint32_t t1, int kind) {
            if (TRACE && warp == 4 && lane == 0 && tile_it < TRACE_SLOTS) {
                uint32_t* t = trc + (2 * TRACE_SLOTS + tile_it) * 4;
                t[0] = t0; t[1] = t1; t[2] = clk32(); t[3] = es.wait_cyc | ((uint32_t)kind << 31);
                es.wait_cyc = 0;
            }
            ++tile_it;
        };
        // per-step row state of the gate tiles (recomputed when the step changes) ...
        int g_step = -1;
        bool in_range = false, valid = false;
        int rl0 = 0, rl1 = 0;
        float w0 = 1.f, w1 = 0.f;
        auto gate_tile = [&](int j, int t) {
            if (g_step != j) {
                g_step = j;
                const int m0 = m0_of(j);
                const int row = m0 + rl;
                in_range = row < (int)p.rows && !(dbg & 8);
                valid = false;
                rl0 = rl1 = 0;
                w0 = 1.f; w1 = 0.f;
                if (in_range) {
                    const int u = p.grid.frame_utt[row / p.steps_per_frame];
                    if (u >= 0) {
                        valid = true;
                        const int hic = p.grid.utt_end[u] * p.steps_per_frame / p.lin_up;
                        const int rc0 = m0 / p.lin_up, rc = row / p.lin_up;
                        const int un = row - rc * p.lin_up;
                        const int rn = rc + 1 < hic ? rc + 1 : hic - 1;
                        rl0 = rc - rc0;
                        rl1 = rn - rc0;
                        w0 = p.lin_w0[un];
                        w1 = p.lin_w1[un];
                    }
                }
            }
            uint32_t t0, t1;
            const TileDesc td = p.t1[t];
            const uint32_t tacc = wait_tile(t0, t1);
            // chunks c0 + k of this tile with (c0 + k) % 4 == part: at most two (a tile has at most eight); eight channels
            // (8 tanh + 8 sigmoid accumulators) at a time.  The TMEM buffer goes back to the MMA warp when the last eight
            // are in registers.
            const int k0 = (part - td.c0) & 3;
            const int nmy = k0 < td.nch ? (k0 + 4 < td.nch ? 2 : 1) : 0;
            if (nmy == 0) release_tile();
#pragma unroll 1
            for (int k = 0; k < 2; ++k) {
                // half-tile stage 2 gt + k: every warp waits for it and hands it back, with or without a chunk in this half
                // (a warp that ran ahead could otherwise arrive twice in one phase of `cond_empty`)
                const uint32_t hs = 2 * gt + k;
                mbar_wait(&cond_full[hs % NCOND], (hs / NCOND) & 1);
                if (k >= nmy) {
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&cond_empty[hs % NCOND]);
                    continue;
                }
                const float* cbuf = cond_stage + (hs % NCOND) * cond_buf - COND_COLS * k;     // indexed by the column inside the tile
                const int cl = 32 * (k0 + 4 * k);
                // chunks behind the tile's MMA width are channel padding that is not computed at all (C = 340: 22 of the 24 chunks
                // of cpad = 384): their warps write the zeros the staging block needs and nothing else
                const bool live = valid && cl < td.w;
                const uint32_t buf = es.blk % NSTG;
                float zt[8], zs[8];
                tmem_ld8(tacc + cl, zt);
                tmem_ld8(tacc + cl + 16, zs);
                {
                    const uint32_t w0c = TRACE ? clk32() : 0u;
                    mbar_wait(epi_bar(es, BAR_STG_AVAIL + buf), (es.blk / NSTG) & 1);
                    if (TRACE) es.wait_cyc += clk32() - w0c;
                }
                uint8_t* t_hi = stg + buf * STG_BYTES + rl * 128;
                tmem_ld_wait();
                if (in_range) gate_half<MODE, GTU>(p, zt, zs, cbuf + rl0 * COND_LD + cl, cbuf + rl1 * COND_LD + cl, w0, w1, live, t_hi, rl & 7, part, 0);
                tmem_ld8(tacc + cl + 8, zt);
                tmem_ld8(tacc + cl + 24, zs);
                if (k == nmy - 1) release_tile();
                else tmem_ld_wait();
                if (in_range) gate_half<MODE, GTU>(p, zt, zs, cbuf + rl0 * COND_LD + cl, cbuf + rl1 * COND_LD + cl, w0, w1, live, t_hi, rl & 7, part, 1);
                fence_proxy_async();
                __syncwarp();
                if (lane == 0) {
                    mbar_arrive(epi_bar(es, BAR_STG_READY + buf));
                    mbar_arrive(&cond_empty[hs % NCOND]);
                }
                ++es.blk;
            }
            end_tile(t0, t1, 0);
            ++gt;
        };
        // ... and of the res tiles
        int r_step = -1, r_row = 0;
        bool r_valid = false;
        auto res_tile = [&](int j, int t) {
            if (r_step != j) {
                r_step = j;
                r_row = m0_of(j - 1) + rl;
                r_valid = false;
                if (r_row < (int)p.rows && !(dbg & 8)) r_valid = p.grid.frame_utt[r_row / p.steps_per_frame] >= 0;
            }
            const int row = r_row;
            const bool valid = r_valid;
            uint32_t t0, t1;
            const TileDesc td = p.t2[t];
            const uint32_t tacc = wait_tile(t0, t1);
            if (t == p.n_t2 - 1 && p.scr_discard) {
                // The accumulator of the LAST res tile is complete: every activation tile of M tile j - 1 has been read
                // out of the scratch.  Its rows are dead until the gate stores of step j + 1 re-write them in full: drop
                // the lines from the L2 instead of letting it write them back (0.64 GB per layer at 64 x 5 s).  Four
                // threads share a row (128-byte lines `part`, part + 4, ...); the proxy fence orders the discards
                // before the TMA stores that will re-use the rows (issued after this thread's next staging arrive).
                uint8_t* srow = p.scr_discard + (size_t)(scr_of(j - 1) + rl) * ((size_t)4 * p.cpad);
                for (int l = part; l * 128 < 4 * p.cpad; l += 4)
                    asm volatile("discard.global.L2 [%0], 128;" ::"l"(srow + l * 128) : "memory");
                asm volatile("fence.proxy.async;" ::: "memory");
            }
            // 16-column chunks c0 + k with (c0 + k) % 4 == part: at most four (a tile has at most sixteen)
            const int k0 = (part - td.c0) & 3;
            float v[4][16];
#pragma unroll
            for (int k = 0; k < 4; ++k)
                if (k0 + 4 * k < td.nch) tmem_ld16(tacc + 16 * (k0 + 4 * k), v[k]);
            release_tile();
#pragma unroll
            for (int k = 0; k < 4; ++k)
                if (k0 + 4 * k < td.nch) {
                    const int n = 16 * (td.c0 + k0 + 4 * k);               // first packed column of the chunk
                    if (n < p.res_cols) res_chunk<TRACE, MODE>(p, es, v[k], n, valid, rl, part, lane);
                    else skip_chunk(p, v[k], n, (long long)row, valid);
                }
            end_tile(t0, t1, 1);
        };
        for (int j = 0; j <= n_j; ++j)
            for (int e = 0; e < p.n_seq; ++e) {
                const int code = p.seq[e], t = code & 15;
                if (code < 16) {
                    if (j < n_j) gate_tile(j, t);
                } else if (j > 0) {
                    res_tile(j, t);
                }
            }
    }

    tc_fence_before();
    __syncthreads();
    cluster_sync_all();                                             // nobody leaves while the cluster still uses its smem / TMEM
    if (warp == 2) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512) : "memory");
    }
}

// N tiles of (almost) equal width: `chunks` chunks of `cw` columns in ceil(chunks * cw / 256) tiles
int split_tiles(TileDesc* td, int chunks, int cw) {
    const int nt = (chunks * cw + ACC_COLS - 1) / ACC_COLS;
    int c0 = 0;
    for (int t = 0; t < nt; ++t) {
        const int nch = chunks / nt + (t < chunks % nt ? 1 : 0);
        td[t] = TileDesc{c0 * cw, nch * cw, 0, c0, nch};
        c0 += nch;
    }
    return nt;
}

}  // namespace

size_t wn_layer_scratch_bytes(int cpad, int sm_count) { return (size_t)(sm_count / 2) * 2 * 2 * TILE_M * 4 * cpad; }

size_t wn_layer_trace_bytes(int sm_count) { return (size_t)sm_count * 3 * TRACE_SLOTS * 4 * sizeof(uint32_t); }

bool wn_layer_supported(const mbexwn_config_t& c, int cpad, int n_terms, int cond_rows) {
    const int nkb = c.wn_k * (cpad / TILE_K);
    const int per = n_terms == 3 ? 3 : (n_terms == 2 ? 2 : 1);
    if (nkb * per > MAX_K1 || (cpad / TILE_K) * per > MAX_K2) return false;
    if (cond_rows <= 0 || cond_rows > COND_ROWS) return false;      // the gate epilogue reads its conditioning from the smem stage
    if (2 * (cpad / 64) > MAX_BLK) return false;
    if ((2 * cpad + ACC_COLS - 1) / ACC_COLS > MAX_TILES || (cpad + 32 + ACC_COLS - 1) / ACC_COLS > MAX_TILES) return false;
    return true;
}

int wn_layer_forward(WnTcState& st, const WnLayerArgs& a, cudaStream_t s, std::string* error) {
    auto fail = [&](const std::string& m, int code) { if (error) *error = m; return code; };
    const int mode = a.n_terms == 2 ? 0 : (a.n_terms == 3 ? 2 : 1);
    const bool gtu = a.gate == GATE_GTU;
    // instantiations: pairs for the three operand modes and gtu / other gates; clusters of 4 and the trace build for the gtu gate only
    // (another gate type runs the plain pair kernel)
    using Kern = void (*)(LayerParams);
    static const Kern kerns[3][3][2] = {
        {{wn_layer_kernel<false, false, 0, false>, wn_layer_kernel<false, false, 0, true>},
         {wn_layer_kernel<false, false, 1, false>, wn_layer_kernel<false, false, 1, true>},
         {wn_layer_kernel<false, false, 2, false>, wn_layer_kernel<false, false, 2, true>}},
        {{nullptr, wn_layer_kernel<false, true, 0, true>}, {nullptr, wn_layer_kernel<false, true, 1, true>}, {nullptr, wn_layer_kernel<false, true, 2, true>}},
        {{nullptr, wn_layer_kernel<true, false, 0, true>}, {nullptr, wn_layer_kernel<true, false, 1, true>}, {nullptr, wn_layer_kernel<true, false, 2, true>}}};
    static unsigned long long attr_set = 0;                    // bit per device: the attribute belongs to the device's function
    int dev = 0;
    cudaGetDevice(&dev);
    if (!((attr_set >> (dev & 63)) & 1ull)) {
        for (int v = 0; v < 3; ++v)
            for (int m = 0; m < 3; ++m)
                for (int g = 0; g < 2; ++g) {
                    if (!kerns[v][m][g]) continue;
                    cudaError_t e = cudaFuncSetAttribute(kerns[v][m][g], cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_LIMIT);
                    if (e != cudaSuccess) return fail(std::string("cudaFuncSetAttribute(layer kernel): ") + cudaGetErrorString(e), MBEXWN_ERR_CUDA);
                }
        attr_set |= 1ull << (dev & 63);
    }
    const int cpad = a.cpad;
    LayerParams p{};
    int rc;
    // A operand of the gate tiles: one slab of pad_l + 128 + pad_r rows per 64-channel block serves every tap (pad_l a multiple
    // of 8 rows keeps the slab's first row on the 1024-byte swizzle pattern); dilations too wide for a slab slot load per tap
    int smin = 0, smax = 0;
    for (int t = 0; t < a.n_taps; ++t) { if (a.shifts[t] < smin) smin = a.shifts[t]; if (a.shifts[t] > smax) smax = a.shifts[t]; }
    const int pad_l = (-smin + 7) / 8 * 8, pad_r = smax;
    const bool slab = st.slab && pad_l + TILE_M + pad_r <= SLAB_ROWS_MAX;
    const int slab_rows = slab ? pad_l + TILE_M + pad_r : TILE_M;
    p.slab_bytes = slab_rows * 128;
    if ((rc = wn_tc_encode_map(st, &p.tm_hs, a.h_in, a.rows, 2 * cpad, slab_rows, error))) return rc;
    if ((rc = wn_tc_encode_map(st, &p.tm_h, a.h_in, a.rows, 2 * cpad, TILE_M, error))) return rc;
    if ((rc = wn_tc_encode_map(st, &p.tm_hout, a.h_out ? a.h_out : a.h_in, a.rows, 2 * cpad, TILE_M, error))) return rc;
    const int groups_max = a.sm_count / 2;
    if ((rc = wn_tc_encode_map(st, &p.tm_scr, a.scratch, (long long)groups_max * 4 * TILE_M, 2 * cpad, TILE_M, error))) return rc;

    // N tiles and their B boxes (at most two distinct widths per GEMM)
    p.n_t1 = split_tiles(p.t1, cpad / 16, 32);
    p.g_first = p.n_t1;
    p.nb0 = cpad / 64;
    if (st.interleave && p.n_t1 > 1) {
        // the res tiles of the previous M tile run behind the FIRST gate tile: that tile has to end on a 64-channel block
        // (4 chunks) for the staging blocks to complete in order -- 8 chunks (256 columns), the rest in equal tiles
        const int chunks = cpad / 16, first = 8;
        TileDesc rest[MAX_TILES];
        const int nr = split_tiles(rest, chunks - first, 32);
        if (nr + 1 <= MAX_TILES) {
            p.t1[0] = TileDesc{0, first * 32, 0, 0, first};
            for (int t = 0; t < nr; ++t) p.t1[t + 1] = TileDesc{first * 32 + rest[t].n0, rest[t].w, 0, first + rest[t].c0, rest[t].nch};
            p.n_t1 = nr + 1;
            p.g_first = 1;
            p.nb0 = first / 4;
        }
    }
    p.n_t2 = split_tiles(p.t2, a.n2 / 16, 16);
    auto make_b_maps = [&](CUtensorMap* tm, TileDesc* td, int nt, const void* w, int n, int k) -> int {
        int widths[2] = {td[0].w, 0};
        for (int t = 0; t < nt; ++t) {
            if (td[t].w == widths[0]) td[t].bmap = 0;
            else if (widths[1] == 0 || td[t].w == widths[1]) { widths[1] = td[t].w; td[t].bmap = 1; }
            else return MBEXWN_ERR_UNSUPPORTED;
        }
        for (int i = 0; i < 2; ++i) {
            const int w_i = widths[i] ? widths[i] : widths[0];
            int r = wn_tc_encode_map(st, &tm[i], w, n, k, w_i / 2, error);
            if (r) return r;
        }
        return (int)MBEXWN_OK;
    };
    if (a.n2 % 16) return fail("layer kernel: res / skip columns must be a multiple of 16", MBEXWN_ERR_UNSUPPORTED);
    // channel padding beyond the last 16-channel chunk that holds a real channel is not computed: the last gate tile's MMA / B box
    // is that much narrower (C = 340: 704 instead of 768 gate columns), its epilogue still covers the chunks of the whole staging block
    {
        const int phantom = cpad / 16 - (a.c + 15) / 16;
        TileDesc& last = p.t1[p.n_t1 - 1];
        if (phantom > 0 && phantom < last.nch) last.w -= 32 * phantom;
    }
    p.n_seq = 0;
    for (int t = 0; t < p.g_first; ++t) p.seq[p.n_seq++] = t;
    for (int t = 0; t < p.n_t2; ++t) p.seq[p.n_seq++] = 16 + t;
    for (int t = p.g_first; t < p.n_t1; ++t) p.seq[p.n_seq++] = t;
    if ((rc = make_b_maps(p.tm_w1, p.t1, p.n_t1, a.w1, a.n1, a.k1))) return rc;
    if ((rc = make_b_maps(p.tm_w2, p.t2, p.n_t2, a.w2, a.n2, a.k2))) return rc;

    // K loops.  Gate: per A plane and 64-channel block one slab, then one B tile per tap (and per B plane); the e4m3
    // correction blocks come first (their products are rescaled by the first 16-bit MMA of the tile).
    const int ncb = cpad / TILE_K;
    const int b1_lo = a.n_taps * cpad;
    int n = 0;
    auto gate_group = [&](int a_off, const int* b_offs, int nb, int kflags) {
        for (int cb = 0; cb < ncb; ++cb) {
            bool first = true;
            for (int tap = 0; tap < a.n_taps; ++tap)
                for (int ib = 0; ib < nb; ++ib) {
                    KEnt e{};
                    e.a_col = cb * TILE_K + a_off;
                    e.b_col = tap * cpad + cb * TILE_K + b_offs[ib];
                    e.flags = kflags;
                    if (slab) {
                        e.a_row = (short)-pad_l;
                        e.a_view = (short)(pad_l + a.shifts[tap]);
                        if (first) e.flags |= KF_NEW_SLAB;
                        if (tap == a.n_taps - 1 && ib == nb - 1) e.flags |= KF_LAST_OF_SLAB;
                    } else {
                        e.a_row = (short)a.shifts[tap];
                        e.a_view = 0;
                        if (ib == 0) e.flags |= KF_NEW_SLAB;
                        if (ib == nb - 1) e.flags |= KF_LAST_OF_SLAB;
                    }
                    first = false;
                    p.k1[n++] = e;
                }
        }
    };
    const int zero = 0;
    if (a.n_terms == 2) {
        gate_group(cpad, &b1_lo, 1, KF_F8);
        const int n8 = n;
        p.n8_1 = n8;
        gate_group(0, &zero, 1, 0);
        p.k1[0].flags |= KF_OVERWRITE;
        p.k1[n8].flags |= KF_SCALE_D;
    } else if (a.n_terms == 3) {
        const int both[2] = {0, b1_lo};
        gate_group(0, both, 2, 0);                                  // A_hi x (B_hi, B_lo) over the taps of every 64-channel block
        gate_group(cpad, &zero, 1, 0);                              // then A_lo x B_hi
        p.k1[0].flags |= KF_OVERWRITE;
    } else {
        gate_group(0, &zero, 1, 0);
        p.k1[0].flags |= KF_OVERWRITE;
    }
    p.n_k1 = n;
    p.n_slab1 = 0;
    for (int i = 0; i < p.n_k1; ++i)
        if (p.k1[i].flags & KF_NEW_SLAB) p.k1[i].flags |= (p.n_slab1++) << 8;
    n = 0;
    auto res_group = [&](int a_off, const int* b_offs, int nb, int kflags) {
        for (int cb = 0; cb < ncb; ++cb)
            for (int ib = 0; ib < nb; ++ib) {
                KEnt e{};
                e.a_col = cb * TILE_K + a_off;
                e.b_col = cb * TILE_K + b_offs[ib];
                e.flags = kflags | (ib == 0 ? KF_NEW_SLAB : 0) | (ib == nb - 1 ? KF_LAST_OF_SLAB : 0);
                p.k2[n++] = e;
            }
    };
    if (a.n_terms == 2) {
        res_group(cpad, &cpad, 1, KF_F8);
        const int n8 = n;
        p.n8_2 = n8;
        res_group(0, &zero, 1, 0);
        p.k2[0].flags |= KF_OVERWRITE;
        p.k2[n8].flags |= KF_SCALE_D;
    } else if (a.n_terms == 3) {
        const int both[2] = {0, cpad};
        res_group(0, both, 2, 0);
        res_group(cpad, &zero, 1, 0);
        p.k2[0].flags |= KF_OVERWRITE;
    } else {
        res_group(0, &zero, 1, 0);
        p.k2[0].flags |= KF_OVERWRITE;
    }
    p.n_k2 = n;
    p.n_slab2 = 0;
    for (int i = 0; i < p.n_k2; ++i)
        if (p.k2[i].flags & KF_NEW_SLAB) p.k2[i].flags |= (p.n_slab2++) << 8;

    // shared memory carve-up: barriers, A slab ring, B tile ring (as deep as the rest allows), staging blocks, conditioning stages
    int wmax = 0, w1max = 0;
    for (int t = 0; t < p.n_t1; ++t) { if (p.t1[t].w > wmax) wmax = p.t1[t].w; if (p.t1[t].w > w1max) w1max = p.t1[t].w; }
    for (int t = 0; t < p.n_t2; ++t) if (p.t2[t].w > wmax) wmax = p.t2[t].w;
    p.slab_slot = (slab_rows * 128 + 1023) / 1024 * 1024;
    p.b_slot = ((wmax / 2) * 128 + 1023) / 1024 * 1024;
    p.cond_ld = COND_LD;
    const int cond_bytes = (COND_ROWS * COND_LD * 4 + 15) / 16 * 16;            // one half-tile stage
    p.n_a = st.n_a >= 2 && st.n_a <= NA ? st.n_a : NA;
    const int fixed = 1024 /* alignment slack */ + OFF_A + p.n_a * p.slab_slot + NSTG * STG_BYTES + NCOND * cond_bytes;
    p.n_b = (SMEM_LIMIT - fixed) / p.b_slot;
    if (p.n_b > MAX_NB) p.n_b = MAX_NB;
    if (p.n_b < 3) return fail("layer kernel: shared memory too small for the operand rings", MBEXWN_ERR_UNSUPPORTED);
    p.off_b = OFF_A + p.n_a * p.slab_slot;
    p.off_stg = p.off_b + p.n_b * p.b_slot;
    p.off_cond = p.off_stg + NSTG * STG_BYTES;
    const int smem_bytes = 1024 + p.off_cond + NCOND * cond_bytes;
    p.f16 = a.n_terms == 2;
    p.rows = a.rows;
    const int tiles_m = (int)((a.rows + TILE_M - 1) / TILE_M);
    p.tiles_mg = (tiles_m + 1) / 2;
    p.n_gate_blk = cpad / 64;
    p.n_res_blk = a.res_cols / 64;
    p.bias1 = a.bias1; p.cond = a.cond; p.cond_total = a.cond_total; p.cond_rows = a.cond_rows;
    p.c = a.c; p.cpad = cpad; p.lin_up = a.lin_up; p.gate = a.gate; p.write_lo = a.n_terms == 3; p.steps_per_frame = a.steps_per_frame;
    p.out_f16f8 = a.n_terms == 2;
    p.act_lo_scale = a.act_lo_scale; p.h_lo_scale = a.h_lo_scale; p.h_lo_inv = 1.f / a.h_lo_scale;
    for (int u = 0; u < a.lin_up && u < MAX_LIN; ++u) { p.lin_w0[u] = a.lin_w0[u]; p.lin_w1[u] = a.lin_w1[u]; }
    p.bias2 = a.bias2; p.skip = a.skip; p.skip_ld = a.skip_ld; p.skip_c = a.skip_c; p.res_cols = a.res_cols; p.first = a.first;
    p.grid = a.grid;
    p.range_flag = a.range_flag;
    p.scr_discard = st.discard && (4 * cpad) % 128 == 0 ? reinterpret_cast<uint8_t*>(a.scratch) : nullptr;
    p.trace = reinterpret_cast<uint32_t*>(a.trace);
    p.debug = a.trace ? st.debug : 0;

    int groups = groups_max;
    if (p.tiles_mg < groups) groups = p.tiles_mg;
    cudaLaunchConfig_t cfg{};
    cfg.blockDim = dim3(L_THREADS);
    cfg.dynamicSmemBytes = (size_t)smem_bytes;
    cfg.stream = s;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 2;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    // Clusters of 4 when every pair has work for several steps.  How many of them the device holds at once depends on the
    // GPC sizes (a cluster lives inside one GPC), so ask: the persistent schedule needs all of them resident.
    p.cluster = 2;
    if (st.cluster == 4 && gtu && p.tiles_mg >= 2 * groups_max) {
        static int max_quads[64] = {0};
        int& mq = max_quads[dev & 63];
        if (mq == 0) {
            cfg.gridDim = dim3(a.sm_count / 4 * 4);
            attr[0].val.clusterDim.x = 4;
            int n = 0;
            cudaError_t e = cudaOccupancyMaxActiveClusters(&n, kerns[1][mode][gtu ? 1 : 0], &cfg);
            mq = (e == cudaSuccess && n > 0) ? n : -1;
            cudaGetLastError();
        }
        if (mq > 0 && 4 * mq >= st.cluster_min_sms) {
            p.cluster = 4;
            groups = 2 * (mq < a.sm_count / 4 ? mq : a.sm_count / 4);
        }
        st.last_quads = mq;
    }
    cfg.gridDim = dim3(groups * 2);
    attr[0].val.clusterDim.x = p.cluster;
    st.last_cluster = p.cluster;
    // a trace run of a cluster-of-4 launch uses the untraced kernel (the trace build exists for pairs only)
    const int variant = p.cluster == 4 ? 1 : (a.trace && gtu ? 2 : 0);
    cudaError_t e = cudaLaunchKernelEx(&cfg, kerns[variant][mode][gtu ? 1 : 0], p);
    if (e != cudaSuccess) return fail(std::string("fused layer kernel: ") + cudaGetErrorString(e), MBEXWN_ERR_CUDA);
    return MBEXWN_OK;
}

}  // namespace mbx
