// One WaveNet layer (custom_AE_layers.py:305-335) as ONE persistent tcgen05 kernel for sm_100a.
//
//   z   = sum_tap h[r + (tap - 1) d, :] @ W1[tap]                 dilated conv, K = k C, N = 2 C        ("gate tiles")
//   act = tanh(z_t + b + cond) * sigmoid(z_s + b + cond)          conditioning interpolated in the epilogue
//   rs  = act @ [R_res | R_skip W_end]                            1x1, K = C, N = C + c_out              ("res tiles")
//   h' = h + rs[:, :C];  wn_out (+)= rs[:, C:]
//
// The two-launch version (k_wavenet_tc.cu) writes `act` (4 bytes per element) to HBM and reads it back, and reads the
// residual stream twice: 3.5 GB per layer at 64 x 5 s against ~1.45 GB that have to move.  Here a CTA pair (cluster of 2,
// cta_group::2, 256 rows) runs the res tiles of M tile m - 1 right after the gate tiles of M tile m:
//
//   accumulator tiles of step j:   G0(m_j) G1(m_j) G2(m_j) | R0(m_j-1) R1(m_j-1)          (C = 320: 256 + 256 + 128 | 256 + 96)
//
// * `act` of an M tile goes through a per-pair, double-buffered scratch in global memory (2 x 256 rows x 4 cpad bytes per
//   pair, 48 MB for 74 pairs) that is re-written every step and therefore lives in the 126 MB L2: TMA store from the
//   gate epilogue's staging tiles, TMA load as the A operand of the res tiles one step later.  The skew by one M tile
//   takes the store -> load round trip and the last gate epilogue off the tensor pipe's critical path.
// * the residual stream ping-pongs between two buffers (the dilated taps of the neighbouring M tiles must keep reading
//   the layer's *input*): old block TMA-loaded from h_in into a staging tile, updated in place by the epilogue warps,
//   TMA-stored to h_out.  Guard rows pass through as the zeros they are.
// * one operand ring of 4 stages, a stage = A tile (128 rows x 64 K) + B tile (this CTA's half of the N rows x 64 K) behind
//   ONE full / empty mbarrier pair: the producer and MMA warps of the two-launch kernel spend ~70 % of their time in
//   the serial latencies of two waits + two arrivals per K block (ncu source page, profiles/r01k), which is what kept
//   its tensor pipe at 70 %.
// * staging tiles (2 x 32 KB) are handed around with mbarriers only -- no CTA-wide named barrier in the epilogue: the
//   16 epilogue warps arrive on `stg_ready`, one manager thread (warp 2) issues every TMA store, recycles the buffer
//   (`stg_avail`: plain arrive for a gate block, expect_tx + TMA load of the old residual block for a res block) one block
//   ahead of the epilogue, and signals `act_ready` to the producer once the scratch writes of a step have completed.
//
// Warp roles (640 threads): 0 TMA producer, 1 MMA issuer (leader CTA), 2 TMEM allocator + staging manager, 3 conditioning
// stager, 4-19 epilogue (TMEM lane quarter = warp % 4, 16-column chunk of every 64-column block = (warp - 4) / 4).
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

constexpr int TILE_M = 128, TILE_N = 256, TILE_K = 64;
constexpr int NST = 4;                                   // operand ring stages
constexpr int A_BYTES = TILE_M * TILE_K * 2;             // 16 KB
constexpr int B_BYTES = (TILE_N / 2) * TILE_K * 2;       // 16 KB: a CTA of the pair loads half of the B rows
constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
constexpr int NSTG = 2;                                  // staging blocks: [16-bit plane tile 16 KB | lo plane tile 16 KB]
constexpr int STG_BYTES = 2 * TILE_M * 128;
constexpr int COND_ROWS = 16, COND_LD = TILE_N + 4;
constexpr int COND_BYTES = COND_ROWS * COND_LD * 4;
constexpr int OFF_STG = NST * STAGE_BYTES;               // 128 KB
constexpr int OFF_COND = OFF_STG + NSTG * STG_BYTES;     // 192 KB
constexpr int OFF_BAR = OFF_COND + 2 * COND_BYTES;
constexpr int L_SMEM_BYTES = 1024 + OFF_BAR + 512;
static_assert(L_SMEM_BYTES <= 232448, "dynamic shared memory budget (227 KB)");
static_assert(OFF_STG % 1024 == 0 && STG_BYTES % 1024 == 0, "swizzled tiles need 1024-byte alignment");
constexpr int EW = 16, L_THREADS = 128 + 32 * EW;
constexpr int MAX_KB1 = 64, MAX_KB2 = 24, MAX_BLK = 24, MAX_LIN = 32;
constexpr int TRACE_SLOTS = 384;                          // tile records per role and CTA

struct LKB { int a_col, a_shift, b_col; };

struct alignas(64) LayerParams {
    CUtensorMap tm_h;        // layer input (rows, 2 cpad), box 64 x 128: gate A operand and the old residual blocks
    CUtensorMap tm_w1;       // (n1, k1) dilated-conv weights, box 64 x 128
    CUtensorMap tm_w2;       // (n2, k2) res / skip weights, box 64 x 128
    CUtensorMap tm_scr;      // act scratch (n_groups * 512, 2 cpad), box 64 x 128: gate stores, res A operand
    CUtensorMap tm_hout;     // layer output (rows, 2 cpad), box 64 x 128
    LKB kb1[MAX_KB1];        // K blocks of a gate tile: first n8_1 e4m3 blocks (K = 128 bytes), then n16_1 16-bit blocks
    LKB kb2[MAX_KB2];        // K blocks of a res tile (a_shift unused)
    int n8_1, n16_1, n8_2, n16_2;
    int f16;                 // 16-bit operands are fp16 (MBEXWN_PREC_F16F8), else bf16
    int n1, n2, tiles_n1, tiles_n2;
    long long rows;
    int tiles_mg;            // 256-row M tiles
    // staged blocks of one step in epilogue order: kind 0 = gate output (to the scratch), 1 = residual read-modify-write
    int n_blk;
    int blk_kind[MAX_BLK], blk_col[MAX_BLK], blk_last_gate[MAX_BLK];
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
    uint32_t* trace;         // TRACE builds: [cta][role 0..2][TRACE_SLOTS][4]
    int debug;               // TRACE builds only, timing experiments (results are wrong): 1 = no B loads, 2 = no A loads,
                             // 4 = no MMAs are issued, 8 = the epilogue skips its math
};

__device__ __forceinline__ uint32_t clk32() {
    uint32_t c;
    asm volatile("mov.u32 %0, %%clock;" : "=r"(c));
    return c;
}

// conditioning rows (+ bias) of a gate tile -> smem stage, by the 32 lanes of warp 3 (see k_wavenet_tc.cu:gate_stage_fill)
__device__ __forceinline__ void cond_stage_fill(const LayerParams& p, float* buf, int m0, int n0, int width, int lane) {
    const int w4 = width >> 2, hw = width >> 1;
    const int rc0 = m0 / p.lin_up;
    const int total = p.cond_rows * w4;
    for (int f0 = lane; f0 < total; f0 += 32 * 4) {
        float4 v[4], b[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int f = f0 + 32 * i;
            const int r = f / w4, j = (f - r * w4) * 4;
            const int ch = (n0 >> 1) + (j < hw ? j : j - hw);
            const int src_col = (j < hw ? 0 : p.c) + ch;
            v[i] = make_float4(0.f, 0.f, 0.f, 0.f);
            b[i] = make_float4(0.f, 0.f, 0.f, 0.f);
            if (f < total && ch < p.c && rc0 + r < p.cond_total) {
                b[i] = __ldg(reinterpret_cast<const float4*>(p.bias1 + n0 + j));
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

struct EpiState {
    uint8_t* stg;
    uint64_t* stg_avail;
    uint64_t* stg_ready;
    uint32_t blk;            // staged blocks of this CTA so far
    uint32_t wait_cyc;       // TRACE: cycles spent waiting for staging tiles
};

// Gate tile: accumulator columns [0, hw) are the tanh pre-activations of channels n0/2 .., columns [hw, 2 hw) their sigmoid
// partners.  One thread = one row; warp part `part` owns the 16-channel chunk `part` of every 64-channel block.
template <bool TRACE>
__device__ __forceinline__ void epi_gate_tile(const LayerParams& p, EpiState& es, const float* cbuf, uint32_t tacc, int row, int m0,
                                              int n0, int width, int part, int lane) {
    const int hw = width >> 1;
    const bool in_range = row < (int)p.rows && !(TRACE && (p.debug & 8));
    bool valid = false;
    int rl0 = 0, rl1 = 0;
    float w0 = 1.f, w1 = 0.f;
    if (in_range) {
        const int f = row / p.steps_per_frame;
        const int u = p.grid.frame_utt[f];
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
    const float* s0 = cbuf + rl0 * COND_LD;
    const float* s1 = cbuf + rl1 * COND_LD;
    const int r = row - m0, sw = r & 7;
    float zt[16], zs[16];
#pragma unroll 1
    for (int k = 0; k < hw / 64; ++k) {
        const int c0 = 64 * k + 16 * part;                         // tile column of this warp's chunk (tanh half)
        tmem_ld16(tacc + c0, zt);
        tmem_ld16(tacc + hw + c0, zs);
        const uint32_t buf = es.blk % NSTG;
        {
            const uint32_t t0 = TRACE ? clk32() : 0u;
            mbar_wait(&es.stg_avail[buf], (es.blk / NSTG) & 1);
            if (TRACE) es.wait_cyc += clk32() - t0;
        }
        tmem_ld_wait();
        uint8_t* t_hi = es.stg + buf * STG_BYTES + r * 128;
        uint8_t* t_lo = t_hi + TILE_M * 128;
#pragma unroll
        for (int i = 0; i < 16; i += 8) {
            if (!in_range) break;
            float a[8];
            if (valid) {
#pragma unroll
                for (int v4 = 0; v4 < 2; ++v4) {
                    const int col = c0 + i + 4 * v4;
                    const float4 x0 = *reinterpret_cast<const float4*>(s0 + col);
                    const float4 x1 = *reinterpret_cast<const float4*>(s1 + col);
                    const float4 y0 = *reinterpret_cast<const float4*>(s0 + hw + col);
                    const float4 y1 = *reinterpret_cast<const float4*>(s1 + hw + col);
                    const float xa[4] = {x0.x, x0.y, x0.z, x0.w}, xb[4] = {x1.x, x1.y, x1.z, x1.w};
                    const float ya[4] = {y0.x, y0.y, y0.z, y0.w}, yb[4] = {y1.x, y1.y, y1.z, y1.w};
#pragma unroll
                    for (int e = 0; e < 4; ++e) {
                        float t = fmaf(xb[e], w1, fmaf(xa[e], w0, zt[i + 4 * v4 + e]));
                        const float sg = fmaf(yb[e], w1, fmaf(ya[e], w0, zs[i + 4 * v4 + e]));
                        switch (p.gate) {
                            case GATE_GTU: t = fast_tanh(t); break;
                            case GATE_GFU: t = t * rcp_approx(1.f + fabsf(t)); break;
                            case GATE_GSU: t = t * rcp_approx(1.f + sqrtf(fabsf(t))); break;
                            default: break;
                        }
                        a[4 * v4 + e] = t * fast_sigmoid(sg);
                    }
                }
            } else {
#pragma unroll
                for (int e = 0; e < 8; ++e) a[e] = 0.f;              // guard rows stay zero
            }
            // staging tiles, SWIZZLE_128B: 16-byte chunk c of row r sits at r * 128 + ((c ^ (r & 7)) << 4)
            if (p.out_f16f8) {
                uint4 h16;
                uint2 l8, h8;
                split_f16f8(a, p.act_lo_scale, h16, l8, h8);
                *reinterpret_cast<uint4*>(t_hi + (((2 * part + (i >> 3)) ^ sw) << 4)) = h16;
                *reinterpret_cast<uint2*>(t_lo + ((part ^ sw) << 4) + i) = l8;                 // lo8: bytes 0 .. 63 of the row
                *reinterpret_cast<uint2*>(t_lo + (((4 + part) ^ sw) << 4) + i) = h8;           // hi8: bytes 64 .. 127
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
                *reinterpret_cast<uint4*>(t_hi + (((2 * part + (i >> 3)) ^ sw) << 4)) = make_uint4(hw4[0], hw4[1], hw4[2], hw4[3]);
                if (p.write_lo)
                    *reinterpret_cast<uint4*>(t_lo + (((2 * part + (i >> 3)) ^ sw) << 4)) = make_uint4(lw4[0], lw4[1], lw4[2], lw4[3]);
            }
        }
        fence_proxy_async();
        __syncwarp();
        if (lane == 0) mbar_arrive(&es.stg_ready[buf]);
        ++es.blk;
    }
}

// Res tile: columns n < res_cols are residual channels (64-column blocks updated inside the staging tile the manager
// filled with the old values), the c_out columns behind them are accumulated into the WaveNet output in global memory.
template <bool TRACE>
__device__ __forceinline__ void epi_res_tile(const LayerParams& p, EpiState& es, uint32_t tacc, int row, int m0, int n0, int width,
                                             int part, int lane) {
    const bool in_range = row < (int)p.rows;
    bool valid = false;
    if (in_range && !(TRACE && (p.debug & 8))) valid = p.grid.frame_utt[row / p.steps_per_frame] >= 0;
    const int r = row - m0, sw = r & 7;
    float v[16];
#pragma unroll 1
    for (int g = 0; g * 64 < width; ++g) {
        const int ct = 64 * g + 16 * part;                          // tile column of this warp's chunk
        const int n = n0 + ct;
        const bool rmw = n0 + 64 * g < p.res_cols;
        if (!rmw && ct >= width) continue;                          // beyond the tile: nothing in TMEM
        tmem_ld16(tacc + ct, v);
        if (rmw) {
            const uint32_t buf = es.blk % NSTG;
            {
                const uint32_t t0 = TRACE ? clk32() : 0u;
                mbar_wait(&es.stg_avail[buf], (es.blk / NSTG) & 1);
                if (TRACE) es.wait_cyc += clk32() - t0;
            }
            tmem_ld_wait();
            if (valid) {
                uint8_t* t_hi = es.stg + buf * STG_BYTES + r * 128;
                uint8_t* t_lo = t_hi + TILE_M * 128;
                const float4 b0 = __ldg(reinterpret_cast<const float4*>(p.bias2 + n)), b1 = __ldg(reinterpret_cast<const float4*>(p.bias2 + n) + 1);
                const float4 b2 = __ldg(reinterpret_cast<const float4*>(p.bias2 + n) + 2), b3 = __ldg(reinterpret_cast<const float4*>(p.bias2 + n) + 3);
                const float bv[16] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w, b2.x, b2.y, b2.z, b2.w, b3.x, b3.y, b3.z, b3.w};
#pragma unroll
                for (int i = 0; i < 16; i += 8) {
                    uint4* ph16 = reinterpret_cast<uint4*>(t_hi + (((2 * part + (i >> 3)) ^ sw) << 4));
                    float prev[8], o[8];
                    if (p.out_f16f8) {
                        uint2* pl8 = reinterpret_cast<uint2*>(t_lo + ((part ^ sw) << 4) + i);
                        join_f16f8(*ph16, *pl8, p.h_lo_inv, prev);
#pragma unroll
                        for (int e = 0; e < 8; ++e) o[e] = (n + i + e < p.c) ? prev[e] + (v[i + e] + bv[i + e]) : 0.f;
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
                                x[e] = (n + i + idx < p.c) ? pv + (v[i + idx] + bv[i + idx]) : 0.f;
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
            if (lane == 0) mbar_arrive(&es.stg_ready[buf]);
            ++es.blk;
        } else {
            tmem_ld_wait();
            const int sc = n - p.res_cols;                          // 16 consecutive WaveNet-output channels of this row
            if (valid && sc < p.skip_c) {
                float* dst = p.skip + (long long)row * p.skip_ld + sc;
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
        }
    }
}

template <bool TRACE>
__global__ void __launch_bounds__(L_THREADS, 1) wn_layer_kernel(const __grid_constant__ LayerParams p) {
    // K-major SWIZZLE_128B smem matrix descriptor without the address field (see k_wavenet_tc.cu)
    constexpr uint64_t DESC_HI = ((uint64_t)1 << 16) | ((uint64_t)(1024 >> 4) << 32) | ((uint64_t)1 << 46) | ((uint64_t)2 << 61);
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    uint8_t* ring = smem;                                           // stage s: A at s * STAGE_BYTES, B behind it
    uint8_t* stg = smem + OFF_STG;
    float* cond_stage = reinterpret_cast<float*>(smem + OFF_COND);
    uint64_t* full = reinterpret_cast<uint64_t*>(smem + OFF_BAR);
    uint64_t* empty = full + NST;
    uint64_t* tmem_full = empty + NST;
    uint64_t* tmem_empty = tmem_full + 2;
    uint64_t* stg_avail = tmem_empty + 2;
    uint64_t* stg_ready = stg_avail + NSTG;
    uint64_t* cond_full = stg_ready + NSTG;
    uint64_t* cond_empty = cond_full + 2;
    uint64_t* act_ready = cond_empty + 2;
    uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(act_ready + 2);

    const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);
    const int lane = threadIdx.x & 31;
    const int rank = (int)cluster_ctarank();
    const bool leader = rank == 0;
    const int group = blockIdx.x >> 1, n_groups = gridDim.x >> 1;
    const int n_j = group < p.tiles_mg ? (p.tiles_mg - group + n_groups - 1) / n_groups : 0;    // M tiles of this pair
    // rows of this CTA in M tile j of the pair / in the act scratch
    auto m0_of = [&](int j) { return ((group + n_groups * j) * 2 + rank) * TILE_M; };
    auto scr_of = [&](int j) { return ((group * 2 + (j & 1)) * 2 + rank) * TILE_M; };
    auto gate_width = [&](int t) { const int w = p.n1 - t * TILE_N; return w > TILE_N ? TILE_N : w; };
    auto res_width = [&](int t) { const int w = p.n2 - t * TILE_N; return w > TILE_N ? TILE_N : ((w + 15) & ~15); };

    if (warp == 0 && elect_one()) {
        asm volatile("prefetch.tensormap [%0];" ::"l"((uint64_t)&p.tm_h) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"((uint64_t)&p.tm_w1) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"((uint64_t)&p.tm_w2) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"((uint64_t)&p.tm_scr) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"((uint64_t)&p.tm_hout) : "memory");
    }
    if (warp == 1 && elect_one()) {
        for (int s = 0; s < NST; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
        for (int s = 0; s < 2; ++s) {
            mbar_init(&tmem_full[s], 1);
            mbar_init(&tmem_empty[s], 2 * EW);
            mbar_init(&cond_full[s], 32);
            mbar_init(&cond_empty[s], EW);
            mbar_init(&act_ready[s], 1);
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

    if (warp == 0) {
        asm volatile("setmaxnreg.dec.sync.aligned.u32 56;");
        // ===== TMA producer =====
        uint32_t it = 0, tile_it = 0;
        const uint32_t lfull0 = map_to_cta(smem_u32(&full[0]), 0);
        auto load_tile = [&](const CUtensorMap* tma, const CUtensorMap* tmb, const LKB* kb, int nkb, int a_row, int b_row) {
            uint32_t wait_cyc = 0;
            for (int i = 0; i < nkb; ++i, ++it) {
                const uint32_t s = it % NST, ph = (it / NST) & 1;
                const int a_col = kb[i].a_col, a_r = a_row + kb[i].a_shift, b_col = kb[i].b_col;
                {
                    const uint32_t t0 = TRACE ? clk32() : 0u;
                    mbar_wait(&empty[s], ph ^ 1);
                    if (TRACE) wait_cyc += clk32() - t0;
                }
                if (elect_one()) {
                    const int dbg = TRACE ? p.debug : 0;
                    const uint32_t bytes = ((dbg & 1) ? 0 : B_BYTES) + ((dbg & 2) ? 0 : A_BYTES);
                    if (leader) {
                        if (bytes) mbar_expect_tx(&full[s], 2 * bytes);
                        else mbar_arrive(&full[s]);
                    }
                    const uint32_t lbar = lfull0 + s * 8;
                    if (!(dbg & 2)) tma_load_2d_2sm(tma, lbar, ring + s * STAGE_BYTES, a_col, a_r);
                    if (!(dbg & 1)) tma_load_2d_2sm(tmb, lbar, ring + s * STAGE_BYTES + A_BYTES, b_col, b_row);
                }
                __syncwarp();
            }
            if (TRACE && lane == 0 && tile_it < TRACE_SLOTS) {
                uint32_t* t = trc + (0 * TRACE_SLOTS + tile_it) * 4;
                t[0] = clk32(); t[1] = wait_cyc; t[2] = (uint32_t)nkb; t[3] = 0;
            }
            ++tile_it;
        };
        for (int j = 0; j <= n_j; ++j) {
            if (j < n_j)
                for (int t = 0; t < p.tiles_n1; ++t)
                    load_tile(&p.tm_h, &p.tm_w1, p.kb1, p.n8_1 + p.n16_1, m0_of(j), t * TILE_N + rank * (gate_width(t) >> 1));
            if (j > 0) {
                // the act of M tile j - 1 must have landed in the scratch (writes of the async proxy, completed by the manager)
                mbar_wait(&act_ready[(j - 1) & 1], ((j - 1) >> 1) & 1);
                asm volatile("fence.proxy.async;" ::: "memory");
                for (int t = 0; t < p.tiles_n2; ++t)
                    load_tile(&p.tm_scr, &p.tm_w2, p.kb2, p.n8_2 + p.n16_2, scr_of(j - 1), t * TILE_N + rank * (res_width(t) >> 1));
            }
        }
    } else if (warp == 1) {
        asm volatile("setmaxnreg.dec.sync.aligned.u32 56;");
        // ===== MMA issuer (leader CTA) =====
        if (leader) {
            uint32_t it = 0, tile_it = 0;
            const uint32_t ring_base = smem_u32(ring) >> 4;
            auto mma_tile = [&](int width, int n8, int n16) {
                const uint32_t idesc = p.f16 ? make_idesc_fmt0(2 * TILE_M, width) : make_idesc(2 * TILE_M, width);
                const uint32_t as = tile_it & 1, aph = (tile_it >> 1) & 1;
                const uint32_t t0 = TRACE ? clk32() : 0u;
                mbar_wait(&tmem_empty[as], aph ^ 1);
                tc_fence_after();
                const uint32_t t1 = TRACE ? clk32() : 0u;
                uint32_t wait_cyc = 0;
                const uint32_t tacc = tmem_base + as * TILE_N;
                for (int i = 0; i < n8; ++i, ++it) {
                    const uint32_t s = it % NST, ph = (it / NST) & 1;
                    {
                        const uint32_t w0 = TRACE ? clk32() : 0u;
                        mbar_wait(&full[s], ph);
                        if (TRACE) wait_cyc += clk32() - w0;
                    }
                    tc_fence_after();
                    if (elect_one()) {
                        const uint64_t da = DESC_HI | (uint64_t)(ring_base + s * (STAGE_BYTES >> 4));
                        const uint64_t db = da + (A_BYTES >> 4);
                        if (!(TRACE && (p.debug & 4))) {
#pragma unroll
                            for (int k = 0; k < 4; ++k) tc_mma_f8_2sm(tacc, da + 2 * k, db + 2 * k, idesc, !(i == 0 && k == 0));
                        }
                        tc_commit_2sm(&empty[s]);
                    }
                    __syncwarp();
                }
                for (int i = 0; i < n16; ++i, ++it) {
                    const uint32_t s = it % NST, ph = (it / NST) & 1;
                    {
                        const uint32_t w0 = TRACE ? clk32() : 0u;
                        mbar_wait(&full[s], ph);
                        if (TRACE) wait_cyc += clk32() - w0;
                    }
                    tc_fence_after();
                    if (elect_one()) {
                        const uint64_t da = DESC_HI | (uint64_t)(ring_base + s * (STAGE_BYTES >> 4));
                        const uint64_t db = da + (A_BYTES >> 4);
                        if (!(TRACE && (p.debug & 4))) {
                            if (i == 0 && n8 > 0) tc_mma_f16_sd_2sm(tacc, da, db, idesc);          // rescales the e4m3 products by 2^-15
                            else tc_mma_bf16_2sm(tacc, da, db, idesc, i != 0);
#pragma unroll
                            for (int k = 1; k < 4; ++k) tc_mma_bf16_2sm(tacc, da + 2 * k, db + 2 * k, idesc, 1u);
                        }
                        tc_commit_2sm(&empty[s]);
                        if (i == n16 - 1) tc_commit_2sm(&tmem_full[as]);
                    }
                    __syncwarp();
                }
                if (TRACE && lane == 0 && tile_it < TRACE_SLOTS) {
                    uint32_t* t = trc + (1 * TRACE_SLOTS + tile_it) * 4;
                    t[0] = t0; t[1] = t1; t[2] = clk32(); t[3] = wait_cyc;
                }
                ++tile_it;
            };
            for (int j = 0; j <= n_j; ++j) {
                if (j < n_j)
                    for (int t = 0; t < p.tiles_n1; ++t) mma_tile(gate_width(t), p.n8_1, p.n16_1);
                if (j > 0)
                    for (int t = 0; t < p.tiles_n2; ++t) mma_tile(res_width(t), p.n8_2, p.n16_2);
            }
        }
    } else if (warp == 2) {
        asm volatile("setmaxnreg.dec.sync.aligned.u32 56;");
        // ===== staging manager: one thread recycles the staging tiles one block ahead of the epilogue and issues every store =====
        if (lane == 0) {
            int pj = 0, pe = -1, sj = 0, se = -1;                   // prepare / store cursors over (step, block-of-step)
            auto advance = [&](int& j, int& e) {
                for (;;) {
                    if (++e >= p.n_blk) { e = 0; ++j; }
                    if (j > n_j) return false;
                    if (p.blk_kind[e] == 0 ? j < n_j : j > 0) return true;
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
                    if (p.blk_kind[pe] == 0) {
                        mbar_arrive(&stg_avail[buf]);
                    } else {
                        const int col = p.blk_col[pe], row0 = m0_of(pj - 1);
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
                    mbar_arrive(&act_ready[pending_act & 1]);
                    pending_act = -1;
                }
                const uint32_t buf = sb % NSTG;
                mbar_wait(&stg_ready[buf], (sb / NSTG) & 1);
                const uint8_t* t_hi = stg + buf * STG_BYTES;
                const int col = p.blk_col[se];
                if (p.blk_kind[se] == 0) {
                    const int row0 = scr_of(sj);
                    tma_store_2d(&p.tm_scr, t_hi, col, row0);
                    if (p.out_f16f8 || p.write_lo) tma_store_2d(&p.tm_scr, t_hi + TILE_M * 128, p.cpad + col, row0);
                    if (p.blk_last_gate[se]) pending_act = sj;
                } else {
                    const int row0 = m0_of(sj - 1);
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
                mbar_arrive(&act_ready[pending_act & 1]);
            }
        }
    } else if (warp == 3) {
        asm volatile("setmaxnreg.dec.sync.aligned.u32 56;");
        // ===== conditioning stager: one gate tile ahead of the epilogue warps =====
        uint32_t gt = 0;
        for (int j = 0; j < n_j; ++j)
            for (int t = 0; t < p.tiles_n1; ++t, ++gt) {
                const uint32_t b = gt & 1;
                mbar_wait(&cond_empty[b], ((gt >> 1) & 1) ^ 1);
                cond_stage_fill(p, cond_stage + b * (COND_ROWS * COND_LD), m0_of(j), t * TILE_N, gate_width(t), lane);
                mbar_arrive(&cond_full[b]);                         // every lane: its own writes are released
            }
    } else {
        asm volatile("setmaxnreg.inc.sync.aligned.u32 104;");
        // ===== epilogue warps =====
        const int q4 = warp & 3, part = (warp - 4) >> 2;
        EpiState es{stg, stg_avail, stg_ready, 0u, 0u};
        uint32_t tile_it = 0, gt = 0;
        const uint32_t lempty0 = map_to_cta(smem_u32(&tmem_empty[0]), 0);
        auto begin_tile = [&](uint32_t& t0, uint32_t& t1) -> uint32_t {
            const uint32_t as = tile_it & 1, aph = (tile_it >> 1) & 1;
            t0 = TRACE ? clk32() : 0u;
            mbar_wait(&tmem_full[as], aph);
            tc_fence_after();
            t1 = TRACE ? clk32() : 0u;
            return tmem_base + ((uint32_t)(q4 * 32) << 16) + as * TILE_N;
        };
        auto end_tile = [&](uint32_t t0, uint32_t t1, int kind) {
            tc_fence_before();
            __syncwarp();
            if (lane == 0) {
                const uint32_t as = tile_it & 1;
                if (leader) mbar_arrive(&tmem_empty[as]);
                else mbar_arrive_cluster(lempty0 + as * 8);
            }
            if (TRACE && warp == 4 && lane == 0 && tile_it < TRACE_SLOTS) {
                uint32_t* t = trc + (2 * TRACE_SLOTS + tile_it) * 4;
                t[0] = t0; t[1] = t1; t[2] = clk32(); t[3] = es.wait_cyc | ((uint32_t)kind << 31);
                es.wait_cyc = 0;
            }
            ++tile_it;
        };
        for (int j = 0; j <= n_j; ++j) {
            if (j < n_j) {
                const int m0 = m0_of(j);
                for (int t = 0; t < p.tiles_n1; ++t, ++gt) {
                    uint32_t t0, t1;
                    mbar_wait(&cond_full[gt & 1], (gt >> 1) & 1);
                    const uint32_t tacc = begin_tile(t0, t1);
                    epi_gate_tile<TRACE>(p, es, cond_stage + (gt & 1) * (COND_ROWS * COND_LD), tacc, m0 + q4 * 32 + lane, m0, t * TILE_N,
                                         gate_width(t), part, lane);
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&cond_empty[gt & 1]);
                    end_tile(t0, t1, 0);
                }
            }
            if (j > 0) {
                const int m0 = m0_of(j - 1);
                for (int t = 0; t < p.tiles_n2; ++t) {
                    uint32_t t0, t1;
                    const uint32_t tacc = begin_tile(t0, t1);
                    epi_res_tile<TRACE>(p, es, tacc, m0 + q4 * 32 + lane, m0, t * TILE_N, res_width(t), part, lane);
                    end_tile(t0, t1, 1);
                }
            }
        }
    }

    tc_fence_before();
    __syncthreads();
    cluster_sync_all();                                             // nobody leaves while the pair still uses its smem / TMEM
    if (warp == 2) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512) : "memory");
    }
}


}  // namespace

size_t wn_layer_scratch_bytes(int cpad, int sm_count) { return (size_t)(sm_count / 2) * 2 * 2 * TILE_M * 4 * cpad; }

size_t wn_layer_trace_bytes(int sm_count) { return (size_t)sm_count * 3 * TRACE_SLOTS * 4 * sizeof(uint32_t); }

bool wn_layer_supported(const mbexwn_config_t& c, int cpad, int n_terms, int cond_rows) {
    const int nkb = c.wn_k * (cpad / TILE_K);
    const int per = n_terms == 3 ? 3 : (n_terms == 2 ? 2 : 1);
    if (nkb * per > MAX_KB1 || (cpad / TILE_K) * per > MAX_KB2) return false;
    if (cond_rows <= 0 || cond_rows > COND_ROWS - 1) return false;      // the gate epilogue reads its conditioning from the smem stage
    if (2 * (cpad / 64) > MAX_BLK) return false;
    return true;
}

int wn_layer_forward(WnTcState& st, const WnLayerArgs& a, cudaStream_t s, std::string* error) {
    auto fail = [&](const std::string& m, int code) { if (error) *error = m; return code; };
    static unsigned long long attr_set = 0;                    // bit per device: the attribute belongs to the device's function
    int dev = 0;
    cudaGetDevice(&dev);
    if (!((attr_set >> (dev & 63)) & 1ull)) {
        cudaError_t e = cudaFuncSetAttribute(wn_layer_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, L_SMEM_BYTES);
        if (e == cudaSuccess) e = cudaFuncSetAttribute(wn_layer_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, L_SMEM_BYTES);
        if (e != cudaSuccess) return fail(std::string("cudaFuncSetAttribute(layer kernel): ") + cudaGetErrorString(e), MBEXWN_ERR_CUDA);
        attr_set |= 1ull << (dev & 63);
    }
    const int cpad = a.cpad;
    LayerParams p{};
    int rc;
    if ((rc = wn_tc_encode_map(st, &p.tm_h, a.h_in, a.rows, 2 * cpad, TILE_M, error))) return rc;
    if ((rc = wn_tc_encode_map(st, &p.tm_hout, a.h_out ? a.h_out : a.h_in, a.rows, 2 * cpad, TILE_M, error))) return rc;
    if ((rc = wn_tc_encode_map(st, &p.tm_w1, a.w1, a.n1, a.k1, TILE_N / 2, error))) return rc;
    if ((rc = wn_tc_encode_map(st, &p.tm_w2, a.w2, a.n2, a.k2, TILE_N / 2, error))) return rc;
    const int groups_max = a.sm_count / 2;
    if ((rc = wn_tc_encode_map(st, &p.tm_scr, a.scratch, (long long)groups_max * 4 * TILE_M, 2 * cpad, TILE_M, error))) return rc;

    // K-block programs: e4m3 correction blocks first (their products are rescaled by the first 16-bit MMA of a tile)
    const int ncb = cpad / TILE_K;
    int n = 0;
    auto add1 = [&](int a_off, int b_off) {
        for (int tap = 0; tap < a.n_taps; ++tap)
            for (int cb = 0; cb < ncb; ++cb) p.kb1[n++] = LKB{cb * TILE_K + a_off, a.shifts[tap], tap * cpad + cb * TILE_K + b_off};
    };
    const int b1_lo = a.n_taps * cpad;
    if (a.n_terms == 2) { add1(cpad, b1_lo); p.n8_1 = n; add1(0, 0); p.n16_1 = n - p.n8_1; }
    else if (a.n_terms == 3) {
        // hi * hi, lo * hi, hi * lo per (tap, channel block): the summation order of the two-launch kernel
        for (int tap = 0; tap < a.n_taps; ++tap)
            for (int cb = 0; cb < ncb; ++cb) {
                const int ac = cb * TILE_K, bc = tap * cpad + cb * TILE_K;
                p.kb1[n++] = LKB{ac, a.shifts[tap], bc};
                p.kb1[n++] = LKB{ac + cpad, a.shifts[tap], bc};
                p.kb1[n++] = LKB{ac, a.shifts[tap], bc + b1_lo};
            }
        p.n8_1 = 0; p.n16_1 = n;
    }
    else { add1(0, 0); p.n8_1 = 0; p.n16_1 = n; }
    n = 0;
    auto add2 = [&](int a_off, int b_off) {
        for (int cb = 0; cb < ncb; ++cb) p.kb2[n++] = LKB{cb * TILE_K + a_off, 0, cb * TILE_K + b_off};
    };
    if (a.n_terms == 2) { add2(cpad, cpad); p.n8_2 = n; add2(0, 0); p.n16_2 = n - p.n8_2; }
    else if (a.n_terms == 3) {
        for (int cb = 0; cb < ncb; ++cb) {
            p.kb2[n++] = LKB{cb * TILE_K, 0, cb * TILE_K};
            p.kb2[n++] = LKB{cb * TILE_K + cpad, 0, cb * TILE_K};
            p.kb2[n++] = LKB{cb * TILE_K, 0, cb * TILE_K + cpad};
        }
        p.n8_2 = 0; p.n16_2 = n;
    }
    else { add2(0, 0); p.n8_2 = 0; p.n16_2 = n; }
    p.f16 = a.n_terms == 2;
    p.n1 = a.n1; p.n2 = a.n2;
    p.tiles_n1 = (a.n1 + TILE_N - 1) / TILE_N;
    p.tiles_n2 = (a.n2 + TILE_N - 1) / TILE_N;
    p.rows = a.rows;
    const int tiles_m = (int)((a.rows + TILE_M - 1) / TILE_M);
    p.tiles_mg = (tiles_m + 1) / 2;
    // staged blocks of a step in the order the epilogue warps meet them
    int nb = 0;
    for (int t = 0; t < p.tiles_n1; ++t) {
        const int w = a.n1 - t * TILE_N > TILE_N ? TILE_N : a.n1 - t * TILE_N;
        for (int k = 0; k < (w / 2) / 64; ++k) { p.blk_kind[nb] = 0; p.blk_col[nb] = t * (TILE_N / 2) + 64 * k; p.blk_last_gate[nb] = 0; ++nb; }
    }
    if (nb == 0) return fail("layer kernel: no gate blocks", MBEXWN_ERR_UNSUPPORTED);
    p.blk_last_gate[nb - 1] = 1;
    for (int t = 0; t < p.tiles_n2; ++t)
        for (int g = 0; g < 4; ++g) {
            const int col = t * TILE_N + 64 * g;
            if (col < a.res_cols) { p.blk_kind[nb] = 1; p.blk_col[nb] = col; p.blk_last_gate[nb] = 0; ++nb; }
        }
    p.n_blk = nb;
    p.bias1 = a.bias1; p.cond = a.cond; p.cond_total = a.cond_total; p.cond_rows = a.cond_rows;
    p.c = a.c; p.cpad = cpad; p.lin_up = a.lin_up; p.gate = a.gate; p.write_lo = a.n_terms == 3; p.steps_per_frame = a.steps_per_frame;
    p.out_f16f8 = a.n_terms == 2;
    p.act_lo_scale = a.act_lo_scale; p.h_lo_scale = a.h_lo_scale; p.h_lo_inv = 1.f / a.h_lo_scale;
    for (int u = 0; u < a.lin_up && u < MAX_LIN; ++u) { p.lin_w0[u] = a.lin_w0[u]; p.lin_w1[u] = a.lin_w1[u]; }
    p.bias2 = a.bias2; p.skip = a.skip; p.skip_ld = a.skip_ld; p.skip_c = a.skip_c; p.res_cols = a.res_cols; p.first = a.first;
    p.grid = a.grid;
    p.trace = reinterpret_cast<uint32_t*>(a.trace);
    p.debug = a.trace ? st.debug : 0;

    int groups = groups_max;
    if (p.tiles_mg < groups) groups = p.tiles_mg;
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(groups * 2);
    cfg.blockDim = dim3(L_THREADS);
    cfg.dynamicSmemBytes = L_SMEM_BYTES;
    cfg.stream = s;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 2;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    cudaError_t e = a.trace ? cudaLaunchKernelEx(&cfg, wn_layer_kernel<true>, p) : cudaLaunchKernelEx(&cfg, wn_layer_kernel<false>, p);
    if (e != cudaSuccess) return fail(std::string("fused layer kernel: ") + cudaGetErrorString(e), MBEXWN_ERR_CUDA);
    return MBEXWN_OK;
}

}  // namespace mbx
