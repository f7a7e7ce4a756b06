// Tensor-core WaveNet path for sm_100a: TMA-fed, tcgen05.mma with TMEM accumulators, fused epilogues.
//
// One warp-specialised persistent kernel ("tap-GEMM") serves both contractions of a WaveNet layer
// (custom_AE_layers.py:305-335):
//
//   GEMM1 (EPI_GATE)    z[r, :]  = sum_tap h[r + (tap-1)*d, :] @ W1[tap]           K = k*C, N = 2C
//                       act[r,c] = tanh(z_t + b_t + cond_t) * sigmoid(z_s + b_s + cond_s)
//                       cond is the x10 linear interpolation of the mel-rate conditioning, evaluated in the
//                       epilogue from the two neighbouring low-rate rows (never materialised at 1.6 kHz).
//   GEMM2 (EPI_RESSKIP) rs[r, :] = act[r, :] @ [R_res | R_skip @ W_end]               K = C,   N = C + c_out
//                       h[r, :] += rs[:, :C] ; wn_out[r, :] (+)= rs[:, C:]
//                       The skip sum only ever feeds the linear `end` 1x1 (custom_AE_layers.py:337-340), so the skip
//                       half of every res_skip matrix is pre-multiplied by W_end on the host: the kernel accumulates
//                       the c_out (30, padded to 32) channels of the WaveNet output instead of C skip channels.
//
// Implicit GEMM: the dilated taps are *row-shifted TMA loads* of the same activation tensor; guard rows between
// utterances hold zeros and TMA zero-fills outside the tensor, which reproduces the reference's per-utterance SAME
// zero padding without any im2col buffer.  Weight columns are permuted at load so that every N tile (256 wide, the
// last one narrower) holds its tanh channels in the first half and the matching sigmoid channels in the second half,
// so the gate needs no cross-tile exchange.
//
// Precision: operands are bf16, accumulation is fp32 in TMEM.  Activations and weights are stored as bf16 (hi, lo)
// pairs (x ~ hi + lo, 16 mantissa bits); MBEXWN_PREC_BF16X3 runs three products per K block (hi*hi + lo*hi + hi*lo)
// simply by listing three times as many K blocks, MBEXWN_PREC_BF16 lists only hi*hi.  The K-block table
// {A column, A row shift, B column} is the whole "program" of a launch.
//
// Roles (384 threads): warp 0 = TMA producer, warp 1 = MMA issuer, warp 2 = TMEM allocator, warps 4-11 = epilogue
// (TMEM lanes 32*(warp%4)..+31, two warps per lane quarter taking alternate 32-column chunks).  smem ring of
// 4 x (A 128x64 + B 256x64 bf16, SWIZZLE_128B), TMEM double buffered (2 x 256 fp32 columns = all 512) so the epilogue
// of tile i overlaps the MMAs of tile i+1.  The MMA N of a tile is min(256, N - n0), so a narrow last tile costs
// proportionally less tensor time.
//
// The producer and MMA warps run warp-uniform control flow (role index broadcast with shfl, one lane elected per
// issue): ptxas then keeps descriptors and barrier addresses in uniform registers instead of emitting an
// ELECT / R2UR.BROADCAST waterfall per tcgen05.mma, which made the single issuing thread the bottleneck (ncu r01d).
// The gate epilogue reads the conditioning rows of its tile (15 low-rate rows x 256 channels, bias folded in) from a
// double-buffered smem stage that is loaded one tile ahead, so its global-load latency is off the critical path.
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cuda_fp8.h>

#include <map>
#include <vector>

#include "kernels.cuh"
#include "wn_tc.cuh"
#include "tc_common.cuh"

namespace mbx {

namespace {

constexpr int TILE_M = 128, TILE_N = 256, TILE_K = 64, UMMA_K = 16;
constexpr int A_BYTES = TILE_M * TILE_K * 2;            // one A operand tile (128 rows x 64 bf16)
constexpr int RING_BYTES = 192 * 1024;                  // A ring + B ring
constexpr int ACC_STAGES = 2;
constexpr int MAX_RING = 8;
constexpr int TMEM_COLS = ACC_STAGES * TILE_N;          // 512 = all of TMEM
constexpr int MAX_KB = 64;
// Epilogue warps: `parts` warps per TMEM lane quarter splitting the columns of a tile.  The gate epilogue is the long one
// (54 -> ~25 instructions per output and latency bound), so it gets 16 warps (4 per scheduler) and the register budget
// is moved from the producer / MMA warps to them with setmaxnreg; the other epilogues keep 8 warps.
__host__ __device__ constexpr int epi_warps(int epi, int cg) {
    return epi == 1 /* EPI_GATE */ || epi == 3 /* EPI_CONV */ || (epi == 2 /* EPI_RESSKIP */ && cg == 2) ? 16 : 8;
}
__host__ __device__ constexpr int tc_threads(int epi, int cg) { return 128 + 32 * epi_warps(epi, cg); }
constexpr int COND_ROWS = 16;                           // staged conditioning rows per tile (<= 15 used at lin_up = 10)
constexpr int COND_LD = TILE_N + 4;                     // floats per staged row: consecutive rows shift by 4 banks
constexpr int COND_BYTES = COND_ROWS * COND_LD * 4;
constexpr int MAX_LIN_UP = 32;
constexpr size_t SMEM_BYTES = 1024 /*align slack*/ + (size_t)RING_BYTES + 512 + 2 * COND_BYTES;
static_assert(SMEM_BYTES <= 232448, "dynamic shared memory budget (227 KB)");
// Gate kernel with CTA pairs: the activations leave through TMA stores from two 32 KB staging buffers (64 channels x 128
// rows: one 16 KB tile of the fp16 / bf16-hi plane + one 16 KB tile of the [lo8 | hi8] / bf16-lo plane, SWIZZLE_128B like the
// operand tiles); the operand ring shrinks to 4 + 4 slots to make room.
// Res/skip kernel with CTA pairs: the read-modify-write of the residual stream runs through three such staging buffers --
// warp 3 TMA-loads the old 64-channel block, the 16 epilogue warps update it in place, one thread TMA-stores it.
constexpr int RS_STAGES = 3;
constexpr int RS_STAGE_OFF = 129 * 1024;                 // ring (128 KB) + barriers, rounded up to 1024
static_assert(1024 + RS_STAGE_OFF + RS_STAGES * 2 * TILE_M * 128 <= 232448, "smem budget of the res/skip kernel");
constexpr int RING_BYTES_TMA_OUT = 128 * 1024;
constexpr int OUT_STAGE_BYTES = 2 * TILE_M * 128;
static_assert(1024 + RING_BYTES_TMA_OUT + 512 + 2 * COND_BYTES + 2 * OUT_STAGE_BYTES <= 232448, "smem budget of the gate kernel");
static_assert((RING_BYTES_TMA_OUT + 512 + 2 * COND_BYTES) % 1024 == 0, "staging tiles need 1024-byte alignment");

enum Epi { EPI_PLAIN = 0, EPI_GATE = 1, EPI_RESSKIP = 2, EPI_CONV = 3 };

struct KBlock {
    int a_col;      // first A column (elements) of the hi plane for this K block
    int a_shift;    // row shift of the A tile (dilated tap)
    int b_col;      // first B column (elements) of the hi plane
};

struct alignas(64) GemmParams {
    CUtensorMap tm_a;
    CUtensorMap tm_b;
    CUtensorMap tm_out;     // EPI_GATE with CTA pairs: the activation buffer (same geometry as tm_a of the res/skip GEMM)
    KBlock kb[MAX_KB];
    int n_kb;
    int n_terms;            // 1: hi*hi;  3: hi*hi + lo*hi + hi*lo, the lo planes sit a_lo_off / b_lo_off columns further;
                            // 2: fp16 hi*hi + 2^-15 (e4m3 lo8*hi8 + e4m3 hi8*lo8) (MBEXWN_PREC_F16F8): the "lo plane" of a
                            //    64-channel K block is 128 bytes of e4m3, A = [lo8 (64) | hi8 (64)], B = [hi8 (64) | lo8 (64)],
                            //    so both correction products are one K = 128 e4m3 block
    int a_lo_off, b_lo_off;
    int f16;                // main product operands are fp16 (else bf16)
    // scales of the e4m3 planes the epilogues write: x_lo8 = e4m3((x - fp16(x)) * lo_scale), x_hi8 = e4m3(x * hi_scale)
    int out_f16f8;
    float out_lo_scale, in_lo_inv;
    long long rows;         // M
    int n_cols;             // N (multiple of 8; tiles are masked)
    int tiles_m, tiles_n;
    int sched_m_major;      // tile schedule: 1 = a CTA walks all N tiles of its M tiles, 0 = round-robin over (m, n)
    // epilogue
    const float* bias;      // (N) in packed column order
    float* out_f32;         // EPI_PLAIN: (rows, n_cols); EPI_CONV: optional (rows, ld_out)
    // EPI_CONV: bias + activation, fp32 and / or bf16 [hi | lo] output with the sub-pixel unfold
    int ld_out;
    __nv_bfloat16* out_hilo;    // optional (rows * subpixel, 2 * out_cpad)
    int out_cpad, subpixel, cout_per;   // cout_per = n_cols / subpixel
    int cv_act, cv_act_mod, rate;
    const float* alpha;
    float leaky;
    // EPI_GATE
    const float* cond;      // (rows / lin_up, 2C) fp32
    __nv_bfloat16* act;     // (rows, ld_act): [hi (cpad) | lo (cpad)]
    int ld_act;
    int c, cpad, lin_up, gate, write_lo, steps_per_frame;
    int cond_rows;          // conditioning rows one 128-row tile touches, or 0 => read them from global memory
    long long cond_total;   // rows of `cond`
    float lin_w0[MAX_LIN_UP], lin_w1[MAX_LIN_UP];   // (U - u) / U and u / U rounded from double (support_layers.py:19-27)
    // EPI_RESSKIP
    __nv_bfloat16* h;       // (rows, ld_h): [hi | lo]
    int ld_h;
    float* skip;            // (rows, skip_ld): accumulated WaveNet output channels
    int skip_ld, skip_c;
    int res_cols;           // cpad, or 0 for the last layer (skip only)
    int first;              // skip = instead of +=
    FrameGrid grid;
    int* range_flag;        // f16f8 range guard (tc_common.cuh: range_check8), or nullptr
    int debug;              // timing experiments only (option "tc_debug"): 1 = epilogues skip their math and stores,
                            // 2 = every tile loads the operands of tile 0 (L2-resident feed), 4 = no MMAs are issued,
                            // 8 = epilogues skip their global stores, 16 = no gate math, 32 = res/skip does not read h
};

using namespace tcx;

// ---- epilogues: one thread = one accumulator row -----------------------------------------------------------------

// `part` (0 .. nparts-1) selects which 32-column chunks of the tile this warp handles (part, part + nparts, ...);
// `width` = valid tile columns.

__device__ __forceinline__ void epi_plain(const GemmParams& p, uint32_t tacc, long long row, int n0, int width, int part, int nparts) {
    float v[32];
    for (int q = part; q < width / 32; q += nparts) {
        tmem_ld32(tacc + q * 32, v);
        tmem_ld_wait();
        if (row < p.rows) {
#pragma unroll
            for (int i = 0; i < 32; ++i) {
                int n = n0 + q * 32 + i;
                if (n < p.n_cols) p.out_f32[row * p.n_cols + n] = v[i] + (p.bias ? p.bias[n] : 0.f);
            }
        }
    }
}

// Conv epilogue of the mel-rate sub-nets (conv_layers.py:149-165 + PReLU / LeakyReLU, custom_pulsed_generator.py:86-124):
// bias, activation, then fp32 rows and / or the bf16 [hi | lo] planes the next tensor-core conv reads.  A sub-pixel conv
// (conv_layers.py:250-255) unfolds channel c' of row t to row t * f + c' / (cout / f), channel c' % (cout / f).
// Guard rows are written as zeros (the next conv's zero padding; mirrored pads are patched by mirror_guards_kernel).
// 16-column chunks; the bias (and PReLU slope) vectors of a chunk are fetched as float4s before the accumulator wait so that
// their latency hides behind the MMAs instead of sitting in front of every add (ncu r01k: the per-element __ldg made this
// epilogue long-scoreboard bound).
__device__ __forceinline__ void epi_conv(const GemmParams& p, uint32_t tacc, long long row, int n0, int width, int part, int nparts) {
    float v[16], bz[16], al[16];
    bool valid = false;
    if (row < p.rows) {
        long long lo, hi;
        valid = utt_bounds(p.grid, p.rate, row, lo, hi);
    }
    for (int q = part; q < width / 16; q += nparts) {
        const int nq = n0 + q * 16;
        const bool full = nq + 15 < p.n_cols;
        if (full) {
#pragma unroll
            for (int i = 0; i < 16; i += 4) {
                const float4 b4 = __ldg(reinterpret_cast<const float4*>(p.bias + nq + i));
                bz[i] = b4.x; bz[i + 1] = b4.y; bz[i + 2] = b4.z; bz[i + 3] = b4.w;
            }
        } else {
#pragma unroll
            for (int i = 0; i < 16; ++i) bz[i] = nq + i < p.n_cols ? __ldg(p.bias + nq + i) : 0.f;
        }
        if (p.cv_act == ACT_PRELU) {
            const int a0 = nq % p.cv_act_mod;
            if (full && a0 + 16 <= p.cv_act_mod && (a0 & 3) == 0) {
#pragma unroll
                for (int i = 0; i < 16; i += 4) {
                    const float4 a4 = __ldg(reinterpret_cast<const float4*>(p.alpha + a0 + i));
                    al[i] = a4.x; al[i + 1] = a4.y; al[i + 2] = a4.z; al[i + 3] = a4.w;
                }
            } else {
#pragma unroll
                for (int i = 0; i < 16; ++i) al[i] = nq + i < p.n_cols ? __ldg(p.alpha + (nq + i) % p.cv_act_mod) : 0.f;
            }
        } else {
#pragma unroll
            for (int i = 0; i < 16; ++i) al[i] = p.leaky;
        }
        tmem_ld16(tacc + q * 16, v);
        tmem_ld_wait();
        if (row >= p.rows) continue;
#pragma unroll
        for (int i = 0; i < 16; ++i) {
            float x = 0.f;
            if (valid && nq + i < p.n_cols) {
                x = v[i] + bz[i];
                if (p.cv_act == ACT_PRELU || p.cv_act == ACT_LEAKY) x = x >= 0.f ? x : __fmul_rn(al[i], x);
            }
            v[i] = x;
        }
        if (p.out_f32) {
#pragma unroll
            for (int i = 0; i < 16; i += 4) {
                const int n = nq + i;
                if (n + 3 < p.n_cols) *reinterpret_cast<float4*>(p.out_f32 + row * p.ld_out + n) = make_float4(v[i], v[i + 1], v[i + 2], v[i + 3]);
                else
                    for (int e = 0; e < 4; ++e)
                        if (n + e < p.n_cols) p.out_f32[row * p.ld_out + n + e] = v[i + e];
            }
        }
        if (p.out_hilo) {
#pragma unroll
            for (int i = 0; i < 16; i += 8) {
                const int n = nq + i;
                if (n >= p.n_cols) break;                       // n_cols and cout_per are multiples of 8
                const int sub = n / p.cout_per, ch = n - sub * p.cout_per;
                uint32_t hw[4], lw[4];
#pragma unroll
                for (int e = 0; e < 8; e += 2) {
                    __nv_bfloat16 h0, l0, h1, l1;
                    split_bf16(v[i + e], h0, l0);
                    split_bf16(v[i + e + 1], h1, l1);
                    hw[e / 2] = pack2(h0, h1);
                    lw[e / 2] = pack2(l0, l1);
                }
                __nv_bfloat16* dst = p.out_hilo + (row * p.subpixel + sub) * 2 * p.out_cpad + ch;
                *reinterpret_cast<uint4*>(dst) = make_uint4(hw[0], hw[1], hw[2], hw[3]);
                *reinterpret_cast<uint4*>(dst + p.out_cpad) = make_uint4(lw[0], lw[1], lw[2], lw[3]);
            }
        }
    }
}

// ---- gate epilogue with the conditioning rows staged in shared memory ---------------------------------------------
// Tile (m0, n0, width): accumulator columns [0, width/2) are tanh channels n0/2 + j, columns [width/2, width) their
// sigmoid partners.  The 128 rows of the tile interpolate between cond rows rc0 .. rc0 + cond_rows - 1 (rc0 = m0 /
// lin_up).  Stage layout: buf[r][j] = cond[rc0 + r][channel of column j] + bias[n0 + j], j in [0, width).
// The stage of tile j + 1 is filled by warp 3 while the epilogue warps work on tile j; one named barrier per tile (epilogue
// warps + warp 3) hands a finished stage over and frees the buffer read two tiles ago.
// Written by the otherwise idle warp 3, one tile ahead of the epilogue warps (double buffered): lane l copies float4 groups
// l, l + 32, ... of the COND_ROWS x width stage, four row / bias pairs in flight at a time.
__device__ __forceinline__ void gate_stage_fill(const GemmParams& p, float* buf, int m0, int n0, int width, int lane) {
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
            if (f < total && ch < p.c && rc0 + r < p.cond_total) {         // channel padding: C is a multiple of 4
                b[i] = __ldg(reinterpret_cast<const float4*>(p.bias + n0 + j));
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

// One thread = one accumulator row, 16-column chunks (part, part + nparts, ...): the conditioning is interpolated with two
// FMAs per value on top of the accumulator (z + c0 w0 + c1 w1; the reference rounds c0 w0 + c1 w1 first -- a difference of
// one fp32 ulp of the conditioning, far inside the tolerance of the split-precision GEMM that produced z).
// STAGED: conditioning rows (+ bias) come from the shared-memory stage `buf`; otherwise straight from global memory.
// TMA_OUT: the 16 epilogue warps fill a 64-channel x 128-row staging buffer per iteration (warp part p owns the 16-channel
// chunk p of the block), meet at a named barrier, and one thread hands the two tiles to the TMA; `blk_it` counts the blocks of
// this CTA so far (staging buffer = blk_it & 1).  Without it every thread stores its own row pieces straight to global memory.
template <bool STAGED, bool TMA_OUT>
__device__ __forceinline__ void epi_gate(const GemmParams& p, const float* buf, uint32_t tacc, int row, int m0,
                                         int n0, int width, int part, int nparts, uint8_t* out_stage, uint32_t& blk_it) {
    const int hw = width >> 1;
    const int ch_tile = n0 >> 1;
    const bool in_range = row < (int)p.rows;                     // the TMEM loads are warp-collective: no early exit
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
            rl0 = STAGED ? rc - rc0 : rc;
            rl1 = STAGED ? rn - rc0 : rn;
            w0 = p.lin_w0[un];
            w1 = p.lin_w1[un];
        }
    }
    // STAGED: stage rows, column j of the tile (tanh half, then sigmoid half).  Global: cond rows [tanh (C) | sigmoid (C)].
    const float* s0 = STAGED ? buf + rl0 * COND_LD : p.cond + (long long)rl0 * 2 * p.c + ch_tile;
    const float* s1 = STAGED ? buf + rl1 * COND_LD : p.cond + (long long)rl1 * 2 * p.c + ch_tile;
    const int sig_off = STAGED ? hw : p.c;
    __nv_bfloat16* arow = p.act + (long long)row * p.ld_act;
    uint8_t* row8 = reinterpret_cast<uint8_t*>(arow);
    float zt[16], zs[16];
#pragma unroll 1
    for (int q = part; q < hw / 16; q += nparts) {
        tmem_ld16(tacc + q * 16, zt);
        tmem_ld16(tacc + hw + q * 16, zs);
        tmem_ld_wait();
        const int ch0 = ch_tile + q * 16;
#pragma unroll
        for (int i = 0; i < 16; i += 8) {
            if (!in_range) break;                                  // rows past the end: nothing to store (the TMA clips them)
            float a[8];
            if (p.debug & 16) {                                    // timing experiment: no gate math
#pragma unroll
                for (int e = 0; e < 8; ++e) a[e] = zt[i + e] + zs[i + e];
            } else if (valid) {
#pragma unroll
                for (int v4 = 0; v4 < 2; ++v4) {
                    const int col = q * 16 + i + 4 * v4;
                    if (!STAGED && ch_tile + col >= p.c) {             // channel padding (C is a multiple of 4): z = 0 -> act = 0
#pragma unroll
                        for (int e = 0; e < 4; ++e) a[4 * v4 + e] = 0.f;
                        continue;
                    }
                    const float4 x0 = *reinterpret_cast<const float4*>(s0 + col);
                    const float4 x1 = *reinterpret_cast<const float4*>(s1 + col);
                    const float4 y0 = *reinterpret_cast<const float4*>(s0 + sig_off + col);
                    const float4 y1 = *reinterpret_cast<const float4*>(s1 + sig_off + col);
                    const float xa[4] = {x0.x, x0.y, x0.z, x0.w}, xb[4] = {x1.x, x1.y, x1.z, x1.w};
                    const float ya[4] = {y0.x, y0.y, y0.z, y0.w}, yb[4] = {y1.x, y1.y, y1.z, y1.w};
                    if (!STAGED) {                                     // the stage has the bias folded in
                        const float4 b0 = __ldg(reinterpret_cast<const float4*>(p.bias + n0 + col));
                        const float4 b1 = __ldg(reinterpret_cast<const float4*>(p.bias + n0 + hw + col));
                        zt[i + 4 * v4] += b0.x; zt[i + 4 * v4 + 1] += b0.y; zt[i + 4 * v4 + 2] += b0.z; zt[i + 4 * v4 + 3] += b0.w;
                        zs[i + 4 * v4] += b1.x; zs[i + 4 * v4 + 1] += b1.y; zs[i + 4 * v4 + 2] += b1.z; zs[i + 4 * v4 + 3] += b1.w;
                    }
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
            __nv_bfloat16* dst = arow + ch0 + i;
            if (TMA_OUT) {
                // staging tiles, SWIZZLE_128B: 16-byte chunk c of row r sits at r * 128 + ((c ^ (r & 7)) << 4)
                const int r = row - m0, cq = q & 3, sw = r & 7;
                uint8_t* t_hi = out_stage + (blk_it & 1) * OUT_STAGE_BYTES + r * 128;
                uint8_t* t_lo = t_hi + TILE_M * 128;
                if (p.out_f16f8) {
                    uint4 h16;
                    uint2 l8, h8;
                    split_f16f8(a, p.out_lo_scale, h16, l8, h8);
                    *reinterpret_cast<uint4*>(t_hi + (((2 * cq + (i >> 3)) ^ sw) << 4)) = h16;
                    *reinterpret_cast<uint2*>(t_lo + ((cq ^ sw) << 4) + i) = l8;                 // lo8: bytes 0 .. 63 of the row
                    *reinterpret_cast<uint2*>(t_lo + (((4 + cq) ^ sw) << 4) + i) = h8;           // hi8: bytes 64 .. 127
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
                    *reinterpret_cast<uint4*>(t_hi + (((2 * cq + (i >> 3)) ^ sw) << 4)) = make_uint4(hw4[0], hw4[1], hw4[2], hw4[3]);
                    *reinterpret_cast<uint4*>(t_lo + (((2 * cq + (i >> 3)) ^ sw) << 4)) = make_uint4(lw4[0], lw4[1], lw4[2], lw4[3]);
                }
            } else if (p.debug & 8) {                              // timing experiment: no global stores
                if (a[0] + a[1] + a[2] + a[3] + a[4] + a[5] + a[6] + a[7] == 123.456f) *reinterpret_cast<float*>(dst) = a[0];
            } else if (p.out_f16f8) {
                uint4 h16;
                uint2 l8, h8;
                split_f16f8(a, p.out_lo_scale, h16, l8, h8);
                uint8_t* d8 = row8 + f8_off(p.cpad, ch0 + i);
                *reinterpret_cast<uint4*>(dst) = h16;
                *reinterpret_cast<uint2*>(d8) = l8;
                *reinterpret_cast<uint2*>(d8 + 64) = h8;
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
                *reinterpret_cast<uint4*>(dst) = make_uint4(hw4[0], hw4[1], hw4[2], hw4[3]);
                if (p.write_lo) *reinterpret_cast<uint4*>(dst + p.cpad) = make_uint4(lw4[0], lw4[1], lw4[2], lw4[3]);
            }
        }
        if (TMA_OUT) {
            // block (q / 4) of the tile is complete once all 16 warps have written their chunk.  The issuing thread first
            // makes sure the stores it issued earlier have finished reading the *other* buffer (it is written next).
            const bool issuer = threadIdx.x == 128;
            if (issuer) tma_store_wait_read();
            fence_proxy_async();
            asm volatile("bar.sync 2, %0;" ::"n"(512) : "memory");
            if (issuer && !(p.debug & 8)) {
                const uint8_t* t_hi = out_stage + (blk_it & 1) * OUT_STAGE_BYTES;
                const int col = ch_tile + (q >> 2) * 64;
                tma_store_2d(&p.tm_out, t_hi, col, m0);
                if (p.out_f16f8 || p.write_lo) tma_store_2d(&p.tm_out, t_hi + TILE_M * 128, p.cpad + col, m0);
                tma_store_commit();
            }
            ++blk_it;
        }
    }
}

// res/skip epilogue.  A TMEM lane is an accumulator row, so "one thread = one row" global accesses touch 32 different
// 128 B lines per warp instruction and the L1 tag stage becomes the bottleneck (16 k cycles per tile, ncu r01e).  Each
// warp therefore transposes its 32 x 32 fp32 chunk through a private, XOR-swizzled 4 KB smem scratch and does the
// read-modify-write with 4 (bf16 planes) or 8 (fp32 output) lanes per row: every warp instruction covers whole 64 B /
// 128 B row segments.  The old values do not depend on the accumulators, so their loads are issued one chunk ahead
// (the first chunk's before the wait on the MMA) and their latency hides behind the tensor work.
struct ResSkipCtx {
    unsigned vmask;          // bit r: row (row0 + r) of this warp is inside an utterance
    long long row0;          // first row of this warp's lane quarter
    int n0, width, part, nparts;
};

// scratch tile: float4 group g (0..7) of row r lives at r * 32 + ((g ^ (r & 7)) << 2): conflict-free row writes and
// conflict-free 8-rows-x-64-B / 4-rows-x-128-B reads
__device__ __forceinline__ float4 xp_read(const float* S, int r, int g) {
    return *reinterpret_cast<const float4*>(S + r * 32 + ((g ^ (r & 7)) << 2));
}

__device__ __forceinline__ void resskip_load_old(const GemmParams& p, const ResSkipCtx& c, int q, int lane, uint4 (&old)[8]) {
    const int n = c.n0 + q * 32;
    if (p.debug & 32) return;                                      // timing experiment: no read of the old values
    if (q >= c.width / 32 || n >= p.n_cols) return;
    if (n < p.res_cols) {
        const int rr = lane >> 2, cg = lane & 3;
#pragma unroll
        for (int ps = 0; ps < 4; ++ps) {
            const int r = ps * 8 + rr;
            if ((c.vmask >> r) & 1u) {
                const __nv_bfloat16* ph = p.h + (c.row0 + r) * p.ld_h + n + cg * 8;
                old[ps] = *reinterpret_cast<const uint4*>(ph);
                if (p.out_f16f8) {
                    const uint2 l8 = *reinterpret_cast<const uint2*>(reinterpret_cast<const uint8_t*>(p.h + (c.row0 + r) * p.ld_h) + f8_off(p.cpad, n + cg * 8));
                    old[4 + ps] = make_uint4(l8.x, l8.y, 0u, 0u);
                } else {
                    old[4 + ps] = *reinterpret_cast<const uint4*>(ph + p.cpad);
                }
            }
        }
    } else if (!p.first) {
        const int rr = lane >> 3, cg = lane & 7;
        const int sc = n - p.res_cols + cg * 4;
#pragma unroll
        for (int ps = 0; ps < 8; ++ps) {
            const int r = ps * 4 + rr;
            if (((c.vmask >> r) & 1u) && sc < p.skip_c)
                old[ps] = *reinterpret_cast<const uint4*>(p.skip + (c.row0 + r) * p.skip_ld + sc);
        }
    }
}

__device__ __forceinline__ void resskip_store(const GemmParams& p, const ResSkipCtx& c, int q, int lane, const float* S,
                                              const uint4 (&old)[8]) {
    const int n = c.n0 + q * 32;
    if (n >= p.n_cols) return;
    if (p.debug & 8) return;                                       // timing experiment: no stores
    if (n < p.res_cols) {
        // residual stream: h <- h + rs, kept as a bf16 (hi, lo) pair (guard rows stay zero: never written)
        const int rr = lane >> 2, cg = lane & 3;
        const int ch0 = n + cg * 8;
        const float4 ba = __ldg(reinterpret_cast<const float4*>(p.bias + ch0));
        const float4 bb = __ldg(reinterpret_cast<const float4*>(p.bias + ch0) + 1);
        const float bv[8] = {ba.x, ba.y, ba.z, ba.w, bb.x, bb.y, bb.z, bb.w};
#pragma unroll
        for (int ps = 0; ps < 4; ++ps) {
            const int r = ps * 8 + rr;
            if (!((c.vmask >> r) & 1u)) continue;
            const float4 x0 = xp_read(S, r, 2 * cg), x1 = xp_read(S, r, 2 * cg + 1);
            const float xv[8] = {x0.x, x0.y, x0.z, x0.w, x1.x, x1.y, x1.z, x1.w};
            if (p.out_f16f8) {
                float prev[8], o[8];
                join_f16f8(old[ps], make_uint2(old[4 + ps].x, old[4 + ps].y), p.in_lo_inv, prev);
#pragma unroll
                for (int idx = 0; idx < 8; ++idx) o[idx] = (ch0 + idx < p.c) ? prev[idx] + (xv[idx] + bv[idx]) : 0.f;
                range_check8(o, p.range_flag);
                uint4 h16;
                uint2 l8, h8;
                split_f16f8(o, p.out_lo_scale, h16, l8, h8);
                __nv_bfloat16* prow = p.h + (c.row0 + r) * p.ld_h;
                uint8_t* p8 = reinterpret_cast<uint8_t*>(prow) + f8_off(p.cpad, ch0);
                *reinterpret_cast<uint4*>(prow + ch0) = h16;
                *reinterpret_cast<uint2*>(p8) = l8;
                *reinterpret_cast<uint2*>(p8 + 64) = h8;
                continue;
            }
            uint32_t hw[4] = {old[ps].x, old[ps].y, old[ps].z, old[ps].w};
            uint32_t lw[4] = {old[4 + ps].x, old[4 + ps].y, old[4 + ps].z, old[4 + ps].w};
#pragma unroll
            for (int w = 0; w < 4; ++w) {
                float o[2];
#pragma unroll
                for (int e = 0; e < 2; ++e) {
                    const int idx = w * 2 + e;
                    const float prev = __bfloat162float(__ushort_as_bfloat16((unsigned short)(hw[w] >> (16 * e)))) +
                                       __bfloat162float(__ushort_as_bfloat16((unsigned short)(lw[w] >> (16 * e))));
                    o[e] = (ch0 + idx < p.c) ? prev + (xv[idx] + bv[idx]) : 0.f;
                }
                __nv_bfloat16 h0, l0, h1, l1;
                split_bf16(o[0], h0, l0);
                split_bf16(o[1], h1, l1);
                hw[w] = pack2(h0, h1);
                lw[w] = pack2(l0, l1);
            }
            __nv_bfloat16* ph = p.h + (c.row0 + r) * p.ld_h + ch0;
            *reinterpret_cast<uint4*>(ph) = make_uint4(hw[0], hw[1], hw[2], hw[3]);
            *reinterpret_cast<uint4*>(ph + p.cpad) = make_uint4(lw[0], lw[1], lw[2], lw[3]);
        }
    } else {
        const int rr = lane >> 3, cg = lane & 7;
        const int sc = n - p.res_cols + cg * 4;
        if (sc >= p.skip_c) return;                               // skip_c is a multiple of 4
        const float4 b4 = __ldg(reinterpret_cast<const float4*>(p.bias + n + cg * 4));
#pragma unroll
        for (int ps = 0; ps < 8; ++ps) {
            const int r = ps * 4 + rr;
            if (!((c.vmask >> r) & 1u)) continue;
            const float4 x = xp_read(S, r, cg);
            float4 nv = make_float4(x.x + b4.x, x.y + b4.y, x.z + b4.z, x.w + b4.w);
            if (!p.first) {
                nv.x += __uint_as_float(old[ps].x); nv.y += __uint_as_float(old[ps].y);
                nv.z += __uint_as_float(old[ps].z); nv.w += __uint_as_float(old[ps].w);
            }
            *reinterpret_cast<float4*>(p.skip + (c.row0 + r) * p.skip_ld + sc) = nv;
        }
    }
}

__device__ __forceinline__ ResSkipCtx resskip_begin(const GemmParams& p, long long row, int n0, int width, int part, int nparts,
                                                    int lane, uint4 (&old)[8]) {
    ResSkipCtx c;
    c.row0 = row - lane; c.n0 = n0; c.width = width; c.part = part; c.nparts = nparts;
    bool valid = false;
    if (row < p.rows) {
        long long lo, hi;
        valid = utt_bounds(p.grid, p.steps_per_frame, row, lo, hi);
    }
    c.vmask = __ballot_sync(0xffffffffu, valid);
    resskip_load_old(p, c, part, lane, old);
    return c;
}

__device__ __forceinline__ void epi_resskip(const GemmParams& p, const ResSkipCtx& c, uint32_t tacc, int lane, float* S, uint4 (&old)[8]) {
    float v[32];
    uint4 nxt[8];
#pragma unroll 1
    for (int q = c.part; q < c.width / 32; q += c.nparts) {
        tmem_ld32(tacc + q * 32, v);
        resskip_load_old(p, c, q + c.nparts, lane, nxt);
        tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 8; ++j)
            *reinterpret_cast<float4*>(S + lane * 32 + ((j ^ (lane & 7)) << 2)) = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
        __syncwarp();
        resskip_store(p, c, q, lane, S, old);
        __syncwarp();
#pragma unroll
        for (int i = 0; i < 8; ++i) old[i] = nxt[i];
    }
}

// Res/skip epilogue of the CTA-pair kernel: 16 warps, 16-column chunks (warp part p owns chunk p of every 64-channel block).
// Residual blocks are updated inside the staging buffer the TMA filled with the old values (SWIZZLE_128B: 16-byte chunk c
// of row r sits at r * 128 + ((c ^ (r & 7)) << 4)), then stored back by the TMA; guard rows are left untouched (zeros).
// The WaveNet-output columns (32 fp32 per row) are accumulated straight in global memory.
struct RsPipe {
    uint64_t* full_h;        // [RS_STAGES] TMA load of a block has landed
    uint64_t* empty_h;       // [RS_STAGES] the store of a block has finished reading the buffer
    uint8_t* stage;
    uint32_t blk;            // blocks consumed by this CTA so far
    int pending;             // issuer thread: buffer whose store may still be reading, or -1
};

__device__ __forceinline__ void epi_resskip_tma(const GemmParams& p, RsPipe& rp, uint32_t tacc, int row, int m0, int n0, int width,
                                                int part, int nparts) {
    const bool in_range = row < (int)p.rows;
    bool valid = false;
    if (in_range) valid = p.grid.frame_utt[row / p.steps_per_frame] >= 0;
    const bool issuer = threadIdx.x == 128;
    float v[16];
#pragma unroll 1
    for (int q = part; q < width / 16; q += nparts) {
        const int n = n0 + q * 16;
        tmem_ld16(tacc + q * 16, v);
        if (n < p.res_cols) {
            const uint32_t buf = rp.blk % RS_STAGES, ph = (rp.blk / RS_STAGES) & 1;
            mbar_wait(&rp.full_h[buf], ph);
            tmem_ld_wait();
            if (valid && !(p.debug & 8)) {
                const int r = row - m0, cq = q & 3, sw = r & 7;
                uint8_t* t_hi = rp.stage + buf * OUT_STAGE_BYTES + r * 128;
                uint8_t* t_lo = t_hi + TILE_M * 128;
                const float4 b0 = __ldg(reinterpret_cast<const float4*>(p.bias + n)), b1 = __ldg(reinterpret_cast<const float4*>(p.bias + n) + 1);
                const float4 b2 = __ldg(reinterpret_cast<const float4*>(p.bias + n) + 2), b3 = __ldg(reinterpret_cast<const float4*>(p.bias + n) + 3);
                const float bv[16] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w, b2.x, b2.y, b2.z, b2.w, b3.x, b3.y, b3.z, b3.w};
#pragma unroll
                for (int i = 0; i < 16; i += 8) {
                    uint4* ph16 = reinterpret_cast<uint4*>(t_hi + (((2 * cq + (i >> 3)) ^ sw) << 4));
                    float prev[8], o[8];
                    if (p.out_f16f8) {
                        uint2* pl8 = reinterpret_cast<uint2*>(t_lo + ((cq ^ sw) << 4) + i);
                        join_f16f8(*ph16, *pl8, p.in_lo_inv, prev);
#pragma unroll
                        for (int e = 0; e < 8; ++e) o[e] = (n + i + e < p.c) ? prev[e] + (v[i + e] + bv[i + e]) : 0.f;
                        range_check8(o, p.range_flag);
                        uint4 h16;
                        uint2 l8, h8;
                        split_f16f8(o, p.out_lo_scale, h16, l8, h8);
                        *ph16 = h16;
                        *pl8 = l8;
                        *reinterpret_cast<uint2*>(t_lo + (((4 + cq) ^ sw) << 4) + i) = h8;
                    } else {
                        uint4* plo = reinterpret_cast<uint4*>(t_lo + (((2 * cq + (i >> 3)) ^ sw) << 4));
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
            // the block is complete once all 16 warps have updated their chunk.  Before the barrier the issuing thread retires
            // the previous store (it has long finished reading) and hands that buffer back to the loader.
            if (issuer && rp.pending >= 0) {
                tma_store_wait_read();
                mbar_arrive(&rp.empty_h[rp.pending]);
                rp.pending = -1;
            }
            fence_proxy_async();
            asm volatile("bar.sync 2, %0;" ::"n"(512) : "memory");
            if (issuer) {
                const uint8_t* t_hi = rp.stage + buf * OUT_STAGE_BYTES;
                const int col = n0 + (q >> 2) * 64;
                tma_store_2d(&p.tm_out, t_hi, col, m0);
                tma_store_2d(&p.tm_out, t_hi + TILE_M * 128, p.cpad + col, m0);
                tma_store_commit();
                rp.pending = (int)buf;
            }
            ++rp.blk;
        } else {
            tmem_ld_wait();
            const int sc = n - p.res_cols;                       // 16 consecutive WaveNet-output channels of this row
            if (valid && sc < p.skip_c && !(p.debug & 8)) {
                float* dst = p.skip + (long long)row * p.skip_ld + sc;
#pragma unroll
                for (int i = 0; i < 16; i += 4) {
                    const float4 b4 = __ldg(reinterpret_cast<const float4*>(p.bias + n + i));
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

// Operand-ring pipeline.  A tiles (128 rows x 64 K) and B tiles (the CTA's share of the N rows x 64 K) travel through
// two independent smem rings, each slot with its own full/empty mbarrier pair, so an operand that several products
// need is loaded once: per K block the split precision issues hi*hi, lo*hi, hi*lo from {A_hi, A_lo, B_hi, B_lo}
// (4 loads for 3 products instead of 6).
//
// CG = 1: one CTA per 128 x 256 tile.  CG = 2: a CTA pair (cluster of 2, cta_group::2) owns a 256 x 256 tile: each CTA
// loads its own 128 rows of A and *half* of the B rows, the leader issues M = 256 MMAs that read both halves, each CTA
// keeps the accumulators of its own rows in its own TMEM and runs its own epilogue.  Halves the per-SM ingest of B.
template <int EPI, int CG>
__global__ void __launch_bounds__(tc_threads(EPI, CG), 1)
wn_gemm_kernel(const __grid_constant__ GemmParams p) {
    constexpr int EW = epi_warps(EPI, CG), ET = 32 * EW, NPARTS = EW / 4;
    constexpr bool TMA_OUT = EPI == EPI_GATE && CG == 2;
    constexpr bool TMA_RS = EPI == EPI_RESSKIP && CG == 2;
    constexpr int RB = (TMA_OUT || TMA_RS) ? RING_BYTES_TMA_OUT : RING_BYTES;
    constexpr int B_BYTES = (TILE_N / CG) * TILE_K * 2;
    constexpr int NA = CG == 1 ? 4 : ((TMA_OUT || TMA_RS) ? 4 : 6);
    constexpr int NB = CG == 1 ? 4 : ((TMA_OUT || TMA_RS) ? 4 : 6);
    static_assert(NA * A_BYTES + NB * B_BYTES <= RB && NA <= MAX_RING && NB <= MAX_RING, "ring sizes");
    // K-major SWIZZLE_128B smem matrix descriptor without the address field: LBO = 1 (ignored), SBO = 1024 B between
    // 8-row groups, descriptor version 1 (Blackwell), swizzle mode 2 (128 B)
    constexpr uint64_t DESC_HI = ((uint64_t)1 << 16) | ((uint64_t)(1024 >> 4) << 32) | ((uint64_t)1 << 46) | ((uint64_t)2 << 61);
    extern __shared__ uint8_t smem_raw[];
    // align up inside the shared window with plain pointer arithmetic: a round trip through uintptr_t makes ptxas treat
    // every later access as generic (LD.E / ST.E instead of LDS / STS)
    uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    uint8_t* ring_a = smem;
    uint8_t* ring_b = smem + NA * A_BYTES;
    uint64_t* full_a = reinterpret_cast<uint64_t*>(smem + RB);
    uint64_t* empty_a = full_a + MAX_RING;
    uint64_t* full_b = empty_a + MAX_RING;
    uint64_t* empty_b = full_b + MAX_RING;
    uint64_t* tmem_full = empty_b + MAX_RING;
    uint64_t* tmem_empty = tmem_full + ACC_STAGES;
    uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(tmem_empty + ACC_STAGES);
    uint64_t* full_h = tmem_empty + ACC_STAGES + 1;                 // TMA_RS only
    uint64_t* empty_h = full_h + RS_STAGES;
    float* cond_stage = reinterpret_cast<float*>(smem + RB + 512);
    uint8_t* out_stage = smem + RB + 512 + 2 * COND_BYTES;          // TMA_OUT only

    const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);      // warp-uniform role index
    const int lane = threadIdx.x & 31;
    const uint32_t rank = CG == 2 ? cluster_ctarank() : 0;
    const bool leader = rank == 0;
    const int tiles_mg = (p.tiles_m + CG - 1) / CG;                 // M tiles per CTA group
    const int group = blockIdx.x / CG, n_groups = gridDim.x / CG;
    // Tile schedule: a CTA (pair) owns M tiles group, group + n_groups, ... and walks all N tiles of one M tile back to
    // back.  Every CTA gets the same mix of wide and narrow N tiles (a round-robin over (m, n) pairs gave the even
    // CTAs all the 256-wide res/skip tiles and the odd ones all the 96-wide ones), and the A operand of an M tile is
    // re-read by the same SM while it is still in L2.
    // Small problems (fewer M tiles than a few waves of CTAs) keep the round-robin over (m, n) pairs so that all SMs get work.
    const int n_seq = p.sched_m_major ? (group < tiles_mg ? ((tiles_mg - group + n_groups - 1) / n_groups) * p.tiles_n : 0)
                                      : (group < tiles_mg * p.tiles_n ? (tiles_mg * p.tiles_n - group + n_groups - 1) / n_groups : 0);
    auto tile_of = [&](int j, int& m_grp, int& n_blk) {
        if (p.sched_m_major) {
            m_grp = group + n_groups * (j / p.tiles_n);
            n_blk = j % p.tiles_n;
        } else {
            const int t = group + n_groups * j;
            m_grp = t / p.tiles_n;
            n_blk = t - m_grp * p.tiles_n;
        }
    };
    const int ops_a = p.n_terms == 3 ? 2 : 1;                       // A / B tiles per K block

    if (warp == 0 && elect_one()) {
        asm volatile("prefetch.tensormap [%0];" ::"l"((uint64_t)&p.tm_a) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"((uint64_t)&p.tm_b) : "memory");
        if (TMA_OUT || TMA_RS) asm volatile("prefetch.tensormap [%0];" ::"l"((uint64_t)&p.tm_out) : "memory");
    }
    if (warp == 1 && elect_one()) {
        // full barriers: one arrival (the leader's expect_tx for the bytes of *all* CTAs of the group); a peer CTA's
        // TMA only reports its bytes to the leader's barrier (no remote arrive: a release.cluster arrive per load costs a
        // MEMBAR + ERRBAR round trip and serialised the peer's producer thread)
        for (int s = 0; s < NA; ++s) { mbar_init(&full_a[s], 1); mbar_init(&empty_a[s], 1); }
        for (int s = 0; s < NB; ++s) { mbar_init(&full_b[s], 1); mbar_init(&empty_b[s], 1); }
        for (int s = 0; s < ACC_STAGES; ++s) { mbar_init(&tmem_full[s], 1); mbar_init(&tmem_empty[s], CG * EW); }
        if (TMA_RS) for (int s = 0; s < RS_STAGES; ++s) { mbar_init(&full_h[s], 1); mbar_init(&empty_h[s], 1); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 2) {
        if (CG == 1) {
            asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_ptr)), "r"(TMEM_COLS) : "memory");
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
        } else {
            asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_ptr)), "r"(TMEM_COLS) : "memory");
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
        }
    }
    tc_fence_before();
    __syncthreads();
    if (CG == 2) cluster_sync_all();                                // peer barriers are initialised before any remote signal
    tc_fence_after();
    const uint32_t tmem_base = *tmem_ptr;

    // EW == 16: 640 threads start with 96 registers each; the producer / MMA warpgroup hands most of its share to the 16
    // epilogue warps (128 x 56 + 512 x 104 <= 64 K).  Every role branch starts with its own setmaxnreg so that ptxas
    // allocates the branch under the new limit.
#define MBX_REG_DEC() do { if (EW == 16) asm volatile("setmaxnreg.dec.sync.aligned.u32 56;"); } while (0)
#define MBX_REG_INC() do { if (EW == 16) asm volatile("setmaxnreg.inc.sync.aligned.u32 104;"); } while (0)

    if (warp == 0) {
        MBX_REG_DEC();
        // ===== TMA producer: the whole warp walks the schedule, one elected lane issues =====
        uint32_t ia = 0, ib = 0;
        for (int j = 0; j < n_seq; ++j) {
            int m_grp, n_blk;
            tile_of(j, m_grp, n_blk);
            const int m0 = (p.debug & 2) ? (int)rank * TILE_M : (m_grp * CG + (int)rank) * TILE_M;
            int width = p.n_cols - n_blk * TILE_N;
            width = width > TILE_N ? TILE_N : ((width + 15) & ~15);
            const int nb0 = n_blk * TILE_N + (int)rank * (width / CG);      // this CTA's share of the B rows
            if (p.n_terms == 2) {
                // e4m3 correction blocks first (their products are rescaled by the first fp16 MMA of the tile)
                for (int kb = 0; kb < p.n_kb; ++kb) {
                    const int a_col = p.kb[kb].a_col + p.a_lo_off, a_row = m0 + p.kb[kb].a_shift, b_col = p.kb[kb].b_col + p.b_lo_off;
                    {
                        const uint32_t s = ia % NA, ph = (ia / NA) & 1;
                        ++ia;
                        mbar_wait(&empty_a[s], ph ^ 1);
                        if (elect_one()) {
                            if (CG == 1) {
                                mbar_expect_tx(&full_a[s], A_BYTES);
                                tma_load_2d(&p.tm_a, &full_a[s], ring_a + s * A_BYTES, a_col, a_row);
                            } else {
                                const uint32_t lbar = map_to_cta(smem_u32(&full_a[s]), 0);
                                if (leader) mbar_expect_tx(&full_a[s], CG * A_BYTES);
                                tma_load_2d_2sm(&p.tm_a, lbar, ring_a + s * A_BYTES, a_col, a_row);
                            }
                        }
                        __syncwarp();
                    }
                    {
                        const uint32_t s = ib % NB, ph = (ib / NB) & 1;
                        ++ib;
                        mbar_wait(&empty_b[s], ph ^ 1);
                        if (elect_one()) {
                            if (CG == 1) {
                                mbar_expect_tx(&full_b[s], B_BYTES);
                                tma_load_2d(&p.tm_b, &full_b[s], ring_b + s * B_BYTES, b_col, nb0);
                            } else {
                                const uint32_t lbar = map_to_cta(smem_u32(&full_b[s]), 0);
                                if (leader) mbar_expect_tx(&full_b[s], CG * B_BYTES);
                                tma_load_2d_2sm(&p.tm_b, lbar, ring_b + s * B_BYTES, b_col, nb0);
                            }
                        }
                        __syncwarp();
                    }
                }
            }
            for (int kb = 0; kb < p.n_kb; ++kb) {
                const int a_col = p.kb[kb].a_col, a_row = m0 + p.kb[kb].a_shift, b_col = p.kb[kb].b_col;
                for (int o = 0; o < ops_a; ++o) {
                    {   // A tile (hi, then lo)
                        const uint32_t s = ia % NA, ph = (ia / NA) & 1;
                        ++ia;
                        mbar_wait(&empty_a[s], ph ^ 1);
                        if (elect_one()) {
                            const int col = a_col + (o ? p.a_lo_off : 0);
                            if (CG == 1) {
                                mbar_expect_tx(&full_a[s], A_BYTES);
                                tma_load_2d(&p.tm_a, &full_a[s], ring_a + s * A_BYTES, col, a_row);
                            } else {
                                const uint32_t lbar = map_to_cta(smem_u32(&full_a[s]), 0);
                                if (leader) mbar_expect_tx(&full_a[s], CG * A_BYTES);
                                tma_load_2d_2sm(&p.tm_a, lbar, ring_a + s * A_BYTES, col, a_row);
                            }
                        }
                        __syncwarp();
                    }
                    {   // B tile (hi, then lo)
                        const uint32_t s = ib % NB, ph = (ib / NB) & 1;
                        ++ib;
                        mbar_wait(&empty_b[s], ph ^ 1);
                        if (elect_one()) {
                            const int col = b_col + (o ? p.b_lo_off : 0);
                            if (CG == 1) {
                                mbar_expect_tx(&full_b[s], B_BYTES);
                                tma_load_2d(&p.tm_b, &full_b[s], ring_b + s * B_BYTES, col, nb0);
                            } else {
                                const uint32_t lbar = map_to_cta(smem_u32(&full_b[s]), 0);
                                if (leader) mbar_expect_tx(&full_b[s], CG * B_BYTES);
                                tma_load_2d_2sm(&p.tm_b, lbar, ring_b + s * B_BYTES, col, nb0);
                            }
                        }
                        __syncwarp();
                    }
                }
            }
        }
    } else if (warp == 1) {
        // ===== MMA issuer (leader CTA): warp-uniform loop, one elected lane issues the tcgen05 instructions =====
        MBX_REG_DEC();
        if (leader) {
            uint32_t ia = 0, ib = 0, tile_it = 0;
            const uint32_t a_base = smem_u32(ring_a) >> 4, b_base = smem_u32(ring_b) >> 4;
            // first: 0 = accumulate, 1 = the first MMA overwrites the accumulator, 2 = the first MMA rescales it by 2^-15
            auto mma4 = [&](uint32_t tacc, uint32_t sa, uint32_t sb, uint32_t idesc, int first) {
                if (p.debug & 4) return;
                const uint64_t da = DESC_HI | (uint64_t)(a_base + sa * (A_BYTES >> 4));
                const uint64_t db = DESC_HI | (uint64_t)(b_base + sb * (B_BYTES >> 4));
#pragma unroll
                for (int k = 0; k < TILE_K / UMMA_K; ++k) {
                    // advance 16 bf16 = 32 B along K inside the 128 B swizzle row: +2 in 16-byte units
                    if (k == 0 && first == 2) {
                        if (CG == 1) tc_mma_f16_sd(tacc, da, db, idesc);
                        else tc_mma_f16_sd_2sm(tacc, da, db, idesc);
                    } else if (CG == 1) tc_mma_bf16(tacc, da + 2 * k, db + 2 * k, idesc, !(first == 1 && k == 0));
                    else tc_mma_bf16_2sm(tacc, da + 2 * k, db + 2 * k, idesc, !(first == 1 && k == 0));
                }
            };
            // e4m3 products of one K block: [lo8 | hi8] (A) x [hi8 | lo8] (B) = 128 bytes along K = 4 instructions of K = 32
            auto mma8 = [&](uint32_t tacc, uint32_t sa, uint32_t sb, uint32_t idesc, bool first) {
                if (p.debug & 4) return;
                const uint64_t da = DESC_HI | (uint64_t)(a_base + sa * (A_BYTES >> 4));
                const uint64_t db = DESC_HI | (uint64_t)(b_base + sb * (B_BYTES >> 4));
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    if (CG == 1) tc_mma_f8(tacc, da + 2 * k, db + 2 * k, idesc, !(first && k == 0));
                    else tc_mma_f8_2sm(tacc, da + 2 * k, db + 2 * k, idesc, !(first && k == 0));
                }
            };
            auto commit = [&](uint64_t* bar) { if (CG == 1) tc_commit(bar); else tc_commit_2sm(bar); };
            for (int j = 0; j < n_seq; ++j, ++tile_it) {
                int m_grp, n_blk;
                tile_of(j, m_grp, n_blk);
                int width = p.n_cols - n_blk * TILE_N;
                width = width > TILE_N ? TILE_N : ((width + 15) & ~15);
                const uint32_t idesc = p.f16 ? make_idesc_fmt0(TILE_M * CG, width) : make_idesc(TILE_M * CG, width);
                const uint32_t as = tile_it % ACC_STAGES, aph = (tile_it / ACC_STAGES) & 1;
                mbar_wait(&tmem_empty[as], aph ^ 1);
                tc_fence_after();
                const uint32_t tacc = tmem_base + as * TILE_N;
                if (p.n_terms == 2) {
                    for (int kb = 0; kb < p.n_kb; ++kb) {
                        const uint32_t sa = ia % NA, pa = (ia / NA) & 1;
                        const uint32_t sb = ib % NB, pb = (ib / NB) & 1;
                        mbar_wait(&full_a[sa], pa);
                        mbar_wait(&full_b[sb], pb);
                        tc_fence_after();
                        if (elect_one()) {
                            mma8(tacc, sa, sb, idesc, kb == 0);
                            commit(&empty_a[sa]);
                            commit(&empty_b[sb]);
                        }
                        __syncwarp();
                        ia += 1;
                        ib += 1;
                    }
                }
                for (int kb = 0; kb < p.n_kb; ++kb) {
                    const uint32_t sa_hi = ia % NA, pa_hi = (ia / NA) & 1;
                    const uint32_t sb_hi = ib % NB, pb_hi = (ib / NB) & 1;
                    mbar_wait(&full_a[sa_hi], pa_hi);
                    mbar_wait(&full_b[sb_hi], pb_hi);
                    tc_fence_after();
                    if (p.n_terms == 3) {
                        const uint32_t sa_lo = (ia + 1) % NA, pa_lo = ((ia + 1) / NA) & 1;
                        const uint32_t sb_lo = (ib + 1) % NB, pb_lo = ((ib + 1) / NB) & 1;
                        if (elect_one()) mma4(tacc, sa_hi, sb_hi, idesc, kb == 0 ? 1 : 0);     // hi * hi
                        __syncwarp();
                        mbar_wait(&full_a[sa_lo], pa_lo);
                        tc_fence_after();
                        if (elect_one()) {
                            mma4(tacc, sa_lo, sb_hi, idesc, 0);                        // lo * hi
                            commit(&empty_a[sa_lo]);
                            commit(&empty_b[sb_hi]);
                        }
                        __syncwarp();
                        mbar_wait(&full_b[sb_lo], pb_lo);
                        tc_fence_after();
                        if (elect_one()) {
                            mma4(tacc, sa_hi, sb_lo, idesc, 0);                        // hi * lo
                            commit(&empty_a[sa_hi]);
                            commit(&empty_b[sb_lo]);
                            if (kb == p.n_kb - 1) commit(&tmem_full[as]);              // accumulator complete
                        }
                        __syncwarp();
                        ia += 2;
                        ib += 2;
                    } else {
                        if (elect_one()) {
                            mma4(tacc, sa_hi, sb_hi, idesc, kb == 0 ? (p.n_terms == 2 ? 2 : 1) : 0);
                            commit(&empty_a[sa_hi]);
                            commit(&empty_b[sb_hi]);
                            if (kb == p.n_kb - 1) commit(&tmem_full[as]);              // both CTAs of a pair are told
                        }
                        __syncwarp();
                        ia += 1;
                        ib += 1;
                    }
                }
            }
        }
    } else if (warp >= 4) {
        // ===== epilogue warps: every CTA drains the accumulators of its own 128 rows =====
        MBX_REG_INC();
        const int q4 = warp & 3, part = (warp - 4) >> 2;
        const bool staged = EPI == EPI_GATE && p.cond_rows > 0;
        uint32_t tile_it = 0, blk_it = 0;
        RsPipe rpipe{full_h, empty_h, smem + RS_STAGE_OFF, 0u, -1};
        for (int j = 0; j < n_seq; ++j, ++tile_it) {
            int m_grp, n_blk;
            tile_of(j, m_grp, n_blk);
            const uint32_t as = tile_it % ACC_STAGES, aph = (tile_it / ACC_STAGES) & 1;
            const uint32_t tacc = tmem_base + ((uint32_t)(q4 * 32) << 16) + as * TILE_N;
            const long long m0 = (long long)(m_grp * CG + (int)rank) * TILE_M;
            const long long row = m0 + q4 * 32 + lane;
            int width = p.n_cols - n_blk * TILE_N;
            width = width > TILE_N ? TILE_N : ((width + 31) & ~31);
            uint4 old[TMA_RS ? 1 : 8];
            ResSkipCtx rctx;
            if constexpr (EPI == EPI_RESSKIP && !TMA_RS)
                rctx = resskip_begin(p, row, n_blk * TILE_N, width, part, NPARTS, lane, old);   // loads fly during the MMAs
            if (staged) epi_bar_sync<ET + 32>();                    // warp 3 has finished this tile's stage
            mbar_wait(&tmem_full[as], aph);
            tc_fence_after();
            if (p.debug & 1) {
                tc_fence_before();
                __syncwarp();
                if (lane == 0) {
                    if (CG == 1 || leader) mbar_arrive(&tmem_empty[as]);
                    else mbar_arrive_cluster(map_to_cta(smem_u32(&tmem_empty[as]), 0));
                }
                continue;
            }
            if (EPI == EPI_PLAIN) epi_plain(p, tacc, row, n_blk * TILE_N, width, part, NPARTS);
            if (EPI == EPI_CONV) epi_conv(p, tacc, row, n_blk * TILE_N, width, part, NPARTS);
            if (EPI == EPI_GATE) {
                if (staged) epi_gate<true, TMA_OUT>(p, cond_stage + (tile_it & 1) * (COND_ROWS * COND_LD), tacc, (int)row, (int)m0, n_blk * TILE_N, width, part, NPARTS, out_stage, blk_it);
                else epi_gate<false, TMA_OUT>(p, nullptr, tacc, (int)row, (int)m0, n_blk * TILE_N, width, part, NPARTS, out_stage, blk_it);
            }
            if constexpr (TMA_RS) epi_resskip_tma(p, rpipe, tacc, (int)row, (int)m0, n_blk * TILE_N, width, part, NPARTS);
            else if constexpr (EPI == EPI_RESSKIP) epi_resskip(p, rctx, tacc, lane, cond_stage + (warp - 4) * 1024, old);
            tc_fence_before();
            __syncwarp();
            if (lane == 0) {
                if (CG == 1 || leader) mbar_arrive(&tmem_empty[as]);
                else mbar_arrive_cluster(map_to_cta(smem_u32(&tmem_empty[as]), 0));
            }
        }
        if ((TMA_OUT || TMA_RS) && threadIdx.x == 128) tma_store_wait_all();     // the staging buffers and the writes outlive the loop
    } else if (warp == 2) {
        MBX_REG_DEC();
    } else if (warp == 3) {
        // ===== conditioning stager of the gate epilogue =====
        MBX_REG_DEC();
        if (TMA_RS && !(p.debug & 1)) {
            // ===== loader of the old residual blocks: same block order as the epilogue warps =====
            uint32_t hb = 0;
            uint8_t* stage = smem + RS_STAGE_OFF;
            for (int j = 0; j < n_seq; ++j) {
                int m_grp, n_blk;
                tile_of(j, m_grp, n_blk);
                const int m0 = (m_grp * CG + (int)rank) * TILE_M, n0 = n_blk * TILE_N;
                for (int c0 = n0; c0 < p.res_cols && c0 < n0 + TILE_N; c0 += 64, ++hb) {
                    const uint32_t buf = hb % RS_STAGES, ph = (hb / RS_STAGES) & 1;
                    mbar_wait(&empty_h[buf], ph ^ 1);
                    if (elect_one()) {
                        mbar_expect_tx(&full_h[buf], OUT_STAGE_BYTES);
                        tma_load_2d(&p.tm_out, &full_h[buf], stage + buf * OUT_STAGE_BYTES, c0, m0);
                        tma_load_2d(&p.tm_out, &full_h[buf], stage + buf * OUT_STAGE_BYTES + TILE_M * 128, p.cpad + c0, m0);
                    }
                    __syncwarp();
                }
            }
        }
        if (EPI == EPI_GATE && p.cond_rows > 0) {
            for (int j = 0; j < n_seq; ++j) {
                int m_grp, n_blk;
                tile_of(j, m_grp, n_blk);
                int width = p.n_cols - n_blk * TILE_N;
                width = width > TILE_N ? TILE_N : ((width + 31) & ~31);
                gate_stage_fill(p, cond_stage + (j & 1) * (COND_ROWS * COND_LD), (m_grp * CG + (int)rank) * TILE_M, n_blk * TILE_N, width, lane);
                epi_bar_sync<ET + 32>();                            // stage j is ready; everyone is done with tile j - 1
            }
        }
    }

    tc_fence_before();
    __syncthreads();
    if (CG == 2) cluster_sync_all();                                // nobody leaves while the pair still uses its smem / TMEM
    if (warp == 2) {
        tc_fence_after();
        if (CG == 1) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS) : "memory");
        else asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS) : "memory");
    }
}

// start 1x1 (custom_AE_layers.py:280) fused with the split into the bf16 [hi | lo] residual stream: one thread = one
// row x 8 channels; guard rows and the channel padding are written as zeros (the tap-GEMM relies on both).
// Weights (cin x cpad, zero padded) and bias sit in shared memory; a block covers START_ROWS rows.
// Thread = (row lane, 8-channel group): its 8 x cin weights and 8 biases live in registers for the whole block, so the
// inner loop reads only the cin inputs of a row from shared memory (the earlier version re-read the weights from shared
// memory for every row and was bound by that traffic: 0.32 ms for 0.66 GB of output).
constexpr int START_ROWS = 256;
constexpr int START_MAX_CIN = 16;                      // register-resident weights up to here
constexpr int START_WIDE_CIN = 40;                     // START_ROWS x cin floats of staging within the default 48 KB
constexpr int START_LANES = 8;                          // rows in flight per block pass
template <int CIN_MAX>
__global__ void __launch_bounds__(START_LANES * 48, (CIN_MAX > 0 && CIN_MAX <= 8) ? 2 : 1)
start_pack_kernel(const float* __restrict__ x, int ld_x, int cin, const float* __restrict__ w, const float* __restrict__ b,
                  __nv_bfloat16* __restrict__ out, long long rows, int c, int cpad, int rate, FrameGrid g,
                  int f16f8, float lo_scale, int* range_flag) {
    extern __shared__ float sx[];                       // [START_ROWS][cin] inputs, zero for guard rows
    __shared__ int svalid[START_ROWS];
    const int groups = cpad >> 3;                       // <= 48 (cpad <= 384)
    const long long r0 = (long long)blockIdx.x * START_ROWS;
    const int grp = threadIdx.x % groups, lane = threadIdx.x / groups;
    const int ch0 = grp * 8;
    const bool active = lane < START_LANES;
    // CIN_MAX = 0: wide inputs (the blocks behind the first of a multi-block stack read n_out_channels); the weights of a row pass
    // come from global memory / L1 instead of registers
    float wr[CIN_MAX > 0 ? CIN_MAX : 1][8], br[8];
#pragma unroll
    for (int ci = 0; ci < CIN_MAX; ++ci)
#pragma unroll
        for (int j = 0; j < 8; ++j) wr[ci][j] = (ci < cin && ch0 + j < c) ? __ldg(w + ci * c + ch0 + j) : 0.f;
#pragma unroll
    for (int j = 0; j < 8; ++j) br[j] = ch0 + j < c ? __ldg(b + ch0 + j) : 0.f;
    for (int i = threadIdx.x; i < START_ROWS; i += blockDim.x) {
        long long lo, hi;
        svalid[i] = (r0 + i < rows) && utt_bounds(g, rate, r0 + i, lo, hi);
    }
    __syncthreads();
    for (int i = threadIdx.x; i < START_ROWS * cin; i += blockDim.x) {
        const int rl = i / cin;
        sx[i] = svalid[rl] ? x[(r0 + rl) * ld_x + (i - rl * cin)] : 0.f;
    }
    __syncthreads();
    if (!active) return;
    for (int rl = lane; rl < START_ROWS; rl += START_LANES) {
        const long long r = r0 + rl;
        if (r >= rows) break;
        float v[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) v[j] = 0.f;
        if (svalid[rl]) {
#pragma unroll
            for (int ci = 0; ci < CIN_MAX; ++ci) {
                if (ci < cin) {
                    const float xv = sx[rl * cin + ci];
#pragma unroll
                    for (int j = 0; j < 8; ++j) v[j] = fmaf(xv, wr[ci][j], v[j]);
                }
            }
            if (CIN_MAX == 0) {
                for (int ci = 0; ci < cin; ++ci) {
                    const float xv = sx[rl * cin + ci];
#pragma unroll
                    for (int j = 0; j < 8; ++j) v[j] = fmaf(xv, ch0 + j < c ? __ldg(w + ci * c + ch0 + j) : 0.f, v[j]);
                }
            }
#pragma unroll
            for (int j = 0; j < 8; ++j) v[j] += br[j];
        }
        if (f16f8) {
            uint4 h16;
            uint2 l8, h8;
            range_check8(v, range_flag);
            split_f16f8(v, lo_scale, h16, l8, h8);
            uint8_t* p8 = reinterpret_cast<uint8_t*>(out + r * 2 * cpad) + f8_off(cpad, ch0);
            *reinterpret_cast<uint4*>(out + r * 2 * cpad + ch0) = h16;
            *reinterpret_cast<uint2*>(p8) = l8;
            *reinterpret_cast<uint2*>(p8 + 64) = h8;
            continue;
        }
        uint32_t hw[4], lw[4];
#pragma unroll
        for (int j = 0; j < 8; j += 2) {
            __nv_bfloat16 h0, l0, h1, l1;
            split_bf16(v[j], h0, l0);
            split_bf16(v[j + 1], h1, l1);
            hw[j / 2] = pack2(h0, h1);
            lw[j / 2] = pack2(l0, l1);
        }
        *reinterpret_cast<uint4*>(out + r * 2 * cpad + ch0) = make_uint4(hw[0], hw[1], hw[2], hw[3]);
        *reinterpret_cast<uint4*>(out + r * 2 * cpad + cpad + ch0) = make_uint4(lw[0], lw[1], lw[2], lw[3]);
    }
}

// fp32 (rows, c) -> bf16 [hi | lo] (rows, 2 * cpad) for the first tensor-core conv of a sub-net.  Guard rows within
// pad_l / pad_r of an utterance take the value the reference's TFPad1d would put there (custom_layers.py:47-71), every
// other guard row and the channel padding are zero.  One thread per output (row, 8-channel group): a pure gather.
__global__ void pack_hilo_pad_kernel(const float* __restrict__ x, __nv_bfloat16* __restrict__ out, long long rows, int c, int cpad,
                                     int rate, int pad_l, int pad_r, int pad_mode, FrameGrid g) {
    const int groups = cpad >> 3;
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= rows * groups) return;
    const long long r = idx / groups;
    const int ch0 = (int)(idx - r * groups) * 8;
    long long src = -1, lo, hi;
    if (utt_bounds(g, rate, r, lo, hi)) src = r;
    else if (pad_mode != PAD_ZERO) {
        if (utt_bounds(g, rate, r + pad_l, lo, hi) && r < lo) src = pad_index(r, lo, hi, pad_mode);
        else if (r - pad_r >= 0 && utt_bounds(g, rate, r - pad_r, lo, hi) && r >= hi) src = pad_index(r, lo, hi, pad_mode);
    }
    uint32_t hw[4], lw[4];
#pragma unroll
    for (int j = 0; j < 8; j += 2) {
        float v0 = (src >= 0 && ch0 + j < c) ? x[src * c + ch0 + j] : 0.f;
        float v1 = (src >= 0 && ch0 + j + 1 < c) ? x[src * c + ch0 + j + 1] : 0.f;
        __nv_bfloat16 h0, l0, h1, l1;
        split_bf16(v0, h0, l0);
        split_bf16(v1, h1, l1);
        hw[j / 2] = pack2(h0, h1);
        lw[j / 2] = pack2(l0, l1);
    }
    *reinterpret_cast<uint4*>(out + r * 2 * cpad + ch0) = make_uint4(hw[0], hw[1], hw[2], hw[3]);
    *reinterpret_cast<uint4*>(out + r * 2 * cpad + cpad + ch0) = make_uint4(lw[0], lw[1], lw[2], lw[3]);
}

// Patch the guard rows next to every utterance of a bf16 [hi | lo] activation buffer with the mirrored / replicated edge
// rows (TFPad1d SYMMETRIC / EDGE) so that the next conv's row-shifted TMA loads see the reference's padding.
// grid = (n_utt, pad_l + pad_r); threads stride over the row's 16-byte groups.
__global__ void mirror_guards_kernel(__nv_bfloat16* __restrict__ buf, int row_elems, int rate, int pad_l, int pad_r, int pad_mode,
                                     FrameGrid g) {
    const int u = blockIdx.x, j = blockIdx.y;
    const long long lo = (long long)g.utt_begin[u] * rate, hi = (long long)g.utt_end[u] * rate;
    const long long dst = j < pad_l ? lo - 1 - j : hi + (j - pad_l);
    if (dst < 0 || dst >= (long long)g.n_frames * rate) return;
    const long long src = pad_index(dst, lo, hi, pad_mode);
    if (src < 0) return;
    const uint4* s4 = reinterpret_cast<const uint4*>(buf + src * row_elems);
    uint4* d4 = reinterpret_cast<uint4*>(buf + dst * row_elems);
    for (int i = threadIdx.x; i < row_elems / 8; i += blockDim.x) d4[i] = s4[i];
}

// ---- host side ------------------------------------------------------------------------------------------------

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

struct Impl {
    EncodeTiledFn encode = nullptr;
    int sm_count = 0;
    int cta_group = 1;          // 1: one CTA per tile;  2: CTA pairs (cluster of 2, tcgen05 cta_group::2)
};

int make_map(Impl* im, CUtensorMap* tm, const void* base, long long rows, long long cols, int box_rows, std::string* err) {
    if (box_rows == TILE_N) box_rows = TILE_N / im->cta_group;      // a CTA of a pair loads half of the B rows
    cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
    cuuint64_t strides[1] = {(cuuint64_t)cols * 2};
    cuuint32_t box[2] = {(cuuint32_t)TILE_K, (cuuint32_t)box_rows};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = im->encode(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, strides, box, estr,
                            CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                            CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        if (err) *err = "cuTensorMapEncodeTiled failed with code " + std::to_string((int)r);
        return MBEXWN_ERR_CUDA;
    }
    return MBEXWN_OK;
}

int ensure_impl(WnTcState& st, std::string* err) {
    if (st.impl) return MBEXWN_OK;
    Impl* im = new Impl();
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult qres;
    cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres);
    if (e != cudaSuccess || !fn || qres != cudaDriverEntryPointSuccess) {
        if (err) *err = "cuTensorMapEncodeTiled is not available from the driver";
        delete im;
        return MBEXWN_ERR_CUDA;
    }
    im->encode = reinterpret_cast<EncodeTiledFn>(fn);
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&im->sm_count, cudaDevAttrMultiProcessorCount, dev);
    cudaError_t a = cudaSuccess;
    const void* fns[8] = {(const void*)wn_gemm_kernel<EPI_CONV, 1>, (const void*)wn_gemm_kernel<EPI_CONV, 2>,
                          (const void*)wn_gemm_kernel<EPI_PLAIN, 1>, (const void*)wn_gemm_kernel<EPI_GATE, 1>,
                          (const void*)wn_gemm_kernel<EPI_RESSKIP, 1>, (const void*)wn_gemm_kernel<EPI_PLAIN, 2>,
                          (const void*)wn_gemm_kernel<EPI_GATE, 2>, (const void*)wn_gemm_kernel<EPI_RESSKIP, 2>};
    for (int i = 0; i < 8 && a == cudaSuccess; ++i)
        a = cudaFuncSetAttribute(fns[i], cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_BYTES);
    if (a != cudaSuccess) {
        if (err) *err = std::string("cudaFuncSetAttribute(smem): ") + cudaGetErrorString(a);
        delete im;
        return MBEXWN_ERR_CUDA;
    }
    st.impl = im;
    return MBEXWN_OK;
}

}  // namespace

int wn_tc_encode_map(WnTcState& st, void* tensor_map, const void* base, long long rows, long long cols, int box_rows,
                     std::string* error) {
    int rc = ensure_impl(st, error);
    if (rc) return rc;
    Impl* im = reinterpret_cast<Impl*>(st.impl);
    cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
    cuuint64_t strides[1] = {(cuuint64_t)cols * 2};
    cuuint32_t box[2] = {(cuuint32_t)TILE_K, (cuuint32_t)box_rows};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = im->encode(reinterpret_cast<CUtensorMap*>(tensor_map), CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base),
                            dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                            CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        if (error) *error = "cuTensorMapEncodeTiled failed with code " + std::to_string((int)r);
        return MBEXWN_ERR_CUDA;
    }
    return MBEXWN_OK;
}

int wn_tc_sm_count(WnTcState& st) {
    std::string e;
    if (ensure_impl(st, &e)) return 0;
    return reinterpret_cast<Impl*>(st.impl)->sm_count;
}

namespace {

template <int EPI>
cudaError_t launch_gemm(Impl* im, GemmParams& p, cudaStream_t s) {
    p.tiles_m = (int)((p.rows + TILE_M - 1) / TILE_M);
    p.tiles_n = (p.n_cols + TILE_N - 1) / TILE_N;
    const int cg = im->cta_group;
    const int tiles_mg = (p.tiles_m + cg - 1) / cg;
    const int n_tiles = tiles_mg * p.tiles_n;
    int groups = im->sm_count / cg;
    if (n_tiles < groups) groups = n_tiles;
    p.sched_m_major = tiles_mg >= 8 * groups ? 1 : 0;
    if (cg == 1) {
        wn_gemm_kernel<EPI, 1><<<groups, tc_threads(EPI, 1), SMEM_BYTES, s>>>(p);
        return cudaGetLastError();
    }
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(groups * 2);
    cfg.blockDim = dim3(tc_threads(EPI, 2));
    cfg.dynamicSmemBytes = SMEM_BYTES;
    cfg.stream = s;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 2;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    return cudaLaunchKernelEx(&cfg, wn_gemm_kernel<EPI, 2>, p);
}

inline int round_up(int x, int m) { return (x + m - 1) / m * m; }

// K-block program of a launch: one entry per (tap, 64-channel block); the lo planes are addressed by fixed offsets.
int build_kblocks(KBlock* kb, int n_taps, const int* shifts, int cpad) {
    int n = 0;
    for (int tap = 0; tap < n_taps; ++tap)
        for (int cb = 0; cb < cpad / TILE_K; ++cb) {
            if (n >= MAX_KB) return -1;
            kb[n++] = KBlock{cb * TILE_K, shifts[tap], tap * cpad + cb * TILE_K};
        }
    return n;
}

}  // namespace

void wn_tc_carve(const mbexwn_config_t& c, long long rows, int precision, const std::function<void(const char*, size_t)>& add) {
    (void)precision;
    const int cpad = round_up(c.wn_c, TILE_K);
    // two-launch path: h2 = residual stream (in place), a2 = gated activations; fused layer kernel: h2 / a2 = residual stream
    // ping-pong, act_scr = per-pair scratch of the gated activations (L2 resident)
    add("h2", (size_t)rows * 2 * cpad * sizeof(__nv_bfloat16));
    add("a2", (size_t)rows * 2 * cpad * sizeof(__nv_bfloat16));
    static int sm_count = 0;
    if (!sm_count) {
        int dev = 0;
        if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&sm_count, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess)
            sm_count = 0;
        if (sm_count <= 0) sm_count = 148;
    }
    add("act_scr", wn_layer_scratch_bytes(cpad, sm_count));
}

int wn_tc_out_pad(const mbexwn_config_t& c) { return round_up(c.wn_cout, 32); }

int wn_tc_forward(WnTcState& st, const mbexwn_config_t& c, const FrameGrid& g, int precision, const float* wn_in, int ld_in,
                  const float* cond, float* wn_out, const std::function<void*(const char*)>& slot,
                  const std::function<const void*(const std::string&, size_t)>& tensor, cudaStream_t s, int* launches,
                  std::string* error) {
    int rc = ensure_impl(st, error);
    if (rc) return rc;
    Impl* im = reinterpret_cast<Impl*>(st.impl);
    const long long rows = (long long)g.n_frames * c.steps_per_frame;
    const int cpad = round_up(c.wn_c, TILE_K);
    const bool f8 = precision == MBEXWN_PREC_F16F8;
    const int n_terms = precision == MBEXWN_PREC_BF16X3 ? 3 : (f8 ? 2 : 1);
    // power-of-two scales of the e4m3 planes (MBEXWN_PREC_F16F8): residual stream h and gated activations
    const float h_lo = ldexpf(1.f, st.sh_h_lo), a_lo = ldexpf(1.f, st.sh_a_lo);
    const std::string n = c.wn_name;
    if (c.wn_c % 4) { if (error) *error = "tensor-core path needs n_channels % 4 == 0"; return MBEXWN_ERR_UNSUPPORTED; }
    if (c.wn_k * (cpad / TILE_K) > MAX_KB) { if (error) *error = "K-block table too small"; return MBEXWN_ERR_UNSUPPORTED; }
    im->cta_group = st.cta_group == 2 ? 2 : 1;

    __nv_bfloat16* h2 = reinterpret_cast<__nv_bfloat16*>(slot("h2"));
    __nv_bfloat16* a2 = reinterpret_cast<__nv_bfloat16*>(slot("a2"));
    auto fail = [&](const std::string& m, int code) { if (error) *error = m; return code; };
    const int out_pad = wn_tc_out_pad(c);

    // start 1x1 on CUDA cores (K = 6), written straight into the bf16 (hi, lo) residual stream
    {
        const float* w = (const float*)tensor(n + "/start/W", (size_t)c.wn_cin * c.wn_c * 4);
        const float* b = (const float*)tensor(n + "/start/b", (size_t)c.wn_c * 4);
        if (!w || !b) return fail("start conv weights missing", MBEXWN_ERR_MISSING);
        if (c.wn_cin > START_WIDE_CIN) return fail("start conv: more than 40 input channels", MBEXWN_ERR_UNSUPPORTED);
        if (cpad > 384) return fail("start conv: more than 384 residual channels", MBEXWN_ERR_UNSUPPORTED);
        const size_t smem = (size_t)START_ROWS * c.wn_cin * sizeof(float);
        const unsigned nblk = (unsigned)((rows + START_ROWS - 1) / START_ROWS), nthr = (unsigned)(START_LANES * (cpad >> 3));
        if (c.wn_cin <= 8)
            start_pack_kernel<8><<<nblk, nthr, smem, s>>>(wn_in, ld_in, c.wn_cin, w, b, h2, rows, c.wn_c, cpad, c.steps_per_frame, g,
                                                          f8 ? 1 : 0, h_lo, f8 ? st.range_flag : nullptr);
        else if (c.wn_cin <= START_MAX_CIN)
            start_pack_kernel<16><<<nblk, nthr, smem, s>>>(wn_in, ld_in, c.wn_cin, w, b, h2, rows, c.wn_c, cpad, c.steps_per_frame, g,
                                                           f8 ? 1 : 0, h_lo, f8 ? st.range_flag : nullptr);
        else
            start_pack_kernel<0><<<nblk, nthr, smem, s>>>(wn_in, ld_in, c.wn_cin, w, b, h2, rows, c.wn_c, cpad, c.steps_per_frame, g,
                                                          f8 ? 1 : 0, h_lo, f8 ? st.range_flag : nullptr);
        cudaError_t e = cudaGetLastError();
        if (e != cudaSuccess) return fail(std::string("start conv: ") + cudaGetErrorString(e), MBEXWN_ERR_CUDA);
        *launches += 1;
    }

    // x lin_up interpolation weights exactly as the reference initialises them (double, rounded to float)
    float lw0[MAX_LIN_UP], lw1[MAX_LIN_UP];
    if (c.wn_cond_lin_up < 1 || c.wn_cond_lin_up > MAX_LIN_UP) return fail("cond_lin_upsampling outside [1, 32]", MBEXWN_ERR_UNSUPPORTED);
    for (int u = 0; u < c.wn_cond_lin_up; ++u) {
        lw0[u] = (float)((double)(c.wn_cond_lin_up - u) / (double)c.wn_cond_lin_up);
        lw1[u] = (float)((double)u / (double)c.wn_cond_lin_up);
    }
    int cond_rows = (TILE_M + c.wn_cond_lin_up - 2) / c.wn_cond_lin_up + 2;
    if (cond_rows > COND_ROWS - 1 || st.cond_stage == 0) cond_rows = 0;      // does not fit the smem stage: read from global

    CUtensorMap tm_h, tm_a;
    if ((rc = make_map(im, &tm_h, h2, rows, 2 * cpad, TILE_M, error))) return rc;
    if ((rc = make_map(im, &tm_a, a2, rows, 2 * cpad, TILE_M, error))) return rc;
    const std::string tc = f8 ? "/tc8/" : "/tc/";

    cudaEvent_t* ev = nullptr;
    st.n_timed = 0;
    if (st.time_launches) {
        if (!st.events) {
            cudaEvent_t* e = new cudaEvent_t[2 * MBEXWN_MAX_LAYERS + 1];
            for (int i = 0; i < 2 * MBEXWN_MAX_LAYERS + 1; ++i) cudaEventCreate(&e[i]);
            st.events = e;
        }
        ev = reinterpret_cast<cudaEvent_t*>(st.events);
        cudaEventRecord(ev[st.n_timed++], s);
    }
    // fused per-layer kernel (k_wavenet_layer.cu) when the problem is large enough to keep every CTA pair busy
    const int tiles_mg_l = (int)((rows + 2 * TILE_M - 1) / (2 * TILE_M));
    const bool fused = st.fused && im->cta_group == 2 && wn_layer_supported(c, cpad, n_terms, cond_rows) &&
                       (st.fused == 1 || tiles_mg_l >= 2 * (im->sm_count / 2));
    st.last_fused = fused ? 1 : 0;
    if (fused) {
        __nv_bfloat16* hbuf[2] = {h2, a2};
        void* scr = slot("act_scr");
        for (int i = 0; i < c.wn_layers; ++i) {
            const std::string li = std::to_string(i);
            const bool last = i == c.wn_layers - 1;
            const int d = c.wn_dilations[i];
            WnLayerArgs a{};
            a.n1 = 2 * cpad; a.k1 = 2 * c.wn_k * cpad;
            a.n2 = (last ? 0 : cpad) + out_pad; a.k2 = 2 * cpad;
            // W1 rows in the chunk order of the fused kernel: [16 tanh channels | their 16 sigmoid partners] (tc_pack.py)
            a.w1 = tensor(n + (f8 ? "/tcf8/" : "/tcf/") + "W1_" + li, (size_t)a.n1 * a.k1 * 2);
            a.bias1 = (const float*)tensor(n + "/tcf/b1_" + li, (size_t)a.n1 * 4);
            a.w2 = tensor(n + tc + "R_" + li, (size_t)a.n2 * a.k2 * 2);
            a.bias2 = (const float*)tensor(n + "/tc/rb_" + li, (size_t)a.n2 * 4);
            if (!a.w1 || !a.bias1 || !a.w2 || !a.bias2) return fail("packed tensor-core weights missing for layer " + li, MBEXWN_ERR_MISSING);
            a.h_in = hbuf[i & 1];
            a.h_out = last ? nullptr : hbuf[(i + 1) & 1];
            a.scratch = scr;
            a.rows = rows; a.c = c.wn_c; a.cpad = cpad; a.n_terms = n_terms; a.n_taps = c.wn_k;
            for (int t = 0; t < c.wn_k; ++t) a.shifts[t] = (t - (c.wn_causal ? c.wn_k - 1 : (c.wn_k - 1) / 2)) * d;
            a.cond = cond; a.cond_total = rows / c.wn_cond_lin_up; a.cond_rows = cond_rows; a.lin_up = c.wn_cond_lin_up;
            a.gate = c.wn_gate; a.steps_per_frame = c.steps_per_frame;
            for (int u = 0; u < c.wn_cond_lin_up; ++u) { a.lin_w0[u] = lw0[u]; a.lin_w1[u] = lw1[u]; }
            a.act_lo_scale = a_lo; a.h_lo_scale = h_lo;
            a.skip = wn_out; a.skip_ld = out_pad; a.skip_c = out_pad; a.res_cols = last ? 0 : cpad; a.first = i == 0;
            a.grid = g; a.sm_count = im->sm_count;
            a.range_flag = f8 ? st.range_flag : nullptr;
            a.trace = nullptr;
            if (st.trace_on == i + 1) {
                if (!st.trace && cudaMalloc(&st.trace, wn_layer_trace_bytes(im->sm_count)) != cudaSuccess) return fail("trace buffer", MBEXWN_ERR_CUDA);
                cudaMemsetAsync(st.trace, 0, wn_layer_trace_bytes(im->sm_count), s);
                a.trace = st.trace;
            }
            if ((rc = wn_layer_forward(st, a, s, error))) return rc;
            if (ev) cudaEventRecord(ev[st.n_timed++], s);
            *launches += 1;
        }
        return MBEXWN_OK;
    }
    for (int i = 0; i < c.wn_layers; ++i) {
        const std::string li = std::to_string(i);
        const bool last = i == c.wn_layers - 1;
        const int d = c.wn_dilations[i];
        const int n1 = 2 * cpad, k1 = 2 * c.wn_k * cpad;                 // W1 packed: (n1, [hi | lo] x k x cpad)
        const int n2 = (last ? 0 : cpad) + out_pad, k2 = 2 * cpad;       // R packed: (n2, [hi | lo] x cpad)
        const void* w1 = tensor(n + tc + "W1_" + li, (size_t)n1 * k1 * 2);
        const float* b1 = (const float*)tensor(n + "/tc/b1_" + li, (size_t)n1 * 4);
        const void* w2 = tensor(n + tc + "R_" + li, (size_t)n2 * k2 * 2);
        const float* b2 = (const float*)tensor(n + "/tc/rb_" + li, (size_t)n2 * 4);
        if (!w1 || !b1 || !w2 || !b2) return fail("packed tensor-core weights missing for layer " + li, MBEXWN_ERR_MISSING);

        GemmParams p1{};
        p1.tm_a = tm_h;
        p1.tm_out = tm_a;
        if ((rc = make_map(im, &p1.tm_b, w1, n1, k1, TILE_N, error))) return rc;
        int shifts[16];
        for (int t = 0; t < c.wn_k; ++t) shifts[t] = (t - (c.wn_causal ? c.wn_k - 1 : (c.wn_k - 1) / 2)) * d;
        p1.n_kb = build_kblocks(p1.kb, c.wn_k, shifts, cpad);
        p1.n_terms = n_terms; p1.a_lo_off = cpad; p1.b_lo_off = c.wn_k * cpad;
        if (f8) { p1.f16 = 1; p1.out_f16f8 = 1; p1.out_lo_scale = a_lo; }
        p1.rows = rows; p1.n_cols = n1; p1.bias = b1; p1.cond = cond; p1.act = a2; p1.ld_act = 2 * cpad;
        p1.c = c.wn_c; p1.cpad = cpad; p1.lin_up = c.wn_cond_lin_up; p1.gate = c.wn_gate; p1.write_lo = n_terms == 3;
        p1.steps_per_frame = c.steps_per_frame; p1.grid = g; p1.debug = st.debug;
        p1.cond_rows = cond_rows; p1.cond_total = rows / c.wn_cond_lin_up;
        for (int u = 0; u < c.wn_cond_lin_up; ++u) { p1.lin_w0[u] = lw0[u]; p1.lin_w1[u] = lw1[u]; }
        cudaError_t e = launch_gemm<EPI_GATE>(im, p1, s);
        if (e != cudaSuccess) return fail(std::string("gate GEMM: ") + cudaGetErrorString(e), MBEXWN_ERR_CUDA);
        if (ev) cudaEventRecord(ev[st.n_timed++], s);

        GemmParams p2{};
        p2.tm_a = tm_a;
        p2.tm_out = tm_h;
        if ((rc = make_map(im, &p2.tm_b, w2, n2, k2, TILE_N, error))) return rc;
        int zero = 0;
        p2.n_kb = build_kblocks(p2.kb, 1, &zero, cpad);
        p2.n_terms = n_terms; p2.a_lo_off = cpad; p2.b_lo_off = cpad;
        if (f8) { p2.f16 = 1; p2.out_f16f8 = 1; p2.out_lo_scale = h_lo; p2.in_lo_inv = 1.f / h_lo; }
        p2.rows = rows; p2.n_cols = n2; p2.bias = b2; p2.h = h2; p2.ld_h = 2 * cpad;
        p2.skip = wn_out; p2.skip_ld = out_pad; p2.skip_c = out_pad;
        p2.c = c.wn_c; p2.cpad = cpad; p2.res_cols = last ? 0 : cpad; p2.first = i == 0;
        p2.steps_per_frame = c.steps_per_frame; p2.grid = g; p2.debug = st.debug;
        p2.range_flag = f8 ? st.range_flag : nullptr;
        e = launch_gemm<EPI_RESSKIP>(im, p2, s);
        if (e != cudaSuccess) return fail(std::string("res/skip GEMM: ") + cudaGetErrorString(e), MBEXWN_ERR_CUDA);
        if (ev) cudaEventRecord(ev[st.n_timed++], s);
        *launches += 2;
    }
    return MBEXWN_OK;
}

int wn_tc_pack(const float* x, void* out_hilo, long long rows, int c, int cpad, int rate, int pad_l, int pad_r, int pad_mode,
               const FrameGrid& g, cudaStream_t s, std::string* error) {
    const long long total = rows * (cpad / 8);
    pack_hilo_pad_kernel<<<(unsigned)((total + 255) / 256), 256, 0, s>>>(x, reinterpret_cast<__nv_bfloat16*>(out_hilo), rows, c, cpad,
                                                                          rate, pad_l, pad_r, pad_mode, g);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) { if (error) *error = std::string("pack: ") + cudaGetErrorString(e); return MBEXWN_ERR_CUDA; }
    return MBEXWN_OK;
}

int wn_tc_mirror(void* hilo, long long rows, int cpad, int rate, int pad_l, int pad_r, int pad_mode, const FrameGrid& g,
                 cudaStream_t s, std::string* error) {
    (void)rows;
    if (pad_mode == PAD_ZERO || pad_l + pad_r <= 0) return MBEXWN_OK;
    dim3 grid((unsigned)g.n_utt, (unsigned)(pad_l + pad_r));
    mirror_guards_kernel<<<grid, 64, 0, s>>>(reinterpret_cast<__nv_bfloat16*>(hilo), 2 * cpad, rate, pad_l, pad_r, pad_mode, g);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) { if (error) *error = std::string("mirror: ") + cudaGetErrorString(e); return MBEXWN_ERR_CUDA; }
    return MBEXWN_OK;
}

int wn_tc_conv(WnTcState& st, const TcConvArgs& a, const FrameGrid& g, cudaStream_t s, std::string* error) {
    int rc = ensure_impl(st, error);
    if (rc) return rc;
    Impl* im = reinterpret_cast<Impl*>(st.impl);
    im->cta_group = st.cta_group == 2 ? 2 : 1;
    auto fail = [&](const std::string& m, int code) { if (error) *error = m; return code; };
    if (a.cin_pad % TILE_K || a.cout % 8 || a.k * (a.cin_pad / TILE_K) > MAX_KB || a.k > 16)
        return fail("tensor-core conv: unsupported geometry", MBEXWN_ERR_UNSUPPORTED);
    if (a.out_hilo && (a.subpixel < 1 || a.cout % a.subpixel || (a.cout / a.subpixel) % 8))
        return fail("tensor-core conv: sub-pixel unfold needs cout / f to be a multiple of 8", MBEXWN_ERR_UNSUPPORTED);
    GemmParams p{};
    if ((rc = make_map(im, &p.tm_a, a.a_hilo, a.rows, 2 * a.cin_pad, TILE_M, error))) return rc;
    if ((rc = make_map(im, &p.tm_b, a.w, a.cout, 2 * a.k * a.cin_pad, TILE_N, error))) return rc;
    int shifts[16];
    for (int t = 0; t < a.k; ++t) shifts[t] = t * a.dilation - a.pad_l;
    p.n_kb = build_kblocks(p.kb, a.k, shifts, a.cin_pad);
    p.n_terms = 3; p.a_lo_off = a.cin_pad; p.b_lo_off = a.k * a.cin_pad;
    p.rows = a.rows; p.n_cols = a.cout; p.bias = a.bias;
    p.out_f32 = a.out_f32; p.ld_out = a.ld_out;
    p.out_hilo = reinterpret_cast<__nv_bfloat16*>(a.out_hilo); p.out_cpad = a.out_cpad;
    p.subpixel = a.subpixel > 0 ? a.subpixel : 1; p.cout_per = a.cout / p.subpixel;
    p.cv_act = a.act; p.cv_act_mod = a.act_mod > 0 ? a.act_mod : a.cout; p.alpha = a.alpha; p.leaky = a.leaky; p.rate = a.rate;
    p.grid = g;
    cudaError_t e = launch_gemm<EPI_CONV>(im, p, s);
    if (e != cudaSuccess) return fail(std::string("tensor-core conv: ") + cudaGetErrorString(e), MBEXWN_ERR_CUDA);
    return MBEXWN_OK;
}

// Stand-alone GEMM for unit tests: out (rows, n) fp32 = sum_kb A[rows + shift, a_col : a_col+64] @ B[:, b_col : b_col+64]^T
int wn_tc_gemm_test(WnTcState& st, const void* a_bf16, long long rows, int a_cols, const void* b_bf16, int n, int b_cols,
                    const int* kblocks, int n_kb, float* out, cudaStream_t s, std::string* error) {
    int rc = ensure_impl(st, error);
    if (rc) return rc;
    Impl* im = reinterpret_cast<Impl*>(st.impl);
    if (n_kb < 1 || n_kb > MAX_KB) return MBEXWN_ERR_INVALID;
    im->cta_group = st.cta_group == 2 ? 2 : 1;
    GemmParams p{};
    p.n_terms = 1;
    if ((rc = make_map(im, &p.tm_a, a_bf16, rows, a_cols, TILE_M, error))) return rc;
    if ((rc = make_map(im, &p.tm_b, b_bf16, n, b_cols, TILE_N, error))) return rc;
    for (int i = 0; i < n_kb; ++i) p.kb[i] = KBlock{kblocks[3 * i], kblocks[3 * i + 1], kblocks[3 * i + 2]};
    p.n_kb = n_kb; p.rows = rows; p.n_cols = n; p.out_f32 = out; p.bias = nullptr;
    cudaError_t e = launch_gemm<EPI_PLAIN>(im, p, s);
    if (e != cudaSuccess) { if (error) *error = cudaGetErrorString(e); return MBEXWN_ERR_CUDA; }
    return MBEXWN_OK;
}

// Stand-alone split-precision tap-GEMM (unit tests): A (rows, 4 * a_cpad bytes) = [fp16 (a_cpad) | per 64 channels: e4m3 lo8
// (64), e4m3 hi8 (64)], B (n, 4 * b_k bytes) = [fp16 (b_k) | per 64 of K: e4m3 hi8 (64), e4m3 lo8 (64)];
// out = sum_kb A16 @ B16^T + 2^-15 sum_kb (A_lo8 @ B_hi8^T + A_hi8 @ B_lo8^T)
int wn_tc_gemm_test_f16f8(WnTcState& st, const void* a, long long rows, int a_cpad, const void* b, int n, int b_k,
                          const int* kblocks, int n_kb, float* out, cudaStream_t s, std::string* error) {
    int rc = ensure_impl(st, error);
    if (rc) return rc;
    Impl* im = reinterpret_cast<Impl*>(st.impl);
    if (n_kb < 1 || n_kb > MAX_KB) return MBEXWN_ERR_INVALID;
    im->cta_group = st.cta_group == 2 ? 2 : 1;
    GemmParams p{};
    p.n_terms = 2; p.f16 = 1;
    if ((rc = make_map(im, &p.tm_a, a, rows, 2 * a_cpad, TILE_M, error))) return rc;
    if ((rc = make_map(im, &p.tm_b, b, n, 2 * b_k, TILE_N, error))) return rc;
    p.a_lo_off = a_cpad; p.b_lo_off = b_k;
    for (int i = 0; i < n_kb; ++i) p.kb[i] = KBlock{kblocks[3 * i], kblocks[3 * i + 1], kblocks[3 * i + 2]};
    p.n_kb = n_kb; p.rows = rows; p.n_cols = n; p.out_f32 = out; p.bias = nullptr;
    cudaError_t e = launch_gemm<EPI_PLAIN>(im, p, s);
    if (e != cudaSuccess) { if (error) *error = cudaGetErrorString(e); return MBEXWN_ERR_CUDA; }
    return MBEXWN_OK;
}

int wn_tc_launch_ms(WnTcState& st, float* gate_ms, float* resskip_ms, int* n_layers) {
    if (!st.events || st.n_timed < 2) return MBEXWN_ERR_INVALID;
    cudaEvent_t* ev = reinterpret_cast<cudaEvent_t*>(st.events);
    if (cudaEventSynchronize(ev[st.n_timed - 1]) != cudaSuccess) return MBEXWN_ERR_CUDA;
    float g = 0.f, r = 0.f;
    for (int i = 0; i + 1 < st.n_timed; ++i) {
        float ms = 0.f;
        if (cudaEventElapsedTime(&ms, ev[i], ev[i + 1]) != cudaSuccess) return MBEXWN_ERR_CUDA;
        (!st.last_fused && (i & 1) ? r : g) += ms;
    }
    *gate_ms = g; *resskip_ms = r; *n_layers = st.last_fused ? st.n_timed - 1 : (st.n_timed - 1) / 2;
    return MBEXWN_OK;
}

void wn_tc_invalidate(WnTcState&) {}

long long wn_tc_read_trace(WnTcState& st, uint32_t* out, long long n_words) {
    if (!st.trace || !st.impl) return -1;
    const long long have = (long long)(wn_layer_trace_bytes(reinterpret_cast<Impl*>(st.impl)->sm_count) / 4);
    const long long n = n_words < have ? n_words : have;
    if (cudaDeviceSynchronize() != cudaSuccess) return -1;
    if (cudaMemcpy(out, st.trace, (size_t)n * 4, cudaMemcpyDeviceToHost) != cudaSuccess) return -1;
    return n;
}

void wn_tc_destroy(WnTcState& st) {
    if (st.trace) { cudaFree(st.trace); st.trace = nullptr; }
    if (st.events) {
        cudaEvent_t* e = reinterpret_cast<cudaEvent_t*>(st.events);
        for (int i = 0; i < 2 * MBEXWN_MAX_LAYERS + 1; ++i) cudaEventDestroy(e[i]);
        delete[] e;
        st.events = nullptr;
    }
    delete reinterpret_cast<Impl*>(st.impl);
    st.impl = nullptr;
}

}  // namespace mbx
