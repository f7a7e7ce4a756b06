// tcgen05 / TMEM / TMA WaveNet layer path (definitions in k_wavenet_tc.cu).
#pragma once
#include <functional>
#include <string>

#include "../../include/mbexwn.h"
#include "common.cuh"

namespace mbx {

struct WnTcState {
    bool ready = false;
    int cta_group = 2;          // option "tc_cta_group": 1 = one CTA per tile, 2 = CTA pairs (cta_group::2)
    int cond_stage = 1;         // option "tc_cond_stage": gate epilogue reads its conditioning rows from a smem stage
    // MBEXWN_PREC_F16F8: log2 scales of the e4m3 lo8 planes ((x - fp16 x) * 2^sh) of the residual stream (h) and of the
    // gated activations (a); the hi8 planes are e4m3(x) unscaled.  The weight hi8 planes carry 2^(15 - sh_*), the weight
    // lo8 planes 2^15: options "tc8_h_lo", "tc8_a_lo" (the host packer must agree).
    int sh_h_lo = 9, sh_a_lo = 10;
    // option "stage_timing": CUDA events around every WaveNet GEMM launch of the last forward (gate / res-skip alternate)
    int time_launches = 0;
    int n_timed = 0;            // events recorded by the last forward (n launches + 1)
    void* events = nullptr;     // cudaEvent_t[2 * MBEXWN_MAX_LAYERS + 1]
    int debug = 0;              // option "tc_debug": timing experiments (results are wrong), see GemmParams::debug
    // option "tc_fused": one persistent kernel per WaveNet layer (k_wavenet_layer.cu) instead of a gate and a res/skip launch:
    // 0 = never, 1 (default) = whenever the geometry allows, 2 = only when every CTA pair gets at least two 256-row M tiles.
    // The two paths sum the same products in a different K order (agreement ~1e-5 of peak), so a batch-size dependent
    // choice would make an utterance's samples depend on the batch it travels in: the default does not switch.
    int fused = 1;
    // MBEXWN_PREC_F16F8 range guard: kernels set bit 0 of this (host-mapped) word when a residual-stream value leaves the range of
    // the unscaled e4m3 hi8 plane (|x| > 448: the correction product would silently lose its meaning), bit 1 beyond 60000 (fp16)
    int* range_flag = nullptr;
    int n_a = 3;                // option "tc_ring_a": A slab ring slots of the fused kernel (2 or 3); the B ring takes the rest
    int slab = 1;               // option "tc_slab": the dilated taps of the fused kernel share one A slab per 64-channel block (0: one
                                // A tile per tap)
    int last_fused = 0;         // the last forward ran the fused kernel
    void* trace = nullptr;      // option "tc_trace": device buffer of per-tile cycle stamps of the last fused launch
    int trace_on = 0;
    int interleave = 1;         // option "tc_interleave": the res tiles of M tile j - 1 run behind the FIRST gate tile of M tile j (0: behind the
                                // last): with "tc_discard" 2.4 instead of 3.6 GB of DRAM traffic per layer at config 2, ~1 % faster
    int discard = 1;            // option "tc_discard": rows of the activation scratch are discarded from the L2 (no write-back) once the res tiles have read them
    int cluster = 2;            // option "tc_cluster": CTAs per cluster of the fused kernel, 2 (one pair) or 4 (two pairs share the B tiles by TMA multicast)
    int cluster_min_sms = 128;  // clusters of 4 only when the device keeps at least this many SMs busy with them
    int last_quads = 0;         // clusters of 4 the device holds at once (cudaOccupancyMaxActiveClusters), -1: query failed
    int last_cluster = 2;       // cluster size of the last fused launch
    void* impl = nullptr;
};

// ---- fused per-layer kernel (k_wavenet_layer.cu) ----------------------------------------------------------------
struct WnLayerArgs {
    const void* h_in;           // (rows, 2 cpad) operand planes of the layer input
    void* h_out;                // same geometry, or nullptr for the last layer (no residual output)
    void* scratch;              // wn_layer_scratch_bytes: per-pair double-buffered gated activations
    const void* w1;             // (n1, k1) packed dilated-conv weights
    const void* w2;             // (n2, k2) packed res / skip weights
    int n1, k1, n2, k2;
    long long rows;
    int c, cpad, n_terms, n_taps;
    int shifts[16];
    const float* bias1;
    const float* cond;
    long long cond_total;
    int cond_rows, lin_up, gate, steps_per_frame;
    float lin_w0[32], lin_w1[32];
    float act_lo_scale, h_lo_scale;
    const float* bias2;
    float* skip;
    int skip_ld, skip_c, res_cols, first;
    FrameGrid grid;
    int sm_count;
    int* range_flag;            // WnTcState::range_flag
    void* trace;                // nullptr, or wn_layer_trace_bytes of device memory
};
size_t wn_layer_scratch_bytes(int cpad, int sm_count);
size_t wn_layer_trace_bytes(int sm_count);
bool wn_layer_supported(const mbexwn_config_t& c, int cpad, int n_terms, int cond_rows);
int wn_layer_forward(WnTcState& st, const WnLayerArgs& a, cudaStream_t s, std::string* error);
// cuTensorMapEncodeTiled of a row-major 16-bit matrix (rows, cols) with a 64-column x box_rows box, SWIZZLE_128B
int wn_tc_encode_map(WnTcState& st, void* tensor_map, const void* base, long long rows, long long cols, int box_rows,
                     std::string* error);
int wn_tc_sm_count(WnTcState& st);
// copies the trace of the last fused launch to the host (n_words uint32); returns the number of words written or < 0
long long wn_tc_read_trace(WnTcState& st, uint32_t* out, long long n_words);

// workspace slots the tensor-core path needs (called from the workspace carver)
void wn_tc_carve(const mbexwn_config_t& c, long long rows, int precision,
                 const std::function<void(const char*, size_t)>& add);

// Channels of the WaveNet output buffer (c_out rounded up to 32; the padding columns are written as zeros).
int wn_tc_out_pad(const mbexwn_config_t& c);

// Runs start conv + all WaveNet layers; leaves end(skip sum) = the WaveNet output (rows, wn_tc_out_pad) fp32 in
// `wn_out` (the linear `end` 1x1 is folded into the skip half of every res_skip matrix at pack time).
// `wn_in` rows hold wn_cin channels with a pitch of `ld_in` floats (a later block of a stack reads the previous block's
// output in place, whose pitch is wn_tc_out_pad).
int wn_tc_forward(WnTcState& st, const mbexwn_config_t& c, const FrameGrid& g, int precision, const float* wn_in, int ld_in,
                  const float* cond, float* wn_out, const std::function<void*(const char*)>& slot,
                  const std::function<const void*(const std::string&, size_t)>& tensor, cudaStream_t s, int* launches,
                  std::string* error);

// One conv of a mel-rate sub-net on the tensor cores (3-product bf16 split, fp32 accumulate): out[r, :] =
// act(sum_t A[r + t * dilation - pad_l, :] @ W[t] + bias), rows outside an utterance read as zero (or as the mirrored
// rows wn_tc_pack / wn_tc_mirror put into the guard rows of A).
struct TcConvArgs {
    const void* a_hilo;         // (rows, 2 * cin_pad) bf16 [hi | lo]
    long long rows;
    int cin_pad;                // multiple of 64
    const void* w;              // (cout, 2 * k * cin_pad) bf16 [hi | lo], (tap, cin)-major
    int cout, k, dilation, pad_l;
    const float* bias;          // (cout)
    int act;                    // Act: none / PReLU / LeakyReLU
    const float* alpha;         // PReLU slopes, index n % act_mod
    int act_mod;
    float leaky;
    int rate;                   // rows per frame of A
    float* out_f32;             // optional (rows, ld_out)
    int ld_out;
    void* out_hilo;             // optional (rows * subpixel, 2 * out_cpad) bf16 [hi | lo]
    int out_cpad, subpixel;
};
int wn_tc_conv(WnTcState& st, const TcConvArgs& a, const FrameGrid& g, cudaStream_t s, std::string* error);
// fp32 (rows, c) -> bf16 [hi | lo] (rows, 2 * cpad) with the reference's pad values in the guard rows next to utterances
int wn_tc_pack(const float* x, void* out_hilo, long long rows, int c, int cpad, int rate, int pad_l, int pad_r, int pad_mode,
               const FrameGrid& g, cudaStream_t s, std::string* error);
// copy mirrored / replicated edge rows into the guard rows of a [hi | lo] buffer (SYMMETRIC / EDGE padding of the next conv)
int wn_tc_mirror(void* hilo, long long rows, int cpad, int rate, int pad_l, int pad_r, int pad_mode, const FrameGrid& g,
                 cudaStream_t s, std::string* error);

// Stand-alone tap-GEMM (unit tests): out (rows, n) fp32 = sum over K blocks {a_col, a_row_shift, b_col} of
// A[rows + shift, a_col : a_col + 64] @ B[:, b_col : b_col + 64]^T, A (rows, a_cols) bf16, B (n, b_cols) bf16.
int wn_tc_gemm_test(WnTcState& st, const void* a_bf16, long long rows, int a_cols, const void* b_bf16, int n, int b_cols,
                    const int* kblocks, int n_kb, float* out, cudaStream_t s, std::string* error);

// Split-precision variant (MBEXWN_PREC_F16F8 operand format, see include/mbexwn.h): fp16 main product plus two e4m3
// correction products scaled by 2^-15.
int wn_tc_gemm_test_f16f8(WnTcState& st, const void* a, long long rows, int a_cpad, const void* b, int n, int b_k,
                          const int* kblocks, int n_kb, float* out, cudaStream_t s, std::string* error);

// Device time of the gate and res/skip GEMM launches of the last forward (needs time_launches); synchronises.
int wn_tc_launch_ms(WnTcState& st, float* gate_ms, float* resskip_ms, int* n_layers);

void wn_tc_invalidate(WnTcState& st);
void wn_tc_destroy(WnTcState& st);

}  // namespace mbx
