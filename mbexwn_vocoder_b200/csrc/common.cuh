// Shared device/host helpers for the MBExWN sm_100a kernels.
//
// Data layout ("padded frame grid"): all utterances of a batch are concatenated in time with `halo`
// guard frames before the first, between neighbours and after the last utterance.  Every buffer at every
// internal rate (mel 1x, cond 2x, WaveNet 20x, pulse 100x, audio 300x rows per mel frame) uses the same
// frame grid, so row r at rate R belongs to padded frame r / R.  frame_utt[f] is the utterance id of
// padded frame f or -1 for a guard frame; utt_begin/utt_end are the [begin, end) padded-frame range of each
// utterance.  Guard rows of activation buffers are kept at zero, which gives the reference's per-utterance
// "SAME" zero padding for free in the tensor-core path (TMA reads zeros) and makes every
// row independent of its tile's position.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace mbx {

struct FrameGrid {
    const int32_t* frame_utt;   // [n_frames]
    const int32_t* utt_begin;   // [n_utt]
    const int32_t* utt_end;     // [n_utt]
    int32_t n_frames;
    int32_t n_utt;
};

enum PadMode { PAD_ZERO = 0, PAD_SYMMETRIC = 1, PAD_EDGE = 2 };
enum Act { ACT_NONE = 0, ACT_PRELU = 1, ACT_LEAKY = 2, ACT_SOFT_SIGMOID_AFFINE = 3, ACT_TANH = 4, ACT_SIGMOID = 5 };
enum Gate { GATE_GTU = 0, GATE_GLU = 1, GATE_GFU = 2, GATE_GSU = 3 };

// Row bounds [lo, hi) at `rate` rows per frame of the utterance owning row r; false for guard rows.
__device__ __forceinline__ bool utt_bounds(const FrameGrid& g, int rate, long long r, long long& lo, long long& hi) {
    int f = (int)(r / rate);
    if (f < 0 || f >= g.n_frames) return false;
    int u = g.frame_utt[f];
    if (u < 0) return false;
    lo = (long long)g.utt_begin[u] * rate;
    hi = (long long)g.utt_end[u] * rate;
    return true;
}

// Source row for tap position `s` (may lie outside [lo, hi)) under the reference's padding modes
// (custom_layers.py:47-71; tf.pad SYMMETRIC mirrors including the edge sample). Returns -1 for a zero tap.
__device__ __forceinline__ long long pad_index(long long s, long long lo, long long hi, int mode) {
    if (s >= lo && s < hi) return s;
    if (mode == PAD_ZERO) return -1;
    if (mode == PAD_EDGE) return s < lo ? lo : hi - 1;
    // symmetric
    long long r = s < lo ? (2 * lo - 1 - s) : (2 * hi - 1 - s);
    if (r < lo) r = lo;
    if (r >= hi) r = hi - 1;
    return r;
}

__device__ __forceinline__ float apply_act(float v, int act, float alpha, float a0, float a1) {
    switch (act) {
        case ACT_PRELU:
        case ACT_LEAKY: return v >= 0.f ? v : __fmul_rn(alpha, v);
        case ACT_SOFT_SIGMOID_AFFINE: {
            // custom_AE_layers.py:99: 0.5 + 0.5 * x / (1 + |x|); then F0 = y * (fmax - fmin) + fmin
            // (custom_pulsed_generator.py:786).  Same operation order as the reference, no FMA contraction.
            float y = __fadd_rn(0.5f, __fdiv_rn(__fmul_rn(0.5f, v), __fadd_rn(1.f, fabsf(v))));
            return __fadd_rn(__fmul_rn(y, a0), a1);
        }
        case ACT_TANH: return tanhf(v);
        case ACT_SIGMOID: return 1.f / (1.f + expf(-v));
        default: return v;
    }
}

}  // namespace mbx

#define MBX_RC(expr)                                                                           \
    do {                                                                                       \
        int _rc = (expr);                                                                      \
        if (_rc) return _rc;                                                                   \
    } while (0)

#define MBX_CUDA_CHECK(expr)                                                                   \
    do {                                                                                       \
        cudaError_t _e = (expr);                                                               \
        if (_e != cudaSuccess) return mbx::set_error(h, #expr, _e);                            \
    } while (0)
