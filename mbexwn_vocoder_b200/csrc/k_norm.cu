// NormMelComponents.normalize_inputs_by_rms (vocoder/model/wavegen_1d.py:638-769) on the padded frame grid: the optional
// mel-derived RMS normaliser around the generator (PaNWaveNet.infer, wavegen_1d.py:493-512).
//   * norm_rms_kernel     -- raw frame RMS from the log-mel (:663-690): sqrt(sum_k (exp(mell) . P[:, k])^2 / norm_fact), floor,
//                            compressor exponent
//   * norm_smooth_kernel  -- one smoothing iteration (:703-717): gain = OLA(padded rms x smooth window) / OLA(smooth window),
//                            rms <- conv(gain, Hann / sum, stride hop); the sample-rate gain is never materialised, each of its
//                            values is a sum of <= Ws / hop + 1 window taps evaluated in registers
//   * norm_mel_kernel     -- mell <- mel_scale log(exp(mell) / max(eps, rms) lin_scale + lin_off)  (:725-730)
//   * norm_apply_kernel   -- out *= max(gain[win / 2 + n], eps) with the gain of the LAST iteration (:734-736, :504-507)
#include "kernels.cuh"

namespace mbx {
namespace {

constexpr float KERAS_EPS = 1e-7f;

// gain value at utterance-local gain sample n (after the slice offset `off`): T frames, padded = [r0 r0 r... rl rl]
__device__ __forceinline__ float norm_gain_at(const float* __restrict__ rms, int T, int n, const NormArgs& a) {
    const int p = n + a.off;
    int j_hi = p / a.hop;
    int j_lo = (p - a.ws + a.hop) / a.hop;               // ceil((p - ws + 1) / hop) for p - ws + 1 > 0
    if (p - a.ws + 1 <= 0) j_lo = 0;
    if (j_hi > T + 3) j_hi = T + 3;
    float num = 0.f, den = 0.f;
    for (int j = j_lo; j <= j_hi; ++j) {
        const float w = __ldg(a.smooth_win + (p - j * a.hop));
        int t = j - 2;
        t = t < 0 ? 0 : (t >= T ? T - 1 : t);
        num = fmaf(rms[t], w, num);
        den += w;
    }
    return num / fmaxf(KERAS_EPS, den);
}

__global__ void __launch_bounds__(128)
norm_rms_kernel(NormArgs a, FrameGrid g, const float* __restrict__ mell, float* __restrict__ rms_out) {
    __shared__ float mel[256];
    __shared__ float red[4];
    const int f = blockIdx.x, j = threadIdx.x;
    if (g.frame_utt[f] < 0) {
        if (j == 0) rms_out[f] = 0.f;
        return;
    }
    for (int b = j; b < a.n_mel; b += 128) mel[b] = expf(mell[(long long)f * a.n_mel + b]);
    __syncthreads();
    float part = 0.f;
    if (a.proj_cols == 0) {                               // sparse-spectrum assumption: mel * inv_enorm (:690)
        for (int b = j; b < a.n_mel; b += 128) {
            const float v = mel[b] * __ldg(a.proj + b);
            part = fmaf(v, v, part);
        }
    } else {                                              // pseudo-inverse of the mel basis (:687-688)
        for (int k = j; k < a.proj_cols; k += 128) {
            float v = 0.f;
            for (int b = 0; b < a.n_mel; ++b) v = fmaf(mel[b], __ldg(a.proj + (long long)b * a.proj_cols + k), v);
            v *= a.proj_scale;
            part = fmaf(v, v, part);
        }
    }
#pragma unroll
    for (int s = 16; s > 0; s >>= 1) part += __shfl_xor_sync(0xffffffffu, part, s);
    if ((j & 31) == 0) red[j >> 5] = part;
    __syncthreads();
    if (j == 0) {
        float r = sqrtf(((red[0] + red[1]) + (red[2] + red[3])) / a.norm_fact);
        if (a.floor > 0.f) r = fmaxf(r, a.floor);
        if (a.compress_exp != 0.f) r = powf(r, a.compress_exp);
        rms_out[f] = r;
    }
}

__global__ void __launch_bounds__(128)
norm_smooth_kernel(NormArgs a, FrameGrid g, const float* __restrict__ rms_in, float* __restrict__ rms_out) {
    __shared__ float red[4];
    const int f = blockIdx.x, j = threadIdx.x;
    const int u = g.frame_utt[f];
    if (u < 0) {
        if (j == 0) rms_out[f] = 0.f;
        return;
    }
    const int fb = g.utt_begin[u], T = g.utt_end[u] - fb, t = f - fb;
    const float* rms = rms_in + fb;
    float part = 0.f;
    for (int m = j; m < a.win; m += 128) part = fmaf(norm_gain_at(rms, T, t * a.hop + m, a), __ldg(a.gwin + m), part);
#pragma unroll
    for (int s = 16; s > 0; s >>= 1) part += __shfl_xor_sync(0xffffffffu, part, s);
    if ((j & 31) == 0) red[j >> 5] = part;
    __syncthreads();
    if (j == 0) rms_out[f] = (red[0] + red[1]) + (red[2] + red[3]);
}

__global__ void norm_mel_kernel(NormArgs a, FrameGrid g, const float* __restrict__ mell, const float* __restrict__ rms,
                                float* __restrict__ out, long long total) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total) return;
    const int f = (int)(i / a.n_mel);
    if (g.frame_utt[f] < 0) {
        out[i] = 0.f;
        return;
    }
    const float mel = expf(mell[i]) / fmaxf(KERAS_EPS, rms[f]) * a.lin_scale;
    out[i] = a.use_max_limit ? a.mel_scale * logf(fmaxf(mel, a.lin_off)) : a.mel_scale * logf(mel + a.lin_off);
}

__global__ void norm_apply_kernel(NormArgs a, FrameGrid g, const float* __restrict__ rms_prev, float* __restrict__ out,
                                  float* __restrict__ gain_tap, long long total, int out_hop) {
    const long long n = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= total) return;
    const int f = (int)(n / out_hop);
    const int u = g.frame_utt[f];
    if (u < 0) {
        if (gain_tap) gain_tap[n] = 0.f;
        return;
    }
    const int fb = g.utt_begin[u], T = g.utt_end[u] - fb;
    const int local = (int)(n - (long long)fb * out_hop);
    const float gv = fmaxf(norm_gain_at(rms_prev + fb, T, a.win / 2 + local, a), KERAS_EPS);
    if (gain_tap) gain_tap[n] = gv;
    out[n] *= gv;
}

}  // namespace

cudaError_t launch_norm_mel(const NormArgs& a, const FrameGrid& g, const float* mell, float* rms_a, float* rms_b,
                            float* mel_out, const float** rms_prev_out, int* launches, cudaStream_t s) {
    if (g.n_frames <= 0) return cudaSuccess;
    if (a.n_mel > 256 || a.iters < 1 || a.hop <= 0 || a.ws < 1) return cudaErrorInvalidValue;
    norm_rms_kernel<<<g.n_frames, 128, 0, s>>>(a, g, mell, rms_a);
    float *cur = rms_a, *nxt = rms_b;
    const float* prev = rms_a;
    for (int it = 0; it < a.iters; ++it) {
        norm_smooth_kernel<<<g.n_frames, 128, 0, s>>>(a, g, cur, nxt);
        prev = cur;
        float* t = cur; cur = nxt; nxt = t;
    }
    const long long total = (long long)g.n_frames * a.n_mel;
    norm_mel_kernel<<<(unsigned)((total + 255) / 256), 256, 0, s>>>(a, g, mell, cur, mel_out, total);
    *rms_prev_out = prev;
    if (launches) *launches += 2 + a.iters;
    return cudaGetLastError();
}

cudaError_t launch_norm_apply(const NormArgs& a, const FrameGrid& g, const float* rms_prev, float* out, float* gain_tap,
                              int out_hop, cudaStream_t s) {
    const long long total = (long long)g.n_frames * out_hop;
    if (total <= 0) return cudaSuccess;
    norm_apply_kernel<<<(unsigned)((total + 255) / 256), 256, 0, s>>>(a, g, rms_prev, out, gain_tap, total, out_hop);
    return cudaGetLastError();
}

}  // namespace mbx
