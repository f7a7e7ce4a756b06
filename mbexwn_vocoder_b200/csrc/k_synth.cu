// Output stage of the MBExWN forward path:
//   * pqmf_kernel        -- TFPQMF.synthesis in polyphase form (tf_preprocess.py:208-226)
//   * stft_filter_kernel -- per mel frame: periodic-Hann STFT of the excitation, vocal-tract filter from the
//                           cepstrum (lifter -> rfft -> exp(r tanh(Re) + i Im)), product, inverse FFT, dual window
//                           (custom_pulsed_generator.py:681-724, :793-836, :507-525)
//   * ola_kernel         -- overlap-add of the synthesis frames, crop to [win/2, win/2 + T*hop) (:716-724)
// One complex FFT of size N carries both real inputs (windowed excitation frame, zero-padded cepstrum); the two
// spectra are separated by symmetry, so each frame costs one forward and one inverse N-point FFT in shared memory.
#include "kernels.cuh"

namespace mbx {

namespace {

constexpr int PQ_ROWS = 64;      // sub-band rows (= blocks of S output samples) per CTA
constexpr int PQ_THREADS = 256;

__global__ void __launch_bounds__(PQ_THREADS)
pqmf_kernel(PqmfArgs a, FrameGrid g) {
    extern __shared__ float sm[];
    float* G = sm;                                   // (Q, S, S)
    float* X = sm + a.Q * a.S * a.S;                 // (PQ_ROWS + Q - 1, S)
    const long long m0 = (long long)blockIdx.x * PQ_ROWS;
    const int tile_rows = PQ_ROWS + a.Q - 1;
    for (int i = threadIdx.x; i < a.Q * a.S * a.S; i += PQ_THREADS) G[i] = a.poly[i];
    // bounds of the utterance are evaluated per row: rows outside the owning utterance contribute zeros
    for (int i = threadIdx.x; i < tile_rows * a.S; i += PQ_THREADS) {
        int rl = i / a.S, k = i - rl * a.S;
        long long r = m0 + rl - a.back;
        float v = 0.f;
        if (r >= 0 && r < a.rows && g.frame_utt[r / a.steps_per_frame] >= 0) v = a.sub[r * a.S + k];
        X[i] = v;
    }
    __syncthreads();
    for (int o = threadIdx.x; o < PQ_ROWS * a.S; o += PQ_THREADS) {
        int ml = o / a.S, p = o - ml * a.S;
        long long m = m0 + ml;
        if (m >= a.rows) continue;
        int fu = g.frame_utt[m / a.steps_per_frame];
        float acc = 0.f;
        if (fu >= 0) {
            const long long lo = (long long)g.utt_begin[fu] * a.steps_per_frame;
            const long long hi = (long long)g.utt_end[fu] * a.steps_per_frame;
            for (int q = 0; q < a.Q; ++q) {
                long long r = m + q - a.back;
                if (r < lo || r >= hi) continue;     // zero padding at the utterance's own ends
                const float* xr = X + (ml + q) * a.S;
                const float* gq = G + (q * a.S) * a.S + p;
#pragma unroll 5
                for (int k = 0; k < a.S; ++k) acc = fmaf(xr[k], gq[k * a.S], acc);
            }
        }
        a.out[m * a.S + p] = acc;
    }
}

// ---- shared-memory FFT ----------------------------------------------------------------------------

__device__ __forceinline__ float2 cmul(float2 x, float2 y) {
    return make_float2(x.x * y.x - x.y * y.y, x.x * y.y + x.y * y.x);
}

// In-place radix-2 decimation-in-time FFT over `s` (length n = 1 << logn) whose input is already in
// bit-reversed order.  tw[j] = exp(-2 pi i j / n), j < n/2.
__device__ void fft_inplace(float2* s, const float2* __restrict__ tw, int n, int logn) {
    for (int st = 1; st <= logn; ++st) {
        const int half = 1 << (st - 1);
        const int tstride = n >> st;
        for (int b = threadIdx.x; b < n / 2; b += blockDim.x) {
            int j = b & (half - 1);
            int i0 = ((b >> (st - 1)) << st) + j;
            int i1 = i0 + half;
            float2 w = tw[j * tstride];
            float2 u = s[i0];
            float2 v = cmul(s[i1], w);
            s[i0] = make_float2(u.x + v.x, u.y + v.y);
            s[i1] = make_float2(u.x - v.x, u.y - v.y);
        }
        __syncthreads();
    }
}

__global__ void __launch_bounds__(256)
stft_filter_kernel(StftFilterArgs a, FrameGrid g) {
    extern __shared__ float2 smem2[];
    float2* A = smem2;                 // forward transform buffer
    float2* Bf = smem2 + a.fft;        // inverse transform buffer
    __shared__ float red[256];
    __shared__ int lifter_row;

    const int f = blockIdx.x;
    const int u = g.frame_utt[f];
    const int tid = threadIdx.x;
    const int N = a.fft;
    int logn = 0;
    while ((1 << logn) < N) ++logn;
    float* frame_out = a.frames_out + (long long)f * a.win;
    if (u < 0) {                                             // guard frame: keep the frame buffer defined
        for (int n = tid; n < a.win; n += blockDim.x) frame_out[n] = 0.f;
        if (a.lifter_index_out && tid == 0) a.lifter_index_out[f] = 0;
        return;
    }
    const long long s_lo = (long long)g.utt_begin[u] * a.hop, s_hi = (long long)g.utt_end[u] * a.hop;

    // ---- cepstral lifter selection (custom_pulsed_generator.py:507-525) ----
    if (a.lifters != nullptr) {
        const long long p_lo = (long long)g.utt_begin[u] * a.pulse_per_frame;
        const long long p_hi = (long long)g.utt_end[u] * a.pulse_per_frame;
        const long long start = (long long)f * a.pulse_per_frame - a.n_smooth / 2;
        float part = 0.f;
        for (int i = tid; i < a.n_smooth; i += blockDim.x) {
            long long j = start + i;
            j = j < p_lo ? p_lo : (j >= p_hi ? p_hi - 1 : j);
            part = fmaf(a.f0_smooth[i], a.f0[j], part);
        }
        red[tid] = part;
        __syncthreads();
        for (int sft = 128; sft > 0; sft >>= 1) {
            if (tid < sft) red[tid] += red[tid + sft];
            __syncthreads();
        }
        if (tid == 0) {
            float lo10 = a.lifter_grid[0], hi10 = a.lifter_grid[a.n_lift - 1];
            float l10 = __fmul_rn(0.43429448190325176f, logf(red[0]));
            l10 = fminf(fmaxf(l10, lo10), hi10);
            float ratio = __fdiv_rn(__fsub_rn(l10, lo10), __fsub_rn(hi10, lo10));
            int idx = (int)rintf(__fmul_rn(ratio, (float)(a.n_lift - 1)));   // round half to even like tf.round
            lifter_row = idx;
            if (a.lifter_index_out) a.lifter_index_out[f] = idx;
        }
        __syncthreads();
    }

    // ---- pack: real = windowed excitation frame, imag = zero-padded cepstrum without c0 ----
    const float* ceps = a.ceps + (long long)f * a.n_ceps;
    const float* lift = a.lifters ? a.lifters + (long long)lifter_row * a.n_ceps : nullptr;
    const long long x0 = (long long)f * a.hop - a.win / 2;      // first excitation sample of this frame
    for (int n = tid; n < N; n += blockDim.x) {
        float re = 0.f, im = 0.f;
        if (n < a.win) {
            long long sidx = x0 + n;
            if (sidx >= s_lo && sidx < s_hi) re = a.exc[sidx] * a.window[n];
        }
        if (n >= 1 && n < a.n_ceps) im = lift ? ceps[n] * lift[n] : ceps[n];
        A[__brev((unsigned)n) >> (32 - logn)] = make_float2(re, im);
    }
    __syncthreads();
    fft_inplace(A, a.twiddle, N, logn);

    // ---- separate the two spectra, apply the vocal-tract filter, build the conjugate Hermitian spectrum ----
    float2* vtf_out = a.vtf_out ? reinterpret_cast<float2*>(a.vtf_out) + (long long)f * (N / 2 + 1) : nullptr;
    for (int k = tid; k <= N / 2; k += blockDim.x) {
        float2 zk = A[k];
        float2 zn = A[(N - k) & (N - 1)];
        float2 X = make_float2(0.5f * (zk.x + zn.x), 0.5f * (zk.y - zn.y));        // spectrum of the real part
        float2 L = make_float2(0.5f * (zk.y + zn.y), 0.5f * (zn.x - zk.x));        // spectrum of the imag part
        float mag = a.max_log_range > 0.f ? expf(a.max_log_range * tanhf(L.x)) : expf(L.x);
        float sn, cs;
        sincosf(L.y, &sn, &cs);
        float2 V = make_float2(mag * cs, mag * sn);
        if (vtf_out) vtf_out[k] = V;
        float2 Y = cmul(X, V);
        // inverse transform as real(fft(conj(W))) with W Hermitian: W[k] = Y, W[N-k] = conj(Y)
        Bf[__brev((unsigned)k) >> (32 - logn)] = make_float2(Y.x, -Y.y);
        if (k > 0 && k < N / 2) Bf[__brev((unsigned)(N - k)) >> (32 - logn)] = Y;
    }
    __syncthreads();
    fft_inplace(Bf, a.twiddle, N, logn);
    const float scale = 1.f / (float)N;
    for (int n = tid; n < a.win; n += blockDim.x) frame_out[n] = Bf[n].x * scale * a.inv_window[n];
}

__global__ void ola_kernel(OlaArgs a, FrameGrid g) {
    const long long n = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= (long long)a.n_frames * a.hop) return;
    const int f = (int)(n / a.hop);
    const int u = g.frame_utt[f];
    float acc = 0.f;
    if (u >= 0) {
        const int fb = g.utt_begin[u], fe = g.utt_end[u];
        const int half = a.win / 2;
        const long long t = n - (long long)fb * a.hop;          // utterance-local output sample
        // frame j covers local samples [j*hop - half, j*hop - half + win)
        long long j_hi = (t + half) / a.hop;
        long long j_lo = (t + half - a.win) / a.hop + 1;
        if (t + half - a.win < 0) j_lo = 0;
        if (j_hi > fe - fb - 1) j_hi = fe - fb - 1;
        for (long long j = j_lo; j <= j_hi; ++j) {
            long long pos = t + half - j * a.hop;
            acc += a.frames[(fb + j) * (long long)a.win + pos];
        }
    }
    a.out[n] = acc;
}

}  // namespace

cudaError_t launch_pqmf(const PqmfArgs& a, const FrameGrid& g, cudaStream_t s) {
    if (a.rows <= 0) return cudaSuccess;
    size_t smem = (size_t)(a.Q * a.S * a.S + (PQ_ROWS + a.Q - 1) * a.S) * sizeof(float);
    if (smem > 48 * 1024) {
        cudaError_t e = cudaFuncSetAttribute(pqmf_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
    }
    pqmf_kernel<<<(unsigned)((a.rows + PQ_ROWS - 1) / PQ_ROWS), PQ_THREADS, smem, s>>>(a, g);
    return cudaGetLastError();
}

cudaError_t launch_stft_filter(const StftFilterArgs& a, const FrameGrid& g, cudaStream_t s) {
    if (a.n_frames <= 0) return cudaSuccess;
    if (a.fft & (a.fft - 1)) return cudaErrorInvalidValue;
    size_t smem = (size_t)2 * a.fft * sizeof(float2);
    if (smem > 48 * 1024) {
        cudaError_t e = cudaFuncSetAttribute(stft_filter_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
    }
    stft_filter_kernel<<<a.n_frames, 256, smem, s>>>(a, g);
    return cudaGetLastError();
}

cudaError_t launch_ola(const OlaArgs& a, const FrameGrid& g, cudaStream_t s) {
    long long total = (long long)a.n_frames * a.hop;
    if (total <= 0) return cudaSuccess;
    ola_kernel<<<(unsigned)((total + 255) / 256), 256, 0, s>>>(a, g);
    return cudaGetLastError();
}

}  // namespace mbx
