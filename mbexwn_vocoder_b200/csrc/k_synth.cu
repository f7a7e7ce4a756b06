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

// ---- fused post 1x1 + PQMF synthesis -------------------------------------------------------------------------------
// One CTA = PP_ROWS sub-band rows (= PP_ROWS * S output samples).  The WaveNet output rows of the tile (with the Q - 1
// halo rows the polyphase filter reaches) are staged coalesced in shared memory, the wn_post_net 1x1
// (custom_pulsed_generator.py:913-914) turns them into the sub-band tile, then thread = row accumulates its S output
// samples over the (Q, S) taps with the polyphase matrix read as warp-wide broadcasts (one LDS.128 feeds four FMAs).
// Same summation order as conv1d_kernel + pqmf_kernel, so the results are bit-identical to the two-kernel path.
constexpr int PP_ROWS = 128, PP_MAXS = 16, PP_MAXC = 32, PP_MAXQ = 24;

__global__ void __launch_bounds__(PP_ROWS)
post_pqmf_kernel(PostPqmfArgs a, FrameGrid g) {
    extern __shared__ float sm[];
    const int S = a.S, Q = a.Q, tile_rows = PP_ROWS + Q - 1;
    float* G = sm;                                        // (Q * S, 16): polyphase taps, phase-padded to 16
    float* Wp = G + Q * S * PP_MAXS;                      // (cin + 1, 16): post weights, then bias
    float* Xw = Wp + (a.cin + 1) * PP_MAXS;               // (tile_rows, ld + 1) staged WaveNet output rows
    float* X = Xw + tile_rows * (a.ld + 1);               // (tile_rows, S) sub-band tile; reused for the output samples
    __shared__ int s_fu[PP_ROWS + PP_MAXQ];               // utterance of every tile row, -1 for guard / out-of-range rows
    const int tid = threadIdx.x;
    const long long m0 = (long long)blockIdx.x * PP_ROWS;
    const long long r0 = m0 - a.back;                     // first staged row

    for (int i = tid; i < Q * S * PP_MAXS; i += PP_ROWS) {
        const int qk = i >> 4, p = i & 15;
        G[i] = p < S ? a.poly[qk * S + p] : 0.f;
    }
    for (int i = tid; i < (a.cin + 1) * PP_MAXS; i += PP_ROWS) {
        const int ci = i >> 4, p = i & 15;
        Wp[i] = p < S ? (ci < a.cin ? a.post_w[ci * S + p] : a.post_b[p]) : 0.f;
    }
    for (int i = tid; i < tile_rows; i += PP_ROWS) {
        const long long r = r0 + i;
        s_fu[i] = (r >= 0 && r < a.rows) ? g.frame_utt[r / a.steps_per_frame] : -1;
    }
    __syncthreads();                                      // the copy below skips guard rows by s_fu
    {   // coalesced float4 copy of the contiguous (tile_rows, ld) block
        const int n4 = tile_rows * a.ld / 4;
        const float4* src = reinterpret_cast<const float4*>(a.wn_out + r0 * a.ld);
        for (int i = tid; i < n4; i += PP_ROWS) {
            const int e = i * 4, rl = e / a.ld, c = e - rl * a.ld;
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            // guard rows of the WaveNet output are never written by the res/skip epilogue: do not read them
            if (s_fu[rl] >= 0) v = __ldg(src + i);
            float* d = Xw + rl * (a.ld + 1) + c;
            d[0] = v.x; d[1] = v.y; d[2] = v.z; d[3] = v.w;
        }
    }
    __syncthreads();
    // post 1x1: sub-bands of every staged row (guard rows are zero)
    for (int rl = tid; rl < tile_rows; rl += PP_ROWS) {
        float acc[PP_MAXS];
#pragma unroll
        for (int p = 0; p < PP_MAXS; ++p) acc[p] = 0.f;
        const bool valid = s_fu[rl] >= 0;
        if (valid) {
            const float* xr = Xw + rl * (a.ld + 1);
            for (int ci = 0; ci < a.cin; ++ci) {
                const float x = xr[ci];
                const float4* w4 = reinterpret_cast<const float4*>(Wp + ci * PP_MAXS);
#pragma unroll
                for (int v = 0; v < 4; ++v) {
                    const float4 w = w4[v];
                    acc[4 * v] = fmaf(x, w.x, acc[4 * v]); acc[4 * v + 1] = fmaf(x, w.y, acc[4 * v + 1]);
                    acc[4 * v + 2] = fmaf(x, w.z, acc[4 * v + 2]); acc[4 * v + 3] = fmaf(x, w.w, acc[4 * v + 3]);
                }
            }
            const float* b = Wp + a.cin * PP_MAXS;
#pragma unroll
            for (int p = 0; p < PP_MAXS; ++p) acc[p] += b[p];
            if (a.log_gain) {
                // multi-band gain: value (r - first row of the utterance) of the x gain_up interpolation of exp(log gain)
                const int fu = s_fu[rl];
                const int fb = g.utt_begin[fu], T = g.utt_end[fu] - fb;
                const long long rloc = (r0 + rl) - (long long)fb * a.steps_per_frame;
                const int t = (int)(rloc / a.gain_up), u = (int)(rloc - (long long)t * a.gain_up);
                const int t1 = t + 1 < T ? t + 1 : T - 1;
                const float* l0 = a.log_gain + (long long)(fb + t) * S;
                const float* l1 = a.log_gain + (long long)(fb + t1) * S;
                float m0g = 0.f, m1g = 0.f;
                if (a.gain_center) {
                    for (int p = 0; p < S; ++p) { m0g += l0[p]; m1g += l1[p]; }
                    m0g /= (float)S; m1g /= (float)S;
                }
                const float w0 = (float)((double)(a.gain_up - u) / a.gain_up), w1 = (float)((double)u / a.gain_up);
#pragma unroll
                for (int p = 0; p < PP_MAXS; ++p)
                    if (p < S) acc[p] *= __fadd_rn(__fmul_rn(expf(l0[p] - m0g), w0), __fmul_rn(expf(l1[p] - m1g), w1));
            }
        }
        const long long r = r0 + rl;
#pragma unroll
        for (int p = 0; p < PP_MAXS; ++p)
            if (p < S) {
                X[rl * S + p] = acc[p];
                // the tile's own rows (not the halo) also go to the sub-band tap
                if (a.sub_out && rl >= a.back && rl < a.back + PP_ROWS && r < a.rows) a.sub_out[r * S + p] = acc[p];
            }
    }
    __syncthreads();
    // polyphase synthesis: thread = row m0 + tid
    float acc[PP_MAXS];
#pragma unroll
    for (int p = 0; p < PP_MAXS; ++p) acc[p] = 0.f;
    const long long m = m0 + tid;
    const int fu = s_fu[tid + a.back];
    if (m < a.rows && fu >= 0) {
        const long long lo = (long long)g.utt_begin[fu] * a.steps_per_frame;
        const long long hi = (long long)g.utt_end[fu] * a.steps_per_frame;
        for (int q = 0; q < Q; ++q) {
            const long long r = m + q - a.back;
            if (r < lo || r >= hi) continue;             // zero padding at the utterance's own ends
            const float* xr = X + (tid + q) * S;
            const float* gq = G + q * S * PP_MAXS;
            for (int k = 0; k < S; ++k) {
                const float x = xr[k];
                const float4* g4 = reinterpret_cast<const float4*>(gq + k * PP_MAXS);
#pragma unroll
                for (int v = 0; v < 4; ++v) {
                    const float4 w = g4[v];
                    acc[4 * v] = fmaf(x, w.x, acc[4 * v]); acc[4 * v + 1] = fmaf(x, w.y, acc[4 * v + 1]);
                    acc[4 * v + 2] = fmaf(x, w.z, acc[4 * v + 2]); acc[4 * v + 3] = fmaf(x, w.w, acc[4 * v + 3]);
                }
            }
        }
    }
    __syncthreads();                                      // everyone is done reading X
#pragma unroll
    for (int p = 0; p < PP_MAXS; ++p)
        if (p < S) X[tid * S + p] = acc[p];
    __syncthreads();
    // coalesced store of the PP_ROWS * S contiguous output samples
    const long long o0 = m0 * S;
    const long long total = a.rows * S;
    for (int i = tid * 4; i < PP_ROWS * S; i += PP_ROWS * 4) {
        if (o0 + i + 3 < total && ((o0 & 3) == 0)) {
            *reinterpret_cast<float4*>(a.out + o0 + i) = make_float4(X[i], X[i + 1], X[i + 2], X[i + 3]);
        } else {
            for (int e = 0; e < 4; ++e)
                if (i + e < PP_ROWS * S && o0 + i + e < total) a.out[o0 + i + e] = X[i + e];
        }
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

// ---- register-resident 2048-point FFT -----------------------------------------------------------------------------
// Stockham autosort, radices 16 x 16 x 8, 128 threads per frame: every thread holds 16 complex points in registers, the
// three passes exchange data through one 17 KB shared buffer (pass 1 -> 2 with one pad element per 16 so that the
// stride-16 writes are conflict free; the later layouts are linear).  Pass p with radix R and Ns = product of the earlier
// radices: butterfly j reads x[j + t N/R], multiplies by exp(-2 pi i t (j mod Ns) / (Ns R)), does an R-point DFT and writes
// y[(j - j mod Ns) R + j mod Ns + t Ns]; after the last pass the spectrum is in natural order.
constexpr int FN = 2048, FT = 128;
constexpr int F_S1 = FN + FN / 16;          // padded pass-1 -> pass-2 exchange

__device__ __forceinline__ float2 cadd(float2 a, float2 b) { return make_float2(a.x + b.x, a.y + b.y); }
__device__ __forceinline__ float2 csub(float2 a, float2 b) { return make_float2(a.x - b.x, a.y - b.y); }

__device__ __forceinline__ void dft4(float2& a0, float2& a1, float2& a2, float2& a3) {
    const float2 t0 = cadd(a0, a2), t1 = csub(a0, a2), t2 = cadd(a1, a3), d = csub(a1, a3);
    const float2 t3 = make_float2(d.y, -d.x);                  // -i (a1 - a3)
    a0 = cadd(t0, t2); a1 = cadd(t1, t3); a2 = csub(t0, t2); a3 = csub(t1, t3);
}

// forward DFT of v[0..15] in place (natural order in, natural order out)
__device__ __forceinline__ void dft16(float2 (&v)[16]) {
    constexpr float C1 = 0.92387953251128674f, S1 = 0.38268343236508977f, R2 = 0.70710678118654752f;
    // step 1: for every n2 a 4-point DFT over n1 of v[4 n1 + n2]; result u[n2][k1] stays in v[4 k1 + n2]
#pragma unroll
    for (int n2 = 0; n2 < 4; ++n2) dft4(v[n2], v[4 + n2], v[8 + n2], v[12 + n2]);
    // step 2: u[n2][k1] *= W16^(n2 k1)
    const float2 w1 = make_float2(C1, -S1), w2 = make_float2(R2, -R2), w3 = make_float2(S1, -C1);
    const float2 w6 = make_float2(-R2, -R2), w9 = make_float2(-C1, S1);
    v[4 + 1] = cmul(v[4 + 1], w1); v[4 + 2] = cmul(v[4 + 2], w2); v[4 + 3] = cmul(v[4 + 3], w3);
    v[8 + 1] = cmul(v[8 + 1], w2); v[8 + 2] = make_float2(v[8 + 2].y, -v[8 + 2].x); v[8 + 3] = cmul(v[8 + 3], w6);
    v[12 + 1] = cmul(v[12 + 1], w3); v[12 + 2] = cmul(v[12 + 2], w6); v[12 + 3] = cmul(v[12 + 3], w9);
    // step 3: for every k1 a 4-point DFT over n2 -> X[k1 + 4 k2] lands in v[4 k1 + k2]; then transpose to natural order
#pragma unroll
    for (int k1 = 0; k1 < 4; ++k1) dft4(v[4 * k1], v[4 * k1 + 1], v[4 * k1 + 2], v[4 * k1 + 3]);
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int b = a + 1; b < 4; ++b) {
            const float2 t = v[4 * a + b];
            v[4 * a + b] = v[4 * b + a];
            v[4 * b + a] = t;
        }
}

__device__ __forceinline__ void dft8(float2 (&v)[8]) {
    constexpr float R2 = 0.70710678118654752f;
    float2 e0 = v[0], e1 = v[2], e2 = v[4], e3 = v[6], o0 = v[1], o1 = v[3], o2 = v[5], o3 = v[7];
    dft4(e0, e1, e2, e3);
    dft4(o0, o1, o2, o3);
    o1 = cmul(o1, make_float2(R2, -R2));
    o2 = make_float2(o2.y, -o2.x);
    o3 = cmul(o3, make_float2(-R2, -R2));
    v[0] = cadd(e0, o0); v[4] = csub(e0, o0);
    v[1] = cadd(e1, o1); v[5] = csub(e1, o1);
    v[2] = cadd(e2, o2); v[6] = csub(e2, o2);
    v[3] = cadd(e3, o3); v[7] = csub(e3, o3);
}

// v[t] *= w^t, t = 1 .. R-1, powers built by squaring / short products (depth <= 4 multiplications)
template <int R>
__device__ __forceinline__ void twiddle_powers(float2 (&v)[R], float2 w) {
    float2 p[R];
    p[1] = w;
#pragma unroll
    for (int t = 2; t < R; ++t) p[t] = (t & 1) ? cmul(p[t - 1], w) : cmul(p[t / 2], p[t / 2]);
#pragma unroll
    for (int t = 1; t < R; ++t) v[t] = cmul(v[t], p[t]);
}

// passes 2 and 3 of the forward transform; pass 1 has already put its butterflies into v.  S: F_S1 float2.
// On return S[0 .. FN) holds the spectrum in natural order (after the trailing barrier).
__device__ __forceinline__ void fft2048_tail(float2 (&v)[16], float2* S, const float2* __restrict__ tw, int j) {
    dft16(v);
    // pass-1 output y[16 j + t] into the padded exchange
#pragma unroll
    for (int t = 0; t < 16; ++t) S[17 * j + t] = v[t];
    __syncthreads();
    // pass 2: Ns = 16, R = 16
#pragma unroll
    for (int t = 0; t < 16; ++t) {
        const int i = j + FT * t;
        v[t] = S[i + (i >> 4)];
    }
    const int k2 = j & 15;
    twiddle_powers<16>(v, __ldg(tw + k2 * (FN / 256)));
    dft16(v);
    __syncthreads();                                         // every read of the padded layout is done
    {
        const int base = ((j - k2) << 4) + k2;
#pragma unroll
        for (int t = 0; t < 16; ++t) S[base + 16 * t] = v[t];
    }
    __syncthreads();
    // pass 3: Ns = 256, R = 8, two butterflies per thread (j and j + 128)
    float2 a[8], b[8];
#pragma unroll
    for (int t = 0; t < 8; ++t) {
        a[t] = S[j + 256 * t];
        b[t] = S[j + FT + 256 * t];
    }
    twiddle_powers<8>(a, __ldg(tw + j));
    twiddle_powers<8>(b, __ldg(tw + j + FT));
    dft8(a);
    dft8(b);
    __syncthreads();
#pragma unroll
    for (int t = 0; t < 8; ++t) {
        S[j + 256 * t] = a[t];
        S[j + FT + 256 * t] = b[t];
    }
    __syncthreads();
}

// STFT-domain filtering of one frame, N = 2048: same arithmetic as stft_filter_kernel, FFTs in registers.
__global__ void __launch_bounds__(FT)
stft_filter2048_kernel(StftFilterArgs a, FrameGrid g) {
    __shared__ float2 S[F_S1];
    __shared__ float red[FT / 32];
    __shared__ int lifter_row;

    const int f = blockIdx.x;
    const int u = g.frame_utt[f];
    const int j = threadIdx.x;
    float* frame_out = a.frames_out + (long long)f * a.win;
    if (u < 0) {                                             // guard frame: keep the frame buffer defined
        for (int n = j; n < a.win; n += FT) frame_out[n] = 0.f;
        if (a.lifter_index_out && j == 0) a.lifter_index_out[f] = 0;
        return;
    }
    const long long s_lo = (long long)g.utt_begin[u] * a.hop, s_hi = (long long)g.utt_end[u] * a.hop;

    // ---- cepstral lifter selection (custom_pulsed_generator.py:507-525) ----
    if (a.lifters != nullptr) {
        const long long p_lo = (long long)g.utt_begin[u] * a.pulse_per_frame;
        const long long p_hi = (long long)g.utt_end[u] * a.pulse_per_frame;
        const long long start = (long long)f * a.pulse_per_frame - a.n_smooth / 2;
        float part = 0.f;
        for (int i = j; i < a.n_smooth; i += FT) {
            long long q = start + i;
            q = q < p_lo ? p_lo : (q >= p_hi ? p_hi - 1 : q);
            part = fmaf(a.f0_smooth[i], a.f0[q], part);
        }
#pragma unroll
        for (int sft = 16; sft > 0; sft >>= 1) part += __shfl_xor_sync(0xffffffffu, part, sft);
        if ((j & 31) == 0) red[j >> 5] = part;
        __syncthreads();
        if (j == 0) {
            const float tot = (red[0] + red[1]) + (red[2] + red[3]);
            float lo10 = a.lifter_grid[0], hi10 = a.lifter_grid[a.n_lift - 1];
            float l10 = __fmul_rn(0.43429448190325176f, logf(tot));
            l10 = fminf(fmaxf(l10, lo10), hi10);
            float ratio = __fdiv_rn(__fsub_rn(l10, lo10), __fsub_rn(hi10, lo10));
            int idx = (int)rintf(__fmul_rn(ratio, (float)(a.n_lift - 1)));   // round half to even like tf.round
            lifter_row = idx;
            if (a.lifter_index_out) a.lifter_index_out[f] = idx;
        }
        __syncthreads();
    }

    // ---- forward pass 1 straight from global memory: real = windowed excitation frame, imag = liftered cepstrum ----
    const float* ceps = a.ceps + (long long)f * a.n_ceps;
    const float* lift = a.lifters ? a.lifters + (long long)lifter_row * a.n_ceps : nullptr;
    const long long x0 = (long long)f * a.hop - a.win / 2;      // first excitation sample of this frame
    float2 v[16];
#pragma unroll
    for (int t = 0; t < 16; ++t) {
        const int n = j + FT * t;
        float re = 0.f, im = 0.f;
        if (n < a.win) {
            const long long sidx = x0 + n;
            if (sidx >= s_lo && sidx < s_hi) re = __ldg(a.exc + sidx) * __ldg(a.window + n);
        }
        if (n >= 1 && n < a.n_ceps) im = lift ? __ldg(ceps + n) * __ldg(lift + n) : __ldg(ceps + n);
        v[t] = make_float2(re, im);
    }
    fft2048_tail(v, S, a.twiddle, j);

    // ---- separate the two spectra, apply the vocal-tract filter, build conj of the Hermitian product spectrum ----
    float2* vtf_out = a.vtf_out ? reinterpret_cast<float2*>(a.vtf_out) + (long long)f * (FN / 2 + 1) : nullptr;
    float2 Y[9];
#pragma unroll
    for (int t = 0; t < 9; ++t) {
        const int k = j + FT * t;                              // t = 8 only for k = 1024 (thread 0)
        if (t == 8 && j != 0) break;
        const float2 zk = S[k], zn = S[(FN - k) & (FN - 1)];
        const float2 X = make_float2(0.5f * (zk.x + zn.x), 0.5f * (zk.y - zn.y));        // spectrum of the real part
        const float2 L = make_float2(0.5f * (zk.y + zn.y), 0.5f * (zn.x - zk.x));        // spectrum of the imag part
        // exp(r tanh(x)) with tanh from one fast exponential: 1 - 2 / (1 + e^(2x)) (abs. error ~1e-7, x clamped where tanh
        // has saturated); sin / cos of the phase through the fast path for |phase| <= 64, the accurate one beyond
        float mag;
        if (a.max_log_range > 0.f) {
            const float e2 = __expf(2.f * fminf(fmaxf(L.x, -15.f), 15.f));
            mag = __expf(a.max_log_range * (1.f - __fdividef(2.f, 1.f + e2)));
        } else {
            mag = expf(L.x);
        }
        float sn, cs;
        if (fabsf(L.y) <= 64.f) __sincosf(L.y, &sn, &cs);
        else sincosf(L.y, &sn, &cs);
        const float2 V = make_float2(mag * cs, mag * sn);
        if (vtf_out) vtf_out[k] = V;
        Y[t] = cmul(X, V);
    }
    __syncthreads();
    // inverse transform as real(fft(conj(W))) with W Hermitian: W[k] = Y, W[N - k] = conj(Y)
#pragma unroll
    for (int t = 0; t < 9; ++t) {
        const int k = j + FT * t;
        if (t == 8 && j != 0) break;
        S[k] = make_float2(Y[t].x, -Y[t].y);
        if (k > 0 && k < FN / 2) S[FN - k] = Y[t];
    }
    __syncthreads();
#pragma unroll
    for (int t = 0; t < 16; ++t) v[t] = S[j + FT * t];
    __syncthreads();                                         // pass-1 reads done before the padded layout is written
    // ---- inverse passes; the last one goes straight to global memory (only the first `win` samples are kept) ----
    dft16(v);
#pragma unroll
    for (int t = 0; t < 16; ++t) S[17 * j + t] = v[t];
    __syncthreads();
#pragma unroll
    for (int t = 0; t < 16; ++t) {
        const int i = j + FT * t;
        v[t] = S[i + (i >> 4)];
    }
    const int k2 = j & 15;
    twiddle_powers<16>(v, __ldg(a.twiddle + k2 * (FN / 256)));
    dft16(v);
    __syncthreads();
    {
        const int base = ((j - k2) << 4) + k2;
#pragma unroll
        for (int t = 0; t < 16; ++t) S[base + 16 * t] = v[t];
    }
    __syncthreads();
    const float scale = 1.f / (float)FN;
#pragma unroll
    for (int hb = 0; hb < 2; ++hb) {
        const int jj = j + FT * hb;
        float2 c[8];
#pragma unroll
        for (int t = 0; t < 8; ++t) c[t] = S[jj + 256 * t];
        twiddle_powers<8>(c, __ldg(a.twiddle + jj));
        dft8(c);
#pragma unroll
        for (int t = 0; t < 8; ++t) {
            const int n = jj + 256 * t;
            if (n < a.win) frame_out[n] = c[t].x * scale * __ldg(a.inv_window + n);
        }
    }
}

// ---- analysis side: audio -> log-mel ---------------------------------------------------------------
// compute_mel_spectrogram_internal (vocoder/model/preprocess.py:479-560) over calc_stft (sig_proc/spec/stft.py:54-94):
// reflect-pad by (win/2, win), len/hop + 1 frames of `win` samples, symmetric Hann, zero-pad to 2048, |rfft|, mel basis,
// log.  One CTA = two consecutive frames of one utterance carried by ONE complex FFT (frame 2p in the real part, frame
// 2p+1 in the imaginary part, spectra separated by symmetry); the 2 x 1025 magnitudes stay in shared memory and the
// band-compressed triangular mel basis is applied by warp-wide dot products, so HBM sees the samples once (neighbouring
// frames overlap in L2) and 80 floats per frame on the way out.
constexpr int MA_MAXMEL = 128;

__device__ __forceinline__ int reflect_index(int s, int L) {
    // numpy.pad(mode="reflect") of any width: periodic extension of period 2 (L - 1) without repeating the edge sample
    if ((unsigned)s < (unsigned)L) return s;
    if (L == 1) return 0;
    const int period = 2 * (L - 1);
    int r = s % period;
    if (r < 0) r += period;
    return r >= L ? period - r : r;
}

__global__ void __launch_bounds__(FT)
mel_analysis2048_kernel(MelAnalysisArgs a) {
    __shared__ float2 S[F_S1];
    __shared__ float M[2][FN / 2 + 1];
    __shared__ float O[2][MA_MAXMEL];

    const int p = blockIdx.x, j = threadIdx.x;
    int lo = 0, hi = a.n_utt;                              // utterance u with pair_first[u] <= p < pair_first[u + 1]
    while (hi - lo > 1) {
        const int mid = (lo + hi) >> 1;
        if (__ldg(a.pair_first + mid) <= p) lo = mid; else hi = mid;
    }
    const int u = lo;
    const int L = a.n_samples[u];
    const int fbeg = a.frame_begin[u];
    const int nfr = a.frame_begin[u + 1] - fbeg;
    const int fr0 = 2 * (p - a.pair_first[u]);
    const bool has2 = fr0 + 1 < nfr;
    const float* x = a.audio + a.sample_begin[u];

    float2 v[16];
    const int s0 = fr0 * a.hop - a.win / 2;           // utterance-local sample index (an utterance holds < 2^31 samples)
#pragma unroll
    for (int t = 0; t < 16; ++t) {
        const int n = j + FT * t;
        float re = 0.f, im = 0.f;
        if (n < a.win) {
            const float w = __ldg(a.window + n);
            re = __ldg(x + reflect_index(s0 + n, L)) * w;
            if (has2) im = __ldg(x + reflect_index(s0 + a.hop + n, L)) * w;
        }
        v[t] = make_float2(re, im);
    }
    fft2048_tail(v, S, a.twiddle, j);

#pragma unroll
    for (int t = 0; t < 9; ++t) {
        const int k = j + FT * t;                              // t = 8 only for k = 1024 (thread 0)
        if (t == 8 && j != 0) break;
        const float2 zk = S[k], zn = S[(FN - k) & (FN - 1)];
        const float2 X0 = make_float2(0.5f * (zk.x + zn.x), 0.5f * (zk.y - zn.y));        // spectrum of the real part
        const float2 X1 = make_float2(0.5f * (zk.y + zn.y), 0.5f * (zn.x - zk.x));        // spectrum of the imag part
        M[0][k] = sqrtf(X0.x * X0.x + X0.y * X0.y);
        M[1][k] = sqrtf(X1.x * X1.x + X1.y * X1.y);
    }
    __syncthreads();
    if (a.mag_out) {
        for (int fr = 0; fr < (has2 ? 2 : 1); ++fr) {
            float* dst = a.mag_out + (long long)(fbeg + fr0 + fr) * (FN / 2 + 1);
            for (int k = j; k <= FN / 2; k += FT) dst[k] = M[fr][k];
        }
    }
    // thread = band, both frames at once (one weight load feeds two FMAs); the bands are served widest first so that the
    // 48 threads left over by n_mel = 80 sit in the warp of the narrow bands
    for (int b = a.n_mel - 1 - j; b >= 0; b -= FT) {
        const int blo = __ldg(a.mel_lo + b), cnt = __ldg(a.mel_cnt + b);
        const float* w = a.mel_w + __ldg(a.mel_off + b);
        float acc0 = 0.f, acc1 = 0.f;
        for (int i = 0; i < cnt; ++i) {
            const float wi = __ldg(w + i);
            acc0 = fmaf(M[0][blo + i], wi, acc0);
            acc1 = fmaf(M[1][blo + i], wi, acc1);
        }
#pragma unroll
        for (int fr = 0; fr < 2; ++fr) {
            const float acc = fr ? acc1 : acc0;
            float r;
            if (a.mode == 0) r = logf(fmaxf(acc, a.floor));                                   // do_post=False (preprocess.py:543)
            else if (a.mode == 1) r = a.log_scale * logf(fmaf(acc, a.lin_scale, a.lin_off));  // preprocess.py:107
            else r = a.log_scale * logf(fmaxf(acc * a.lin_scale, a.lin_off));                 // use_max_limit, :102
            O[fr][b] = r;
        }
    }
    __syncthreads();
    for (int i = j; i < (has2 ? 2 : 1) * a.n_mel; i += FT) {
        const int fr = i >= a.n_mel ? 1 : 0, b = i - fr * a.n_mel;
        a.mel_out[(long long)(fbeg + fr0 + fr) * a.n_mel + b] = O[fr][b];
    }
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

bool post_pqmf_supported(const PostPqmfArgs& a) {
    return a.S <= PP_MAXS && a.cin <= PP_MAXC && a.ld <= PP_MAXC && a.ld % 4 == 0 && a.Q <= PP_MAXQ && (PP_ROWS * a.S) % 4 == 0;
}

cudaError_t launch_post_pqmf(const PostPqmfArgs& a, const FrameGrid& g, cudaStream_t s) {
    if (a.rows <= 0) return cudaSuccess;
    const int tile_rows = PP_ROWS + a.Q - 1;
    const size_t smem = (size_t)(a.Q * a.S * PP_MAXS + (a.cin + 1) * PP_MAXS + tile_rows * (a.ld + 1) + tile_rows * a.S) * sizeof(float);
    if (smem > 48 * 1024) {
        cudaError_t e = cudaFuncSetAttribute(post_pqmf_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
    }
    post_pqmf_kernel<<<(unsigned)((a.rows + PP_ROWS - 1) / PP_ROWS), PP_ROWS, smem, s>>>(a, g);
    return cudaGetLastError();
}

cudaError_t launch_stft_filter(const StftFilterArgs& a, const FrameGrid& g, cudaStream_t s) {
    if (a.n_frames <= 0) return cudaSuccess;
    if (a.fft & (a.fft - 1)) return cudaErrorInvalidValue;
    if (a.fft == FN && a.win <= FN && a.n_ceps <= FN / 2) {           // the scheme configuration: register-resident FFT
        stft_filter2048_kernel<<<a.n_frames, FT, 0, s>>>(a, g);
        return cudaGetLastError();
    }
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

bool mel_analysis_supported(const MelAnalysisArgs& a) {
    return a.fft == FN && a.win <= FN && a.win > 0 && a.hop > 0 && a.n_mel > 0 && a.n_mel <= MA_MAXMEL;
}

cudaError_t launch_mel_analysis(const MelAnalysisArgs& a, cudaStream_t s) {
    if (a.n_pairs <= 0) return cudaSuccess;
    if (!mel_analysis_supported(a)) return cudaErrorInvalidValue;
    mel_analysis2048_kernel<<<a.n_pairs, FT, 0, s>>>(a);
    return cudaGetLastError();
}

}  // namespace mbx
