// fp32 SIMT kernels for the small convolutions of the MBExWN forward path and the fp32 reference variant
// of the WaveNet layer:
//   * conv1d_kernel     -- Keras Conv1D (cross-correlation, channels-last) with the reference's padding modes
//                          folded in, bias + PReLU/LeakyReLU/soft-sigmoid epilogue
//                          (conv_layers.py:133-165, custom_layers.py:47-71, custom_pulsed_generator.py:38-148)
//   * lininterp_kernel  -- TF2C_LinInterpLayer with num_pad_end=1, drop_last=True (support_layers.py:99-121)
//   * gate_kernel       -- tanh/sigmoid gate on conv + linearly interpolated conditioning
//                          (custom_AE_layers.py:307-321)
//   * resskip_kernel    -- residual add + skip accumulation (custom_AE_layers.py:324-335)
// All kernels work on the padded frame grid (common.cuh): per-utterance boundaries, guard rows written as 0.
#include "kernels.cuh"

namespace mbx {

namespace {

constexpr int BM = 64, BN = 64, BK = 16, TM = 4, TN = 4;
constexpr int CONV_THREADS = (BM / TM) * (BN / TN);   // 256

__global__ void __launch_bounds__(CONV_THREADS)
conv1d_kernel(ConvArgs a, FrameGrid g) {
    __shared__ float As[BK][BM + 4];
    __shared__ float Bs[BK][BN];
    __shared__ long long row_lo[BM], row_hi[BM];

    const int tid = threadIdx.x;
    const long long m0 = (long long)blockIdx.x * BM;
    const int n0 = blockIdx.y * BN;
    const int ktot = a.k * a.cin;

    if (tid < BM) {
        long long r = m0 + tid, lo = 0, hi = -1;
        if (r < a.rows && !utt_bounds(g, a.rate, r, lo, hi)) { lo = 0; hi = -1; }
        row_lo[tid] = lo;
        row_hi[tid] = hi;                                   // hi < lo marks a guard / out-of-range row
    }
    __syncthreads();

    float acc[TM][TN];
#pragma unroll
    for (int i = 0; i < TM; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j) acc[i][j] = 0.f;

    const int ty = tid / (BN / TN), tx = tid % (BN / TN);

    for (int k0 = 0; k0 < ktot; k0 += BK) {
        // A tile: gather with the layer's padding rule
#pragma unroll
        for (int i = 0; i < (BM * BK) / CONV_THREADS; ++i) {
            int e = tid + i * CONV_THREADS;
            int kl = e % BK, ml = e / BK;
            int kk = k0 + kl;
            float v = 0.f;
            if (kk < ktot && row_hi[ml] > row_lo[ml]) {
                int j = kk / a.cin, ci = kk - j * a.cin;
                long long s = m0 + ml + (long long)j * a.dilation - a.pad_l;
                long long src = pad_index(s, row_lo[ml], row_hi[ml], a.pad_mode);
                if (src >= 0) v = a.x[src * a.ld_x + ci];
            }
            As[kl][ml] = v;
        }
#pragma unroll
        for (int i = 0; i < (BK * BN) / CONV_THREADS; ++i) {
            int e = tid + i * CONV_THREADS;
            int kl = e / BN, nl = e % BN;
            int kk = k0 + kl, co = n0 + nl;
            Bs[kl][nl] = (kk < ktot && co < a.cout) ? a.w[(long long)kk * a.cout + co] : 0.f;
        }
        __syncthreads();
#pragma unroll
        for (int kl = 0; kl < BK; ++kl) {
            float av[TM], bv[TN];
#pragma unroll
            for (int i = 0; i < TM; ++i) av[i] = As[kl][ty * TM + i];
#pragma unroll
            for (int j = 0; j < TN; ++j) bv[j] = Bs[kl][tx * TN + j];
#pragma unroll
            for (int i = 0; i < TM; ++i)
#pragma unroll
                for (int j = 0; j < TN; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
        }
        __syncthreads();
    }

#pragma unroll
    for (int i = 0; i < TM; ++i) {
        int ml = ty * TM + i;
        long long r = m0 + ml;
        if (r >= a.rows) continue;
        bool valid = row_hi[ml] > row_lo[ml];
#pragma unroll
        for (int j = 0; j < TN; ++j) {
            int co = n0 + tx * TN + j;
            if (co >= a.cout) continue;
            float v = 0.f;
            if (valid) {
                v = acc[i][j] + a.bias[co];
                float al = a.act == ACT_PRELU ? a.alpha[co % a.act_mod] : a.leaky;
                v = apply_act(v, a.act, al, a.a0, a.a1);
            }
            a.out[r * a.ld_out + co] = v;
        }
    }
}

__global__ void lininterp_kernel(LinInterpArgs a, FrameGrid g) {
    // one thread per output element; w0 = (U-u)/U, w1 = u/U evaluated in double and rounded to float like the
    // reference's float64 -> float32 kernel initialiser (support_layers.py:19-27)
    long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    long long total = a.rows_in * a.up * a.ch;
    if (idx >= total) return;
    int c = (int)(idx % a.ch);
    long long ro = idx / a.ch;
    long long ri = ro / a.up;
    int u = (int)(ro - ri * a.up);
    long long lo, hi;
    float v = 0.f;
    if (utt_bounds(g, a.rate_in, ri, lo, hi)) {
        long long rn = ri + 1 < hi ? ri + 1 : hi - 1;
        float w0 = (float)((double)(a.up - u) / (double)a.up);
        float w1 = (float)((double)u / (double)a.up);
        v = __fadd_rn(__fmul_rn(a.x[ri * a.ch + c], w0), __fmul_rn(a.x[rn * a.ch + c], w1));
        float al = a.act == ACT_PRELU ? a.alpha[c] : a.leaky;
        v = apply_act(v, a.act, al, a.a0, a.a1);
    }
    a.out[idx] = v;
}

__global__ void gate_kernel(GateArgs a, FrameGrid g) {
    long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= a.rows * a.c) return;
    int c = (int)(idx % a.c);
    long long r = idx / a.c;
    long long lo, hi;
    float v = 0.f;
    if (utt_bounds(g, a.rate, r, lo, hi)) {
        long long rc = r / a.lin_up;
        int u = (int)(r - rc * a.lin_up);
        long long hic = hi / a.lin_up;
        long long rn = rc + 1 < hic ? rc + 1 : hic - 1;
        float w0 = (float)((double)(a.lin_up - u) / (double)a.lin_up);
        float w1 = (float)((double)u / (double)a.lin_up);
        const float* c0 = a.cond + rc * 2 * a.c;
        const float* c1 = a.cond + rn * 2 * a.c;
        float ct = __fadd_rn(__fmul_rn(c0[c], w0), __fmul_rn(c1[c], w1));
        float cs = __fadd_rn(__fmul_rn(c0[a.c + c], w0), __fmul_rn(c1[a.c + c], w1));
        float zt = a.z[r * 2 * a.c + c] + ct;
        float zs = a.z[r * 2 * a.c + a.c + c] + cs;
        float t;
        switch (a.gate) {
            case GATE_GTU: t = tanhf(zt); break;
            case GATE_GFU: t = zt / (1.f + fabsf(zt)); break;
            case GATE_GSU: t = zt / (1.f + sqrtf(fabsf(zt))); break;
            default: t = zt; break;
        }
        v = t * (1.f / (1.f + expf(-zs)));
    }
    a.act[idx] = v;
}

__global__ void resskip_kernel(ResSkipArgs a, FrameGrid g) {
    long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= a.rows * a.c) return;
    int c = (int)(idx % a.c);
    long long r = idx / a.c;
    long long lo, hi;
    if (!utt_bounds(g, a.rate, r, lo, hi)) {
        if (a.first) a.skip[idx] = 0.f;
        return;                                            // guard rows of h stay zero
    }
    const float* rs = a.rs + r * a.n_rs;
    float sk;
    if (a.n_rs == 2 * a.c) {
        a.h[idx] = a.h[idx] + rs[c];
        sk = rs[a.c + c];
    } else {
        sk = rs[c];
    }
    a.skip[idx] = a.first ? sk : a.skip[idx] + sk;
}

}  // namespace

cudaError_t launch_conv1d(const ConvArgs& a, const FrameGrid& g, cudaStream_t s) {
    if (a.rows <= 0) return cudaSuccess;
    dim3 grid((unsigned)((a.rows + BM - 1) / BM), (unsigned)((a.cout + BN - 1) / BN));
    conv1d_kernel<<<grid, CONV_THREADS, 0, s>>>(a, g);
    return cudaGetLastError();
}

cudaError_t launch_lininterp(const LinInterpArgs& a, const FrameGrid& g, cudaStream_t s) {
    long long total = a.rows_in * a.up * a.ch;
    if (total <= 0) return cudaSuccess;
    lininterp_kernel<<<(unsigned)((total + 255) / 256), 256, 0, s>>>(a, g);
    return cudaGetLastError();
}

cudaError_t launch_gate(const GateArgs& a, const FrameGrid& g, cudaStream_t s) {
    long long total = a.rows * a.c;
    if (total <= 0) return cudaSuccess;
    gate_kernel<<<(unsigned)((total + 255) / 256), 256, 0, s>>>(a, g);
    return cudaGetLastError();
}

cudaError_t launch_resskip(const ResSkipArgs& a, const FrameGrid& g, cudaStream_t s) {
    long long total = a.rows * a.c;
    if (total <= 0) return cudaSuccess;
    resskip_kernel<<<(unsigned)((total + 255) / 256), 256, 0, s>>>(a, g);
    return cudaGetLastError();
}

}  // namespace mbx
