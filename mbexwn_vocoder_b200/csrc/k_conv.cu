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
            if (a.sub_ch > 0) {
                const int sub = co / a.sub_ch;
                a.out[(r * (a.cout / a.sub_ch) + sub) * a.ld_out + (co - sub * a.sub_ch)] = v;
            } else {
                a.out[r * a.ld_out + co] = v;
            }
        }
    }
}

constexpr int LIN_MAX_UP = 512;

__global__ void lininterp_kernel(LinInterpArgs a, FrameGrid g) {
    // one thread per output element; w0 = (U-u)/U, w1 = u/U evaluated in double and rounded to float like the
    // reference's float64 -> float32 kernel initialiser (support_layers.py:19-27), once per block
    __shared__ float sw0[LIN_MAX_UP], sw1[LIN_MAX_UP];
    const bool tab = a.up <= LIN_MAX_UP;
    if (tab) {
        for (int u = threadIdx.x; u < a.up; u += blockDim.x) {
            sw0[u] = (float)((double)(a.up - u) / (double)a.up);
            sw1[u] = (float)((double)u / (double)a.up);
        }
        __syncthreads();
    }
    long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    long long total = a.rows_in * a.up * a.ch;
    if (idx >= total) return;
    int c;
    long long ro, ri;
    int u;
    if (total < (1ll << 31)) {                            // 32-bit divisions on the common path
        const unsigned i32 = (unsigned)idx;
        const unsigned ro32 = i32 / (unsigned)a.ch, ri32 = ro32 / (unsigned)a.up;
        c = (int)(i32 - ro32 * (unsigned)a.ch);
        u = (int)(ro32 - ri32 * (unsigned)a.up);
        ro = ro32; ri = ri32;
    } else {
        c = (int)(idx % a.ch);
        ro = idx / a.ch;
        ri = ro / a.up;
        u = (int)(ro - ri * a.up);
    }
    long long lo, hi;
    float v = 0.f;
    if (utt_bounds(g, a.rate_in, ri, lo, hi)) {
        long long rn = ri + 1 < hi ? ri + 1 : hi - 1;
        const float w0 = tab ? sw0[u] : (float)((double)(a.up - u) / (double)a.up);
        const float w1 = tab ? sw1[u] : (float)((double)u / (double)a.up);
        v = __fadd_rn(__fmul_rn(a.x[ri * a.ch + c], w0), __fmul_rn(a.x[rn * a.ch + c], w1));
        float al = a.act == ACT_PRELU ? a.alpha[c] : a.leaky;
        v = apply_act(v, a.act, al, a.a0, a.a1);
    }
    a.out[idx] = v;
}

// Fused tail of a sub-net: LinInterp(up1) + activation -> 1x1 conv to ONE channel -> LinInterp(up2) + activation, the shape of
// the F0 sub-net's end (custom_pulsed_generator.py:86-90, :134-146: "[k, C, 'L5']" layer, PulsPar_Layer_final, the missing
// up-sampling factor, soft-sigmoid).  One block = TAIL_Z rows at the intermediate rate (+ 1 look-ahead row): thread = row computes
// the interpolated, activated C-vector on the fly and its dot product with the 1x1 kernel; the block then writes the
// TAIL_Z * up2 outputs.  Same per-element arithmetic as lininterp_kernel / conv1d_kernel / lininterp_kernel in sequence.
constexpr int TAIL_Z = 128, TAIL_MAX_C = 128;

__global__ void __launch_bounds__(TAIL_Z)
subnet_tail_kernel(SubnetTailArgs a, FrameGrid g) {
    __shared__ float z[TAIL_Z + 1];
    __shared__ float w[TAIL_MAX_C], al[TAIL_MAX_C];
    __shared__ float w0a[LIN_MAX_UP], w1a[LIN_MAX_UP], w0b[LIN_MAX_UP], w1b[LIN_MAX_UP];
    __shared__ long long jhi[TAIL_Z + 1];                 // utterance end at the intermediate rate, -1 for guard rows
    const int t = threadIdx.x;
    for (int c = t; c < a.ch; c += TAIL_Z) {
        w[c] = a.w[c];
        al[c] = a.act1 == ACT_PRELU ? a.alpha1[c] : a.leaky;
    }
    for (int u = t; u < a.up1; u += TAIL_Z) {
        w0a[u] = (float)((double)(a.up1 - u) / (double)a.up1);
        w1a[u] = (float)((double)u / (double)a.up1);
    }
    for (int u = t; u < a.up2; u += TAIL_Z) {
        w0b[u] = (float)((double)(a.up2 - u) / (double)a.up2);
        w1b[u] = (float)((double)u / (double)a.up2);
    }
    __syncthreads();
    const long long rows_mid = a.rows_in * a.up1;
    const long long j0 = (long long)blockIdx.x * TAIL_Z;
    for (int i = t; i < TAIL_Z + 1; i += TAIL_Z) {
        const long long j = j0 + i;
        float acc = 0.f;
        long long end = -1;
        if (j < rows_mid) {
            const long long ri = j / a.up1;
            const int u = (int)(j - ri * a.up1);
            long long lo, hi;
            if (utt_bounds(g, a.rate_in, ri, lo, hi)) {
                const long long rn = ri + 1 < hi ? ri + 1 : hi - 1;
                const float* x0 = a.x + ri * a.ch;
                const float* x1 = a.x + rn * a.ch;
                const float f0 = w0a[u], f1 = w1a[u];
                for (int c = 0; c < a.ch; ++c) {
                    float v = __fadd_rn(__fmul_rn(__ldg(x0 + c), f0), __fmul_rn(__ldg(x1 + c), f1));
                    v = apply_act(v, a.act1, al[c], 0.f, 0.f);
                    acc = fmaf(v, w[c], acc);
                }
                acc += a.bias;
                end = hi * a.up1;
            }
        }
        z[i] = acc;
        jhi[i] = end;
    }
    __syncthreads();
    const long long total = rows_mid * a.up2;
    const long long o0 = j0 * a.up2;
    for (int o = t; o < TAIL_Z * a.up2; o += TAIL_Z) {
        if (o0 + o >= total) break;
        const int jl = o / a.up2, v = o - jl * a.up2;
        float r = 0.f;
        if (jhi[jl] >= 0) {
            const int jn = (j0 + jl + 1 < jhi[jl]) ? jl + 1 : jl;
            r = __fadd_rn(__fmul_rn(z[jl], w0b[v]), __fmul_rn(z[jn], w1b[v]));
            r = apply_act(r, a.act2, a.leaky, a.a0, a.a1);
        }
        a.out[o0 + o] = r;
    }
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

bool subnet_tail_supported(const SubnetTailArgs& a) {
    return a.ch >= 1 && a.ch <= TAIL_MAX_C && a.up1 >= 1 && a.up1 <= LIN_MAX_UP && a.up2 >= 1 && a.up2 <= LIN_MAX_UP &&
           a.act2 != ACT_PRELU;
}

cudaError_t launch_subnet_tail(const SubnetTailArgs& a, const FrameGrid& g, cudaStream_t s) {
    const long long rows_mid = a.rows_in * a.up1;
    if (rows_mid <= 0) return cudaSuccess;
    if (!subnet_tail_supported(a)) return cudaErrorInvalidValue;
    subnet_tail_kernel<<<(unsigned)((rows_mid + TAIL_Z - 1) / TAIL_Z), TAIL_Z, 0, s>>>(a, g);
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
