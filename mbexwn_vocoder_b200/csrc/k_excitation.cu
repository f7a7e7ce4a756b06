// Glottal-pulse excitation generator: F0 -> phase velocity -> chunked cumulative phase -> wavetable lookup ->
// F0-grid cross-fade -> (20T x 5) reshape + noise channel = WaveNet input.
//
// Reference: PulseWaveTable.call / stable_cumsum_and_wrap / _linear_lookup (tf_wavetable.py:429-492, :495-560,
// :605-638) and the head of MBExWN.generate_excitation (custom_pulsed_generator.py:886-906).
//
// Bit-exactness contract (integer wavetable index): the reference accumulates float32 phase velocities
// *sequentially* inside chunks of 1000 samples anchored at the utterance start, wraps the chunk totals with a
// floor-mod 1, accumulates those (unwrapped) sequentially over chunks, wraps, adds, wraps.  The three kernels below
// keep exactly that association order with __fadd_rn/__fmul_rn/__fdiv_rn (no FMA contraction, no tree scan):
//   1. phase_chunk_kernel : one warp per chunk; lanes stage the chunk in shared memory (coalesced), one lane runs
//                           the sequential sum, lanes write the running sums back (coalesced)
//   2. chunk_offset_kernel: one warp per utterance; sequential (warp-uniform) scan over the chunk totals
//   3. pulse_kernel       : fully parallel; wrap, 2-tap table lookup, 2-table cross-fade, reshape, noise
// Parallelism comes from the B*T/10 independent chunks, not from splitting a chain.
#include "kernels.cuh"

namespace mbx {

namespace {

constexpr int CHUNK_WARPS = 4;
constexpr int MAX_CHUNK = 1024;

__device__ __forceinline__ float wrap1(float x) { return x - floorf(x); }   // floor-mod 1 for x >= 0 (exact)

__device__ __forceinline__ int find_utt(const int32_t* chunk_first, int n_utt, int c) {
    int lo = 0, hi = n_utt;            // chunk_first has n_utt + 1 entries, find u with first[u] <= c < first[u+1]
    while (hi - lo > 1) {
        int mid = (lo + hi) >> 1;
        if (chunk_first[mid] <= c) lo = mid; else hi = mid;
    }
    return lo;
}

__global__ void __launch_bounds__(CHUNK_WARPS * 32)
phase_chunk_kernel(ExcitationArgs a, FrameGrid g, int n_chunks_total) {
    // separate input / output stages: the sequential lane streams float4s in and out without load-after-store hazards,
    // so the 1000-long chain runs at the latency of its dependent adds
    __shared__ __align__(16) float buf_in[CHUNK_WARPS][MAX_CHUNK];
    __shared__ __align__(16) float buf_out[CHUNK_WARPS][MAX_CHUNK];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int c = blockIdx.x * CHUNK_WARPS + warp;
    if (c >= n_chunks_total) return;
    const int u = find_utt(a.chunk_first, g.n_utt, c);
    const int j = c - a.chunk_first[u];
    const long long base = (long long)g.utt_begin[u] * a.pulse_per_frame;
    const long long n_u = (long long)(g.utt_end[u] - g.utt_begin[u]) * a.pulse_per_frame;
    const long long s0 = (long long)j * a.chunk;
    const int n = (int)min((long long)a.chunk, n_u - s0);
    const int n8 = (n + 7) & ~7;
    float* __restrict__ si = buf_in[warp];
    float* __restrict__ so = buf_out[warp];
    // zero padding up to a multiple of 8 leaves the running sum unchanged (x + 0 is exact)
    for (int i = lane; i < n8; i += 32) si[i] = i < n ? __fdiv_rn(__ldg(a.f0 + base + s0 + i), a.pulse_rate) : 0.f;
    __syncwarp();
    if (lane == 0) {
        float acc = 0.f;
#pragma unroll 4
        for (int i = 0; i < n8; i += 8) {
            const float4 x0 = *reinterpret_cast<const float4*>(si + i), x1 = *reinterpret_cast<const float4*>(si + i + 4);
            float4 y0, y1;
            acc = __fadd_rn(acc, x0.x); y0.x = acc;
            acc = __fadd_rn(acc, x0.y); y0.y = acc;
            acc = __fadd_rn(acc, x0.z); y0.z = acc;
            acc = __fadd_rn(acc, x0.w); y0.w = acc;
            acc = __fadd_rn(acc, x1.x); y1.x = acc;
            acc = __fadd_rn(acc, x1.y); y1.y = acc;
            acc = __fadd_rn(acc, x1.z); y1.z = acc;
            acc = __fadd_rn(acc, x1.w); y1.w = acc;
            *reinterpret_cast<float4*>(so + i) = y0;
            *reinterpret_cast<float4*>(so + i + 4) = y1;
        }
        // zero padding up to the chunk size leaves the last running sum unchanged
        a.chunk_off[c] = wrap1(acc);
    }
    __syncwarp();
    for (int i = lane; i < n; i += 32) a.cum[base + s0 + i] = so[i];
}

__global__ void chunk_offset_kernel(ExcitationArgs a, FrameGrid g) {
    // chunk_off[c] holds (chunk total mod 1) on entry and the offset to add to chunk c on exit
    const int u = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (u >= g.n_utt) return;
    const int first = a.chunk_first[u], n = a.chunk_first[u + 1] - first;
    // unwrapped running sum of the wrapped totals (tf.cumsum(offsets, axis=1)); a window of a longer signal continues
    // the sum of the chunks before it
    float run = a.phase_carry ? a.phase_carry[u] : 0.f;
    for (int b0 = 0; b0 < n; b0 += 32) {
        int i = b0 + lane;
        float tot = i < n ? a.chunk_off[first + i] : 0.f;
        float mine = 0.f;
#pragma unroll
        for (int l = 0; l < 32; ++l) {
            // offset of chunk b0 + l = wrapped sum of the totals of chunks 0 .. b0 + l - 1 (0 for the first chunk);
            // every lane carries the same `run`, lane l keeps its own chunk's value
            float off_l = wrap1(run);
            if (l == lane) mine = off_l;
            run = __fadd_rn(run, __shfl_sync(0xffffffffu, tot, l));
        }
        if (i < n) a.chunk_off[first + i] = mine;
    }
}

// Philox4x32-10 counter-based generator (Salmon et al. 2011) for the in-kernel noise channel.
__device__ __forceinline__ uint4 philox4x32(uint4 ctr, uint2 key) {
    const unsigned M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        unsigned hi0 = __umulhi(M0, ctr.x), lo0 = M0 * ctr.x;
        unsigned hi1 = __umulhi(M1, ctr.z), lo1 = M1 * ctr.z;
        ctr = make_uint4(hi1 ^ ctr.y ^ key.x, lo1, hi0 ^ ctr.w ^ key.y, lo0);
        key.x += W0;
        key.y += W1;
    }
    return ctr;
}

__device__ __forceinline__ float philox_normal(unsigned long long seed, unsigned utt, unsigned long long pos) {
    uint4 r = philox4x32(make_uint4((unsigned)pos, (unsigned)(pos >> 32), utt, 0x4d425857u),
                         make_uint2((unsigned)seed, (unsigned)(seed >> 32)));
    float u1 = ((float)(r.x >> 8) + 0.5f) * (1.f / 16777216.f);
    float u2 = ((float)(r.y >> 8) + 0.5f) * (1.f / 16777216.f);
    return sqrtf(-2.f * logf(u1)) * cospif(2.f * u2);
}

// wavetable pulse sample at wrapped phase `phase` for fundamental f0: _linear_lookup + F0-grid cross-fade
// (tf_wavetable.py:621-638, :539-548); same operation order as the in-line code of pulse_kernel
__device__ __forceinline__ float pulse_value(const ExcitationArgs& a, float phase, float f0, int& i0) {
    const float p = __fmul_rn(phase, (float)a.n_period);
    const float pq = floorf(p);
    const float frac = __fsub_rn(p, pq);
    i0 = (int)pq;
    const float one_m = __fsub_rn(1.f, frac);
    const float ratio = fmaxf(a.min_tr, fminf(a.max_tr, __fdiv_rn(f0, a.nominal_f0)));
    const float x = __fmul_rn(logf(ratio), a.grid_norm);
    int k0 = (int)floorf(x);
    if (k0 < 0) k0 = 0;
    if (k0 > a.n_tables - 1) k0 = a.n_tables - 1;
    const int i0c = min(max(i0, 0), a.n_period - 1);
    const float* t0 = a.tables + (long long)i0c * a.n_tables;
    const float* t1 = t0 + a.n_tables;
    float acc = 0.f;
#pragma unroll
    for (int dk = 0; dk < 2; ++dk) {
        int k = k0 + dk;
        if (k < a.n_tables) {
            float w = fmaxf(__fsub_rn(1.f, fabsf(__fsub_rn(x, (float)k))), 0.f);
            float s = __fadd_rn(__fmul_rn(__ldg(t0 + k), one_m), __fmul_rn(__ldg(t1 + k), frac));
            acc = __fadd_rn(acc, __fmul_rn(s, w));
        }
    }
    return acc;
}

// PQMF analysis of the pulse train (pulse_channels_use_pqmf; TFPQMF.analysis, tf_preprocess.py:192-202): band k of WaveNet
// row m = sum_j pulse[m S + j - taps / 2] h_k[j] with zeros outside the utterance; one thread per (row, band).
__global__ void pulse_pqmf_kernel(ExcitationArgs a, FrameGrid g) {
    const int S = a.pulse_channels;
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long rows = (long long)g.n_frames * a.steps_per_frame;
    if (idx >= rows * S) return;
    const long long m = idx / S;
    const int k = (int)(idx - m * S);
    float* row = a.wn_in + m * a.ld_wn_in;
    long long lo, hi;
    float acc = 0.f;
    if (utt_bounds(g, a.pulse_per_frame, m * S, lo, hi)) {
        const float* hk = a.pqmf_ana + (long long)k * (a.pqmf_taps + 1);
        const long long base = m * S - a.pqmf_taps / 2;
        for (int j = 0; j <= a.pqmf_taps; ++j) {
            const long long n = base + j;
            if (n >= lo && n < hi) acc = fmaf(a.pulse_out[n], __ldg(hk + j), acc);
        }
    }
    row[k] = acc;
}

// IDX = unsigned when the grid holds fewer than 2^31 pulse samples (32-bit divisions), long long otherwise
template <class IDX>
__global__ void pulse_kernel(ExcitationArgs a, FrameGrid g) {
    const IDX n = (IDX)blockIdx.x * (IDX)blockDim.x + (IDX)threadIdx.x;        // pulse-rate sample on the grid
    const IDX total = (IDX)g.n_frames * (IDX)a.pulse_per_frame;
    if (n >= total) return;
    const IDX step = n / (IDX)a.pulse_channels;
    const int ch = (int)(n - step * (IDX)a.pulse_channels);
    float* row = a.wn_in + (long long)step * a.ld_wn_in;
    const int per = 1 + a.subharm;                     // values per pulse sample: pulse [+ sub-harmonic sinusoids]
    if (a.pqmf_taps > 0) {
        // pulse_channels_use_pqmf: the row is [pulse_channels analysis bands (pulse_pqmf_kernel) | sub-harmonic values of the
        // pulse_channels samples | noise] (custom_pulsed_generator.py:895-900); the pulse itself only goes to pulse_out
        const int nsub = a.subharm;
        const int fq = (int)(n / (IDX)a.pulse_per_frame);
        const int uq = g.frame_utt[fq];
        float phase_q = 0.f, acc_q = 0.f;
        int i0q = 0;
        if (uq >= 0) {
            const IDX loq = (IDX)g.utt_begin[uq] * (IDX)a.pulse_per_frame;
            const IDX localq = n - loq;
            const int cq = a.chunk_first[uq] + (int)(localq / (IDX)a.chunk);
            phase_q = wrap1(__fadd_rn(a.cum[n], a.chunk_off[cq]));
            acc_q = pulse_value(a, phase_q, a.f0[n], i0q);
            if (nsub > 0) {
                const float w2pi = __fmul_rn(__fmul_rn(phase_q, 2.f), 3.14159274101257324f);
                for (int j = 1; j <= nsub; ++j) row[a.pulse_channels + ch * nsub + (j - 1)] = sinf(__fdiv_rn(w2pi, (float)(j + 1)));
            }
            if (ch == 0 && a.sigma != 0.f) {
                const IDX lstep = step - loq / (IDX)a.pulse_channels;
                float z = a.noise ? a.noise[step] : philox_normal(a.seed, (unsigned)(a.utt_ids ? a.utt_ids[uq] : uq), (unsigned long long)lstep);
                row[a.pulse_channels * per] = __fmul_rn(a.sigma, z);
            }
        } else {
            for (int j = 1; j <= nsub; ++j) row[a.pulse_channels + ch * nsub + (j - 1)] = 0.f;
            if (ch == 0 && a.sigma != 0.f) row[a.pulse_channels * per] = 0.f;
        }
        if (a.phase_out) a.phase_out[n] = phase_q;
        if (a.index_out) a.index_out[n] = i0q;
        a.pulse_out[n] = acc_q;
        return;
    }
    const int f = (int)(n / (IDX)a.pulse_per_frame);
    const int u = g.frame_utt[f];
    if (u < 0) {
        for (int j = 0; j < per; ++j) row[ch * per + j] = 0.f;
        if (ch == 0 && a.sigma != 0.f) row[a.pulse_channels * per] = 0.f;
        if (a.phase_out) a.phase_out[n] = 0.f;
        if (a.index_out) a.index_out[n] = 0;
        if (a.pulse_out) a.pulse_out[n] = 0.f;
        return;
    }
    const IDX lo = (IDX)g.utt_begin[u] * (IDX)a.pulse_per_frame;
    const IDX local = n - lo;
    const int c = a.chunk_first[u] + (int)(local / (IDX)a.chunk);
    // phase = ((cum + offset) mod 1)   (tf_wavetable.py:483-486)
    const float phase = wrap1(__fadd_rn(a.cum[n], a.chunk_off[c]));
    // _linear_lookup (tf_wavetable.py:621-638)
    const float p = __fmul_rn(phase, (float)a.n_period);
    const float pq = floorf(p);
    const float frac = __fsub_rn(p, pq);
    const int i0 = (int)pq;
    const float one_m = __fsub_rn(1.f, frac);
    // table cross-fade (tf_wavetable.py:539-548); only the two grid neighbours have non-zero weight
    const float f0 = a.f0[n];
    const float ratio = fmaxf(a.min_tr, fminf(a.max_tr, __fdiv_rn(f0, a.nominal_f0)));
    const float x = __fmul_rn(logf(ratio), a.grid_norm);
    int k0 = (int)floorf(x);
    if (k0 < 0) k0 = 0;
    if (k0 > a.n_tables - 1) k0 = a.n_tables - 1;
    // memory safety only: phase < 1 and a power-of-two period keep i0 < n_period (the reference would raise otherwise)
    const int i0c = min(max(i0, 0), a.n_period - 1);
    const float* t0 = a.tables + (long long)i0c * a.n_tables;
    const float* t1 = t0 + a.n_tables;
    float acc = 0.f;
#pragma unroll
    for (int dk = 0; dk < 2; ++dk) {
        int k = k0 + dk;
        if (k < a.n_tables) {
            float w = fmaxf(__fsub_rn(1.f, fabsf(__fsub_rn(x, (float)k))), 0.f);
            float s = __fadd_rn(__fmul_rn(__ldg(t0 + k), one_m), __fmul_rn(__ldg(t1 + k), frac));
            acc = __fadd_rn(acc, __fmul_rn(s, w));
        }
    }
    row[ch * per] = acc;
    if (a.subharm > 0) {
        // add_subharm_chans (tf_wavetable.py:520-521, :554-559): sin(2 pi phase / ii), ii = 2 .. subharm + 1, folded into
        // the WaveNet row together with the pulse sample (custom_pulsed_generator.py:893)
        const float w2pi = __fmul_rn(__fmul_rn(phase, 2.f), 3.14159274101257324f);
        for (int j = 1; j < per; ++j) row[ch * per + j] = sinf(__fdiv_rn(w2pi, (float)(j + 1)));
    }
    if (ch == 0 && a.sigma != 0.f) {
        const IDX lstep = step - lo / (IDX)a.pulse_channels;
        float z = a.noise ? a.noise[step] : philox_normal(a.seed, (unsigned)(a.utt_ids ? a.utt_ids[u] : u), (unsigned long long)lstep);
        row[a.pulse_channels * per] = __fmul_rn(a.sigma, z);
    }
    if (a.phase_out) a.phase_out[n] = phase;
    if (a.index_out) a.index_out[n] = i0;
    if (a.pulse_out) a.pulse_out[n] = acc;
}

// ---- long-form helpers (chunked synthesis of one long signal, long_form.py) ------------------------------------------
// Wrapped total of every 1000-sample chunk of ONE signal, in the association order of phase_chunk_kernel: the velocities
// f0 / rate are summed sequentially in float32 inside the chunk.  One warp per chunk; lane 0 runs the chain.
__global__ void __launch_bounds__(CHUNK_WARPS * 32)
chunk_total_kernel(const float* __restrict__ f0, long long n, float pulse_rate, int chunk, float* __restrict__ tot, int n_chunks) {
    __shared__ __align__(16) float buf[CHUNK_WARPS][MAX_CHUNK];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int c = blockIdx.x * CHUNK_WARPS + warp;
    if (c >= n_chunks) return;
    const long long s0 = (long long)c * chunk;
    const int m = (int)min((long long)chunk, n - s0);
    const int m8 = (m + 7) & ~7;
    float* si = buf[warp];
    for (int i = lane; i < m8; i += 32) si[i] = i < m ? __fdiv_rn(__ldg(f0 + s0 + i), pulse_rate) : 0.f;
    __syncwarp();
    if (lane == 0) {
        float acc = 0.f;
        for (int i = 0; i < m8; ++i) acc = __fadd_rn(acc, si[i]);      // x + 0 is exact: the padding does not change the sum
        tot[c] = wrap1(acc);
    }
}

// run[c] = unwrapped float32 running sum of the wrapped totals of chunks 0 .. c - 1 (tf.cumsum over the chunk axis,
// tf_wavetable.py:470-486; the same sequential order as chunk_offset_kernel): the phase carry a window starting at chunk c
// continues from.  In place; one thread (a few thousand dependent adds for a ten-minute signal).
__global__ void chunk_run_kernel(float* __restrict__ tot_run, int n_chunks) {
    if (blockIdx.x || threadIdx.x) return;
    float run = 0.f;
    for (int c = 0; c < n_chunks; ++c) {
        const float t = tot_run[c];
        tot_run[c] = run;
        run = __fadd_rn(run, t);
    }
}

// dst[dst_row[s] + r, :] = src[src_row[s] + r, :] for r < n_rows[s]; rows of row_vec elements of type V (float4 when the
// row length allows it, else float)
template <typename V>
__global__ void gather_rows_kernel(const V* __restrict__ src, V* __restrict__ dst, int row_vec, const long long* __restrict__ seg, int n_seg) {
    const int s = blockIdx.y;
    if (s >= n_seg) return;
    const long long src_row = seg[3 * s], dst_row = seg[3 * s + 1], n_rows = seg[3 * s + 2];
    const long long total = n_rows * row_vec;
    const V* ps = src + src_row * row_vec;
    V* pd = dst + dst_row * row_vec;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) pd[i] = ps[i];
}

}  // namespace

cudaError_t launch_phase_carry(const float* f0, long long n_samples, float pulse_rate, int chunk, float* run_out, cudaStream_t s) {
    if (chunk > MAX_CHUNK || chunk < 1 || n_samples < 1) return cudaErrorInvalidValue;
    const int n_chunks = (int)((n_samples + chunk - 1) / chunk);
    chunk_total_kernel<<<(n_chunks + CHUNK_WARPS - 1) / CHUNK_WARPS, CHUNK_WARPS * 32, 0, s>>>(f0, n_samples, pulse_rate, chunk, run_out, n_chunks);
    chunk_run_kernel<<<1, 32, 0, s>>>(run_out, n_chunks);
    return cudaGetLastError();
}

cudaError_t launch_gather_rows(const float* src, float* dst, int row_elems, const long long* seg, int n_seg, int max_rows, cudaStream_t s) {
    if (row_elems < 1 || n_seg < 1 || n_seg > 65535) return cudaErrorInvalidValue;
    const bool vec = (row_elems & 3) == 0 && ((reinterpret_cast<uintptr_t>(src) | reinterpret_cast<uintptr_t>(dst)) & 15) == 0;
    const int row_vec = vec ? row_elems / 4 : row_elems;
    const long long per = (long long)max_rows * row_vec;
    int bx = (int)((per + 255) / 256);
    if (bx > 64) bx = 64;
    if (bx < 1) bx = 1;
    if (vec)
        gather_rows_kernel<float4><<<dim3((unsigned)bx, (unsigned)n_seg), 256, 0, s>>>(reinterpret_cast<const float4*>(src),
                                                                                 reinterpret_cast<float4*>(dst), row_vec, seg, n_seg);
    else
        gather_rows_kernel<float><<<dim3((unsigned)bx, (unsigned)n_seg), 256, 0, s>>>(src, dst, row_vec, seg, n_seg);
    return cudaGetLastError();
}

cudaError_t launch_excitation(const ExcitationArgs& a, const FrameGrid& g, int n_chunks_total, cudaStream_t s) {
    if (a.chunk > MAX_CHUNK) return cudaErrorInvalidValue;
    if (g.n_frames <= 0 || n_chunks_total <= 0) return cudaSuccess;
    phase_chunk_kernel<<<(n_chunks_total + CHUNK_WARPS - 1) / CHUNK_WARPS, CHUNK_WARPS * 32, 0, s>>>(a, g, n_chunks_total);
    chunk_offset_kernel<<<(g.n_utt * 32 + 127) / 128, 128, 0, s>>>(a, g);
    long long total = (long long)g.n_frames * a.pulse_per_frame;
    if (a.pqmf_taps > 0 && (!a.pulse_out || !a.pqmf_ana)) return cudaErrorInvalidValue;
    if (total < (1ll << 31) - 256) pulse_kernel<unsigned><<<(unsigned)((total + 255) / 256), 256, 0, s>>>(a, g);
    else pulse_kernel<long long><<<(unsigned)((total + 255) / 256), 256, 0, s>>>(a, g);
    if (a.pqmf_taps > 0) {
        const long long work = (long long)g.n_frames * a.steps_per_frame * a.pulse_channels;
        pulse_pqmf_kernel<<<(unsigned)((work + 255) / 256), 256, 0, s>>>(a, g);
    }
    return cudaGetLastError();
}

}  // namespace mbx
