// Launcher declarations of the MBExWN forward kernels (definitions in k_*.cu).
#pragma once
#include "common.cuh"

namespace mbx {

// ---- k_conv.cu -----------------------------------------------------------------------------------
struct ConvArgs {
    const float* x;      // (rows, ld_x) input activations at `rate` rows per frame
    int ld_x;
    const float* w;      // (k, cin, cout) folded weight-norm kernel, row-major
    const float* bias;   // (cout)
    const float* alpha;  // PReLU slopes (act_mod) or nullptr
    float* out;          // (rows, ld_out)
    int ld_out;
    long long rows;
    int rate;
    int k, cin, cout, dilation, pad_l, pad_mode;
    int act;             // Act
    int act_mod;         // alpha index = co % act_mod (sub-pixel convs share alpha across the unfold)
    float leaky;         // LeakyReLU slope
    float a0, a1;        // affine of ACT_SOFT_SIGMOID_AFFINE
    int sub_ch;          // > 0: sub-pixel unfold with its own row pitch -- channel co of row r goes to row r * (cout / sub_ch) +
                         // co / sub_ch, column co % sub_ch of (rows * cout / sub_ch, ld_out); 0: out[r * ld_out + co]
};
cudaError_t launch_conv1d(const ConvArgs& a, const FrameGrid& g, cudaStream_t s);

struct LinInterpArgs {
    const float* x;      // (rows_in, ch)
    float* out;          // (rows_in * up, ch)
    long long rows_in;
    int rate_in;
    int ch;
    int up;
    int act;
    const float* alpha;
    float leaky;
    float a0, a1;
};
cudaError_t launch_lininterp(const LinInterpArgs& a, const FrameGrid& g, cudaStream_t s);

// LinInterp(up1) + act1 -> 1x1 conv to one channel (+ bias) -> LinInterp(up2) + act2, fused (tail of the F0 sub-net)
struct SubnetTailArgs {
    const float* x;        // (rows_in, ch) at rate_in rows per frame
    float* out;            // (rows_in * up1 * up2)
    long long rows_in;
    int rate_in, ch, up1, up2;
    int act1;              // activation after the first interpolation (PReLU / LeakyReLU / none)
    const float* alpha1;   // PReLU slopes (ch)
    const float* w;        // (ch) 1x1 kernel
    float bias;
    int act2;              // activation after the second interpolation
    float leaky, a0, a1;
};
bool subnet_tail_supported(const SubnetTailArgs& a);
cudaError_t launch_subnet_tail(const SubnetTailArgs& a, const FrameGrid& g, cudaStream_t s);

struct GateArgs {
    const float* z;        // (rows, 2C) dilated conv output incl. bias
    const float* cond;     // (rows / lin_up, 2C) conditioning at the low rate
    float* act;            // (rows, C)
    long long rows;
    int rate;              // rows per frame
    int c;
    int lin_up;
    int gate;
};
cudaError_t launch_gate(const GateArgs& a, const FrameGrid& g, cudaStream_t s);

struct ResSkipArgs {
    const float* rs;       // (rows, n_rs) 1x1 output incl. bias; n_rs = 2C (res | skip) or C (skip only)
    float* h;              // (rows, C) residual stream, updated in place when n_rs == 2C
    float* skip;           // (rows, C)
    long long rows;
    int rate;
    int c;
    int n_rs;
    int first;             // 1: skip = ..., 0: skip += ...
};
cudaError_t launch_resskip(const ResSkipArgs& a, const FrameGrid& g, cudaStream_t s);

// ---- k_excitation.cu -----------------------------------------------------------------------------
struct ExcitationArgs {
    const float* f0;        // (frames * pulse_per_frame) Hz at the pulse rate
    const float* noise;     // (frames * steps_per_frame) standard normal, or nullptr => in-kernel Philox
    unsigned long long seed;
    const int32_t* utt_ids; // global utterance ids for the Philox counter, or nullptr => batch index
    const float* tables;    // (n_period + 1, n_tables)
    int n_period, n_tables;
    float pulse_rate;
    float nominal_f0, min_tr, max_tr, grid_norm;
    float sigma;
    int pulse_per_frame, steps_per_frame, pulse_channels;
    int subharm;            // add_subharm_chans: extra sin(2 pi phase / ii) values per pulse sample (tf_wavetable.py:554-559)
    int pqmf_taps;          // pulse_channels_use_pqmf: taps of the pulse analysis bank (0 = off; needs pulse_out)
    const float* pqmf_ana;  // (pulse_channels, pqmf_taps + 1) analysis filters
    int chunk;              // cumsum chunk (1000, tf_wavetable.py:429)
    float* cum;             // scratch (frames * pulse_per_frame): in-chunk running sums
    float* chunk_off;       // scratch (n_chunks_total): per chunk offsets
    const int32_t* chunk_first;  // [n_utt + 1] first chunk slot of each utterance (exclusive scan)
    const float* phase_carry;    // [n_utt] or nullptr: running sum of the chunk totals before the utterance's first chunk
    float* wn_in;           // (frames * steps_per_frame, ld_wn_in): pulse_channels x (1 + subharm) values [+ noise]
    int ld_wn_in;
    float* phase_out;       // optional taps (frames * pulse_per_frame)
    int32_t* index_out;
    float* pulse_out;
};
cudaError_t launch_excitation(const ExcitationArgs& a, const FrameGrid& g, int n_chunks_total, cudaStream_t s);

// long-form helpers: phase carry before every cumsum chunk of one signal; row-segment gather between device buffers
cudaError_t launch_phase_carry(const float* f0, long long n_samples, float pulse_rate, int chunk, float* run_out, cudaStream_t s);
cudaError_t launch_gather_rows(const float* src, float* dst, int row_elems, const long long* seg, int n_seg, int max_rows, cudaStream_t s);

// ---- k_synth.cu ----------------------------------------------------------------------------------
struct PqmfArgs {
    const float* sub;       // (frames * steps, S)
    const float* poly;      // (Q, S, S) polyphase bank
    float* out;             // (frames * steps * S)
    long long rows;         // frames * steps
    int steps_per_frame;
    int S, Q, back;
};
cudaError_t launch_pqmf(const PqmfArgs& a, const FrameGrid& g, cudaStream_t s);

// wn_post_net 1x1 (custom_pulsed_generator.py:913-914) fused with the polyphase PQMF synthesis
struct PostPqmfArgs {
    const float* wn_out;    // (rows, ld) WaveNet output, channel padding beyond cin
    int ld, cin;
    const float* post_w;    // (cin, S)
    const float* post_b;    // (S)
    const float* poly;      // (Q, S, S) polyphase bank
    float* sub_out;         // optional tap (rows, S)
    float* out;             // (rows * S)
    long long rows;
    int steps_per_frame;
    int S, Q, back;
    // ps_use_stft = False: per-band log gains of the PS sub-net at frame rate (frames, S), or nullptr.  The reference
    // interpolates exp(gain) linearly by `gain_up` (= hop) and multiplies sub-band row r of an utterance with value r of
    // that sequence (custom_pulsed_generator.py:453, :669-670, :916-917)
    const float* log_gain;
    int gain_up;
    int gain_center;        // spect_filters_preserve_energy: subtract the mean over the bands first (:867-876)
};
bool post_pqmf_supported(const PostPqmfArgs& a);
cudaError_t launch_post_pqmf(const PostPqmfArgs& a, const FrameGrid& g, cudaStream_t s);

struct StftFilterArgs {
    const float* exc;       // (frames * hop)
    const float* ceps;      // (frames, n_ceps)
    const float* f0;        // (frames * pulse_per_frame) for the lifter selection, or nullptr
    const float* lifters;   // (n_lift, n_ceps) or nullptr
    const float* lifter_grid;   // (n_lift) log10 f0
    const float* f0_smooth; // (n_smooth) normalised Bartlett kernel
    int n_lift, n_smooth, pulse_per_frame;
    const float* window;    // (win)
    const float* inv_window;// (win)
    const float2* twiddle;  // (fft/2) exp(-2 pi i k / fft)
    float* frames_out;      // (frames, win) windowed synthesis frames
    float* vtf_out;         // optional tap (frames, fft/2+1) complex
    int32_t* lifter_index_out;  // optional tap (frames)
    int n_frames, hop, win, fft, n_ceps;
    float max_log_range;    // 0 => plain exp
};
cudaError_t launch_stft_filter(const StftFilterArgs& a, const FrameGrid& g, cudaStream_t s);

struct OlaArgs {
    const float* frames;    // (frames, win)
    float* out;             // (frames * hop)
    int n_frames, hop, win;
};
cudaError_t launch_ola(const OlaArgs& a, const FrameGrid& g, cudaStream_t s);

// ---- k_norm.cu: NormMelComponents (wavegen_1d.py:578-769) --------------------------------------------------------
struct NormArgs {
    const float* proj;        // (n_mel) inv_enorm when proj_cols == 0, else (n_mel, proj_cols) pinv(mel basis)^T
    const float* smooth_win;  // (ws) Hann [squared] smoothing window
    const float* gwin;        // (win) Hann / sum
    int n_mel, proj_cols, hop, win, ws, off, iters, use_max_limit;
    float proj_scale, norm_fact, floor, compress_exp, lin_scale, lin_off, mel_scale;
};
// mell (frames, n_mel) -> mel_out normalised; rms_a / rms_b (frames) scratch; *rms_prev_out = the frame RMS the last
// smoothing iteration started from (its gain is what the output is multiplied with)
cudaError_t launch_norm_mel(const NormArgs& a, const FrameGrid& g, const float* mell, float* rms_a, float* rms_b,
                            float* mel_out, const float** rms_prev_out, int* launches, cudaStream_t s);
cudaError_t launch_norm_apply(const NormArgs& a, const FrameGrid& g, const float* rms_prev, float* out, float* gain_tap,
                              int out_hop, cudaStream_t s);

// audio -> log-mel (analysis side); all pointers are device pointers
struct MelAnalysisArgs {
    const float* audio;             // utterances back to back
    const long long* sample_begin;  // [n_utt] first sample of each utterance in `audio`
    const int32_t* n_samples;       // [n_utt]
    const int32_t* frame_begin;     // [n_utt + 1] exclusive scan of n_samples / hop + 1
    const int32_t* pair_first;      // [n_utt + 1] exclusive scan of ceil(frames / 2): one CTA per frame pair
    int n_utt, n_pairs;
    const float* window;            // (win) symmetric Hann of the reference's window generator
    const float2* twiddle;          // (fft/2) exp(-2 pi i k / fft)
    const int32_t* mel_lo;          // (n_mel) first bin of each band
    const int32_t* mel_cnt;         // (n_mel) bins per band
    const int32_t* mel_off;         // (n_mel) offset of the band's weights in mel_w
    const float* mel_w;             // packed band weights
    int hop, win, fft, n_mel;
    int mode;                       // 0 log(max(mel, floor)); 1 log_scale log(mel lin_scale + lin_off); 2 log_scale log(max(mel lin_scale, lin_off))
    float lin_scale, lin_off, log_scale, floor;
    float* mel_out;                 // (frames, n_mel)
    float* mag_out;                 // optional tap (frames, fft/2 + 1)
};
bool mel_analysis_supported(const MelAnalysisArgs& a);
cudaError_t launch_mel_analysis(const MelAnalysisArgs& a, cudaStream_t s);

}  // namespace mbx
