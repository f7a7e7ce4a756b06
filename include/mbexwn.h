/* mbexwn.h -- C ABI of the B200-native MBExWN mel-inversion forward pass.
 *
 * Drop-in boundary for ONE path of roebel/MBExWN_Vocoder: MELInverter.synth_from_mel
 * (MBExWN_NVoc/mel_inverter.py:151-154) -> PaNWaveNet.infer (vocoder/model/wavegen_1d.py:483-526)
 * -> MBExWN.call (vocoder/model/custom_pulsed_generator.py:556-771).  The reference has no native layer
 * (it is 100 % Python over TensorFlow), so each entry point names the Python interface it stands in for.
 *
 * Conventions
 *   - plain pointers and sizes, no framework types; every function returns 0 on success or a negative
 *     mbexwn_status code, with a message available from mbexwn_last_error().
 *   - the library never allocates device memory after mbexwn_create(): the caller (PyTorch on the Python
 *     side) owns weights, inputs, outputs and one workspace buffer sized by mbexwn_workspace_bytes().
 *   - a handle is bound to the CUDA device that is current in mbexwn_create(); calls on one handle are
 *     serialised on the caller's stream.  One handle per GPU; no global state.
 *   - there is NO CPU fallback: without a CUDA device every compute entry point returns MBEXWN_ERR_CUDA.
 *
 * Batch layout ("padded frame grid"): utterances are concatenated in time with `halo_frames` all-zero guard
 * frames before the first, between neighbours and after the last one.  frame_utt[f] = utterance index of padded
 * frame f or -1 for guard frames; utt_begin/utt_end = [begin, end) padded-frame range per utterance.  Every
 * buffer at every rate uses this grid: row r at R rows per frame belongs to padded frame r / R.
 */
#ifndef MBEXWN_H_
#define MBEXWN_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(__GNUC__)
#define MBEXWN_API __attribute__((visibility("default")))
#else
#define MBEXWN_API
#endif

#define MBEXWN_ABI_VERSION 3
#define MBEXWN_MAX_LAYERS 64
#define MBEXWN_MAX_OPS 32
#define MBEXWN_MAX_BLOCKS 4

typedef struct mbexwn_handle_s* mbexwn_handle_t;

enum mbexwn_status {
    MBEXWN_OK = 0,
    MBEXWN_ERR_INVALID = -1,   /* bad argument / inconsistent config   (Python: RuntimeError / ValueError) */
    MBEXWN_ERR_MISSING = -2,   /* a tensor was not registered           (Python: KeyError)                */
    MBEXWN_ERR_CUDA = -3,      /* CUDA runtime / driver failure         (Python: RuntimeError)            */
    MBEXWN_ERR_UNSUPPORTED = -4 /* configuration outside the built path (Python: NotImplementedError)     */
};

/* precision of the WaveNet contractions */
enum mbexwn_precision {
    MBEXWN_PREC_FP32_SIMT = 0,  /* fp32 FMA on CUDA cores: bit-for-bit the reference's arithmetic type          */
    MBEXWN_PREC_BF16X3 = 1,     /* tcgen05 bf16 with hi/lo operand split (3 products), fp32 accumulate in TMEM */
    MBEXWN_PREC_BF16 = 2,       /* tcgen05 bf16 operands, fp32 accumulate                                      */
    MBEXWN_PREC_F16F8 = 3       /* tcgen05 fp16 main product + two e4m3 correction products (lo x hi, hi x lo) at the
                                   fp8 rate, folded in with scale-input-d 2^-15; fp32 accumulate; fp32-accurate    */
};

/* One step of a mel-rate conv sub-net (generate_subnet_from_specs, custom_pulsed_generator.py:38-148) with the
 * TFPad1d layer folded into the conv and the activation fused behind the op. */
typedef struct {
    int32_t kind;        /* 0 = conv1d, 1 = linear interpolation */
    int32_t k, cin, cout, dilation, pad_l, pad_r, pad_mode; /* conv; pad_mode 0 zero, 1 symmetric, 2 edge */
    int32_t subpixel;    /* depth-to-time unfold factor of the conv output (conv_layers.py:250-255) */
    int32_t up;          /* interpolation factor (kind 1) */
    int32_t act;         /* 0 none, 1 PReLU, 2 LeakyReLU, 3 soft-sigmoid + F0 affine */
    int32_t act_channels;
    int32_t rate_in, rate_out, ch_out;
    char name[96];       /* layer name: tensors <name>/W (k,cin,cout) and <name>/b */
    char act_name[96];   /* PReLU tensor <act_name>/alpha */
} mbexwn_op_t;

/* One WaveNetAEBlock of pp_waveNetBlocks (custom_AE_layers.py:457-575, custom_pulsed_generator.py:459-488): a WaveNetAE and,
 * when up > 1, the sub-pixel up-sampling conv TF2C_Conv1DUpDownSample(n_out_channels, kernel_size 3) behind it. */
typedef struct {
    int32_t c;              /* residual channels = n_channels x the block's channel factor */
    int32_t cond_conv_up;   /* sub-pixel factor of the block's conditioning conv: block rate / (frame rate x cond_lin_upsampling) */
    int32_t up;             /* up_down_factor; 1 = no up-sampling conv */
    int32_t reserved;
    char name[96];          /* tensor prefix of the block's WaveNetAE, used like wn_name */
    char up_name[96];       /* up-sampling conv: <up_name>/W (3, wn_cout, wn_cout * up) and <up_name>/b */
} mbexwn_wn_block_t;

/* Flat model description == the keyword arguments of MBExWN.__init__ that matter at inference
 * (custom_pulsed_generator.py:155-230) after the rate algebra of :256-267 and :459-488. */
typedef struct {
    int32_t abi_version;
    int32_t sample_rate, hop, mel_channels;
    int32_t pulse_per_frame;    /* spect_to_pulse_upsampling_factor */
    int32_t steps_per_frame;    /* WaveNet rows per mel frame */
    int32_t pulse_channels, subbands;
    float pulse_rate, f0_min, f0_max, f0_span /* float(f0_max - f0_min) */, noise_sigma, leaky_alpha;
    int32_t n_pp_ops, n_ps_ops;
    mbexwn_op_t pp_ops[MBEXWN_MAX_OPS];     /* F0 sub-net  ("PulsPar") */
    mbexwn_op_t ps_ops[MBEXWN_MAX_OPS];     /* VTF sub-net ("PS")      */
    /* WaveNetAE (custom_AE_layers.py:114-346) */
    int32_t wn_c, wn_cin, wn_cout, wn_layers, wn_k, wn_gate;
    int32_t wn_cond_k, wn_cond_conv_up, wn_cond_lin_up;
    int32_t wn_dilations[MBEXWN_MAX_LAYERS];
    char wn_name[96];           /* tensors <wn_name>/start/W, /cond_/W, /conv1D_<i>/W, /res_skip_<i>/W, /end/W (+ /b) */
    char post_name[96];         /* wn_post_net 1x1 (custom_pulsed_generator.py:490-493) */
    /* spectral envelope / STFT (custom_pulsed_generator.py:391-400, :793-836) */
    int32_t n_ceps, stft_win, fft_size, n_lifters, n_smooth;
    float filter_max_log_range; /* 0 => exp(L) without the tanh limiter */
    /* wavetable (tf_wavetable.py:182-307) */
    int32_t wt_n_period, wt_n_tables;
    float wt_nominal_f0, wt_min_transposition, wt_max_transposition, wt_grid_norm;
    int32_t cumsum_chunk;       /* 1000 (tf_wavetable.py:429) */
    /* PQMF polyphase bank (tf_preprocess.py:120-161, :208-226) */
    int32_t pqmf_q, pqmf_back;
    int32_t halo_frames;        /* guard frames between utterances */
    /* NormMelComponents (wavegen_1d.py:578-769, model key normalize_rms_from_mell); norm_enable 0 = off.  Tensors:
     * "norm/proj" (mel_channels) inv_enorm, or (mel_channels, norm_proj_cols) pinv(mel basis)^T when norm_proj_cols > 0;
     * "norm/smooth_win" (norm_smooth_win); "norm/gwin" (norm_win) */
    int32_t norm_enable, norm_iters, norm_win, norm_smooth_win, norm_proj_cols, norm_use_max_limit;
    float norm_fact;            /* fft_size * win_size / 2 (:598) */
    float norm_floor;           /* 1 / max_norm_fact, 0 = none (:691-692) */
    float norm_compress_exp;    /* normalize_compressor_exp, 0 = none (:693-694) */
    float norm_proj_scale;      /* 1 / win_norm of the pinv variant (:603, :687) */
    float norm_lin_scale, norm_lin_off, norm_mel_scale;   /* re-scaling of the normalised mel (:725-730) */
    /* spectral shaping of the excitation (custom_pulsed_generator.py:666-724): 0 = STFT-domain vocal-tract filter,
     * 1 = ps_use_stft False: the PS sub-net ends in `subbands` log gains per frame; exp, linear interpolation x hop
     *     (ps_gain_interpolator :453) and product with the sub-band rows before the PQMF (:857-884, :916-917),
     * 2 = ps_off: the PQMF output is the signal (n_ps_ops may be 0) */
    int32_t ps_mode;
    int32_t ps_preserve_energy; /* ps_mode 1: subtract the mean log gain over the bands (:867-876) */
    int32_t wt_subharm;         /* wavetable_config.add_subharm_chans: sin(2 pi phase / ii), ii = 2 .. n + 1, beside every pulse
                                   sample (tf_wavetable.py:554-559); wn_cin = pulse_channels * (1 + n) [+ 1 noise] */
    int32_t pulse_pqmf_taps;    /* pulse_channels_use_pqmf: the pulse train enters the WaveNet as the pulse_channels bands of a PQMF
                                   analysis bank (tensor "pulse_pqmf" (pulse_channels, taps + 1); TFPQMF.analysis,
                                   tf_preprocess.py:192-202, custom_pulsed_generator.py:895) instead of being folded; 0 = off */
    int32_t wn_causal;          /* force_causal: the dilated WaveNet convs and the conditioning conv pad (k - 1) d zeros on the left
                                   only (Keras padding "causal", custom_pulsed_generator.py:474-475); the sub-net ops carry their
                                   own pad_l / pad_r */
    /* pp_waveNetBlocks with more than one block or with up-sampling (pp_mod_subnet_upsampling_factors / _channel_factors,
     * custom_pulsed_generator.py:465-488).  0 = the single WaveNetAE of the wn_* fields, no up-sampling.  n >= 1: the blocks run
     * in sequence and share wn_layers / wn_k / wn_dilations / wn_gate / wn_cond_k / wn_cond_lin_up / wn_cout (one pp_mod_subnet
     * dict builds them all); wn_c, wn_cond_conv_up and wn_name are ignored.  Block 0 reads wn_cin channels at steps_per_frame
     * rows per frame, block i + 1 reads the wn_cout channels block i (and its up-sampling conv) left at up_i times that rate;
     * steps_per_frame x prod(up_i) = hop / subbands is the rate of the post net and the PQMF bank. */
    int32_t wn_n_blocks;
    mbexwn_wn_block_t wn_blocks[MBEXWN_MAX_BLOCKS];
} mbexwn_config_t;

/* One batch on the padded frame grid; all pointers are DEVICE pointers. */
typedef struct {
    int32_t n_utt;
    int32_t n_frames;               /* padded frames, guards included */
    int32_t n_chunks;               /* total cumsum chunks = chunk_first[n_utt] */
    const int32_t* frame_utt;       /* [n_frames] */
    const int32_t* utt_begin;       /* [n_utt] */
    const int32_t* utt_end;         /* [n_utt] */
    const int32_t* chunk_first;     /* [n_utt + 1] exclusive scan of ceil(T_u * pulse_per_frame / cumsum_chunk) */
    const float* mel;               /* (n_frames, mel_channels) scaled log-mel, guard frames zero */
    const float* noise;             /* (n_frames * steps_per_frame) N(0,1) draws, or NULL => in-kernel Philox(seed) */
    const float* f0_override;       /* (n_frames * pulse_per_frame) Hz, or NULL => F0 sub-net (infer_components, wavegen_1d.py:528-557) */
    const int32_t* utt_ids;         /* [n_utt] global utterance ids for the Philox stream, or NULL => batch index */
    uint64_t seed;
    float* out;                     /* (n_frames * hop) waveform on the grid */
    const float* phase_carry;       /* [n_utt] or NULL: unwrapped running sum of the wrapped 1000-sample chunk totals that
                                       precede the utterance (stable_cumsum_and_wrap, tf_wavetable.py:470-486) when the
                                       "utterance" is a window of a longer signal that starts on a chunk boundary */
} mbexwn_batch_t;

/* ---- life cycle: stands in for create_model + build_model + load_weights (mel_inverter.py:184-210) ---- */
MBEXWN_API int mbexwn_abi_version(void);
MBEXWN_API int mbexwn_create(const mbexwn_config_t* cfg, mbexwn_handle_t* out);
MBEXWN_API void mbexwn_destroy(mbexwn_handle_t h);
MBEXWN_API const char* mbexwn_last_error(mbexwn_handle_t h);

/* Register a device tensor by name (folded weights, biases, PReLU slopes, DSP constants).  The pointer must stay
 * valid for the life of the handle.  Names: see mbexwn_op_t / mbexwn_config_t, plus the constants "wavetable"
 * (n_period+1, n_tables), "pqmf_poly" (Q, S, S), "window" (win), "inv_window" (win), "twiddle" (fft/2, 2),
 * "lifters" (n_lifters, n_ceps), "lifter_grid" (n_lifters), "f0_smooth" (n_smooth).
 * Tensor-core path (precision != FP32_SIMT), per WaveNet layer i, C padded to cpad = ceil(C / 64) * 64:
 *   "<wn_name>/tc/W1_<i>" bf16 (2*cpad, 2*k*cpad): rows = output channels permuted so that each tile of 256 rows
 *        (the last tile may be narrower) is [tanh channels | the matching sigmoid channels]; columns = [hi plane |
 *        lo plane], each (tap, cin)-major
 *   "<wn_name>/tc/b1_<i>" fp32 (2*cpad) in the same row order
 *   "<wn_name>/tc/R_<i>"  bf16 (cpad + opad, or opad for the last layer; 2*cpad columns), opad = wn_cout rounded up to
 *        32: rows = [res channels | res_skip's skip half pre-multiplied by the linear `end` 1x1 (C x wn_cout)], so the
 *        kernel accumulates end(skip sum) -- the WaveNet output -- directly (custom_AE_layers.py:324-340)
 *   "<wn_name>/tc/rb_<i>" fp32 in the same row order (layer 0 carries all skip biases @ W_end + b_end) */
MBEXWN_API int mbexwn_set_tensor(mbexwn_handle_t h, const char* name, const void* dev_ptr, size_t n_bytes);

/* Host-side copy of a one-element tensor (e.g. the bias "<name>/b" of a conv with one output channel) so that fused kernels
 * can take it as a launch argument; optional -- without it the un-fused kernels run. */
MBEXWN_API int mbexwn_set_scalar(mbexwn_handle_t h, const char* name, float value);

/* ---- forward: stands in for MELInverter.synth_from_mel / PaNWaveNet.infer ---- */
MBEXWN_API size_t mbexwn_workspace_bytes(mbexwn_handle_t h, int32_t n_frames, int32_t n_chunks, int32_t precision);
MBEXWN_API int mbexwn_forward(mbexwn_handle_t h, const mbexwn_batch_t* batch, int32_t precision,
                   void* workspace, size_t workspace_bytes, void* cuda_stream);

/* Same call with HOST buffers (pinned for asynchronous copies): copies mel (and noise if given) host->device into
 * the staging pointers of `batch`, runs the forward, copies the waveform device->host and synchronises the stream.
 * mel_host: (n_frames, mel_channels); out_host: (n_frames * hop). */
MBEXWN_API int mbexwn_forward_host(mbexwn_handle_t h, const mbexwn_batch_t* batch, int32_t precision,
                        const float* mel_host, const float* noise_host, float* out_host,
                        void* workspace, size_t workspace_bytes, void* cuda_stream);

/* Pipelined form of mbexwn_forward_host for a stream of batches (throughput serving): `slot` 0 / 1 alternates between two
 * sets of caller-owned device input / output buffers (two mbexwn_batch_t with their own mel / noise / out pointers; the
 * workspace is shared).  _begin enqueues H2D (copy stream) -> forward (caller's stream) -> D2H (second copy stream) with
 * event dependencies and returns at once, so the copies of neighbouring calls run under the kernels of this one;
 * _wait blocks until the waveform of the slot's last _begin is in out_host.  Host buffers must be pinned. */
MBEXWN_API int mbexwn_forward_host_begin(mbexwn_handle_t h, int32_t slot, const mbexwn_batch_t* batch, int32_t precision,
                                         const float* mel_host, const float* noise_host, float* out_host,
                                         void* workspace, size_t workspace_bytes, void* cuda_stream);
MBEXWN_API int mbexwn_forward_host_wait(mbexwn_handle_t h, int32_t slot);

/* Per-stage taps (return_F0 / return_components of PaNWaveNet.infer, custom_pulsed_generator.py:756-771, plus the
 * stage boundaries of SURVEY.md 8a): after mbexwn_forward the named intermediate lives in the workspace at
 * [*offset_bytes, *offset_bytes + *n_bytes).  Names: "F0", "phase", "index", "pulse", "wn_in", "cond", "wn_out"
 * (rows, wn_cout rounded up to 32), "skip" (fp32 variant only), "subbands", "excitation", "ceps" (ps_mode 1: the (frames, subbands) log gains), "frames", "vtf", "lifter_index";
 * with norm_enable also "mel_norm" (frames, mel_channels), "norm_rms_a" / "norm_rms_b" (frames), "norm_gain" (frames * hop). */
MBEXWN_API int mbexwn_tap(mbexwn_handle_t h, const char* name, int32_t n_frames, int32_t n_chunks, int32_t precision,
               size_t* offset_bytes, size_t* n_bytes);

/* Number of kernel launches issued by the last mbexwn_forward on this handle. */
MBEXWN_API int mbexwn_last_launch_count(mbexwn_handle_t h);

/* Options: "debug_taps" (default 1): keep the phase / index / pulse / vtf / lifter_index taps in the workspace;
 * "stage_timing" (default 0): record CUDA events on the caller's stream at the stage boundaries of each forward;
 * "tc_cta_group" (1 or 2): tensor-core tiles owned by one CTA or by a CTA pair (cluster of 2, tcgen05 cta_group::2);
 * "tc_cond_stage" (default 1): the gate epilogue reads its conditioning rows from a shared-memory stage (0: global);
 * "fuse_tail" (default 1): the LinInterp -> 1x1 -> LinInterp end of the F0 sub-net runs as one kernel;
 * "stop_after_f0" (default 0): return after the F0 sub-net (tap "F0"), for the F0 pass of chunked long-form synthesis;
 * "tc8_h_lo" / "tc8_a_lo": log2 scale of the e4m3 lo8 planes of the residual stream / gated activations (F16F8);
 * "tc_fused" (default 1): one persistent kernel per WaveNet layer (dilated conv, gate, res/skip 1x1 and the residual update in one
 *   launch; the gated activations stay in L2) -- 0: never (a gate and a res/skip launch per layer), 1: whenever the geometry allows
 *   (CTA pairs, conditioning rows fit the shared-memory stage), 2: only when every CTA pair gets at least two 256-row tiles.  The two
 *   paths agree to ~1e-5 of peak (different K order), so the default never switches with the batch size;
 * "tc_slab" (default 1): the dilated taps of the fused kernel share one A slab per 64-channel block (row-shifted MMA operand
 *   views of one shared-memory tile); 0: one TMA tile per tap (bit-identical results, more L2 -> SM traffic);
 * "tc_cluster" (default 2): CTAs per cluster of the fused kernel.  4: two CTA pairs walk the same weight-tile sequence and
 *   share every weight (B) tile load by TMA multicast (15 % fewer L2 sectors read, bit-identical results); used when every
 *   pair has at least two 256-row tiles and the device keeps >= 128 SMs busy with clusters of 4
 *   (cudaOccupancyMaxActiveClusters: 33 clusters = 132 of 148 SMs on a B200, which costs more than the multicast saves);
 * "tc_interleave" (default 1): the res tiles of a 256-row tile run right behind the first gate tile of the next one (gate tiles cut
 *   256 + 192 + 192 instead of 224 + 224 + 192), so that they read the gated activations and the old residual rows back from the L2;
 *   0: behind the last gate tile (bit-identical results);
 * "tc_discard" (default 1): rows of the activation scratch are discarded from the L2 (discard.global.L2, no write-back) once the res
 *   tiles have read them.  Together: 2.4 instead of 3.6 GB of DRAM traffic per layer at 64 x 5 s, ~1 % faster;
 * "tc_trace" (default 0): k > 0 records per-tile cycle stamps of the fused kernel of layer k - 1 (mbexwn_tc_trace_read). */
MBEXWN_API int mbexwn_set_option(mbexwn_handle_t h, const char* name, int32_t value);

/* What the last forward did: "tc_last_fused" (1: fused layer kernel), "tc_last_cluster" (CTAs per cluster of the fused
 * kernel, 2 or 4), "tc_max_quads" (clusters of 4 the device holds at once; 0: not asked yet, -1: the query failed). */
MBEXWN_API int mbexwn_get_info(mbexwn_handle_t h, const char* name, int32_t* value);

/* Device time of each stage of the last forward (needs "stage_timing"); ms[MBEXWN_N_STAGES] in the order
 * f0_net, excitation, cond_conv, wavenet, post_pqmf, vtf_net, stft_ola.  Synchronises on the last event. */
#define MBEXWN_N_STAGES 7
MBEXWN_API int mbexwn_stage_ms(mbexwn_handle_t h, float* ms);

/* Device time of the WaveNet tap-GEMM launches of the last forward (needs "stage_timing" and a tensor-core precision):
 * sum over the layers of the gate launches (dilated conv + gate epilogue) and of the res/skip launches, from CUDA events
 * recorded on the caller's stream around every launch.  Synchronises on the last event.  A stack of several WaveNet blocks
 * (wn_n_blocks > 1) reports the launches of its last block; mbexwn_stage_ms covers all blocks in the "wavenet" stage. */
MBEXWN_API int mbexwn_wavenet_launch_ms(mbexwn_handle_t h, float* gate_ms, float* resskip_ms, int32_t* n_layers);
/* With "tc_fused" in effect a layer is ONE launch: gate_ms then holds the sum of the fused launches and resskip_ms is 0. */

/* ---- chunked long-form synthesis (BASELINE.json configs[4]; long_form.py) without the host in the loop ----
 * mbexwn_phase_carry: f0_dev = F0 of ONE whole signal at the pulse rate (n_samples, device); run_out_dev[c] (device,
 * ceil(n_samples / cumsum_chunk) floats) = the unwrapped float32 running sum of the wrapped chunk totals of chunks 0 .. c - 1,
 * i.e. mbexwn_batch_t.phase_carry of a window that starts at chunk c.  Same association order as the forward's own phase
 * kernels (PulseWaveTable.stable_cumsum_and_wrap, tf_wavetable.py:429-492).
 * mbexwn_gather_rows: for every segment s: dst[dst_row + r, :] = src[src_row + r, :], r < n_rows, rows of row_elems floats
 * (float4 copies when row_elems % 4 == 0); seg_dev = n_seg x {src_row, dst_row, n_rows} int64 on the device (n_seg <= 65535),
 * max_rows = the longest segment.  Cuts the
 * windows of a long signal out of device-resident mel / noise / F0 buffers into a batch grid and the cores back out of it. */
MBEXWN_API int mbexwn_phase_carry(mbexwn_handle_t h, const float* f0_dev, int64_t n_samples, float* run_out_dev, void* cuda_stream);
MBEXWN_API int mbexwn_gather_rows(mbexwn_handle_t h, const float* src_dev, float* dst_dev, int32_t row_elems, const int64_t* seg_dev,
                                  int32_t n_seg, int32_t max_rows, void* cuda_stream);

/* Range guard of MBEXWN_PREC_F16F8.  The main operand plane of that path is fp16 and the hi8 correction plane is e4m3(x),
 * unscaled and saturating at 448: a model whose residual stream leaves that range would silently drop to plain-fp16 accuracy
 * (no non-finite sample marks it).  The kernels that write the residual stream (start conv, res/skip epilogue) therefore OR
 * their findings into a sticky word: bit 0 = |x| > 448 seen, bit 1 = |x| > 60000 seen.  Valid once the forward's stream
 * has been synchronised (mbexwn_forward_host, mbexwn_forward_host_wait); `reset` != 0 clears it.  The Python MELInverter
 * re-runs a flagged batch on MBEXWN_PREC_BF16X3 (same accuracy class, fp32 exponent range) with a note on stderr. */
MBEXWN_API int mbexwn_range_status(mbexwn_handle_t h, int32_t* flags, int32_t reset);

/* Profiling aid ("tc_trace"): copies the cycle stamps of the traced fused-layer launch to `out` (host, n_words uint32);
 * layout [cta][role: 0 producer, 1 MMA issuer, 2 epilogue warp][384 tiles][4 words].  Returns the words written or < 0. */
MBEXWN_API int64_t mbexwn_tc_trace_read(mbexwn_handle_t h, uint32_t* out, int64_t n_words);

/* ---- single kernels on caller-provided buffers (stage-level parity tests; same kernels the forward uses) ---- */
MBEXWN_API int mbexwn_k_conv1d(mbexwn_handle_t h, const mbexwn_batch_t* grid, const mbexwn_op_t* op, int32_t rate,
                    const float* x, const float* w, const float* bias, const float* alpha, float* out,
                    void* cuda_stream);
MBEXWN_API int mbexwn_k_lininterp(mbexwn_handle_t h, const mbexwn_batch_t* grid, const mbexwn_op_t* op,
                       const float* x, const float* alpha, float* out, void* cuda_stream);

/* The tensor-core tap-GEMM on its own: out (rows, n) fp32 = sum over K blocks of
 * A[rows + shift_b, a_col_b : a_col_b + 64] @ B[0:n, b_col_b : b_col_b + 64]^T with A (rows, a_cols) and B (n, b_cols)
 * row-major bf16; kblocks = n_kb x {a_col, row_shift, b_col}.  Rows shifted outside [0, rows) read as zero. */
MBEXWN_API int mbexwn_k_tc_gemm(mbexwn_handle_t h, const void* a_bf16, int64_t rows, int32_t a_cols, const void* b_bf16,
                                int32_t n, int32_t b_cols, const int32_t* kblocks, int32_t n_kb, float* out,
                                void* cuda_stream);

/* The split-precision tap-GEMM of MBEXWN_PREC_F16F8 on its own.  A (rows, 4 * a_cpad bytes per row) = [fp16 x (a_cpad) |
 * for every 64 channels: e4m3 lo8 (64 B), e4m3 hi8 (64 B)], B (n, 4 * b_k bytes per row) = [fp16 w (b_k) | for every 64 of K:
 * e4m3 hi8 (64 B), e4m3 lo8 (64 B)] -- the two correction products of a K block are one 128-byte e4m3 block;
 * out = sum_b A16 @ B16^T + 2^-15 * sum_b (A_lo8 @ B_hi8^T + A_hi8 @ B_lo8^T) over the same K-block list as above. */
MBEXWN_API int mbexwn_k_tc_gemm_f16f8(mbexwn_handle_t h, const void* a, int64_t rows, int32_t a_cpad, const void* b,
                                      int32_t n, int32_t b_k, const int32_t* kblocks, int32_t n_kb, float* out,
                                      void* cuda_stream);

/* ---- analysis side: audio -> log-mel; stands in for MELInverter.generate_mel_from_snd (mel_inverter.py:156-182) ->
 * compute_mel_spectrogram_internal (vocoder/model/preprocess.py:417-560, band_limit=None) -> calc_stft
 * (sig_proc/spec/stft.py:14-96), and for scale_mel_spectrogram (preprocess.py:80-108) when mode != 0.
 * Stateless (no handle): every table is a caller-owned DEVICE buffer. ---- */
typedef struct {
    int32_t hop, win, fft_size, n_mel;   /* fft_size must be 2048 (the scheme configuration), win <= fft_size, n_mel <= 128 */
    int32_t mode;                /* 0: log(max(mel, floor))                                  (do_post=False, preprocess.py:543)
                                    1: log_scale * log(mel * lin_scale + lin_off)            (preprocess.py:107)
                                    2: log_scale * log(max(mel * lin_scale, lin_off))        (use_max_limit, preprocess.py:102) */
    float lin_scale, lin_off, log_scale, floor;
    const float* window;         /* (win) symmetric Hann (sig_proc/Mwindows.py:65-67,192-196) */
    const float* twiddle;        /* (fft_size / 2, 2) exp(-2 pi i k / fft_size) */
    const int32_t* mel_lo;       /* (n_mel) first FFT bin of each triangular band */
    const int32_t* mel_cnt;      /* (n_mel) number of bins of each band */
    const int32_t* mel_off;      /* (n_mel) offset of the band's weights in mel_w */
    const float* mel_w;          /* packed band weights (librosa.filters.mel, Slaney; preprocess.py:52-74) */
} mbexwn_analysis_config_t;

typedef struct {
    int32_t n_utt;
    int32_t n_pairs;             /* pair_first[n_utt]: one CTA per pair of consecutive frames */
    int32_t n_frames;            /* frame_begin[n_utt] */
    int64_t n_samples_total;     /* samples in `audio` */
    const int64_t* sample_begin; /* [n_utt] first sample of each utterance in `audio` */
    const int32_t* n_samples;    /* [n_utt] >= 1 */
    const int32_t* frame_begin;  /* [n_utt + 1] exclusive scan of n_samples / hop + 1 (stft.py:56) */
    const int32_t* pair_first;   /* [n_utt + 1] exclusive scan of ceil(frames / 2) */
    const float* audio;          /* utterances back to back */
    float* mel;                  /* (n_frames, n_mel) */
    float* mag_tap;              /* optional (n_frames, fft_size / 2 + 1) |STFT|, or NULL */
} mbexwn_analysis_batch_t;

MBEXWN_API int mbexwn_mel_analysis(const mbexwn_analysis_config_t* cfg, const mbexwn_analysis_batch_t* batch,
                                   void* cuda_stream);
/* Same call with HOST audio / mel buffers (pinned for asynchronous copies): H2D of the samples into batch->audio, the
 * kernel, D2H of batch->mel, stream synchronised. */
MBEXWN_API int mbexwn_mel_analysis_host(const mbexwn_analysis_config_t* cfg, const mbexwn_analysis_batch_t* batch,
                                        const float* audio_host, float* mel_host, void* cuda_stream);
/* Message of the last failed handle-less call on this thread (the analysis entry points). */
MBEXWN_API const char* mbexwn_global_error(void);

#ifdef __cplusplus
}
#endif
#endif /* MBEXWN_H_ */
